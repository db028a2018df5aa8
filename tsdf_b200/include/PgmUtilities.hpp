// PgmUtilities.hpp — binary 16-bit PGM reader (reference src/include/PgmUtilities.hpp).
#ifndef PGM_UTILITIES_H
#define PGM_UTILITIES_H
#include <cstdint>
#include <string>

uint16_t *read_pgm(const std::string &file_name, uint32_t &width, uint32_t &height);   // raw sample bytes, file order
#endif

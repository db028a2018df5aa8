// DepthMapUtilities.hpp — depth map readers (reference src/include/DepthMapUtilities.hpp).
#ifndef DepthMapUtilities_h
#define DepthMapUtilities_h

#include <cstdint>
#include <string>

uint16_t *load_depth_map(std::string file_name, uint16_t &width, uint16_t &height);    // declared, never defined in the reference
uint16_t *read_nyu_depth_map(const std::string &file_name, uint32_t &width, uint32_t &height);   // big-endian 16-bit PGM, mm
uint16_t *read_tum_depth_map(const std::string &file_name, uint32_t &width, uint32_t &height);   // PNG, 1/5000 m -> mm
#endif

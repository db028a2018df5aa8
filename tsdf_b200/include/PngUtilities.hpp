// PngUtilities.hpp — PNG read/write (reference src/include/PngUtilities.hpp).  16-bit samples are host-order u16 in
// memory and big-endian in the file, as in the reference (Utilities/PngUtilities.cpp:63-66).  Loaders return new[]
// arrays the caller deletes, or nullptr.
#ifndef PNG_UTILITIES_H
#define PNG_UTILITIES_H

#include <cstdint>
#include <string>

uint16_t *load_png_from_file(const std::string file_name, uint32_t &width, uint32_t &height);
uint8_t *load_colour_png_from_file(const std::string file_name, uint32_t &width, uint32_t &height);
bool save_png_to_file(const std::string file_name, uint32_t width, uint32_t height, const uint16_t *pixel_data);
bool save_png_to_file(const std::string file_name, uint32_t width, uint32_t height, const uint8_t *pixel_data);
bool save_colour_png_to_file(const std::string file_name, uint32_t width, uint32_t height, const uint8_t *pixel_data);
#endif

// C entry points around the REFERENCE's own BilateralFilter class (compiled from /root/reference/src/BilateralFilter.cpp
// by oracle/build_ref.sh into oracle/_ref/libref_bilateral.so).  TEST INFRASTRUCTURE: pins the oracle's restatement.
#include <cstdint>
#include <cstring>
#include "include/BilateralFilter.hpp"

extern "C" void ref_bilateral_u8(uint8_t *image, int width, int height, float sigma_colour, float sigma_space) {
    BilateralFilter f(sigma_colour, sigma_space);
    f.filter(image, width, height);
}

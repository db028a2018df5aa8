// BilateralFilter.hpp — the reference's bilateral filter class (src/include/BilateralFilter.hpp), filtering on the GPU
// (tsdf_b200_bilateral_host).  Same surface: constructed from the two sigmas, filter() works in place on a host image.
// The look-up tables are the constructor's own (src/BilateralFilter.cpp:15-42); for 16-bit images the range table covers
// every possible difference instead of the reference's 256 entries (indexing past them is undefined behaviour there).
#ifndef BilateralFilter_hpp
#define BilateralFilter_hpp

#include <cstdint>

class BilateralFilter {
public:
    BilateralFilter(float sigma_colour, float sigma_space);
    ~BilateralFilter();
    BilateralFilter(const BilateralFilter &) = delete;
    BilateralFilter &operator=(const BilateralFilter &) = delete;

    void filter(const uint8_t *depth_image, int width, int height) const;      // in place, like the reference
    void filter(const uint16_t *depth_image, int width, int height) const;

private:
    float m_sigma_colour;
    float m_sigma_space;
    float *m_kernel;            // m_kernel_size^2 spatial weights
    float *m_similarity;        // 65536 range weights
    int m_kernel_size;
};
#endif /* BilateralFilter_hpp */

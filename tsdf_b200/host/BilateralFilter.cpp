// BilateralFilter.cpp — reference src/BilateralFilter.cpp over tsdf_b200_bilateral_host.
#include "../include/BilateralFilter.hpp"
#include "../../include/tsdf_b200.h"

#include <cmath>
#include <cstdlib>
#include <iostream>

BilateralFilter::BilateralFilter(float sigma_colour, float sigma_space) : m_sigma_colour{sigma_colour}, m_sigma_space{sigma_space} {
    const int radius = static_cast<int>(std::ceil(sigma_space * 1.5f));
    const float inv_sigma_colour_squared = 1.0f / (sigma_colour * sigma_colour);
    const float inv_sigma_space_squared = 1.0f / (sigma_space * sigma_space);
    m_kernel_size = radius * 2 + 1;                     // always odd
    m_kernel = new float[m_kernel_size * m_kernel_size];
    int idx = 0;
    for (int x = -radius; x <= radius; x++)
        for (int y = -radius; y <= radius; y++) {
            const float dist_squared = static_cast<float>(x * x + y * y);
            m_kernel[idx++] = std::exp(-dist_squared * inv_sigma_space_squared);
        }
    m_similarity = new float[65536];
    for (int i = 0; i < 65536; i++) m_similarity[i] = std::exp(-i * inv_sigma_colour_squared);
}

BilateralFilter::~BilateralFilter() {
    delete[] m_similarity;
    delete[] m_kernel;
}

static void run(void *image, int bits, int width, int height, const float *kernel, int kernel_size, const float *similarity, int n) {
    const int rc = tsdf_b200_bilateral_host(image, bits, static_cast<uint32_t>(width), static_cast<uint32_t>(height), kernel,
                                            static_cast<uint32_t>(kernel_size), similarity, static_cast<uint32_t>(n));
    if (rc != 0) {
        std::cerr << "Bilateral filter failed" << std::endl << tsdf_b200_strerror(rc) << std::endl;
        std::exit(-1);
    }
}

void BilateralFilter::filter(const uint8_t *image, int width, int height) const {
    // 8-bit differences never exceed 255: the first 256 entries are the reference's table
    run(const_cast<uint8_t *>(image), 8, width, height, m_kernel, m_kernel_size, m_similarity, 256);
}

void BilateralFilter::filter(const uint16_t *image, int width, int height) const {
    run(const_cast<uint16_t *>(image), 16, width, height, m_kernel, m_kernel_size, m_similarity, 65536);
}

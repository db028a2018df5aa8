// DepthImage.cpp — reference src/DataLoader/DepthImage.cpp.
#include "../include/DepthImage.hpp"
#include "../include/FileUtilities.hpp"
#include "../include/PngUtilities.hpp"

#include "../../include/tsdf_b200.h"

#include <cstring>
#include <stdexcept>

// The pixels live in the library's pinned pool (tsdf_b200_host_alloc): the upload inside TSDFVolume::integrate is then an
// asynchronous DMA from the caller's buffer instead of a staged copy.
static uint16_t *pixels_alloc(size_t n) {
    uint16_t *p = static_cast<uint16_t *>(tsdf_b200_host_alloc(n * sizeof(uint16_t)));
    if (!p) throw std::bad_alloc();
    return p;
}

DepthImage::DepthImage(std::string file_name) : m_width(0), m_height(0), m_data(nullptr) {
    bool is_directory = false;
    if (!file_exists(file_name, is_directory) || is_directory) throw std::invalid_argument("File not found or is directory " + file_name);
    uint32_t w = 0, h = 0;
    uint16_t *loaded = load_png_from_file(file_name, w, h);
    if (!loaded) throw std::invalid_argument("Problem reading depth image " + file_name);
    m_width = static_cast<uint16_t>(w);
    m_height = static_cast<uint16_t>(h);
    m_data = pixels_alloc(static_cast<size_t>(w) * h);
    std::memcpy(m_data, loaded, static_cast<size_t>(w) * h * sizeof(uint16_t));
    delete[] loaded;
}

DepthImage::DepthImage(const uint16_t width, const uint16_t height, const uint16_t *const data) : m_width(width), m_height(height), m_data(nullptr) {
    if (width == 0 || height == 0 || data == nullptr) throw std::invalid_argument("width and height must be non-zero and data must not be null");
    const size_t n = static_cast<size_t>(width) * height;
    m_data = pixels_alloc(n);
    std::memcpy(m_data, data, n * sizeof(uint16_t));
}

DepthImage::~DepthImage() { tsdf_b200_host_free(m_data); }

void DepthImage::scale_depth(const float factor) {
    const size_t n = static_cast<size_t>(m_width) * m_height;
    for (size_t i = 0; i < n; i++) m_data[i] = static_cast<uint16_t>(static_cast<float>(m_data[i]) * factor);
}

void DepthImage::truncate_depth_to(const int mm) {
    const size_t n = static_cast<size_t>(m_width) * m_height;
    for (size_t i = 0; i < n; i++) if (m_data[i] > mm) m_data[i] = 0;
}

void DepthImage::min_max(uint16_t &min, uint16_t &max) {
    min = 0xFFFF;
    max = 0;
    const size_t n = static_cast<size_t>(m_width) * m_height;
    for (size_t i = 0; i < n; i++) {
        if (m_data[i] > max) max = m_data[i];
        if (m_data[i] < min) min = m_data[i];
    }
}

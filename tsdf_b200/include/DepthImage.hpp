// DepthImage.hpp — u16 depth frame (reference src/include/DepthImage.hpp:6-61).
#ifndef DEPTH_IMAGE_H
#define DEPTH_IMAGE_H

#include <cstdint>
#include <string>

class DepthImage {
public:
    DepthImage(std::string file_name);                                                // 16-bit greyscale PNG
    DepthImage(const uint16_t width, const uint16_t height, const uint16_t *const data);  // copies data
    ~DepthImage();
    DepthImage(const DepthImage &) = delete;
    DepthImage &operator=(const DepthImage &) = delete;

    void scale_depth(const float factor);          // v <- (uint16_t)((float)v * factor)
    void truncate_depth_to(const int mm);          // v > mm -> 0
    void min_max(uint16_t &min, uint16_t &max);
    uint16_t width() const { return m_width; }
    uint16_t height() const { return m_height; }
    const uint16_t *data() const { return m_data; }

private:
    uint16_t m_width;
    uint16_t m_height;
    uint16_t *m_data;
};
#endif

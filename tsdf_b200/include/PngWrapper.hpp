// PngWrapper.hpp — owning wrapper of an 8-bit grey, 16-bit grey or 8-bit RGB image that can be written as PNG
// (reference src/include/PngWrapper.hpp:6-32).  The codec is tsdf_b200/host/PngUtilities.cpp (zlib), not libpng.
#ifndef PNGWRAPPER_H
#define PNGWRAPPER_H

#include <cstdint>
#include <string>

class PngWrapper {
public:
    enum PNG_TYPE { GREYSCALE_8, GREYSCALE_16, COLOUR };

    PngWrapper(const std::string &file_name, PNG_TYPE type = GREYSCALE_16);
    PngWrapper(const uint16_t width, const uint16_t height, const uint8_t *data, PNG_TYPE);   // copies data
    virtual ~PngWrapper();

    inline uint32_t width() const { return m_width; }
    inline uint32_t height() const { return m_height; }
    bool save_to(const std::string &file_name) const;

private:
    uint32_t m_width;
    uint32_t m_height;
    const uint8_t *m_data;
    PNG_TYPE m_type;
};
#endif  // PNGWRAPPER_H

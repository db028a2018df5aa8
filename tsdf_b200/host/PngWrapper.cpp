// PngWrapper.cpp — reference src/Utilities/PngWrapper.cpp.
#include "../include/PngWrapper.hpp"
#include "../include/PngUtilities.hpp"

#include <cstring>
#include <stdexcept>

namespace {
size_t bytes_per_pixel(PngWrapper::PNG_TYPE type) { return type == PngWrapper::GREYSCALE_8 ? 1 : (type == PngWrapper::GREYSCALE_16 ? 2 : 3); }
}

PngWrapper::PngWrapper(const std::string &file_name, PNG_TYPE type) : m_width(0), m_height(0), m_data(nullptr), m_type(type) {
    // (8-bit greyscale files cannot be loaded through this class in the reference either, PngWrapper.cpp:6-20)
    if (type == COLOUR) m_data = load_colour_png_from_file(file_name, m_width, m_height);
    else if (type == GREYSCALE_16) m_data = reinterpret_cast<uint8_t *>(load_png_from_file(file_name, m_width, m_height));
    if (!m_data) throw std::invalid_argument("Failed to create PNGWrapper");
}

PngWrapper::PngWrapper(const uint16_t width, const uint16_t height, const uint8_t *data, PNG_TYPE type)
    : m_width(width), m_height(height), m_data(nullptr), m_type(type) {
    const size_t n = size_t(width) * height * bytes_per_pixel(type);
    uint8_t *copy = new uint8_t[n];
    std::memcpy(copy, data, n);
    m_data = copy;
}

PngWrapper::~PngWrapper() {
    // the 16-bit loader allocates uint16_t[]; release through the type it was allocated with
    if (m_type == GREYSCALE_16) delete[] reinterpret_cast<const uint16_t *>(m_data);
    else delete[] m_data;
    m_data = nullptr;
}

bool PngWrapper::save_to(const std::string &file_name) const {
    switch (m_type) {
        case COLOUR: return save_colour_png_to_file(file_name, m_width, m_height, m_data);
        case GREYSCALE_8: return save_png_to_file(file_name, m_width, m_height, m_data);
        case GREYSCALE_16: return save_png_to_file(file_name, m_width, m_height, reinterpret_cast<const uint16_t *>(m_data));
    }
    return false;
}

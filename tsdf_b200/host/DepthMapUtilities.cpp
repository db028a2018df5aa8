// DepthMapUtilities.cpp — reference src/Utilities/DepthMapUtilities.cpp and PgmUtilities.cpp.
#include "../include/DepthMapUtilities.hpp"
#include "../include/PgmUtilities.hpp"
#include "../include/PngUtilities.hpp"

#include <cstdio>
#include <fstream>

uint16_t *read_tum_depth_map(const std::string &file_name, uint32_t &width, uint32_t &height) {
    uint16_t *map = load_png_from_file(file_name, width, height);
    if (!map) return nullptr;
    // 5000 units per metre -> millimetres
    for (size_t i = 0; i < size_t(width) * height; i++) map[i] = map[i] / 5;
    return map;
}

uint16_t *read_nyu_depth_map(const std::string &file_name, uint32_t &width, uint32_t &height) {
    uint16_t *map = read_pgm(file_name, width, height);
    if (!map) return nullptr;
    // already millimetres, stored big-endian
    for (size_t i = 0; i < size_t(width) * height; i++) map[i] = static_cast<uint16_t>((map[i] >> 8) + ((map[i] & 0xFF) * 256));
    return map;
}

uint16_t *load_depth_map(std::string file_name, uint16_t &width, uint16_t &height) {
    uint32_t w = 0, h = 0;
    uint16_t *map = load_png_from_file(file_name, w, h);
    width = static_cast<uint16_t>(w);
    height = static_cast<uint16_t>(h);
    return map;
}

uint16_t *read_pgm(const std::string &file_name, uint32_t &width, uint32_t &height) {
    width = height = 0;
    std::ifstream f(file_name, std::ios::binary);
    std::string magic;
    uint32_t maxval = 0;
    auto next_token = [&f]() {
        std::string t;
        while (f >> t) {
            if (t[0] != '#') return t;
            std::getline(f, t);          // comment to end of line
        }
        return std::string();
    };
    magic = next_token();
    if (magic != "P5") return nullptr;
    width = static_cast<uint32_t>(std::stoul("0" + next_token()));
    height = static_cast<uint32_t>(std::stoul("0" + next_token()));
    maxval = static_cast<uint32_t>(std::stoul("0" + next_token()));
    f.get();                             // the single whitespace byte after maxval
    if (width == 0 || height == 0 || maxval < 256) return nullptr;
    uint16_t *data = new uint16_t[size_t(width) * height];
    f.read(reinterpret_cast<char *>(data), std::streamsize(size_t(width) * height * 2));
    if (!f) { delete[] data; return nullptr; }
    return data;
}

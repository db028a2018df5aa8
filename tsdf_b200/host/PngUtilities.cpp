// PngUtilities.cpp — minimal PNG codec on zlib for the three formats the TSDF tools use: 8-bit grey, 16-bit grey,
// 8-bit RGB (the reference uses libpng for the same five entry points, src/Utilities/PngUtilities.cpp:13-355).
// Reader: non-interlaced images of exactly those formats, all five scanline filters.  Writer: filter 0, one IDAT.
#include "../include/PngUtilities.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>
#include <zlib.h>

namespace {
const unsigned char kSignature[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

uint32_t be32(const unsigned char *p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
void put_be32(unsigned char *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

struct Image { uint32_t width = 0, height = 0; int bit_depth = 0, colour_type = 0; std::vector<unsigned char> bytes; };

bool read_file(const std::string &name, std::vector<unsigned char> &out) {
    FILE *f = std::fopen(name.c_str(), "rb");
    if (!f) { std::cerr << "Couldn't open file " << name << std::endl; return false; }
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? n : 0);
    const bool ok = n > 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Decodes into raw sample bytes, rows packed, file byte order.
bool decode(const std::string &name, Image &img) {
    std::vector<unsigned char> file;
    if (!read_file(name, file)) return false;
    if (file.size() < 8 || std::memcmp(file.data(), kSignature, 8) != 0) { std::cerr << "File " << name << " is not a PNG" << std::endl; return false; }
    std::vector<unsigned char> compressed;
    int interlace = 0;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t len = be32(&file[pos]);
        if (pos + 12 + len > file.size()) return false;
        const unsigned char *type = &file[pos + 4], *data = &file[pos + 8];
        if (std::memcmp(type, "IHDR", 4) == 0 && len >= 13) {
            img.width = be32(data); img.height = be32(data + 4);
            img.bit_depth = data[8]; img.colour_type = data[9]; interlace = data[12];
        } else if (std::memcmp(type, "IDAT", 4) == 0) {
            compressed.insert(compressed.end(), data, data + len);
        } else if (std::memcmp(type, "IEND", 4) == 0) {
            break;
        }
        pos += 12 + len;
    }
    if (img.width == 0 || img.height == 0 || interlace != 0) return false;
    const int channels = img.colour_type == 0 ? 1 : (img.colour_type == 2 ? 3 : 0);
    if (channels == 0 || (img.bit_depth != 8 && img.bit_depth != 16)) return false;
    const size_t bpp = size_t(channels) * img.bit_depth / 8, row = bpp * img.width;
    std::vector<unsigned char> raw((row + 1) * img.height);
    uLongf raw_len = raw.size();
    if (uncompress(raw.data(), &raw_len, compressed.data(), compressed.size()) != Z_OK || raw_len != raw.size()) return false;
    img.bytes.assign(row * img.height, 0);
    for (uint32_t y = 0; y < img.height; y++) {
        const unsigned char filter = raw[(row + 1) * y], *in = &raw[(row + 1) * y + 1];
        unsigned char *out = &img.bytes[row * y];
        const unsigned char *up = y ? out - row : nullptr;
        for (size_t i = 0; i < row; i++) {
            const int a = i >= bpp ? out[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
            int v = in[i];
            switch (filter) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: return false;
            }
            out[i] = static_cast<unsigned char>(v);
        }
    }
    return true;
}

void append_chunk(std::vector<unsigned char> &out, const char *type, const unsigned char *data, size_t len) {
    unsigned char head[8];
    put_be32(head, static_cast<uint32_t>(len));
    std::memcpy(head + 4, type, 4);
    out.insert(out.end(), head, head + 8);
    if (len) out.insert(out.end(), data, data + len);
    uLong crc = crc32(0L, reinterpret_cast<const Bytef *>(type), 4);
    if (len) crc = crc32(crc, data, static_cast<uInt>(len));
    unsigned char tail[4];
    put_be32(tail, static_cast<uint32_t>(crc));
    out.insert(out.end(), tail, tail + 4);
}

// rows: packed sample bytes in file byte order
bool encode(const std::string &name, uint32_t width, uint32_t height, int bit_depth, int colour_type, const unsigned char *rows, size_t row_bytes) {
    std::vector<unsigned char> raw((row_bytes + 1) * height);
    for (uint32_t y = 0; y < height; y++) {
        raw[(row_bytes + 1) * y] = 0;
        std::memcpy(&raw[(row_bytes + 1) * y + 1], rows + row_bytes * y, row_bytes);
    }
    uLongf clen = compressBound(raw.size());
    std::vector<unsigned char> compressed(clen);
    if (compress2(compressed.data(), &clen, raw.data(), raw.size(), Z_BEST_SPEED) != Z_OK) return false;
    std::vector<unsigned char> out(kSignature, kSignature + 8);
    unsigned char ihdr[13];
    put_be32(ihdr, width); put_be32(ihdr + 4, height);
    ihdr[8] = static_cast<unsigned char>(bit_depth); ihdr[9] = static_cast<unsigned char>(colour_type);
    ihdr[10] = ihdr[11] = ihdr[12] = 0;
    append_chunk(out, "IHDR", ihdr, 13);
    append_chunk(out, "IDAT", compressed.data(), clen);
    append_chunk(out, "IEND", nullptr, 0);
    FILE *f = std::fopen(name.c_str(), "wb");
    if (!f) { std::cerr << "Couldn't open file " << name << " for writing" << std::endl; return false; }
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    return std::fclose(f) == 0 && ok;
}
}  // namespace

uint16_t *load_png_from_file(const std::string file_name, uint32_t &width, uint32_t &height) {
    width = height = 0;
    Image img;
    if (!decode(file_name, img)) return nullptr;
    width = img.width; height = img.height;
    if (img.bit_depth != 16 || img.colour_type != 0) { std::cerr << "Expected 16bpp greyscale file" << std::endl; return nullptr; }
    uint16_t *pixels = new uint16_t[size_t(width) * height];
    for (size_t i = 0; i < size_t(width) * height; i++) pixels[i] = static_cast<uint16_t>(img.bytes[2 * i] * 256 + img.bytes[2 * i + 1]);
    return pixels;
}

uint8_t *load_colour_png_from_file(const std::string file_name, uint32_t &width, uint32_t &height) {
    width = height = 0;
    Image img;
    if (!decode(file_name, img)) return nullptr;
    width = img.width; height = img.height;
    if (img.bit_depth != 8 || img.colour_type != 2) { std::cerr << "Expected 24bpp RGB file" << std::endl; return nullptr; }
    uint8_t *pixels = new uint8_t[img.bytes.size()];
    std::memcpy(pixels, img.bytes.data(), img.bytes.size());
    return pixels;
}

bool save_png_to_file(const std::string file_name, uint32_t width, uint32_t height, const uint16_t *pixel_data) {
    std::vector<unsigned char> rows(size_t(width) * height * 2);
    for (size_t i = 0; i < size_t(width) * height; i++) { rows[2 * i] = pixel_data[i] >> 8; rows[2 * i + 1] = pixel_data[i] & 0xff; }
    return encode(file_name, width, height, 16, 0, rows.data(), size_t(width) * 2);
}

bool save_png_to_file(const std::string file_name, uint32_t width, uint32_t height, const uint8_t *pixel_data) {
    return encode(file_name, width, height, 8, 0, pixel_data, width);
}

bool save_colour_png_to_file(const std::string file_name, uint32_t width, uint32_t height, const uint8_t *pixel_data) {
    return encode(file_name, width, height, 8, 2, pixel_data, size_t(width) * 3);
}

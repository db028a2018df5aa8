// RenderUtilities.cpp — reference src/Utilities/RenderUtilities.cpp:39-112.
#include "../include/RenderUtilities.hpp"
#include "../include/Camera.hpp"

#include <cmath>
#include <vector>

void save_normals_as_colour_png(std::string filename, uint16_t width, uint16_t height, const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) {
    PngWrapper *p = normals_as_png(width, height, normals);
    p->save_to(filename);
    delete p;
}

void save_rendered_scene_as_png(std::string filename, uint16_t width, uint16_t height, const Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                                const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals, const Camera &camera, const Eigen::Vector3f &light_source) {
    PngWrapper *p = scene_as_png(width, height, vertices, normals, camera, light_source);
    p->save_to(filename);
    delete p;
}

// Lambertian shading: 0.2 ambient + 0.8 * max(0, n . unit(light - vertex)), as 8-bit grey.
PngWrapper *scene_as_png(uint16_t width, uint16_t height, const Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                         const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals, const Camera &, const Eigen::Vector3f &light_source) {
    const size_t pixels = size_t(width) * height;
    std::vector<uint8_t> image(pixels);
    const float ambient = 0.2f, diffuse = 1.0f - ambient;
    for (size_t i = 0; i < pixels; i++) {
        const Eigen::Vector3f v{vertices(0, i), vertices(1, i), vertices(2, i)}, n{normals(0, i), normals(1, i), normals(2, i)};
        const Eigen::Vector3f to_light = (light_source - v).normalized();
        const float shade = ambient + diffuse * static_cast<float>(std::fmax(0.0, n.dot(to_light)));     // fmax drops NaN (no surface) -> ambient
        image[i] = static_cast<uint8_t>(std::floor(shade * 255));
    }
    return new PngWrapper(width, height, image.data(), PngWrapper::GREYSCALE_8);
}

// (nx, ny, |nz|) mapped from [-1,1] to [0,255] as RGB.
PngWrapper *normals_as_png(uint16_t width, uint16_t height, const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) {
    const size_t pixels = size_t(width) * height;
    std::vector<uint8_t> image(pixels * 3);
    for (size_t i = 0; i < pixels; i++) {
        float n[3] = {normals(0, i), normals(1, i), normals(2, i)};
        if (n[2] < 0) n[2] = -n[2];
        for (int c = 0; c < 3; c++) {
            const float v = std::floor(((n[c] / 2.0f) + 0.5f) * 255);
            image[3 * i + c] = (v == v) ? static_cast<uint8_t>(v) : 0;           // NaN normals (no surface) -> 0
        }
    }
    return new PngWrapper(width, height, image.data(), PngWrapper::COLOUR);
}

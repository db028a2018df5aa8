// RenderUtilities.hpp — Lambertian and normal-map renderings of a raycast (reference src/include/RenderUtilities.hpp).
#ifndef RenderUtilities_h
#define RenderUtilities_h

#include "PngWrapper.hpp"
#include <Eigen/Dense>
#include <string>

class Camera;

PngWrapper *normals_as_png(uint16_t width, uint16_t height, const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals);
PngWrapper *scene_as_png(uint16_t width, uint16_t height, const Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                         const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals, const Camera &camera,
                         const Eigen::Vector3f &light_source);
void save_normals_as_colour_png(std::string filename, uint16_t width, uint16_t height,
                                const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals);
void save_rendered_scene_as_png(std::string filename, uint16_t width, uint16_t height,
                                const Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                                const Eigen::Matrix<float, 3, Eigen::Dynamic> &normals, const Camera &camera,
                                const Eigen::Vector3f &light_source);
#endif  // RenderUtilities_h

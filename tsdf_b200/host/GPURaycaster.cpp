// GPURaycaster.cpp — reference src/RayCaster/GPURaycaster.cu:519-606 over the C-ABI.
#include "../include/GPURaycaster.hpp"
#include "../../include/tsdf_b200.h"

#include <cmath>
#include <cstdlib>
#include <iostream>
#include <vector>

void GPURaycaster::raycast(const TSDFVolume &volume, const Camera &camera, Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                           Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) const {
    const size_t pixels = static_cast<size_t>(m_width) * m_height;
    vertices.resize(3, pixels);
    normals.resize(3, pixels);
    const Eigen::Matrix4f pose = camera.pose();
    const Eigen::Matrix3f kinv = camera.kinv();
    // column-major 3 x N == packed xyz per pixel: the layout tsdf_b200_volume_raycast writes
    const int rc = tsdf_b200_volume_raycast(volume.c_abi(), m_width, m_height, pose.data(), kinv.data(), vertices.data(), normals.data());
    if (rc != 0) {
        std::cerr << "Raycast failed" << std::endl << tsdf_b200_strerror(rc) << std::endl;
        std::exit(-1);
    }
}

DepthImage *GPURaycaster::render_to_depth_image(const TSDFVolume &volume, const Camera &camera) const {
    Eigen::Matrix<float, 3, Eigen::Dynamic> vertices, normals;
    raycast(volume, camera, vertices, normals);
    const size_t pixels = static_cast<size_t>(m_width) * m_height;
    std::vector<uint16_t> depth(pixels);
    for (size_t i = 0; i < pixels; i++) {
        const Eigen::Vector3f cam = camera.world_to_camera(Eigen::Vector3f{vertices(0, i), vertices(1, i), vertices(2, i)});
        // (uint16_t)roundf(z) as the reference writes it (GPURaycaster.cu:579); NaN (no surface) is made an explicit 0
        const float z = std::round(cam.z());
        depth[i] = (z == z && z > 0.0f && z < 65536.0f) ? static_cast<uint16_t>(z) : 0;
    }
    return new DepthImage(m_width, m_height, depth.data());
}

// GPURaycaster.cpp — reference src/RayCaster/GPURaycaster.cu:519-606 over the C-ABI.
#include "../include/GPURaycaster.hpp"
#include "../../include/tsdf_b200.h"

#include <cmath>
#include <cstdint>
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
        // (uint16_t)roundf(z) as the reference's host compiler evaluates it (GPURaycaster.cu:579): a truncating float ->
        // int32 conversion (cvttss2si: NaN and out-of-range give INT32_MIN) whose low 16 bits are kept.  Written out so that
        // the result does not depend on how this compiler treats the (formally undefined) out-of-range cast: a pixel
        // without a surface (NaN vertex) is 0, a vertex behind the camera wraps like it does in the reference.
        const float z = std::round(cam.z());
        const int32_t zi = (z >= -2147483648.0f && z < 2147483648.0f) ? static_cast<int32_t>(z) : INT32_MIN;
        depth[i] = static_cast<uint16_t>(static_cast<uint32_t>(zi) & 0xffffu);
    }
    return new DepthImage(m_width, m_height, depth.data());
}

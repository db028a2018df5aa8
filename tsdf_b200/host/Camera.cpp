// Camera.cpp — pin-hole camera maths on the host (behaviour of reference src/Camera.cpp:20-390).
#include "../include/Camera.hpp"
#include "../include/Definitions.hpp"

#include <cmath>

using Eigen::Matrix3f;
using Eigen::Matrix4f;
using Eigen::Vector2f;
using Eigen::Vector2i;
using Eigen::Vector3f;
using Eigen::Vector4f;

namespace {
const float kFacingEps = 1e-6f;      // "looking straight up or down" threshold of look_at (Camera.cpp:15,157)

Vector3f dehomogenise(const Vector4f &h) { return Vector3f{h[0] / h[3], h[1] / h[3], h[2] / h[3]}; }
Vector4f homogeneous(const Vector3f &p) { return Vector4f{p.x(), p.y(), p.z(), 1.0f}; }
}  // namespace

void Camera::finish_construction() {
    m_k_inverse = m_k.inverse();
    set_pose(Matrix4f::Identity());
}

void Camera::pose_changed() { m_pose_inverse = m_pose.inverse(); }

Camera::Camera(const float focal_x, const float focal_y, const float centre_x, const float centre_y) {
    m_k = Matrix3f::Zero();
    m_k(0, 0) = focal_x;
    m_k(1, 1) = focal_y;
    m_k(0, 2) = centre_x;
    m_k(1, 2) = centre_y;
    m_k(2, 2) = 1.0f;
    finish_construction();
}

Camera::Camera(const Matrix3f &k) : m_k(k) { finish_construction(); }

Camera::Camera(const int image_width, const int image_height, const float fov_x, const float fov_y) {
    // focal = size / (2 tan(fov/2)); the reference negates twice (Camera.cpp:65-68), the net sign is positive
    const float fx = image_width / (2 * std::tan(fov_x / 2.0f));
    const float fy = image_height / (2 * std::tan(fov_y / 2.0f));
    m_k << fx, 0.0f, (image_width / 2.0f), 0.0f, fy, (image_height / 2.0f), 0.0f, 0.0f, 1.0f;
    finish_construction();
}

void Camera::set_pose(const Matrix4f &pose) {
    m_pose = pose;
    pose_changed();
}

void Camera::set_pose(float vars[7]) {
    // unit quaternion (x, y, z, w) = vars[3..6] -> rotation, vars[0..2] -> translation
    const float x = vars[3], y = vars[4], z = vars[5], w = vars[6];
    Matrix4f p = Matrix4f::Identity();
    p(0, 0) = 1 - 2 * (y * y + z * z); p(0, 1) = 2 * (x * y - w * z);     p(0, 2) = 2 * (x * z + w * y);
    p(1, 0) = 2 * (x * y + w * z);     p(1, 1) = 1 - 2 * (x * x + z * z); p(1, 2) = 2 * (y * z - w * x);
    p(2, 0) = 2 * (x * z - w * y);     p(2, 1) = 2 * (y * z + w * x);     p(2, 2) = 1 - 2 * (x * x + y * y);
    p(0, 3) = vars[0]; p(1, 3) = vars[1]; p(2, 3) = vars[2];
    set_pose(p);
}

void Camera::move_to(const Vector3f &world_coordinate) { move_to(world_coordinate.x(), world_coordinate.y(), world_coordinate.z()); }

void Camera::move_to(float wx, float wy, float wz) {
    // keeps the facing; does not keep looking at an earlier look_at point
    m_pose(0, 3) = wx;
    m_pose(1, 3) = wy;
    m_pose(2, 3) = wz;
    pose_changed();
}

void Camera::look_at(const Vector3f &world_coordinate) {
    // gluLookAt with +Y up; pose columns become (left, up, forward)
    const Vector3f here = position();
    Vector3f forward = world_coordinate - here;
    forward.normalize();

    Vector3f up;
    const bool vertical = std::fabs(forward.x()) < kFacingEps && std::fabs(forward.z()) < kFacingEps;
    if (!vertical) up << 0.0f, 1.0f, 0.0f;
    else if (forward.y() < 0) up << 0.0f, 0.0f, 1.0f;        // straight down: up is +z
    else if (forward.y() > 0) up << 0.0f, 0.0f, -1.0f;       // straight up: up is -z
    // (forward.y() == 0 with a vertical facing means forward is the zero/NaN vector: up stays zero, as in the reference)

    Vector3f left = up.cross(forward);
    left.normalize();
    up = forward.cross(left);
    up.normalize();

    for (int r = 0; r < 3; r++) {
        m_pose(r, 0) = left[r];
        m_pose(r, 1) = up[r];
        m_pose(r, 2) = forward[r];
    }
    m_pose(3, 0) = m_pose(3, 1) = m_pose(3, 2) = 0.0f;
    m_pose(3, 3) = 1.0f;
    pose_changed();
}

void Camera::look_at(float wx, float wy, float wz) { look_at(Vector3f{wx, wy, wz}); }

Vector3f Camera::position() const { return Vector3f{m_pose(0, 3), m_pose(1, 3), m_pose(2, 3)}; }

Vector2f Camera::pixel_to_image_plane(const Vector2i &image_coordinate) const {
    return pixel_to_image_plane(static_cast<uint16_t>(image_coordinate.x()), static_cast<uint16_t>(image_coordinate.y()));
}

Vector2f Camera::pixel_to_image_plane(const uint16_t x, const uint16_t y) const {
    const Vector3f h = m_k_inverse * Vector3f{static_cast<float>(x), static_cast<float>(y), 1.0f};
    return Vector2f{h[0] / h[2], h[1] / h[2]};
}

Vector2i Camera::image_plane_to_pixel(const Vector2f &camera_coordinate) const {
    const Vector3f h = m_k * Vector3f{camera_coordinate.x(), camera_coordinate.y(), 1.0f};
    Vector2i pixel;
    pixel.x() = static_cast<int>(std::round(h.x()));
    pixel.y() = static_cast<int>(std::round(h.y()));
    return pixel;
}

Vector3f Camera::camera_to_world(const Vector3f &camera_coordinate) const { return dehomogenise(m_pose * homogeneous(camera_coordinate)); }

Vector3f Camera::world_to_camera_normal(const Vector3f &world_normal) const {
    Matrix3f r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = m_pose_inverse(i, j);
    return r * world_normal;
}

Vector3f Camera::world_to_camera(const Vector3f &world_coordinate) const { return dehomogenise(m_pose_inverse * homogeneous(world_coordinate)); }

Vector2i Camera::world_to_pixel(const Vector3f &world_coordinate) const {
    Vector3f img = m_k * world_to_camera(world_coordinate);
    img = img / img[2];
    Vector2i pixel;
    pixel.x() = static_cast<int>(std::round(img[0]));
    pixel.y() = static_cast<int>(std::round(img[1]));
    return pixel;
}

void Camera::depth_image_to_vertices_and_normals(const uint16_t *depth_image, const uint32_t width, const uint32_t height,
                                                 Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                                                 Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) const {
    vertices.resize(3, width * height);
    normals.resize(3, width * height);
    // Bottom-right to top-left, so the right and lower neighbours of a pixel are already back-projected when its
    // normal (right - v) x (below - v) is formed (Camera.cpp:336-390).  Pixels without depth get BAD_VERTEX and a zero
    // normal; so do pixels in the last row/column or next to a BAD_VERTEX.
    for (int64_t idx = static_cast<int64_t>(width) * height - 1; idx >= 0; idx--) {
        const uint32_t x = static_cast<uint32_t>(idx % width), y = static_cast<uint32_t>(idx / width);
        Vector3f vertex = BAD_VERTEX, normal{0.0f, 0.0f, 0.0f};
        const uint16_t depth = depth_image[idx];
        if (depth != 0) {
            const Vector2f plane = pixel_to_image_plane(static_cast<uint16_t>(x), static_cast<uint16_t>(y));
            vertex = Vector3f{plane.x(), plane.y(), 1.0f} * depth;
            if (y + 1 < height && x + 1 < width) {
                Vector3f right{vertices(0, idx + 1), vertices(1, idx + 1), vertices(2, idx + 1)};
                Vector3f below{vertices(0, idx + width), vertices(1, idx + width), vertices(2, idx + width)};
                if (right != BAD_VERTEX && below != BAD_VERTEX) normal = (right - vertex).cross(below - vertex).normalized();
            }
        }
        for (int i = 0; i < 3; i++) {
            vertices(i, idx) = vertex[i];
            normals(i, idx) = normal[i];
        }
    }
}

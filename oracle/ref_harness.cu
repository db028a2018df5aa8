// ref_harness.cu — C entry points around the REFERENCE's own classes (TSDFVolume, Camera, GPURaycaster as
// compiled from /root/reference/src by oracle/build_ref.sh), so tests and bench.py can drive the reference
// CUDA path through ctypes.  TEST INFRASTRUCTURE ONLY; never linked into the product.
#include "TSDFVolume.hpp"
#include "Camera.hpp"
#include "GPURaycaster.hpp"
#include "PngUtilities.hpp"
#include "DepthImage.hpp"
#include <cuda_runtime.h>

// libpng is not available here; GPURaycaster.cu references these only from render_to_depth_image (unused).
bool save_png_to_file(const std::string, uint32_t, uint32_t, const uint16_t *) { return false; }
DepthImage::DepthImage(uint16_t width, uint16_t height, const uint16_t *data) {
    m_width = width; m_height = height;
    m_data = new uint16_t[(size_t)width * height];
    memcpy(m_data, data, (size_t)width * height * sizeof(uint16_t));
}

static Camera make_camera(const float *k9, const float *pose16) {
    Eigen::Matrix3f k;
    memcpy(k.data(), k9, 9 * sizeof(float));
    Camera cam(k);
    Eigen::Matrix4f pose;
    memcpy(pose.data(), pose16, 16 * sizeof(float));
    cam.set_pose(pose);
    return cam;
}

extern "C" {

void *ref_volume_create(uint32_t nx, uint32_t ny, uint32_t nz, float px, float py, float pz) {
    try { return new TSDFVolume(TSDFVolume::UInt3{nx, ny, nz}, TSDFVolume::Float3{px, py, pz}); }
    catch (...) { return nullptr; }
}
void ref_volume_destroy(void *v) { delete (TSDFVolume *)v; }
void ref_volume_offset(void *v, float ox, float oy, float oz) { ((TSDFVolume *)v)->offset(ox, oy, oz); }
void ref_volume_clear(void *v) { ((TSDFVolume *)v)->clear(); }
float ref_volume_trunc(void *v) { return ((TSDFVolume *)v)->truncation_distance(); }
void ref_volume_voxel(void *v, float out[3]) {
    TSDFVolume::Float3 s = ((TSDFVolume *)v)->voxel_size();
    out[0] = s.x; out[1] = s.y; out[2] = s.z;
}

// What the reference Camera derives from (K, pose): K^-1 and pose^-1 (column-major), as fed to the kernels.
void ref_camera_matrices(const float *k9, const float *pose16, float *kinv9, float *inv_pose16) {
    Camera cam = make_camera(k9, pose16);
    Eigen::Matrix3f kinv = cam.kinv();
    memcpy(kinv9, kinv.data(), 9 * sizeof(float));
    memcpy(inv_pose16, cam.inverse_pose().data(), 16 * sizeof(float));
}

// TSDFVolume::integrate exactly as kinfu.cpp calls it (host depth map, Camera object).
void ref_volume_integrate(void *v, const uint16_t *depth, uint32_t w, uint32_t h, const float *k9, const float *pose16) {
    Camera cam = make_camera(k9, pose16);
    ((TSDFVolume *)v)->integrate(depth, w, h, cam);
}

// TSDFVolume::raycast -> GPURaycaster::raycast; outputs are 3*w*h floats each.
void ref_volume_raycast(void *v, uint32_t w, uint32_t h, const float *k9, const float *pose16, float *vertices, float *normals) {
    Camera cam = make_camera(k9, pose16);
    Eigen::Matrix<float, 3, Eigen::Dynamic> V, N;
    ((TSDFVolume *)v)->raycast((uint16_t)w, (uint16_t)h, cam, V, N);
    memcpy(vertices, V.data(), (size_t)3 * w * h * sizeof(float));
    memcpy(normals, N.data(), (size_t)3 * w * h * sizeof(float));
}

int ref_volume_read(void *v, float *dist, float *weight) {
    TSDFVolume *vol = (TSDFVolume *)v;
    TSDFVolume::UInt3 s = vol->size();
    size_t n = (size_t)s.x * s.y * s.z;
    cudaError_t e = cudaMemcpy(dist, vol->distance_data(), n * sizeof(float), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(weight, vol->weight_data(), n * sizeof(float), cudaMemcpyDeviceToHost);
    return (int)e;
}
void ref_volume_set_distance_data(void *v, const float *dist) { ((TSDFVolume *)v)->set_distance_data(dist); }
int ref_volume_read_deformation(void *v, float *nodes) {
    TSDFVolume *vol = (TSDFVolume *)v;
    TSDFVolume::UInt3 s = vol->size();
    size_t n = (size_t)s.x * s.y * s.z;
    return (int)cudaMemcpy(nodes, vol->deformation(), n * 6 * sizeof(float), cudaMemcpyDeviceToHost);
}
int ref_volume_save(void *v, const char *path) { return ((TSDFVolume *)v)->save_to_file(path) ? 0 : 1; }
void *ref_volume_load(const char *path) {
    try { return new TSDFVolume(std::string(path)); } catch (...) { return nullptr; }
}

}  // extern "C"

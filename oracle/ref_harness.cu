// ref_harness.cu — C entry points around the REFERENCE's own classes (TSDFVolume, Camera, GPURaycaster as
// compiled from /root/reference/src by oracle/build_ref.sh), so tests and bench.py can drive the reference
// CUDA path through ctypes.  TEST INFRASTRUCTURE ONLY; never linked into the product.
#include "TSDFVolume.hpp"
#include "Camera.hpp"
#include "GPURaycaster.hpp"
#include "PngUtilities.hpp"
#include "DepthImage.hpp"
#include "MarkAndSweepMC.hpp"
#include <cuda_runtime.h>
#include <vector>

// libpng is not available here.  GPURaycaster::render_to_depth_image (GPURaycaster.cu:555-606) hands its depth map to
// save_png_to_file (with a hard-wired 640x480, :592) and to the DepthImage constructor: the constructor and accessors of
// DataLoader/DepthImage.cpp (which needs libpng for its other constructor) are restated here, the PNG writer is a no-op.
bool save_png_to_file(const std::string, uint32_t, uint32_t, const uint16_t *) { return false; }
DepthImage::DepthImage(uint16_t width, uint16_t height, const uint16_t *data) {
    m_width = width; m_height = height;
    m_data = new uint16_t[(size_t)width * height];
    memcpy(m_data, data, (size_t)width * height * sizeof(uint16_t));
}
DepthImage::~DepthImage() { delete[] m_data; }
uint16_t DepthImage::width() const { return m_width; }
uint16_t DepthImage::height() const { return m_height; }
const uint16_t *DepthImage::data() const { return m_data; }

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
// raw device pointers (distance_data() / weight_data()): full-size volumes are compared on the device
const void *ref_volume_distance_ptr(void *v) { return ((TSDFVolume *)v)->distance_data(); }
const void *ref_volume_weight_ptr(void *v) { return ((TSDFVolume *)v)->weight_data(); }
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


// GPURaycaster::render_to_depth_image (GPURaycaster.cu:555-606); out holds w*h uint16.
int ref_render_depth(void *v, uint32_t w, uint32_t h, const float *k9, const float *pose16, uint16_t *out) {
    Camera cam = make_camera(k9, pose16);
    GPURaycaster caster((int)w, (int)h);
    DepthImage *d = caster.render_to_depth_image(*(TSDFVolume *)v, cam);
    if (!d) return 1;
    memcpy(out, d->data(), (size_t)w * h * sizeof(uint16_t));
    delete d;
    return 0;
}

// extract_surface (MarkAndSweepMC.cu:506-555): returns the vertex count; *out receives a malloc'ed array of 3 floats
// per vertex (free with ref_free).  NOTE: the reference exit()s when the surface is empty (:426-429).
long long ref_extract_surface(void *v, float **out) {
    std::vector<float3> vertices;
    std::vector<int3> triangles;
    extract_surface((TSDFVolume *)v, vertices, triangles);
    // triangles are (i, i+2, i+1) for every third vertex (:546-551): checked here so that the test need not carry them
    for (size_t t = 0; t < triangles.size(); t++)
        if (triangles[t].x != (int)(3 * t) || triangles[t].y != (int)(3 * t + 2) || triangles[t].z != (int)(3 * t + 1)) return -1;
    *out = (float *)malloc(vertices.size() * 3 * sizeof(float) + 4);
    for (size_t i = 0; i < vertices.size(); i++) { (*out)[3 * i] = vertices[i].x; (*out)[3 * i + 1] = vertices[i].y; (*out)[3 * i + 2] = vertices[i].z; }
    return (long long)vertices.size();
}
void ref_free(void *p) { free(p); }
int ref_volume_set_weight_data(void *v, const float *w) { ((TSDFVolume *)v)->set_weight_data(w); return 0; }

}  // extern "C"

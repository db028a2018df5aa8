// TSDFVolume.cpp — the reference's TSDFVolume methods (src/TSDF/TSDFVolume.cu:396-1058) as calls into the tsdf_b200
// C-ABI.  No CUDA in this file: device memory, streams and kernels live behind include/tsdf_b200.h.
#include "../include/TSDFVolume.hpp"
#include "../include/GPURaycaster.hpp"
#include "../../include/tsdf_b200.h"

#include <cassert>
#include <cstdlib>
#include <stdexcept>

namespace {
// check_cuda_error of the reference (Utilities/cuda_utilities.cu:5-11): message, reason, exit(-1).
void die_on(int code, const char *what) {
    if (code == 0) return;
    std::cerr << what << std::endl;
    std::cerr << tsdf_b200_strerror(code) << std::endl;
    std::exit(-1);
}
}  // namespace

void TSDFVolume::refresh() {
    uint32_t size[3];
    float physical[3], voxel[3], off[3], trunc = 0, max_weight = 0;
    tsdf_b200_volume_get(m_impl, size, physical, voxel, off, &trunc, &max_weight);
    m_size = dim3{size[0], size[1], size[2]};
    m_physical_size = float3{physical[0], physical[1], physical[2]};
    m_voxel_size = float3{voxel[0], voxel[1], voxel[2]};
    m_offset = float3{off[0], off[1], off[2]};
    m_truncation_distance = trunc;
    m_max_weight = max_weight;
    float gt[3] = {0, 0, 0}, gr[3] = {0, 0, 0};
    tsdf_b200_volume_get_global(m_impl, gt, gr);
    m_global_translation = float3{gt[0], gt[1], gt[2]};
    m_global_rotation = float3{gr[0], gr[1], gr[2]};
}

void TSDFVolume::release() {
    if (m_impl) tsdf_b200_volume_destroy(m_impl);
    m_impl = nullptr;
}

TSDFVolume::TSDFVolume(const UInt3 &size, const Float3 &physical_size)
    : m_impl(nullptr), m_offset{0.0f, 0.0f, 0.0f}, m_global_translation{0, 0, 0}, m_global_rotation{0, 0, 0} {
    if (!((size.x > 0) && (size.y > 0) && (size.z > 0) && (physical_size.x > 0) && (physical_size.y > 0) && (physical_size.z > 0)))
        throw std::invalid_argument("Attempt to construct TSDFVolume with zero or negative size");
    set_size(static_cast<uint16_t>(size.x), static_cast<uint16_t>(size.y), static_cast<uint16_t>(size.z), physical_size.x,
             physical_size.y, physical_size.z);
}

TSDFVolume::TSDFVolume(uint16_t volume_x, uint16_t volume_y, uint16_t volume_z, float psize_x, float psize_y, float psize_z)
    : m_impl(nullptr), m_offset{0.0f, 0.0f, 0.0f}, m_global_translation{0, 0, 0}, m_global_rotation{0, 0, 0} {
    if (!((volume_x > 0) && (volume_y > 0) && (volume_z > 0) && (psize_x > 0) && (psize_y > 0) && (psize_z > 0)))
        throw std::invalid_argument("Attempt to construct TSDFVolume with zero or negative size");
    set_size(volume_x, volume_y, volume_z, psize_x, psize_y, psize_z);
}

TSDFVolume::TSDFVolume(const std::string &file_name)
    : m_impl(nullptr), m_offset{0.0f, 0.0f, 0.0f}, m_global_translation{0, 0, 0}, m_global_rotation{0, 0, 0} {
    std::cout << "Reading TSDF from " << file_name << std::endl;
    if (tsdf_b200_volume_load(file_name.c_str(), &m_impl) != 0 || !m_impl) throw std::invalid_argument("Unable to load file");
    refresh();
}

TSDFVolume::~TSDFVolume() {
    std::cout << "Destroying TSDFVolume" << std::endl;
    release();
}

void TSDFVolume::set_size(uint16_t volume_x, uint16_t volume_y, uint16_t volume_z, float psize_x, float psize_y, float psize_z) {
    if (!((volume_x != 0 && volume_y != 0 && volume_z != 0) && (psize_x != 0 && psize_y != 0 && psize_z != 0)))
        throw std::invalid_argument("Attempt to set TSDFVolume size or physical size to zero");
    const float3 keep = m_offset;
    release();
    die_on(tsdf_b200_volume_create(volume_x, volume_y, volume_z, psize_x, psize_y, psize_z, &m_impl), "Couldn't allocate TSDF volume");
    // the reference clears with m_offset already in place (TSDFVolume.cu:714): keep that order
    die_on(tsdf_b200_volume_set_offset(m_impl, keep.x, keep.y, keep.z), "Couldn't set offset");
    if (keep.x != 0.0f || keep.y != 0.0f || keep.z != 0.0f) die_on(tsdf_b200_volume_clear(m_impl), "Couldn't clear TSDF volume");
    refresh();
}

void TSDFVolume::offset(float ox, float oy, float oz) {
    m_offset = float3{ox, oy, oz};
    die_on(tsdf_b200_volume_set_offset(m_impl, ox, oy, oz), "Couldn't set offset");
}

void TSDFVolume::clear() { die_on(tsdf_b200_volume_clear(m_impl), "Couldn't clear TSDF volume"); }

const float *TSDFVolume::distance_data() const { return tsdf_b200_volume_distance_data(m_impl); }
const float *TSDFVolume::weight_data() const { return tsdf_b200_volume_weight_data(m_impl); }

TSDFVolume::DeformationNode *TSDFVolume::deformation() const {
    return reinterpret_cast<DeformationNode *>(tsdf_b200_volume_deformation(m_impl));
}

void TSDFVolume::set_deformation(DeformationNode *deformation) {
    die_on(tsdf_b200_volume_set_deformation(m_impl, reinterpret_cast<const float *>(deformation)), "Couldn't set deformation");
}

void TSDFVolume::set_distance_data(const float *distance_data) {
    die_on(tsdf_b200_volume_set_distance_data(m_impl, distance_data), "Couldn't set distance data");
}

void TSDFVolume::set_weight_data(const float *weight_data) {
    die_on(tsdf_b200_volume_set_weight_data(m_impl, weight_data), "Couldn't set weight data");
}

void TSDFVolume::deform_mesh(const int, float3 *) const {
    std::cerr << "TSDFVolume::deform_mesh: the non-rigid (SceneFusion) path is outside tsdf_b200's scope; points unchanged" << std::endl;
}

void TSDFVolume::integrate(const uint16_t *depth_map, uint32_t width, uint32_t height, const Camera &camera) {
    assert(depth_map);
    std::cout << "Integrating depth map size " << width << "x" << height << std::endl;
    const Eigen::Matrix4f inv_pose = camera.inverse_pose();
    const Eigen::Matrix3f k = camera.k(), kinv = camera.kinv();
    die_on(tsdf_b200_volume_integrate(m_impl, depth_map, width, height, inv_pose.data(), k.data(), kinv.data()), "Integrate kernel failed");
    std::cout << "Integration finished" << std::endl;
}

bool TSDFVolume::save_to_file(const std::string &file_name) const { return tsdf_b200_volume_save(m_impl, file_name.c_str()) == 0; }

bool TSDFVolume::load_from_file(const std::string &) { return false; }

void TSDFVolume::raycast(uint16_t width, uint16_t height, const Camera &camera, Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                         Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) const {
    GPURaycaster raycaster(width, height);
    raycaster.raycast(*this, camera, vertices, normals);
}

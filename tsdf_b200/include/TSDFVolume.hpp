// TSDFVolume.hpp — drop-in replacement for the reference's TSDFVolume class (reference src/include/TSDFVolume.hpp:21-304),
// implemented as a handle over the tsdf_b200 C-ABI (include/tsdf_b200.h, level 2).  The public surface — nested
// Float3/UInt3/Int3/DeformationNode with their implicit conversions, constructors, accessors, integrate, raycast,
// save_to_file — is what src/Tools/kinfu.cpp, GPURaycaster and the marching cubes code compile against.  Everything
// behind it is different: device arrays live in the C-ABI object, integration and raycast are the sm_100a kernels of
// tsdf_b200/csrc, the 24 B/voxel deformation grid exists only once somebody asks for it.
//
// Error convention kept from the reference: constructors throw std::invalid_argument (TSDFVolume.cu:435,455,662);
// failures of device work print a message and exit(-1) (Utilities/cuda_utilities.cu:5-11).
#ifndef TSDFVolume_hpp
#define TSDFVolume_hpp

#include "Camera.hpp"

#include <Eigen/Core>
#include "vector_types.h"

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <string>

struct tsdf_b200_volume;

class TSDFVolume {
public:
    // One node of the deformation grid as stored in .tsdf files and handed to set_deformation().
    struct DeformationNode {
        float3 translation;
        float3 rotation;
    };

    // float3 with arithmetic; converts to and from CUDA's float3.
    struct Float3 {
        float x, y, z;
        Float3(const float3 &v) : x(v.x), y(v.y), z(v.z) {}
        Float3(float fx = 0.0f, float fy = 0.0f, float fz = 0.0f) : x(fx), y(fy), z(fz) {}
        operator float3() const { return float3{x, y, z}; }
        Float3 operator-(const Float3 &o) const { return Float3(x - o.x, y - o.y, z - o.z); }
        Float3 operator+(const Float3 &o) const { return Float3(x + o.x, y + o.y, z + o.z); }
        Float3 operator/(const float s) const { return Float3(x / s, y / s, z / s); }
        Float3 operator*(const Float3 &o) const { return Float3(x * o.x, y * o.y, z * o.z); }
        float norm() const { return std::sqrt(x * x + y * y + z * z); }
    };

    struct Int3 {
        int16_t x, y, z;
    };

    // Voxel counts; converts to and from CUDA's dim3.
    struct UInt3 {
        unsigned int x, y, z;
        UInt3(const dim3 &d) : x(d.x), y(d.y), z(d.z) {}
        UInt3(uint32_t ux, uint32_t uy, uint32_t uz) : x(ux), y(uy), z(uz) {}
        operator dim3() const { return dim3{x, y, z}; }
    };

    // size in voxels, physical size in mm (TSDFVolume.cu:430-437)
    TSDFVolume(const UInt3 &size = UInt3{64, 64, 64}, const Float3 &physical_size = Float3{3000.0f, 3000.0f, 3000.0f});
    TSDFVolume(uint16_t volume_x, uint16_t volume_y, uint16_t volume_z, float psize_x, float psize_y, float psize_z);
    // load a volume written by save_to_file (TSDFVolume.cu:463-664)
    TSDFVolume(const std::string &file_name);
    ~TSDFVolume();
    TSDFVolume(const TSDFVolume &) = delete;               // owns device memory (the reference's is not copyable in practice)
    TSDFVolume &operator=(const TSDFVolume &) = delete;

    // Reallocates and clears; the offset is kept (TSDFVolume.cu:679-722).
    void set_size(uint16_t volume_x, uint16_t volume_y, uint16_t volume_z, float psize_x, float psize_y, float psize_z);

    UInt3 size() const { return m_size; }
    Float3 voxel_size() const { return m_voxel_size; }
    Float3 physical_size() const { return m_physical_size; }
    float truncation_distance() const { return m_truncation_distance; }

    // World position (mm) of the front-left-bottom corner of voxel (0,0,0).
    void offset(float ox, float oy, float oz);
    Float3 offset() const { return m_offset; }

    // weights <- 0, distances <- truncation distance, deformation grid <- voxel centres (TSDFVolume.cu:812-845)
    void clear();

    size_t index(int x, int y, int z) const { return x + (y * m_size.x) + (z * m_size.x * m_size.y); }

    // DEVICE pointers, x fastest, valid on any stream once the call that produced the data has returned.
    const float *distance_data() const;
    const float *weight_data() const;
    // Materialises the node array on first use; the volume stops assuming the identity grid from then on.
    DeformationNode *deformation() const;

    // HOST -> device, vx*vy*vz elements each (TSDFVolume.cu:729-755).
    void set_deformation(DeformationNode *deformation);
    void set_distance_data(const float *distance_data);
    void set_weight_data(const float *weight_data);

    float3 global_rotation() const { return m_global_rotation; }
    float3 global_translation() const { return m_global_translation; }

    // Non-rigid SceneFusion only (TSDFVolume.cu:263-291): not part of the rigid hot path; reports and leaves the points alone.
    void deform_mesh(const int num_points, float3 *points) const;

    // depth_map: width*height u16 millimetres in HOST memory, 0 = no measurement (TSDFVolume.cu:861-902).
    void integrate(const uint16_t *depth_map, uint32_t width, uint32_t height, const Camera &camera);

    bool save_to_file(const std::string &file_name) const;
    bool load_from_file(const std::string &file_name);      // a stub returning false in the reference too (:1035-1047)

    // vertices / normals: 3 x (width*height), pixel index y*width + x, NaN vertex = no surface (TSDFVolume.cu:1054-1058).
    void raycast(uint16_t width, uint16_t height, const Camera &camera, Eigen::Matrix<float, 3, Eigen::Dynamic> &vertices,
                 Eigen::Matrix<float, 3, Eigen::Dynamic> &normals) const;

    // Not in the reference: the C-ABI object behind this volume (GPURaycaster and extract_surface forward through it).
    tsdf_b200_volume *c_abi() const { return m_impl; }

private:
    void refresh();                // re-read size / voxel size / truncation distance from the C-ABI object
    void release();

    tsdf_b200_volume *m_impl;
    dim3 m_size;
    float3 m_physical_size;
    float3 m_offset;
    float3 m_voxel_size;
    float m_truncation_distance;
    float m_max_weight;
    float3 m_global_translation;
    float3 m_global_rotation;
};
#endif /* TSDFVolume_hpp */

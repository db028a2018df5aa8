// MarkAndSweepMC.cpp — extract_surface / extract_surface_ms of the reference (src/MarchingCubes/MarkAndSweepMC.cu:390-555)
// over tsdf_b200_mc_extract.
#include "../include/MarkAndSweepMC.hpp"
#include "../../include/tsdf_b200.h"

#include <cstdlib>
#include <iostream>

void extract_surface_ms(const TSDFVolume *const volume, int &num_vertices, float3 *&d_mesh_vertices,
                        int *&d_mesh_vertex_voxel_indices, uint8_t *&d_mesh_vertex_voxel_count) {
    std::cout << "Extracting surface" << std::endl;
    num_vertices = 0;
    d_mesh_vertices = nullptr;
    d_mesh_vertex_voxel_indices = nullptr;
    d_mesh_vertex_voxel_count = nullptr;
    float *d_vertices = nullptr;
    unsigned long long count = 0;
    // through the volume object: a volume sharded over several GPUs (TSDF_NGPUS) extracts per slab
    const int rc = tsdf_b200_volume_extract_mesh(volume->c_abi(), &d_vertices, &count);
    if (rc != 0) {
        std::cerr << "Marching cubes failed" << std::endl << tsdf_b200_strerror(rc) << std::endl;
        std::exit(-1);
    }
    std::cout << "-- found " << count << " vertices" << std::endl;
    if (count == 0) {
        // the reference treats an empty surface as fatal (MarkAndSweepMC.cu:426-429)
        std::cout << "Either no occupied cubes or no vertices. Either way a bit sus." << std::endl;
        std::exit(-1);
    }
    num_vertices = static_cast<int>(count);
    d_mesh_vertices = reinterpret_cast<float3 *>(d_vertices);
}

void extract_surface(const TSDFVolume *volume, std::vector<float3> &vertices, std::vector<int3> &triangles) {
    int num_vertices = 0;
    float3 *d_vertices = nullptr;
    int *d_indices = nullptr;
    uint8_t *d_counts = nullptr;
    extract_surface_ms(volume, num_vertices, d_vertices, d_indices, d_counts);
    const size_t first = vertices.size();
    vertices.resize(first + num_vertices);
    const int rc = tsdf_b200_copy_to_host(vertices.data() + first, d_vertices, size_t(num_vertices) * sizeof(float3));
    tsdf_b200_device_free(d_vertices);
    if (rc != 0) {
        std::cerr << "Couldn't copy mesh vertices to host" << std::endl << tsdf_b200_strerror(rc) << std::endl;
        std::exit(-1);
    }
    // every three consecutive vertices are one triangle, wound (i, i+2, i+1) (MarkAndSweepMC.cu:546-551)
    for (int i = 0; i + 2 < num_vertices; i += 3) triangles.push_back(int3{i, i + 2, i + 1});
}

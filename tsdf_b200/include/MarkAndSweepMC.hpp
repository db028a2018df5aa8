// MarkAndSweepMC.hpp — iso-surface extraction (reference src/include/MarkAndSweepMC.hpp).  extract_surface keeps the
// reference's output contract (MarchingCubes/MarkAndSweepMC.cu:506-555): vertices in ascending cube order (x fastest),
// within a cube in triangle-table order, triangle i = {3i, 3i+2, 3i+1}; the classification, the prefix sum and the
// vertex generation all run on the GPU (tsdf_b200/csrc/mc.cu) instead of a host scan between two kernels.
#ifndef MARK_AND_SWEEP_MC_H
#define MARK_AND_SWEEP_MC_H
#include "../include/TSDFVolume.hpp"

#include <vector>

void extract_surface(const TSDFVolume *volume, std::vector<float3> &vertices, std::vector<int3> &triangles);

// Device-side result for callers that keep working on the GPU: num_vertices and a cudaMalloc'ed array the caller
// frees with cudaFree.  The two voxel bookkeeping outputs of the reference feed only the non-rigid SceneFusion
// pipeline and are returned as nullptr.
void extract_surface_ms(const TSDFVolume *const volume, int &num_vertices, float3 *&d_mesh_vertices,
                        int *&d_mesh_vertex_voxel_indices, uint8_t *&d_mesh_vertex_voxel_count);
#endif

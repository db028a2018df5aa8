// volume.cu — level-2 C-ABI: the TSDF volume object with host-buffer entry points.
//
// Mirrors the host side of the reference's TSDFVolume (src/TSDF/TSDFVolume.cu:396-1058) and
// GPURaycaster::raycast (src/RayCaster/GPURaycaster.cu:432-547): same state, same
// synchronous semantics, but device buffers are persistent (no per-call cudaMalloc/cudaFree),
// work runs on one private stream, and the 24 B/voxel deformation array exists only when
// somebody asks for it.
#include "volume_internal.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <map>
#include <mutex>
#include <unordered_map>

// (tsdf_b200_raycast_ex is declared in include/tsdf_b200.h)

namespace {

size_t nvox(const tsdf_b200_volume *v) { return (size_t)v->nx * v->ny * v->nz; }

// Is `p` host memory the device can address directly (cudaHostAlloc / cudaHostRegister)?  Returns its device alias or null.
void *pinned_alias(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer) return attr.devicePointer;
    cudaGetLastError();          // an unregistered pointer is not an error of the caller's
    return nullptr;
}

void release(tsdf_b200_volume *v) {
    if (v->multi) tsdf::multi_destroy(v);
    cudaFree(v->d_dist); cudaFree(v->d_weight); cudaFree(v->d_deform); cudaFree(v->d_occ);
    cudaFree(v->d_table); cudaFree(v->d_depth); cudaFree(v->d_staged); cudaFree(v->d_vn); cudaFree(v->d_counters); cudaFree(v->d_tiles);
    v->d_tiles = nullptr;
    free(v->h_colour);
    if (v->ev_depth) cudaEventDestroy(v->ev_depth);
    if (v->stream) cudaStreamDestroy(v->stream);
    v->ev_depth = nullptr;
    v->d_dist = v->d_weight = v->d_deform = nullptr; v->d_occ = nullptr; v->d_table = nullptr;
    v->d_depth = nullptr; v->d_staged = nullptr; v->d_vn = nullptr; v->d_counters = nullptr; v->h_colour = nullptr; v->stream = nullptr;
}

int allocate(tsdf_b200_volume *v, uint32_t nx, uint32_t ny, uint32_t nz, float px, float py, float pz, int ngpus) {
    // set_size (TSDFVolume.cu:679-722) takes uint16_t dimensions; the raycast indexes in 32 bits.
    if (nx == 0 || ny == 0 || nz == 0 || nx > 65535 || ny > 65535 || nz > 65535) return TSDF_B200_EINVAL;
    if (!(px != 0 && py != 0 && pz != 0)) return TSDF_B200_EINVAL;
    if ((uint64_t)nx * ny * nz > 0xffffffffull) return TSDF_B200_EINVAL;
    v->nx = nx; v->ny = ny; v->nz = nz;
    v->phys[0] = px; v->phys[1] = py; v->phys[2] = pz;
    tsdf_b200_volume_params(nx, ny, nz, v->phys, v->vs, &v->trunc);
    if (ngpus > 1) {
        // Z-slabs over several GPUs (multi.cu); falls back to one GPU when the box or the volume has room for one slab only
        const int rc = tsdf::multi_create(v, ngpus);
        if (rc != TSDF_B200_EINVAL) return rc;
    }
    const size_t n = nvox(v);
    TSDF_CUDA_TRY(cudaGetDevice(&v->device));
    // A BLOCKING stream: distance_data() / weight_data() / deformation() hand raw device pointers to callers that work on
    // the legacy default stream (the reference's marching cubes and raycaster kernels, SceneFusion writing the deformation
    // grid); the legacy stream and a blocking stream order each other implicitly, as if everything ran on one stream
    // like in the reference (which synchronises the device after every launch).
    TSDF_CUDA_TRY(cudaStreamCreateWithFlags(&v->stream, cudaStreamDefault));
    TSDF_CUDA_TRY(cudaEventCreateWithFlags(&v->ev_depth, cudaEventDisableTiming));
    TSDF_CUDA_TRY(cudaMalloc(&v->d_dist, n * sizeof(float)));
    TSDF_CUDA_TRY(cudaMalloc(&v->d_weight, n * sizeof(float)));
    TSDF_CUDA_TRY(cudaMalloc(&v->d_occ, tsdf_b200_occupancy_bytes(nx, ny, nz)));
    TSDF_CUDA_TRY(cudaMalloc(&v->d_table, TSDF_B200_RAY_TABLE_LEN * sizeof(float)));
    TSDF_CUDA_TRY(cudaMalloc(&v->d_counters, 2 * sizeof(unsigned long long)));
    TSDF_CUDA_TRY(cudaMemsetAsync(v->d_counters, 0, 2 * sizeof(unsigned long long), v->stream));
    int rc = tsdf_b200_ray_table(v->trunc, v->d_table, v->stream);
    if (rc) return rc;
    // Prove the reciprocal division for this volume's voxel sizes (once per distinct size).
    v->fastdiv = 1;
    for (int a = 0; a < 3 && v->fastdiv; a++) {
        bool seen = false;
        for (int b = 0; b < a; b++) seen |= (v->vs[b] == v->vs[a]);
        if (seen) continue;
        unsigned long long bad = 1;
        rc = tsdf_b200_selftest_division(v->vs[a], &bad);
        if (rc) return rc;
        if (bad) v->fastdiv = 0;
    }
    return 0;
}

int ensure_deformation(tsdf_b200_volume *v) {
    if (v->d_deform) return 0;
    if (v->multi) TSDF_CUDA_TRY(cudaSetDevice(v->device));     // materialised on GPU 0 (save_to_file writes it)
    TSDF_CUDA_TRY(cudaMalloc(&v->d_deform, nvox(v) * 6 * sizeof(float)));
    int rc = tsdf_b200_init_deformation(v->d_deform, v->nx, v->ny, v->nz, v->vs, v->off_clear, v->stream);
    if (rc) return rc;
    TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return 0;
}

}  // namespace

extern "C" int tsdf_b200_volume_clear(tsdf_b200_volume *v) {
    if (!v) return TSDF_B200_EINVAL;
    if (v->multi) {
        int rc = tsdf::multi_clear(v);
        if (rc) return rc;
        for (int i = 0; i < 3; i++) v->off_clear[i] = v->off[i];
        v->deform_identity = true;
        cudaSetDevice(v->device);
        if (v->d_deform) { cudaFree(v->d_deform); v->d_deform = nullptr; }
        return 0;
    }
    int rc = tsdf_b200_clear(v->d_dist, v->d_weight, v->nx, v->ny, v->nz, v->trunc, v->d_occ, v->stream);
    if (rc) return rc;
    // clear() rewrites the deformation grid with the CURRENT offset (TSDFVolume.cu:839-841).
    for (int i = 0; i < 3; i++) v->off_clear[i] = v->off[i];
    v->deform_identity = true;
    if (v->d_deform) {
        rc = tsdf_b200_init_deformation(v->d_deform, v->nx, v->ny, v->nz, v->vs, v->off_clear, v->stream);
        if (rc) return rc;
    }
    TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return 0;
}

// GPUs a volume is created on when the caller does not say: TSDF_NGPUS (default 1) — this is how the unchanged kinfu.cpp,
// which only knows `TSDFVolume(size, physical_size)`, gets a volume sharded over the GPUs of the box.
static int env_gpus() {
    const char *e = getenv("TSDF_NGPUS");
    const int n = e ? atoi(e) : 1;
    return n > 1 ? n : 1;
}

extern "C" int tsdf_b200_volume_create_sharded(uint32_t nx, uint32_t ny, uint32_t nz, float px, float py, float pz, int ngpus,
                                               tsdf_b200_volume **out) {
    if (!out) return TSDF_B200_EINVAL;
    *out = nullptr;
    tsdf_b200_volume *v = new (std::nothrow) tsdf_b200_volume();
    if (!v) return TSDF_B200_ENOMEM;
    int rc = allocate(v, nx, ny, nz, px, py, pz, ngpus);
    if (!rc) rc = tsdf_b200_volume_clear(v);
    if (rc) { release(v); delete v; return rc; }
    *out = v;
    return 0;
}

extern "C" int tsdf_b200_volume_gpus(const tsdf_b200_volume *v) {
    if (!v) return 0;
    return v->multi ? (int)v->multi->shards.size() : 1;
}

extern "C" int tsdf_b200_volume_create(uint32_t nx, uint32_t ny, uint32_t nz, float px, float py, float pz,
                                       tsdf_b200_volume **out) {
    if (!out) return TSDF_B200_EINVAL;
    *out = nullptr;
    tsdf_b200_volume *v = new (std::nothrow) tsdf_b200_volume();
    if (!v) return TSDF_B200_ENOMEM;
    int rc = allocate(v, nx, ny, nz, px, py, pz, env_gpus());
    if (!rc) rc = tsdf_b200_volume_clear(v);
    if (rc) { release(v); delete v; return rc; }
    *out = v;
    return 0;
}

extern "C" void tsdf_b200_volume_destroy(tsdf_b200_volume *v) {
    if (!v) return;
    release(v);
    delete v;
}

extern "C" int tsdf_b200_volume_get(const tsdf_b200_volume *v, uint32_t size[3], float physical[3], float voxel[3],
                                    float offset[3], float *trunc, float *max_weight) {
    if (!v) return TSDF_B200_EINVAL;
    if (size) { size[0] = v->nx; size[1] = v->ny; size[2] = v->nz; }
    for (int i = 0; i < 3; i++) {
        if (physical) physical[i] = v->phys[i];
        if (voxel) voxel[i] = v->vs[i];
        if (offset) offset[i] = v->off[i];
    }
    if (trunc) *trunc = v->trunc;
    if (max_weight) *max_weight = v->max_weight;
    return 0;
}

extern "C" int tsdf_b200_volume_get_global(const tsdf_b200_volume *v, float translation[3], float rotation[3]) {
    if (!v) return TSDF_B200_EINVAL;
    for (int i = 0; i < 3; i++) {
        if (translation) translation[i] = v->gtrans[i];
        if (rotation) rotation[i] = v->grot[i];
    }
    return 0;
}

extern "C" int tsdf_b200_volume_set_offset(tsdf_b200_volume *v, float ox, float oy, float oz) {
    if (!v) return TSDF_B200_EINVAL;
    v->off[0] = ox; v->off[1] = oy; v->off[2] = oz;
    return 0;
}

// Sharded volume: the slabs are gathered into a full-size array on GPU 0 at every call (the reference's callers of these
// pointers — its marching cubes and raycaster kernels — are replaced by tsdf_b200_volume_extract_mesh / _raycast, which work
// on the slabs where they are; the gather exists for completeness of the class surface).
extern "C" const float *tsdf_b200_volume_distance_data(const tsdf_b200_volume *cv) {
    if (!cv) return nullptr;
    if (!cv->multi) return cv->d_dist;
    tsdf_b200_volume *v = const_cast<tsdf_b200_volume *>(cv);
    return tsdf::multi_gather_device(v) ? nullptr : v->multi->d_full_dist;
}
extern "C" const float *tsdf_b200_volume_weight_data(const tsdf_b200_volume *cv) {
    if (!cv) return nullptr;
    if (!cv->multi) return cv->d_weight;
    tsdf_b200_volume *v = const_cast<tsdf_b200_volume *>(cv);
    return tsdf::multi_gather_device(v) ? nullptr : v->multi->d_full_weight;
}

extern "C" float *tsdf_b200_volume_deformation(tsdf_b200_volume *v) {
    // a sharded volume is rigid-only: the deformation grid (SceneFusion's non-rigid path) is not distributed
    if (!v || v->multi || ensure_deformation(v)) return nullptr;
    // The caller holds a writable device pointer from now on (SceneFusion writes through it):
    // stop assuming the identity grid.
    v->deform_identity = false;
    return v->d_deform;
}

extern "C" int tsdf_b200_volume_set_distance_data(tsdf_b200_volume *v, const float *host) {
    if (!v || !host) return TSDF_B200_EINVAL;
    if (v->multi) return tsdf::multi_write(v, host, nullptr);
    TSDF_CUDA_TRY(cudaMemcpyAsync(v->d_dist, host, nvox(v) * sizeof(float), cudaMemcpyHostToDevice, v->stream));
    int rc = tsdf_b200_occupancy_rebuild(v->d_dist, v->nx, v->ny, v->nz, v->trunc, v->d_occ, v->stream);
    if (rc) return rc;
    TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return 0;
}

extern "C" int tsdf_b200_volume_set_weight_data(tsdf_b200_volume *v, const float *host) {
    if (!v || !host) return TSDF_B200_EINVAL;
    if (v->multi) return tsdf::multi_write(v, nullptr, host);
    TSDF_CUDA_TRY(cudaMemcpyAsync(v->d_weight, host, nvox(v) * sizeof(float), cudaMemcpyHostToDevice, v->stream));
    TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return 0;
}

extern "C" int tsdf_b200_volume_set_deformation(tsdf_b200_volume *v, const float *host_nodes) {
    if (!v || !host_nodes) return TSDF_B200_EINVAL;
    if (v->multi) return TSDF_B200_ESTATE;                 // rigid-only when sharded
    if (!v->d_deform) TSDF_CUDA_TRY(cudaMalloc(&v->d_deform, nvox(v) * 6 * sizeof(float)));
    TSDF_CUDA_TRY(cudaMemcpyAsync(v->d_deform, host_nodes, nvox(v) * 6 * sizeof(float), cudaMemcpyHostToDevice, v->stream));
    TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    v->deform_identity = false;
    return 0;
}

extern "C" int tsdf_b200_volume_read(const tsdf_b200_volume *v, float *host_dist, float *host_weight) {
    if (!v) return TSDF_B200_EINVAL;
    if (v->multi) return tsdf::multi_read(v, host_dist, host_weight);
    if (host_dist) TSDF_CUDA_TRY(cudaMemcpyAsync(host_dist, v->d_dist, nvox(v) * sizeof(float), cudaMemcpyDeviceToHost, v->stream));
    if (host_weight) TSDF_CUDA_TRY(cudaMemcpyAsync(host_weight, v->d_weight, nvox(v) * sizeof(float), cudaMemcpyDeviceToHost, v->stream));
    TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return 0;
}

extern "C" int tsdf_b200_volume_integrate(tsdf_b200_volume *v, const uint16_t *host_depth, uint32_t width, uint32_t height,
                                          const float inv_pose[16], const float k[9], const float kinv[9]) {
    if (!v || !host_depth || !inv_pose || !k || !kinv || width == 0 || height == 0) return TSDF_B200_EINVAL;
    if (v->multi) return tsdf::multi_integrate(v, host_depth, width, height, inv_pose, k, kinv);
    const size_t npix = (size_t)width * height;
    if (npix > v->depth_cap) {
        cudaFree(v->d_depth); v->d_depth = nullptr; v->depth_cap = 0;
        TSDF_CUDA_TRY(cudaMalloc(&v->d_depth, npix * sizeof(uint16_t)));
        v->depth_cap = npix;
    }
    const size_t staged_bytes = tsdf_b200_depth_staged_bytes(width, height);
    if (staged_bytes > v->staged_cap) {
        // (the texture object the library caches for the old buffer is dropped when its address is reused)
        cudaFree(v->d_staged); v->d_staged = nullptr; v->staged_cap = 0;
        TSDF_CUDA_TRY(cudaMalloc(&v->d_staged, staged_bytes));
        v->staged_cap = staged_bytes;
    }
    const uint16_t *src = host_depth;
    TSDF_CUDA_TRY(cudaMemcpyAsync(v->d_depth, src, npix * sizeof(uint16_t), cudaMemcpyHostToDevice, v->stream));
    TSDF_CUDA_TRY(cudaEventRecord(v->ev_depth, v->stream));
    int rc = tsdf_b200_depth_stage(v->d_depth, width, height, v->d_staged, v->stream);
    if (rc) return rc;
    if (v->counting) TSDF_CUDA_TRY(cudaMemsetAsync(v->d_counters, 0, sizeof(unsigned long long), v->stream));
    const float *deform = v->deform_identity ? nullptr : v->d_deform;
    rc = tsdf_b200_integrate(v->d_dist, v->d_weight, deform, v->nx, v->ny, v->nz, v->vs, v->off_clear, v->off, v->trunc,
                                 inv_pose, k, kinv, width, height, v->d_depth, v->d_staged, 0, v->nz, 0, v->d_occ,
                                 v->counting ? v->d_counters : nullptr, v->stream);
    if (rc) return rc;
    v->counters_stale = true;
    // The call returns once the caller's depth buffer has been read.  The fusion kernels complete in stream order before any
    // later call on this volume reads or returns data (every entry point works on v->stream, a blocking stream that also
    // orders the legacy default stream), so the result is the one of a synchronous call — without a host round trip between
    // integrate and the raycast that follows it.  TSDF_B200_SYNC=1 waits for the kernels here (their errors then surface in
    // this call instead of the next one).
    static const bool sync_all = getenv("TSDF_B200_SYNC") && atoi(getenv("TSDF_B200_SYNC")) != 0;
    if (sync_all) TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    else          TSDF_CUDA_TRY(cudaEventSynchronize(v->ev_depth));
    return 0;
}

extern "C" int tsdf_b200_volume_raycast(const tsdf_b200_volume *cv, uint32_t width, uint32_t height, const float pose[16],
                                        const float kinv[9], float *host_vertices, float *host_normals) {
    tsdf_b200_volume *v = const_cast<tsdf_b200_volume *>(cv);   // scratch buffers only; logical state is untouched
    if (!v || !pose || !kinv || !host_vertices || !host_normals || width == 0 || height == 0) return TSDF_B200_EINVAL;
    if (v->multi) return tsdf::multi_raycast(v, width, height, pose, kinv, host_vertices, host_normals);
    const size_t npix = (size_t)width * height;
    if (npix > v->pix_cap) {
        cudaFree(v->d_vn); v->d_vn = nullptr; v->pix_cap = 0;
        TSDF_CUDA_TRY(cudaMalloc(&v->d_vn, npix * 6 * sizeof(float)));
        v->pix_cap = npix;
    }
    const size_t tile_words = tsdf_b200_raycast_tile_counters(width, height);
    if (tile_words > v->tile_cap) {
        cudaFree(v->d_tiles); v->d_tiles = nullptr; v->tile_cap = 0;
        TSDF_CUDA_TRY(cudaMalloc(&v->d_tiles, tile_words * sizeof(unsigned int)));
        TSDF_CUDA_TRY(cudaMemsetAsync(v->d_tiles, 0, tile_words * sizeof(unsigned int), v->stream));
        v->tile_cap = tile_words;
    }
    float *d_vert = v->d_vn, *d_norm = v->d_vn + 3 * npix;
    // get_vertices (GPURaycaster.cu:432-470): origin = pose translation, rot = top-left 3x3,
    // space_min = offset, space_max = offset + physical size.
    const float origin[3] = { pose[12], pose[13], pose[14] };
    const float rot[9] = { pose[0], pose[1], pose[2], pose[4], pose[5], pose[6], pose[8], pose[9], pose[10] };
    float smin[3], smax[3];
    for (int i = 0; i < 3; i++) { smin[i] = v->off[i]; smax[i] = v->off[i] + v->phys[i]; }
    if (v->counting) TSDF_CUDA_TRY(cudaMemsetAsync(v->d_counters + 1, 0, sizeof(unsigned long long), v->stream));
    // Pinned result buffers that the device can address receive the vertex map and the normal map while the march runs
    // (tile by tile: the transfers overlap the kernel instead of following it); any other buffer gets a copy afterwards.
    float *mirror_v = nullptr, *mirror_n = nullptr;
    if (width % 8 == 0 && height % 4 == 0) {
        void *av = pinned_alias(host_vertices), *an = pinned_alias(host_normals);
        if (av && ((uintptr_t)av & 15u) == 0) mirror_v = static_cast<float *>(av);
        if (an && ((uintptr_t)an & 15u) == 0) mirror_n = static_cast<float *>(an);
    }
    // Pageable result buffers (the Eigen matrices of TSDFVolume::raycast) are filled by the driver's staged copy below.  An own
    // pinned ring (vertex map mirrored during the march, normals by DMA, four host threads copying out) was measured in round
    // 2 and was SLOWER on the bench host (671 against 846 frames/s end to end): one pass over 7.4 MB of host memory is the cost
    // either way, and the driver overlaps its DMA with that pass.
    int rc = tsdf_b200_raycast_fused(v->d_dist, v->nx, v->ny, v->nz, v->vs, smin, smax, v->trunc, origin, rot, kinv, width, height,
                                     v->d_table, v->skipping ? v->d_occ : nullptr, d_vert, d_norm, mirror_v, mirror_n, v->d_tiles,
                                     v->counting ? v->d_counters + 1 : nullptr, v->fastdiv, v->stream);
    if (rc) return rc;
    if (!mirror_v)
        TSDF_CUDA_TRY(cudaMemcpyAsync(host_vertices, d_vert, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, v->stream));
    if (!mirror_n)
        TSDF_CUDA_TRY(cudaMemcpyAsync(host_normals, d_norm, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, v->stream));
    v->counters_stale = true;
    TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
    return 0;
}

extern "C" int tsdf_b200_volume_stats(const tsdf_b200_volume *cv, unsigned long long *n_updated, unsigned long long *n_samples) {
    tsdf_b200_volume *v = const_cast<tsdf_b200_volume *>(cv);
    if (!v) return TSDF_B200_EINVAL;
    if (!v->multi && v->counters_stale && v->counting) {
        // the counters stay on the device until somebody asks (a copy per call would put a host round trip into every frame)
        TSDF_CUDA_TRY(cudaMemcpyAsync(v->h_counters, v->d_counters, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, v->stream));
        TSDF_CUDA_TRY(cudaStreamSynchronize(v->stream));
        v->counters_stale = false;
    }
    if (n_updated) *n_updated = v->h_counters[0];
    if (n_samples) *n_samples = v->h_counters[1];
    return 0;
}

extern "C" int tsdf_b200_volume_set_skipping(tsdf_b200_volume *v, int enabled) {
    if (!v) return TSDF_B200_EINVAL;
    v->skipping = enabled ? 1 : 0;
    return 0;
}

// ---- .tsdf files (TSDFVolume.cu:911-1027 save, :463-664 load) ---------------------------------
// Little-endian: dim3 size, float3 physical, float3 offset, float trunc, float max_weight,
// float3 global_translation, float3 global_rotation (68 B), then float dist[N], float
// weight[N], uchar3 colour[N], DeformationNode{float3 t; float3 r}[N].

extern "C" int tsdf_b200_volume_save(const tsdf_b200_volume *cv, const char *path) {
    tsdf_b200_volume *v = const_cast<tsdf_b200_volume *>(cv);
    if (!v || !path) return TSDF_B200_EINVAL;
    const size_t n = nvox(v);
    const bool had_deform = v->d_deform != nullptr;
    int rc = ensure_deformation(v);
    if (rc) return rc;
    float *h = (float *)malloc(n * 6 * sizeof(float));
    if (!h) return TSDF_B200_ENOMEM;
    FILE *f = fopen(path, "wb");
    if (!f) { free(h); return TSDF_B200_EIO; }
    bool ok = true;
    uint32_t size[3] = { v->nx, v->ny, v->nz };
    ok &= fwrite(size, 4, 3, f) == 3;
    ok &= fwrite(v->phys, 4, 3, f) == 3;
    ok &= fwrite(v->off, 4, 3, f) == 3;
    ok &= fwrite(&v->trunc, 4, 1, f) == 1;
    ok &= fwrite(&v->max_weight, 4, 1, f) == 1;
    ok &= fwrite(v->gtrans, 4, 3, f) == 3;
    ok &= fwrite(v->grot, 4, 3, f) == 3;
    ok &= tsdf_b200_volume_read(v, h, nullptr) == 0 && fwrite(h, 4, n, f) == n;        // (gathers the slabs of a sharded volume)
    ok &= tsdf_b200_volume_read(v, nullptr, h) == 0 && fwrite(h, 4, n, f) == n;
    cudaError_t e = cudaSuccess;
    if (v->h_colour) {
        ok &= fwrite(v->h_colour, 3, n, f) == n;
    } else {   // the reference never initialises colours (TSDFVolume.cu:835); write zeros
        memset(h, 0, n * 3);
        ok &= fwrite(h, 3, n, f) == n;
    }
    e = cudaMemcpy(h, v->d_deform, n * 24, cudaMemcpyDeviceToHost);
    ok &= e == cudaSuccess && fwrite(h, 24, n, f) == n;
    ok &= fclose(f) == 0;
    free(h);
    if (!had_deform && v->deform_identity) { cudaFree(v->d_deform); v->d_deform = nullptr; }
    return ok ? 0 : TSDF_B200_EIO;
}

extern "C" int tsdf_b200_volume_load(const char *path, tsdf_b200_volume **out) {
    if (!path || !out) return TSDF_B200_EINVAL;
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return TSDF_B200_EIO;
    uint32_t size[3]; float phys[3], off[3], trunc, maxw, gt[3], gr[3];
    bool ok = fread(size, 4, 3, f) == 3 && fread(phys, 4, 3, f) == 3 && fread(off, 4, 3, f) == 3 &&
              fread(&trunc, 4, 1, f) == 1 && fread(&maxw, 4, 1, f) == 1 && fread(gt, 4, 3, f) == 3 && fread(gr, 4, 3, f) == 3;
    if (!ok) { fclose(f); return TSDF_B200_EIO; }
    tsdf_b200_volume *v = new (std::nothrow) tsdf_b200_volume();
    if (!v) { fclose(f); return TSDF_B200_ENOMEM; }
    // a file carries its own deformation grid, which is used verbatim: one GPU (sharded volumes are rigid-only)
    int rc = allocate(v, size[0], size[1], size[2], phys[0], phys[1], phys[2], 1);
    if (rc) { fclose(f); release(v); delete v; return rc; }
    // The load constructor recomputes only the voxel size; everything else is taken from the file.
    for (int i = 0; i < 3; i++) { v->off[i] = off[i]; v->off_clear[i] = 0.f; v->gtrans[i] = gt[i]; v->grot[i] = gr[i]; }
    if (trunc != v->trunc) {
        v->trunc = trunc;
        rc = tsdf_b200_ray_table(v->trunc, v->d_table, v->stream);
    }
    v->max_weight = maxw;
    const size_t n = nvox(v);
    float *h = (float *)malloc(n * 6 * sizeof(float));
    v->h_colour = (uint8_t *)malloc(n * 3);
    ok = h && v->h_colour;
    cudaError_t e = cudaSuccess;
    if (ok) ok = fread(h, 4, n, f) == n && (e = cudaMemcpy(v->d_dist, h, n * 4, cudaMemcpyHostToDevice)) == cudaSuccess;
    if (ok) ok = fread(h, 4, n, f) == n && (e = cudaMemcpy(v->d_weight, h, n * 4, cudaMemcpyHostToDevice)) == cudaSuccess;
    if (ok) ok = fread(v->h_colour, 3, n, f) == n;
    if (ok) ok = fread(h, 24, n, f) == n && (e = cudaMalloc(&v->d_deform, n * 24)) == cudaSuccess &&
                 (e = cudaMemcpy(v->d_deform, h, n * 24, cudaMemcpyHostToDevice)) == cudaSuccess;
    fclose(f);
    free(h);
    v->deform_identity = false;   // the stored field is used verbatim (TSDFVolume.cu:618-652)
    if (ok && !rc) rc = tsdf_b200_occupancy_rebuild(v->d_dist, v->nx, v->ny, v->nz, v->trunc, v->d_occ, v->stream);
    if (ok && !rc && cudaStreamSynchronize(v->stream) != cudaSuccess) ok = false;
    if (!ok || rc) { release(v); delete v; return rc ? rc : (e != cudaSuccess ? (int)e : TSDF_B200_EIO); }
    *out = v;
    return 0;
}

// extract_surface_ms of the reference (MarchingCubes/MarkAndSweepMC.cu:390-497) on the volume object: one call for a whole
// volume on one GPU, per slab and concatenated in slab order (= the reference's cube order) when the volume is sharded.
extern "C" int tsdf_b200_volume_extract_mesh(const tsdf_b200_volume *cv, float **d_vertices_out, unsigned long long *n_vertices_out) {
    tsdf_b200_volume *v = const_cast<tsdf_b200_volume *>(cv);
    if (!v || !d_vertices_out || !n_vertices_out) return TSDF_B200_EINVAL;
    *d_vertices_out = nullptr;
    *n_vertices_out = 0;
    if (v->multi) return tsdf::multi_extract_mesh(v, d_vertices_out, n_vertices_out);
    return tsdf_b200_mc_extract(v->d_dist, v->nx, v->ny, v->nz, 0, 0, v->nz - 1, v->vs, v->off, d_vertices_out, n_vertices_out, v->stream);
}

// ---- pooled pinned host memory (tsdf_b200_host_alloc / _free) ----------------------------------------------------------------
namespace {
struct HostPool {
    std::mutex m;
    std::unordered_map<void *, size_t> pinned;          // every live pinned block, in use or pooled
    std::multimap<size_t, void *> free_blocks;          // pooled blocks by size
    size_t pooled_bytes = 0;
};
HostPool &host_pool() { static HostPool *p = new HostPool(); return *p; }     // never destroyed: outlives static destructors
constexpr size_t kPinnedMin = 256 * 1024, kPoolMax = (size_t)1 << 30;
}  // namespace

extern "C" void *tsdf_b200_host_alloc(size_t bytes) {
    if (bytes < kPinnedMin) return malloc(bytes ? bytes : 1);
    HostPool &P = host_pool();
    {
        std::lock_guard<std::mutex> lk(P.m);
        auto it = P.free_blocks.lower_bound(bytes);
        if (it != P.free_blocks.end() && it->first <= bytes + bytes / 4) {      // a pooled block of (nearly) this size
            void *p = it->second;
            P.pooled_bytes -= it->first;
            P.free_blocks.erase(it);
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();                          // no device, or no pinnable memory left: ordinary memory does the job
        return malloc(bytes);
    }
    std::lock_guard<std::mutex> lk(P.m);
    P.pinned[p] = bytes;
    return p;
}

extern "C" void tsdf_b200_host_free(void *p) {
    if (!p) return;
    HostPool &P = host_pool();
    {
        std::lock_guard<std::mutex> lk(P.m);
        auto it = P.pinned.find(p);
        if (it == P.pinned.end()) { free(p); return; }
        if (P.pooled_bytes + it->second <= kPoolMax) {
            P.free_blocks.emplace(it->second, p);
            P.pooled_bytes += it->second;
            return;
        }
        P.pinned.erase(it);
    }
    cudaFreeHost(p);
}

// exchange.cu — multi-GPU surface exchange over NVLink peer memory.
//
// The reference is single-GPU; the sharded engine keeps the reference's volume semantics
// (src/TSDF/TSDFVolume.cu:889-893 integrate, src/RayCaster/GPURaycaster.cu:479-482 raycast) and adds one exchange:
// after the integrate every rank PUSHES the voxels of the surface bricks it owns straight into a full-size, mostly
// untouched copy of the distance volume on every GPU of the box (peer stores, no staging buffer, no sizes to agree
// on), so that the raycast can be sharded over the IMAGE instead of the volume: each rank marches its pixel tiles
// against its own copy with the single-GPU kernel — bit-identical samples, balanced work whatever slab the surface
// happens to lie in.
//
// What a ray can read (raycast.cu): away from the low faces of the volume a sample whose voxel lies in brick b is
// evaluated only when b is flagged, and it then reads voxels of [8b-1, 8b+8]^3: every brick with a flagged brick in its
// 27-neighbourhood is published.  Voxel layer 0 of each axis is the exception — the march never skips it, flagged or not,
// because the reference extrapolates there (u in [-0.5, 0), raycast.cu "off_low_edge") and an extrapolated sample of two
// positive corners can be <= 0 — and its samples read voxel layers 0 and 1 of that axis: every brick on a low face
// (bx == 0, by == 0 or bz == 0) is therefore published every frame as well.  Everything else in a replica may be stale
// or never written.
#include "common.cuh"
#include <string.h>

namespace tsdf {

struct PushParams {
    const float *src;            // this rank's slabs, back to back, slab_planes + 1 planes each (owned + halo)
    uint32_t nx, ny, nz;
    uint32_t slab_planes, world, rank;
    const uint8_t *occ;          // brick flags of the WHOLE volume, merged over the ranks
    uint32_t nbx, nby, nbz;
    uint32_t n_dst;
    float *dst[TSDF_B200_MAX_PEERS];
    unsigned long long *n_bricks;
};

// One block per owned brick: 128 threads = 64 rows (y, z) x 2 halves of 4 voxels.
__global__ void __launch_bounds__(128)
bricks_push_kernel(const __grid_constant__ PushParams P) {
    const uint32_t bx = blockIdx.x % P.nbx, by = blockIdx.x / P.nbx;
    const uint32_t layers_per_slab = P.slab_planes / TSDF_B200_BRICK;
    const uint32_t j = blockIdx.y / layers_per_slab, lb = blockIdx.y % layers_per_slab;    // owned slab, layer inside it
    const uint32_t bz = (j * P.world + P.rank) * layers_per_slab + lb;                      // global brick layer
    if (bz >= P.nbz) return;
    int wanted = (bx == 0u || by == 0u || bz == 0u) ? 1 : 0;          // low-face bricks: always read by the march
    if (!wanted && threadIdx.x < 27) {
        const int cx = (int)bx + (int)(threadIdx.x % 3) - 1, cy = (int)by + (int)(threadIdx.x / 3 % 3) - 1,
                  cz = (int)bz + (int)(threadIdx.x / 9) - 1;
        if (cx >= 0 && cy >= 0 && cz >= 0 && cx < (int)P.nbx && cy < (int)P.nby && cz < (int)P.nbz)
            wanted = P.occ[((size_t)cz * P.nby + cy) * P.nbx + cx];
    }
    if (!__syncthreads_or(wanted)) return;
    if (threadIdx.x == 0 && P.n_bricks) atomicAdd(P.n_bricks, 1ull);

    const uint32_t row = threadIdx.x >> 1;
    const uint32_t x = bx * TSDF_B200_BRICK + (threadIdx.x & 1u) * 4u;
    const uint32_t y = by * TSDF_B200_BRICK + (row & 7u);
    const uint32_t zi = row >> 3;
    const uint32_t zg = bz * TSDF_B200_BRICK + zi;                                       // global plane
    const uint32_t zl = j * (P.slab_planes + 1u) + lb * TSDF_B200_BRICK + zi;          // plane in this rank's arrays
    if (y >= P.ny || zg >= P.nz || x >= P.nx) return;
    const size_t from = ((size_t)zl * P.ny + y) * P.nx + x, to = ((size_t)zg * P.ny + y) * P.nx + x;
    if ((P.nx & 3u) == 0) {
        const float4 v = *reinterpret_cast<const float4 *>(P.src + from);
        for (uint32_t d = 0; d < P.n_dst; d++) *reinterpret_cast<float4 *>(P.dst[d] + to) = v;
    } else {
        for (uint32_t i = 0; i < 4 && x + i < P.nx; i++) {
            const float v = P.src[from + i];
            for (uint32_t d = 0; d < P.n_dst; d++) P.dst[d][to + i] = v;
        }
    }
}

__global__ void __launch_bounds__(256) fill_kernel(float *p, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void __launch_bounds__(256) fill_i64_kernel(long long *p, size_t n, long long v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace tsdf

using namespace tsdf;

extern "C" int tsdf_b200_fill_i64(long long *d_ptr, size_t count, long long value, void *stream) {
    if (!d_ptr) return TSDF_B200_EINVAL;
    fill_i64_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>(d_ptr, count, value);
    return (int)cudaGetLastError();
}

extern "C" int tsdf_b200_bricks_push(const float *d_dist_local, uint32_t nx, uint32_t ny, uint32_t nz,
                                     uint32_t slab_planes, uint32_t world, uint32_t rank,
                                     const uint8_t *d_occ_global, uint32_t n_dst, float *const *d_dst,
                                     unsigned long long *d_n_bricks, void *stream) {
    if (!d_dist_local || !d_occ_global || !d_dst || n_dst == 0 || n_dst > TSDF_B200_MAX_PEERS) return TSDF_B200_EINVAL;
    if (world == 0 || rank >= world || slab_planes == 0 || slab_planes % TSDF_B200_BRICK != 0) return TSDF_B200_EINVAL;
    if (nx == 0 || ny == 0 || nz == 0) return TSDF_B200_EINVAL;
    PushParams P;
    P.src = d_dist_local; P.nx = nx; P.ny = ny; P.nz = nz;
    P.slab_planes = slab_planes; P.world = world; P.rank = rank;
    P.occ = d_occ_global;
    const BrickDims nb = brick_dims(nx, ny, nz);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.n_dst = n_dst;
    for (uint32_t i = 0; i < TSDF_B200_MAX_PEERS; i++) P.dst[i] = i < n_dst ? d_dst[i] : nullptr;
    for (uint32_t i = 0; i < n_dst; i++) if (!P.dst[i]) return TSDF_B200_EINVAL;
    P.n_bricks = d_n_bricks;
    // slabs this rank owns: global slabs rank, rank + world, ...
    const uint32_t n_slabs = (nz + slab_planes - 1) / slab_planes;
    const uint32_t owned = n_slabs > rank ? (n_slabs - rank + world - 1) / world : 0;
    if (owned == 0) return 0;
    const uint32_t layers = owned * (slab_planes / TSDF_B200_BRICK);
    if (layers > 65535 || (uint64_t)nb.bx * nb.by > 0x7fffffffull) return TSDF_B200_EINVAL;
    dim3 grid(nb.bx * nb.by, layers);
    bricks_push_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P);
    return (int)cudaGetLastError();
}

// ---- peer memory: plain cudaMalloc blocks exported / opened with CUDA IPC ---------------------------------------------------
extern "C" int tsdf_b200_peer_alloc(size_t bytes, void **d_ptr, unsigned char handle[TSDF_B200_PEER_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == TSDF_B200_PEER_HANDLE_BYTES, "handle size");
    if (!d_ptr || bytes == 0) return TSDF_B200_EINVAL;
    *d_ptr = nullptr;
    TSDF_CUDA_TRY(cudaMalloc(d_ptr, bytes));
    if (handle) {
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, *d_ptr);
        if (e != cudaSuccess) { cudaFree(*d_ptr); *d_ptr = nullptr; return (int)e; }
        memcpy(handle, &h, sizeof(h));
    }
    return 0;
}

extern "C" int tsdf_b200_peer_open(const unsigned char handle[TSDF_B200_PEER_HANDLE_BYTES], void **d_ptr) {
    if (!handle || !d_ptr) return TSDF_B200_EINVAL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *d_ptr = nullptr;
    TSDF_CUDA_TRY(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int tsdf_b200_peer_close(void *d_ptr) {
    if (!d_ptr) return 0;
    return (int)cudaIpcCloseMemHandle(d_ptr);
}

extern "C" int tsdf_b200_peer_free(void *d_ptr) {
    if (!d_ptr) return 0;
    return (int)cudaFree(d_ptr);
}

extern "C" int tsdf_b200_fill_f32(float *d_ptr, size_t count, float value, void *stream) {
    if (!d_ptr) return TSDF_B200_EINVAL;
    fill_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(d_ptr, count, value);
    return (int)cudaGetLastError();
}

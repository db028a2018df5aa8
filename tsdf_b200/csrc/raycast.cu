// raycast.cu — per-pixel raycast of the TSDF for sm_100a.
//
// Replaces process_ray + compute_normals (reference src/RayCaster/GPURaycaster.cu:265-377,
// 393-427).  The reference marches every ray with a fixed step and re-reads 8 voxels per
// sample.  Three facts make a faster march that returns the SAME bits:
//   1. the parameter sequence t_k (t += step from 0, :316,324,360) does not depend on the
//      ray, so it is tabulated once (tsdf_b200_ray_table) and a ray can jump to any k;
//   2. the loop ends at the FIRST sample <= 0 (the back-face branch is dead code because
//      `float tsdf` shadows the outer variable, :311,329,332), so samples that are provably
//      positive need not be evaluated: bricks whose voxels (with a 1-voxel apron) are all
//      inside a positive band make every trilinear sample inside them positive;
//   3. ~10 consecutive samples share their 8 corner voxels, which are kept in registers.
// Every sample that IS evaluated uses the reference's exact operation order.
#include "common.cuh"
#include <math_constants.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <mutex>

namespace tsdf {

struct RayParams {
    const float *dist;
    uint32_t nx, ny, nz;
    float vs[3];
    float rvs[3];            // RN(1 / vs)
    float smin[3], smax[3];
    float trunc, step;
    float origin[3];
    M33 rot, kinv;
    uint32_t width, height;
    const float *table;
    const uint8_t *occ;          // per brick: a voxel of the brick or of its 1-voxel apron is outside the positive band
    const uint8_t *occ_d;        // brick distance grid (written by distance_pass_kernel right before the march)
    uint32_t nbx, nby, nbz;      // occupancy grid dimensions (local planes when sharded)
    uint32_t nbz_planes;         // planes held by the array (slab variants; nz for a whole volume)
    float occ_lo, occ_hi;        // positive band
    uint32_t z_base, z_lo, z_hi; // Z-slab: global z of array plane 0; cells owned by this rank start in [z_lo, z_hi)
    // Interleaved slabs (cyc_g > 0, SLAB kernels): global slabs of cyc_s planes are dealt to cyc_g ranks round robin, this
    // rank (cyc_r) stores its slabs back to back, each followed by one halo plane (cyc_s + 1 planes per slab); the occupancy
    // grid covers the WHOLE volume and only bricks of owned slabs (and the halo's) are ever flagged in it
    uint32_t cyc_s, cyc_g, cyc_r;
    float *vertices;
    int32_t *khit;
    long long *keys;
    long long *keys_min;         // SLAB: hits are min-merged into this key map with atomics (may live on a peer GPU) instead of
                                 // being written to `keys`; pixels without a hit in this slab write nothing
    int reset_keys;              // resolve_kernel: put INT64_MAX back into every key it has read
    unsigned long long *n_samples;
    unsigned int *tile_counter;  // work counter for dynamic tile scheduling (zeroed before the launch) or nullptr
    // Continuation queue: the kernel's run time used to be the march of its slowest ray (~400 dependent loop iterations
    // at ~1.3 us each against a median of ~25, tools/ray_iters.py).  A ray still marching after max_iters iterations
    // appends (pixel, next sample) here and a warp that has run out of tiles finishes it (continue_entry), each lane on
    // 1/32 of the remaining samples.
    int2 *queue;                 // every slot holds (-1, -1) before the launch
    unsigned int *queue_count;   // slots handed out to producers
    unsigned int *queue_head;    // slots claimed by consumers
    unsigned int *tiles_finished;
    const unsigned int *cap_word; // iteration cap chosen by march_reset for this frame (0 / nullptr: max_iters)
    uint32_t queue_cap;
    int max_iters;
    int low_skip;                // unflagged bricks on a low face may be skipped through their extrapolated first voxel layer
    int debug_iters;             // TSDF_B200_DEBUG_ITERS: khit receives loop iterations per ray (tuning aid)
    // Image sharding (tsdf_b200_raycast_tiles): of every `tile_stride` consecutive tiles this rank (tile_first) marches one,
    // and stores each vertex into all n_out vertex maps (its own and, through peer memory, the other GPUs')
    uint32_t tile_first, tile_stride;
    uint32_t n_out;
    float *out[TSDF_B200_MAX_PEERS];
    // Second copy of the vertex map in memory that is slow to write in small pieces (pinned host memory seen through the
    // bus): each warp writes its 8x4 tile as 24 aligned 16-byte stores.  Needs width % 8 == 0, height % 4 == 0, base % 16 == 0.
    float *mirror;
    // Fused normals (tsdf_b200_raycast_fused): a tile's pixels may be finished by different warps (rays set aside for the continuation),
    // so a counter per tile says when its 32 vertices are final; the warp that completes a tile tells the tiles whose
    // normals read it (itself, its left and its upper neighbour), and the warp whose signal is the last one a tile waits for
    // computes that tile's normals into `normals` and `mirror_n`.  Both counter arrays are the caller's and are zero between launches.
    float *normals, *mirror_n;
    unsigned int *tile_done, *tile_deps;
};

template <bool FASTDIV>
__device__ __forceinline__ float div_vs(float a, float b, float r) {
    return FASTDIV ? fdiv_recip(a, b, r) : fdiv(a, b);
}

// can_intersect_in_dimension (GPURaycaster.cu:138-181)
__device__ __forceinline__ bool can_intersect(float smin, float smax, float o, float d, float &near_t, float &far_t) {
    bool ok = true;
    if (d == 0) {
        if (o < smin || o > smax) ok = false;
    } else {
        float d0 = fdiv(fsub(smin, o), d);
        float d1 = fdiv(fsub(smax, o), d);
        if (d0 > d1) { float t = d0; d0 = d1; d1 = t; }
        if (d0 > near_t) near_t = d0;
        if (d1 < far_t) far_t = d1;
        if (near_t > far_t) ok = false;
        else if (far_t < 0) ok = false;
    }
    return ok;
}

// compute_near_and_far_t (GPURaycaster.cu:197-251)
__device__ __forceinline__ bool near_far(const float o[3], const float d[3], const float smin[3], const float smax[3],
                                         float &near_t, float &far_t) {
    if (o[0] >= smin[0] && o[0] <= smax[0] && o[1] >= smin[1] && o[1] <= smax[1] && o[2] >= smin[2] && o[2] <= smax[2]) {
        near_t = 0;
        float t[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            t[a] = CUDART_NAN_F;
            if (d[a] > 0) t[a] = fdiv(fsub(smax[a], o[a]), d[a]);
            else if (d[a] < 0) t[a] = fdiv(fsub(smin[a], o[a]), d[a]);
        }
        if (t[0] < t[1]) { far_t = (t[0] < t[2]) ? t[0] : t[2]; }
        else             { far_t = (t[1] < t[2]) ? t[1] : t[2]; }
        return true;
    }
    near_t = -CUDART_INF_F;
    far_t = CUDART_INF_F;
    return can_intersect(smin[0], smax[0], o[0], d[0], near_t, far_t) &&
           can_intersect(smin[1], smax[1], o[1], d[1], near_t, far_t) &&
           can_intersect(smin[2], smax[2], o[2], d[2], near_t, far_t);
}

// Chebyshev distance transform of the brick flags, one separable pass per launch:
//   out(B) = min over j in [-R, R] of max(in(B + j*axis), |j|)        (missing bricks count as empty)
// with in = (flag ? 0 : R) for the first pass.  After the x, y and z passes out(B) is the distance, in bricks and
// capped at R, from B to the nearest flagged brick (min-max over a cube separates because max distributes).
constexpr int kDistCap = 16;
template <int AXIS, bool FIRST>
__global__ void __launch_bounds__(256)
distance_pass_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int nbx, int nby, int nbz) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= nbx) return;
    const size_t stride = AXIS == 0 ? 1 : (AXIS == 1 ? (size_t)nbx : (size_t)nbx * nby);
    const int pos = AXIS == 0 ? x : (AXIS == 1 ? y : z), len = AXIS == 0 ? nbx : (AXIS == 1 ? nby : nbz);
    const size_t base = ((size_t)z * nby + y) * nbx + x;
    auto value = [&](int j) -> int {
        const uint8_t v = in[base + (ptrdiff_t)j * (ptrdiff_t)stride];
        return FIRST ? (v ? 0 : kDistCap) : (int)v;
    };
    int best = value(0);
    for (int j = 1; j < best; j++) {          // a candidate at offset j can only give max(.., j) >= j
        if (pos - j >= 0)  best = min(best, max(value(-j), j));
        if (pos + j < len) best = min(best, max(value(j), j));
    }
    out[base] = (uint8_t)best;
}

// Start-of-frame reset of the march's words in the scratch third of the occupancy buffer, done by the kernel that builds the
// distance grid: [0] work counter, [1] queue slots handed out, [2] queue slots claimed, [3] tiles finished <- 0; every slot of
// the continuation queue <- (-1, -1).  Word [4] persists from frame to frame: the iteration cap of the march.  A low cap
// shortens the kernel as long as few rays reach it (the warps that are out of tiles finish them while the last tiles are
// still marching); when many do, their continuation is the longer part.  The number of rays set aside in the previous frame —
// poses change slowly — says which case this is: cap_lo unless more than thr_hi rays were set aside with it, cap_hi until
// fewer than thr_lo are.  (Results do not depend on the cap.)
struct MarchReset { unsigned int *words; int n_queue_words; int cap_lo, cap_hi; unsigned int thr_lo, thr_hi; };
constexpr int kMarchWords = 8;
__device__ __forceinline__ void march_reset(const MarchReset &M, int global_thread, int global_threads) {
    if (!M.words) return;
    if (global_thread == 0) {
        const unsigned int set_aside = M.words[1], cap_was = M.words[4];
        unsigned int cap = (cap_was == (unsigned int)M.cap_lo || cap_was == (unsigned int)M.cap_hi) ? cap_was : (unsigned int)M.cap_hi;
        if (cap == (unsigned int)M.cap_hi && cap_was == cap && set_aside < M.thr_lo) cap = (unsigned int)M.cap_lo;
        else if (cap == (unsigned int)M.cap_lo && set_aside > M.thr_hi) cap = (unsigned int)M.cap_hi;
        M.words[0] = 0u; M.words[1] = 0u; M.words[2] = 0u; M.words[3] = 0u;
        M.words[4] = cap;
    }
    for (int i = global_thread; i < M.n_queue_words; i += global_threads) M.words[kMarchWords + i] = 0xffffffffu;
}

// The same transform with the passes done in shared memory (the global version is bound by ~16 dependent L2 round trips per
// pass): one block per z-slice does the x and y passes, one block per y-row does the z pass in place.  The x/y kernel
// also resets the words of the march (march_reset).
__device__ __forceinline__ int distance_scan(const uint8_t *v, int i, int pos, int len, int stride) {
    int best = v[i];
    for (int j = 1; j < best; j++) {
        if (pos - j >= 0)  best = min(best, max((int)v[i - j * stride], j));
        if (pos + j < len) best = min(best, max((int)v[i + j * stride], j));
    }
    return best;
}

__global__ void __launch_bounds__(1024)
distance_xy_kernel(const uint8_t *__restrict__ flags, uint8_t *__restrict__ out, int nbx, int nby, const MarchReset reset) {
    extern __shared__ uint8_t s_grid[];
    const int n = nbx * nby;
    uint8_t *a = s_grid, *b = s_grid + n;
    const size_t base = (size_t)blockIdx.x * n;
    march_reset(reset, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = flags[base + i] ? 0 : kDistCap;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = (uint8_t)distance_scan(a, i, i % nbx, nbx, 1);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[base + i] = (uint8_t)distance_scan(b, i, i / nbx, nby, nbx);
}

__global__ void __launch_bounds__(1024)
distance_z_kernel(uint8_t *__restrict__ grid, int nbx, int nby, int nbz) {
    extern __shared__ uint8_t s_grid[];
    const int n = nbx * nbz, y = blockIdx.x;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_grid[i] = grid[((size_t)(i / nbx) * nby + y) * nbx + i % nbx];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        grid[((size_t)(i / nbx) * nby + y) * nbx + i % nbx] = (uint8_t)distance_scan(s_grid, i, i / nbx, nbz, nbx);
}

// The same transform on bit rows, in ONE launch (round 2; 8 us per frame less than the two shared-memory kernels above at 64^3
// bricks, which are bound by the latency of their dependent byte scans on 64 blocks).  A row of bricks along x is a W x 64-bit mask.
// D_0 = the flags; D_{r+1} = D_r dilated by one brick along x (shifts), y and z (OR with the neighbouring rows): D_r holds the
// bricks within Chebyshev distance r of a flagged one, and the capped distance of a brick is the number of r in 0..15 with
// the brick NOT in D_r — counted in five bit-sliced planes with a ripple-carry add.  A warp owns one z-slice (lane l rows
// [l R, (l + 1) R): the y-neighbours are in the lane's own registers or one shuffle away), slices meet through shared memory
// once per dilation.  A block of 32 warps produces kDistOut consecutive slices from the 32 slices around them (fifteen
// dilations reach fifteen slices), so blocks do not talk.  Rows, columns and slices past the end of the grid behave like empty
// bricks that happen to exist: whatever a dilation puts into them is within the claimed distance of a flagged brick, so what
// they hand back is right as well, and nothing needs masking.
constexpr int kDistOut = 32 - 2 * (kDistCap - 1);         // output slices per block (2)
template <int R, int W>
__global__ void __launch_bounds__(1024)
distance_bits_kernel(const uint8_t *__restrict__ flags, uint8_t *__restrict__ out, int nbx, int nby, int nbz, const MarchReset reset) {
    typedef unsigned long long u64;
    extern __shared__ u64 s_bits[];                    // two buffers of [32 slices][32 R rows][W]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    march_reset(reset, blockIdx.x * blockDim.x + tid, gridDim.x * blockDim.x);
    const int zo0 = blockIdx.x * kDistOut;
    const int z = zo0 - (kDistCap - 1) + warp;          // this warp's slice
    const bool is_out = warp >= kDistCap - 1 && warp < kDistCap - 1 + kDistOut && z < nbz;
    constexpr int kSlice = 32 * R * W;                 // words per slice
    u64 m[R][W];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int y = lane * R + j;
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 v = 0;
            if (z >= 0 && z < nbz && y < nby && w * 64 < nbx) {
                const uint8_t *src = flags + ((size_t)z * nby + y) * nbx + w * 64;
                const int n = min(64, nbx - w * 64);
                if (nbx % 16 == 0) {
                    uint32_t half[2] = { 0u, 0u };
                    for (int q = 0; q < n / 16; q++) {
                        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(src) + q);
                        const uint32_t part[4] = { b.x, b.y, b.z, b.w };
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            // one bit per non-zero byte: 0xff per such byte, its top bit moved to bit 0 of the byte, and the
                            // four of them gathered into the top nibble by one multiplication (no two partial products
                            // share a bit position)
                            const uint32_t nz = (__vcmpne4(part[c], 0u) >> 7) & 0x01010101u;
                            half[q >> 1] |= ((nz * 0x10204080u) >> 28) << (16 * (q & 1) + 4 * c);
                        }
                    }
                    v = (u64)half[0] | ((u64)half[1] << 32);
                } else {
                    for (int x = 0; x < n; x++) v |= (u64)(src[x] != 0) << x;
                }
            }
            m[j][w] = v;
        }
    }
    u64 planes[5][R][W];
#pragma unroll
    for (int p = 0; p < 5; p++)
#pragma unroll
        for (int j = 0; j < R; j++)
#pragma unroll
            for (int w = 0; w < W; w++) planes[p][j][w] = 0ull;

    // A slice past either end of the grid holds no flagged brick and hands back nothing its real neighbours do not reach
    // by themselves (a path through it is never shorter): its warp only keeps the barriers company.
    const bool real = z >= 0 && z < nbz;
    if (!real) {
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int i = 0; i < R * W; i++) s_bits[b * (32 * kSlice) + warp * kSlice + lane * (R * W) + i] = 0ull;
    }
    for (int r = 0; r < kDistCap; r++) {
        if (!real) {
            if (r < kDistCap - 1) __syncthreads();
            continue;
        }
        if (is_out) {
            // m holds D_r: one more for every brick outside it
#pragma unroll
            for (int j = 0; j < R; j++)
#pragma unroll
                for (int w = 0; w < W; w++) {
                    u64 carry = ~m[j][w];
#pragma unroll
                    for (int p = 0; p < 5; p++) {
                        const u64 v = planes[p][j][w];
                        planes[p][j][w] = v ^ carry;
                        carry &= v;
                    }
                }
        }
        if (r == kDistCap - 1) break;
        // along y: the rows of the neighbouring lanes that touch this lane's block of rows
        u64 t[R][W];
#pragma unroll
        for (int w = 0; w < W; w++) {
            u64 below = __shfl_up_sync(0xffffffffu, m[R - 1][w], 1), above = __shfl_down_sync(0xffffffffu, m[0][w], 1);
            if (lane == 0) below = 0ull;
            if (lane == 31) above = 0ull;
#pragma unroll
            for (int j = 0; j < R; j++)
                t[j][w] = m[j][w] | (j > 0 ? m[j - 1][w] : below) | (j + 1 < R ? m[j + 1][w] : above);
        }
        // along x, then out to the other slices
        u64 *buf = s_bits + (r & 1) * (32 * kSlice) + warp * kSlice + lane * (R * W);
#pragma unroll
        for (int j = 0; j < R; j++)
#pragma unroll
            for (int w = 0; w < W; w++) {
                u64 left = t[j][w] << 1, right = t[j][w] >> 1;
                if (w > 0) left |= t[j][w - 1] >> 63;
                if (w + 1 < W) right |= t[j][w + 1] << 63;
                m[j][w] = t[j][w] | left | right;
                buf[j * W + w] = m[j][w];
            }
        __syncthreads();
        // along z (a slice at the end of the block's range misses a neighbour: the error creeps inwards one slice per
        // dilation and never reaches the output slices)
#pragma unroll
        for (int j = 0; j < R; j++)
#pragma unroll
            for (int w = 0; w < W; w++) {
                if (warp > 0) m[j][w] |= buf[j * W + w - kSlice];
                if (warp < 31) m[j][w] |= buf[j * W + w + kSlice];
            }
    }

    // bit planes -> one byte per brick, all threads of the block
    __syncthreads();
    u64 *pl = s_bits;                                  // [kDistOut][5][32 R][W]
    if (is_out) {
        const int o = warp - (kDistCap - 1);
#pragma unroll
        for (int p = 0; p < 5; p++)
#pragma unroll
            for (int j = 0; j < R; j++)
#pragma unroll
                for (int w = 0; w < W; w++) pl[((o * 5 + p) * 32 * R + lane * R + j) * W + w] = planes[p][j][w];
    }
    __syncthreads();
    const int n_out = min(kDistOut, nbz - zo0);
    if (nbx % 8 == 0) {
        // eight bricks per thread: a byte of each plane is spread into eight bytes holding one bit each (replicate the
        // byte, keep bit j in byte j, turn "non-zero" into 1 through the carry into bit 7), the planes are summed with
        // their weights, and the eight distances leave as one 8-byte store
        const int per_row = nbx / 8, n_chunks = n_out * nby * per_row;
        for (int i = tid; i < n_chunks; i += blockDim.x) {
            const int c = i % per_row, row = i / per_row, y = row % nby, o = row / nby;
            u64 acc = 0;
#pragma unroll
            for (int p = 0; p < 5; p++) {
                const u64 bits = (pl[((o * 5 + p) * 32 * R + y) * W + (c >> 3)] >> (8 * (c & 7))) & 0xffull;
                const u64 spread = (((bits * 0x0101010101010101ull) & 0x8040201008040201ull) + 0x7f7f7f7f7f7f7f7full) >> 7;
                acc += (spread & 0x0101010101010101ull) << p;
            }
            *reinterpret_cast<u64 *>(out + ((size_t)(zo0 + o) * nby + y) * nbx + 8 * c) = acc;
        }
    } else {
        const int out_bytes = n_out * nby * nbx;
        for (int i = tid; i < out_bytes; i += blockDim.x) {
            const int x = i % nbx, row = i / nbx, y = row % nby, o = row / nby;
            uint32_t v = 0;
#pragma unroll
            for (int p = 0; p < 5; p++) v |= (uint32_t)((pl[((o * 5 + p) * 32 * R + y) * W + (x >> 6)] >> (x & 63)) & 1ull) << p;
            out[(size_t)zo0 * nby * nbx + i] = (uint8_t)v;
        }
    }
}

// Per-ray set-up shared by the march and the resolve kernel: direction, clip, start point.
struct RaySetup { float dir[3], start[3], max_t; bool intersects; };

__device__ __forceinline__ RaySetup ray_setup(const RayParams &P, uint32_t imx, uint32_t imy) {
    RaySetup r;
    // compute_ray_direction_at_pixel (:24-44): uint16 pixel coords, K^-1 then R, NOT normalised.
    const float fx = (float)(int)(uint16_t)imx, fy = (float)(int)(uint16_t)imy;
    float rc[3];
#pragma unroll
    for (int i = 1; i <= 3; i++)
        rc[i - 1] = fadd(fadd(fmul(fx, T33(P.kinv, i, 1)), fmul(fy, T33(P.kinv, i, 2))), T33(P.kinv, i, 3));
#pragma unroll
    for (int i = 1; i <= 3; i++)
        r.dir[i - 1] = fadd(fadd(fmul(T33(P.rot, i, 1), rc[0]), fmul(T33(P.rot, i, 2), rc[1])), fmul(T33(P.rot, i, 3), rc[2]));
    float near_t, far_t;
    r.intersects = near_far(P.origin, r.dir, P.smin, P.smax, near_t, far_t);
#pragma unroll
    for (int a = 0; a < 3; a++) r.start[a] = fsub(fadd(fmul(r.dir[a], near_t), P.origin[a]), P.smin[a]);   // :306
    r.max_t = fsub(far_t, near_t);                                                                         // :317
    return r;
}

// Vertex of a hit at sample k with interpolated value s (:336-348); previous_tsdf == trunc always.
__device__ __forceinline__ void hit_vertex(const RayParams &P, const float dir[3], const float start[3], float t, float s, float ip[3]) {
    float th = t;
    if (s < 0) {
        th = fsub(th, P.step);                                                     // :338
        th = fadd(th, fmul(fdiv(P.trunc, fsub(P.trunc, s)), P.step));              // :341
    }
#pragma unroll
    for (int a = 0; a < 3; a++) ip[a] = fadd(fadd(fmul(dir[a], th), start[a]), P.smin[a]);  // :345-348
}
__device__ __forceinline__ void hit_vertex(const RayParams &P, const RaySetup &r, float t, float s, float ip[3]) {
    hit_vertex(P, r.dir, r.start, t, s, ip);
}

// Largest j >= 0 such that samples k+1 .. k+j all have t <= t_lim (0 if none), using the monotone table.
__device__ __forceinline__ int safe_steps(const float *s_t, int k, float t, float t_gain, float inv_step) {
    if (!(t_gain > 0.0f) || !(t_gain < 1.0e9f)) return 0;
    int j = (int)(t_gain * inv_step * 0.999f);
    if (j <= 0) return 0;
    if (k + j > TSDF_B200_MAX_SAMPLES) j = TSDF_B200_MAX_SAMPLES - k;
    const float t_lim = t + t_gain;
    while (j > 0 && !(s_t[k + j] <= t_lim)) j--;
    return j;
}

// ---- The march, one sample at a time ------------------------------------------------------------------------------------
// State of one ray between two loop iterations.  Every sample that is evaluated uses the reference's operation order;
// everything else only decides which samples need no evaluation.
struct RayDebug { int iters, l1, l2, l3, eval; };

struct RayState {
    float dir[3], start[3], max_t;      // ray_setup (exact)
    float ainv[3];                      // 1 / |dir| (approximate, skipping only); 0 for an axis the ray does not move along
    int k, k_stop;                      // next sample, last sample of the range
    int clx, cly, clz;                  // corner cache key
    float c000, c001, c010, c011, c100, c101, c110, c111;
    float lip_inv, lip_margin;          // level 3: 1 / (Lipschitz bound per step), absolute slack
    bool cpos;                          // all 8 cached corners inside the positive band
    int kh;                             // first sample <= 0 (march_step returned kRayHit) and its value
    float s_hit;
};

// Per-volume constants of the march (warp-uniform).
struct MarchConst {
    float mx[3], hi_adj[3], inv_step;
};

__device__ __forceinline__ MarchConst march_const(const RayParams &P) {
    MarchConst C;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float n = (float)(a == 0 ? P.nx : (a == 1 ? P.ny : P.nz));
        C.mx[a] = fmul(n, P.vs[a]);
        C.hi_adj[a] = fsub(C.mx[a], fdiv(P.vs[a], 10.0f));
    }
    C.inv_step = __frcp_rn(P.step);
    return C;
}

// A ray about to march samples [k_first, k_last] (further narrowed to the slab's parameter interval when SLAB).
template <bool SKIP, bool SLAB>
__device__ __forceinline__ void ray_begin(const RayParams &P, const float *s_t, const MarchConst &C, const RaySetup &R,
                                          int k_first, int k_last, RayState &S) {
#pragma unroll
    for (int a = 0; a < 3; a++) { S.dir[a] = R.dir[a]; S.start[a] = R.start[a]; }
    S.max_t = R.max_t;
#pragma unroll
    for (int a = 0; a < 3; a++) {
        const float ad = fabsf(R.dir[a]);
        S.ainv[a] = (SKIP && ad > 0.0f && ad < 3.0e38f) ? __frcp_rn(ad) : 0.0f;
    }
    S.clx = S.cly = S.clz = -1;
    S.c000 = S.c001 = S.c010 = S.c011 = S.c100 = S.c101 = S.c110 = S.c111 = 0.0f;
    S.cpos = false;
    S.lip_inv = 0.0f; S.lip_margin = 3.0e38f;
    S.kh = -1; S.s_hit = 0.0f;
    int k = k_first, k_stop = k_last;
    if (SLAB && P.cyc_g == 0) {
        // Parameter interval in which a sample's cell can start inside [z_lo, z_hi), with a voxel of slack.
        const float zl = (P.z_lo == 0) ? -3.0e38f : ((float)P.z_lo - 1.0f) * P.vs[2];
        const float zh = (P.z_hi >= P.nz) ? 3.0e38f : ((float)P.z_hi + 1.5f) * P.vs[2];
        float ta = 0.0f, tb = 3.0e38f;
        if (R.dir[2] > 0.0f)      { ta = (zl - R.start[2]) / R.dir[2]; tb = (zh - R.start[2]) / R.dir[2]; }
        else if (R.dir[2] < 0.0f) { ta = (zh - R.start[2]) / R.dir[2]; tb = (zl - R.start[2]) / R.dir[2]; }
        else if (R.start[2] < zl || R.start[2] > zh) { tb = -1.0f; }
        if (tb < 0.0f) k = k_stop + 1;                       // never inside this slab
        else {
            if (ta > 0.0f && ta < 1.0e9f) {
                int ka = (int)(ta * C.inv_step) - 3;
                if (ka > k_stop) ka = k_stop + 1;
                while (ka > 0 && ka <= k_stop && s_t[ka] > ta) ka--;
                if (ka > k) k = ka;
            }
            if (tb < 1.0e9f) {
                int kb = (int)(tb * C.inv_step) + 4;
                if (kb < k_stop) k_stop = kb;
            }
        }
    }
    S.k = k; S.k_stop = k_stop;
}

enum { kRayContinue = 0, kRayHit = 1, kRayEnd = 2 };

// One loop iteration: looks at sample S.k and either reports it as the ray's first sample <= 0 (kRayHit: S.kh, S.s_hit),
// or finds the range exhausted (kRayEnd), or advances S.k past it and past every following sample that is PROVEN positive.
// Every level of skipping ends in the same exit computation — "how many of the next samples stay inside this axis-aligned
// box":
//   level 1  box = the brick (plus cd - 2 whole bricks beyond its exit)        brick distance grid says: empty space
//   level 2  box = the interpolation cell, pulled in by a guard band           all 8 corners in the positive band
//   level 3  box = the same cell, and at most j samples                        the evaluated sample is s > 0 and the
//                                                                              interpolant cannot fall by more than s in j steps
// The cell levels (and the cells of another rank's slab) share one copy of it, which the lanes of a warp run together
// whichever of them they are on (10 % off the march of views dominated by cell steps); level 1 keeps its own.
template <bool FASTDIV, bool SKIP, bool SLAB>
__device__ __forceinline__ int march_step(const RayParams &P, const float *s_t, const MarchConst &C, RayState &S,
                                          uint32_t &samples, RayDebug &dbg) {
    const int k = S.k;
    if (k > S.k_stop) return kRayEnd;
    const float t = s_t[k];
    if (k > 0 && t >= S.max_t) return kRayEnd;                  // sample k>0 exists only if t_k < max_t (:360-365)

    float p[3];
#pragma unroll
    for (int a = 0; a < 3; a++) p[a] = fadd(fmul(S.dir[a], t), S.start[a]);        // :326

    // trilinearly_interpolate (:53-124)
    int vox[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float adj = p[a];
        if (p[a] >= C.mx[a]) adj = C.hi_adj[a];
        if (p[a] < 0.0f) adj = 0.0f;
        vox[a] = f2i(floorf(div_vs<FASTDIV>(adj, P.vs[a], P.rvs[a])));
    }
    const bool oob = vox[0] < 0 || vox[1] < 0 || vox[2] < 0 ||
                     (uint32_t)vox[0] >= P.nx || (uint32_t)vox[1] >= P.ny || (uint32_t)vox[2] >= P.nz;

    // The box of the exit computation: distances from the sample to its low / high faces (guard band already taken off),
    // whole bricks beyond it, whether it may be used at all, and the most samples it may skip.
    float d_lo[3] = { 0.0f, 0.0f, 0.0f }, d_hi[3] = { 0.0f, 0.0f, 0.0f };
    bool box = false;
    int j_cap = 0x7fffffff;

    if (oob) {
        samples++;                                                                  // :77-80: NaN, never a hit
    } else {
        // ---- level 1: empty space, sphere-traced on the brick distance grid -----------------------------
        // cd[B] = Chebyshev distance (in bricks, capped) from brick B to the nearest brick whose voxels or
        // 1-voxel apron leave the positive band.  A sample's fp32 position is within ~1e-3 mm of the real
        // line start + t*dir, i.e. in the brick the real line is in or in a face/edge/corner neighbour of it.
        //   cd >= 2: while the real line stays within cd-2 bricks of B (per axis), every sample on it lies in
        //            a brick at distance <= cd-1 of B, hence empty, hence > 0 — no guard band is needed and
        //            rays grazing brick faces are handled like any other;
        //   cd == 1: B itself is empty but a neighbour is not: skip to the exit of B pulled in by a guard
        //            band (2% of a voxel, >> the rounding error of a sample position), provided the landing
        //            point is itself clear of every face by that band.
        // Voxel layer 0 of each axis, where the reference extrapolates (:87-99, u in [-0.5, 0)): bricks on a low face of the
        // volume are unflagged only if their voxels lie in the TIGHT band [0.8, 1.0001] * trunc (common.cuh), which makes every
        // sample in them positive for u, v, w >= -0.51 — so they are skipped like any other empty brick (P.low_skip; every ray
        // that enters through such a face used to evaluate the ~11 samples of its first voxel).  Without that guarantee
        // (!P.low_skip: coordinates so large that a sample may lie more than 1% of a voxel outside the volume) the landing
        // point and every skipped sample keep 1.05 voxels away from the low faces.
        if (SKIP) {
            const int b[3] = { vox[0] / TSDF_B200_BRICK, vox[1] / TSDF_B200_BRICK, vox[2] / TSDF_B200_BRICK };
            const int bz_local = b[2] - ((SLAB && P.cyc_g == 0) ? (int)(P.z_base / TSDF_B200_BRICK) : 0);
            const bool in_grid = !SLAB || (bz_local >= 0 && bz_local < (int)P.nbz);
            const int cd = in_grid ? (int)__ldg(P.occ_d + ((size_t)bz_local * P.nby + b[1]) * P.nbx + b[0]) : 0;
            if (cd >= 1) {
                bool clear = true, off_low_edge = true;
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float lo = (float)(b[a] * TSDF_B200_BRICK) * P.vs[a];
                    const float hi = (float)((b[a] + 1) * TSDF_B200_BRICK) * P.vs[a];
                    const float g = 0.02f * P.vs[a];
                    const float dlo = p[a] - lo, dhi = hi - p[a];
                    clear = clear && dlo >= g && dhi >= g;
                    if (!P.low_skip) off_low_edge = off_low_edge && (p[a] - 1.05f * P.vs[a] >= 0.0f);
                    d_lo[a] = cd >= 2 ? dlo : dlo - g;
                    d_hi[a] = cd >= 2 ? dhi : dhi - g;
                }
                if (off_low_edge && (cd >= 2 || clear)) {
                    // (its own copy of the exit computation: lanes in empty space go round the loop without waiting for the
                    // lanes of their warp that work cell by cell — measured 2 % faster on frames dominated by empty space
                    // than joining the computation below)
                    const float extra = cd >= 2 ? (float)(cd - 2) : 0.0f;      // whole bricks beyond the exit of B
                    float t_gain = 3.0e30f;
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        if (S.ainv[a] > 0.0f) {
                            float ta = (S.dir[a] > 0.0f ? d_hi[a] : d_lo[a]) * S.ainv[a] + extra * ((float)TSDF_B200_BRICK * P.vs[a] * S.ainv[a]);
                            if (!P.low_skip && S.dir[a] < 0.0f) ta = fminf(ta, (p[a] - 1.05f * P.vs[a]) * S.ainv[a]);
                            t_gain = fminf(t_gain, ta);
                        }
                    }
                    S.k = k + 1 + safe_steps(s_t, k, t, t_gain, C.inv_step);
#ifdef TSDF_RAY_DEBUG
                    dbg.l1++;
#endif
                    return kRayContinue;
                }
            }
        }

        {
            int low[3];
            float uvw[3], lcs[3];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float ctr = fadd(fmul(fadd((float)vox[a], 0.5f), P.vs[a]), 0.0f);      // TSDF_utilities.cu:10-17
                int l = (p[a] < ctr) ? vox[a] - 1 : vox[a];                               // :87-89
                l = max(l, 0);                                                            // :92-94
                const float lc = fadd(fmul(fadd((float)l, 0.5f), P.vs[a]), 0.0f);
                uvw[a] = div_vs<FASTDIV>(fsub(p[a], lc), P.vs[a], P.rvs[a]);             // :98-102
                low[a] = l;
                lcs[a] = lc;
            }
            const float u = uvw[0], v = uvw[1], w = uvw[2];
            const bool uvw_in_cell = u >= 0.0f && u <= 1.0f && v >= 0.0f && v <= 1.0f && w >= 0.0f && w <= 1.0f;
            // the cell as a box, pulled in by the guard band so that `lower` and the weights' range cannot flip inside it
            bool inside = true;
            if (SKIP) {
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    const float g = 0.02f * P.vs[a];
                    d_lo[a] = p[a] - (lcs[a] + g);
                    d_hi[a] = (lcs[a] + P.vs[a] - g) - p[a];
                    inside = inside && d_lo[a] >= 0.0f && d_hi[a] >= 0.0f;
                }
            }
            bool mine = true;
            if (SLAB) {
                if (P.cyc_g > 0) mine = ((uint32_t)low[2] / P.cyc_s) % P.cyc_g == P.cyc_r;
                else             mine = (uint32_t)low[2] >= P.z_lo && (uint32_t)low[2] < P.z_hi;
            }
            if (!mine) {
                // another rank's cell: nothing of it is evaluated here.  It is left the way level 2 leaves a cell that is
                // certainly positive instead of one sample at a time — with interleaved slabs a ray meets such cells at
                // every slab boundary.
                box = SKIP && uvw_in_cell && inside;
            } else {
                if (low[0] != S.clx || low[1] != S.cly || low[2] != S.clz) {
                    S.clx = low[0]; S.cly = low[1]; S.clz = low[2];
                    // tsdf_value_at (TSDF_utilities.cu:29-37): upper clamp, 32-bit index arithmetic
                    const uint32_t x0 = min((uint32_t)S.clx, P.nx - 1), x1 = min((uint32_t)S.clx + 1, P.nx - 1);
                    const uint32_t y0 = P.nx * min((uint32_t)S.cly, P.ny - 1), y1 = P.nx * min((uint32_t)S.cly + 1, P.ny - 1);
                    // array plane of global plane z: whole volume, contiguous slab, or interleaved slabs (+1 halo each)
                    const uint32_t zg0 = min((uint32_t)S.clz, P.nz - 1), zg1 = min((uint32_t)S.clz + 1, P.nz - 1);
                    uint32_t zl0, zl1;
                    if (SLAB && P.cyc_g > 0) {
                        const uint32_t sg = zg0 / P.cyc_s;
                        zl0 = (sg / P.cyc_g) * (P.cyc_s + 1u) + (zg0 - sg * P.cyc_s);
                        zl1 = zl0 + (zg1 - zg0);                      // the halo plane follows the slab's last plane
                    } else {
                        const uint32_t zb = SLAB ? P.z_base : 0u;
                        zl0 = zg0 - zb; zl1 = zg1 - zb;
                    }
                    const uint32_t z0 = P.nx * P.ny * zl0, z1 = P.nx * P.ny * zl1;
                    const float c000 = __ldg(P.dist + (size_t)(z0 + y0 + x0));
                    const float c001 = __ldg(P.dist + (size_t)(z1 + y0 + x0));
                    const float c010 = __ldg(P.dist + (size_t)(z0 + y1 + x0));
                    const float c011 = __ldg(P.dist + (size_t)(z1 + y1 + x0));
                    const float c100 = __ldg(P.dist + (size_t)(z0 + y0 + x1));
                    const float c101 = __ldg(P.dist + (size_t)(z1 + y0 + x1));
                    const float c110 = __ldg(P.dist + (size_t)(z0 + y1 + x1));
                    const float c111 = __ldg(P.dist + (size_t)(z1 + y1 + x1));
                    S.c000 = c000; S.c001 = c001; S.c010 = c010; S.c011 = c011;
                    S.c100 = c100; S.c101 = c101; S.c110 = c110; S.c111 = c111;
                    if (SKIP) {
                        const float cmin = fminf(fminf(fminf(c000, c001), fminf(c010, c011)), fminf(fminf(c100, c101), fminf(c110, c111)));
                        const float cmax = fmaxf(fmaxf(fmaxf(c000, c001), fmaxf(c010, c011)), fmaxf(fmaxf(c100, c101), fmaxf(c110, c111)));
                        const bool finite = (c000 == c000) && (c001 == c001) && (c010 == c010) && (c011 == c011) &&
                                            (c100 == c100) && (c101 == c101) && (c110 == c110) && (c111 == c111);
                        S.cpos = finite && cmin >= P.occ_lo && cmax <= P.occ_hi;
                        // level 3 (below): the interpolant is multilinear, so with weights in [0,1] its derivative
                        // along u is a convex combination of the four corner differences along x, and likewise for
                        // v and w: |ds/dt| <= sum_a G_a * |dir_a| / voxel_a with G_a the largest corner difference
                        // (not needed for a cell that level 2 skips as a whole)
                        S.lip_inv = 0.0f;
                        if (!S.cpos) {
                            const float gx = fmaxf(fmaxf(fabsf(c100 - c000), fabsf(c101 - c001)), fmaxf(fabsf(c110 - c010), fabsf(c111 - c011)));
                            const float gy = fmaxf(fmaxf(fabsf(c010 - c000), fabsf(c011 - c001)), fmaxf(fabsf(c110 - c100), fabsf(c111 - c101)));
                            const float gz = fmaxf(fmaxf(fabsf(c001 - c000), fabsf(c011 - c010)), fmaxf(fabsf(c101 - c100), fabsf(c111 - c110)));
                            const float lt = (gx * fabsf(S.dir[0]) * P.rvs[0] + gy * fabsf(S.dir[1]) * P.rvs[1] + gz * fabsf(S.dir[2]) * P.rvs[2]) * P.step;
                            // per-step bound inflated by 1% (rounding of the bound itself)
                            S.lip_inv = (finite && lt < 3.0e37f) ? 0.99f / fmaxf(lt, 1.0e-30f) : 0.0f;
                            // an evaluated sample differs from the ideal interpolant at the ideal position by the rounding of
                            // p (a few ulps of a coordinate as large as the volume: < 1e-6 * n voxels, four times the estimate)
                            // times the gradient bound, plus ~10 roundings of terms no larger than the largest corner
                            S.lip_margin = (gx + gy + gz) * (1.0e-6f * (float)(max(max(P.nx, P.ny), P.nz) + 2u)) +
                                           1.0e-5f * fmaxf(fabsf(cmin), fabsf(cmax));
                        }
                    }
                }
                // ---- level 2: a cell whose 8 corners are all in the positive band ----------------------------
                // With weights in [0,1] every product is >= 0 and one is >= corner/8 > 0, so the sample is > 0
                // without evaluating it; the same holds for every following sample that stays inside the cell.
                if (SKIP && S.cpos && uvw_in_cell) {
                    box = inside;
#ifdef TSDF_RAY_DEBUG
                    dbg.l2++;
#endif
                } else {
#ifdef TSDF_RAY_DEBUG
                    dbg.eval++;
#endif
                    const float u1 = fsub(1.0f, u), v1 = fsub(1.0f, v), w1 = fsub(1.0f, w);
                    float s = fmul(fmul(fmul(S.c000, u1), v1), w1);                              // :114-121
                    s = fadd(s, fmul(fmul(fmul(S.c001, u1), v1), w));
                    s = fadd(s, fmul(fmul(fmul(S.c010, u1), v), w1));
                    s = fadd(s, fmul(fmul(fmul(S.c011, u1), v), w));
                    s = fadd(s, fmul(fmul(fmul(S.c100, u), v1), w1));
                    s = fadd(s, fmul(fmul(fmul(S.c101, u), v1), w));
                    s = fadd(s, fmul(fmul(fmul(S.c110, u), v), w1));
                    s = fadd(s, fmul(fmul(fmul(S.c111, u), v), w));
                    samples++;
                    if (s <= 0) {
                        S.kh = k;
                        S.s_hit = s;
                        return kRayHit;
                    }
                    // ---- level 3: samples that cannot have reached zero yet -----------------------------------
                    // While the ray stays in this cell sample k+j is at least s - j * (Lipschitz bound per step) -
                    // rounding slack: the first j for which that is still positive need no evaluation.  This is what
                    // bounds the cost of a ray that skims a surface at a small positive distance for thousands of samples.
                    if (SKIP && S.lip_inv > 0.0f) {
                        const int j = (int)fminf((s - S.lip_margin) * S.lip_inv, 5000.0f);       // NaN / negative -> 0 or less
                        if (j >= 1 && uvw_in_cell) {
                            box = inside;
                            j_cap = j;
#ifdef TSDF_RAY_DEBUG
                            dbg.l3++;
#endif
                        }
                    }
                }
            }
        }
    }

    // ---- the exit computation of the cell levels ---------------------------------------------------------------------
    int j = 0;
    if (SKIP && box) {
        float t_gain = 3.0e30f;
#pragma unroll
        for (int a = 0; a < 3; a++)
            if (S.ainv[a] > 0.0f) t_gain = fminf(t_gain, (S.dir[a] > 0.0f ? d_hi[a] : d_lo[a]) * S.ainv[a]);
        j = min(safe_steps(s_t, k, t, t_gain, C.inv_step), j_cap);
    }
    S.k = k + 1 + j;
    return kRayContinue;
}

// The march of one ray over samples [k_first, k_last].  Returns -1 when the range is finished (kh >= 0: first sample <= 0
// and its value), or, after max_iters loop iterations, the sample to resume from.
template <bool FASTDIV, bool SKIP, bool SLAB>
__device__ __forceinline__ int march_ray(const RayParams &P, const float *s_t, const RaySetup &R, int k_first, int k_last,
                                         int max_iters, int &kh, float &s_hit, uint32_t &samples, RayDebug &dbg) {
    const MarchConst C = march_const(P);
    RayState S;
    ray_begin<SKIP, SLAB>(P, s_t, C, R, k_first, k_last, S);
    int iters = 0, resume = -1;
    while (true) {
        iters++;
        // (the range test comes first: a ray at the end of its range is finished, not set aside)
        if (iters > max_iters && S.k <= S.k_stop && !(S.k > 0 && s_t[S.k] >= S.max_t)) { resume = S.k; break; }
        const int st = march_step<FASTDIV, SKIP, SLAB>(P, s_t, C, S, samples, dbg);
        if (st != kRayContinue) break;
    }
    kh = S.kh; s_hit = S.s_hit;
    dbg.iters += iters;
    return resume;
}

// ---- Tile bookkeeping of the fused raycast (normals and pinned mirrors written while the march runs) ---------------------
// compute_normals (GPURaycaster.cu:393-427) for one pixel from its own vertex a, its right neighbour r and the one below, b.
__device__ __forceinline__ void normal_of(const float a[3], const float r[3], const float b[3], float n[3]) {
    const float v2x = fsub(r[0], a[0]), v2y = fsub(r[1], a[1]), v2z = fsub(r[2], a[2]);
    const float v1x = fsub(b[0], a[0]), v1y = fsub(b[1], a[1]), v1z = fsub(b[2], a[2]);
    float nx = fsub(fmul(v1y, v2z), fmul(v1z, v2y));
    float ny = fsub(fmul(v1z, v2x), fmul(v1x, v2z));
    float nz = fsub(fmul(v1x, v2y), fmul(v1y, v2x));
    const float l = __fsqrt_rn(fadd(fadd(fmul(nx, nx), fmul(ny, ny)), fmul(nz, nz)));
    n[0] = fdiv(nx, l); n[1] = fdiv(ny, l); n[2] = fdiv(nz, l);
}

// Normals of the 8x4 tile `tile` by one warp; its own vertices and those of its right and lower neighbour tiles are final
// and visible (the caller has seen all three completion signals and fenced).  Loads go to L2 (other SMs wrote the data).
__device__ __forceinline__ void tile_normals(const RayParams &P, uint32_t tile, float *st, int lane) {
    const uint32_t tiles_x = P.width / 8;
    const uint32_t imx = (tile % tiles_x) * 8 + (lane & 7), imy = (tile / tiles_x) * 4 + (lane >> 3);
    const size_t idx = (size_t)imy * P.width + imx;
    float n[3] = { 0.0f, 0.0f, 0.0f };
    if (imy != P.height - 1 && imx != P.width - 1) {
        const float *pa = P.vertices + 3 * idx, *pr = pa + 3, *pb = pa + 3 * (size_t)P.width;
        const float a[3] = { __ldcg(pa), __ldcg(pa + 1), __ldcg(pa + 2) };
        const float r[3] = { __ldcg(pr), __ldcg(pr + 1), __ldcg(pr + 2) };
        const float b[3] = { __ldcg(pb), __ldcg(pb + 1), __ldcg(pb + 2) };
        normal_of(a, r, b, n);
    }
    __syncwarp();
    st[3 * lane + 0] = n[0]; st[3 * lane + 1] = n[1]; st[3 * lane + 2] = n[2];
    __syncwarp();
    if (lane < 24) {
        const uint32_t row = lane / 6, q = lane % 6;
        const float4 v = *reinterpret_cast<const float4 *>(st + 24 * row + 4 * q);
        const size_t first = (size_t)((tile / tiles_x) * 4 + row) * P.width + (tile % tiles_x) * 8;
        *reinterpret_cast<float4 *>(P.normals + 3 * first + 4 * q) = v;
        if (P.mirror_n) *reinterpret_cast<float4 *>(P.mirror_n + 3 * first + 4 * q) = v;
    }
    __syncwarp();
}

// Called by a whole warp once the last vertex of `tile` has been written (the caller observed the 32nd completion).
__device__ __forceinline__ void tile_completed(const RayParams &P, uint32_t tile, float *st, int lane) {
    const uint32_t tiles_x = P.width / 8, tiles_y = P.height / 4;
    const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
    __threadfence();
    if (lane == 0) P.tile_done[tile] = 0u;                    // nobody touches it again in this launch
    if (P.mirror && lane < 24) {
        const uint32_t row = lane / 6, q = lane % 6;
        const size_t first = (size_t)(ty * 4 + row) * P.width + tx * 8;
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(P.vertices + 3 * first + 4 * q));
        *reinterpret_cast<float4 *>(P.mirror + 3 * first + 4 * q) = v;
    }
    if (P.normals) {
        // this tile's vertices are read by the normals of: itself, its left neighbour, its upper neighbour
        bool fired = false;
        uint32_t target = 0;
        if (lane < 3) {
            const int cx = (int)tx - (lane == 1 ? 1 : 0), cy = (int)ty - (lane == 2 ? 1 : 0);
            if (cx >= 0 && cy >= 0) {
                target = (uint32_t)cy * tiles_x + (uint32_t)cx;
                const uint32_t need = 1u + ((uint32_t)cx + 1 < tiles_x ? 1u : 0u) + ((uint32_t)cy + 1 < tiles_y ? 1u : 0u);
                const uint32_t old = atomicAdd(P.tile_deps + target, 1u);
                if (old + 1 == need) { fired = true; P.tile_deps[target] = 0u; }
            }
        }
        unsigned f = __ballot_sync(0xffffffffu, fired);
        while (f) {
            const int src = __ffs(f) - 1;
            f &= f - 1;
            const uint32_t t = __shfl_sync(0xffffffffu, target, src);
            __threadfence();
            tile_normals(P, t, st, lane);
        }
    }
}

#ifndef TSDF_RAY_ROUNDS
#define TSDF_RAY_ROUNDS 4
#endif

#ifndef TSDF_RAY_MINB
#define TSDF_RAY_MINB 6
#endif
// Finishes one ray that the march set aside — queue entry (pixel, first remaining sample) — with a whole warp: lane l marches
// the l-th of 32 equal pieces of the remaining sample range; the ray's first hit is the smallest hit sample over the lanes.
template <bool FASTDIV, bool SKIP, bool SLAB>
__device__ __noinline__ void continue_entry(const RayParams &P, const float *s_t, int2 entry, int lane, float *st, uint32_t &samples) {
    const size_t pix = (size_t)entry.x;
    const uint32_t imx = (uint32_t)entry.x % P.width, imy = (uint32_t)entry.x / P.width;
    const RaySetup R = ray_setup(P, imx, imy);
    // samples the ray still has: k_end is only an estimate of the last one, the final piece runs to the end of the table
    const int k0 = entry.y;
    const int k_end = min(max((int)fminf(R.max_t * __frcp_rn(P.step), 5000.0f) + 1, k0), TSDF_B200_MAX_SAMPLES - 1);
    const int len = k_end - k0 + 1;
    // The remaining samples are cut into kRounds * 32 chunks; in round r lane l marches chunk 32 r + l.  Neighbouring
    // chunks run side by side, so a stretch of expensive samples (a ray skimming a surface) spreads over many lanes
    // instead of landing in one lane's piece; a round whose chunks all start behind the best hit so far is skipped.
    constexpr int kRounds = TSDF_RAY_ROUNDS, kChunks = 32 * kRounds;
    int kh = -1;
    float s_hit = 0.0f;
    RayDebug dbg = { 0, 0, 0, 0, 0 };
    long long key = 0x7fffffffffffffffLL;
    for (int r = 0; r < kRounds; r++) {
        const int c = 32 * r + lane;
        const int first = k0 + (int)((long long)len * c / kChunks);
        const int last = c == kChunks - 1 ? TSDF_B200_MAX_SAMPLES - 1 : k0 + (int)((long long)len * (c + 1) / kChunks) - 1;
        if (first <= last && kh < 0) march_ray<FASTDIV, SKIP, SLAB>(P, s_t, R, first, last, 0x7fffffff, kh, s_hit, samples, dbg);
        key = (kh >= 0) ? (((long long)kh << 32) | (long long)(uint32_t)__float_as_uint(s_hit)) : 0x7fffffffffffffffLL;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, o));
        if (key != 0x7fffffffffffffffLL) break;           // later rounds only hold later samples
    }
    if (lane == 0) {
        if (SLAB) {
            if (P.keys_min) { if (key != 0x7fffffffffffffffLL) atomicMin(P.keys_min + pix, key); }
            else P.keys[pix] = key;
        } else {
            float ip[3] = { CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F };
            int k_hit = -1;
            if (key != 0x7fffffffffffffffLL) {
                k_hit = (int)(key >> 32);
                hit_vertex(P, R, s_t[k_hit], __uint_as_float((uint32_t)(key & 0xffffffffLL)), ip);
            }
            if (P.n_out) {
                for (uint32_t d = 0; d < P.n_out; d++) {
                    float *v = P.out[d] + 3 * pix;
                    v[0] = ip[0]; v[1] = ip[1]; v[2] = ip[2];
                }
            } else {
                P.vertices[3 * pix + 0] = ip[0];
                P.vertices[3 * pix + 1] = ip[1];
                P.vertices[3 * pix + 2] = ip[2];
                if (P.khit && !P.debug_iters) P.khit[pix] = k_hit;
                if (P.mirror && !P.tile_done) { P.mirror[3 * pix + 0] = ip[0]; P.mirror[3 * pix + 1] = ip[1]; P.mirror[3 * pix + 2] = ip[2]; }
            }
        }
    }
    __syncwarp();
    if (!SLAB && P.tile_done) {
        // fused normals: this ray may have been the last one its tile waited for
        const uint32_t tile = (imy / 4) * (P.width / 8) + imx / 8;
        bool completed = false;
        if (lane == 0) { __threadfence(); completed = atomicAdd(P.tile_done + tile, 1u) == 31u; }
        if (__shfl_sync(0xffffffffu, (int)completed, 0)) tile_completed(P, tile, st, lane);
    }
}

// The march without a cap, out of line: for the ray that finds the continuation queue full.
template <bool FASTDIV, bool SKIP, bool SLAB>
__device__ __noinline__ void march_ray_cold(const RayParams &P, const float *s_t, const RaySetup &R, int k_first,
                                            int &kh, float &s_hit, uint32_t &samples, RayDebug &dbg) {
    march_ray<FASTDIV, SKIP, SLAB>(P, s_t, R, k_first, TSDF_B200_MAX_SAMPLES - 1, 0x7fffffff, kh, s_hit, samples, dbg);
}

// One thread per pixel.  SLAB: the volume arrays hold planes [z_base, z_base + nz_local) of a Z-sharded volume,
// only samples whose interpolation cell starts in [z_lo, z_hi) are evaluated, and the result is a key.
template <bool FASTDIV, bool SKIP, bool SLAB>
__global__ void __launch_bounds__(128, TSDF_RAY_MINB)
raycast_kernel(const __grid_constant__ RayParams P) {
    __shared__ float s_t[TSDF_B200_RAY_TABLE_LEN];
    __shared__ __align__(16) float s_tile[4][96];
    for (int i = threadIdx.x; i < TSDF_B200_RAY_TABLE_LEN; i += blockDim.x) s_t[i] = P.table[i];
    __syncthreads();

    // A warp marches one 8x4 pixel tile at a time.  The grid is sized to what is resident on the chip and warps draw
    // tiles from a work counter: a tile costs between ~10 and ~400 loop iterations (tools/ray_iters.py), so with one
    // block per 16x8 tile the kernel ran 2.3 waves at 23% achieved occupancy, and with a static interleaved assignment
    // the slowest warp still took twice the mean.  Without a counter (no occupancy buffer to keep it in) warps take
    // tiles w, w + W, w + 2W, ...
    const int lane = threadIdx.x & 31;
    const uint32_t tiles_x = (P.width + 7) / 8, tiles_y = (P.height + 3) / 4, n_tiles = tiles_x * tiles_y;
    const uint32_t warps_total = gridDim.x * (blockDim.x >> 5);
    const uint32_t n_work = P.tile_stride > 1 ? (n_tiles + P.tile_stride - 1) / P.tile_stride : n_tiles;
    uint32_t samples = 0;
    int max_iters = 0x7fffffff;
    if (P.queue) {
        max_iters = P.max_iters;
        if (P.cap_word) { const unsigned int c = __ldg(P.cap_word); if (c) max_iters = (int)c; }
    }
    auto next_tile = [&](uint32_t previous, bool first) -> uint32_t {
        if (!P.tile_counter) return first ? blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : previous + warps_total;
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(P.tile_counter, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };

    for (uint32_t work = next_tile(0, true); work < n_work; work = next_tile(work, false)) {
    uint32_t tile = work;
    if (P.tile_stride > 1) {
        // one tile of each group of tile_stride; the pick rotates from tile row to tile row so that a rank's tiles do
        // not line up in image columns
        const uint32_t base = tile * P.tile_stride;
        tile = base + (P.tile_first + base / tiles_x) % P.tile_stride;
        if (tile >= n_tiles) {
            if (P.queue && lane == 0) atomicAdd(P.tiles_finished, 1u);
            continue;
        }
    }
    const uint32_t imx = (tile % tiles_x) * 8 + (lane & 7);
    const uint32_t imy = (tile / tiles_x) * 4 + (lane >> 3);
    float ip[3] = { CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F };
    bool queued = false;
    uint32_t queue_slot = 0;
    int queue_resume = 0;
    if (imx < P.width && imy < P.height) {
        const size_t pix = (size_t)imy * P.width + imx;
        const RaySetup R = ray_setup(P, imx, imy);
        int kh = -1, dbg_iters = 0;
        float s_hit = 0.0f;
        RayDebug dbg = { 0, 0, 0, 0, 0 };
#ifdef TSDF_RAY_DEBUG
        unsigned long long dbg_t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
#endif

        if (R.intersects) {
            // a ray that is still marching after max_iters iterations hands the rest of its samples to the continuation
            // queue (continue_entry); with the queue full it finishes here (out-of-line copy of the march: an outer loop
            // around the inlined one made the compiler give up reconverging the warp inside the march, 4x the instructions)
            const int resume = march_ray<FASTDIV, SKIP, SLAB>(P, s_t, R, 0, TSDF_B200_MAX_SAMPLES - 1, max_iters, kh, s_hit, samples, dbg);
            if (resume >= 0) {
                // (the slot is taken now, the entry is written after this tile's own stores — see below)
                const uint32_t slot = atomicAdd(P.queue_count, 1u);
                if (slot < P.queue_cap) { queued = true; queue_slot = slot; queue_resume = resume; }
                else march_ray_cold<FASTDIV, SKIP, SLAB>(P, s_t, R, resume, kh, s_hit, samples, dbg);
            }
            if (kh >= 0 && !SLAB) hit_vertex(P, R, s_t[kh], s_hit, ip);
            dbg_iters = dbg.iters;
        }
#ifdef TSDF_RAY_DEBUG
        {
            unsigned long long dbg_t1;
            __syncwarp(__activemask());
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t1));
            if (P.debug_iters == 2) dbg_iters = dbg.eval;
            if (P.debug_iters == 3) dbg_iters = dbg.l1;
            if (P.debug_iters == 4) dbg_iters = dbg.l2;
            if (P.debug_iters == 5) dbg_iters = (int)(dbg_t1 - dbg_t0);
            if (P.debug_iters == 6) dbg_iters = (int)(dbg_t0 & 0x7fffffffull);
            if (P.debug_iters == 7) dbg_iters = dbg.l3;
        }
#endif
        if (queued) {
            // the continuation writes this pixel
        } else if (SLAB) {
            // key: first hit along the ray wins the min over the ranks; the sample value rides in the low word
            const long long key = (kh >= 0) ? (((long long)kh << 32) | (long long)(uint32_t)__float_as_uint(s_hit)) : 0x7fffffffffffffffLL;
            if (P.keys_min) { if (kh >= 0) atomicMin(P.keys_min + pix, key); }
            else P.keys[pix] = key;
        } else if (P.n_out) {
            for (uint32_t d = 0; d < P.n_out; d++) {
                float *v = P.out[d] + 3 * pix;
                v[0] = ip[0]; v[1] = ip[1]; v[2] = ip[2];
            }
        } else {
            P.vertices[3 * pix + 0] = ip[0];
            P.vertices[3 * pix + 1] = ip[1];
            P.vertices[3 * pix + 2] = ip[2];
            if (P.khit) P.khit[pix] = P.debug_iters ? dbg_iters : kh;
        }
    }
    __syncwarp();
    if (!SLAB && P.tile_done) {
        // fused normals (and mirrors): the tile is complete once its set-aside rays have been finished as well
        const unsigned nq = (unsigned)__popc(__ballot_sync(0xffffffffu, queued));
        bool completed = false;
        if (lane == 0) { __threadfence(); completed = atomicAdd(P.tile_done + tile, 32u - nq) + (32u - nq) == 32u; }
        if (__shfl_sync(0xffffffffu, (int)completed, 0)) tile_completed(P, tile, s_tile[threadIdx.x >> 5], lane);
    } else if (!SLAB && P.mirror) {
        // the tile as 4 rows of 24 floats; pixels handed to the continuation hold NaN here and are rewritten by it
        float *st = s_tile[threadIdx.x >> 5];
        st[3 * lane + 0] = ip[0]; st[3 * lane + 1] = ip[1]; st[3 * lane + 2] = ip[2];
        __syncwarp();
        if (lane < 24) {
            const uint32_t row = lane / 6, q = lane % 6;
            const float4 v = *reinterpret_cast<const float4 *>(st + 24 * row + 4 * q);
            const size_t first = (size_t)((tile / tiles_x) * 4 + row) * P.width + (tile % tiles_x) * 8;
            *reinterpret_cast<float4 *>(P.mirror + 3 * first + 4 * q) = v;
        }
        __syncwarp();
    }
    if (P.queue) {
        // The rays of this tile that are set aside enter the queue only now: whoever finishes one of them may start at once,
        // and its single-pixel store into the mirror has to come after this warp's store of the whole tile (which holds NaN
        // for that pixel).
        if (!SLAB && P.mirror && !P.tile_done) __threadfence_system();
        if (queued) P.queue[queue_slot] = make_int2((int)(imy * P.width + imx), queue_resume);
        __syncwarp();
        if (lane == 0) { __threadfence(); atomicAdd(P.tiles_finished, 1u); }
    }
    }

    // ---- no tile left: this warp turns to the rays that were set aside (round 2; a second kernel used to do this after the
    // march had drained, 50 us during which the chip was all but idle).  It claims the next queue slot and waits for a ray to
    // appear in it, or for the last tile to finish — from then on a slot that is still empty stays empty.  Whether a block of
    // this grid is resident or not does not matter: only warps that hold a tile are waited for.
    if (P.queue) {
        float *st = s_tile[threadIdx.x >> 5];
        while (true) {
            long long raw = -1;
            if (lane == 0) {
                const uint32_t c = atomicAdd(P.queue_head, 1u);
                if (c < P.queue_cap) {
                    const volatile long long *slot = reinterpret_cast<const volatile long long *>(P.queue + c);
                    for (uint32_t polls = 0; polls < (1u << 22); polls++) {       // (bounded: ~1 s)
                        raw = *slot;
                        if ((int)(raw >> 32) >= 0) break;                           // int2 (pixel, sample): the sample is never negative
                        if (*reinterpret_cast<const volatile unsigned int *>(P.tiles_finished) >= n_work) {
                            __threadfence();
                            raw = *slot;
                            break;
                        }
                        __nanosleep(200);
                    }
                }
            }
            raw = __shfl_sync(0xffffffffu, raw, 0);
            const int2 entry = make_int2((int)(uint32_t)(raw & 0xffffffffLL), (int)(raw >> 32));
            if (entry.y < 0) break;
            continue_entry<FASTDIV, SKIP, SLAB>(P, s_t, entry, lane, st, samples);
        }
    }

    if (P.n_samples) {
        for (int o = 16; o > 0; o >>= 1) samples += __shfl_down_sync(0xffffffffu, samples, o);
        if (lane == 0 && samples) atomicAdd(P.n_samples, (unsigned long long)samples);
    }
}

// Keys (after the min-reduction over ranks) -> vertices: the same per-ray set-up and hit formula as the march.
__global__ void __launch_bounds__(128)
resolve_kernel(const __grid_constant__ RayParams P) {
    const uint32_t imx = blockIdx.x * 16 + (threadIdx.x & 15);
    const uint32_t imy = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (imx >= P.width || imy >= P.height) return;
    if (P.tile_stride > 1) {       // image sharding: only the pixels of this rank's tiles (same pick as the march)
        const uint32_t tiles_x = (P.width + 7) / 8, tile = (imy / 4) * tiles_x + imx / 8;
        const uint32_t base = tile / P.tile_stride * P.tile_stride;
        if (tile != base + (P.tile_first + base / tiles_x) % P.tile_stride) return;
    }
    const size_t pix = (size_t)imy * P.width + imx;
    const long long key = P.keys[pix];
    if (P.reset_keys) P.keys[pix] = 0x7fffffffffffffffLL;
    float ip[3] = { CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F };
    int kh = -1;
    if (key != 0x7fffffffffffffffLL) {
        kh = (int)(key >> 32);
        const float s = __uint_as_float((uint32_t)(key & 0xffffffffLL));
        const RaySetup R = ray_setup(P, imx, imy);
        hit_vertex(P, R, __ldg(P.table + kh), s, ip);
    }
    if (P.n_out) {
        for (uint32_t d = 0; d < P.n_out; d++) {
            float *v = P.out[d] + 3 * pix;
            v[0] = ip[0]; v[1] = ip[1]; v[2] = ip[2];
        }
        return;
    }
    P.vertices[3 * pix + 0] = ip[0];
    P.vertices[3 * pix + 1] = ip[1];
    P.vertices[3 * pix + 2] = ip[2];
    if (P.khit) P.khit[pix] = kh;
}

// compute_normals (GPURaycaster.cu:393-427)
__global__ void normals_kernel(uint32_t width, uint32_t height, const float *__restrict__ V, float *__restrict__ N) {
    const uint32_t imx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t imy = blockIdx.y * blockDim.y + threadIdx.y;
    if (imx >= width || imy >= height) return;
    const size_t idx = (size_t)imy * width + imx;
    float nx = 0, ny = 0, nz = 0;
    if (imy != height - 1 && imx != width - 1) {
        const float *a = V + 3 * idx, *r = V + 3 * (idx + 1), *b = V + 3 * (idx + width);
        const float ax = a[0], ay = a[1], az = a[2];
        const float v2x = fsub(r[0], ax), v2y = fsub(r[1], ay), v2z = fsub(r[2], az);
        const float v1x = fsub(b[0], ax), v1y = fsub(b[1], ay), v1z = fsub(b[2], az);
        nx = fsub(fmul(v1y, v2z), fmul(v1z, v2y));
        ny = fsub(fmul(v1z, v2x), fmul(v1x, v2z));
        nz = fsub(fmul(v1x, v2y), fmul(v1y, v2x));
        const float l = __fsqrt_rn(fadd(fadd(fmul(nx, nx), fmul(ny, ny)), fmul(nz, nz)));
        nx = fdiv(nx, l); ny = fdiv(ny, l); nz = fdiv(nz, l);
    }
    N[3 * idx + 0] = nx; N[3 * idx + 1] = ny; N[3 * idx + 2] = nz;
}

// t_0 = 0, t_{k+1} = t_k + step (GPURaycaster.cu:316,324,360): a serial chain of 4415 adds.
__global__ void ray_table_kernel(float trunc, float *table) {
    const float step = (float)((double)trunc * 0.05);
    float t = 0;
    for (int k = 0; k < TSDF_B200_RAY_TABLE_LEN; k++) { table[k] = t; t = fadd(t, step); }
}

// Exhaustive check of fdiv_recip against IEEE division for one divisor, over every numerator
// the raycast can feed it: +-0 and 2^-100 <= |a| <= 2^100.  (Smaller |a| only ever reach the
// floor() of voxel_for_point, where any value in [0,1) gives voxel 0; p - centre is either 0 or
// at least voxel*2^-25; larger |a| are excluded by the host-side magnitude screen.)
__global__ void selftest_div_kernel(float b, float r, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    const uint64_t total = 1ull << 32;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const float a = __uint_as_float((uint32_t)i);
        const uint32_t ex = ((uint32_t)i >> 23) & 0xffu;
        if (((uint32_t)i << 1) != 0u && (ex < 27u || ex > 227u)) continue;
        const float q0 = fdiv(a, b), q1 = fdiv_recip(a, b, r);
        // NaN payloads aside, results must be bit-identical.  (-0 numerator: the sequence returns +0 where
        // IEEE gives -0; the only consumer of a possible -0 numerator is floor() -> voxel 0 either way.)
        if (__float_as_uint(q0) != __float_as_uint(q1) && !(q0 != q0 && q1 != q1) && !(q0 == 0.0f && q1 == 0.0f)) bad++;
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, o);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, bad);
}

}  // namespace tsdf

using namespace tsdf;

extern "C" int tsdf_b200_ray_table(float trunc, float *d_table, void *stream) {
    if (!d_table) return TSDF_B200_EINVAL;
    ray_table_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(trunc, d_table);
    return (int)cudaGetLastError();
}

extern "C" int tsdf_b200_selftest_division(float divisor, unsigned long long *mismatches) {
    if (!mismatches) return TSDF_B200_EINVAL;
    // The proof is a property of the divisor's bit pattern and of the (identical) GPUs: it is run once per divisor and
    // process — 4 ms of device time that every volume of the same voxel size used to pay again at creation.
    static std::mutex cache_lock;
    static std::map<uint32_t, unsigned long long> cache;
    uint32_t key;
    memcpy(&key, &divisor, sizeof(key));
    {
        std::lock_guard<std::mutex> lk(cache_lock);
        auto hit = cache.find(key);
        if (hit != cache.end()) { *mismatches = hit->second; return 0; }
    }
    unsigned long long *d = nullptr;
    TSDF_CUDA_TRY(cudaMalloc(&d, sizeof(*d)));
    cudaMemset(d, 0, sizeof(*d));
    selftest_div_kernel<<<148 * 8, 256>>>(divisor, 1.0f / divisor, d);
    cudaError_t e = cudaMemcpy(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(cache_lock);
        cache[key] = *mismatches;
    }
    return (int)e;
}

// Host-side magnitude screen for fdiv_recip: with every geometric input finite and of sane
// magnitude no numerator leaves the domain selftest_div_kernel covers; otherwise IEEE division.
static bool fastdiv_range_ok(float b) { return b > 1.0e-6f && b < 1.0e6f; }
static bool sane(float x, float bound) { return x == x && fabsf(x) < bound; }

static int fill_params(RayParams &P, const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz, const float voxel[3],
                       const float space_min[3], const float space_max[3], float trunc, const float origin[3],
                       const float rot[9], const float kinv[9], uint32_t width, uint32_t height, const float *d_table,
                       int *fastdiv) {
    if (!voxel || !space_min || !space_max || !origin || !rot || !kinv || !d_table) return TSDF_B200_EINVAL;
    if (nx == 0 || ny == 0 || nz == 0 || width == 0 || height == 0) return TSDF_B200_EINVAL;
    if (nx > 65535 || ny > 65535 || nz > 65535 || width > 65535 || height > 65535) return TSDF_B200_EINVAL;
    if ((uint64_t)nx * ny * nz > 0xffffffffull) return TSDF_B200_EINVAL;   // reference indexes voxels in 32 bits
    P.dist = d_dist; P.nx = nx; P.ny = ny; P.nz = nz;
    for (int i = 0; i < 3; i++) {
        P.vs[i] = voxel[i]; P.rvs[i] = 1.0f / voxel[i];
        P.smin[i] = space_min[i]; P.smax[i] = space_max[i]; P.origin[i] = origin[i];
        if (!fastdiv_range_ok(voxel[i]) || !sane(space_min[i], 1e12f) || !sane(space_max[i], 1e12f) || !sane(origin[i], 1e12f))
            *fastdiv = 0;
    }
    for (int i = 0; i < 9; i++) if (!sane(rot[i], 1e6f) || !sane(kinv[i], 1e6f)) *fastdiv = 0;
    P.trunc = trunc;
    P.step = (float)((double)trunc * 0.05);
    for (int i = 0; i < 9; i++) { P.rot.m[i] = rot[i]; P.kinv.m[i] = kinv[i]; }
    P.width = width; P.height = height; P.table = d_table;
    P.occ = nullptr; P.occ_d = nullptr; P.nbx = P.nby = P.nbz = 0; P.nbz_planes = nz;
    P.occ_lo = trunc * kCellLoFrac; P.occ_hi = trunc * kCellHiFrac;      // level 2 (corners of one cell): the wide band
    P.z_base = 0; P.z_lo = 0; P.z_hi = nz;
    P.cyc_s = 0; P.cyc_g = 0; P.cyc_r = 0;
    P.vertices = nullptr; P.khit = nullptr; P.keys = nullptr; P.keys_min = nullptr; P.reset_keys = 0; P.n_samples = nullptr; P.tile_counter = nullptr;
    P.debug_iters = getenv("TSDF_B200_DEBUG_ITERS") ? atoi(getenv("TSDF_B200_DEBUG_ITERS")) : 0;
    P.tile_first = 0; P.tile_stride = 1; P.n_out = 0; P.mirror = nullptr;
    P.normals = nullptr; P.mirror_n = nullptr; P.tile_done = nullptr; P.tile_deps = nullptr;
    P.queue = nullptr; P.queue_count = nullptr; P.queue_head = nullptr; P.tiles_finished = nullptr; P.cap_word = nullptr; P.queue_cap = 0;
    // (round 2, after the march_step rewrite: 48 / 64 / 80 / 96 / 128 / 160 iterations -> 283 / 259 / 257 / 273 / 302 / 327 us on
    // the bench frames, 470 / 431 / 404 / 398 / 417 / 433 us on frame 500 — profiles/r02u_ray_caps_march_step.txt)
    static const int cap = getenv("TSDF_B200_RAY_CAP") ? atoi(getenv("TSDF_B200_RAY_CAP")) : 80;
    P.max_iters = cap > 0 ? cap : 0x7fffffff;
    // The tight band of low-face bricks covers weights down to -0.51: a sample position may undershoot the low face by 1 %
    // of a voxel.  It undershoots by rounding only (start = origin + near_t * dir - space_min: the rounding of near_t and of the
    // product, ~3e-7 of the largest coordinate involved; 2e-6 is budgeted); with coordinates too large for that bound the old
    // rule (never skip layer 0) applies.
    {
        static const int env_low = getenv("TSDF_B200_LOW_SKIP") ? atoi(getenv("TSDF_B200_LOW_SKIP")) : 1;
        float big = 0.0f, small = 3.0e38f;
        for (int i = 0; i < 3; i++) {
            big = fmaxf(big, fmaxf(fabsf(origin[i]), fmaxf(fabsf(space_min[i]), fabsf(space_max[i]))));
            small = fminf(small, voxel[i]);
        }
        P.low_skip = (env_low && big * 2.0e-6f < 0.01f * small) ? 1 : 0;
    }
    for (int i = 0; i < TSDF_B200_MAX_PEERS; i++) P.out[i] = nullptr;
    return 0;
}

template <bool SLAB>
static int launch_march(RayParams &P, int fastdiv, cudaStream_t s, unsigned int *tile_words = nullptr) {
    // P.normals (a device buffer) asks for the normal map as well; P.mirror / P.mirror_n for copies in pinned host memory.
    float *want_normals = P.normals, *want_mirror_n = P.mirror_n;
    P.normals = nullptr; P.mirror_n = nullptr;
    const uint32_t n_tiles_all = ((P.width + 7) / 8) * ((P.height + 3) / 4);
    const bool aligned = P.width % 8 == 0 && P.height % 4 == 0;
    if (P.occ) {
        // occupancy buffer = [brick flags | distance grid | scratch], each one byte per brick
        const size_t nb = (size_t)P.nbx * P.nby * P.nbz;
        uint8_t *cd = const_cast<uint8_t *>(P.occ) + nb, *tmp = cd + nb;
        if (P.nby > 65535 || P.nbz > 65535) return TSDF_B200_EINVAL;
        // scratch third of the buffer: [work counter | queue slots handed out | queue slots claimed | tiles finished | queue]
        uint8_t *word = (uint8_t *)(((uintptr_t)tmp + 7) & ~(uintptr_t)7);
        const size_t n_words = kMarchWords;
        const bool have_words = word + n_words * 4 <= tmp + nb;
        size_t queue_cap = 0;
        if (have_words && P.max_iters != 0x7fffffff) {
            queue_cap = (size_t)(tmp + nb - (word + n_words * 4)) / sizeof(int2);
            if (queue_cap > 65536) queue_cap = 65536;
            if (queue_cap < 64) queue_cap = 0;
        }
        const int n_queue_words = (int)(2 * queue_cap);
        // iteration cap: fixed (TSDF_B200_RAY_CAP), or 64 / 80 by the number of rays the previous frame set aside (march_reset;
        // thresholds from the orbit, profiles/r02v_adaptive_cap.txt: with 64, more than 6.3 % of the rays; with 80, fewer than 3.7 %)
        MarchReset reset;
        reset.words = have_words ? reinterpret_cast<unsigned int *>(word) : nullptr;
        reset.n_queue_words = n_queue_words;
        // (a Z-slab keeps the fixed cap: most rays cross a slab in a few iterations, so even a few thousand set-aside rays
        // are the longer part — slowest of 8 slabs on frame 10: 298 us with 64, 220 us with 80, 241 us with 96; 263 us before
        // the continuation moved into the march kernel, tools/slab_march_time.py)
        // A slab of at least half the volume behaves like the whole volume (2 GPUs: 2103 frames/s with the adaptive cap against
        // 2005 with the fixed one).
        const bool thick = !SLAB || (P.cyc_g == 0 && 2u * (P.z_hi - P.z_lo) >= P.nz);
        const bool adaptive = !getenv("TSDF_B200_RAY_CAP") && queue_cap && thick;
        // A slab thinner than a quarter of the volume that lies in front of a surface holds almost only skimming rays: cap 96
        // (slowest of 8 slabs on frames 5 / 30 / 54: 212 / 308 / 347 us with 80, 228 / 247 / 247 us with 96; of 4 slabs:
        // 245 / 253 / 253 against 263 / 273 / 278 us — profiles/r02w_slab_caps.txt)
        if (SLAB && !getenv("TSDF_B200_RAY_CAP") && P.cyc_g == 0 && 4u * (P.z_hi - P.z_lo) < P.nz && P.max_iters != 0x7fffffff) P.max_iters = 96;
        reset.cap_lo = adaptive ? 64 : P.max_iters; reset.cap_hi = adaptive ? 80 : P.max_iters;
        const double rays = (double)P.width * P.height / (P.tile_stride > 1 ? P.tile_stride : 1);
        reset.thr_lo = (unsigned int)(0.037 * rays); reset.thr_hi = (unsigned int)(0.063 * rays);
        // completion counters (the caller's, zero between launches) for fused normals
        if (have_words && !SLAB && aligned && tile_words && P.tile_stride <= 1 && want_normals) {
            P.tile_done = tile_words;
            P.tile_deps = tile_words + n_tiles_all;
            P.normals = want_normals; P.mirror_n = want_mirror_n;
        }
        const size_t smem_xy = 2 * (size_t)P.nbx * P.nby, smem_z = (size_t)P.nbx * P.nbz;
        static const bool global_passes = getenv("TSDF_B200_DIST_GLOBAL") != nullptr;      // A/B switches (tuning aids)
        static const bool byte_passes = getenv("TSDF_B200_DIST_BYTES") != nullptr;
        // bit rows: up to 128 bricks along x and y
        const int wpr = (int)((P.nbx + 63) / 64), rpl = (int)((P.nby + 31) / 32);
        void (*bits_kernel)(const uint8_t *, uint8_t *, int, int, int, MarchReset) = nullptr;
        int R = 0;
        if (!global_passes && !byte_passes && wpr <= 2 && rpl <= 4) {
            R = rpl <= 1 ? 1 : (rpl <= 2 ? 2 : 4);
            // (R = 4 with two words per row — 128 x 128 bricks per slice, a 1024^3 volume — needs more registers than a
            // 1024-thread block has: measured 41 us slower than the byte kernels, which keep that case)
            if (wpr == 1)    bits_kernel = R == 1 ? distance_bits_kernel<1, 1> : (R == 2 ? distance_bits_kernel<2, 1> : distance_bits_kernel<4, 1>);
            else if (R <= 2) bits_kernel = R == 1 ? distance_bits_kernel<1, 2> : distance_bits_kernel<2, 2>;
        }
        if (bits_kernel) {
            // two exchange buffers of 32 slices; the bit planes of the output slices reuse them
            const size_t smem_bits = std::max<size_t>(2 * 32, 5 * kDistOut) * 32 * R * wpr * sizeof(unsigned long long);
            // (per device and cheap: set on every call rather than remembered per device and thread)
            if (smem_bits > 48 * 1024)
                TSDF_CUDA_TRY(cudaFuncSetAttribute(bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bits));
            bits_kernel<<<(P.nbz + kDistOut - 1) / kDistOut, 1024, smem_bits, s>>>(
                P.occ, cd, (int)P.nbx, (int)P.nby, (int)P.nbz, reset);
        } else if (smem_xy <= 48 * 1024 && smem_z <= 48 * 1024 && !global_passes) {
            distance_xy_kernel<<<P.nbz, 1024, smem_xy, s>>>(P.occ, cd, (int)P.nbx, (int)P.nby, reset);
            distance_z_kernel<<<P.nby, 1024, smem_z, s>>>(cd, (int)P.nbx, (int)P.nby, (int)P.nbz);
        } else {
            const dim3 g((P.nbx + 255) / 256, P.nby, P.nbz);
            distance_pass_kernel<0, true><<<g, 256, 0, s>>>(P.occ, cd, (int)P.nbx, (int)P.nby, (int)P.nbz);
            distance_pass_kernel<1, false><<<g, 256, 0, s>>>(cd, tmp, (int)P.nbx, (int)P.nby, (int)P.nbz);
            distance_pass_kernel<2, false><<<g, 256, 0, s>>>(tmp, cd, (int)P.nbx, (int)P.nby, (int)P.nbz);
            if (have_words) {      // (no kernel to carry the cap's state here: the fixed cap)
                TSDF_CUDA_TRY(cudaMemsetAsync(word, 0, n_words * 4, s));
                if (n_queue_words) TSDF_CUDA_TRY(cudaMemsetAsync(word + n_words * 4, 0xff, (size_t)n_queue_words * 4, s));
            }
        }
        P.occ_d = cd;
        if (have_words) {
            P.tile_counter = reinterpret_cast<unsigned int *>(word);
            if (queue_cap) {
                P.queue_count = P.tile_counter + 1;
                P.queue_head = P.tile_counter + 2;
                P.tiles_finished = P.tile_counter + 3;
                P.cap_word = P.tile_counter + 4;
                P.queue = reinterpret_cast<int2 *>(word + n_words * 4);
                P.queue_cap = (uint32_t)queue_cap;
            }
        }
    }
    dim3 block(128);
    uint32_t n_tiles = n_tiles_all;
    if (P.tile_stride > 1) n_tiles = (n_tiles + P.tile_stride - 1) / P.tile_stride;      // tiles this rank marches
    auto launch = [&](auto kernel) -> int {
        // resident blocks on this device (queried once per kernel variant)
        static int resident = 0;
        if (resident == 0) {
            int dev = 0, sms = 0, per_sm = 0;
            TSDF_CUDA_TRY(cudaGetDevice(&dev));
            TSDF_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            TSDF_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0));
            resident = sms * (per_sm > 0 ? per_sm : 1);
        }
        const uint32_t blocks = (n_tiles + 3) / 4 < (uint32_t)resident ? (n_tiles + 3) / 4 : (uint32_t)resident;
        kernel<<<blocks, block, 0, s>>>(P);
        return (int)cudaGetLastError();
    };
    int rc;
    if (fastdiv) rc = P.occ ? launch(raycast_kernel<true, true, SLAB>) : launch(raycast_kernel<true, false, SLAB>);
    else         rc = P.occ ? launch(raycast_kernel<false, true, SLAB>) : launch(raycast_kernel<false, false, SLAB>);
    if (rc) return rc;
    if (want_normals && !P.normals) {
        // not fused (no counters given, no occupancy grid, image not tile-aligned): the normals kernel, and a copy
        rc = tsdf_b200_normals(P.width, P.height, P.vertices, want_normals, s);
        if (rc) return rc;
        if (want_mirror_n)
            TSDF_CUDA_TRY(cudaMemcpyAsync(want_mirror_n, want_normals, (size_t)P.width * P.height * 3 * sizeof(float), cudaMemcpyDefault, s));
    }
    return 0;
}

extern "C" int tsdf_b200_raycast_ex(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                    const float voxel[3], const float space_min[3], const float space_max[3],
                                    float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                    uint32_t width, uint32_t height, const float *d_table,
                                    const uint8_t *d_occ, float *d_vertices, int32_t *d_khit,
                                    unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist || !d_vertices) return TSDF_B200_EINVAL;
    RayParams P;
    int rc = fill_params(P, d_dist, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.occ = d_occ;
    const BrickDims nb = brick_dims(nx, ny, nz);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.vertices = d_vertices; P.khit = d_khit; P.n_samples = d_n_samples;
    return launch_march<false>(P, fastdiv, (cudaStream_t)stream);
}

extern "C" int tsdf_b200_raycast_mirrored(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                          const float voxel[3], const float space_min[3], const float space_max[3],
                                          float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                          uint32_t width, uint32_t height, const float *d_table,
                                          const uint8_t *d_occ, float *d_vertices, float *mirror,
                                          unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist || !d_vertices) return TSDF_B200_EINVAL;
    if (mirror && (width % 8 != 0 || height % 4 != 0 || ((uintptr_t)mirror & 15u) != 0)) return TSDF_B200_EINVAL;
    RayParams P;
    int rc = fill_params(P, d_dist, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.occ = d_occ;
    const BrickDims nb = brick_dims(nx, ny, nz);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.vertices = d_vertices; P.n_samples = d_n_samples; P.mirror = mirror;
    return launch_march<false>(P, fastdiv, (cudaStream_t)stream);
}

extern "C" size_t tsdf_b200_raycast_tile_counters(uint32_t width, uint32_t height) {
    return 2 * (size_t)((width + 7) / 8) * ((height + 3) / 4);
}

extern "C" int tsdf_b200_raycast_fused(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                       const float voxel[3], const float space_min[3], const float space_max[3],
                                       float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                       uint32_t width, uint32_t height, const float *d_table,
                                       const uint8_t *d_occ, float *d_vertices, float *d_normals,
                                       float *mirror_vertices, float *mirror_normals, unsigned int *d_tile_counters,
                                       unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist || !d_vertices || !d_normals) return TSDF_B200_EINVAL;
    if ((mirror_vertices || mirror_normals) &&
        (width % 8 != 0 || height % 4 != 0 || ((uintptr_t)mirror_vertices & 15u) != 0 || ((uintptr_t)mirror_normals & 15u) != 0))
        return TSDF_B200_EINVAL;
    RayParams P;
    int rc = fill_params(P, d_dist, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.occ = d_occ;
    const BrickDims nb = brick_dims(nx, ny, nz);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.vertices = d_vertices; P.n_samples = d_n_samples; P.mirror = mirror_vertices;
    P.normals = d_normals; P.mirror_n = mirror_normals;
    return launch_march<false>(P, fastdiv, (cudaStream_t)stream, d_tile_counters);
}

extern "C" int tsdf_b200_raycast_tiles(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                       const float voxel[3], const float space_min[3], const float space_max[3],
                                       float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                       uint32_t width, uint32_t height, const float *d_table,
                                       const uint8_t *d_occ, uint32_t world, uint32_t rank,
                                       uint32_t n_out, float *const *d_vertices_out,
                                       unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist || !d_vertices_out || n_out == 0 || n_out > TSDF_B200_MAX_PEERS || world == 0 || rank >= world)
        return TSDF_B200_EINVAL;
    RayParams P;
    int rc = fill_params(P, d_dist, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.occ = d_occ;
    const BrickDims nb = brick_dims(nx, ny, nz);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.n_samples = d_n_samples;
    P.tile_first = rank; P.tile_stride = world;
    P.n_out = n_out;
    for (uint32_t i = 0; i < n_out; i++) {
        if (!d_vertices_out[i]) return TSDF_B200_EINVAL;
        P.out[i] = d_vertices_out[i];
    }
    return launch_march<false>(P, fastdiv, (cudaStream_t)stream);
}

extern "C" int tsdf_b200_raycast(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                 const float voxel[3], const float space_min[3], const float space_max[3],
                                 float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                 uint32_t width, uint32_t height, const float *d_table,
                                 const uint8_t *d_occ, float *d_vertices, int32_t *d_khit,
                                 unsigned long long *d_n_samples, void *stream) {
    // IEEE division unless a caller (the level-2 volume) has proven the reciprocal form.
    return tsdf_b200_raycast_ex(d_dist, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv,
                                width, height, d_table, d_occ, d_vertices, d_khit, d_n_samples, 0, stream);
}

extern "C" int tsdf_b200_raycast_slab(const float *d_dist_slab, uint32_t nx, uint32_t ny, uint32_t nz,
                                      uint32_t z_base, uint32_t z_planes, uint32_t z_lo, uint32_t z_hi,
                                      const float voxel[3], const float space_min[3], const float space_max[3],
                                      float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                      uint32_t width, uint32_t height, const float *d_table,
                                      const uint8_t *d_occ_slab, long long *d_keys,
                                      unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist_slab || !d_keys) return TSDF_B200_EINVAL;
    if (z_lo < z_base || z_hi > nz || z_lo > z_hi || z_base + z_planes > nz || z_planes == 0) return TSDF_B200_EINVAL;
    // every cell starting in [z_lo, z_hi) needs planes l and min(l+1, nz-1)
    if (z_hi > z_lo && (z_hi < nz ? z_hi : nz - 1) > z_base + z_planes - 1) return TSDF_B200_EINVAL;
    if (d_occ_slab && z_base % TSDF_B200_BRICK != 0) return TSDF_B200_EINVAL;
    RayParams P;
    int rc = fill_params(P, d_dist_slab, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.occ = d_occ_slab;
    const BrickDims nb = brick_dims(nx, ny, z_planes);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.z_base = z_base; P.z_lo = z_lo; P.z_hi = z_hi; P.nbz_planes = z_planes;
    P.keys = d_keys; P.n_samples = d_n_samples;
    return launch_march<true>(P, fastdiv, (cudaStream_t)stream);
}

extern "C" int tsdf_b200_raycast_slab_min(const float *d_dist_slab, uint32_t nx, uint32_t ny, uint32_t nz,
                                          uint32_t z_base, uint32_t z_planes, uint32_t z_lo, uint32_t z_hi,
                                          const float voxel[3], const float space_min[3], const float space_max[3],
                                          float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                          uint32_t width, uint32_t height, const float *d_table,
                                          const uint8_t *d_occ_slab, long long *d_keys_min,
                                          unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist_slab || !d_keys_min) return TSDF_B200_EINVAL;
    if (z_lo < z_base || z_hi > nz || z_lo > z_hi || z_base + z_planes > nz || z_planes == 0) return TSDF_B200_EINVAL;
    if (z_hi > z_lo && (z_hi < nz ? z_hi : nz - 1) > z_base + z_planes - 1) return TSDF_B200_EINVAL;
    if (d_occ_slab && z_base % TSDF_B200_BRICK != 0) return TSDF_B200_EINVAL;
    RayParams P;
    int rc = fill_params(P, d_dist_slab, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.occ = d_occ_slab;
    const BrickDims nb = brick_dims(nx, ny, z_planes);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.z_base = z_base; P.z_lo = z_lo; P.z_hi = z_hi; P.nbz_planes = z_planes;
    P.keys_min = d_keys_min; P.n_samples = d_n_samples;
    return launch_march<true>(P, fastdiv, (cudaStream_t)stream);
}

extern "C" int tsdf_b200_raycast_interleaved(const float *d_dist_local, uint32_t nx, uint32_t ny, uint32_t nz,
                                             uint32_t slab_planes, uint32_t world, uint32_t rank,
                                             const float voxel[3], const float space_min[3], const float space_max[3],
                                             float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                             uint32_t width, uint32_t height, const float *d_table,
                                             const uint8_t *d_occ_global, long long *d_keys,
                                             unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist_local || !d_keys || world == 0 || rank >= world) return TSDF_B200_EINVAL;
    if (slab_planes == 0 || slab_planes % TSDF_B200_BRICK != 0) return TSDF_B200_EINVAL;
    RayParams P;
    int rc = fill_params(P, d_dist_local, nx, ny, nz, voxel, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.occ = d_occ_global;
    const BrickDims nb = brick_dims(nx, ny, nz);
    P.nbx = nb.bx; P.nby = nb.by; P.nbz = nb.bz;
    P.cyc_s = slab_planes; P.cyc_g = world; P.cyc_r = rank;
    P.keys = d_keys; P.n_samples = d_n_samples;
    return launch_march<true>(P, fastdiv, (cudaStream_t)stream);
}

extern "C" int tsdf_b200_raycast_resolve_reset(long long *d_keys, const float space_min[3], const float space_max[3],
                                               float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                               uint32_t width, uint32_t height, const float *d_table,
                                               float *d_vertices, int32_t *d_khit, void *stream) {
    if (!d_keys || !d_vertices) return TSDF_B200_EINVAL;
    RayParams P;
    int fastdiv = 0;
    const float one[3] = { 1.f, 1.f, 1.f };
    int rc = fill_params(P, nullptr, 1, 1, 1, one, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.keys = d_keys; P.reset_keys = 1;
    P.vertices = d_vertices; P.khit = d_khit;
    dim3 block(128);
    dim3 grid((width + 15) / 16, (height + 7) / 8);
    resolve_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P);
    return (int)cudaGetLastError();
}

extern "C" int tsdf_b200_raycast_resolve(const long long *d_keys, const float space_min[3], const float space_max[3],
                                         float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                         uint32_t width, uint32_t height, const float *d_table,
                                         float *d_vertices, int32_t *d_khit, void *stream) {
    if (!d_keys || !d_vertices) return TSDF_B200_EINVAL;
    RayParams P;
    int fastdiv = 0;
    const float one[3] = { 1.f, 1.f, 1.f };
    int rc = fill_params(P, nullptr, 1, 1, 1, one, space_min, space_max, trunc, origin, rot, kinv, width, height, d_table, &fastdiv);
    if (rc) return rc;
    P.keys = const_cast<long long *>(d_keys);
    P.vertices = d_vertices; P.khit = d_khit;
    dim3 block(128);
    dim3 grid((width + 15) / 16, (height + 7) / 8);
    resolve_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P);
    return (int)cudaGetLastError();
}

extern "C" int tsdf_b200_normals(uint32_t width, uint32_t height, const float *d_vertices, float *d_normals, void *stream) {
    if (!d_vertices || !d_normals || width == 0 || height == 0) return TSDF_B200_EINVAL;
    dim3 block(32, 8);
    dim3 grid((width + 31) / 32, (height + 7) / 8);
    normals_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(width, height, d_vertices, d_normals);
    return (int)cudaGetLastError();
}

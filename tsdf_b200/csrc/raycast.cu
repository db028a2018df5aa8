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
    const uint8_t *occ;
    float *vertices;
    int32_t *khit;
    unsigned long long *n_samples;
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

template <bool FASTDIV, bool SKIP>
__global__ void __launch_bounds__(128)
raycast_kernel(const __grid_constant__ RayParams P) {
    __shared__ float s_t[TSDF_B200_RAY_TABLE_LEN];
    for (int i = threadIdx.x; i < TSDF_B200_RAY_TABLE_LEN; i += blockDim.x) s_t[i] = P.table[i];
    __syncthreads();

    // 128 threads = 4 warps, each an 8x4 pixel tile; block tile 16x8.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t imx = blockIdx.x * 16 + (warp & 1) * 8 + (lane & 7);
    const uint32_t imy = blockIdx.y * 8 + (warp >> 1) * 4 + (lane >> 3);
    uint32_t samples = 0;

    if (imx < P.width && imy < P.height) {
        const size_t pix = (size_t)imy * P.width + imx;
        // compute_ray_direction_at_pixel (:24-44): uint16 pixel coords, K^-1 then R, NOT normalised.
        const float fx = (float)(int)(uint16_t)imx, fy = (float)(int)(uint16_t)imy;
        float rc[3], dir[3];
#pragma unroll
        for (int r = 1; r <= 3; r++)
            rc[r - 1] = fadd(fadd(fmul(fx, T33(P.kinv, r, 1)), fmul(fy, T33(P.kinv, r, 2))), T33(P.kinv, r, 3));
#pragma unroll
        for (int r = 1; r <= 3; r++)
            dir[r - 1] = fadd(fadd(fmul(T33(P.rot, r, 1), rc[0]), fmul(T33(P.rot, r, 2), rc[1])), fmul(T33(P.rot, r, 3), rc[2]));

        float near_t, far_t;
        const bool intersects = near_far(P.origin, dir, P.smin, P.smax, near_t, far_t);
        float ip[3] = { CUDART_NAN_F, CUDART_NAN_F, CUDART_NAN_F };
        int kh = -1;

        if (intersects) {
            float start[3];
#pragma unroll
            for (int a = 0; a < 3; a++) start[a] = fsub(fadd(fmul(dir[a], near_t), P.origin[a]), P.smin[a]);   // :306
            const float max_t = fsub(far_t, near_t);                                                           // :317
            const float step = P.step;
            const float mx[3] = { fmul((float)P.nx, P.vs[0]), fmul((float)P.ny, P.vs[1]), fmul((float)P.nz, P.vs[2]) };
            const float hi_adj[3] = { fsub(mx[0], fdiv(P.vs[0], 10.0f)), fsub(mx[1], fdiv(P.vs[1], 10.0f)), fsub(mx[2], fdiv(P.vs[2], 10.0f)) };
            const BrickDims nb = brick_dims(P.nx, P.ny, P.nz);
            float inv_dir[3];
            if (SKIP) {
#pragma unroll
                for (int a = 0; a < 3; a++) inv_dir[a] = __frcp_rn(dir[a]);    // approximate use only
            }
            const float inv_step = __frcp_rn(step);

            int clx = -1, cly = -1, clz = -1;     // corner cache key
            float c000 = 0, c001 = 0, c010 = 0, c011 = 0, c100 = 0, c101 = 0, c110 = 0, c111 = 0;

            int k = 0;
            while (true) {
                // Samples k = 0..4401 exist (:369); sample k>0 exists only if t_k < max_t (:360-365).
                if (k > TSDF_B200_MAX_SAMPLES - 1) break;
                const float t = s_t[k];
                if (k > 0 && t >= max_t) break;

                float p[3];
#pragma unroll
                for (int a = 0; a < 3; a++) p[a] = fadd(fmul(dir[a], t), start[a]);        // :326

                // trilinearly_interpolate (:53-124)
                int vox[3];
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    float adj = p[a];
                    if (p[a] >= mx[a]) adj = hi_adj[a];
                    if (p[a] < 0.0f) adj = 0.0f;
                    vox[a] = f2i(floorf(div_vs<FASTDIV>(adj, P.vs[a], P.rvs[a])));
                }
                const bool oob = vox[0] < 0 || vox[1] < 0 || vox[2] < 0 ||
                                 (uint32_t)vox[0] >= P.nx || (uint32_t)vox[1] >= P.ny || (uint32_t)vox[2] >= P.nz;

                if (SKIP && !oob) {
                    const int bx = vox[0] / TSDF_B200_BRICK, by = vox[1] / TSDF_B200_BRICK, bz = vox[2] / TSDF_B200_BRICK;
                    // Low-edge half voxel extrapolates (:87-99): never skip a sample in voxel layer 0.
                    if (vox[0] >= 1 && vox[1] >= 1 && vox[2] >= 1 &&
                        __ldg(P.occ + ((size_t)bz * nb.by + by) * nb.bx + bx) == 0) {
                        // This sample is > 0 for sure.  How many of the following samples stay inside the
                        // brick (conservatively)?  Exit distance along the ray, minus one step of margin.
                        const int b[3] = { bx, by, bz };
                        float t_exit = CUDART_INF_F;
#pragma unroll
                        for (int a = 0; a < 3; a++) {
                            // Brick faces pulled in by 1% of a voxel: >100x the rounding error of p
                            // and of floor(p/voxel) for grids up to 65535 voxels per side.
                            if (dir[a] > 0.0f) {
                                float bound = ((float)((b[a] + 1) * TSDF_B200_BRICK) - 0.01f) * P.vs[a];
                                t_exit = fminf(t_exit, (bound - p[a]) * inv_dir[a]);
                            } else if (dir[a] < 0.0f) {
                                float bound = ((b[a] == 0) ? 1.01f : (float)(b[a] * TSDF_B200_BRICK) + 0.01f) * P.vs[a];
                                t_exit = fminf(t_exit, (bound - p[a]) * inv_dir[a]);
                            }
                        }
                        int j = 0;
                        if (t_exit > 2.0f * step && t_exit < 1.0e9f) {
                            j = (int)(t_exit * inv_step * 0.999f) - 2;
                            if (j < 0) j = 0;
                            if (k + j > TSDF_B200_MAX_SAMPLES) j = TSDF_B200_MAX_SAMPLES - k;
                            const float t_lim = t + t_exit - step;
                            while (j > 0 && !(s_t[k + j] <= t_lim)) j--;
                        }
                        k += 1 + j;
                        continue;
                    }
                }

                float s;
                if (oob) {
                    s = CUDART_NAN_F;                                                          // :77-80
                } else {
                    int low[3];
                    float uvw[3];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        const float ctr = fadd(fmul(fadd((float)vox[a], 0.5f), P.vs[a]), 0.0f);      // TSDF_utilities.cu:10-17
                        int l = (p[a] < ctr) ? vox[a] - 1 : vox[a];                               // :87-89
                        l = max(l, 0);                                                            // :92-94
                        const float lc = fadd(fmul(fadd((float)l, 0.5f), P.vs[a]), 0.0f);
                        uvw[a] = div_vs<FASTDIV>(fsub(p[a], lc), P.vs[a], P.rvs[a]);             // :98-102
                        low[a] = l;
                    }
                    if (low[0] != clx || low[1] != cly || low[2] != clz) {
                        clx = low[0]; cly = low[1]; clz = low[2];
                        // tsdf_value_at (TSDF_utilities.cu:29-37): upper clamp, 32-bit index arithmetic
                        const uint32_t x0 = min((uint32_t)clx, P.nx - 1), x1 = min((uint32_t)clx + 1, P.nx - 1);
                        const uint32_t y0 = P.nx * min((uint32_t)cly, P.ny - 1), y1 = P.nx * min((uint32_t)cly + 1, P.ny - 1);
                        const uint32_t z0 = P.nx * P.ny * min((uint32_t)clz, P.nz - 1), z1 = P.nx * P.ny * min((uint32_t)clz + 1, P.nz - 1);
                        c000 = __ldg(P.dist + (size_t)(z0 + y0 + x0));
                        c001 = __ldg(P.dist + (size_t)(z1 + y0 + x0));
                        c010 = __ldg(P.dist + (size_t)(z0 + y1 + x0));
                        c011 = __ldg(P.dist + (size_t)(z1 + y1 + x0));
                        c100 = __ldg(P.dist + (size_t)(z0 + y0 + x1));
                        c101 = __ldg(P.dist + (size_t)(z1 + y0 + x1));
                        c110 = __ldg(P.dist + (size_t)(z0 + y1 + x1));
                        c111 = __ldg(P.dist + (size_t)(z1 + y1 + x1));
                    }
                    const float u = uvw[0], v = uvw[1], w = uvw[2];
                    const float u1 = fsub(1.0f, u), v1 = fsub(1.0f, v), w1 = fsub(1.0f, w);
                    s = fmul(fmul(fmul(c000, u1), v1), w1);                                      // :114-121
                    s = fadd(s, fmul(fmul(fmul(c001, u1), v1), w));
                    s = fadd(s, fmul(fmul(fmul(c010, u1), v), w1));
                    s = fadd(s, fmul(fmul(fmul(c011, u1), v), w));
                    s = fadd(s, fmul(fmul(fmul(c100, u), v1), w1));
                    s = fadd(s, fmul(fmul(fmul(c101, u), v1), w));
                    s = fadd(s, fmul(fmul(fmul(c110, u), v), w1));
                    s = fadd(s, fmul(fmul(fmul(c111, u), v), w));
                }
                samples++;

                if (s <= 0) {
                    float th = t;
                    if (s < 0) {
                        th = fsub(th, step);                                                     // :338
                        th = fadd(th, fmul(fdiv(P.trunc, fsub(P.trunc, s)), step));              // :341 (previous_tsdf == trunc)
                    }
#pragma unroll
                    for (int a = 0; a < 3; a++) ip[a] = fadd(fadd(fmul(dir[a], th), start[a]), P.smin[a]);  // :345-348
                    kh = k;
                    break;
                }
                k++;
            }
        }
        P.vertices[3 * pix + 0] = ip[0];
        P.vertices[3 * pix + 1] = ip[1];
        P.vertices[3 * pix + 2] = ip[2];
        if (P.khit) P.khit[pix] = kh;
    }

    if (P.n_samples) {
        for (int o = 16; o > 0; o >>= 1) samples += __shfl_down_sync(0xffffffffu, samples, o);
        if (lane == 0 && samples) atomicAdd(P.n_samples, (unsigned long long)samples);
    }
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
    unsigned long long *d = nullptr;
    TSDF_CUDA_TRY(cudaMalloc(&d, sizeof(*d)));
    cudaMemset(d, 0, sizeof(*d));
    selftest_div_kernel<<<148 * 8, 256>>>(divisor, 1.0f / divisor, d);
    cudaError_t e = cudaMemcpy(mismatches, d, sizeof(*d), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return (int)e;
}

// Host-side magnitude screen for fdiv_recip: with every geometric input finite and of sane
// magnitude no numerator leaves the domain selftest_div_kernel covers; otherwise IEEE division.
static bool fastdiv_range_ok(float b) { return b > 1.0e-6f && b < 1.0e6f; }
static bool sane(float x, float bound) { return x == x && fabsf(x) < bound; }

extern "C" int tsdf_b200_raycast_ex(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                    const float voxel[3], const float space_min[3], const float space_max[3],
                                    float trunc, const float origin[3], const float rot[9], const float kinv[9],
                                    uint32_t width, uint32_t height, const float *d_table,
                                    const uint8_t *d_occ, float *d_vertices, int32_t *d_khit,
                                    unsigned long long *d_n_samples, int fastdiv, void *stream) {
    if (!d_dist || !voxel || !space_min || !space_max || !origin || !rot || !kinv || !d_table || !d_vertices)
        return TSDF_B200_EINVAL;
    if (nx == 0 || ny == 0 || nz == 0 || width == 0 || height == 0) return TSDF_B200_EINVAL;
    if (nx > 65535 || ny > 65535 || nz > 65535 || width > 65535 || height > 65535) return TSDF_B200_EINVAL;
    if ((uint64_t)nx * ny * nz > 0xffffffffull) return TSDF_B200_EINVAL;   // reference indexes voxels in 32 bits

    RayParams P;
    P.dist = d_dist; P.nx = nx; P.ny = ny; P.nz = nz;
    for (int i = 0; i < 3; i++) {
        P.vs[i] = voxel[i]; P.rvs[i] = 1.0f / voxel[i];
        P.smin[i] = space_min[i]; P.smax[i] = space_max[i]; P.origin[i] = origin[i];
        if (!fastdiv_range_ok(voxel[i]) || !sane(space_min[i], 1e12f) || !sane(space_max[i], 1e12f) || !sane(origin[i], 1e12f))
            fastdiv = 0;
    }
    for (int i = 0; i < 9; i++) if (!sane(rot[i], 1e6f) || !sane(kinv[i], 1e6f)) fastdiv = 0;
    P.trunc = trunc;
    P.step = (float)((double)trunc * 0.05);
    for (int i = 0; i < 9; i++) { P.rot.m[i] = rot[i]; P.kinv.m[i] = kinv[i]; }
    P.width = width; P.height = height; P.table = d_table; P.occ = d_occ;
    P.vertices = d_vertices; P.khit = d_khit; P.n_samples = d_n_samples;

    dim3 block(128);
    dim3 grid((width + 15) / 16, (height + 7) / 8);
    cudaStream_t s = (cudaStream_t)stream;
    if (fastdiv) {
        if (d_occ) raycast_kernel<true, true><<<grid, block, 0, s>>>(P);
        else       raycast_kernel<true, false><<<grid, block, 0, s>>>(P);
    } else {
        if (d_occ) raycast_kernel<false, true><<<grid, block, 0, s>>>(P);
        else       raycast_kernel<false, false><<<grid, block, 0, s>>>(P);
    }
    return (int)cudaGetLastError();
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

extern "C" int tsdf_b200_normals(uint32_t width, uint32_t height, const float *d_vertices, float *d_normals, void *stream) {
    if (!d_vertices || !d_normals || width == 0 || height == 0) return TSDF_B200_EINVAL;
    dim3 block(32, 8);
    dim3 grid((width + 31) / 32, (height + 7) / 8);
    normals_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(width, height, d_vertices, d_normals);
    return (int)cudaGetLastError();
}

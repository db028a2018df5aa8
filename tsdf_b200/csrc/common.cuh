// common.cuh — shared device helpers for the sm_100a TSDF kernels.
//
// Parity rule: every fp32 operation of the reference hot path is
// evaluated in the reference's order with NO fused multiply-add.  The reference builds
// with -G (kinfu.make:64), which never contracts; here the arithmetic that must match is
// written with the explicit round-to-nearest intrinsics below so that it stays exact
// whatever -fmad says, and FMAs appear only where they are proven result-neutral.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/tsdf_b200.h"

namespace tsdf {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// Column-major matrices: element (row r, column c), 1-based like the reference's mRC names
// (src/include/cuda_utilities.hpp:12-23).
struct M44 { float m[16]; };
struct M33 { float m[9]; };
#define T44(M, r, c) ((M).m[((c) - 1) * 4 + ((r) - 1)])
#define T33(M, r, c) ((M).m[((c) - 1) * 3 + ((r) - 1)])

// (int) of a float as the reference's device code performs it: cvt.rzi.s32.f32
// (NaN -> 0, saturating).
__device__ __forceinline__ int f2i(float f) { return __float2int_rz(f); }

// a / b through the correctly rounded reciprocal r = RN(1/b): q = RN(a*r), then one
// Markstein correction q' = RN(q + (a - b*q) * r) with the residual exact in an FMA.
// q' == RN(a/b) whenever nothing under/overflows; tsdf_b200_selftest_division() proves it
// exhaustively for a given b, and the kernels are only instantiated with FASTDIV when it
// passed for all three voxel sizes.
__device__ __forceinline__ float fdiv_recip(float a, float b, float r) {
    float q = __fmul_rn(a, r);
    float e = __fmaf_rn(-b, q, a);
    return __fmaf_rn(e, r, q);
}

// Positive bands.  A brick is FLAGGED as soon as a voxel of it or of its 1-voxel apron leaves [0.8, 1.0001] * trunc.  Inside
// an unflagged brick every trilinear sample is positive, even in the first voxel layer of the volume where the reference
// EXTRAPOLATES (u in [-0.5, 0): weights 1 - u in (1, 1.5] and u < 0): with every corner in [a, b], split the weights (they sum
// to 1) into the positive ones (sum W+) and the negative ones (sum -W-), W+ - W- = 1 and W- <= 3.7 for u, v, w >= -0.51; then
// sample >= a W+ - b W- = a - W- (b - a) >= a - 3.7 (b - a) = 0.059 trunc for a = 0.8 trunc, b = 1.0001 trunc — four orders of
// magnitude above the rounding error of the evaluation.  Free space a frame has seen (every update was +trunc) and untouched
// voxels (trunc) are in the band; the ramp in front of a surface and everything behind it are not.  (Round 1 flagged against
// [1e-3, 1e3] * trunc, which proves positivity only for weights in [0, 1]: every ray that entered the volume through a low
// face had to evaluate the ~11 samples of its first voxel.)  Inside a flagged brick the march works cell by cell, and a
// cell whose eight corners are in the WIDE band [1e-3, 1e3] * trunc is positive for weights in [0, 1] (level 2).
static constexpr float kOccLoFrac = 0.8f;
static constexpr float kOccHiFrac = 1.0001f;
static constexpr float kCellLoFrac = 1.0e-3f;
static constexpr float kCellHiFrac = 1.0e3f;

struct BrickDims { uint32_t bx, by, bz; };
__host__ __device__ inline BrickDims brick_dims(uint32_t nx, uint32_t ny, uint32_t nz) {
    return BrickDims{ (nx + TSDF_B200_BRICK - 1) / TSDF_B200_BRICK,
                      (ny + TSDF_B200_BRICK - 1) / TSDF_B200_BRICK,
                      (nz + TSDF_B200_BRICK - 1) / TSDF_B200_BRICK };
}

// Mark every brick whose 1-voxel apron [8b-1, 8b+8]^3 contains voxel (x,y,z).
__device__ __forceinline__ void occ_mark(uint8_t *occ, BrickDims nb, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t B = TSDF_B200_BRICK;
    uint32_t b0[3] = { x / B, y / B, z / B };
    uint32_t lo[3], hi[3];
    const uint32_t v[3] = { x, y, z };
    const uint32_t n[3] = { nb.bx, nb.by, nb.bz };
#pragma unroll
    for (int a = 0; a < 3; a++) {
        uint32_t l = v[a] % B;
        lo[a] = (l == 0 && b0[a] > 0) ? b0[a] - 1 : b0[a];
        hi[a] = (l == B - 1 && b0[a] + 1 < n[a]) ? b0[a] + 1 : b0[a];
    }
    for (uint32_t k = lo[2]; k <= hi[2]; k++)
        for (uint32_t j = lo[1]; j <= hi[1]; j++)
            for (uint32_t i = lo[0]; i <= hi[0]; i++)
                occ[((size_t)k * nb.by + j) * nb.bx + i] = 1;
}

}  // namespace tsdf

#define TSDF_CUDA_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

// misc.cu — clear / deformation grid / occupancy kernels and small host helpers.
#include "common.cuh"
#include <math.h>

namespace tsdf {

// set_memory_to_value x2 of TSDFVolume::clear (reference src/TSDF/TSDFVolume.cu:797-832) as
// one streaming pass with 128-bit stores: weight <- 0, dist <- trunc.
__global__ void __launch_bounds__(256)
clear_kernel(float *__restrict__ dist, float *__restrict__ weight, size_t n, float trunc) {
    const size_t n4 = n / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 d4 = make_float4(trunc, trunc, trunc, trunc);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        reinterpret_cast<float4 *>(dist)[i] = d4;
        reinterpret_cast<float4 *>(weight)[i] = z4;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        dist[i] = trunc;
        weight[i] = 0.f;
    }
}

// initialise_deformation (TSDFVolume.cu:768-794), lanes along X instead of a serial X loop.
__global__ void __launch_bounds__(256)
init_deformation_kernel(float *__restrict__ deform, uint32_t nx, uint32_t ny, uint32_t nz,
                        float vx, float vy, float vz, float ox, float oy, float oz) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y, z = blockIdx.z;
    if (x >= nx) return;
    const size_t idx = ((size_t)nx * ny) * z + (size_t)nx * y + x;
    float *n = deform + 6 * idx;
    n[0] = fadd(fmul(fadd((float)(int)x, 0.5f), vx), ox);
    n[1] = fadd(fmul(fadd((float)(int)y, 0.5f), vy), oy);
    n[2] = fadd(fmul(fadd((float)(int)z, 0.5f), vz), oz);
    n[3] = 0.f; n[4] = 0.f; n[5] = 0.f;
}

__global__ void __launch_bounds__(256)
occupancy_rebuild_kernel(const float *__restrict__ dist, uint32_t nx, uint32_t ny, uint32_t nz,
                         float trunc, uint8_t *occ) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t y = blockIdx.y, z = blockIdx.z;
    if (x >= nx) return;
    const float d = dist[((size_t)nx * ny) * z + (size_t)nx * y + x];
    if (!(d >= trunc * kOccLoFrac && d <= trunc * kOccHiFrac)) occ_mark(occ, brick_dims(nx, ny, nz), x, y, z);
}

}  // namespace tsdf

using namespace tsdf;

extern "C" const char *tsdf_b200_version(void) { return "tsdf_b200 0.1 (sm_100a)"; }

extern "C" const char *tsdf_b200_strerror(int code) {
    switch (code) {
        case 0: return "ok";
        case TSDF_B200_EINVAL: return "invalid argument";
        case TSDF_B200_ENOMEM: return "out of host memory";
        case TSDF_B200_EIO: return "file i/o error";
        case TSDF_B200_ESTATE: return "invalid state";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

extern "C" int tsdf_b200_volume_params(uint32_t nx, uint32_t ny, uint32_t nz, const float physical[3],
                                       float voxel_out[3], float *trunc_out) {
    if (!physical || !voxel_out || !trunc_out || nx == 0 || ny == 0 || nz == 0) return TSDF_B200_EINVAL;
    // f3_div_elem(float3, dim3) and 1.1f * f3_norm (cuda_utilities.hpp:78-81, 99-102; TSDFVolume.cu:690-693).
    // volatile keeps the host compiler from contracting or reassociating.
    volatile float vx = physical[0] / (float)nx, vy = physical[1] / (float)ny, vz = physical[2] / (float)nz;
    volatile float xx = vx * vx, yy = vy * vy, zz = vz * vz;
    volatile float s = xx + yy;
    s = s + zz;
    voxel_out[0] = vx; voxel_out[1] = vy; voxel_out[2] = vz;
    *trunc_out = 1.1f * sqrtf(s);
    return 0;
}

extern "C" size_t tsdf_b200_occupancy_bytes(uint32_t nx, uint32_t ny, uint32_t nz) {
    // [brick flags, maintained by integrate / rebuild | brick distance grid | scratch]: the last two are
    // rewritten by every raycast
    BrickDims nb = brick_dims(nx, ny, nz);
    return 3 * (size_t)nb.bx * nb.by * nb.bz;
}

extern "C" int tsdf_b200_clear(float *d_dist, float *d_weight, uint32_t nx, uint32_t ny, uint32_t nz,
                               float trunc, uint8_t *d_occ, void *stream) {
    if (!d_dist || !d_weight || nx == 0 || ny == 0 || nz == 0) return TSDF_B200_EINVAL;
    if ((((uintptr_t)d_dist | (uintptr_t)d_weight) & 15) != 0) return TSDF_B200_EINVAL;
    const size_t n = (size_t)nx * ny * nz;
    cudaStream_t s = (cudaStream_t)stream;
    clear_kernel<<<148 * 8, 256, 0, s>>>(d_dist, d_weight, n, trunc);
    TSDF_CUDA_TRY(cudaGetLastError());
    if (d_occ) TSDF_CUDA_TRY(cudaMemsetAsync(d_occ, 0, tsdf_b200_occupancy_bytes(nx, ny, nz), s));
    return 0;
}

extern "C" int tsdf_b200_init_deformation(float *d_deform, uint32_t nx, uint32_t ny, uint32_t nz,
                                          const float voxel[3], const float grid_offset[3], void *stream) {
    if (!d_deform || !voxel || !grid_offset || nx == 0 || ny == 0 || nz == 0) return TSDF_B200_EINVAL;
    if (ny > 65535 || nz > 65535) return TSDF_B200_EINVAL;
    dim3 block(256);
    dim3 grid((nx + 255) / 256, ny, nz);
    init_deformation_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(d_deform, nx, ny, nz, voxel[0], voxel[1], voxel[2],
                                                                       grid_offset[0], grid_offset[1], grid_offset[2]);
    return (int)cudaGetLastError();
}

extern "C" int tsdf_b200_occupancy_rebuild(const float *d_dist, uint32_t nx, uint32_t ny, uint32_t nz,
                                           float trunc, uint8_t *d_occ, void *stream) {
    if (!d_dist || !d_occ || nx == 0 || ny == 0 || nz == 0) return TSDF_B200_EINVAL;
    if (ny > 65535 || nz > 65535) return TSDF_B200_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    TSDF_CUDA_TRY(cudaMemsetAsync(d_occ, 0, tsdf_b200_occupancy_bytes(nx, ny, nz), s));
    dim3 block(256);
    dim3 grid((nx + 255) / 256, ny, nz);
    occupancy_rebuild_kernel<<<grid, block, 0, s>>>(d_dist, nx, ny, nz, trunc, d_occ);
    return (int)cudaGetLastError();
}

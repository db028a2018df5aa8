// membench.cu — how fast can a read-modify-write stream over dist+weight go on this GPU, as a function of the ORDER in
// which thread blocks walk the volume?  (Tuning aid for the integrate kernel; trivial arithmetic, same bytes.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu && tools/membench
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__global__ void linear(float4 *d, float4 *w, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 a = d[i], b = w[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; b.x += 1.f; b.y += 1.f; b.z += 1.f; b.w += 1.f;
        d[i] = a; w[i] = b;
    }
}
// block = 128 threads along x (one 2 KB row when nx = 512), grid (1, ny, nz / zpt); each thread walks zpt planes
__global__ void zwalk(float4 *d, float4 *w, int nx4, int ny, int zpt) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= nx4) return;
    size_t i = ((size_t)blockIdx.z * zpt * ny + y) * nx4 + x;
    for (int z = 0; z < zpt; z++, i += (size_t)ny * nx4) {
        float4 a = d[i], b = w[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; b.x += 1.f; b.y += 1.f; b.z += 1.f; b.w += 1.f;
        d[i] = a; w[i] = b;
    }
}
// block = 128 threads along x, grid (1, ny / rows, nz); each thread walks `rows` consecutive rows of one plane
__global__ void ywalk(float4 *d, float4 *w, int nx4, int ny, int rows) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx4) return;
    size_t i = ((size_t)blockIdx.z * ny + (size_t)blockIdx.y * rows) * nx4 + x;
    for (int r = 0; r < rows; r++, i += nx4) {
        float4 a = d[i], b = w[i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; b.x += 1.f; b.y += 1.f; b.z += 1.f; b.w += 1.f;
        d[i] = a; w[i] = b;
    }
}
// z-walk, but a block owns `ry` consecutive rows of every plane (ry * 2 KB contiguous per plane visit)
__global__ void zwalk_rows(float4 *d, float4 *w, int nx4, int ny, int zpt, int ry) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nx4) return;
    for (int z = 0; z < zpt; z++) {
        size_t i = (((size_t)blockIdx.z * zpt + z) * ny + (size_t)blockIdx.y * ry) * nx4 + x;
        for (int r = 0; r < ry; r++, i += nx4) {
            float4 a = d[i], b = w[i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; b.x += 1.f; b.y += 1.f; b.z += 1.f; b.w += 1.f;
            d[i] = a; w[i] = b;
        }
    }
}

template <typename F> static float time_ms(F launch, int reps = 10) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int n = 512, nx4 = n / 4;
    const size_t nv = (size_t)n * n * n;
    float4 *d, *w;
    cudaMalloc(&d, nv * 4); cudaMalloc(&w, nv * 4);
    cudaMemset(d, 0, nv * 4); cudaMemset(w, 0, nv * 4);
    const double gb = 16.0 * nv / 1e9;
    auto report = [&](const char *name, float ms) { printf("%-28s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, gb / (ms * 1e-3)); };
    report("linear 148x16 blocks", time_ms([&] { linear<<<148 * 16, 256>>>(d, w, nv / 4); }));
    report("linear 148x8 blocks x512", time_ms([&] { linear<<<148 * 8, 512>>>(d, w, nv / 4); }));
    for (int zpt : { 8, 16, 32, 64 }) {
        char name[64]; snprintf(name, sizeof name, "zwalk zpt=%d", zpt);
        report(name, time_ms([&] { zwalk<<<dim3(1, n, n / zpt), 128>>>(d, w, nx4, n, zpt); }));
    }
    for (int rows : { 1, 4, 8, 16 }) {
        char name[64]; snprintf(name, sizeof name, "ywalk rows=%d", rows);
        report(name, time_ms([&] { ywalk<<<dim3(1, n / rows, n), 128>>>(d, w, nx4, n, rows); }));
    }
    for (int ry : { 2, 4, 8 }) for (int zpt : { 8, 16 }) {
        char name[64]; snprintf(name, sizeof name, "zwalk_rows ry=%d zpt=%d", ry, zpt);
        report(name, time_ms([&] { zwalk_rows<<<dim3(1, n / ry, n / zpt), 128>>>(d, w, nx4, n, zpt, ry); }));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

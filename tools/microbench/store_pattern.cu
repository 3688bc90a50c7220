// Dev microbenchmark (GPU box): what a WRITE-ONLY kernel reaches on planes of 257 x 257 floats depending on how its
// warps' stores fall on the 32-byte sectors.  nvcc -arch=sm_100a -O3 -o store_pattern store_pattern.cu && ./store_pattern
//   A  thread = column, rows in sequence (the layout of the fused pde_du / up-sampling kernels): a warp's 128-byte
//      store starts on an arbitrary 4-byte phase -> five sectors, two of them partial
//   B  the same bytes, warps write 128-byte ALIGNED lines of the linear plane (lane = consecutive float)
//   C  aligned float4 per thread
#include <cstdio>
#include <cuda_runtime.h>
constexpr int F = 257, PL = F * F;
__global__ void __launch_bounds__(288) kA(float* out, int cs) {
    const int j = threadIdx.x;
    if (j >= F) return;
    float* o = out + (size_t)blockIdx.x * PL + j;
    for (int i = 0; i < F; ++i, o += F) { const float v = (float)(i + j); if (cs) __stcs(o, v); else *o = v; }
}
// A with only the 256 columns of the eight full warps (no lone 4-byte store per row), and with a row pitch of 256
// floats (every warp store a whole aligned line): separates the access ORDER from the ALIGNMENT
__global__ void __launch_bounds__(256) kA2(float* out, int pitch) {
    const int j = threadIdx.x;
    float* o = out + (size_t)blockIdx.x * PL + j;
    for (int i = 0; i < 256; ++i, o += pitch) *o = (float)(i + j);
}
__global__ void __launch_bounds__(288) kB(float* out, size_t total) {
    // plane per CTA, lines of 32 floats aligned to 128 bytes of the whole buffer
    const size_t lo = (size_t)blockIdx.x * PL, hi = lo + PL;
    const size_t l0 = lo & ~(size_t)31;
    for (size_t e = l0 + threadIdx.x; e < hi; e += blockDim.x) {
        if (e >= lo) { const int r = (int)(e - lo); out[e] = (float)(r / F + r % F); }
    }
}
__global__ void __launch_bounds__(256) kC(float4* out, size_t n4) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (size_t)gridDim.x * blockDim.x) {
        const float v = (float)(k & 1023);
        out[k] = make_float4(v, v + 1.f, v + 2.f, v + 3.f);
    }
}
template <class Fn> static float timeit(Fn fn) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) fn();
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) fn();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20;
}
int main() {
    const int planes = 4320;                               // 1.14 GB: the fused pde_du output at B = 2160
    const size_t total = (size_t)planes * PL;
    float* d; cudaMalloc(&d, total * 4 + 256);
    const double gb = total * 4 / 1e9;
    float t;
    t = timeit([&] { kA<<<planes, 288>>>(d, 1); }); printf("A  thread = column, streaming stores : %.4f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { kA<<<planes, 288>>>(d, 0); }); printf("A  thread = column, write-back stores: %.4f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { kA<<<planes, 288>>>(d + 1, 0); }); printf("A  the same, buffer shifted by 4 bytes : %.4f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { kA2<<<planes, 256>>>(d, 257); }); printf("A2 256 columns of 257-float rows       : %.4f ms  %.0f GB/s\n", t, planes * 65536.0 * 4 / 1e9 / t * 1e3);
    t = timeit([&] { kA2<<<planes, 256>>>(d, 256); }); printf("A3 256 columns of 256-float rows       : %.4f ms  %.0f GB/s\n", t, planes * 65536.0 * 4 / 1e9 / t * 1e3);
    t = timeit([&] { kB<<<planes, 288>>>(d, total); }); printf("B  aligned 128-byte lines, scalar    : %.4f ms  %.0f GB/s\n", t, gb / t * 1e3);
    t = timeit([&] { kC<<<148 * 16, 256>>>((float4*)d, total / 4); }); printf("C  aligned float4                    : %.4f ms  %.0f GB/s\n", t, gb / t * 1e3);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

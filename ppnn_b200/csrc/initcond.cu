// Random-Fourier initial condition of the reference's generator (src/operators.py:423-445):
//
//   f[b,i,j] = sum_{k,l=-m..m} a[b,k,l] sin(k x_i + l y_j) + c[b,k,l] cos(k x_i + l y_j),
//   x = y = linspace(0, 2 pi, n)          (duplicated endpoint included)
//   out = (f - min f) / (max f - min f) + 0.1                     (per sample)
//
// The reference adds 2 * 81 full-size sin/cos fields per sample (324 element-wise launches for the
// two velocity components).  Here the angle-sum identities make the sum separable,
//
//   f[i,j] = sum_l cos(l y_j) P_l(i) + sin(l y_j) Q_l(i),
//   P_l(i) = sum_k a_kl sin(k x_i) + c_kl cos(k x_i),   Q_l(i) = sum_k a_kl cos(k x_i) - c_kl sin(k x_i)
//
// so a point costs 2 (2m+1) FMAs: one kernel evaluates f (sin / cos tables of the grid and the row
// factors P, Q in shared memory) and reduces min / max per sample with order-independent integer
// atomics, a second one normalises in place.
#include "common.cuh"

namespace ppnn {

constexpr int kIcMaxModes = 8;
constexpr int kIcRows = 8;          // rows of one sample per CTA
constexpr int kIcThreads = 256;

// monotone float <-> int key, so that atomicMin / atomicMax on ints order floats
__device__ __forceinline__ int float_key(float f) {
    int b = __float_as_int(f);
    return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key_float(int k) {
    return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff);
}

// torch.linspace(0, end, n)[i] in fp32 (symmetric evaluation, as ATen's kernel does)
__device__ __forceinline__ float linspace_at(float end, int n, int i) {
    const float step = end / (float)(n - 1);
    return i < n / 2 ? step * (float)i : end - step * (float)(n - i - 1);
}

__global__ void ic_init_kernel(int* keys, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { keys[2 * b] = 0x7fffffff; keys[2 * b + 1] = (int)0x80000000; }
}

__global__ void __launch_bounds__(kIcThreads)
ic_eval_kernel(const float* __restrict__ coef, float* __restrict__ out, int* __restrict__ keys, int n, int modes) {
    extern __shared__ float sm[];
    const int M = 2 * modes + 1;
    float* sy = sm;                       // [modes + 1][n]  sin(l y_j), l = 0..modes
    float* cy = sy + (modes + 1) * n;     // [modes + 1][n]
    float* PQ = cy + (modes + 1) * n;     // [kIcRows][2][M]
    float* cf = PQ + kIcRows * 2 * M;     // [M*M][2] coefficients of this sample
    __shared__ int kmin, kmax;
    const int b = blockIdx.y, i0 = blockIdx.x * kIcRows, tid = threadIdx.x;
    const float two_pi = 6.283185307179586f;        // float(2*pi), what torch.linspace receives
    if (tid == 0) { kmin = 0x7fffffff; kmax = (int)0x80000000; }
    for (int e = tid; e < (modes + 1) * n; e += kIcThreads) {
        const int l = e / n, j = e - l * n;
        float s, c;
        sincosf((float)l * linspace_at(two_pi, n, j), &s, &c);
        sy[e] = s; cy[e] = c;
    }
    for (int e = tid; e < 2 * M * M; e += kIcThreads) cf[e] = __ldg(coef + (size_t)b * 2 * M * M + e);
    __syncthreads();
    // row factors: P_l, Q_l for the CTA's rows
    for (int e = tid; e < kIcRows * M; e += kIcThreads) {
        const int r = e / M, li = e - r * M;
        const int i = i0 + r;
        float P = 0.f, Q = 0.f;
        if (i < n) {
            for (int ki = 0; ki < M; ++ki) {
                const int k = ki - modes;
                const float s = k < 0 ? -sy[-k * n + i] : sy[k * n + i];     // x and y share the grid
                const float c = cy[(k < 0 ? -k : k) * n + i];
                const float a = cf[2 * (ki * M + li)], cc = cf[2 * (ki * M + li) + 1];
                P += a * s + cc * c;
                Q += a * c - cc * s;
            }
        }
        PQ[(r * 2) * M + li] = P;
        PQ[(r * 2 + 1) * M + li] = Q;
    }
    __syncthreads();
    float lo = INFINITY, hi = -INFINITY;
    for (int e = tid; e < kIcRows * n; e += kIcThreads) {
        const int r = e / n, j = e - r * n;
        const int i = i0 + r;
        if (i >= n) break;
        const float* P = PQ + (r * 2) * M;
        const float* Q = P + M;
        float f = 0.f;
        for (int li = 0; li < M; ++li) {
            const int l = li - modes;
            const float s = l < 0 ? -sy[-l * n + j] : sy[l * n + j];
            const float c = cy[(l < 0 ? -l : l) * n + j];
            f += c * P[li] + s * Q[li];
        }
        out[((size_t)b * n + i) * n + j] = f;
        lo = fminf(lo, f); hi = fmaxf(hi, f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((tid & 31) == 0) { atomicMin(&kmin, float_key(lo)); atomicMax(&kmax, float_key(hi)); }
    __syncthreads();
    if (tid == 0) { atomicMin(keys + 2 * b, kmin); atomicMax(keys + 2 * b + 1, kmax); }
}

__global__ void __launch_bounds__(256)
ic_normalise_kernel(float* __restrict__ out, const int* __restrict__ keys, size_t plane, float offset) {
    const int b = blockIdx.y;
    const float lo = key_float(keys[2 * b]), hi = key_float(keys[2 * b + 1]);
    const float span = hi - lo;
    float* o = out + (size_t)b * plane;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += (size_t)gridDim.x * blockDim.x)
        o[e] = (o[e] - lo) / span + offset;
}

}  // namespace ppnn

using namespace ppnn;

extern "C" int ppnn_fourier_ic(const float* coef, float* out, float* minmax_scratch,
                               int B, int n, int modes, float offset, void* stream) {
    if (!coef || !out || !minmax_scratch) return PPNN_E_NULL;
    if (B < 0 || n < 2 || modes < 0 || modes > kIcMaxModes) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    const int M = 2 * modes + 1;
    const size_t smem = ((size_t)2 * (modes + 1) * n + (size_t)kIcRows * 2 * M + (size_t)2 * M * M) * sizeof(float);
    if (smem > 200 * 1024) return PPNN_E_TOOBIG;
    cudaStream_t s = (cudaStream_t)stream;
    int* keys = reinterpret_cast<int*>(minmax_scratch);
    for (int b0 = 0; b0 < B; b0 += 65535) {                     // samples ride on gridDim.y
        const int nb = B - b0 < 65535 ? B - b0 : 65535;
        ic_init_kernel<<<(nb + 255) / 256, 256, 0, s>>>(keys + 2 * (size_t)b0, nb);
        cudaError_t e = cudaFuncSetAttribute(ic_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((n + kIcRows - 1) / kIcRows, nb);
        ic_eval_kernel<<<grid, kIcThreads, smem, s>>>(coef + (size_t)b0 * 2 * M * M, out + (size_t)b0 * n * n,
                                                      keys + 2 * (size_t)b0, n, modes);
        const size_t plane = (size_t)n * n;
        int nblk = (int)((plane + 256 * 8 - 1) / (256 * 8));
        if (nblk > 1024) nblk = 1024;
        ic_normalise_kernel<<<dim3(nblk, nb), 256, 0, s>>>(out + (size_t)b0 * n * n, keys + 2 * (size_t)b0, plane, offset);
    }
    return launch_status();
}

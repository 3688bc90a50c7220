// Tier A: the finite-difference operator modules of src/operators.py as
// stand-alone kernels on the caller's PRE-PADDED input ('valid' correlation),
// forward and adjoint.  d2udx2_2D / d2udy2_2D / dudx_2D / dudy_2D
// (operators.py:228-273) and, with Hin = 1, diffusion1D / dudx_1D (:74-128).
//
// Bandwidth-bound: 4 B read + 4 B written per output point.  Along H a thread
// owns one column and a run of ROWS outputs so every input row is loaded once
// per run; along W a thread computes VEC consecutive outputs from K+VEC-1 loads.
#include "common.cuh"

namespace ppnn {

constexpr int kRows = 8;   // outputs per thread along H
constexpr int kVec = 4;    // outputs per thread along W

struct StencilParams {
    const float* in; float* out;
    int Hin, Win, Hout, Wout;
    float a[kMaxTaps], b[kMaxTaps];
    int centre;            // offset of the mask value inside the window (Upwind1)
};

template <int K, bool UPWIND>
__device__ __forceinline__ float apply_taps(const StencilParams& p, const float* w) {
    float f = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) f += p.a[k] * w[k];
    if (!UPWIND) return f;
    float bsum = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) bsum += p.b[k] * w[k];
    // (c <= 0) * fwd + (c > 0) * bwd   (operators.py:257-258)
    float c = w[(K - 1) / 2];
    return c <= 0.f ? f : (c > 0.f ? bsum : 0.f);
}

// ---- along H (axis 2): grid (col chunks, row runs, planes) -------------------
template <int K, bool UPWIND>
__global__ void __launch_bounds__(128)
stencil_x_kernel(const StencilParams p) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.Win) return;
    int i0 = blockIdx.y * kRows;
    const float* in = p.in + (size_t)blockIdx.z * p.Hin * p.Win + j;
    float* out = p.out + (size_t)blockIdx.z * p.Hout * p.Wout + j;
    float w[K + kRows - 1];
#pragma unroll
    for (int k = 0; k < K + kRows - 1; ++k) {
        int r = i0 + k;
        w[k] = r < p.Hin ? ld_stream(in + (size_t)r * p.Win) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < kRows; ++q)
        if (i0 + q < p.Hout) st_stream(out + (size_t)(i0 + q) * p.Wout, apply_taps<K, UPWIND>(p, w + q));
}

// ---- along W (axis 3): grid (col groups, row runs, planes) ---------------------
template <int K, bool UPWIND>
__global__ void __launch_bounds__(128)
stencil_y_kernel(const StencilParams p) {
    int jg = (blockIdx.x * blockDim.x + threadIdx.x) * kVec;
    if (jg >= p.Wout) return;
    const int i0 = blockIdx.y * kRows;
    const float* in = p.in + ((size_t)blockIdx.z * p.Hin + i0) * p.Win + jg;
    float* out = p.out + ((size_t)blockIdx.z * p.Hout + i0) * p.Wout + jg;
    const int rows = min(kRows, p.Hout - i0);
    for (int r = 0; r < rows; ++r, in += p.Win, out += p.Wout) {
        float w[K + kVec - 1];
#pragma unroll
        for (int k = 0; k < K + kVec - 1; ++k) w[k] = (jg + k < p.Win) ? __ldg(in + k) : 0.f;
#pragma unroll
        for (int q = 0; q < kVec; ++q)
            if (jg + q < p.Wout) st_stream(out + q, apply_taps<K, UPWIND>(p, w + q));
    }
}

// ---- adjoint: grad_in[m] = sum_k coef(o = m-k)[k] * g[o], thread per element --
struct StencilBwdParams {
    const float* g; const float* in; float* gin;
    int Hin, Win, Hout, Wout, axis;
    long total;
    float a[kMaxTaps], b[kMaxTaps];
};

template <int K, bool UPWIND>
__global__ void __launch_bounds__(256)
stencil_bwd_kernel(const StencilBwdParams p) {
    constexpr int C = (K - 1) / 2;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long)gridDim.x * blockDim.x) {
        int j = (int)(idx % p.Win);
        long t = idx / p.Win;
        int i = (int)(t % p.Hin);
        long pl = t / p.Hin;
        const float* g = p.g + (size_t)pl * p.Hout * p.Wout;
        const float* in = UPWIND ? p.in + (size_t)pl * p.Hin * p.Win : nullptr;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            int oi = p.axis == 2 ? i - k : i, oj = p.axis == 2 ? j : j - k;
            if (oi < 0 || oj < 0 || oi >= p.Hout || oj >= p.Wout) continue;
            float coef = p.a[k];
            if (UPWIND) {
                float c = p.axis == 2 ? __ldg(in + (size_t)(oi + C) * p.Win + oj) : __ldg(in + (size_t)oi * p.Win + oj + C);
                coef = c <= 0.f ? p.a[k] : (c > 0.f ? p.b[k] : 0.f);
            }
            s += coef * __ldg(g + (size_t)oi * p.Wout + oj);
        }
        p.gin[idx] = s;
    }
}

template <int K, bool UPWIND>
static int launch_fwd(const StencilParams& p, long planes, int axis, cudaStream_t s) {
    if (axis == 2) {
        dim3 grid((p.Win + 127) / 128, (p.Hout + kRows - 1) / kRows, (unsigned)planes);
        stencil_x_kernel<K, UPWIND><<<grid, 128, 0, s>>>(p);
    } else {
        int groups = (p.Wout + kVec - 1) / kVec;
        dim3 grid((groups + 127) / 128, (p.Hout + kRows - 1) / kRows, (unsigned)planes);
        stencil_y_kernel<K, UPWIND><<<grid, 128, 0, s>>>(p);
    }
    return launch_status();
}

template <int K, bool UPWIND>
static int launch_bwd(const StencilBwdParams& p, cudaStream_t s) {
    long blocks = (p.total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    stencil_bwd_kernel<K, UPWIND><<<(unsigned)blocks, 256, 0, s>>>(p);
    return launch_status();
}

#define PPNN_DISPATCH_K(K_, UP_, CALL)                         \
    switch (K_) {                                              \
    case 2: if (!(UP_)) return CALL(2, false); break;          \
    case 3: return (UP_) ? CALL(3, true) : CALL(3, false);     \
    case 4: if (!(UP_)) return CALL(4, false); break;          \
    case 5: return (UP_) ? CALL(5, true) : CALL(5, false);     \
    case 7: return (UP_) ? CALL(7, true) : CALL(7, false);     \
    case 9: if (!(UP_)) return CALL(9, false); break;          \
    }                                                          \
    return PPNN_E_SCHEME;

static int check_geometry(long planes, int Hin, int Win, int axis, int K, int* Hout, int* Wout) {
    if (planes < 0 || Hin < 1 || Win < 1) return PPNN_E_SHAPE;
    if (axis != 2 && axis != 3) return PPNN_E_MODE;
    *Hout = axis == 2 ? Hin - K + 1 : Hin;
    *Wout = axis == 3 ? Win - K + 1 : Win;
    if (*Hout < 1 || *Wout < 1) return PPNN_E_SHAPE;
    if (planes > 65535 || *Hout > 65535 * kRows) return PPNN_E_SHAPE;
    return PPNN_OK;
}

}  // namespace ppnn

using namespace ppnn;

extern "C" int ppnn_stencil_fwd(const float* in, float* out, long planes, int Hin, int Win,
                                int axis, int scheme, int accuracy, void* stream) {
    if (!in || !out) return PPNN_E_NULL;
    Taps ta, tb;
    if (!make_taps(scheme, accuracy, &ta, &tb)) return PPNN_E_SCHEME;
    StencilParams p;
    int st = check_geometry(planes, Hin, Win, axis, ta.n, &p.Hout, &p.Wout);
    if (st != PPNN_OK) return st;
    if (planes == 0) return PPNN_OK;
    p.in = in; p.out = out; p.Hin = Hin; p.Win = Win; p.centre = accuracy;
    for (int k = 0; k < kMaxTaps; ++k) { p.a[k] = ta.c[k]; p.b[k] = scheme == PPNN_UPWIND1 ? tb.c[k] : 0.f; }
    const bool up = scheme == PPNN_UPWIND1;
    cudaStream_t s = (cudaStream_t)stream;
#define CALL_FWD(K, U) launch_fwd<K, U>(p, planes, axis, s)
    PPNN_DISPATCH_K(ta.n, up, CALL_FWD)
#undef CALL_FWD
}

extern "C" int ppnn_stencil_bwd(const float* grad_out, const float* in, float* grad_in,
                                long planes, int Hin, int Win, int axis, int scheme,
                                int accuracy, void* stream) {
    if (!grad_out || !grad_in) return PPNN_E_NULL;
    if (scheme == PPNN_UPWIND1 && !in) return PPNN_E_NULL;
    Taps ta, tb;
    if (!make_taps(scheme, accuracy, &ta, &tb)) return PPNN_E_SCHEME;
    StencilBwdParams p;
    int st = check_geometry(planes, Hin, Win, axis, ta.n, &p.Hout, &p.Wout);
    if (st != PPNN_OK) return st;
    if (planes == 0) return PPNN_OK;
    p.g = grad_out; p.in = in; p.gin = grad_in; p.Hin = Hin; p.Win = Win; p.axis = axis;
    p.total = planes * Hin * Win;
    for (int k = 0; k < kMaxTaps; ++k) { p.a[k] = ta.c[k]; p.b[k] = scheme == PPNN_UPWIND1 ? tb.c[k] : 0.f; }
    const bool up = scheme == PPNN_UPWIND1;
    cudaStream_t s = (cudaStream_t)stream;
#define CALL_BWD(K, U) launch_bwd<K, U>(p, s)
    PPNN_DISPATCH_K(ta.n, up, CALL_BWD)
#undef CALL_BWD
}

// Tier A: the finite-difference operator modules of src/operators.py as
// stand-alone kernels on the caller's PRE-PADDED input ('valid' correlation),
// forward and adjoint.  d2udx2_2D / d2udy2_2D / dudx_2D / dudy_2D
// (operators.py:228-273) and, with Hin = 1, diffusion1D / dudx_1D (:74-128).
//
// Bandwidth-bound: 4 B read + 4 B written per output point.  Along H a thread
// owns one column and a run of ROWS outputs so every input row is loaded once
// per run; along W a thread computes VEC consecutive outputs from K+VEC-1 loads.
#include "common.cuh"
#include "tma.cuh"

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

// ---- along H, staged: (rows + K - 1) x 512 input tile by TMA, 4 columns per thread ----
// Same idea as the W-direction kernel below: no global load instructions, the whole tile is
// in flight at once; a thread marches down its 4 columns with a rolling window of K float4
// read by LDS.128 and stores float4.  Rows are 16-byte aligned here (width % 4 == 0).
constexpr int kXTileIn = 22;                        // input rows per tile: 22 - K + 1 output rows
constexpr int kXTileCols = 512;

template <int K, bool UPWIND>
__global__ void __launch_bounds__(128)
stencil_x_tma_kernel(const StencilParams p) {
    constexpr int TR = kXTileIn - K + 1;
    __shared__ __align__(16) float tile[kXTileIn * kXTileCols];
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x;
    const int j0 = blockIdx.x * kXTileCols;
    const int i0 = blockIdx.y * TR;
    const int rows = min(TR, p.Hout - i0);
    const int wseg = min(kXTileCols, p.Win - j0);               // multiple of 4
    const float* in = p.in + ((size_t)blockIdx.z * p.Hin + i0) * p.Win + j0;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&bar, (uint32_t)(rows + K - 1) * wseg * 4);
        for (int r = 0; r < rows + K - 1; ++r)
            bulk_g2s(tile + r * kXTileCols, in + (size_t)r * p.Win, (uint32_t)wseg * 4, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    const int jg = j0 + 4 * t;
    if (jg >= p.Win) return;
    const float4* sp = reinterpret_cast<const float4*>(tile) + t;
    float* out = p.out + ((size_t)blockIdx.z * p.Hout + i0) * p.Wout + jg;
    float w[4][K];
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
        const float4 q = sp[k * (kXTileCols / 4)];
        w[0][k + 1] = q.x; w[1][k + 1] = q.y; w[2][k + 1] = q.z; w[3][k + 1] = q.w;
    }
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        if (r >= rows) break;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int k = 0; k < K - 1; ++k) w[c][k] = w[c][k + 1];
        const float4 q = sp[(r + K - 1) * (kXTileCols / 4)];
        w[0][K - 1] = q.x; w[1][K - 1] = q.y; w[2][K - 1] = q.z; w[3][K - 1] = q.w;
        const float o0 = apply_taps<K, UPWIND>(p, w[0]), o1 = apply_taps<K, UPWIND>(p, w[1]),
                    o2 = apply_taps<K, UPWIND>(p, w[2]), o3 = apply_taps<K, UPWIND>(p, w[3]);
        asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + (size_t)r * p.Wout), "f"(o0), "f"(o1), "f"(o2), "f"(o3) : "memory");
    }
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

// ---- along W, staged: TMA row segments -> shared memory -> 4 outputs per thread ----
// A 'valid' stencil along W reads K consecutive floats per output from rows whose pitch
// (Wout + K - 1) is not a multiple of 16 bytes, so neither vector loads nor a K-fold reuse in
// registers come for free.  Here one thread issues a `cp.async.bulk` per row of the tile (from the
// enclosing 16-byte boundary; `lead` = 0..3 floats to skip, per row), the CTA waits on one
// mbarrier, and every thread reads FOUR aligned float4 per row and picks its K+3 window values out
// of those 16 registers (the four `lead` cases are separate, fully unrolled code paths), computes
// 4 consecutive outputs and stores them as one float4.  LSU work per output: 1 LDS.128 + 1/4
// STG.128, no global load instructions at all.
constexpr int kYTileRows = 8;
constexpr int kYTileCols = 512;                     // outputs per row of a tile: 4 per thread, 128 threads
constexpr int kYPitch = kYTileCols + 16;            // floats; chunk loads reach 4*t + 15

template <int K, bool UPWIND, int LEAD>
__device__ __forceinline__ void ytile_row(const StencilParams& p, const float* srow, float* orow, int t, int jg) {
    const float4* c = reinterpret_cast<const float4*>(srow) + t;
    const float4 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    const float v[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
    float w[K + 3];
#pragma unroll
    for (int k = 0; k < K + 3; ++k) w[k] = v[LEAD + k];
    float4 o;
    o.x = apply_taps<K, UPWIND>(p, w);
    o.y = apply_taps<K, UPWIND>(p, w + 1);
    o.z = apply_taps<K, UPWIND>(p, w + 2);
    o.w = apply_taps<K, UPWIND>(p, w + 3);
    if (jg + 3 < p.Wout) {
        asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(orow), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
    } else {
        if (jg < p.Wout) st_stream(orow, o.x);
        if (jg + 1 < p.Wout) st_stream(orow + 1, o.y);
        if (jg + 2 < p.Wout) st_stream(orow + 2, o.z);
    }
}

template <int K, bool UPWIND>
__global__ void __launch_bounds__(128)
stencil_y_tma_kernel(const StencilParams p, const float* in_end) {
    __shared__ __align__(16) float tile[kYTileRows * kYPitch];
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x;
    const int j0 = blockIdx.x * kYTileCols;
    const int i0 = blockIdx.y * kYTileRows;
    const int rows = min(kYTileRows, p.Hout - i0);
    const int wseg = min(kYTileCols + K - 1, p.Win - j0);      // input floats of this tile per row
    const float* in = p.in + ((size_t)blockIdx.z * p.Hin + i0) * p.Win + j0;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        const uintptr_t end_al = reinterpret_cast<uintptr_t>(in_end) & ~(uintptr_t)15;
        uint32_t total = 0, nb[kYTileRows]; uintptr_t a0[kYTileRows];
#pragma unroll
        for (int r = 0; r < kYTileRows; ++r) {
            nb[r] = 0;
            if (r >= rows) continue;
            const uintptr_t a = reinterpret_cast<uintptr_t>(in + (size_t)r * p.Win);
            a0[r] = a & ~(uintptr_t)15;
            nb[r] = (uint32_t)((a - a0[r]) + (size_t)wseg * 4 + 15) & ~15u;
            if (a0[r] + nb[r] > end_al) {                        // last bytes of the tensor: stay inside it
                float* dst = tile + r * kYPitch;
                for (const float* q = reinterpret_cast<const float*>(end_al); q < in_end; ++q)
                    dst[(reinterpret_cast<uintptr_t>(q) - a0[r]) >> 2] = *q;
                nb[r] = (uint32_t)(end_al - a0[r]);
            }
            total += nb[r];
        }
        mbar_expect_tx(&bar, total);
#pragma unroll
        for (int r = 0; r < kYTileRows; ++r)
            if (nb[r]) bulk_g2s(tile + r * kYPitch, reinterpret_cast<const void*>(a0[r]), nb[r], &bar);
    }
    __syncthreads();                                             // barrier initialised (and the tail floats written)
    mbar_wait(&bar, 0);
    const int jg = j0 + 4 * t;
    if (jg >= p.Wout) return;
    float* out = p.out + ((size_t)blockIdx.z * p.Hout + i0) * p.Wout + jg;
    for (int r = 0; r < rows; ++r, out += p.Wout) {
        const float* srow = tile + r * kYPitch;
        const int lead = (int)(reinterpret_cast<uintptr_t>(in + (size_t)r * p.Win) & 15) >> 2;   // warp-uniform
        switch (lead) {
        case 0: ytile_row<K, UPWIND, 0>(p, srow, out, t, jg); break;
        case 1: ytile_row<K, UPWIND, 1>(p, srow, out, t, jg); break;
        case 2: ytile_row<K, UPWIND, 2>(p, srow, out, t, jg); break;
        default: ytile_row<K, UPWIND, 3>(p, srow, out, t, jg); break;
        }
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
    if (axis == 2 && p.Win % 4 == 0 && ((reinterpret_cast<uintptr_t>(p.in) | reinterpret_cast<uintptr_t>(p.out)) & 15) == 0 &&
        (p.Hout + 7) / 8 <= 65535) {
        constexpr int TR = kXTileIn - K + 1;
        dim3 grid((p.Win + kXTileCols - 1) / kXTileCols, (p.Hout + TR - 1) / TR, (unsigned)planes);
        stencil_x_tma_kernel<K, UPWIND><<<grid, 128, 0, s>>>(p);
    } else if (axis == 2) {
        dim3 grid((p.Win + 127) / 128, (p.Hout + kRows - 1) / kRows, (unsigned)planes);
        stencil_x_kernel<K, UPWIND><<<grid, 128, 0, s>>>(p);
    } else if (p.Wout % 4 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && K + 3 + 3 <= 16 &&
               (p.Hout + kYTileRows - 1) / kYTileRows <= 65535) {
        // float4 stores need 16-byte aligned output rows; everything else takes the scalar kernel below
        dim3 grid((p.Wout + kYTileCols - 1) / kYTileCols, (p.Hout + kYTileRows - 1) / kYTileRows, (unsigned)planes);
        stencil_y_tma_kernel<K, UPWIND><<<grid, 128, 0, s>>>(p, p.in + (size_t)planes * p.Hin * p.Win);
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

static int stencil_fwd_chunk(const StencilParams& p, long planes, int axis, int K, bool up, cudaStream_t s) {
#define CALL_FWD(K, U) launch_fwd<K, U>(p, planes, axis, s)
    PPNN_DISPATCH_K(K, up, CALL_FWD)
#undef CALL_FWD
}

constexpr long kMaxGridZ = 65535;

static int check_geometry(long planes, int Hin, int Win, int axis, int K, int* Hout, int* Wout) {
    if (planes < 0 || Hin < 1 || Win < 1) return PPNN_E_SHAPE;
    if (axis != 2 && axis != 3) return PPNN_E_MODE;
    *Hout = axis == 2 ? Hin - K + 1 : Hin;
    *Wout = axis == 3 ? Win - K + 1 : Win;
    if (*Hout < 1 || *Wout < 1) return PPNN_E_SHAPE;
    if (*Hout > 65535 * kRows) return PPNN_E_SHAPE;
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
    // planes ride on gridDim.z (<= 65535): larger batches go out as several launches
    for (long p0 = 0; p0 < planes; p0 += kMaxGridZ) {
        const long n = planes - p0 < kMaxGridZ ? planes - p0 : kMaxGridZ;
        p.in = in + (size_t)p0 * Hin * Win;
        p.out = out + (size_t)p0 * p.Hout * p.Wout;
        st = stencil_fwd_chunk(p, n, axis, ta.n, up, (cudaStream_t)stream);
        if (st != PPNN_OK) return st;
    }
    return PPNN_OK;
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

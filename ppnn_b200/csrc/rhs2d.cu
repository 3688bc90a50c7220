// Tier A: one rhs.py function (src/rhs.py:1-47,52-56) evaluated in a single
// pass on the caller's pre-padded halos ax [B,H+2h,W] / ay [B,H,W+2h] -- what
// train2D.py:75-77 and operators.py:474-475 compute with ~30 ATen launches.
// 20 B per output point (u1, v1, ax, ay read, out written).
#include "common.cuh"

namespace ppnn {

struct RhsParams {
    const float* u1; const float* v1; const float* mu; const float* ax; const float* ay;
    float* out;
    int H, W, halo, K;         // K = 2*halo + 1 taps for every operator
    int rhs_id, adv_upwind;
    float aa[kMaxTaps], ab[kMaxTaps], d[kMaxTaps];
    float dx, dy, dx2, dy2;
    long total;
};

constexpr int kRhsRows = 4;     // output rows per thread: the x-direction window is loaded once for all of them

// thread = column j, kRhsRows consecutive output rows; grid (column chunks, row groups, samples)
template <int K>
__global__ void __launch_bounds__(128)
rhs2d_kernel(const RhsParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.W) return;
    const int i0 = blockIdx.y * kRhsRows;
    const long b = blockIdx.z;
    const int Wy = p.W + 2 * p.halo;
    const float mu = __ldg(p.mu + b);
    const float* cx = p.ax + ((size_t)b * (p.H + 2 * p.halo) + i0) * p.W + j;     // padded rows i0 .. i0+K+kRhsRows-2
    float xw[K + kRhsRows - 1];
#pragma unroll
    for (int k = 0; k < K + kRhsRows - 1; ++k)
        xw[k] = (i0 + k < p.H + 2 * p.halo) ? ld_stream(cx + (size_t)k * p.W) : 0.f;
#pragma unroll
    for (int q = 0; q < kRhsRows; ++q) {
        const int i = i0 + q;
        if (i >= p.H) break;
        const size_t o = ((size_t)b * p.H + i) * p.W + j;
        const float u1 = ld_stream(p.u1 + o), v1 = ld_stream(p.v1 + o);
        const float* cy = p.ay + ((size_t)b * p.H + i) * Wy + j;                   // padded cols j .. j+K-1
        float yw[K];
#pragma unroll
        for (int k = 0; k < K; ++k) yw[k] = __ldg(cy + k);
        float acc = 0.f;
        if (p.rhs_id & PPNN_RHS_ADV) {
            float fx = 0.f, bx = 0.f, fy = 0.f, by = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fx += p.aa[k] * xw[q + k]; fy += p.aa[k] * yw[k];
                bx += p.ab[k] * xw[q + k]; by += p.ab[k] * yw[k];
            }
            if (p.adv_upwind) {
                const float c1 = xw[q + (K - 1) / 2], c2 = yw[(K - 1) / 2];
                fx = c1 <= 0.f ? fx : (c1 > 0.f ? bx : 0.f);
                fy = c2 <= 0.f ? fy : (c2 > 0.f ? by : 0.f);
            }
            acc = (-u1 * fx) / p.dx - (v1 * fy) / p.dy;
        }
        if (p.rhs_id & PPNN_RHS_DIFF) {
            float lx = 0.f, ly = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) { lx += p.d[k] * xw[q + k]; ly += p.d[k] * yw[k]; }
            const float lap = mu * (lx / p.dx2 + ly / p.dy2);
            acc = (p.rhs_id & PPNN_RHS_ADV) ? acc + lap : lap;
        }
        st_stream(p.out + o, add_reaction(acc, p.rhs_id & ~3, u1, v1));
    }
}

template <int K>
static int launch_rhs2d(const RhsParams& p, int B, cudaStream_t s) {
    dim3 grid((p.W + 127) / 128, (p.H + kRhsRows - 1) / kRhsRows, (unsigned)B);
    rhs2d_kernel<K><<<grid, 128, 0, s>>>(p);
    return launch_status();
}

}  // namespace ppnn

using namespace ppnn;

extern "C" int ppnn_rhs2d_fwd(const float* u1, const float* v1, const float* mu,
                              const float* ax, const float* ay, float* out,
                              int B, int H, int W, int halo, int rhs_id,
                              int adv_scheme, int adv_accuracy, int diff_scheme, int diff_accuracy,
                              float dx, float dy, float dx2, float dy2, void* stream) {
    if (!u1 || !v1 || !mu || !ax || !ay || !out) return PPNN_E_NULL;
    if (B < 0 || H < 1 || W < 1 || halo < 0) return PPNN_E_SHAPE;
    if (!valid_rhs_id(rhs_id)) return PPNN_E_RHS;
    RhsParams p;
    p.K = 2 * halo + 1;
    if (p.K > kMaxTaps) return PPNN_E_HALO;
    Taps ta, tb, td;
    for (int k = 0; k < kMaxTaps; ++k) p.aa[k] = p.ab[k] = p.d[k] = 0.f;
    if (rhs_id & PPNN_RHS_ADV) {
        if (!make_taps(adv_scheme, adv_accuracy, &ta, &tb)) return PPNN_E_SCHEME;
        if (ta.n != p.K) return PPNN_E_HALO;
        for (int k = 0; k < p.K; ++k) { p.aa[k] = ta.c[k]; p.ab[k] = adv_scheme == PPNN_UPWIND1 ? tb.c[k] : 0.f; }
    }
    if (rhs_id & PPNN_RHS_DIFF) {
        if (!make_taps(diff_scheme, diff_accuracy, &td, nullptr)) return PPNN_E_SCHEME;
        if (td.n != p.K || diff_scheme == PPNN_UPWIND1) return PPNN_E_HALO;
        for (int k = 0; k < p.K; ++k) p.d[k] = td.c[k];
    }
    if (B == 0) return PPNN_OK;
    p.u1 = u1; p.v1 = v1; p.mu = mu; p.ax = ax; p.ay = ay; p.out = out;
    p.H = H; p.W = W; p.halo = halo; p.rhs_id = rhs_id; p.adv_upwind = adv_scheme == PPNN_UPWIND1;
    p.dx = dx; p.dy = dy; p.dx2 = dx2; p.dy2 = dy2;
    p.total = (long)B * H * W;
    if (B > 65535 || (H + kRhsRows - 1) / kRhsRows > 65535) return PPNN_E_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    switch (p.K) {
    case 1: return launch_rhs2d<1>(p, B, s);
    case 3: return launch_rhs2d<3>(p, B, s);
    case 5: return launch_rhs2d<5>(p, B, s);
    case 7: return launch_rhs2d<7>(p, B, s);
    case 9: return launch_rhs2d<9>(p, B, s);
    }
    return PPNN_E_HALO;
}

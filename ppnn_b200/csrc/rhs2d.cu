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

__global__ void __launch_bounds__(256)
rhs2d_kernel(const RhsParams p) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long)gridDim.x * blockDim.x) {
        int j = (int)(idx % p.W);
        long t = idx / p.W;
        int i = (int)(t % p.H);
        long b = t / p.H;
        const float mu = __ldg(p.mu + b);
        const float u1 = ld_stream(p.u1 + idx), v1 = ld_stream(p.v1 + idx);
        const float* cx = p.ax + ((size_t)b * (p.H + 2 * p.halo) + i) * p.W + j;        // rows i .. i+K-1
        const float* cy = p.ay + ((size_t)b * p.H + i) * (p.W + 2 * p.halo) + j;        // cols j .. j+K-1
        float acc = 0.f;
        if (p.rhs_id & PPNN_RHS_ADV) {
            float fx = 0.f, bx = 0.f, fy = 0.f, by = 0.f;
            for (int k = 0; k < p.K; ++k) {
                float xv = __ldg(cx + (size_t)k * p.W), yv = __ldg(cy + k);
                fx += p.aa[k] * xv; fy += p.aa[k] * yv;
                bx += p.ab[k] * xv; by += p.ab[k] * yv;
            }
            if (p.adv_upwind) {
                float c1 = __ldg(cx + (size_t)p.halo * p.W), c2 = __ldg(cy + p.halo);
                fx = c1 <= 0.f ? fx : (c1 > 0.f ? bx : 0.f);
                fy = c2 <= 0.f ? fy : (c2 > 0.f ? by : 0.f);
            }
            acc = (-u1 * fx) / p.dx - (v1 * fy) / p.dy;
        }
        if (p.rhs_id & PPNN_RHS_DIFF) {
            float lx = 0.f, ly = 0.f;
            for (int k = 0; k < p.K; ++k) {
                lx += p.d[k] * __ldg(cx + (size_t)k * p.W);
                ly += p.d[k] * __ldg(cy + k);
            }
            float lap = mu * (lx / p.dx2 + ly / p.dy2);
            acc = (p.rhs_id & PPNN_RHS_ADV) ? acc + lap : lap;
        }
        st_stream(p.out + idx, add_reaction(acc, p.rhs_id & ~3, u1, v1));
    }
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
    long blocks = (p.total + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    rhs2d_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return launch_status();
}

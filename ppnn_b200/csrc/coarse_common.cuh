// Device-side pieces shared by the fused coarse-grid kernels (forward, adjoint,
// coarse-only solver): the bilinear restriction gather, the periodic RHS on the
// N x N coarse grid held in shared memory, and the bicubic index/weight tables.
//
// Semantics follow SURVEY.md appendix A.2, i.e. src/train2D.py:66-79 on top of
// src/rhs.py, src/operators.py:228-273 and src/utility/utils.py:24,30-31.
#pragma once
#include "common.cuh"

namespace ppnn {

// idx / n for 0 <= idx < 2^21 and 1 <= n without the ~20-instruction integer-division sequence: (idx + 0.5) / n is
// at least 0.5 / n away from an integer and the fp32 product is off by less than idx / n * 2^-22.
__device__ __forceinline__ int small_div(int idx, int n) {
    return __float2int_rz(((float)idx + 0.5f) * (1.0f / (float)n));
}


// Kernel-parameter block (lives in the constant bank; taps are read from there).
struct CoarseParams {
    int FH, FW;          // fine points incl. duplicated endpoint
    int CH, CW;          // coarse points incl. duplicated endpoint
    int NH, NW;          // periodic coarse grid, C - 1
    int rhs_u, rhs_v;
    int adv_n;           // taps per one-sided filter = accuracy + 1
    int diff_n, diff_off;
    float adv_f[4];      // Forward1[acc]  at offsets 0..acc
    float adv_b[4];      // Backward1[acc] at offsets -acc..0
    float diff[kMaxTaps];
    float dt, dx, dy, dx2, dy2;
    float rdx, rdy, rdx2, rdy2;   // fp32 reciprocals (for the correctly rounded fast division)
    float sdh, sdw;      // restriction scales (F-1)/(C-1)
    float suh, suw;      // prolongation scales (C-1)/(F-1)
};

inline int fill_coarse_params(const ppnn_coarse_cfg* cfg, CoarseParams* p) {
    if (!cfg) return PPNN_E_NULL;
    if (cfg->CH < 3 || cfg->CW < 3 || cfg->FH < 2 || cfg->FW < 2) return PPNN_E_SHAPE;
    if (!valid_rhs_id(cfg->rhs_u) || !valid_rhs_id(cfg->rhs_v)) return PPNN_E_RHS;
    Taps f, b, d;
    if (!make_taps(PPNN_FORWARD1, cfg->adv_accuracy, &f, nullptr) ||
        !make_taps(PPNN_BACKWARD1, cfg->adv_accuracy, &b, nullptr) ||
        !make_taps(PPNN_CENTRAL2, cfg->diff_accuracy, &d, nullptr))
        return PPNN_E_SCHEME;
    p->FH = cfg->FH; p->FW = cfg->FW; p->CH = cfg->CH; p->CW = cfg->CW;
    p->NH = cfg->CH - 1; p->NW = cfg->CW - 1;
    int halo = cfg->adv_accuracy > -d.off ? cfg->adv_accuracy : -d.off;
    if (p->NH <= halo || p->NW <= halo) return PPNN_E_SHAPE;
    p->rhs_u = cfg->rhs_u; p->rhs_v = cfg->rhs_v;
    p->adv_n = f.n;
    for (int k = 0; k < 4; ++k) { p->adv_f[k] = k < f.n ? f.c[k] : 0.f; p->adv_b[k] = k < b.n ? b.c[k] : 0.f; }
    p->diff_n = d.n; p->diff_off = d.off;
    for (int k = 0; k < kMaxTaps; ++k) p->diff[k] = d.c[k];
    p->dt = cfg->dt; p->dx = cfg->dx; p->dy = cfg->dy; p->dx2 = cfg->dx2; p->dy2 = cfg->dy2;
    p->rdx = 1.f / cfg->dx; p->rdy = 1.f / cfg->dy; p->rdx2 = 1.f / cfg->dx2; p->rdy2 = 1.f / cfg->dy2;
    p->sdh = resample_scale(cfg->FH, cfg->CH, true);
    p->sdw = resample_scale(cfg->FW, cfg->CW, true);
    p->suh = resample_scale(cfg->CH, cfg->FH, true);
    p->suw = resample_scale(cfg->CW, cfg->FW, true);
    return PPNN_OK;
}

#ifdef __CUDACC__

// ---------------------------------------------------------------------------
// bilinear restriction of one coarse point (I, J), I < NH, J < NW, from a fine
// plane [FH, FW]   (mesh_convertor.down then [:, :, :-1, :-1], train2D.py:67-68;
// torch upsample_bilinear2d, align_corners=True).  Zero-weight taps are skipped:
// for F-1 a multiple of C-1 the restriction is an exact decimation (quirk Q6).
// ---------------------------------------------------------------------------
struct BilinearTap {
    int   i0, j0, di, dj;     // di, dj in {0, 1}: whether the +1 neighbour exists
    float li, lj;             // weights of the +1 neighbours
};

__device__ __forceinline__ BilinearTap bilinear_tap(const CoarseParams& p, int I, int J) {
    BilinearTap t;
    float si = p.sdh * (float)I, sj = p.sdw * (float)J;
    t.i0 = (int)si; t.j0 = (int)sj;
    t.di = t.i0 < p.FH - 1 ? 1 : 0;
    t.dj = t.j0 < p.FW - 1 ? 1 : 0;
    t.li = si - (float)t.i0;
    t.lj = sj - (float)t.j0;
    return t;
}

__device__ __forceinline__ float restrict_point(const CoarseParams& p, const float* __restrict__ plane,
                                                int I, int J) {
    BilinearTap t = bilinear_tap(p, I, J);
    const float* q = plane + (size_t)t.i0 * p.FW + t.j0;
    float h0 = 1.f - t.li, w0 = 1.f - t.lj;
    float a = __ldg(q);
    float b = (t.lj != 0.f) ? __ldg(q + t.dj) : 0.f;
    float top = w0 * a + t.lj * b;
    if (t.li == 0.f) return h0 * top;
    const float* r = q + (size_t)t.di * p.FW;
    float c = __ldg(r);
    float d = (t.lj != 0.f) ? __ldg(r + t.dj) : 0.f;
    return h0 * top + t.li * (w0 * c + t.lj * d);
}

// ---------------------------------------------------------------------------
// One-sided / central differences on the periodic coarse grid in shared memory.
// f points at an [NH, NW] array; wrap is the true periodic wrap of
// train2D.py:38-42 generalised to `halo` cells.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float upwind_x(const CoarseParams& p, const float* f, int I, int J) {
    float c = f[I * p.NW + J];
    float s = 0.f;
    if (c <= 0.f) {
        for (int k = 0; k < p.adv_n; ++k) s += p.adv_f[k] * f[wrap1(I + k, p.NH) * p.NW + J];
    } else if (c > 0.f) {
        for (int k = 0; k < p.adv_n; ++k) s += p.adv_b[k] * f[wrap1(I - (p.adv_n - 1) + k, p.NH) * p.NW + J];
    }
    return s;
}

__device__ __forceinline__ float upwind_y(const CoarseParams& p, const float* f, int I, int J) {
    float c = f[I * p.NW + J];
    float s = 0.f;
    if (c <= 0.f) {
        for (int k = 0; k < p.adv_n; ++k) s += p.adv_f[k] * f[I * p.NW + wrap1(J + k, p.NW)];
    } else if (c > 0.f) {
        for (int k = 0; k < p.adv_n; ++k) s += p.adv_b[k] * f[I * p.NW + wrap1(J - (p.adv_n - 1) + k, p.NW)];
    }
    return s;
}

__device__ __forceinline__ float central_x(const CoarseParams& p, const float* f, int I, int J) {
    float s = 0.f;
    for (int k = 0; k < p.diff_n; ++k) s += p.diff[k] * f[wrap1(I + p.diff_off + k, p.NH) * p.NW + J];
    return s;
}

__device__ __forceinline__ float central_y(const CoarseParams& p, const float* f, int I, int J) {
    float s = 0.f;
    for (int k = 0; k < p.diff_n; ++k) s += p.diff[k] * f[I * p.NW + wrap1(J + p.diff_off + k, p.NW)];
    return s;
}

// rhs.py function `id` applied to field f (its halos are what the caller of the
// reference passes as ax / ay), with u1 = cu, v1 = cv.  Operation order mirrors
// rhs.py: (-u1*Dx)/dx - (v1*Dy)/dy + mu*(Lx/dx2 + Ly/dy2) + reaction.
__device__ __forceinline__ float eval_rhs(const CoarseParams& p, int id, const float* f,
                                          const float* cu, const float* cv, int I, int J, float mu) {
    float u1 = cu[I * p.NW + J], v1 = cv[I * p.NW + J];
    float acc = 0.f;
    if (id & PPNN_RHS_ADV)
        acc = div_by(-u1 * upwind_x(p, f, I, J), p.dx, p.rdx) - div_by(v1 * upwind_y(p, f, I, J), p.dy, p.rdy);
    if (id & PPNN_RHS_DIFF) {
        float lap = mu * (div_by(central_x(p, f, I, J), p.dx2, p.rdx2) + div_by(central_y(p, f, I, J), p.dy2, p.rdy2));
        acc = (id & PPNN_RHS_ADV) ? acc + lap : lap;
    }
    return add_reaction(acc, id & ~3, u1, v1);
}

// Both right-hand sides of one coarse point for the stencils every train2D.py case
// uses (Upwind1 accuracy 1, Central2 accuracy 2; train2D.py:53-56): the four
// neighbours are fetched once per field and the tap loops disappear.  Same
// arithmetic as eval_rhs (taps +-1, -2 are exact; sums in tap order).
__device__ __forceinline__ void eval_rhs_pair_3pt(const CoarseParams& p, const float* cu, const float* cv,
                                                  int I, int J, float mu, float* out_u, float* out_v) {
    const int Im = (I == 0 ? p.NH : I) - 1, Ip = I + 1 == p.NH ? 0 : I + 1;
    const int Jm = (J == 0 ? p.NW : J) - 1, Jp = J + 1 == p.NW ? 0 : J + 1;
    const int o = I * p.NW + J;
    const float u1 = cu[o], v1 = cv[o];
#pragma unroll
    for (int f = 0; f < 2; ++f) {
        const float* a = f ? cv : cu;
        const int id = f ? p.rhs_v : p.rhs_u;
        const float c = f ? v1 : u1;
        float acc = 0.f;
        if (id & 3) {
            const float xm = a[Im * p.NW + J], xp = a[Ip * p.NW + J], ym = a[I * p.NW + Jm], yp = a[I * p.NW + Jp];
            if (id & PPNN_RHS_ADV) {
                const float ddx = c <= 0.f ? (-c + xp) : (c > 0.f ? (-xm + c) : 0.f);
                const float ddy = c <= 0.f ? (-c + yp) : (c > 0.f ? (-ym + c) : 0.f);
                acc = div_by(-u1 * ddx, p.dx, p.rdx) - div_by(v1 * ddy, p.dy, p.rdy);
            }
            if (id & PPNN_RHS_DIFF) {
                const float lx = (xm + -2.f * c) + xp, ly = (ym + -2.f * c) + yp;
                const float lap = mu * (div_by(lx, p.dx2, p.rdx2) + div_by(ly, p.dy2, p.rdy2));
                acc = (id & PPNN_RHS_ADV) ? acc + lap : lap;
            }
        }
        acc = p.dt * add_reaction(acc, id & ~3, u1, v1);
        if (f) *out_v = acc; else *out_u = acc;
    }
}

// dt * rhs of ONE field at one coarse point, same 3-point fast path
__device__ __forceinline__ float eval_rhs_3pt(const CoarseParams& p, int id, const float* a,
                                              const float* cu, const float* cv, int I, int J, float mu) {
    const int o = I * p.NW + J;
    const float u1 = cu[o], v1 = cv[o];
    float acc = 0.f;
    if (id & 3) {
        const int Im = (I == 0 ? p.NH : I) - 1, Ip = I + 1 == p.NH ? 0 : I + 1;
        const int Jm = (J == 0 ? p.NW : J) - 1, Jp = J + 1 == p.NW ? 0 : J + 1;
        const float c = a[o];
        const float xm = a[Im * p.NW + J], xp = a[Ip * p.NW + J], ym = a[I * p.NW + Jm], yp = a[I * p.NW + Jp];
        if (id & PPNN_RHS_ADV) {
            const float ddx = c <= 0.f ? (-c + xp) : (c > 0.f ? (-xm + c) : 0.f);
            const float ddy = c <= 0.f ? (-c + yp) : (c > 0.f ? (-ym + c) : 0.f);
            acc = div_by(-u1 * ddx, p.dx, p.rdx) - div_by(v1 * ddy, p.dy, p.rdy);
        }
        if (id & PPNN_RHS_DIFF) {
            const float lx = (xm + -2.f * c) + xp, ly = (ym + -2.f * c) + yp;
            const float lap = mu * (div_by(lx, p.dx2, p.rdx2) + div_by(ly, p.dy2, p.rdy2));
            acc = (id & PPNN_RHS_ADV) ? acc + lap : lap;
        }
    }
    return p.dt * add_reaction(acc, id & ~3, u1, v1);
}

// dt * rhs of both fields over the whole coarse grid, by `nthr` threads
__device__ __forceinline__ void coarse_rhs_phase(const CoarseParams& p, const float* cu, const float* cv,
                                                 float* ru, float* rv, float mu, int tid, int nthr) {
    const int nn2 = p.NH * p.NW;
    if (p.adv_n == 2 && p.diff_n == 3) {
        for (int idx = tid; idx < nn2; idx += nthr) {
            int I = small_div(idx, p.NW), J = idx - I * p.NW;
            eval_rhs_pair_3pt(p, cu, cv, I, J, mu, ru + idx, rv + idx);
        }
    } else {
        for (int idx = tid; idx < 2 * nn2; idx += nthr) {
            int f = idx >= nn2;
            int r = idx - f * nn2;
            int I = small_div(r, p.NW), J = r - I * p.NW;
            ru[idx] = p.dt * eval_rhs(p, f ? p.rhs_v : p.rhs_u, f ? cv : cu, cu, cv, I, J, mu);   // rv follows ru
        }
    }
}

// ---------------------------------------------------------------------------
// Bicubic prolongation tables (torch upsample_bicubic2d, align_corners=True):
// for output index o: i0 = floor(o*scale), t = o*scale - i0, taps i0-1..i0+2.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cubic_entry(float scale, int o, int* i0, float w[4]) {
    float s = scale * (float)o;
    float fl = floorf(s);
    *i0 = (int)fl;
    cubic_weights(s - fl, w);
}

// index of extended coarse row/col (clamped to [0, C-1]) in the periodic N-array:
// padBC_rd (train2D.py:45-47) makes entry N a copy of entry 0.
__device__ __forceinline__ int ext_index(int I, int C) {
    I = clampi(I, 0, C - 1);
    return I == C - 1 ? 0 : I;
}

#endif  // __CUDACC__
}  // namespace ppnn

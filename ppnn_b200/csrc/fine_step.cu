// Fine-grid explicit solver: one fused Euler step of the 2-D Burgers dataset
// generator (src/operators.py:416-420, 466-475) and of the generic
// f <- f + dt*rhs(f) generators (operators.py:494-748), on periodic grids with
// or without the duplicated endpoint.
//
// Three kernels, tried in this order (launch_fine):
//   fine_step_tile.cuh  TMA-fed persistent kernel on column tiles: true-periodic meshes whose
//                       rows are 16-byte aligned (256 ... 8192 columns, e.g. 1024^2)
//   fine_step_ep.cuh    TMA-fed persistent kernel for rows on arbitrary 4-byte phases: the
//                       duplicated-endpoint meshes of the reference (513^2, 1025^2)
//   below               register-window kernel: every other size (257^2, tiny meshes)
//
// Layout: u, v are [B, H, W] fp32, W contiguous.  One CTA = (32k-column chunk,
// PAIR of row bands, sample).  A thread owns one column and walks down both bands
// in lockstep, the two rows packed in float2 so every stencil FMA is one
// Blackwell FFMA2 (packed fp32x2: the kernel is issue-slot bound, not FMA bound):
//   * x-direction (rows) taps live in a register window of 2*HALO+ROWS values,
//     fed by one coalesced global load per new row -> each value is read from
//     HBM/L2 once per band (+ 2*HALO rows of band overlap);
//   * y-direction (columns) taps come from a double-buffered shared-memory row
//     exchange (ROWS rows per __syncthreads), chunk-edge halos are fetched by
//     the first threads of the CTA with the periodic wrap applied to the index.
// Everything of one step (4 derivative stencils per field, upwind switch, RHS
// algebra, Euler update of both fields) happens in that single pass: 8 B per
// field-point of compulsory traffic, which is what bounds it (HBM roofline).
#include <cstdlib>
#include "common.cuh"

namespace ppnn {

enum { FINE_AS_WRITTEN = 0, FINE_INTENDED = 1, FINE_GENERIC = 2 };

struct FineParams {
    const float* u; const float* v; const float* mu;
    float* uo; float* vo;
    int H, W, PH, PW;          // points and periods along rows / columns
    int band_rows;
    int rhs_u, rhs_v;          // FINE_GENERIC only
    float af[4], ab[4];        // one-sided filters, offsets 0..a and -a..0
    float d[kMaxTaps];         // central second derivative, 2*HALO+1 taps
    float dt, rdx, rdy, rdx2, rdy2;
};

#ifndef PPNN_FINE_ROWS
#define PPNN_FINE_ROWS 4
#endif
#ifndef PPNN_FINE_MINCTA
#define PPNN_FINE_MINCTA 3
#endif
constexpr int kFineRows = PPNN_FINE_ROWS;

__device__ __forceinline__ int wrap_pt(int i, int n, int period) {
    if (i < 0) i += period; else if (i >= n) i -= period;
    return clampi(i, 0, n - 1);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot, two rows ----
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float c, float2 x, float2 acc) { return __ffma2_rn(f2(c), x, acc); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
// (c <= 0) * fwd + (c > 0) * bwd for finite operands (a NaN centre poisons both branches anyway)
__device__ __forceinline__ float sel1(float c, float f, float b) { return c <= 0.f ? f : b; }

// sign-switched one-sided derivative of both rows; the switch is the sign of `c`
// (the differentiated field itself at the output point, operators.py:255-258)
template <int HL>
__device__ __forceinline__ float2 upwind2(const FineParams& p, float2 c, const float2* w /* centre at HL */) {
    float2 f = f2(0.f), b = f2(0.f);
#pragma unroll
    for (int k = 0; k <= HL; ++k) { f = fma2(p.af[k], w[HL + k], f); b = fma2(p.ab[k], w[k], b); }
    return make_float2(sel1(c.x, f.x, b.x), sel1(c.y, f.y, b.y));
}

// backward-biased stencil alone: what upwind2 returns where the centre value is positive
template <int HL>
__device__ __forceinline__ float2 upwind2_bwd(const FineParams& p, const float2* w /* centre at HL */) {
    float2 b = f2(0.f);
#pragma unroll
    for (int k = 0; k <= HL; ++k) b = fma2(p.ab[k], w[k], b);
    return b;
}

// Warp-uniform test for the upwind fast path: every lane sees positive centres in both packed rows (of u, and of v
// when v's derivatives are needed).  Always true on the reference's data, whose initial conditions are normalised to
// [0.1, 1.1] (operators.py:436-445): then only the backward stencils are evaluated -- half the FFMA2s of the upwind
// part and no selects.  Same values as the general path (the unselected branch never contributes).
__device__ __forceinline__ bool all_positive(float2 c, float2 v, bool with_v) {
    const bool pos = c.x > 0.f && c.y > 0.f && (!with_v || (v.x > 0.f && v.y > 0.f));
    return __all_sync(__activemask(), pos);
}

template <int HL>
__device__ __forceinline__ float2 central2(const FineParams& p, const float2* w) {
    float2 s = f2(0.f);
#pragma unroll
    for (int k = 0; k <= 2 * HL; ++k) s = fma2(p.d[k], w[k], s);
    return s;
}

__device__ __forceinline__ float2 react2(float2 acc, int react, float2 u1, float2 v1) {
    return make_float2(add_reaction(acc.x, react, u1.x, v1.x), add_reaction(acc.y, react, u1.y, v1.y));
}

// One CTA = (column chunk, PAIR of row bands A/B, sample).  A thread owns one
// column and advances row i of band A and row i of band B in lockstep, the two
// rows packed in float2 so that every stencil FMA is one FFMA2.
// WC: compile-time row pitch (0 = runtime) so that the streamed accesses of a batch are
// one pointer + immediates; pairs that stay clear of the plane ends never wrap rows and
// take a branch-free addressing path.
template <int HL, int MODE, int WC>
__global__ void __launch_bounds__(384, PPNN_FINE_MINCTA)
fine_step_kernel(const FineParams p) {
    const int W = WC ? WC : p.W;
    constexpr int R = kFineRows;
    constexpr int WIN = 2 * HL + R;
    constexpr bool VWIN = MODE != FINE_AS_WRITTEN;
    constexpr int NF = VWIN ? 2 : 1;                 // fields exchanged through smem
    extern __shared__ float2 sm2[];                  // [2][NF][R][pitch] of (band A, band B)
    const int TX = blockDim.x, tid = threadIdx.x;
    const int pitch = TX + 2 * HL;
    const int b = blockIdx.z;
    const int j0 = blockIdx.x * TX;
    const int j = j0 + tid;
    const bool active = j < W;
    const int jl = wrap_pt(j, W, p.PW);            // column this thread loads (wrapped past the edge)
    const int BH = p.band_rows;                      // rows per band; the pair covers 2*BH rows
    const int rA = blockIdx.y * 2 * BH;              // first row of band A
    const int rB = rA + BH;                          // first row of band B (may be >= H: empty band)
    const int endA = min(rA + BH, p.H), endB = min(rB + BH, p.H);
    const size_t plane = (size_t)p.H * W;
    const float* ub = p.u + (size_t)b * plane;
    const float* vb = p.v + (size_t)b * plane;
    float* uob = p.uo + (size_t)b * plane;
    float* vob = p.vo + (size_t)b * plane;
    const float mu = __ldg(p.mu + b);
    // rows never wrap inside a pair that stays clear of both ends of the plane
    const bool interior = rA - HL >= 0 && rB + BH + HL <= p.H;

    // chunk-edge halo duty: thread t < R*2*HL fetches one (row, side, offset) of both bands
    const bool halo_duty = tid < R * 2 * HL;
    const int hq = tid / (2 * HL), hs = tid % (2 * HL);
    const int hcol = wrap_pt(hs < HL ? j0 - HL + hs : j0 + TX + (hs - HL), W, p.PW);
    const int hslot = hs < HL ? hs : TX + hs;        // position in the smem row

    auto rowptr = [&](const float* base, int r) -> const float* {   // row r of the periodic plane
        return base + (size_t)wrap_pt(r, p.H, p.PH) * W;
    };

    float2 wu[WIN], wv[VWIN ? WIN : 1];
#pragma unroll
    for (int k = 0; k < 2 * HL; ++k) {
        wu[k] = make_float2(__ldg(rowptr(ub, rA - HL + k) + jl), __ldg(rowptr(ub, rB - HL + k) + jl));
        if (VWIN) wv[k] = make_float2(__ldg(rowptr(vb, rA - HL + k) + jl), __ldg(rowptr(vb, rB - HL + k) + jl));
    }

    // fast-path pointers: window-feed rows (r + HL) and centre rows r of both bands, this thread's column
    const float* fuA = ub + (size_t)(rA + HL) * W + jl;
    const float* fuB = ub + (size_t)(rB + HL) * W + jl;
    const float* fvA = vb + (size_t)(rA + (VWIN ? HL : 0)) * W + jl;
    const float* fvB = vb + (size_t)(rB + (VWIN ? HL : 0)) * W + jl;
    const float* huA = ub + (size_t)(rA + hq) * W + hcol;      // halo duty: centre row rA + i + hq
    const float* huB = ub + (size_t)(rB + hq) * W + hcol;
    const float* hvA = vb + (size_t)(rA + hq) * W + hcol;
    const float* hvB = vb + (size_t)(rB + hq) * W + hcol;

    // output pointers (rows rA + i, rB + i of this thread's column; only dereferenced when in range)
    float* ouA = uob + (size_t)rA * W + j;
    float* ouB = uob + (size_t)rB * W + j;
    float* ovA = vob + (size_t)rA * W + j;
    float* ovB = vob + (size_t)rB * W + j;

    int buf = 0;
    for (int i = 0; i < BH; i += R, buf ^= 1) {
        float2 vc[R];
        float2 hu = f2(0.f), hv = f2(0.f);
        if (interior) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                wu[2 * HL + q] = make_float2(__ldg(fuA + (size_t)q * W), __ldg(fuB + (size_t)q * W));
                if (VWIN) wv[2 * HL + q] = make_float2(__ldg(fvA + (size_t)q * W), __ldg(fvB + (size_t)q * W));
                else vc[q] = make_float2(__ldg(fvA + (size_t)q * W), __ldg(fvB + (size_t)q * W));
            }
            if (halo_duty) {
                hu = make_float2(__ldg(huA), __ldg(huB));
                if (VWIN) hv = make_float2(__ldg(hvA), __ldg(hvB));
            }
            fuA += (size_t)R * W; fuB += (size_t)R * W; fvA += (size_t)R * W; fvB += (size_t)R * W;
            huA += (size_t)R * W; huB += (size_t)R * W;
            if (VWIN) { hvA += (size_t)R * W; hvB += (size_t)R * W; }
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int ra = rA + i + q + HL, rb = rB + i + q + HL;
                wu[2 * HL + q] = make_float2(__ldg(rowptr(ub, ra) + jl), __ldg(rowptr(ub, rb) + jl));
                if (VWIN) wv[2 * HL + q] = make_float2(__ldg(rowptr(vb, ra) + jl), __ldg(rowptr(vb, rb) + jl));
                else {
                    const int ca = min(rA + i + q, p.H - 1), cb = min(rB + i + q, p.H - 1);
                    vc[q] = make_float2(__ldg(vb + (size_t)ca * W + jl), __ldg(vb + (size_t)cb * W + jl));
                }
            }
            if (halo_duty) {
                const int ca = min(rA + i + hq, p.H - 1), cb = min(rB + i + hq, p.H - 1);
                hu = make_float2(__ldg(ub + (size_t)ca * W + hcol), __ldg(ub + (size_t)cb * W + hcol));
                if (VWIN) hv = make_float2(__ldg(vb + (size_t)ca * W + hcol), __ldg(vb + (size_t)cb * W + hcol));
            }
        }
        float2* su = sm2 + (size_t)buf * NF * R * pitch;
        float2* sv = su + R * pitch;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            su[q * pitch + HL + tid] = wu[HL + q];
            if (VWIN) sv[q * pitch + HL + tid] = wv[HL + q];
        }
        if (halo_duty) {
            su[hq * pitch + hslot] = hu;
            if (VWIN) sv[hq * pitch + hslot] = hv;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int ia = rA + i + q, ib = rB + i + q;
            if (ia >= endA) break;                   // band A is never shorter than band B
            float2 yu[2 * HL + 1];
#pragma unroll
            for (int k = 0; k <= 2 * HL; ++k) yu[k] = su[q * pitch + tid + k];
            const float2 c = wu[HL + q];
            float2 unew, vnew;
            if (MODE == FINE_GENERIC) {
                float2 yv[2 * HL + 1];
#pragma unroll
                for (int k = 0; k <= 2 * HL; ++k) yv[k] = sv[q * pitch + tid + k];
                const float2 cv = wv[HL + q];
                float2 ru = f2(0.f), rv = f2(0.f);
                if (p.rhs_u & PPNN_RHS_ADV)
                    ru = add2(mul2(mul2(neg2(c), upwind2<HL>(p, c, wu + q)), f2(p.rdx)),
                              neg2(mul2(mul2(cv, upwind2<HL>(p, c, yu)), f2(p.rdy))));
                if (p.rhs_u & PPNN_RHS_DIFF) {
                    float2 lap = mul2(f2(mu), add2(mul2(central2<HL>(p, wu + q), f2(p.rdx2)), mul2(central2<HL>(p, yu), f2(p.rdy2))));
                    ru = (p.rhs_u & PPNN_RHS_ADV) ? add2(ru, lap) : lap;
                }
                ru = react2(ru, p.rhs_u & ~3, c, cv);
                if (p.rhs_v & PPNN_RHS_ADV)
                    rv = add2(mul2(mul2(neg2(c), upwind2<HL>(p, cv, wv + q)), f2(p.rdx)),
                              neg2(mul2(mul2(cv, upwind2<HL>(p, cv, yv)), f2(p.rdy))));
                if (p.rhs_v & PPNN_RHS_DIFF) {
                    float2 lap = mul2(f2(mu), add2(mul2(central2<HL>(p, wv + q), f2(p.rdx2)), mul2(central2<HL>(p, yv), f2(p.rdy2))));
                    rv = (p.rhs_v & PPNN_RHS_ADV) ? add2(rv, lap) : lap;
                }
                rv = react2(rv, p.rhs_v & ~3, c, cv);
                unew = fma2(p.dt, ru, c);
                vnew = fma2(p.dt, rv, cv);
            } else {
                // burgers2Dpu then burgers2Dpv with the updated u (operators.py:474-475)
                const float2 dxu = upwind2<HL>(p, c, wu + q), dyu = upwind2<HL>(p, c, yu);
                const float2 lapu = mul2(f2(mu), add2(mul2(central2<HL>(p, wu + q), f2(p.rdx2)), mul2(central2<HL>(p, yu), f2(p.rdy2))));
                const float2 vcen = VWIN ? wv[HL + q] : vc[q];
                const float2 advy = neg2(mul2(mul2(vcen, dyu), f2(p.rdy)));
                unew = fma2(p.dt, add2(add2(mul2(mul2(neg2(c), dxu), f2(p.rdx)), advy), lapu), c);
                if (MODE == FINE_AS_WRITTEN) {
                    // quirk Q3: v is advected and diffused with u's derivative fields
                    vnew = fma2(p.dt, add2(add2(mul2(mul2(neg2(unew), dxu), f2(p.rdx)), advy), lapu), vcen);
                } else {
                    float2 yv[2 * HL + 1];
#pragma unroll
                    for (int k = 0; k <= 2 * HL; ++k) yv[k] = sv[q * pitch + tid + k];
                    const float2 dxv = upwind2<HL>(p, vcen, wv + q), dyv = upwind2<HL>(p, vcen, yv);
                    const float2 lapv = mul2(f2(mu), add2(mul2(central2<HL>(p, wv + q), f2(p.rdx2)), mul2(central2<HL>(p, yv), f2(p.rdy2))));
                    vnew = fma2(p.dt, add2(add2(mul2(mul2(neg2(unew), dxv), f2(p.rdx)),
                                                neg2(mul2(mul2(vcen, dyv), f2(p.rdy)))), lapv), vcen);
                }
            }
            if (active) {
                ouA[(size_t)q * W] = unew.x;
                ovA[(size_t)q * W] = vnew.x;
                if (ib < endB) {
                    ouB[(size_t)q * W] = unew.y;
                    ovB[(size_t)q * W] = vnew.y;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2 * HL; ++k) {
            wu[k] = wu[k + R];
            if (VWIN) wv[k] = wv[k + R];
        }
        ouA += (size_t)R * W; ouB += (size_t)R * W; ovA += (size_t)R * W; ovB += (size_t)R * W;
    }
}

}  // namespace ppnn
#include "fine_step_tile.cuh"
#include "fine_step_ep.cuh"
namespace ppnn {

template <int HL, int MODE>
static int launch_fine(const FineParams& p, int B, void* stream) {
    {
        bool used = false;
        int st = launch_fine_tile<HL, MODE>(p, B, stream, &used);     // true-periodic meshes with aligned rows
        if (used || st != PPNN_OK) return st;
    }
    {
        bool used = false;                                            // duplicated endpoint / odd widths up to 1088 columns
        int st;
        if (HL == 3 && p.W == 1025) st = launch_fine_ep<HL, MODE, 1025>(p, B, stream, &used);
        else if (HL == 3 && p.W == 257) st = launch_fine_ep<HL, MODE, 257>(p, B, stream, &used);
        else st = launch_fine_ep<HL, MODE, 0>(p, B, stream, &used);
        if (used || st != PPNN_OK) return st;
    }
    // columns per CTA: whole warps, as few idle lanes as possible, <= 384 threads
    int warps_total = (p.W + 31) / 32;
    int nchunk = (warps_total + 11) / 12;
    int wpc = (warps_total + nchunk - 1) / nchunk;
    if (wpc < 1) wpc = 1;
    int TX = wpc * 32;
    if (TX < kFineRows * 2 * HL) TX = ((kFineRows * 2 * HL + 31) / 32) * 32;
    nchunk = (p.W + TX - 1) / TX;
    // band pairs: enough CTAs for ~4 waves of 148 SMs x 3 resident CTAs, bands >= 16 rows
    long want = 148L * 3 * 4;
    long per_pair_ctas = (long)nchunk * B;
    int npair = (int)((want + per_pair_ctas - 1) / per_pair_ctas);
    int max_pair = (p.H + 31) / 32;
    if (npair > max_pair) npair = max_pair;
    if (npair < 1) npair = 1;
    int band_rows = (p.H + 2 * npair - 1) / (2 * npair);
    band_rows = (band_rows + kFineRows - 1) / kFineRows * kFineRows;
    npair = (p.H + 2 * band_rows - 1) / (2 * band_rows);
    FineParams q = p;
    q.band_rows = band_rows;
    constexpr int NF = MODE == FINE_AS_WRITTEN ? 1 : 2;
    size_t smem = (size_t)2 * NF * kFineRows * (TX + 2 * HL) * sizeof(float2);
    dim3 grid(nchunk, npair, B);
    if (npair > 65535 || B > 65535) return PPNN_E_SHAPE;
    void (*kern)(FineParams) = fine_step_kernel<HL, MODE, 0>;
    if (HL == 3) {                       // the generator's stencil: specialise the usual mesh widths
        if (p.W == 257) kern = fine_step_kernel<HL == 3 ? 3 : HL, MODE, 257>;
        else if (p.W == 1025) kern = fine_step_kernel<HL == 3 ? 3 : HL, MODE, 1025>;
        else if (p.W == 1024) kern = fine_step_kernel<HL == 3 ? 3 : HL, MODE, 1024>;
        else if (p.W == 256) kern = fine_step_kernel<HL == 3 ? 3 : HL, MODE, 256>;
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<grid, TX, smem, (cudaStream_t)stream>>>(q);
    return launch_status();
}

template <int MODE>
static int dispatch_fine(FineParams& p, int B, int adv_accuracy, int diff_accuracy, void* stream) {
    Taps f, bk, d;
    if (!make_taps(PPNN_FORWARD1, adv_accuracy, &f, nullptr) ||
        !make_taps(PPNN_BACKWARD1, adv_accuracy, &bk, nullptr) ||
        !make_taps(PPNN_CENTRAL2, diff_accuracy, &d, nullptr))
        return PPNN_E_SCHEME;
    int halo = -d.off;
    if (adv_accuracy != halo) return PPNN_E_HALO;       // the generator pads ONE halo for both stencils
    if (p.PH <= 2 * halo || p.PW <= 2 * halo) return PPNN_E_SHAPE;
    for (int k = 0; k < 4; ++k) { p.af[k] = k < f.n ? f.c[k] : 0.f; p.ab[k] = k < bk.n ? bk.c[k] : 0.f; }
    for (int k = 0; k < kMaxTaps; ++k) p.d[k] = d.c[k];
    switch (halo) {
    case 1: return launch_fine<1, MODE>(p, B, stream);
    case 2: return launch_fine<2, MODE>(p, B, stream);
    case 3: return launch_fine<3, MODE>(p, B, stream);
    }
    return PPNN_E_SCHEME;
}

static int fill_fine(FineParams* p, const float* u, const float* v, const float* mu, float* uo, float* vo,
                     int B, int H, int W, int endpoint, float dt, float dx, float dy) {
    if (!u || !v || !mu || !uo || !vo) return PPNN_E_NULL;
    if (B < 0 || H < 2 || W < 2) return PPNN_E_SHAPE;
    if (u == uo || v == vo || u == vo || v == uo) return PPNN_E_MODE;
    p->u = u; p->v = v; p->mu = mu; p->uo = uo; p->vo = vo;
    p->H = H; p->W = W;
    p->PH = endpoint ? H - 1 : H;
    p->PW = endpoint ? W - 1 : W;
    p->band_rows = H;
    p->rhs_u = p->rhs_v = 0;
    p->dt = dt;
    // torch on CUDA divides by a Python scalar as a multiply by its fp32 reciprocal
    p->rdx = 1.f / dx; p->rdy = 1.f / dy;
    p->rdx2 = 1.f / (float)((double)dx * (double)dx);
    p->rdy2 = 1.f / (float)((double)dy * (double)dy);
    return PPNN_OK;
}

}  // namespace ppnn

using namespace ppnn;

#ifdef PPNN_TRACE
extern "C" int ppnn_trace_read_fine(unsigned long long* host, long n) {
    return (int)cudaMemcpyFromSymbol(host, ppnn::g_trace_fine, (size_t)n * sizeof(unsigned long long));
}
#endif

extern "C" int ppnn_fine_burgers_step(const float* u, const float* v, const float* mu,
                                      float* u_out, float* v_out, int B, int H, int W,
                                      int endpoint, int mode, int adv_accuracy, int diff_accuracy,
                                      float dt, float dx, float dy, void* stream) {
    FineParams p;
    int st = fill_fine(&p, u, v, mu, u_out, v_out, B, H, W, endpoint, dt, dx, dy);
    if (st != PPNN_OK) return st;
    if (B == 0) return PPNN_OK;
    if (mode == PPNN_FINE_AS_WRITTEN) return dispatch_fine<FINE_AS_WRITTEN>(p, B, adv_accuracy, diff_accuracy, stream);
    if (mode == PPNN_FINE_INTENDED) return dispatch_fine<FINE_INTENDED>(p, B, adv_accuracy, diff_accuracy, stream);
    return PPNN_E_MODE;
}

extern "C" int ppnn_fine_rhs_step(const float* u, const float* v, const float* mu,
                                  float* u_out, float* v_out, int B, int H, int W,
                                  int endpoint, int rhs_u, int rhs_v,
                                  int adv_accuracy, int diff_accuracy,
                                  float dt, float dx, float dy, void* stream) {
    FineParams p;
    int st = fill_fine(&p, u, v, mu, u_out, v_out, B, H, W, endpoint, dt, dx, dy);
    if (st != PPNN_OK) return st;
    if (!valid_rhs_id(rhs_u) || !valid_rhs_id(rhs_v)) return PPNN_E_RHS;
    if (B == 0) return PPNN_OK;
    p.rhs_u = rhs_u; p.rhs_v = rhs_v;
    return dispatch_fine<FINE_GENERIC>(p, B, adv_accuracy, diff_accuracy, stream);
}

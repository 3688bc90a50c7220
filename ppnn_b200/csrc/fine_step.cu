// Fine-grid explicit solver: one fused Euler step of the 2-D Burgers dataset
// generator (src/operators.py:416-420, 466-475) and of the generic
// f <- f + dt*rhs(f) generators (operators.py:494-748), on periodic grids with
// or without the duplicated endpoint.
//
// Layout: u, v are [B, H, W] fp32, W contiguous.  One CTA = (32k-column chunk,
// row band, sample).  A thread owns one column and walks down the band:
//   * x-direction (rows) taps live in a register window of 2*HALO+ROWS values,
//     fed by one coalesced global load per new row -> each value is read from
//     HBM/L2 once per band (+ 2*HALO rows of band overlap);
//   * y-direction (columns) taps come from a double-buffered shared-memory row
//     exchange (ROWS rows per __syncthreads), chunk-edge halos are fetched by
//     the first threads of the CTA with the periodic wrap applied to the index.
// Everything of one step (4 derivative stencils per field, upwind switch, RHS
// algebra, Euler update of both fields) happens in that single pass: 8 B per
// field-point of compulsory traffic, which is what bounds it (HBM roofline).
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

constexpr int kFineRows = 4;

__device__ __forceinline__ int wrap_pt(int i, int n, int period) {
    if (i < 0) i += period; else if (i >= n) i -= period;
    return clampi(i, 0, n - 1);
}

template <int HL>
__device__ __forceinline__ float upwind_sel(const FineParams& p, float c, const float* w /* w[0..2HL], centre at HL */) {
    float f = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k <= HL; ++k) { f += p.af[k] * w[HL + k]; b += p.ab[k] * w[k]; }
    return c <= 0.f ? f : (c > 0.f ? b : 0.f);
}

template <int HL>
__device__ __forceinline__ float central_sum(const FineParams& p, const float* w) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k <= 2 * HL; ++k) s += p.d[k] * w[k];
    return s;
}

template <int HL, int MODE>
__global__ void __launch_bounds__(384)
fine_step_kernel(const FineParams p) {
    constexpr int R = kFineRows;
    constexpr int WIN = 2 * HL + R;
    constexpr bool VWIN = MODE != FINE_AS_WRITTEN;
    constexpr int NF = VWIN ? 2 : 1;                 // fields exchanged through smem
    extern __shared__ float sm[];                    // [2][NF][R][pitch]
    const int TX = blockDim.x, tid = threadIdx.x;
    const int pitch = TX + 2 * HL;
    const int b = blockIdx.z;
    const int j0 = blockIdx.x * TX;
    const int j = j0 + tid;
    const bool active = j < p.W;
    const int jl = wrap_pt(j, p.W, p.PW);            // column this thread loads (wrapped past the edge)
    const int r0 = blockIdx.y * p.band_rows;
    const int r1 = min(r0 + p.band_rows, p.H);
    const size_t plane = (size_t)p.H * p.W;
    const float* ub = p.u + (size_t)b * plane;
    const float* vb = p.v + (size_t)b * plane;
    float* uob = p.uo + (size_t)b * plane;
    float* vob = p.vo + (size_t)b * plane;
    const float mu = __ldg(p.mu + b);

    // chunk-edge halo duty: thread t < R*2*HL fetches one (row, side, offset)
    const bool halo_duty = tid < R * 2 * HL;
    const int hq = tid / (2 * HL), hs = tid % (2 * HL);
    const int hcol = wrap_pt(hs < HL ? j0 - HL + hs : j0 + TX + (hs - HL), p.W, p.PW);
    const int hslot = hs < HL ? hs : TX + hs;        // position in the smem row

    float wu[WIN], wv[VWIN ? WIN : 1];
#pragma unroll
    for (int k = 0; k < 2 * HL; ++k) {
        int r = wrap_pt(r0 - HL + k, p.H, p.PH);
        wu[k] = __ldg(ub + (size_t)r * p.W + jl);
        if (VWIN) wv[k] = __ldg(vb + (size_t)r * p.W + jl);
    }

    int buf = 0;
    for (int i = r0; i < r1; i += R, buf ^= 1) {
        float vc[R];
        float hu = 0.f, hv = 0.f;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            int r = wrap_pt(i + q + HL, p.H, p.PH);
            wu[2 * HL + q] = __ldg(ub + (size_t)r * p.W + jl);
            if (VWIN) wv[2 * HL + q] = __ldg(vb + (size_t)r * p.W + jl);
            else vc[q] = (active && i + q < r1) ? __ldg(vb + (size_t)(i + q) * p.W + j) : 0.f;
        }
        if (halo_duty) {
            int r = min(i + hq, p.H - 1);
            hu = __ldg(ub + (size_t)r * p.W + hcol);
            if (VWIN) hv = __ldg(vb + (size_t)r * p.W + hcol);
        }
        float* su = sm + (size_t)buf * NF * R * pitch;
        float* sv = su + R * pitch;
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
            int ii = i + q;
            if (ii >= r1) break;
            float yu[2 * HL + 1];
#pragma unroll
            for (int k = 0; k <= 2 * HL; ++k) yu[k] = su[q * pitch + tid + k];
            const float c = wu[HL + q];
            float unew, vnew;
            if (MODE == FINE_GENERIC) {
                float yv[2 * HL + 1];
#pragma unroll
                for (int k = 0; k <= 2 * HL; ++k) yv[k] = sv[q * pitch + tid + k];
                const float cv = wv[HL + q];
                float ru = 0.f, rv = 0.f;
                if (p.rhs_u & PPNN_RHS_ADV)
                    ru = (-c * upwind_sel<HL>(p, c, wu + q)) * p.rdx - (cv * upwind_sel<HL>(p, c, yu)) * p.rdy;
                if (p.rhs_u & PPNN_RHS_DIFF) {
                    float lap = mu * (central_sum<HL>(p, wu + q) * p.rdx2 + central_sum<HL>(p, yu) * p.rdy2);
                    ru = (p.rhs_u & PPNN_RHS_ADV) ? ru + lap : lap;
                }
                ru = add_reaction(ru, p.rhs_u & ~3, c, cv);
                if (p.rhs_v & PPNN_RHS_ADV)
                    rv = (-c * upwind_sel<HL>(p, cv, wv + q)) * p.rdx - (cv * upwind_sel<HL>(p, cv, yv)) * p.rdy;
                if (p.rhs_v & PPNN_RHS_DIFF) {
                    float lap = mu * (central_sum<HL>(p, wv + q) * p.rdx2 + central_sum<HL>(p, yv) * p.rdy2);
                    rv = (p.rhs_v & PPNN_RHS_ADV) ? rv + lap : lap;
                }
                rv = add_reaction(rv, p.rhs_v & ~3, c, cv);
                unew = c + p.dt * ru;
                vnew = cv + p.dt * rv;
            } else {
                // burgers2Dpu then burgers2Dpv with the updated u (operators.py:474-475)
                const float dxu = upwind_sel<HL>(p, c, wu + q), dyu = upwind_sel<HL>(p, c, yu);
                const float lapu = mu * (central_sum<HL>(p, wu + q) * p.rdx2 + central_sum<HL>(p, yu) * p.rdy2);
                const float vcen = VWIN ? wv[HL + q] : vc[q];
                unew = c + p.dt * ((-c * dxu) * p.rdx - (vcen * dyu) * p.rdy + lapu);
                if (MODE == FINE_AS_WRITTEN) {
                    // quirk Q3: v is advected and diffused with u's derivative fields
                    vnew = vcen + p.dt * ((-unew * dxu) * p.rdx - (vcen * dyu) * p.rdy + lapu);
                } else {
                    float yv[2 * HL + 1];
#pragma unroll
                    for (int k = 0; k <= 2 * HL; ++k) yv[k] = sv[q * pitch + tid + k];
                    const float dxv = upwind_sel<HL>(p, vcen, wv + q), dyv = upwind_sel<HL>(p, vcen, yv);
                    const float lapv = mu * (central_sum<HL>(p, wv + q) * p.rdx2 + central_sum<HL>(p, yv) * p.rdy2);
                    vnew = vcen + p.dt * ((-unew * dxv) * p.rdx - (vcen * dyv) * p.rdy + lapv);
                }
            }
            if (active) {
                uob[(size_t)ii * p.W + j] = unew;
                vob[(size_t)ii * p.W + j] = vnew;
            }
        }
#pragma unroll
        for (int k = 0; k < 2 * HL; ++k) {
            wu[k] = wu[k + R];
            if (VWIN) wv[k] = wv[k + R];
        }
    }
}

template <int HL, int MODE>
static int launch_fine(const FineParams& p, int B, void* stream) {
    // columns per CTA: whole warps, as few idle lanes as possible, <= 384 threads
    int warps_total = (p.W + 31) / 32;
    int nchunk = (warps_total + 11) / 12;
    int wpc = (warps_total + nchunk - 1) / nchunk;
    if (wpc < 1) wpc = 1;
    int TX = wpc * 32;
    if (TX < kFineRows * 2 * HL) TX = ((kFineRows * 2 * HL + 31) / 32) * 32;
    nchunk = (p.W + TX - 1) / TX;
    // row bands: enough CTAs for ~4 waves of 148 SMs x 3 resident CTAs, bands >= 32 rows
    long want = 148L * 3 * 4;
    long per_band_ctas = (long)nchunk * B;
    int nband = (int)((want + per_band_ctas - 1) / per_band_ctas);
    int max_band = (p.H + 31) / 32;
    if (nband > max_band) nband = max_band;
    if (nband < 1) nband = 1;
    int band_rows = (p.H + nband - 1) / nband;
    band_rows = (band_rows + kFineRows - 1) / kFineRows * kFineRows;
    nband = (p.H + band_rows - 1) / band_rows;
    FineParams q = p;
    q.band_rows = band_rows;
    constexpr int NF = MODE == FINE_AS_WRITTEN ? 1 : 2;
    size_t smem = (size_t)2 * NF * kFineRows * (TX + 2 * HL) * sizeof(float);
    dim3 grid(nchunk, nband, B);
    if (nband > 65535 || B > 65535) return PPNN_E_SHAPE;
    fine_step_kernel<HL, MODE><<<grid, TX, smem, (cudaStream_t)stream>>>(q);
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

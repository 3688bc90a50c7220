// TMA-fed variant of the fine-grid solver step (included by fine_step.cu).
//
// Persistent CTAs, one contiguous range of (sample, row) pairs each, so every SM
// gets the same number of rows whatever the batch size.  A producer warp streams
// 4-row bands of BOTH fields through a shared-memory ring with `cp.async.bulk`
// (one copy per band and field; bands that wrap around the periodic plane, or that
// end the tensor, go row by row: 16-byte-aligned interior by TMA, the <= 3 edge
// floats by ordinary loads, so nothing outside a row is ever read).  Consumer
// warps own two columns per lane (t and t + ceil(W/2), packed in float2 so each
// stencil FMA is one FFMA2); the x-direction (row) taps live in a register window
// fed from the ring, the y-direction (column) taps are read straight from the ring
// row, the periodic column wrap being index arithmetic inside that row.  No
// __syncthreads, no per-batch global-load latency: completion is signalled on
// mbarriers (full / empty per ring slot).
//
// Needs (period * W) % 4 == 0 so that all rows of a band keep one 16-byte phase.
#pragma once
#include <type_traits>
#include "tma.cuh"

namespace ppnn {

constexpr int kTmaRows = 4;            // rows per ring band = rows per compute batch
constexpr int kTmaMaxRing = 8;

struct FineTmaPlan {
    long total_rows;       // B * H
    long rows_per_cta;
    int lanes;             // ceil(W / 2): lane t owns columns t and t + lanes
    int cwarps;            // consumer warps
    int ring;              // ring slots
    int band_floats;       // floats per (slot, field): 4 * W + 8, multiple of 4
    const float* u_end;    // one past the last element of u / v (bounds of the tensors)
    const float* v_end;
};

template <int HL, int MODE, int WC>
__global__ void __launch_bounds__(576)
fine_step_tma_kernel(const FineParams p, const FineTmaPlan pl) {
    static_assert(HL <= 3, "band geometry assumes a halo of at most 3 rows");
    constexpr int R = kTmaRows;
    constexpr bool VWIN = MODE != FINE_AS_WRITTEN;
    const int W = WC ? WC : p.W;
    extern __shared__ float4 smem4[];
    float* ringbuf = reinterpret_cast<float*>(smem4);            // [ring][2][band_floats]
    __shared__ __align__(8) uint64_t full[kTmaMaxRing];
    __shared__ __align__(8) uint64_t empty[kTmaMaxRing];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == pl.cwarps;
    const size_t plane = (size_t)p.H * W;
    const long g0 = (long)blockIdx.x * pl.rows_per_cta;
    const long g1 = min(g0 + pl.rows_per_cta, pl.total_rows);

    if (tid == 0) {
        for (int s = 0; s < pl.ring; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], pl.cwarps); }
        fence_barrier_init();
    }
    __syncthreads();

    auto band_ptr = [&](int slot, int f) -> float* { return ringbuf + ((size_t)slot * 2 + f) * pl.band_floats; };
    // 16-byte phase (in floats) of plane row k of a field plane
    auto lead_of = [&](const float* fplane, int k) -> int {
        return (int)((reinterpret_cast<uintptr_t>(fplane + (size_t)k * W) >> 2) & 3);
    };

    if (producer) {
        // ------------------------------ producer warp ------------------------------
        long cidx = 0;
        int pslot = 0; uint32_t pphase = 0;       // slot / parity of use of band `cidx`
        for (long g = g0; g < g1;) {
            const int b = (int)(g / p.H), r0 = (int)(g % p.H);
            const int r1 = (int)min((long)p.H, r0 + (g1 - g));
            const int nb = (r1 - r0 + R - 1) / R;
            const float* fpl[2] = {p.u + (size_t)b * plane, p.v + (size_t)b * plane};
            const float* fend[2] = {pl.u_end, pl.v_end};
            for (int c = 0; c < nb + 2; ++c, ++cidx) {
                const int slot = pslot;
                if (cidx >= pl.ring) mbar_wait(&empty[slot], pphase ^ 1);   // previous use of the slot was released
                if (++pslot == pl.ring) { pslot = 0; pphase ^= 1; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const int kf = r0 - R + c * R;                       // first (unwrapped) plane row of the band
                const int k0 = wrap_pt(kf, p.H, p.PH);
                bool contiguous = kf >= 0 && kf + R <= p.H;
                uint32_t nbytes[2] = {0, 0};
                uintptr_t a0[2] = {0, 0};
                if (contiguous) {
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const uintptr_t a = reinterpret_cast<uintptr_t>(fpl[f] + (size_t)kf * W);
                        a0[f] = a & ~(uintptr_t)15;
                        nbytes[f] = (uint32_t)((a - a0[f]) + (size_t)R * W * 4 + 15) & ~15u;
                        if (a0[f] + nbytes[f] > (reinterpret_cast<uintptr_t>(fend[f]) & ~(uintptr_t)15)) contiguous = false;
                    }
                }
                if (contiguous) {
                    if (lane == 0) {
                        mbar_expect_tx(&full[slot], nbytes[0] + nbytes[1]);
                        bulk_g2s(band_ptr(slot, 0), reinterpret_cast<const void*>(a0[0]), nbytes[0], &full[slot]);
                        bulk_g2s(band_ptr(slot, 1), reinterpret_cast<const void*>(a0[1]), nbytes[1], &full[slot]);
                    }
                } else {
                    // row by row: aligned interior by TMA, the <= 3 + 3 edge floats of each row by lanes
                    uint32_t total = 0;
#pragma unroll
                    for (int f = 0; f < 2; ++f) {
                        const int lead = lead_of(fpl[f], k0);
                        for (int j = 0; j < R; ++j) {
                            const int k = wrap_pt(kf + j, p.H, p.PH);
                            const float* src = fpl[f] + (size_t)k * W;
                            float* dst = band_ptr(slot, f) + lead + j * W;
                            const uintptr_t a = reinterpret_cast<uintptr_t>(src);
                            const int head = (int)(((16 - (a & 15)) & 15) >> 2);          // floats before the aligned interior
                            const int body = ((W - head) >> 2) << 2;                      // floats copied by TMA
                            const int tail = W - head - body;
                            if (lane < head) dst[lane] = __ldg(src + lane);
                            if (lane >= 4 && lane - 4 < tail) dst[head + body + lane - 4] = __ldg(src + head + body + lane - 4);
                            total += (uint32_t)body * 4;
                        }
                    }
                    __syncwarp();                                    // edge stores ordered before the arrive
                    if (lane == 0) {
                        mbar_expect_tx(&full[slot], total);
#pragma unroll
                        for (int f = 0; f < 2; ++f) {
                            const int lead = lead_of(fpl[f], k0);
                            for (int j = 0; j < R; ++j) {
                                const int k = wrap_pt(kf + j, p.H, p.PH);
                                const float* src = fpl[f] + (size_t)k * W;
                                float* dst = band_ptr(slot, f) + lead + j * W;
                                const uintptr_t a = reinterpret_cast<uintptr_t>(src);
                                const int head = (int)(((16 - (a & 15)) & 15) >> 2);
                                const int body = ((W - head) >> 2) << 2;
                                if (body) bulk_g2s(dst + head, src + head, (uint32_t)body * 4, &full[slot]);
                            }
                        }
                    }
                }
                __syncwarp();
            }
            g += r1 - r0;
        }
        return;
    }

    // -------------------------------- consumer warps --------------------------------
    const bool active = tid < pl.lanes;
    const int colA = active ? tid : 0;
    const int cB = tid + pl.lanes;
    const bool hasB = active && cB < W;
    const int colB = hasB ? cB : colA;
    // lanes whose y-direction taps cross the periodic seam take (warp-uniformly) the path
    // with wrapped tap columns held in registers; everybody else uses immediates
    const bool seam = colA < HL || colA + HL >= W || colB < HL || colB + HL >= W;
    const bool seam_warp = __any_sync(0xffffffffu, seam);
    int ja[2 * HL + 1], jb[2 * HL + 1];
#pragma unroll
    for (int k = 0; k <= 2 * HL; ++k) {
        ja[k] = wrap_pt(colA + k - HL, W, p.PW) - colA;
        jb[k] = wrap_pt(colB + k - HL, W, p.PW) - colB;
    }

    const int slot_floats = 2 * pl.band_floats;
    int slot = 0;            // ring slot of the current segment's band 0 ... advanced band by band
    uint32_t phase = 0;      // parity of `slot`'s current use
    auto next_slot = [&](int& sl, uint32_t& ph) { if (++sl == pl.ring) { sl = 0; ph ^= 1; } };

    for (long g = g0; g < g1;) {
        const int b = (int)(g / p.H), r0 = (int)(g % p.H);
        const int r1 = (int)min((long)p.H, r0 + (g1 - g));
        const int nb = (r1 - r0 + R - 1) / R;
        const float* upl = p.u + (size_t)b * plane;
        const float* vpl = p.v + (size_t)b * plane;
        float* ouA = p.uo + (size_t)b * plane + (size_t)r0 * W + colA;
        float* ovA = p.vo + (size_t)b * plane + (size_t)r0 * W + colA;
        const int dB = colB - colA;
        const float mu = __ldg(p.mu + b);
        // R*W and period*W are multiples of 4 floats: every band of this plane has the same 16-byte phase
        const int leadU = lead_of(upl, wrap_pt(r0 - R, p.H, p.PH));
        const int leadV = lead_of(vpl, wrap_pt(r0 - R, p.H, p.PH)) + pl.band_floats;

        // bands 0 and 1: window rows r0-3 .. r0+2 (band 0 rows 1..3, band 1 rows 0..2)
        int s0 = slot; uint32_t ph0 = phase;
        int s1 = s0; uint32_t ph1 = ph0; next_slot(s1, ph1);
        mbar_wait(&full[s0], ph0);
        mbar_wait(&full[s1], ph1);
        float2 wu[2 * HL + R], wv[VWIN ? 2 * HL + R : 1];
        {
            const float* b0 = ringbuf + (size_t)s0 * slot_floats;
            const float* b1 = ringbuf + (size_t)s1 * slot_floats;
#pragma unroll
            for (int k = 0; k < HL; ++k) {
                const float* ua = b0 + leadU + (R - HL + k) * W + colA;
                const float* ub2 = b1 + leadU + k * W + colA;
                wu[k] = make_float2(ua[0], ua[dB]);
                wu[HL + k] = make_float2(ub2[0], ub2[dB]);
                if (VWIN) {
                    const float* va = b0 + leadV + (R - HL + k) * W + colA;
                    const float* vb2 = b1 + leadV + k * W + colA;
                    wv[k] = make_float2(va[0], va[dB]);
                    wv[HL + k] = make_float2(vb2[0], vb2[dB]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s0]);          // the leading band is consumed
        // sc = slot of the band holding this batch's centre rows, sn = the band after it
        int sc = s1; uint32_t phc = ph1;
        int sn = sc; uint32_t phn = phc; next_slot(sn, phn);

        auto batch = [&](auto seam_tag, int n) {
            constexpr bool SEAM = decltype(seam_tag)::value;
            const float* cb = ringbuf + (size_t)sc * slot_floats + colA;     // centre band, this lane's column A
            const float* nbp = ringbuf + (size_t)sn * slot_floats + colA;    // next band (feed rows)
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int jr = q + HL;
                const float* fu = (jr < R ? cb + jr * W : nbp + (jr - R) * W) + leadU;
                wu[2 * HL + q] = make_float2(fu[0], fu[dB]);
                if (VWIN) {
                    const float* fv = (jr < R ? cb + jr * W : nbp + (jr - R) * W) + leadV;
                    wv[2 * HL + q] = make_float2(fv[0], fv[dB]);
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const float* ur = cb + leadU + q * W;          // centre row, column A
                const float* vr = cb + leadV + q * W;
                float2 yu[2 * HL + 1], yv[VWIN ? 2 * HL + 1 : 1];
#pragma unroll
                for (int k = 0; k <= 2 * HL; ++k) {
                    if (SEAM) {
                        yu[k] = make_float2(ur[ja[k]], ur[dB + jb[k]]);
                        if (VWIN) yv[k] = make_float2(vr[ja[k]], vr[dB + jb[k]]);
                    } else {
                        yu[k] = make_float2(ur[k - HL], ur[dB + k - HL]);
                        if (VWIN) yv[k] = make_float2(vr[k - HL], vr[dB + k - HL]);
                    }
                }
                const float2 c = wu[HL + q];
                const float2 vcen = VWIN ? wv[HL + q] : make_float2(vr[0], vr[dB]);
                float2 unew, vnew;
                if (MODE == FINE_GENERIC) {
                    float2 ru = f2(0.f), rv = f2(0.f);
                    if (p.rhs_u & PPNN_RHS_ADV)
                        ru = add2(mul2(mul2(neg2(c), upwind2<HL>(p, c, wu + q)), f2(p.rdx)),
                                  neg2(mul2(mul2(vcen, upwind2<HL>(p, c, yu)), f2(p.rdy))));
                    if (p.rhs_u & PPNN_RHS_DIFF) {
                        float2 lap = mul2(f2(mu), add2(mul2(central2<HL>(p, wu + q), f2(p.rdx2)), mul2(central2<HL>(p, yu), f2(p.rdy2))));
                        ru = (p.rhs_u & PPNN_RHS_ADV) ? add2(ru, lap) : lap;
                    }
                    ru = react2(ru, p.rhs_u & ~3, c, vcen);
                    if (p.rhs_v & PPNN_RHS_ADV)
                        rv = add2(mul2(mul2(neg2(c), upwind2<HL>(p, vcen, wv + q)), f2(p.rdx)),
                                  neg2(mul2(mul2(vcen, upwind2<HL>(p, vcen, yv)), f2(p.rdy))));
                    if (p.rhs_v & PPNN_RHS_DIFF) {
                        float2 lap = mul2(f2(mu), add2(mul2(central2<HL>(p, wv + q), f2(p.rdx2)), mul2(central2<HL>(p, yv), f2(p.rdy2))));
                        rv = (p.rhs_v & PPNN_RHS_ADV) ? add2(rv, lap) : lap;
                    }
                    rv = react2(rv, p.rhs_v & ~3, c, vcen);
                    unew = fma2(p.dt, ru, c);
                    vnew = fma2(p.dt, rv, vcen);
                } else {
                    // burgers2Dpu then burgers2Dpv with the updated u (operators.py:474-475)
                    const float2 dxu = upwind2<HL>(p, c, wu + q), dyu = upwind2<HL>(p, c, yu);
                    const float2 lapu = mul2(f2(mu), add2(mul2(central2<HL>(p, wu + q), f2(p.rdx2)), mul2(central2<HL>(p, yu), f2(p.rdy2))));
                    const float2 advy = neg2(mul2(mul2(vcen, dyu), f2(p.rdy)));
                    unew = fma2(p.dt, add2(add2(mul2(mul2(neg2(c), dxu), f2(p.rdx)), advy), lapu), c);
                    if (MODE == FINE_AS_WRITTEN) {
                        // quirk Q3: v is advected and diffused with u's derivative fields
                        vnew = fma2(p.dt, add2(add2(mul2(mul2(neg2(unew), dxu), f2(p.rdx)), advy), lapu), vcen);
                    } else {
                        const float2 dxv = upwind2<HL>(p, vcen, wv + q), dyv = upwind2<HL>(p, vcen, yv);
                        const float2 lapv = mul2(f2(mu), add2(mul2(central2<HL>(p, wv + q), f2(p.rdx2)), mul2(central2<HL>(p, yv), f2(p.rdy2))));
                        vnew = fma2(p.dt, add2(add2(mul2(mul2(neg2(unew), dxv), f2(p.rdx)),
                                                    neg2(mul2(mul2(vcen, dyv), f2(p.rdy)))), lapv), vcen);
                    }
                }
                if (active && r0 + n * R + q < r1) {              // rows past r1 of a ragged last batch are dropped
                    ouA[(size_t)q * W] = unew.x; ovA[(size_t)q * W] = vnew.x;
                    if (hasB) { ouA[(size_t)q * W + dB] = unew.y; ovA[(size_t)q * W + dB] = vnew.y; }
                }
            }
#pragma unroll
            for (int k = 0; k < 2 * HL; ++k) {
                wu[k] = wu[k + R];
                if (VWIN) wv[k] = wv[k + R];
            }
        };

        for (int n = 0; n < nb; ++n) {
            mbar_wait(&full[sn], phn);                    // band n+2
            if (seam_warp) batch(std::true_type{}, n); else batch(std::false_type{}, n);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[sc]);      // centre band n+1 is consumed
            sc = sn; phc = phn;
            next_slot(sn, phn);
            ouA += (size_t)R * W; ovA += (size_t)R * W;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[sc]);          // the trailing band (feed rows only)
        slot = sn; phase = phn;                           // first band of the next segment
        g += r1 - r0;
    }
}

}  // namespace ppnn

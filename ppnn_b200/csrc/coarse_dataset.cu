// Dataset-side precompute of the PPNN training set (src/train2D.py:88-90,109-151,
// class `myset`): for every sample (u_t, u_{t+1}, mu)
//
//     pdeu  = pde_du(u_t, mu)                              train2D.py:121
//     label = (u_{t+1} - u_t) - pdeu                       train2D.py:90,129
//
// the per-channel mean / std of  u_t[:, :, :-1, :-1],  pdeu (FULL 257 x 257, as the
// reference takes them, :126-127)  and  label[:, :, :-1, :-1],  and the three
// normalised, cropped training arrays u0_normd / pdeu_normd / du_normd.
//
// The reference does this with a .to(device) ... .cpu() round trip of the whole
// 2 x 5.7 GB data set and ~10 full-size temporaries.  Here the data set stays in
// HBM (180 GB) and is streamed exactly twice by the same (sample, field)-per-CTA
// TMA pipeline as the rollout step (coarse_step_staged.cuh):
//
//   pass STATS  reads u_t, u_{t+1} (8 B/pt), writes 6 partial sums per plane
//   pass EMIT   reads u_t, u_{t+1}, writes the three normalised arrays (8 + 12 B/pt)
//
// pdeu and label are never materialised un-normalised (recomputing the coarse step
// is free under the HBM bound; storing them would cost 16 + 24 B/pt instead).
#include "coarse_step_staged.cuh"

namespace ppnn {

#ifdef PPNN_TRACE
// dev-only timeline of the CTAs (entry, end of the prologue, end of the stream, SM id); never in the shipped library
__device__ unsigned long long g_trace[6 * 8192];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

struct DatasetIO {
    const float* u;          // sample b = planes [slot(b)*2, slot(b)*2+2) of this array
    const float* x;          // the state one step later, same slots
    const float* mu;         // [B]
    float* u_norm;           // [B,2,FH-1,FW-1]   (EMIT)
    float* p_norm;           // [B,2,FH-1,FW-1]   (EMIT)
    float* l_norm;           // [B,2,FH-1,FW-1]   (EMIT)
    double* partial;         // [B,2,6] sum u, sum u^2, sum pdeu, sum pdeu^2, sum label, sum label^2   (STATS)
    float m[3][2], s[3][2];  // mean / std per channel of (u, pdeu, label)   (EMIT)
    long T;                  // samples per trajectory (slot(b) = b + b / T skips each trajectory's last state); 0 = dense
    const float* u_end;      // one past the last float the kernel may read from u / x
    const float* x_end;
};

// PER > 0: compile-time prolongation period for the reference's meshes (see coarse_step_staged.cuh)
#ifndef PPNN_NST_DS49
#define PPNN_NST_DS49 4
#endif
#ifndef PPNN_NST_DS33
#define PPNN_NST_DS33 4
#endif
// ring depths (see coarse_step.cu): 5 / 6 stages of 16.5 KB fit but do not pay (statistics 0.402 vs 0.407 ms)
constexpr int kNstDs49 = PPNN_NST_DS49, kNstDs33 = PPNN_NST_DS33;

template <int FWC, bool EMIT, int PER = 0, int ADV = 0, int NST = kStages>
__global__ void __launch_bounds__(kStagedThreads, PPNN_MINCTA)
dataset_pass_kernel(const CoarseParams p, const DatasetIO io, const StepPlan pl, const StagedLayout lay, const PhaseWeights pw) {
    extern __shared__ float4 smem4[];
    const int trows = PER > 0 ? 0 : p.FH, tints = PER > 0 ? 0 : (p.FH + 2 * lay.nsteps + 3) & ~3;
    float4* roww = smem4;
    int* rowi = reinterpret_cast<int*>(roww + trows);
    int* rowadv = rowi + trows;
    const int nn2 = p.NH * p.NW;
    float* cu = reinterpret_cast<float*>(rowi + tints);
    float* cv = cu + nn2;
    float* rf = cv + nn2;
    float* sU = cu + ((3 * nn2 + 3) & ~3);                      // [NST][2][ustage]
    float* sX = sU + (size_t)NST * 2 * lay.ustage;          // [NST][2][ustage]
    __shared__ __align__(8) uint64_t full[NST];
    __shared__ __align__(8) uint64_t empty[NST];
    __shared__ double red[PPNN_STEP_WARPS][6];

    const int FW = FWC ? FWC : p.FW;
    const int PW = FW - 1, PH = p.FH - 1;
    const long b = blockIdx.x >> 1;
    const int f = blockIdx.x & 1;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)p.FH * FW;
    const size_t nplane = (size_t)PH * PW;
    const size_t slot = (size_t)(io.T > 0 ? b + b / io.T : b);
    const float* ub = io.u + slot * 2 * plane;
    const float* uf = ub + f * plane;
    const float* xf = io.x + (slot * 2 + f) * plane;
    const float mu = __ldg(io.mu + b);
    const int ngroups = lay.split < p.FH ? 2 : 1;
    const int nconsumers = min(ngroups * pl.nchunk, PPNN_STEP_WARPS);
    const int rid = f ? p.rhs_v : p.rhs_u;

    // one band of one operand: aligned bulk copy, the tail of the tensor finished by scalar loads
    auto band = [&](const float* src, int nrows, const float* end, float* dst, uint64_t* bar) -> uint32_t {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        const uintptr_t a0 = a & ~(uintptr_t)15;
        const uintptr_t end_al = reinterpret_cast<uintptr_t>(end) & ~(uintptr_t)15;
        uint32_t nb = (uint32_t)((a - a0) + (size_t)nrows * FW * 4 + 15) & ~15u;
        if (a0 + nb > end_al) {
            for (const float* q = reinterpret_cast<const float*>(end_al); q < end; ++q)
                dst[(reinterpret_cast<uintptr_t>(q) - a0) >> 2] = *q;
            nb = (uint32_t)(end_al - a0);
        }
        if (nb) bulk_g2s(dst, reinterpret_cast<const void*>(a0), nb, bar);
        return nb;
    };
    auto issue = [&](int t) {
        const int s = t % NST;
        // expected bytes first (same arithmetic as band()), then the copies
        uint32_t total = 0;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int rs = g ? lay.split : 0, re = g ? p.FH : lay.split;
                const int r0 = rs + t * kStageRows;
                if (g >= ngroups || r0 >= re) continue;
                const int nrows = min(kStageRows, re - r0);
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const float* src = (o ? xf : uf) + (size_t)r0 * FW;
                    const float* end = o ? io.x_end : io.u_end;
                    if (pass == 0) {
                        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
                        const uintptr_t a0 = a & ~(uintptr_t)15;
                        const uintptr_t end_al = reinterpret_cast<uintptr_t>(end) & ~(uintptr_t)15;
                        uint32_t nb = (uint32_t)((a - a0) + (size_t)nrows * FW * 4 + 15) & ~15u;
                        if (a0 + nb > end_al) nb = (uint32_t)(end_al - a0);
                        total += nb;
                    } else {
                        band(src, nrows, end, (o ? sX : sU) + ((size_t)s * 2 + g) * lay.ustage, &full[s]);
                    }
                }
            }
            if (pass == 0) mbar_expect_tx(&full[s], total);
        }
    };

    const bool producer = warp == PPNN_STEP_WARPS;
#ifdef PPNN_TRACE
    if (tid == 0 && blockIdx.x < 8192) { unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); g_trace[6 * blockIdx.x] = gtime(); g_trace[6 * blockIdx.x + 3] = sm; }
#endif
    if (producer) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nconsumers); }
            fence_barrier_init();
            for (int t = 0; t < NST && t < lay.nsteps; ++t) issue(t);
        }
    } else {
        // ---- phase 1: restriction gather + row tables (as coarse_step_staged_kernel) ----
        const bool both = (rid & PPNN_RHS_ADV) || (rid & ~3);
        const int lo = both ? 0 : f * nn2, hi = both ? 2 * nn2 : (f + 1) * nn2;
        gather_coarse<true>(p, ub, plane, FW, lo, hi, cu, tid, kStepThreads);
        if (PER == 0) build_row_tables(p, lay.split, lay.nsteps, kStageRows, roww, rowi, rowadv, tid, kStepThreads);
    }
    __syncthreads();
#ifdef PPNN_TRACE
    if (tid == 0 && blockIdx.x < 8192) g_trace[6 * blockIdx.x + 4] = gtime();
#endif

    // ---- phase 2: dt * rhs of this CTA's field ----
    if (!producer) {
        if (p.adv_n == 2 && p.diff_n == 3) {
            for (int idx = tid; idx < nn2; idx += kStepThreads) {
                const int I = small_div(idx, p.NW), J = idx - I * p.NW;
                rf[idx] = eval_rhs_3pt(p, rid, f ? cv : cu, cu, cv, I, J, mu);
            }
        } else {
            for (int idx = tid; idx < nn2; idx += kStepThreads) {
                const int I = small_div(idx, p.NW), J = idx - I * p.NW;
                rf[idx] = p.dt * eval_rhs(p, rid, f ? cv : cu, cu, cv, I, J, mu);
            }
        }
    }
    __syncthreads();
#ifdef PPNN_TRACE
    if (tid == 0 && blockIdx.x < 8192) g_trace[6 * blockIdx.x + 1] = gtime();
#endif

    if (producer) {
        if (lane == 0) {
            for (int t = NST; t < lay.nsteps; ++t) {
                mbar_wait(&empty[t % NST], ((t / NST) - 1) & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t);
            }
        }
        return;
    }

    // per-thread partial sums (fp32 over the <= FH/2 rows of one column, then fp64)
    float a_u = 0.f, a_uu = 0.f, a_p = 0.f, a_pp = 0.f, a_l = 0.f, a_ll = 0.f;

    // ---- phase 3a: the duplicated-endpoint column only enters the pdeu statistics ----
    // (pde(i, FW-1) == pde(i, 0) bit for bit when the last column sits on the last coarse node: then
    //  the lane that owns column 0 counts it twice, see coarse_step_staged.cuh)
    const float slast = p.suw * (float)(FW - 1);
    const bool coldup = pl.wmain < FW && floorf(slast) == slast && ext_index((int)slast, p.CW) == 0;
    if (!EMIT && pl.wmain < FW && !coldup) {
        const int jc = FW - 1;
        float wy[4]; int j0;
        cubic_entry(p.suw, jc, &j0, wy);
        const int c0 = ext_index(j0 - 1, p.CW), c1 = ext_index(j0, p.CW),
                  c2 = ext_index(j0 + 1, p.CW), c3 = ext_index(j0 + 2, p.CW);
        for (int i = tid; i < p.FH; i += kStepThreads) {
            const int i0 = rowi[i];
            const float4 wx = roww[i];
            float h[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float* r = rf + ext_index(i0 - 1 + a, p.CH) * p.NW;
                h[a] = r[c0] * wy[0] + r[c1] * wy[1] + r[c2] * wy[2] + r[c3] * wy[3];
            }
            const float pde = h[0] * wx.x + h[1] * wx.y + h[2] * wx.z + h[3] * wx.w;
            a_p += pde; a_pp += pde * pde;
        }
    }

    // ---- phase 3b: warp = (row group, 32-column chunk), operands from the ring ----
    if constexpr (PER > 0) {
        constexpr int R = kStageRows, NACC = ADV + 3, SPLIT = 128;
        static_assert(SPLIT % PER == 0 && PER % R == 0 && FWC == 257 && PPNN_STEP_WARPS == 16, "period layout");
        const int g = warp >> 3;
        const int rs = g ? SPLIT : 0;
        const int j = (warp & 7) * 32 + lane;                   // 0 .. 255
        float wy[4]; int j0;
        cubic_entry(p.suw, j, &j0, wy);
        const float* rc0 = rf + ext_index(j0 - 1, p.CW);
        const float* rc1 = rf + ext_index(j0, p.CW);
        const float* rc2 = rf + ext_index(j0 + 1, p.CW);
        const float* rc3 = rf + ext_index(j0 + 2, p.CW);
        auto yrow = [&](int I) -> float {
            const int o = ext_index(I, p.CH) * p.NW;
            return rc0[o] * wy[0] + rc1[o] * wy[1] + rc2[o] * wy[2] + rc3[o] * wy[3];
        };
        const float m0 = f ? io.m[0][1] : io.m[0][0], s0 = f ? io.s[0][1] : io.s[0][0];
        const float m1 = f ? io.m[1][1] : io.m[1][0], s1 = f ? io.s[1][1] : io.s[1][0];
        const float m2 = f ? io.m[2][1] : io.m[2][0], s2 = f ? io.s[2][1] : io.s[2][0];
        const float rs0 = 1.f / s0, rs1 = 1.f / s1, rs2 = 1.f / s2;
        const size_t o0 = ((size_t)b * 2 + f) * nplane + (size_t)rs * PW + j;
        const int lead_u = (int)(reinterpret_cast<uintptr_t>(uf) & 15) >> 2;
        const int lead_x = (int)(reinterpret_cast<uintptr_t>(xf) & 15) >> 2;
        const float* su_base = sU + (size_t)g * lay.ustage + lead_u + j;
        const float* sx_base = sX + (size_t)g * lay.ustage + lead_x + j;
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const size_t ustep = (size_t)2 * lay.ustage;
        auto stream_rows = [&](auto dupl_tag) {
            constexpr bool DUPL = decltype(dupl_tag)::value;
            const bool dupl = DUPL && lane == 0;                // this lane counts the duplicated endpoint column too
            float* unp = EMIT ? io.u_norm + o0 : nullptr;
            float* pnp = EMIT ? io.p_norm + o0 : nullptr;
            float* lnp = EMIT ? io.l_norm + o0 : nullptr;
            const float* su = su_base;
            const float* sx = sx_base;
            int s = 0; uint32_t ph = 0;
            int base = g ? SPLIT * ADV / PER : 0;
            float2 A_u = make_float2(0.f, 0.f), A_uu = A_u, A_p = A_u, A_pp = A_u, A_l = A_u, A_ll = A_u;
            float h[NACC];
#pragma unroll
            for (int k = 0; k < NACC; ++k) h[k] = yrow(base - 1 + k);
            auto emit = [&](int q, float pde, bool last_row) {
                if (EMIT) {
                    if (!last_row) {
                        const float uu = su[q * FW];
                        const float lab = (sx[q * FW] - uu) - pde;    // (u_{t+1} - u_t) - pdeu, in the reference's order
                        st_stream(unp + (size_t)q * PW, div_by(uu - m0, s0, rs0));
                        st_stream(pnp + (size_t)q * PW, div_by(pde - m1, s1, rs1));
                        st_stream(lnp + (size_t)q * PW, div_by(lab - m2, s2, rs2));
                    }
                } else {
                    a_p += pde; a_pp += pde * pde;
                    if (DUPL && dupl) { a_p += pde; a_pp += pde * pde; }
                    if (!last_row) {
                        const float uu = su[q * FW];
                        const float lab = (sx[q * FW] - uu) - pde;
                        a_u += uu; a_uu += uu * uu; a_l += lab; a_ll += lab * lab;
                    }
                }
            };
            auto advance = [&]() {
                __syncwarp();
                if (lane == 0) mbar_arrive_addr(empty_a + 8u * s);
                if (++s == NST) { s = 0; ph ^= 1u; su = su_base; sx = sx_base; }
                else { su += ustep; sx += ustep; }
                if (EMIT) { unp += (size_t)R * PW; pnp += (size_t)R * PW; lnp += (size_t)R * PW; }
            };
            for (int per = 0; per < SPLIT / PER; ++per) {
#pragma unroll
                for (int bnd = 0; bnd < PER / R; ++bnd) {
                    mbar_wait_addr(full_a + 8u * s, ph);
                    auto pde_row = [&](int q) -> float {
                        const int r = bnd * R + q;
                        const int o = (r * ADV) / PER;
                        return h[o] * pw.w[r][0] + h[o + 1] * pw.w[r][1] + h[o + 2] * pw.w[r][2] + h[o + 3] * pw.w[r][3];
                    };
#ifdef PPNN_DBG_NOCOMPUTE
                    // dev-only timing build (profiles/r2_cta_timelines.txt): consumers only wait for and release the bands
                    constexpr bool kCompute = false;
                    A_p.x += su[0];
#else
                    constexpr bool kCompute = true;
#endif
                    if constexpr (!kCompute) {
                    } else if constexpr (EMIT) {
#pragma unroll
                        for (int q = 0; q < R; ++q) emit(q, pde_row(q), false);
                    } else {
                        // statistics: two rows per packed fp32x2 instruction (8 issue slots per row pair
                        // instead of 16 for the two differences and six accumulations)
#pragma unroll
                        for (int q = 0; q < R; q += 2) {
                            const float2 pde = make_float2(pde_row(q), pde_row(q + 1));
                            const float2 uu = make_float2(su[q * FW], su[(q + 1) * FW]);
                            const float2 xx = make_float2(sx[q * FW], sx[(q + 1) * FW]);
                            const float2 m1 = make_float2(-1.f, -1.f);
                            // (u_{t+1} - u_t) - pdeu in the reference's order: a * -1 + b rounds once, like b - a
                            const float2 lab = __ffma2_rn(pde, m1, __ffma2_rn(uu, m1, xx));
                            A_p = __fadd2_rn(A_p, pde);  A_pp = __ffma2_rn(pde, pde, A_pp);
                            if (DUPL && dupl) { A_p = __fadd2_rn(A_p, pde);  A_pp = __ffma2_rn(pde, pde, A_pp); }
                            A_u = __fadd2_rn(A_u, uu);   A_uu = __ffma2_rn(uu, uu, A_uu);
                            A_l = __fadd2_rn(A_l, lab);  A_ll = __ffma2_rn(lab, lab, A_ll);
                        }
                    }
                    advance();
                }
                base += ADV;
#pragma unroll
                for (int k = 0; k < NACC; ++k) h[k] = k + ADV < NACC ? h[k + ADV] : yrow(base - 1 + k);
            }
            if (g == 1) {                                       // row 256: on the last coarse node, outside the cropped arrays
                mbar_wait_addr(full_a + 8u * s, ph);
                emit(0, h[1], true);
            }
            if (!EMIT) {
                a_u += A_u.x + A_u.y; a_uu += A_uu.x + A_uu.y; a_p += A_p.x + A_p.y;
                a_pp += A_pp.x + A_pp.y; a_l += A_l.x + A_l.y; a_ll += A_ll.x + A_ll.y;
            }
        };
        if ((warp & 7) == 0) stream_rows(std::true_type{}); else stream_rows(std::false_type{});
    } else if (warp < nconsumers) {
        const int g = warp / pl.nchunk;
        const int chunk = warp - g * pl.nchunk;
        const int rs = g ? lay.split : 0, re = g ? p.FH : lay.split;
        const int j = chunk * 32 + lane;
        const bool active = j < pl.wmain;
        const int jc = active ? j : pl.wmain - 1;
        const bool colok = active && jc < PW;

        float wy[4]; int j0;
        cubic_entry(p.suw, jc, &j0, wy);
        const float* rc0 = rf + ext_index(j0 - 1, p.CW);
        const float* rc1 = rf + ext_index(j0, p.CW);
        const float* rc2 = rf + ext_index(j0 + 1, p.CW);
        const float* rc3 = rf + ext_index(j0 + 2, p.CW);
        auto yrow = [&](int I) -> float {
            const int o = ext_index(I, p.CH) * p.NW;
            return rc0[o] * wy[0] + rc1[o] * wy[1] + rc2[o] * wy[2] + rc3[o] * wy[3];
        };
        int base = rowi[rs];
        float h0 = yrow(base - 1), h1 = yrow(base), h2 = yrow(base + 1), h3 = yrow(base + 2);
        auto prolong = [&](const float4 wx, bool adv) -> float {
            if (adv) { ++base; h0 = h1; h1 = h2; h2 = h3; h3 = yrow(base + 2); }
            return h0 * wx.x + h1 * wx.y + h2 * wx.z + h3 * wx.w;
        };

        const float m0 = f ? io.m[0][1] : io.m[0][0], s0 = f ? io.s[0][1] : io.s[0][0];
        const float m1 = f ? io.m[1][1] : io.m[1][0], s1 = f ? io.s[1][1] : io.s[1][0];
        const float m2 = f ? io.m[2][1] : io.m[2][0], s2 = f ? io.s[2][1] : io.s[2][0];
        const float rs0 = 1.f / s0, rs1 = 1.f / s1, rs2 = 1.f / s2;
        const size_t o0 = ((size_t)b * 2 + f) * nplane + (size_t)rs * PW + jc;
        float* unp = EMIT ? io.u_norm + o0 : nullptr;
        float* pnp = EMIT ? io.p_norm + o0 : nullptr;
        float* lnp = EMIT ? io.l_norm + o0 : nullptr;
        const int lead_u = (int)(reinterpret_cast<uintptr_t>(uf) & 15) >> 2;
        const int lead_x = (int)(reinterpret_cast<uintptr_t>(xf) & 15) >> 2;
        const float* su_base = sU + (size_t)g * lay.ustage + lead_u + jc;
        const float* sx_base = sX + (size_t)g * lay.ustage + lead_x + jc;
        const float4* wrow = roww + rs;
        const int* advp = rowadv + g * lay.nsteps;

        const bool dupl = coldup && j == 0;
        auto row_out = [&](int q, int row, int advbits, const float* su, const float* sx) {
            const float pde = prolong(wrow[q], (advbits >> q) & 1);
            const float uu = su[q * FW];
            const float lab = (sx[q * FW] - uu) - pde;            // (u_{t+1} - u_t) - pdeu, in the reference's order
            const bool in = colok && row < PH;
            if (EMIT) {
                if (in) {
                    st_stream(unp + (size_t)q * PW, div_by(uu - m0, s0, rs0));
                    st_stream(pnp + (size_t)q * PW, div_by(pde - m1, s1, rs1));
                    st_stream(lnp + (size_t)q * PW, div_by(lab - m2, s2, rs2));
                }
            } else {
                if (active) { a_p += pde; a_pp += pde * pde; }
                if (dupl) { a_p += pde; a_pp += pde * pde; }
                if (in) { a_u += uu; a_uu += uu * uu; a_l += lab; a_ll += lab * lab; }
            }
        };

        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const size_t ustep = (size_t)2 * lay.ustage;
        const float* su = su_base;
        const float* sx = sx_base;
        int s = 0; uint32_t ph = 0;
        const int nfull = (re - rs) / kStageRows;
        auto advance = [&]() {
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty_a + 8u * s);
            if (++s == NST) { s = 0; ph ^= 1u; su = su_base; sx = sx_base; }
            else { su += ustep; sx += ustep; }
            wrow += kStageRows;
            if (EMIT) { unp += (size_t)kStageRows * PW; pnp += (size_t)kStageRows * PW; lnp += (size_t)kStageRows * PW; }
        };
        int t = 0;
        for (; t < nfull; ++t) {
            const int advbits = advp[t];
            mbar_wait_addr(full_a + 8u * s, ph);
#pragma unroll
            for (int q = 0; q < kStageRows; ++q) row_out(q, rs + t * kStageRows + q, advbits, su, sx);
            advance();
        }
        for (; t < lay.nsteps; ++t) {
            const int r0 = rs + t * kStageRows;
            const int advbits = advp[t];
            mbar_wait_addr(full_a + 8u * s, ph);
            for (int q = 0; r0 + q < re; ++q) row_out(q, r0 + q, advbits, su, sx);
            advance();
        }
    }

#ifdef PPNN_TRACE
    if (tid == 0 && blockIdx.x < 8192) g_trace[6 * blockIdx.x + 2] = gtime();
#endif
    if (!EMIT) {
        // deterministic reduction: lanes -> warp (shuffles, fp64), warps -> CTA in warp order
        double v[6] = {(double)a_u, (double)a_uu, (double)a_p, (double)a_pp, (double)a_l, (double)a_ll};
#pragma unroll
        for (int k = 0; k < 6; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
            if (lane == 0) red[warp][k] = v[k];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(kStepThreads) : "memory");     // the 16 consumer-side warps
        if (tid < 6) {
            double acc = 0.0;
            for (int w = 0; w < PPNN_STEP_WARPS; ++w) acc += red[w][tid];
            io.partial[((size_t)b * 2 + f) * 6 + tid] = acc;
        }
    }
}

static int run_dataset(const ppnn_coarse_cfg* cfg, DatasetIO io, long B, bool emit, void* stream) {
    CoarseParams p;
    int st = fill_coarse_params(cfg, &p);
    if (st != PPNN_OK) return st;
    if (B < 0 || io.T < 0) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    if (B >= (1L << 30)) return PPNN_E_TOOBIG;
    StepPlan pl;
    pl.wmain = ((p.FW - 1) % 32 == 0) ? p.FW - 1 : p.FW;
    pl.nchunk = (pl.wmain + 31) / 32;
    pl.nband = 1; pl.band_rows = p.FH;
    // the streamed path needs bands that start on a constant 16-byte phase and one warp per 32-column chunk
    if (pl.nchunk > PPNN_STEP_WARPS || (p.FW - 1) % 4 != 0 || p.CH > p.FH) return PPNN_E_SHAPE;
    StagedLayout lay;
    lay.ustage = (kStageRows * p.FW + 8 + 3) & ~3;
    lay.nstage = 0;
    lay.split = p.FH;
    if (2 * pl.nchunk <= PPNN_STEP_WARPS && p.FH >= 4 * kStageRows)
        lay.split = ((p.FH / 2 + kStageRows - 1) / kStageRows) * kStageRows;
    int n0 = (lay.split + kStageRows - 1) / kStageRows;
    int n1 = (p.FH - lay.split + kStageRows - 1) / kStageRows;
    lay.nsteps = n0 > n1 ? n0 : n1;
    const size_t plane2 = (size_t)2 * p.FH * p.FW;
    const size_t last = (size_t)(io.T > 0 ? (B - 1) + (B - 1) / io.T : B - 1);
    io.u_end = io.u + (last + 1) * plane2;
    io.x_end = io.x + (last + 1) * plane2;
    void (*kern)(CoarseParams, DatasetIO, StepPlan, StagedLayout, PhaseWeights);
    const bool square = p.FH == 257 && p.FW == 257 && p.CH == p.CW && lay.split == 128 && pl.nchunk == 8 &&
                        !PPNN_ENV_OFF("PPNN_NO_PERIODIC");
    // compile-time period: no run-time tables, ring as deep as two resident CTAs allow (stages of 16.5 KB)
    int per = 0, nst = kStages;
    if (square && p.CH == 49)      { per = 16; nst = kNstDs49; kern = emit ? dataset_pass_kernel<257, true, 16, 3, kNstDs49> : dataset_pass_kernel<257, false, 16, 3, kNstDs49>; }
    else if (square && p.CH == 33) { per = 8;  nst = kNstDs33; kern = emit ? dataset_pass_kernel<257, true, 8, 1, kNstDs33> : dataset_pass_kernel<257, false, 8, 1, kNstDs33>; }
    else if (p.FW == 257) kern = emit ? dataset_pass_kernel<257, true> : dataset_pass_kernel<257, false>;
    else                  kern = emit ? dataset_pass_kernel<0, true> : dataset_pass_kernel<0, false>;
    const size_t tables = per ? 0 : (size_t)p.FH * sizeof(float4) + (size_t)((p.FH + 2 * lay.nsteps + 3) & ~3) * sizeof(int);
    const size_t smem = tables + (size_t)((3 * p.NH * p.NW + 3) & ~3) * sizeof(float) +
                        (size_t)nst * 2 * 2 * lay.ustage * sizeof(float);
    if (smem > 200 * 1024) return PPNN_E_TOOBIG;
#ifdef PPNN_DBG_SMEM      // dev-only timing build: padded shared memory => one CTA per SM (profiles/r2_cta_timelines.txt)
    const size_t smem_launch = PPNN_DBG_SMEM;
#else
    const size_t smem_launch = smem;
#endif
    const PhaseWeights pw = make_phase_weights(p.suh, per);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch);
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)(2 * B), kStagedThreads, smem_launch, (cudaStream_t)stream>>>(p, io, pl, lay, pw);
    return launch_status();
}

}  // namespace ppnn

using namespace ppnn;

#ifdef PPNN_TRACE
extern "C" int ppnn_trace_read(unsigned long long* host, long n) {
    return (int)cudaMemcpyFromSymbol(host, g_trace, (size_t)n * sizeof(unsigned long long));
}
#endif

extern "C" int ppnn_dataset_stats(const ppnn_coarse_cfg* cfg, const float* u, const float* u_next, const float* mu,
                                  long B, long traj_len, double* partial, void* stream) {
    if (!cfg || !u || !u_next || !mu || !partial) return PPNN_E_NULL;
    DatasetIO io = {};
    io.u = u; io.x = u_next; io.mu = mu; io.partial = partial; io.T = traj_len;
    return run_dataset(cfg, io, B, false, stream);
}

extern "C" int ppnn_dataset_emit(const ppnn_coarse_cfg* cfg, const float* u, const float* u_next, const float* mu,
                                 long B, long traj_len, const float* mean6, const float* std6,
                                 float* u_norm, float* pdeu_norm, float* label_norm, void* stream) {
    if (!cfg || !u || !u_next || !mu || !mean6 || !std6 || !u_norm || !pdeu_norm || !label_norm) return PPNN_E_NULL;
    DatasetIO io = {};
    io.u = u; io.x = u_next; io.mu = mu; io.T = traj_len;
    io.u_norm = u_norm; io.p_norm = pdeu_norm; io.l_norm = label_norm;
    for (int k = 0; k < 3; ++k)
        for (int f = 0; f < 2; ++f) { io.m[k][f] = mean6[2 * k + f]; io.s[k][f] = std6[2 * k + f]; }
    return run_dataset(cfg, io, B, true, stream);
}

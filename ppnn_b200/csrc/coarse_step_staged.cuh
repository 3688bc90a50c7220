// TMA-staged variant of the fused coarse step (MODE_UPDATE / MODE_INPUTS) for
// meshes of at most 256 + 1 columns -- the reference mesh, 257.
//
// Same phases as coarse_step_kernel, but the streamed operands (u and nn_out) do
// not go through registers: one elected thread feeds a 2-stage shared-memory
// ring with `cp.async.bulk` (TMA, 1-D) copies of kStageRows-row bands of both
// field planes, completion is signalled on an mbarrier, and the 16 warps (field,
// 32-column chunk) read their operands with LDS.  The first two stages are
// issued at kernel entry, so they are in flight while the latency-bound
// restriction gather and the coarse RHS run.  Long-scoreboard (global-load)
// stalls in the hot loop become short shared-memory ones, and the prefetch
// depth is no longer bounded by registers or by the six scoreboard slots.
//
// Alignment: a band of rows of a [.., 257, 257] plane starts at a 4-byte, not a
// 16-byte, boundary.  Each copy therefore starts at the enclosing 16-byte
// boundary and the consumer skips `lead` = 0..3 floats; the last copy of the
// tensor is clamped to its 16-byte-aligned end and the few trailing floats are
// fetched with ordinary loads, so nothing outside the tensor is ever read.
#pragma once
#include <type_traits>
#include "coarse_step_types.cuh"
#include "tma.cuh"

namespace ppnn {

#ifndef PPNN_STAGE_ROWS
#define PPNN_STAGE_ROWS 4
#endif
#ifndef PPNN_STAGES
#define PPNN_STAGES 4
#endif
constexpr int kStageRows = PPNN_STAGE_ROWS;   // kStageRows * FW must be a multiple of 4 (constant lead)
constexpr int kStages = PPNN_STAGES;

// L2 eviction-priority policies: the restriction gather touches sectors that the
// TMA stream re-reads tens of microseconds later -> keep them (evict_last); the
// streamed bands themselves are dead after one use (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#ifndef PPNN_HINT_GATHER
#define PPNN_HINT_GATHER 1
#endif
#ifndef PPNN_HINT_STREAM
#define PPNN_HINT_STREAM 0
#endif
__device__ __forceinline__ float ld_hint(const float* p, uint64_t pol) {
#if PPNN_HINT_GATHER
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
#else
    return __ldg(p);
#endif
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
#if !PPNN_HINT_STREAM
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
    return;
#endif
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

constexpr int kGatherBatch = 5;                     // coarse points per thread whose loads are issued together

// Bilinear restriction gather (mesh_convertor.down + [:, :, :-1, :-1]) of the coarse points
// [lo, hi) of the index space [field][I][J] into dst[idx], by `nthr` threads: the up-to-4 taps
// of kGatherBatch points per thread are all in flight before the first is used; zero-weight
// taps are skipped (exact decimation for 257 -> 33).  KEEP: tag the loads L2::evict_last
// because the same sectors are streamed again a few microseconds later.
template <bool KEEP, bool FASTDIV = true>
__device__ __forceinline__ void gather_coarse(const CoarseParams& p, const float* __restrict__ ub, size_t plane, int FW,
                                              int lo, int hi, float* dst, int tid, int nthr) {
    const int nn2 = p.NH * p.NW;
    const uint64_t pol_keep = KEEP ? l2_policy_evict_last() : 0;
    auto ld = [&](const float* q) -> float { return KEEP ? ld_hint(q, pol_keep) : __ldg(q); };
    for (int i0 = lo + tid; i0 < hi; i0 += kGatherBatch * nthr) {
        float a[kGatherBatch], bq[kGatherBatch], c[kGatherBatch], d[kGatherBatch];
        BilinearTap tp[kGatherBatch];
#pragma unroll
        for (int k = 0; k < kGatherBatch; ++k) {
            const int idx = i0 + k * nthr;
            a[k] = bq[k] = c[k] = d[k] = 0.f;
            if (idx < hi) {
                const int g = idx >= nn2;
                const int r = idx - g * nn2;
                const int I = FASTDIV ? small_div(r, p.NW) : r / p.NW, J = r - I * p.NW;
                tp[k] = bilinear_tap(p, I, J);
                const float* q = ub + g * plane + (size_t)tp[k].i0 * FW + tp[k].j0;
                a[k] = ld(q);
                if (tp[k].lj != 0.f) bq[k] = ld(q + tp[k].dj);
                if (tp[k].li != 0.f) {
                    const float* r2 = q + (size_t)tp[k].di * FW;
                    c[k] = ld(r2);
                    if (tp[k].lj != 0.f) d[k] = ld(r2 + tp[k].dj);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kGatherBatch; ++k) {
            const int idx = i0 + k * nthr;
            if (idx < hi) {
                const float h0 = 1.f - tp[k].li, w0 = 1.f - tp[k].lj;
                const float top = w0 * a[k] + tp[k].lj * bq[k];
                dst[idx] = tp[k].li == 0.f ? h0 * top : h0 * top + tp[k].li * (w0 * c[k] + tp[k].lj * d[k]);
            }
        }
    }
}

// Per-CTA tables of the prolongation: H-pass weights and source row of every fine row, and per
// ring step a bit mask "row q of this band moves the 4-row coarse window on" for both row groups.
__device__ __forceinline__ void build_row_tables(const CoarseParams& p, int split, int nsteps, int rows_per_step,
                                                 float4* roww, int* rowi, int* rowadv, int tid, int nthr) {
    for (int i = tid; i < p.FH; i += nthr) {
        float w[4]; int i0;
        cubic_entry(p.suh, i, &i0, w);
        rowi[i] = i0;
        roww[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
    for (int e = tid; e < 2 * nsteps; e += nthr) {
        const int g = e >= nsteps, t = e - g * nsteps;
        const int rs = g ? split : 0, re = g ? p.FH : split;
        int bits = 0;
        for (int q = 0; q < rows_per_step; ++q) {
            const int i = rs + t * rows_per_step + q;
            if (i > rs && i < re && floorf(p.suh * (float)i) != floorf(p.suh * (float)(i - 1))) bits |= 1 << q;
        }
        rowadv[e] = bits;
    }
}

struct StagedLayout {
    int ustage;      // floats per (stage, group) u buffer: kStageRows*FW + 8 (lead / tail slack), 16 B multiple
    int nstage;      // floats per (stage, group) nn buffer: kStageRows*PW
    int split;       // first row of the second warp group (multiple of kStageRows); FH = no second group
    int nsteps;      // ring steps = max over groups of ceil(rows / kStageRows)
};


constexpr int kStagedThreads = kStepThreads + 32;   // 16 consumer warps + 1 producer warp

// One CTA = one (sample, field) plane.  Consumer warp = (row group g, 32-column
// chunk): group 0 walks rows [0, split), group 1 rows [split, FH), both fed by
// the same ring step.  HAS_NN: the update adds the network output (MODE_UPDATE).
//
// PER > 0: the reference's own meshes (257 fine points, 33 / 49 coarse points), where the prolongation is
// periodic in the fine row index -- every PER rows the 4-row window of W-interpolated coarse rows moves on
// by exactly ADV rows and the PER weight phases repeat.  The consumer loop is then unrolled over one period:
// which window entries a row uses, and when new ones are computed, is decided at compile time, and the
// weights (`pw`, a kernel parameter) are constant-bank operands.  PER == 0: any mesh, run-time tables.
// NST: ring depth (chosen per instantiation in the launchers from same-box timings).
template <int MODE, int FWC, bool HAS_NN, int PER = 0, int ADV = 0, int NST = kStages>
__global__ void __launch_bounds__(kStagedThreads, PPNN_MINCTA)
coarse_step_staged_kernel(const CoarseParams p, const StepIO io, const StepPlan pl, const StagedLayout lay,
                          const float* u_end, const PhaseWeights pw) {
    // Index splits in the per-coarse-point phases go through small_div (fp32 reciprocal) except in the memory-bound
    // update with the network operand: there the shorter sequence changes ptxas' allocation under the 56-register
    // cap (12 bytes of spills in the band loop, 0.485 -> 0.494 ms) while the issue-bound modes gain 2-10 %.
    constexpr bool FASTDIV = !(MODE == MODE_UPDATE && HAS_NN);
    extern __shared__ float4 smem4[];
    // run-time prolongation tables (PER == 0 only; the compile-time period needs none)
    const int trows = PER > 0 ? 0 : p.FH, tints = PER > 0 ? 0 : (p.FH + 2 * lay.nsteps + 3) & ~3;
    float4* roww = smem4;                                       // [FH] H-pass weights
    int* rowi = reinterpret_cast<int*>(roww + trows);           // [FH] floor(i * suh)
    int* rowadv = rowi + trows;                                 // [2][nsteps] bit q: row rs_g + R*t + q moves the window
    const int nn2 = p.NH * p.NW;
    float* cu = reinterpret_cast<float*>(rowi + tints);
    float* cv = cu + nn2;
    float* rf = cv + nn2;                                       // dt * rhs of this CTA's field
    float* sU = cu + ((3 * nn2 + 3) & ~3);                      // [NST][2][ustage], 16-byte aligned
    float* sN = sU + (size_t)NST * 2 * lay.ustage;          // [NST][2][nstage]
    __shared__ __align__(8) uint64_t full[NST];
    __shared__ __align__(8) uint64_t empty[NST];

    const int FW = FWC ? FWC : p.FW;
    const int PW = FW - 1, PH = p.FH - 1;
    const int b = blockIdx.x >> 1, f = blockIdx.x & 1;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)p.FH * FW;
    const size_t nplane = (size_t)PH * PW;
    const float* ub = io.u + (size_t)b * 2 * plane;             // both fields of the sample (gather)
    const float* uf = ub + f * plane;                           // this CTA's plane
    const float* nf = HAS_NN ? io.nn + ((size_t)b * 2 + f) * nplane : nullptr;
    const float mu = __ldg(io.mu + b);
    const int ngroups = lay.split < p.FH ? 2 : 1;
    const int nconsumers = min(ngroups * pl.nchunk, PPNN_STEP_WARPS);
    const int rid = f ? p.rhs_v : p.rhs_u;

    const uint64_t pol_stream = l2_policy_evict_first();
    // producer: bulk copies of ring step t (kStageRows rows of each row group) into slot t % NST
    auto issue = [&](int t) {
        const int s = t % NST;
        uintptr_t a0[2]; uint32_t nb[2] = {0, 0}, nnb[2] = {0, 0}; int nr0[2] = {0, 0};
        uint32_t total = 0;
        const uintptr_t end_al = reinterpret_cast<uintptr_t>(u_end) & ~(uintptr_t)15;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int rs = g ? lay.split : 0, re = g ? p.FH : lay.split;
            const int r0 = rs + t * kStageRows;
            if (g >= ngroups || r0 >= re) continue;
            const int nrows = min(kStageRows, re - r0);
            const uintptr_t a = reinterpret_cast<uintptr_t>(uf + (size_t)r0 * FW);
            a0[g] = a & ~(uintptr_t)15;
            nb[g] = (uint32_t)((a - a0[g]) + (size_t)nrows * FW * 4 + 15) & ~15u;
            if (a0[g] + nb[g] > end_al) {                       // last copy of the tensor: stay inside it
                float* dst = sU + ((size_t)s * 2 + g) * lay.ustage;
                for (const float* q = reinterpret_cast<const float*>(end_al); q < u_end; ++q)
                    dst[(reinterpret_cast<uintptr_t>(q) - a0[g]) >> 2] = *q;
                nb[g] = (uint32_t)(end_al - a0[g]);
            }
            nr0[g] = r0 < PH ? r0 : 0;                          // the endpoint row uses network row 0 (padBC_rd)
            nnb[g] = HAS_NN ? (uint32_t)(r0 < PH ? min(kStageRows, PH - r0) : 1) * PW * 4 : 0;
            total += nb[g] + nnb[g];
        }
        mbar_expect_tx(&full[s], total);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            if (nb[g]) bulk_g2s_hint(sU + ((size_t)s * 2 + g) * lay.ustage, reinterpret_cast<const void*>(a0[g]), nb[g], &full[s], pol_stream);
            if (nnb[g]) bulk_g2s_hint(sN + ((size_t)s * 2 + g) * lay.nstage, nf + (size_t)nr0[g] * PW, nnb[g], &full[s], pol_stream);
        }
    };

    const bool producer = warp == PPNN_STEP_WARPS;
    if (producer) {
        if (lane == 0 && MODE != MODE_PDEU) {
#pragma unroll
            for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nconsumers); }
            fence_barrier_init();
            for (int t = 0; t < NST && t < lay.nsteps; ++t) issue(t);   // in flight during phases 1 and 2
        }
    } else {
        // ---- phase 1: restriction gather (loads of kGatherBatch points in flight) + row tables ----
        // a purely diffusive right-hand side never reads the other field
        const bool both = (rid & PPNN_RHS_ADV) || (rid & ~3);
        const int lo = both ? 0 : f * nn2, hi = both ? 2 * nn2 : (f + 1) * nn2;
        gather_coarse<true, FASTDIV>(p, ub, plane, FW, lo, hi, cu, tid, kStepThreads);
        if (PER == 0) build_row_tables(p, lay.split, lay.nsteps, kStageRows, roww, rowi, rowadv, tid, kStepThreads);
    }
    __syncthreads();

    // ---- phase 2: dt * rhs of this CTA's field on the periodic coarse grid ------
    if (!producer) {
        if (io.coarse) {                                        // own field only: the sibling CTA writes the other one
            const float* own = f ? cv : cu;
            float* dst = io.coarse + ((size_t)b * 2 + f) * nn2;
            for (int idx = tid; idx < nn2; idx += kStepThreads) dst[idx] = own[idx];
        }
        if (p.adv_n == 2 && p.diff_n == 3) {
            for (int idx = tid; idx < nn2; idx += kStepThreads) {
                const int I = FASTDIV ? small_div(idx, p.NW) : idx / p.NW, J = idx - I * p.NW;
                rf[idx] = eval_rhs_3pt(p, rid, f ? cv : cu, cu, cv, I, J, mu);
            }
        } else {
            for (int idx = tid; idx < nn2; idx += kStepThreads) {
                const int I = FASTDIV ? small_div(idx, p.NW) : idx / p.NW, J = idx - I * p.NW;
                rf[idx] = p.dt * eval_rhs(p, rid, f ? cv : cu, cu, cv, I, J, mu);
            }
        }
    }
    __syncthreads();

    if (producer) {
        // ---- producer warp: refill a slot as soon as every consumer warp released it ----
        if (lane == 0 && MODE != MODE_PDEU) {
            for (int t = NST; t < lay.nsteps; ++t) {
                mbar_wait(&empty[t % NST], ((t / NST) - 1) & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t);
            }
        }
        return;
    }

    const float m0 = f ? io.m0[1] : io.m0[0], s0 = f ? io.s0[1] : io.s0[0];
    const float m1 = f ? io.m1[1] : io.m1[0], s1 = f ? io.s1[1] : io.s1[0];

    // ---- phase 3a: the duplicated-endpoint column, one point per thread --------
    const bool want_out = MODE == MODE_INPUTS ? io.out != nullptr : true;
    // When the last fine column sits exactly on the last coarse node (always for 257 -> 33 / 49), its
    // W-pass weights are (0,1,0,0) on the wrapped copy of coarse column 0, so pde(i, FW-1) is
    // bit-identical to pde(i, 0): the lane that owns column 0 stores it from the ring (see row_out).
    const float slast = p.suw * (float)(FW - 1);
    const bool coldup = pl.wmain < FW && floorf(slast) == slast && ext_index((int)slast, p.CW) == 0;
    if (pl.wmain < FW && want_out && !coldup) {
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
            const size_t o = ((size_t)b * 2 + f) * plane + (size_t)i * FW + jc;
            if (MODE == MODE_UPDATE) {
                const float uu = ld_stream(io.u + o);
                float v = uu + pde;
                if (HAS_NN) v = (ld_stream(nf + (size_t)(i == PH ? 0 : i) * PW) * s0 + m0) + uu + pde;
                st_stream(io.out + o, v);
            } else {
                st_stream(io.out + o, pde);
            }
        }
    }

    // ---- phase 3b: warp = (row group, 32-column chunk), operands from the ring ----
    if constexpr (PER > 0) {
        // compile-time period: FH = FW = 257, two row groups [0,128) | [128,257), eight 32-column chunks each
        constexpr int R = kStageRows, NACC = ADV + 3, SPLIT = 128;
        static_assert(SPLIT % PER == 0 && PER % R == 0 && FWC == 257 && PPNN_STEP_WARPS == 16, "period layout");
        const int g = warp >> 3;
        const int rs = g ? SPLIT : 0;
        const int j = (warp & 7) * 32 + lane;                   // 0 .. 255: every lane owns a column
        float wy[4]; int j0;
        cubic_entry(p.suw, j, &j0, wy);
        const float* rc0 = rf + ext_index(j0 - 1, p.CW);
        const float* rc1 = rf + ext_index(j0, p.CW);
        const float* rc2 = rf + ext_index(j0 + 1, p.CW);
        const float* rc3 = rf + ext_index(j0 + 2, p.CW);
        auto yrow = [&](int I) -> float {                       // W pass of row I of the extended increment
            const int o = ext_index(I, p.CH) * p.NW;
            return rc0[o] * wy[0] + rc1[o] * wy[1] + rc2[o] * wy[2] + rc3[o] * wy[3];
        };
        const float rs0 = 1.f / s0, rs1 = 1.f / s1;
        const int lead = (int)(reinterpret_cast<uintptr_t>(uf) & 15) >> 2;
        const float* su_base = sU + (size_t)g * lay.ustage + lead + j;
        const float* sn_base = sN + (size_t)g * lay.nstage + j;
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const size_t ustep = (size_t)2 * lay.ustage, nstep = (size_t)2 * lay.nstage;
        // chunk 0 also owns the duplicated endpoint column (its increment is bit-identical to column 0's, see above):
        // a separate instantiation of the loop, so the other seven chunks carry no instructions for it
        auto stream_rows = [&](auto dupl_tag) {
            constexpr bool DUPL = decltype(dupl_tag)::value;
            const bool dupl = DUPL && lane == 0;
            float* op = io.out ? io.out + ((size_t)b * 2 + f) * plane + (size_t)rs * FW + j : nullptr;
            float* unp = MODE == MODE_INPUTS ? io.u_norm + ((size_t)b * 2 + f) * nplane + (size_t)rs * PW + j : nullptr;
            float* pnp = (MODE == MODE_INPUTS && io.p_norm) ? io.p_norm + ((size_t)b * 2 + f) * nplane + (size_t)rs * PW + j : nullptr;
            const float* su = su_base;
            const float* sn = sn_base;
            int s = 0; uint32_t ph = 0;
            int base = g ? SPLIT * ADV / PER : 0;               // h[k] = W-interpolated coarse row base - 1 + k
            float h[NACC];
#pragma unroll
            for (int k = 0; k < NACC; ++k) h[k] = yrow(base - 1 + k);
            auto emit = [&](int q, float pde, bool last_row) {   // epilogue of one fine row (q: row inside the band)
                if (MODE == MODE_PDEU) {
                    st_stream(op + (size_t)q * FW, pde);
                    if (DUPL && dupl) st_stream(op + (size_t)q * FW + PW, pde);
                    return;
                }
                const float uu = su[q * FW];
                if (MODE == MODE_UPDATE) {
                    const float nv = HAS_NN ? sn[q * PW] * s0 + m0 : 0.f;
                    st_stream(op + (size_t)q * FW, HAS_NN ? nv + uu + pde : uu + pde);
                    if (DUPL && dupl) {
                        const float ue = su[q * FW + PW];
                        st_stream(op + (size_t)q * FW + PW, HAS_NN ? nv + ue + pde : ue + pde);
                    }
                } else {
                    if (op) st_stream(op + (size_t)q * FW, pde);
                    if (op && DUPL && dupl) st_stream(op + (size_t)q * FW + PW, pde);
                    if (!last_row) {
                        st_stream(unp + (size_t)q * PW, div_by(uu - m0, s0, rs0));
                        if (pnp) st_stream(pnp + (size_t)q * PW, div_by(pde - m1, s1, rs1));
                    }
                }
            };
            auto advance = [&]() {
                if (MODE != MODE_PDEU) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_addr(empty_a + 8u * s);
                    if (++s == NST) { s = 0; ph ^= 1u; su = su_base; sn = sn_base; }
                    else { su += ustep; sn += nstep; }
                }
                if (op) op += (size_t)R * FW;
                if (MODE == MODE_INPUTS) { unp += (size_t)R * PW; if (pnp) pnp += (size_t)R * PW; }
            };
            for (int per = 0; per < SPLIT / PER; ++per) {
#pragma unroll
                for (int bnd = 0; bnd < PER / R; ++bnd) {
                    if (MODE != MODE_PDEU) mbar_wait_addr(full_a + 8u * s, ph);
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int r = bnd * R + q;              // compile-time after unrolling
                        const int o = (r * ADV) / PER;
                        emit(q, h[o] * pw.w[r][0] + h[o + 1] * pw.w[r][1] + h[o + 2] * pw.w[r][2] + h[o + 3] * pw.w[r][3], false);
                    }
                    advance();
                }
                base += ADV;
#pragma unroll
                for (int k = 0; k < NACC; ++k) h[k] = k + ADV < NACC ? h[k + ADV] : yrow(base - 1 + k);
            }
            if (g == 1) {                                       // row 256 sits on the last coarse node: weights (0,1,0,0)
                if (MODE != MODE_PDEU) mbar_wait_addr(full_a + 8u * s, ph);
                emit(0, h[1], true);
            }
        };
        if ((warp & 7) == 0) stream_rows(std::true_type{}); else stream_rows(std::false_type{});
        return;
    }
    if (warp >= nconsumers) return;
    const int g = warp / pl.nchunk;
    const int chunk = warp - g * pl.nchunk;
    const int rs = g ? lay.split : 0, re = g ? p.FH : lay.split;
    const int j = chunk * 32 + lane;
    const bool active = j < pl.wmain;
    const int jc = active ? j : pl.wmain - 1;
    const int jn = jc == PW ? 0 : jc;                           // padBC_rd column fold

    float wy[4]; int j0;
    cubic_entry(p.suw, jc, &j0, wy);
    const float* rc0 = rf + ext_index(j0 - 1, p.CW);
    const float* rc1 = rf + ext_index(j0, p.CW);
    const float* rc2 = rf + ext_index(j0 + 1, p.CW);
    const float* rc3 = rf + ext_index(j0 + 2, p.CW);
    auto yrow = [&](int I) -> float {                           // W pass of row I of the extended increment
        const int o = ext_index(I, p.CH) * p.NW;
        return rc0[o] * wy[0] + rc1[o] * wy[1] + rc2[o] * wy[2] + rc3[o] * wy[3];
    };
    int base = rowi[rs];
    float h0 = yrow(base - 1), h1 = yrow(base), h2 = yrow(base + 1), h3 = yrow(base + 2);
    // H pass of one fine row; prolongation never skips a coarse row (F >= C), so the
    // window moves on by at most one row (flag from the per-band bit mask; warp-uniform)
    auto prolong = [&](const float4 wx, bool adv) -> float {
        if (adv) { ++base; h0 = h1; h1 = h2; h2 = h3; h3 = yrow(base + 2); }
        return h0 * wx.x + h1 * wx.y + h2 * wx.z + h3 * wx.w;
    };

    const float rs0 = 1.f / s0, rs1 = 1.f / s1;
    float* op = io.out ? io.out + ((size_t)b * 2 + f) * plane + (size_t)rs * FW + jc : nullptr;
    float* unp = MODE == MODE_INPUTS ? io.u_norm + ((size_t)b * 2 + f) * nplane + (size_t)rs * PW + jc : nullptr;
    float* pnp = (MODE == MODE_INPUTS && io.p_norm) ? io.p_norm + ((size_t)b * 2 + f) * nplane + (size_t)rs * PW + jc : nullptr;
    const bool colok = active && jc < PW;
    // kStageRows * FW and split * FW are multiples of 4 floats, so every band of this plane has the same lead
    const int lead = (int)(reinterpret_cast<uintptr_t>(uf) & 15) >> 2;
    const float* su_base = sU + (size_t)g * lay.ustage + lead + jc;
    const float* sn_base = sN + (size_t)g * lay.nstage + jn;
    const float4* wrow = roww + rs;
    const int* advp = rowadv + g * lay.nsteps;

    const bool dupl = coldup && j == 0;                         // this lane also owns the duplicated endpoint column
    auto row_out = [&](int q, int row, int advbits, const float* su, const float* sn) {
        const float pde = prolong(wrow[q], (advbits >> q) & 1);
        if (MODE == MODE_PDEU) {
            if (active) st_stream(op + (size_t)q * FW, pde);
            if (dupl) st_stream(op + (size_t)q * FW + PW, pde);
            return;
        }
        const float uu = su[q * FW];
        if (MODE == MODE_UPDATE) {
            const float v = HAS_NN ? (sn[q * PW] * s0 + m0) + uu + pde : uu + pde;
            if (active) st_stream(op + (size_t)q * FW, v);
            if (dupl) {
                const float ue = su[q * FW + PW];
                st_stream(op + (size_t)q * FW + PW, HAS_NN ? (sn[q * PW] * s0 + m0) + ue + pde : ue + pde);
            }
        } else {
            if (op && active) st_stream(op + (size_t)q * FW, pde);
            if (op && dupl) st_stream(op + (size_t)q * FW + PW, pde);
            if (colok && row < PH) {
                st_stream(unp + (size_t)q * PW, div_by(uu - m0, s0, rs0));
                if (pnp) st_stream(pnp + (size_t)q * PW, div_by(pde - m1, s1, rs1));
            }
        }
    };

    // ring position kept incrementally; barrier addresses converted to the shared window once
    const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
    const size_t ustep = (size_t)2 * lay.ustage, nstep = (size_t)2 * lay.nstage;
    const float* su = su_base;
    const float* sn = sn_base;
    int s = 0; uint32_t ph = 0;
    const int nfull = (re - rs) / kStageRows;                   // whole bands of this row group
    auto advance = [&]() {
        if (MODE != MODE_PDEU) {
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty_a + 8u * s);  // this warp is done with slot s
            if (++s == NST) { s = 0; ph ^= 1u; su = su_base; sn = sn_base; }
            else { su += ustep; sn += nstep; }
        }
        wrow += kStageRows;
        if (op) op += (size_t)kStageRows * FW;
        if (MODE == MODE_INPUTS) { unp += (size_t)kStageRows * PW; if (pnp) pnp += (size_t)kStageRows * PW; }
    };
    int t = 0;
    for (; t < nfull; ++t) {
        const int advbits = advp[t];
        if (MODE != MODE_PDEU) mbar_wait_addr(full_a + 8u * s, ph);
#pragma unroll
        for (int q = 0; q < kStageRows; ++q) row_out(q, rs + t * kStageRows + q, advbits, su, sn);
        advance();
    }
    for (; t < lay.nsteps; ++t) {                               // ragged last band; steps only the other group has
        const int r0 = rs + t * kStageRows;
        const int advbits = advp[t];
        if (MODE != MODE_PDEU) mbar_wait_addr(full_a + 8u * s, ph);
        for (int q = 0; r0 + q < re; ++q) row_out(q, r0 + q, advbits, su, sn);
        advance();
    }
}

}  // namespace ppnn

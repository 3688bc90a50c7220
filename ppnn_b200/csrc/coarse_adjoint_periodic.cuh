// Adjoint of pde_du on the reference's own meshes (257 fine points, 33 or 49 coarse points), where
// the prolongation is PERIODIC in the fine index: every PER fine rows the bicubic window moves on by
// exactly ADV coarse rows and the PER phases repeat (PER/ADV = 16/3 for 257 -> 49, 8/1 for 257 -> 33).
//
// Same decomposition as coarse_adjoint_staged.cuh (one CTA = one (sample, field) plane of grad_out,
// 2-CTA cluster per sample, producer warp + TMA ring), but everything the generic kernel decides at
// run time per row -- "does the 4-row coarse window move on?", which weights, which accumulator --
// is resolved at COMPILE time by unrolling one period:
//
//   phase 1  H-pass of bicubic^T while streaming: the consumer of column j keeps ADV+3 accumulators
//            (coarse rows base-1 .. base+ADV+1); the PER x 4 weights are kernel parameters, i.e.
//            constant-bank operands of the FFMAs (no weight loads, no window shifts, no bit tests);
//            after a period the ADV oldest accumulators retire into A[I][j] (shared memory).
//            Rows that phase 4 will scatter into are stored with an L2 evict_last hint, so the
//            atomics find them on chip; all other rows are streaming stores.
//   phase 2  W-pass of bicubic^T in the same compile-time form: work item = (coarse row I, period k)
//            produces the ADV columns J = ADV*k .. ADV*k+ADV-1 from the 2*PER fine columns whose
//            windows reach them; A rows are padded with zeros so no item needs a bounds test.  The
//            clamped borders (windows reaching below column 0 / beyond column C-1) are two short
//            compile-time sums per coarse row.  The duplicated endpoint row/column are folded
//            (padBC_rd^T) before / after this pass because the pass is linear.
//   phase 0' coarse state: one bulk copy of the state the forward pass saved (optional input) that
//            flies during phase 2, or the restriction gather of u (nothing saved).
//   phases 3-4 as in the generic kernel, with unrolled 3-point stencils and reciprocal multiplies.
#pragma once
#include <type_traits>
#include "coarse_adjoint_staged.cuh"

namespace ppnn {

__device__ __forceinline__ void st_keep(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// bit r set: the bilinear restriction of some coarse row touches fine row r of a period (the rows that
// receive scatter-adds in phase 4)
template <int PER, int ADV>
__host__ __device__ constexpr unsigned scatter_row_mask() {
    unsigned m = 0;
    for (int e = 0; e < ADV; ++e) {
        const int lo = e * PER / ADV;
        m |= 1u << lo;
        if ((e * PER) % ADV != 0) m |= 1u << (lo + 1);
    }
    return m;
}

template <int ADVN>
__device__ __forceinline__ float upwind_x_t(const CoarseParams& p, const float* f, int I, int J) {
    const float c = f[I * p.NW + J];
    float s = 0.f;
    if (c <= 0.f) {
#pragma unroll
        for (int k = 0; k < ADVN; ++k) s += p.adv_f[k] * f[wrap1(I + k, p.NH) * p.NW + J];
    } else if (c > 0.f) {
#pragma unroll
        for (int k = 0; k < ADVN; ++k) s += p.adv_b[k] * f[wrap1(I - (ADVN - 1) + k, p.NH) * p.NW + J];
    }
    return s;
}
template <int ADVN>
__device__ __forceinline__ float upwind_y_t(const CoarseParams& p, const float* f, int I, int J) {
    const float c = f[I * p.NW + J];
    float s = 0.f;
    if (c <= 0.f) {
#pragma unroll
        for (int k = 0; k < ADVN; ++k) s += p.adv_f[k] * f[I * p.NW + wrap1(J + k, p.NW)];
    } else if (c > 0.f) {
#pragma unroll
        for (int k = 0; k < ADVN; ++k) s += p.adv_b[k] * f[I * p.NW + wrap1(J - (ADVN - 1) + k, p.NW)];
    }
    return s;
}
template <int DIFFN>
__device__ __forceinline__ float laplace_t(const CoarseParams& p, const float* f, int I, int J) {
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int k = 0; k < DIFFN; ++k) {
        sx += p.diff[k] * f[wrap1(I + p.diff_off + k, p.NH) * p.NW + J];
        sy += p.diff[k] * f[I * p.NW + wrap1(J + p.diff_off + k, p.NW)];
    }
    return sx * p.rdx2 + sy * p.rdy2;
}
// stencil_adjoint (coarse_adjoint.cu) with compile-time tap counts and reciprocal multiplies
template <int ADVN, int DIFFN>
__device__ __forceinline__ float stencil_adjoint_t(const CoarseParams& p, int id, const float* f,
                                                   const float* cu, const float* cv, const float* q,
                                                   int M, int J, float mu) {
    float acc = 0.f;
    constexpr int a = ADVN - 1;
    if (id & PPNN_RHS_ADV) {
#pragma unroll
        for (int k = 0; k <= a; ++k) {
            int I = wrap1(M - k, p.NH);
            int o = I * p.NW + J;
            if (f[o] <= 0.f) acc += p.adv_f[k] * (-cu[o] * q[o] * p.rdx);
            I = wrap1(M + a - k, p.NH);
            o = I * p.NW + J;
            if (f[o] > 0.f) acc += p.adv_b[k] * (-cu[o] * q[o] * p.rdx);
            int Jc = wrap1(J - k, p.NW);
            o = M * p.NW + Jc;
            if (f[o] <= 0.f) acc += p.adv_f[k] * (-cv[o] * q[o] * p.rdy);
            Jc = wrap1(J + a - k, p.NW);
            o = M * p.NW + Jc;
            if (f[o] > 0.f) acc += p.adv_b[k] * (-cv[o] * q[o] * p.rdy);
        }
    }
    if (id & PPNN_RHS_DIFF) {
        float sx = 0.f, sy = 0.f;
#pragma unroll
        for (int k = 0; k < DIFFN; ++k) {
            sx += p.diff[k] * q[wrap1(M - p.diff_off - k, p.NH) * p.NW + J];
            sy += p.diff[k] * q[M * p.NW + wrap1(J - p.diff_off - k, p.NW)];
        }
        acc += mu * (sx * p.rdx2 + sy * p.rdy2);
    }
    return acc;
}

#ifdef PPNN_TRACE
// dev-only CTA timeline (tools/trace_adjoint.py); never in the shipped library
__device__ unsigned long long g_trace_adj[8 * 8192];
__device__ __forceinline__ unsigned long long gtime_adj() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PPNN_STAMP(k) do { if (threadIdx.x == 0 && blockIdx.x < 8192) g_trace_adj[8 * blockIdx.x + (k)] = gtime_adj(); } while (0)
#else
#define PPNN_STAMP(k) do { } while (0)
#endif

constexpr int kPerF = 257;                                  // fine points of the meshes this kernel takes
constexpr int kPerApitch = 273;                             // A row pitch: odd, 16 zero columns after the 257
constexpr int kPerApad = 16;                                // zero floats in front of A (left pad of row 0)
constexpr int kPerSplit = 128;                              // row groups [0,128) | [128,257)
constexpr int kPerGstage = kAdjStageRows * kPerF + 8;       // floats per (slot, group) band buffer

template <int PER, int ADV>
constexpr size_t adjoint_periodic_smem() {
    constexpr int NC = (kPerF - 1) * ADV / PER, NN2 = NC * NC, AROWS = NC + 1 + 3;
    constexpr int ring = kAdjStages * 2 * kPerGstage, cst = 2 * NN2;
    return (size_t)(NN2 + kPerApad + ((AROWS * kPerApitch + 3) & ~3) + (ring > cst ? ring : cst)) * sizeof(float);
}

template <int PER, int ADV, bool IDENT, bool WITH_GNN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAdjThreads, 2)
coarse_adjoint_periodic_kernel(const CoarseParams p, const AdjParams a, const PhaseWeights pw, const float* g_end) {
    constexpr int F = kPerF, P = F - 1, NC = P * ADV / PER, CH = NC + 1, NN2 = NC * NC;
    constexpr int R = kAdjStageRows, NACC = ADV + 3;
    constexpr int AP = kPerApitch, AROWS = CH + 3;
    constexpr int GST = kPerGstage, SPLIT = kPerSplit;
    constexpr int NSTEPS = (F - SPLIT + R - 1) / R;              // 33: row group 1 has the extra row 256
    constexpr int OVL_LO = SPLIT * ADV / PER - 1;                // first of the 3 coarse rows both row groups feed
    constexpr unsigned KEEP_ROWS = scatter_row_mask<PER, ADV>();
    static_assert(SPLIT % PER == 0 && PER % R == 0 && P % PER == 0 && PER <= 16, "period layout");
    static_assert(kAdjWarps == 2 * (P / 32), "two row groups of eight 32-column chunks");

    extern __shared__ float4 smem4[];
    float* q = reinterpret_cast<float*>(smem4);                  // [NN2]; same offset in both CTAs of the cluster
    float* Az = q + NN2;                                         // kPerApad zeros, then A
    float* A = Az + kPerApad;                                    // [AROWS][AP]: CH rows + 3 overlap rows of group 1
    float* un = A + ((AROWS * AP + 3) & ~3);                     // union: ring [stages][2][GST] | cu, cv
    float* ring = un;
    float* cu = un;
    float* cv = cu + NN2;
    __shared__ __align__(8) uint64_t full[kAdjStages];
    __shared__ __align__(8) uint64_t empty[kAdjStages];
    __shared__ __align__(8) uint64_t cbar;                       // arrival of the saved coarse state
    __shared__ float red[32];

    const int b = blockIdx.x >> 1, f = blockIdx.x & 1;           // f == rank in the cluster
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr size_t plane = (size_t)F * F;
    const float* gf = a.g + ((size_t)b * 2 + f) * plane;         // this CTA's plane of grad_out
    float* of = a.gu + ((size_t)b * 2 + f) * plane;              // ... and of grad_u
    const float mu = __ldg(a.mu + b);
    const bool producer = warp == kAdjWarps;
    PPNN_STAMP(0);
#ifdef PPNN_TRACE
    if (tid == 0 && blockIdx.x < 8192) { unsigned sm; asm("mov.u32 %0, %smid;" : "=r"(sm)); g_trace_adj[8 * blockIdx.x + 7] = sm; }
#endif

    auto issue = [&](int t) {
        const int s = t % kAdjStages;
        uintptr_t a0[2]; uint32_t nb[2] = {0, 0}; uint32_t total = 0;
        const uintptr_t end_al = reinterpret_cast<uintptr_t>(g_end) & ~(uintptr_t)15;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int rs = g ? SPLIT : 0, re = g ? F : SPLIT;
            const int r0 = rs + t * R;
            if (r0 >= re) continue;
            const int nrows = min(R, re - r0);
            const uintptr_t ad = reinterpret_cast<uintptr_t>(gf + (size_t)r0 * F);
            a0[g] = ad & ~(uintptr_t)15;
            nb[g] = (uint32_t)((ad - a0[g]) + (size_t)nrows * F * 4 + 15) & ~15u;
            if (a0[g] + nb[g] > end_al) {                        // last copy of the tensor: stay inside it
                float* dst = ring + ((size_t)s * 2 + g) * GST;
                for (const float* qq = reinterpret_cast<const float*>(end_al); qq < g_end; ++qq)
                    dst[(reinterpret_cast<uintptr_t>(qq) - a0[g]) >> 2] = *qq;
                nb[g] = (uint32_t)(end_al - a0[g]);
            }
            total += nb[g];
        }
        mbar_expect_tx(&full[s], total);
#pragma unroll
        for (int g = 0; g < 2; ++g)
            if (nb[g]) bulk_g2s(ring + ((size_t)s * 2 + g) * GST, reinterpret_cast<const void*>(a0[g]), nb[g], &full[s]);
    };

    if (producer) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kAdjStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kAdjWarps); }
            mbar_init(&cbar, 1);
            fence_barrier_init();
            for (int t = 0; t < kAdjStages; ++t) issue(t);
        }
    } else {
        // clear the accumulator (and its zero pads) while the first bands fly
        float4* z = reinterpret_cast<float4*>(Az);
        constexpr int NZ = (kPerApad + ((AROWS * AP + 3) & ~3)) / 4;
        for (int idx = tid; idx < NZ; idx += kAdjWarps * 32) z[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    // ---- phase 1: stream grad_out; H-pass of bicubic^T with a compile-time period ------------
    if (producer) {
        if (lane == 0) {
            for (int t = kAdjStages; t < NSTEPS; ++t) {
                mbar_wait(&empty[t % kAdjStages], ((t / kAdjStages) - 1) & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t);
            }
        }
    } else {
        // chunk 0 also owns the duplicated endpoint column: a separate instantiation of the loop, so that the
        // other seven chunks carry no predicated-off instructions for it
        auto stream_rows = [&](auto dupl_tag) {
        constexpr bool DUPL = decltype(dupl_tag)::value;
        const int g = warp >> 3, chunk = warp & 7;
        const int rs = g ? SPLIT : 0;
        const int j = chunk * 32 + lane;                         // 0 .. 255: every lane owns a column
        const int lead = (int)(reinterpret_cast<uintptr_t>(gf) & 15) >> 2;
        const float* sg_base = ring + (size_t)g * GST + lead + j;
        float* op = of + (size_t)rs * F + j;
        float* Af = A + j;
        const uint64_t pol_keep = l2_policy_evict_last();
        // The last fine column sits on the last coarse node, so pde(i, 256) and pde(i, 0) are the same function
        // of the coarse increment: the lane of column 0 adds grad_out(i, 256) to its cotangent straight from the
        // ring (and writes that column's base value); A[.][256] stays zero.
        const bool dupl = DUPL && lane == 0;
        const float os_f = f ? a.os[1] : a.os[0];
        float* np = WITH_GNN ? a.gnn + ((size_t)b * 2 + f) * (size_t)P * P + (size_t)rs * P + j : nullptr;
        float gtop = 0.f;                                        // grad_out(256, j) (+ the corner for column 0): folds onto row 0
        if (WITH_GNN && g == 0) {
            gtop = ld_stream(gf + (size_t)P * F + j);
            if (DUPL && dupl) gtop += ld_stream(gf + (size_t)P * F + P);
        }
        float c[NACC];                                           // c[k]: coarse row base - 1 + k
#pragma unroll
        for (int k = 0; k < NACC; ++k) c[k] = 0.f;
        int base = g ? SPLIT * ADV / PER : 0;
        auto flush = [&](int I, float v) {
            const int r = clampi(I, 0, CH - 1);
            // each (coarse row, column) of A has ONE writer per row group; the groups meet on rows OVL_LO..OVL_LO+2,
            // which group 1 sends to the overlap rows A[CH..CH+2] (folded after the stream)
            const int rr = (g == 1 && r - OVL_LO < 3) ? CH + (r - OVL_LO) : r;
            Af[rr * AP] += v;
        };
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const float* sg = sg_base;
        int s = 0; uint32_t ph = 0;
        auto advance = [&]() {
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty_a + 8u * s);
            if (++s == kAdjStages) { s = 0; ph ^= 1u; sg = sg_base; } else { sg += 2 * GST; }
            op += (size_t)R * F;
            if (WITH_GNN) np += (size_t)R * P;
        };
        bool first = g == 0;                                     // the very first row of the plane is next
        for (int per = 0; per < SPLIT / PER; ++per) {
#pragma unroll
            for (int bnd = 0; bnd < PER / R; ++bnd) {
                mbar_wait_addr(full_a + 8u * s, ph);
#pragma unroll
                for (int qq = 0; qq < R; ++qq) {
                    const int r = bnd * R + qq;                  // compile-time after unrolling
                    const int o = (r * ADV) / PER;
                    float gv = sg[qq * F];
                    const float bv = IDENT ? gv : 0.f;
                    if ((KEEP_ROWS >> r) & 1) st_keep(op + (size_t)qq * F, bv, pol_keep);
                    else st_stream(op + (size_t)qq * F, bv);
                    if (DUPL && dupl) {
                        const float ge = sg[qq * F + P];
                        st_stream(op + (size_t)qq * F + P, IDENT ? ge : 0.f);
                        gv += ge;
                    }
                    if (WITH_GNN) {
                        // padBC_rd^T on the network output: column 256 folds onto column 0 (gv holds it for the lane of
                        // column 0), row 256 onto row 0 (gtop), times outstd
                        const float t0 = (r == 0 && first) ? gv + gtop : gv;
                        st_stream(np + (size_t)qq * P, t0 * os_f);
                    }
                    c[o] += pw.w[r][0] * gv; c[o + 1] += pw.w[r][1] * gv;
                    c[o + 2] += pw.w[r][2] * gv; c[o + 3] += pw.w[r][3] * gv;
                }
                advance();
            }
            first = false;
#pragma unroll
            for (int k = 0; k < ADV; ++k) flush(base - 1 + k, c[k]);
#pragma unroll
            for (int k = 0; k < NACC; ++k) c[k] = k + ADV < NACC ? c[k + ADV] : 0.f;
            base += ADV;
        }
        if (g == 1) {                                            // row 256 sits on the last coarse node: weights (0,1,0,0)
            mbar_wait_addr(full_a + 8u * s, ph);
            float gv = sg[0];
            st_stream(op, IDENT ? gv : 0.f);
            if (DUPL && dupl) {
                const float ge = sg[P];
                st_stream(op + P, IDENT ? ge : 0.f);
                gv += ge;
            }
            c[1] += gv;
        }
        flush(base - 1, c[0]); flush(base, c[1]); flush(base + 1, c[2]);
            };
        if ((warp & 7) == 0) stream_rows(std::true_type{}); else stream_rows(std::false_type{});
    }
    __syncthreads();              // A complete, ring dead, this CTA's base values of grad_u written
    PPNN_STAMP(1);

    // ---- coarse state: bulk copy of the saved one (flies during phase 2) --------------------
    const int rid_own = f ? p.rhs_v : p.rhs_u, rid_oth = f ? p.rhs_u : p.rhs_v;
    const bool both = ((rid_own | rid_oth) & PPNN_RHS_ADV) || ((rid_own | rid_oth) & ~3);
    if (a.coarse && tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&cbar, 2 * NN2 * 4);
        bulk_g2s(cu, a.coarse + (size_t)b * 2 * NN2, 2 * NN2 * 4, &cbar);
    }
    // ---- fold: overlap rows of row group 1, then the duplicated endpoint ROW (padBC_rd^T; the W-pass is linear) ----
    for (int jj = tid; jj < F; jj += kAdjThreads) {
#pragma unroll
        for (int k = 0; k < 3; ++k) A[(OVL_LO + k) * AP + jj] += A[(CH + k) * AP + jj];
        A[jj] += A[NC * AP + jj];
    }
    __syncthreads();

    // ---- phase 2: W-pass, compile-time period; item = (coarse row I, period k) -> columns ADV*k .. ADV*k+ADV-1 ----
    {
        constexpr int NK = P / PER;
        // fine columns (relative to PER*k) whose window [j0-1, j0+2] meets [0, ADV): floor(rp*ADV/PER) in [-2, ADV]
        constexpr int RP_LO = -(2 * PER / ADV);                  // smallest rp with floor(rp*ADV/PER) >= -2
        constexpr int RP_HI = ((ADV + 1) * PER - 1) / ADV;       // largest rp with floor(rp*ADV/PER) <= ADV
        for (int idx = tid; idx < NK * NC; idx += kAdjThreads) {
            const int k = small_div(idx, NC), I = idx - k * NC;
            const float* row = A + I * AP + PER * k;
            float acc[ADV];
#pragma unroll
            for (int m = 0; m < ADV; ++m) acc[m] = 0.f;
#pragma unroll
            for (int rp = RP_LO; rp <= RP_HI; ++rp) {
                const int phs = ((rp % PER) + PER) % PER;
                const int j0 = (rp * ADV - (rp < 0 ? PER - 1 : 0)) / PER;    // floor(rp * ADV / PER)
                const float x = row[rp];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int Jr = j0 - 1 + t;
                    if (Jr >= 0 && Jr < ADV) acc[Jr] += pw.w[phs][t] * x;
                }
            }
#pragma unroll
            for (int m = 0; m < ADV; ++m) q[I * NC + ADV * k + m] = p.dt * acc[m];
        }
    }
    if (!a.coarse) {
        // nothing saved: restriction gather of u into the dead ring (u is not streamed here: no L2 hint)
        const float* ub = a.u + (size_t)b * 2 * plane;
        const int lo = both ? 0 : f * NN2, hi = both ? 2 * NN2 : (f + 1) * NN2;
        gather_coarse<false>(p, ub, plane, F, lo, hi, cu, tid, kAdjThreads);
    }
    __syncthreads();
    // clamped borders, one thread per coarse row: windows reaching below column 0 pile up on column 0, and the
    // duplicated endpoint COLUMN C-1 (everything at or beyond it, clamped) folds onto column 0 as well
    if (tid < NC) {
        const float* row = A + tid * AP;
        float e = 0.f;
#pragma unroll
        for (int jj = 0; jj * ADV < PER; ++jj) e += pw.w[jj][0] * row[jj];          // j0 == 0: tap -1
#pragma unroll
        for (int jj = 0; jj * ADV <= 2 * PER; ++jj) {                                 // column 256 - jj: j0 >= NC - 2
            const int j0r = -((jj * ADV + PER - 1) / PER);                            // j0 - NC = -ceil(jj*ADV/PER)
            const int phs = (PER - jj % PER) % PER;
            float w = 0.f;
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (j0r - 1 + t >= 0) w += pw.w[phs][t];
            e += w * row[P - jj];
        }
        q[tid * NC] += p.dt * e;
    }
    if (a.coarse) mbar_wait(&cbar, 0);
    __syncthreads();
    PPNN_STAMP(2);

    const float* own = f ? cv : cu;
    const bool fast = p.adv_n == 2 && p.diff_n == 3;
    const bool want_mu = a.gmu && (rid_own & PPNN_RHS_DIFF);
    // grad_mu partial of this field (needs only local data): summed inside the scatter loop below -- or, when the
    // fields are coupled, here, where it fills the wait for the sibling CTA at the cluster barrier; reduced at the end
    float part = 0.f;
    if (want_mu && both) {
        for (int idx = tid; idx < NN2; idx += kAdjThreads) {
            const int I = small_div(idx, NC), J = idx - I * NC;
            part += (fast ? laplace_t<3>(p, own, I, J)
                          : central_x(p, own, I, J) * p.rdx2 + central_y(p, own, I, J) * p.rdy2) * q[idx];
        }
    }

    // ---- exchange of the coarse cotangents (only when a right-hand side couples the fields) ----
    const float* q_oth = A;
    if (both) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();                                           // q of both CTAs complete
        const float* q_peer = cluster.map_shared_rank(q, f ^ 1);
        for (int idx = tid; idx < NN2; idx += kAdjThreads) A[idx] = q_peer[idx];
        cluster.sync();                                           // both are done reading each other
    }

    PPNN_STAMP(3);
    // ---- phase 3 + 4: own field's coarse cotangent, scatter through D_down^T ------
    for (int idx = tid; idx < NN2; idx += kAdjThreads) {
        const int I = small_div(idx, NC), J = idx - I * NC;
        const float q_own = q[idx];
        // (same per-thread order of terms as the separate pass)
        if (want_mu && !both) part += (fast ? laplace_t<3>(p, own, I, J)
                                   : central_x(p, own, I, J) * p.rdx2 + central_y(p, own, I, J) * p.rdy2) * q_own;
        float cot = fast ? stencil_adjoint_t<2, 3>(p, rid_own, own, cu, cv, q, I, J, mu)
                         : stencil_adjoint(p, rid_own, own, cu, cv, q, I, J, mu);
        if (both) {
            const float q_o = q_oth[idx];
            const float u1 = cu[idx], v1 = cv[idx];
            const float* oth = f ? cu : cv;
            // point-wise factors: rhs_g contains (-u1*Dx(g))/dx - (v1*Dy(g))/dy
            if (f == 0) {        // d/du1
                if (rid_own & PPNN_RHS_ADV) cot += -(fast ? upwind_x_t<2>(p, own, I, J) : upwind_x(p, own, I, J)) * p.rdx * q_own;
                if (rid_oth & PPNN_RHS_ADV) cot += -(fast ? upwind_x_t<2>(p, oth, I, J) : upwind_x(p, oth, I, J)) * p.rdx * q_o;
            } else {             // d/dv1
                if (rid_own & PPNN_RHS_ADV) cot += -(fast ? upwind_y_t<2>(p, own, I, J) : upwind_y(p, own, I, J)) * p.rdy * q_own;
                if (rid_oth & PPNN_RHS_ADV) cot += -(fast ? upwind_y_t<2>(p, oth, I, J) : upwind_y(p, oth, I, J)) * p.rdy * q_o;
            }
            float du, dv;
            reaction_jac(rid_own & ~3, u1, v1, &du, &dv);
            cot += (f ? dv : du) * q_own;
            reaction_jac(rid_oth & ~3, u1, v1, &du, &dv);
            cot += (f ? dv : du) * q_o;
        }
        const BilinearTap t = bilinear_tap(p, I, J);
        const float h0 = 1.f - t.li, w0 = 1.f - t.lj;
        float* o = of + (size_t)t.i0 * F + t.j0;
        atomicAdd(o, h0 * w0 * cot);
        if (t.lj != 0.f) atomicAdd(o + t.dj, h0 * t.lj * cot);
        if (t.li != 0.f) {
            float* o2 = o + (size_t)t.di * F;
            atomicAdd(o2, t.li * w0 * cot);
            if (t.lj != 0.f) atomicAdd(o2 + t.dj, t.li * t.lj * cot);
        }
    }
    // grad_mu[b] = partial(u) + partial(v): fixed-order sum inside the CTA (lanes, then warps); the launcher zeroed
    // the entry and a sum of two floats does not depend on the order of the two atomics
    if (a.gmu) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (warp == 0) {
            float v = lane < (kAdjThreads >> 5) ? red[lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) atomicAdd(a.gmu + b, v);
        }
    }
    PPNN_STAMP(4);
}

}  // namespace ppnn

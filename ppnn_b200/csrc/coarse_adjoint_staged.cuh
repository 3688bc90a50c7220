// Staged adjoint of pde_du for meshes the staged forward kernel takes (the reference
// mesh, 257): the mirror image of coarse_step_staged.cuh.
//
// One CTA = one (sample, field) plane of grad_out; the two CTAs of a sample form a
// 2-CTA thread-block CLUSTER and exchange their coarse cotangents through
// distributed shared memory, so each CTA writes only its own plane of grad_u:
//
//   phase 1  a producer warp streams 4-row bands of grad_out (two row groups)
//            through a shared-memory ring with cp.async.bulk + mbarriers; consumer
//            warps (row group, 32-column chunk) write grad_u = grad_out (the identity
//            of u + pde_du(u)) or 0 and accumulate the H-pass of the bicubic
//            transpose in a 4-slot register window, flushed into A[I][j] (shared,
//            atomics only where the two row groups meet)
//   phase 2  W-pass of the bicubic transpose through an inverse weight table, fold of
//            the duplicated endpoint (transpose of padBC_rd), times dt  -> q_f
//   phase 0' restriction gather of the coarse state (recomputed, nothing was saved);
//            it reuses the ring's memory, which is dead by now
//   -------- cluster.sync: q_u and q_v are visible to both CTAs --------
//   phase 3  for the CTA's own field: transposed stencils of its right-hand side,
//            point-wise Jacobians of BOTH right-hand sides (the other one's q read
//            from the peer CTA's shared memory), grad_mu partial sums
//   phase 4  transpose of the bilinear restriction: <= 4 atomic adds per coarse point
//            into the CTA's own plane (after its own base values, same CTA => ordered)
#pragma once
#include <cooperative_groups.h>
#include "coarse_step_staged.cuh"

namespace ppnn {

namespace cg = cooperative_groups;

constexpr int kAdjStageRows = 4;
constexpr int kAdjStages = 4;
constexpr int kAdjWarps = 16;                       // consumer warps
constexpr int kAdjThreads = (kAdjWarps + 1) * 32;   // + producer warp

struct AdjStagedLayout {
    int gstage;     // floats per (slot, group) band buffer: rows*FW + 8, multiple of 4
    int split;      // first row of the second row group (multiple of kAdjStageRows); FH = single group
    int nsteps;
    int apitch;     // odd pitch of A rows
    int support;    // inverse W-pass table width
    int nchunk, wmain;
};

template <int FWC, bool IDENT, bool WITH_GNN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAdjThreads, 2)
coarse_adjoint_staged_kernel(const CoarseParams p, const AdjParams a, const AdjStagedLayout lay,
                             const float* g_end) {
    constexpr int R = kAdjStageRows;
    extern __shared__ float4 smem4[];
    float4* roww = smem4;                                        // [FH]
    int* rowi = reinterpret_cast<int*>(roww + p.FH);             // [FH]
    int* rowadv = rowi + p.FH;                                   // [2][nsteps]
    int* jstart = rowadv + 2 * lay.nsteps;                       // [CW]
    const int nn2 = p.NH * p.NW;
    float* wt = reinterpret_cast<float*>(rowi + ((p.FH + 2 * lay.nsteps + p.CW + 3) & ~3));   // [CW][support]
    float* q = wt + ((p.CW * lay.support + 3) & ~3);             // [nn2]   same offset in both CTAs of the cluster
    float* A = q + ((nn2 + 3) & ~3);                             // [CH + 4][apitch]: 4 extra rows = overlap buffer of row group 1
    float* un = A + (((size_t)(p.CH + 4) * lay.apitch + 3) & ~(size_t)3);   // union: ring [stages][2][gstage] | cu, cv
    float* ring = un;
    float* cu = un;
    float* cv = cu + nn2;
    __shared__ __align__(8) uint64_t full[kAdjStages];
    __shared__ __align__(8) uint64_t empty[kAdjStages];
    __shared__ float red[32];
    __shared__ float mu_partial;

    const int FW = FWC ? FWC : p.FW;
    const int b = blockIdx.x >> 1, f = blockIdx.x & 1;           // f == rank in the cluster
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)p.FH * FW;
    const float* ub = a.u + (size_t)b * 2 * plane;
    const float* gf = a.g + ((size_t)b * 2 + f) * plane;         // this CTA's plane of grad_out
    float* of = a.gu + ((size_t)b * 2 + f) * plane;              // ... and of grad_u
    const float mu = __ldg(a.mu + b);
    const int ngroups = lay.split < p.FH ? 2 : 1;
    const int nconsumers = min(ngroups * lay.nchunk, kAdjWarps);
    const bool producer = warp == kAdjWarps;

    auto issue = [&](int t) {
        const int s = t % kAdjStages;
        uintptr_t a0[2]; uint32_t nb[2] = {0, 0}; uint32_t total = 0;
        const uintptr_t end_al = reinterpret_cast<uintptr_t>(g_end) & ~(uintptr_t)15;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int rs = g ? lay.split : 0, re = g ? p.FH : lay.split;
            const int r0 = rs + t * R;
            if (g >= ngroups || r0 >= re) continue;
            const int nrows = min(R, re - r0);
            const uintptr_t ad = reinterpret_cast<uintptr_t>(gf + (size_t)r0 * FW);
            a0[g] = ad & ~(uintptr_t)15;
            nb[g] = (uint32_t)((ad - a0[g]) + (size_t)nrows * FW * 4 + 15) & ~15u;
            if (a0[g] + nb[g] > end_al) {                        // last copy of the tensor: stay inside it
                float* dst = ring + ((size_t)s * 2 + g) * lay.gstage;
                for (const float* qq = reinterpret_cast<const float*>(end_al); qq < g_end; ++qq)
                    dst[(reinterpret_cast<uintptr_t>(qq) - a0[g]) >> 2] = *qq;
                nb[g] = (uint32_t)(end_al - a0[g]);
            }
            total += nb[g];
        }
        mbar_expect_tx(&full[s], total);
#pragma unroll
        for (int g = 0; g < 2; ++g)
            if (nb[g]) bulk_g2s(ring + ((size_t)s * 2 + g) * lay.gstage, reinterpret_cast<const void*>(a0[g]), nb[g], &full[s]);
    };

    if (producer) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kAdjStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nconsumers); }
            fence_barrier_init();
            for (int t = 0; t < kAdjStages && t < lay.nsteps; ++t) issue(t);
        }
    } else {
        // tables and a cleared accumulator while the first bands fly
        build_row_tables(p, lay.split, lay.nsteps, R, roww, rowi, rowadv, tid, kAdjWarps * 32);
        for (int J = tid; J < p.CW; J += kAdjWarps * 32) {
            int js = (int)ceilf((float)(J - 2) / p.suw) - 1;
            jstart[J] = js < 0 ? 0 : js;
        }
        for (int idx = tid; idx < (p.CH + 4) * lay.apitch; idx += kAdjWarps * 32) A[idx] = 0.f;
    }
    __syncthreads();
    if (!producer) {
        for (int idx = tid; idx < p.CW * lay.support; idx += kAdjWarps * 32) {
            int J = small_div(idx, lay.support), k = idx - J * lay.support;
            int j = jstart[J] + k;
            float wsum = 0.f;
            if (j < FW) {
                float w[4]; int j0;
                cubic_entry(p.suw, j, &j0, w);
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (clampi(j0 - 1 + t, 0, p.CW - 1) == J) wsum += w[t];
            }
            wt[idx] = wsum;
        }
    }

    // ---- phase 1a: the duplicated-endpoint column (one element per row) ------------------
    // base values point-wise; its H-pass as a gather per coarse row (fixed summation order,
    // so the result does not depend on thread scheduling)
    // When the last fine column sits exactly on the last coarse node (257 -> 33 / 49), pde(i, FW-1)
    // and pde(i, 0) are the same function of the coarse increment (see coarse_step_staged.cuh), so
    // the lane that owns column 0 adds grad_out(i, FW-1) to its own cotangent straight from the
    // ring and this point-wise phase is skipped; A[.][FW-1] then stays zero.
    const float slast = p.suw * (float)(FW - 1);
    const bool coldup = lay.wmain < FW && floorf(slast) == slast && ext_index((int)slast, p.CW) == 0;
    if (!producer && lay.wmain < FW && !coldup) {
        const int jc = FW - 1;
        for (int i = tid; i < p.FH; i += kAdjWarps * 32)
            st_stream(of + (size_t)i * FW + jc, IDENT ? ld_stream(gf + (size_t)i * FW + jc) : 0.f);
        for (int I = tid; I < p.CH; I += kAdjWarps * 32) {
            // fine rows whose 4-tap window can reach coarse row I: floor(i*suh) in [I-2, I+1]
            int ilo = (int)ceilf((float)(I - 2) / p.suh) - 1, ihi = (int)floorf((float)(I + 2) / p.suh) + 1;
            ilo = max(ilo, 0); ihi = min(ihi, p.FH - 1);
            float sum = 0.f;
            for (int i = ilo; i <= ihi; ++i) {
                const int i0 = rowi[i];
                const float4 wx = roww[i];
                float w = 0.f;
                if (clampi(i0 - 1, 0, p.CH - 1) == I) w += wx.x;
                if (clampi(i0,     0, p.CH - 1) == I) w += wx.y;
                if (clampi(i0 + 1, 0, p.CH - 1) == I) w += wx.z;
                if (clampi(i0 + 2, 0, p.CH - 1) == I) w += wx.w;
                if (w != 0.f) sum += w * __ldg(gf + (size_t)i * FW + jc);
            }
            A[I * lay.apitch + jc] = sum;
        }
    }

    // ---- phase 1b: the other columns, streamed through the ring ---------------------
    if (producer) {
        if (lane == 0) {
            for (int t = kAdjStages; t < lay.nsteps; ++t) {
                mbar_wait(&empty[t % kAdjStages], ((t / kAdjStages) - 1) & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(t);
            }
        }
    } else if (warp < nconsumers) {
        const int g = warp / lay.nchunk;
        const int chunk = warp - g * lay.nchunk;
        const int rs = g ? lay.split : 0, re = g ? p.FH : lay.split;
        const int j = chunk * 32 + lane;
        const bool active = j < lay.wmain;
        const int jc = active ? j : lay.wmain - 1;
        const int lead = (int)(reinterpret_cast<uintptr_t>(gf) & 15) >> 2;
        const float* sg_base = ring + (size_t)g * lay.gstage + lead + jc;
        float* op = of + (size_t)rs * FW + jc;
        float* Af = A + jc;
        const float4* wrow = roww + rs;
        const int* advp = rowadv + g * lay.nsteps;
        int base = rowi[rs];
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;            // accumulators of coarse rows base-1 .. base+2
        // Every (coarse row, column) entry of A has ONE writer per row group, so flushes are plain
        // read-modify-writes (shared-memory float atomics are CAS loops).  The two groups only meet
        // on the 4 coarse rows around the split: group 1 sends those to the overlap rows A[CH..CH+3],
        // which phase 2 folds back.
        const int ovl_lo = max(rowi[lay.split < p.FH ? lay.split : 0] - 1, 0);
        auto flush = [&](int I, float v) {
            const int r = clampi(I, 0, p.CH - 1);
            const int rr = (g == 1 && r - ovl_lo < 4) ? p.CH + (r - ovl_lo) : r;
            if (active) Af[rr * lay.apitch] += v;
        };
        const bool dupl = coldup && j == 0;
        const bool GNN = WITH_GNN && coldup;                     // (the launcher only passes gnn for such meshes)
        const float os_f = f ? a.os[1] : a.os[0];
        float* np = GNN ? a.gnn + ((size_t)b * 2 + f) * (size_t)(p.FH - 1) * (FW - 1) + (size_t)rs * (FW - 1) + jc : nullptr;
        float gtop = 0.f;                                        // grad_out(FH-1, jc) (+ the corner for column 0): folds onto row 0
        if (GNN && g == 0 && active) {
            gtop = ld_stream(gf + (size_t)(p.FH - 1) * FW + jc);
            if (dupl) gtop += ld_stream(gf + (size_t)(p.FH - 1) * FW + FW - 1);
        }
        auto row_in = [&](int qq, int row, int advbits, const float* sg) {
            if ((advbits >> qq) & 1) {                           // retire the oldest slot (warp-uniform)
                flush(base - 1, c0);
                c0 = c1; c1 = c2; c2 = c3; c3 = 0.f;
                ++base;
            }
            float gv = sg[qq * FW];
            if (active) st_stream(op + (size_t)qq * FW, IDENT ? gv : 0.f);
            if (dupl) {
                const float ge = sg[qq * FW + FW - 1];
                st_stream(op + (size_t)qq * FW + FW - 1, IDENT ? ge : 0.f);
                gv += ge;
            }
            if (GNN) {
                // transpose of padBC_rd on the network output: column FW-1 folds onto column 0 (gv already holds
                // it for the lane of column 0), row FH-1 onto row 0 (gtop), times outstd
                if (active && jc < FW - 1 && row < p.FH - 1)
                    st_stream(np + (size_t)qq * (FW - 1), (row == 0 ? gv + gtop : gv) * os_f);
            }
            const float4 wx = wrow[qq];
            c0 += wx.x * gv; c1 += wx.y * gv; c2 += wx.z * gv; c3 += wx.w * gv;
        };
        const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty);
        const size_t gstep = (size_t)2 * lay.gstage;
        const float* sg = sg_base;
        int s = 0; uint32_t ph = 0;
        const int nfull = (re - rs) / R;
        auto advance = [&]() {
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty_a + 8u * s);
            if (++s == kAdjStages) { s = 0; ph ^= 1u; sg = sg_base; } else { sg += gstep; }
            wrow += R;
            op += (size_t)R * FW;
            if (WITH_GNN) np += (size_t)R * (FW - 1);
        };
        int t = 0;
        for (; t < nfull; ++t) {
            const int advbits = advp[t];
            mbar_wait_addr(full_a + 8u * s, ph);
#pragma unroll
            for (int qq = 0; qq < R; ++qq) row_in(qq, rs + t * R + qq, advbits, sg);
            advance();
        }
        for (; t < lay.nsteps; ++t) {                            // ragged last band; steps only the other group has
            const int r0 = rs + t * R;
            const int advbits = advp[t];
            mbar_wait_addr(full_a + 8u * s, ph);
            for (int qq = 0; r0 + qq < re; ++qq) row_in(qq, r0 + qq, advbits, sg);
            advance();
        }
        flush(base - 1, c0); flush(base, c1); flush(base + 1, c2); flush(base + 2, c3);
    }
    __syncthreads();              // A complete, ring dead, this CTA's base values of grad_u written
    if (ngroups == 2) {           // fold row group 1's overlap rows back
        const int ovl_lo = max(rowi[lay.split] - 1, 0);
        for (int idx = tid; idx < 4 * lay.apitch; idx += kAdjThreads) {
            const int k = small_div(idx, lay.apitch), j = idx - k * lay.apitch;
            if (ovl_lo + k < p.CH) A[(ovl_lo + k) * lay.apitch + j] += A[(p.CH + k) * lay.apitch + j];
        }
        __syncthreads();
    }

    // ---- phase 2 (W-pass, fold, dt) and phase 0' (gather into the dead ring) ----
    const int rid_own = f ? p.rhs_v : p.rhs_u, rid_oth = f ? p.rhs_u : p.rhs_v;
    const bool both = ((rid_own | rid_oth) & PPNN_RHS_ADV) || ((rid_own | rid_oth) & ~3);
    const int nthr = kAdjThreads;
    for (int idx = tid; idx < nn2; idx += nthr) {
        int Jp = small_div(idx, p.NH), Ip = idx - Jp * p.NH;                // lanes vary in I: A pitch is odd
        float sum = 0.f;
        for (int ej = 0; ej < 2; ++ej) {
            if (ej && Jp != 0) break;
            int J = ej ? p.NW : Jp;
            const float* w = wt + J * lay.support;
            int js = jstart[J];
            int cnt = min(lay.support, FW - js);
            for (int ei = 0; ei < 2; ++ei) {
                if (ei && Ip != 0) break;
                int I = ei ? p.NH : Ip;
                const float* row = A + I * lay.apitch + js;
                float t = 0.f;
#pragma unroll 4
                for (int k = 0; k < cnt; ++k) t += w[k] * row[k];
                sum += t;
            }
        }
        q[Ip * p.NW + Jp] = p.dt * sum;
    }
    {
        // (u is not streamed by this kernel: no L2 retention hint)
        const int lo = both ? 0 : f * nn2, hi = both ? 2 * nn2 : (f + 1) * nn2;
        if (a.coarse) { for (int idx = lo + tid; idx < hi; idx += nthr) cu[idx] = __ldg(a.coarse + (size_t)b * 2 * nn2 + idx); }
        else gather_coarse<false>(p, ub, plane, FW, lo, hi, cu, tid, nthr);
    }
    // grad_mu partial of this field (needs only local data)
    __syncthreads();
    const float* own = f ? cv : cu;
    float part = 0.f;
    if (a.gmu && (rid_own & PPNN_RHS_DIFF))
        for (int idx = tid; idx < nn2; idx += nthr) {
            int I = small_div(idx, p.NW), J = idx - I * p.NW;
            part += (central_x(p, own, I, J) / p.dx2 + central_y(p, own, I, J) / p.dy2) * q[idx];
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (warp == 0) {
        float v = lane < (kAdjThreads >> 5) ? red[lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) mu_partial = v;
    }

    // grad_mu[b] = partial(u) + partial(v): the launcher zeroed it; a sum of two floats does not depend on order
    if (tid == 0 && a.gmu) atomicAdd(a.gmu + b, mu_partial);

    // ---- exchange of the coarse cotangents (only when a right-hand side couples the fields) ----
    // The peer's q is copied into A (dead after the W-pass) between two back-to-back cluster barriers,
    // so that phases 3 and 4 run, and the CTAs retire, independently of each other.
    const float* q_oth = A;
    if (both) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();                                           // q of both CTAs complete
        const float* q_peer = cluster.map_shared_rank(q, f ^ 1);
        for (int idx = tid; idx < nn2; idx += nthr) A[idx] = q_peer[idx];
        cluster.sync();                                           // both are done reading each other
    }

    // ---- phase 3 + 4: own field's coarse cotangent, scatter through D_down^T ------
    for (int idx = tid; idx < nn2; idx += nthr) {
        int I = small_div(idx, p.NW), J = idx - I * p.NW;
        const float q_own = q[idx];
        float cot = stencil_adjoint(p, rid_own, own, cu, cv, q, I, J, mu);
        float q_o = 0.f;
        if (both) {
            q_o = q_oth[idx];
            const float u1 = cu[idx], v1 = cv[idx];
            const float* oth = f ? cu : cv;
            // point-wise factors: rhs_g contains (-u1*Dx(g))/dx - (v1*Dy(g))/dy
            if (f == 0) {        // d/du1
                if (rid_own & PPNN_RHS_ADV) cot += -upwind_x(p, own, I, J) / p.dx * q_own;
                if (rid_oth & PPNN_RHS_ADV) cot += -upwind_x(p, oth, I, J) / p.dx * q_o;
            } else {             // d/dv1
                if (rid_own & PPNN_RHS_ADV) cot += -upwind_y(p, own, I, J) / p.dy * q_own;
                if (rid_oth & PPNN_RHS_ADV) cot += -upwind_y(p, oth, I, J) / p.dy * q_o;
            }
            float du, dv;
            reaction_jac(rid_own & ~3, u1, v1, &du, &dv);
            cot += (f ? dv : du) * q_own;
            reaction_jac(rid_oth & ~3, u1, v1, &du, &dv);
            cot += (f ? dv : du) * q_o;
        }
        BilinearTap t = bilinear_tap(p, I, J);
        const float h0 = 1.f - t.li, w0 = 1.f - t.lj;
        float* o = of + (size_t)t.i0 * FW + t.j0;
        atomicAdd(o, h0 * w0 * cot);
        if (t.lj != 0.f) atomicAdd(o + t.dj, h0 * t.lj * cot);
        if (t.li != 0.f) {
            float* o2 = o + (size_t)t.di * FW;
            atomicAdd(o2, t.li * w0 * cot);
            if (t.lj != 0.f) atomicAdd(o2 + t.dj, t.li * t.lj * cot);
        }
    }
}

}  // namespace ppnn

// Fused coarse-grid PPNN step, forward:  pde_du (src/train2D.py:66-79) and the
// two kernels of one rollout step around the trainable network
// (src/train2D.py:230-245).  One CTA (16 warps, 2 resident per SM) owns one sample:
//
//   phase 1  bilinear restriction gather of both fields into shared memory
//            (mesh_convertor.down + [:, :, :-1, :-1])
//   phase 2  dt * rhs on the periodic N x N coarse grid, still in shared memory
//            (padbcx/padbcy wrap, Upwind1 / Central2 stencils, rhs.py algebra)
//   phase 3  stream the fine grid once.  The duplicated-endpoint column (one
//            element per row) is evaluated directly, one point per thread; the
//            other FW-1 columns are whole warps wide: every warp takes (field,
//            32-column chunk, row band) strips, lane = column, and walks down the
//            rows with a 4-row register window of W-interpolated coarse values
//            (padBC_rd + bicubic, W pass then H pass like torch), a rolling
//            kDepth-row register prefetch of the streamed operands and the
//            epilogue (pdeu | normalised net inputs | explicit update with the
//            network output) fused into the store.  The fine width is a template
//            parameter for the reference mesh (257) so that every streamed access
//            is base + immediate.
//
// HBM-bound by construction: per fine point the update kernel reads u and nn_out
// once and writes u_next once (12 B); the restriction gather re-reads
// (F-1)/(C-1)-strided rows only (L2 hits).  No tensor cores: 1-channel 3..9-tap
// stencils and a separable 4x4 interpolation have no dense contraction.
#include <cstdlib>
#include "coarse_step_staged.cuh"
namespace ppnn {

template <int MODE, int FWC>
__global__ void __launch_bounds__(kStepThreads, PPNN_MINCTA)
coarse_step_kernel(const CoarseParams p, const StepIO io, const StepPlan pl) {
    extern __shared__ float4 smem4[];
    float4* roww = smem4;                                       // [FH] H-pass weights
    int* rowi = reinterpret_cast<int*>(roww + p.FH);            // [FH] floor(i * suh)
    const int nn2 = p.NH * p.NW;
    float* cu = reinterpret_cast<float*>(rowi + ((p.FH + 3) & ~3));
    float* cv = cu + nn2;
    float* ru = cv + nn2;
    float* rv = ru + nn2;

    const int FW = FWC ? FWC : p.FW;                            // compile-time pitch when specialised
    const int PW = FW - 1, PH = p.FH - 1;                       // network grid (no endpoint)
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)p.FH * FW;
    const size_t nplane = (size_t)PH * PW;
    const float* ub = io.u + (size_t)b * 2 * plane;
    const float mu = __ldg(io.mu + b);

    // ---- phase 1: restriction gather + row tables --------------------------
#pragma unroll 4
    for (int idx = tid; idx < 2 * nn2; idx += kStepThreads) {
        int f = idx >= nn2;
        int r = idx - f * nn2;
        int I = r / p.NW, J = r - I * p.NW;
        cu[idx] = restrict_point(p, ub + f * plane, I, J);       // cv follows cu contiguously
        if (io.coarse) io.coarse[(size_t)b * 2 * nn2 + idx] = cu[idx];   // saved for the adjoint pass
    }
    for (int i = tid; i < p.FH; i += kStepThreads) {
        float w[4]; int i0;
        cubic_entry(p.suh, i, &i0, w);
        rowi[i] = i0;
        roww[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();

    // ---- phase 1': the "iter" variant of pde_du (post_process/test2D2.py:72-79): `inner` sequential Euler sub-steps
    // of the coarse state first -- u1 <- u1 + dt*rhsu(u1, v1), then v1 <- v1 + dt*rhsv(NEW u1, v1) on the halos of the
    // OLD v1, exactly as the script writes it
    for (int it = 0; it < io.inner; ++it) {
        for (int idx = tid; idx < nn2; idx += kStepThreads) {
            const int I = idx / p.NW, J = idx - I * p.NW;
            ru[idx] = cu[idx] + p.dt * eval_rhs(p, p.rhs_u, cu, cu, cv, I, J, mu);
        }
        __syncthreads();
        for (int idx = tid; idx < nn2; idx += kStepThreads) {
            const int I = idx / p.NW, J = idx - I * p.NW;
            rv[idx] = cv[idx] + p.dt * eval_rhs(p, p.rhs_v, cv, ru, cv, I, J, mu);
        }
        __syncthreads();
        for (int idx = tid; idx < 2 * nn2; idx += kStepThreads) cu[idx] = ru[idx];     // cv follows cu, rv follows ru
        __syncthreads();
    }

    // ---- phase 2: dt * rhs on the periodic coarse grid ----------------------
    coarse_rhs_phase(p, cu, cv, ru, rv, mu, tid, kStepThreads);
    __syncthreads();

    // ---- phase 3a: the duplicated-endpoint column, one point per thread -----
    const bool want_out = MODE == MODE_INPUTS ? io.out != nullptr : true;
    if (pl.wmain < FW && want_out) {
        const int jc = FW - 1;
        float wy[4]; int j0;
        cubic_entry(p.suw, jc, &j0, wy);
        const int c0 = ext_index(j0 - 1, p.CW), c1 = ext_index(j0, p.CW),
                  c2 = ext_index(j0 + 1, p.CW), c3 = ext_index(j0 + 2, p.CW);
        for (int idx = tid; idx < 2 * p.FH; idx += kStepThreads) {
            const int f = idx >= p.FH;
            const int i = idx - f * p.FH;
            const float* rf = f ? rv : ru;
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
                if (io.nn) {
                    const float n = ld_stream(io.nn + ((size_t)b * 2 + f) * nplane + (size_t)(i == PH ? 0 : i) * PW);
                    v = (n * (f ? io.s0[1] : io.s0[0]) + (f ? io.m0[1] : io.m0[0])) + uu + pde;
                }
                st_stream(io.out + o, v);
            } else {
                st_stream(io.out + o, pde);
            }
        }
    }

    // ---- phase 3b: (field, chunk, band) strips, one per warp at a time ------
    const int nstrip = 2 * pl.nchunk * pl.nband;
    for (int s = warp; s < nstrip; s += PPNN_STEP_WARPS) {
        // consecutive warps take consecutive column chunks of the same band
        const int chunk = s % pl.nchunk;
        const int t2 = s / pl.nchunk;
        const int f = t2 & 1;
        const int band = t2 >> 1;
        const int j = chunk * 32 + lane;
        const bool active = j < pl.wmain;
        const int jc = active ? j : pl.wmain - 1;
        const float* rf = f ? rv : ru;

        float wy[4]; int j0;
        cubic_entry(p.suw, jc, &j0, wy);
        const float* rc0 = rf + ext_index(j0 - 1, p.CW);
        const float* rc1 = rf + ext_index(j0, p.CW);
        const float* rc2 = rf + ext_index(j0 + 1, p.CW);
        const float* rc3 = rf + ext_index(j0 + 2, p.CW);
        auto yrow = [&](int I) -> float {     // W pass of row I of the extended coarse increment
            const int o = ext_index(I, p.CH) * p.NW;
            return rc0[o] * wy[0] + rc1[o] * wy[1] + rc2[o] * wy[2] + rc3[o] * wy[3];
        };
        const int r0 = band * pl.band_rows;
        const int r1 = band == pl.nband - 1 ? p.FH : r0 + pl.band_rows;
        int base = rowi[r0];
        float h0 = yrow(base - 1), h1 = yrow(base), h2 = yrow(base + 1), h3 = yrow(base + 2);
        // H pass of fine row ii; the window advances when floor(ii*suh) moves on (warp-uniform)
        auto prolong = [&](int ii) -> float {
            const int i0 = rowi[ii];
            const float4 wx = roww[ii];
            if (i0 != base) {
                if (i0 == base + 1) { h0 = h1; h1 = h2; h2 = h3; h3 = yrow(i0 + 2); }
                else { h0 = yrow(i0 - 1); h1 = yrow(i0); h2 = yrow(i0 + 1); h3 = yrow(i0 + 2); }
                base = i0;
            }
            return h0 * wx.x + h1 * wx.y + h2 * wx.z + h3 * wx.w;
        };

        const float m0 = f ? io.m0[1] : io.m0[0], s0 = f ? io.s0[1] : io.s0[0];
        const float m1 = f ? io.m1[1] : io.m1[0], s1 = f ? io.s1[1] : io.s1[0];
        const size_t foff = ((size_t)b * 2 + f) * plane + (size_t)r0 * FW + jc;
        const size_t noff = ((size_t)b * 2 + f) * nplane + (size_t)r0 * PW;
        const float* up = io.u + foff;                         // row-r0 pointers, bumped kDepth rows at a time
        float* op = io.out ? io.out + foff : nullptr;

        if (MODE == MODE_PDEU) {
            for (int ii = r0; ii < r1; ++ii, op += FW) {
                const float pde = prolong(ii);
                if (active) st_stream(op, pde);
            }
        } else if (MODE == MODE_UPDATE) {
            const bool has_nn = io.nn != nullptr;
            const int jn = jc == PW ? 0 : jc;                  // padBC_rd column fold
            const float* np = has_nn ? io.nn + noff + jn : nullptr;
            const int rn = min(r1, PH);                        // rows whose network row is the row itself
            // double-buffered batches of kDepth rows: batch B is in flight while batch A is consumed
            float ua[kDepth], na[kDepth], ub2[kDepth], nb[kDepth];
            auto load = [&](float (&uu)[kDepth], float (&nv)[kDepth], int first, int rowoff) {
#pragma unroll
                for (int q = 0; q < kDepth; ++q) {
                    const bool ok = active && first + q < rn;
                    uu[q] = ok ? ld_stream(up + (size_t)(rowoff + q) * FW) : 0.f;
                    nv[q] = (ok && has_nn) ? ld_stream(np + (size_t)(rowoff + q) * PW) : 0.f;
                }
            };
            auto consume = [&](const float (&uu)[kDepth], const float (&nv)[kDepth], int first, int rowoff) {
#pragma unroll
                for (int q = 0; q < kDepth; ++q) {
                    if (first + q < rn) {                       // warp-uniform
                        const float pde = prolong(first + q);
                        const float v = has_nn ? (nv[q] * s0 + m0) + uu[q] + pde : uu[q] + pde;
                        if (active) st_stream(op + (size_t)(rowoff + q) * FW, v);
                    }
                }
            };
            load(ua, na, r0, 0);
            for (int ii = r0; ii < rn; ii += 2 * kDepth) {
                load(ub2, nb, ii + kDepth, kDepth);
                consume(ua, na, ii, 0);
                load(ua, na, ii + 2 * kDepth, 2 * kDepth);
                consume(ub2, nb, ii + kDepth, kDepth);
                up += (size_t)2 * kDepth * FW; op += (size_t)2 * kDepth * FW;
                if (has_nn) np += (size_t)2 * kDepth * PW;
            }
            if (r1 > rn) {                                      // the endpoint row: network row 0 (padBC_rd)
                const float pde = prolong(PH);
                if (active) {
                    const size_t o = ((size_t)b * 2 + f) * plane + (size_t)PH * FW + jc;
                    const float u0 = ld_stream(io.u + o);
                    float v = u0 + pde;
                    if (has_nn) v = (ld_stream(io.nn + ((size_t)b * 2 + f) * nplane + jn) * s0 + m0) + u0 + pde;
                    st_stream(io.out + o, v);
                }
            }
        } else {   // MODE_INPUTS
            const float rs0 = 1.f / s0, rs1 = 1.f / s1;
            const bool colok = active && jc < PW;
            float* unp = io.u_norm + noff + jc;
            float* pnp = io.p_norm ? io.p_norm + noff + jc : nullptr;
            const int rn = min(r1, PH);                        // rows of the network grid in this band
            float ua[kDepth], ub2[kDepth];
            auto load = [&](float (&uu)[kDepth], int first, int rowoff) {
#pragma unroll
                for (int q = 0; q < kDepth; ++q)
                    uu[q] = (colok && first + q < rn) ? ld_stream(up + (size_t)(rowoff + q) * FW) : 0.f;
            };
            auto consume = [&](const float (&uu)[kDepth], int first, int rowoff) {
#pragma unroll
                for (int q = 0; q < kDepth; ++q) {
                    if (first + q < rn) {
                        const float pde = prolong(first + q);
                        if (op && active) st_stream(op + (size_t)(rowoff + q) * FW, pde);
                        if (colok) {
                            st_stream(unp + (size_t)(rowoff + q) * PW, div_by(uu[q] - m0, s0, rs0));
                            if (pnp) st_stream(pnp + (size_t)(rowoff + q) * PW, div_by(pde - m1, s1, rs1));
                        }
                    }
                }
            };
            load(ua, r0, 0);
            for (int ii = r0; ii < rn; ii += 2 * kDepth) {
                load(ub2, ii + kDepth, kDepth);
                consume(ua, ii, 0);
                load(ua, ii + 2 * kDepth, 2 * kDepth);
                consume(ub2, ii + kDepth, kDepth);
                up += (size_t)2 * kDepth * FW; unp += (size_t)2 * kDepth * PW;
                if (op) op += (size_t)2 * kDepth * FW;
                if (pnp) pnp += (size_t)2 * kDepth * PW;
            }
            if (r1 > rn && op) {                                // endpoint row exists only in pdeu
                const float pde = prolong(PH);
                if (active) st_stream(io.out + ((size_t)b * 2 + f) * plane + (size_t)PH * FW + jc, pde);
            }
        }
    }
}

// Coarse-only solver increment (post_process/test_2D_csolver.py:56-66): the
// C -> C 'down' is the identity, so the gather is a strided copy.
__global__ void __launch_bounds__(256)
coarse_increment_kernel(const CoarseParams p, const float* __restrict__ uc, const float* __restrict__ mu_,
                        float* __restrict__ out, int add_state) {
    extern __shared__ float smem[];
    const int nn2 = p.NH * p.NW;
    float* cu = smem;
    float* cv = cu + nn2;
    const int b = blockIdx.x;
    const size_t plane = (size_t)p.CH * p.CW;
    const float* ub = uc + (size_t)b * 2 * plane;
    const float mu = __ldg(mu_ + b);
    for (int idx = threadIdx.x; idx < 2 * nn2; idx += blockDim.x) {
        int f = idx >= nn2;
        int r = idx - f * nn2;
        int I = r / p.NW, J = r - I * p.NW;
        cu[idx] = __ldg(ub + f * plane + (size_t)I * p.CW + J);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2 * (int)plane; idx += blockDim.x) {
        int f = idx >= (int)plane;
        int r = idx - f * (int)plane;
        int I = r / p.CW, J = r - I * p.CW;
        int Ie = I == p.NH ? 0 : I, Je = J == p.NW ? 0 : J;
        float v = p.dt * eval_rhs(p, f ? p.rhs_v : p.rhs_u, f ? cv : cu, cu, cv, Ie, Je, mu);
        if (add_state) v += __ldg(ub + idx);
        out[(size_t)b * 2 * plane + idx] = v;
    }
}

static size_t step_smem_bytes(const CoarseParams& p) {
    return (size_t)p.FH * sizeof(float4) + (size_t)((p.FH + 3) & ~3) * sizeof(int) +
           (size_t)4 * p.NH * p.NW * sizeof(float);
}

template <int MODE, int FWC>
static int launch_step_fw(const CoarseParams& p, const StepIO& io, const StepPlan& pl, int B, size_t smem, void* stream) {
    auto kern = coarse_step_kernel<MODE, FWC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<B, kStepThreads, smem, (cudaStream_t)stream>>>(p, io, pl);
    return launch_status();
}

// Ring depths of the compile-time-period instantiations (they carry no run-time tables, which leaves room for more
// stages next to the 27 / 12 KB of coarse state at two CTAs per SM).  Same-box timings, 1920 samples: stages of
// 16.5 KB (u + network operand) gain nothing beyond 4 (0.4716 / 0.4715 ms with 4 / 5); the single-operand modes
// (8.3 KB per stage) want 8: PDE-only update 0.357 -> 0.334 ms (10 stages: 0.354, the elected lane issues them
// serially at entry).
#ifndef PPNN_NST_NN49
#define PPNN_NST_NN49 4
#endif
#ifndef PPNN_NST_NN33
#define PPNN_NST_NN33 4
#endif
#ifndef PPNN_NST_U
#define PPNN_NST_U 8
#endif
constexpr int kNstNN49 = PPNN_NST_NN49, kNstNN33 = PPNN_NST_NN33, kNstU = PPNN_NST_U;

template <int MODE>
static int launch_step(const ppnn_coarse_cfg* cfg, const StepIO& io, int B, void* stream) {
    CoarseParams p;
    int st = fill_coarse_params(cfg, &p);
    if (st != PPNN_OK) return st;
    if (B < 0) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    size_t smem = step_smem_bytes(p);
    if (smem > 200 * 1024) return PPNN_E_TOOBIG;
    StepPlan pl;
    // the duplicated-endpoint column is evaluated point-wise when the rest is whole warps wide
    pl.wmain = ((p.FW - 1) % 32 == 0) ? p.FW - 1 : p.FW;
    pl.nchunk = (pl.wmain + 31) / 32;
    // bands so that 2 * nchunk * nband strips fill the warps evenly (2 strips per warp)
    pl.nband = (2 * PPNN_STEP_WARPS + 2 * pl.nchunk - 1) / (2 * pl.nchunk);
    int max_band = (p.FH + 4 * kDepth - 1) / (4 * kDepth);
    if (pl.nband > max_band) pl.nband = max_band;
    if (pl.nband < 1) pl.nband = 1;
    pl.band_rows = (p.FH - 1) / pl.nband;                      // the last band takes the remainder
    if (pl.band_rows < 1) { pl.band_rows = p.FH; pl.nband = 1; }
    // TMA-staged path: operands through a shared-memory ring (see coarse_step_staged.cuh)
    const bool nn_ok = !io.nn || (reinterpret_cast<uintptr_t>(io.nn) & 15) == 0;
    if (io.inner == 0 && pl.nchunk <= PPNN_STEP_WARPS && B < (1 << 30) && (p.FW - 1) % 4 == 0 && p.CH <= p.FH && nn_ok &&
        !PPNN_ENV_OFF("PPNN_NO_TMA")) {
        StagedLayout lay;
        lay.ustage = MODE == MODE_PDEU ? 0 : (kStageRows * p.FW + 8 + 3) & ~3;
        lay.nstage = (MODE == MODE_UPDATE && io.nn) ? kStageRows * (p.FW - 1) : 0;
        // two row groups of 8 chunks each when the mesh is narrow enough for both to fit the CTA
        lay.split = p.FH;
        if (2 * pl.nchunk <= PPNN_STEP_WARPS && p.FH >= 4 * kStageRows)
            lay.split = ((p.FH / 2 + kStageRows - 1) / kStageRows) * kStageRows;
        {
            int n0 = (lay.split + kStageRows - 1) / kStageRows;
            int n1 = (p.FH - lay.split + kStageRows - 1) / kStageRows;
            lay.nsteps = n0 > n1 ? n0 : n1;
        }
        const float* u_end = io.u + (size_t)B * 2 * p.FH * p.FW;
        const bool has_nn = MODE == MODE_UPDATE && io.nn != nullptr;
        void (*kern)(CoarseParams, StepIO, StepPlan, StagedLayout, const float*, PhaseWeights);
        // the reference's own meshes: prolongation period unrolled at compile time (16 fine rows per 3 coarse
        // rows for 257 -> 49, 8 per 1 for 257 -> 33), no run-time tables, and the ring as deep as two resident
        // CTAs allow (kNst*: stages of 16.5 KB with the network operand, 8.3 KB without)
        const bool square = p.FH == 257 && p.FW == 257 && p.CH == p.CW && lay.split == 128 && pl.nchunk == 8 &&
                            !PPNN_ENV_OFF("PPNN_NO_PERIODIC");
        int per = 0, nst = kStages;
        if (square && p.CH == 49) {
            per = 16; nst = has_nn ? kNstNN49 : kNstU;
            kern = has_nn ? coarse_step_staged_kernel<MODE, 257, true, 16, 3, kNstNN49> : coarse_step_staged_kernel<MODE, 257, false, 16, 3, kNstU>;
        } else if (square && p.CH == 33) {
            per = 8; nst = has_nn ? kNstNN33 : kNstU;
            kern = has_nn ? coarse_step_staged_kernel<MODE, 257, true, 8, 1, kNstNN33> : coarse_step_staged_kernel<MODE, 257, false, 8, 1, kNstU>;
        } else if (p.FW == 257) {
            kern = has_nn ? coarse_step_staged_kernel<MODE, 257, true> : coarse_step_staged_kernel<MODE, 257, false>;
        } else {
            kern = has_nn ? coarse_step_staged_kernel<MODE, 0, true> : coarse_step_staged_kernel<MODE, 0, false>;
        }
        const size_t tables = per ? 0 : (size_t)p.FH * sizeof(float4) + (size_t)((p.FH + 2 * lay.nsteps + 3) & ~3) * sizeof(int);
        const size_t smem2 = tables + (size_t)((3 * p.NH * p.NW + 3) & ~3) * sizeof(float) +
                             (size_t)nst * 2 * (lay.ustage + lay.nstage) * sizeof(float);
        if (smem2 <= 200 * 1024) {
            const PhaseWeights pw = make_phase_weights(p.suh, per);
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e != cudaSuccess) return (int)e;
            kern<<<2 * B, kStagedThreads, smem2, (cudaStream_t)stream>>>(p, io, pl, lay, u_end, pw);
            return launch_status();
        }
    }
    if (p.FW == 257) return launch_step_fw<MODE, 257>(p, io, pl, B, smem, stream);
    return launch_step_fw<MODE, 0>(p, io, pl, B, smem, stream);
}

}  // namespace ppnn

using namespace ppnn;

extern "C" int ppnn_pde_du_fwd_save(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                                    float* pdeu, float* coarse_state, int B, void* stream) {
    if (!u || !mu || !pdeu) return PPNN_E_NULL;
    StepIO io = {};
    io.u = u; io.mu = mu; io.out = pdeu; io.coarse = coarse_state;
    return launch_step<MODE_PDEU>(cfg, io, B, stream);
}

extern "C" int ppnn_pde_du_fwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                               float* pdeu, int B, void* stream) {
    return ppnn_pde_du_fwd_save(cfg, u, mu, pdeu, nullptr, B, stream);
}

extern "C" int ppnn_pde_du_iter_fwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                                    float* pdeu, int B, int inner, void* stream) {
    if (!u || !mu || !pdeu) return PPNN_E_NULL;
    if (inner < 0) return PPNN_E_SHAPE;
    StepIO io = {};
    io.u = u; io.mu = mu; io.out = pdeu; io.inner = inner;
    return launch_step<MODE_PDEU>(cfg, io, B, stream);
}

extern "C" int ppnn_step_inputs(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                                const float* inmean, const float* instd,
                                const float* pdemean, const float* pdestd,
                                float* u_norm, float* pdeu_norm, float* pdeu, int B, void* stream) {
    if (!u || !mu || !u_norm || !inmean || !instd) return PPNN_E_NULL;
    if (pdeu_norm && (!pdemean || !pdestd)) return PPNN_E_NULL;
    StepIO io = {};
    io.u = u; io.mu = mu; io.out = pdeu; io.u_norm = u_norm; io.p_norm = pdeu_norm;
    for (int c = 0; c < 2; ++c) {
        io.m0[c] = inmean[c]; io.s0[c] = instd[c];
        io.m1[c] = pdeu_norm ? pdemean[c] : 0.f; io.s1[c] = pdeu_norm ? pdestd[c] : 1.f;
    }
    return launch_step<MODE_INPUTS>(cfg, io, B, stream);
}

extern "C" int ppnn_step_update_save(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                                     const float* nn_out, const float* outmean, const float* outstd,
                                     float* u_next, float* coarse_state, int B, void* stream) {
    if (!u || !mu || !u_next) return PPNN_E_NULL;
    if (nn_out && (!outmean || !outstd)) return PPNN_E_NULL;
    if (u_next == u) return PPNN_E_MODE;      // planes of one sample run on different CTAs: no in-place update
    StepIO io = {};
    io.u = u; io.mu = mu; io.nn = nn_out; io.out = u_next; io.coarse = coarse_state;
    for (int c = 0; c < 2; ++c) { io.m0[c] = nn_out ? outmean[c] : 0.f; io.s0[c] = nn_out ? outstd[c] : 1.f; }
    return launch_step<MODE_UPDATE>(cfg, io, B, stream);
}

extern "C" int ppnn_step_update(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                                const float* nn_out, const float* outmean, const float* outstd,
                                float* u_next, int B, void* stream) {
    return ppnn_step_update_save(cfg, u, mu, nn_out, outmean, outstd, u_next, nullptr, B, stream);
}

extern "C" int ppnn_coarse_increment(const ppnn_coarse_cfg* cfg, const float* uc, const float* mu,
                                     float* out, int B, int add_state, void* stream) {
    if (!uc || !mu || !out) return PPNN_E_NULL;
    ppnn_coarse_cfg c2;
    if (!cfg) return PPNN_E_NULL;
    c2 = *cfg; c2.FH = cfg->CH; c2.FW = cfg->CW;
    CoarseParams p;
    int st = fill_coarse_params(&c2, &p);
    if (st != PPNN_OK) return st;
    if (B < 0) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    size_t smem = (size_t)2 * p.NH * p.NW * sizeof(float);
    if (smem > 200 * 1024) return PPNN_E_TOOBIG;
    cudaError_t e = cudaFuncSetAttribute(coarse_increment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    coarse_increment_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(p, uc, mu, out, add_state);
    return launch_status();
}

// Adjoint (vector-Jacobian product) of pde_du (src/train2D.py:66-79), so that
// gradients flow through autoregressive rollouts.  SURVEY.md appendix A.4:
//
//   d pde_du / d u = M_up . E . J_rhs . W . D_down
//
// One CTA owns one sample and applies the transposes right to left on
// grad_out:
//   phase 0  recompute the coarse state (restriction gather) -- nothing is saved
//            by the forward pass
//   phase 1  stream grad_out once: each warp owns (field, 32-column chunk) and
//            walks all rows, accumulating the H-pass of the bicubic transpose in
//            a 4-slot register window that is flushed into A[I][j] (shared);
//            the same pass writes grad_u = grad_out (add_identity) or 0
//   phase 2  W-pass of the bicubic transpose through an inverse weight table,
//            folded over the duplicated endpoint (transpose of padBC_rd), times dt
//   phase 3  transposed stencils / pointwise Jacobians of rhs.py on the periodic
//            coarse grid (upwind masks are constants, as for autograd), grad_mu
//   phase 4  transpose of the bilinear restriction: <= 4 atomic adds per coarse
//            point into grad_u (no two coarse points share a fine point when
//            (F-1)/(C-1) >= 2, so the result is deterministic there)
#include "coarse_common.cuh"

namespace ppnn {

constexpr int kAdjUnroll = 8;
constexpr int kMaxSupport = 96;     // fine columns that can touch one coarse column

struct AdjParams {
    const float* u; const float* mu; const float* g;
    const float* coarse;   // restricted coarse state [B,2,NH,NW] saved by the forward pass, or null (then u is gathered)
    float* gu; float* gmu;
    float* gnn;        // grad of the network output [B,2,FH-1,FW-1] (padBC_rd^T of grad_out, times outstd) or null
    float os[2];       // outstd per channel
    int add_identity;
    int apitch;        // pitch of A rows (odd => conflict-free column walks)
    int support;       // table width (<= kMaxSupport)
};

// Pointwise Jacobian of the reaction terms: d react / d u1, d react / d v1.
__device__ __forceinline__ void reaction_jac(int react, float u1, float v1, float* du, float* dv) {
    float uv = u1 * u1 + v1 * v1;
    switch (react) {
    case PPNN_REACT_BFU: *du = 1.f - uv - 2.f * u1 * u1 + 2.f * u1 * v1; *dv = uv - 2.f * u1 * v1 + 2.f * v1 * v1; break;
    case PPNN_REACT_BFV: *du = -uv - 2.f * u1 * u1 - 2.f * u1 * v1;      *dv = 1.f - uv - 2.f * u1 * v1 - 2.f * v1 * v1; break;
    case PPNN_REACT_RDU: *du = 1.f - 3.f * u1 * u1; *dv = -1.f; break;
    case PPNN_REACT_RDV: *du = 0.25f; *dv = -0.25f; break;
    default: *du = 0.f; *dv = 0.f;
    }
}

// Transposed stencils acting on the cotangent q of dt*rhs_f at coarse point
// (M, J): returns d/d f[M,J] through the derivative operators only.
__device__ __forceinline__ float stencil_adjoint(const CoarseParams& p, int id, const float* f,
                                                 const float* cu, const float* cv, const float* q,
                                                 int M, int J, float mu) {
    float acc = 0.f;
    const int a = p.adv_n - 1;
    if (id & PPNN_RHS_ADV) {
        const float ix = 1.f / p.dx, iy = 1.f / p.dy;
        for (int k = 0; k <= a; ++k) {
            // forward-biased rows: Dx(f)[I] reads f[I+k]  => I = M-k
            int I = wrap1(M - k, p.NH);
            int o = I * p.NW + J;
            if (f[o] <= 0.f) acc += p.adv_f[k] * (-cu[o] * q[o] * ix);
            // backward-biased: Dx(f)[I] reads f[I-a+k]  => I = M+a-k
            I = wrap1(M + a - k, p.NH);
            o = I * p.NW + J;
            if (f[o] > 0.f) acc += p.adv_b[k] * (-cu[o] * q[o] * ix);
            int Jc = wrap1(J - k, p.NW);
            o = M * p.NW + Jc;
            if (f[o] <= 0.f) acc += p.adv_f[k] * (-cv[o] * q[o] * iy);
            Jc = wrap1(J + a - k, p.NW);
            o = M * p.NW + Jc;
            if (f[o] > 0.f) acc += p.adv_b[k] * (-cv[o] * q[o] * iy);
        }
    }
    if (id & PPNN_RHS_DIFF) {
        float sx = 0.f, sy = 0.f;
        for (int k = 0; k < p.diff_n; ++k) {
            sx += p.diff[k] * q[wrap1(M - p.diff_off - k, p.NH) * p.NW + J];
            sy += p.diff[k] * q[M * p.NW + wrap1(J - p.diff_off - k, p.NW)];
        }
        acc += mu * (sx / p.dx2 + sy / p.dy2);
    }
    return acc;
}

__global__ void __launch_bounds__(576)
coarse_adjoint_kernel(const CoarseParams p, const AdjParams a) {
    extern __shared__ float smem[];
    const int nn2 = p.NH * p.NW;
    float* cu = smem;
    float* cv = cu + nn2;
    float* qu = cv + nn2;
    float* qv = qu + nn2;
    float4* roww = reinterpret_cast<float4*>(qv + nn2);           // [FH]
    int* rowi = reinterpret_cast<int*>(roww + p.FH);              // [FH]
    int* jstart = rowi + p.FH;                                    // [CW]
    float* wt = reinterpret_cast<float*>(jstart + p.CW);          // [CW][support]
    float* A = wt + p.CW * a.support;                             // [2][CH][apitch]
    __shared__ float red[32];

    const int b = blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t plane = (size_t)p.FH * p.FW;
    const float* ub = a.u + (size_t)b * 2 * plane;
    const float* gb = a.g + (size_t)b * 2 * plane;
    float* gub = a.gu + (size_t)b * 2 * plane;
    const float mu = __ldg(a.mu + b);

    // ---- phase 0: coarse state, tables, clear A ----------------------------
    for (int idx = tid; idx < 2 * nn2; idx += nthr) {
        if (a.coarse) { cu[idx] = __ldg(a.coarse + (size_t)b * 2 * nn2 + idx); continue; }
        int f = idx >= nn2;
        int r = idx - f * nn2;
        int I = r / p.NW, J = r - I * p.NW;
        cu[idx] = restrict_point(p, ub + f * plane, I, J);
    }
    for (int i = tid; i < p.FH; i += nthr) {
        float w[4]; int i0;
        cubic_entry(p.suh, i, &i0, w);
        rowi[i] = i0;
        roww[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
    // inverse W-pass table: for extended coarse column J the fine columns
    // j in [jstart, jstart+support) and the weight with which j feeds J
    for (int J = tid; J < p.CW; J += nthr) {
        int js = (int)ceilf((float)(J - 2) / p.suw) - 1;
        jstart[J] = js < 0 ? 0 : js;
    }
    for (int idx = tid; idx < 2 * p.CH * a.apitch; idx += nthr) A[idx] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < p.CW * a.support; idx += nthr) {
        int J = idx / a.support, k = idx - J * a.support;
        int j = jstart[J] + k;
        float wsum = 0.f;
        if (j < p.FW) {
            float w[4]; int j0;
            cubic_entry(p.suw, j, &j0, w);
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (clampi(j0 - 1 + t, 0, p.CW - 1) == J) wsum += w[t];
        }
        wt[idx] = wsum;
    }

    // ---- phase 1: H-pass of the bicubic transpose while streaming grad_out --
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int nchunk = (p.FW + 31) >> 5;
    for (int s = warp; s < 2 * nchunk; s += nwarp) {
        int chunk = s % nchunk, f = s / nchunk;
        int j = chunk * 32 + lane;
        bool active = j < p.FW;
        int jc = active ? j : p.FW - 1;
        const float* gp = gb + f * plane + jc;
        float* op = gub + f * plane + jc;
        float* Af = A + (size_t)f * p.CH * a.apitch + jc;
        int base = rowi[0];
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;            // slots for rows base-1 .. base+2
        for (int i = 0; i < p.FH; i += kAdjUnroll) {
            float gv[kAdjUnroll];
#pragma unroll
            for (int q = 0; q < kAdjUnroll; ++q)
                gv[q] = (active && i + q < p.FH) ? ld_stream(gp + (size_t)(i + q) * p.FW) : 0.f;
#pragma unroll
            for (int q = 0; q < kAdjUnroll; ++q) {
                int ii = i + q;
                if (ii >= p.FH) break;
                int i0 = rowi[ii];
                while (i0 != base) {                            // warp-uniform: retire the oldest slot
                    if (active) Af[clampi(base - 1, 0, p.CH - 1) * a.apitch] += c0;
                    c0 = c1; c1 = c2; c2 = c3; c3 = 0.f;
                    ++base;
                }
                float4 wx = roww[ii];
                c0 += wx.x * gv[q]; c1 += wx.y * gv[q]; c2 += wx.z * gv[q]; c3 += wx.w * gv[q];
                if (active) st_stream(op + (size_t)ii * p.FW, a.add_identity ? gv[q] : 0.f);
            }
        }
        if (active) {
            Af[clampi(base - 1, 0, p.CH - 1) * a.apitch] += c0;
            Af[clampi(base,     0, p.CH - 1) * a.apitch] += c1;
            Af[clampi(base + 1, 0, p.CH - 1) * a.apitch] += c2;
            Af[clampi(base + 2, 0, p.CH - 1) * a.apitch] += c3;
        }
    }
    __syncthreads();

    // ---- phase 2: W-pass + endpoint fold + dt -----------------------------
    // q_f[I', J'] = dt * sum over extended (I, J) folding onto (I', J')
    for (int idx = tid; idx < 2 * nn2; idx += nthr) {
        int f = idx >= nn2;
        int r = idx - f * nn2;
        int Jp = r / p.NH, Ip = r - Jp * p.NH;                  // lanes vary in I: A pitch is odd
        const float* Af = A + (size_t)f * p.CH * a.apitch;
        float sum = 0.f;
        for (int ej = 0; ej < 2; ++ej) {
            if (ej && Jp != 0) break;
            int J = ej ? p.NW : Jp;
            const float* w = wt + J * a.support;
            int js = jstart[J];
            int cnt = min(a.support, p.FW - js);
            for (int ei = 0; ei < 2; ++ei) {
                if (ei && Ip != 0) break;
                int I = ei ? p.NH : Ip;
                const float* row = Af + I * a.apitch + js;
                float t = 0.f;
                for (int k = 0; k < cnt; ++k) t += w[k] * row[k];
                sum += t;
            }
        }
        (f ? qv : qu)[Ip * p.NW + Jp] = p.dt * sum;
    }
    __syncthreads();

    // ---- phase 3 + 4: coarse Jacobian transpose, scatter through D_down^T ---
    float gmu_part = 0.f;
    for (int idx = tid; idx < nn2; idx += nthr) {
        int I = idx / p.NW, J = idx - I * p.NW;
        float u1 = cu[idx], v1 = cv[idx], q_u = qu[idx], q_v = qv[idx];
        float ubar = stencil_adjoint(p, p.rhs_u, cu, cu, cv, qu, I, J, mu);
        float vbar = stencil_adjoint(p, p.rhs_v, cv, cu, cv, qv, I, J, mu);
        if (p.rhs_u & PPNN_RHS_ADV) {
            ubar += -upwind_x(p, cu, I, J) / p.dx * q_u;
            vbar += -upwind_y(p, cu, I, J) / p.dy * q_u;
        }
        if (p.rhs_v & PPNN_RHS_ADV) {
            ubar += -upwind_x(p, cv, I, J) / p.dx * q_v;
            vbar += -upwind_y(p, cv, I, J) / p.dy * q_v;
        }
        float du, dv;
        reaction_jac(p.rhs_u & ~3, u1, v1, &du, &dv);
        ubar += du * q_u; vbar += dv * q_u;
        reaction_jac(p.rhs_v & ~3, u1, v1, &du, &dv);
        ubar += du * q_v; vbar += dv * q_v;
        if (p.rhs_u & PPNN_RHS_DIFF)
            gmu_part += (central_x(p, cu, I, J) / p.dx2 + central_y(p, cu, I, J) / p.dy2) * q_u;
        if (p.rhs_v & PPNN_RHS_DIFF)
            gmu_part += (central_x(p, cv, I, J) / p.dx2 + central_y(p, cv, I, J) / p.dy2) * q_v;

        BilinearTap t = bilinear_tap(p, I, J);
        float h0 = 1.f - t.li, w0 = 1.f - t.lj;
#pragma unroll
        for (int f = 0; f < 2; ++f) {
            float gbar = f ? vbar : ubar;
            float* o = gub + f * plane + (size_t)t.i0 * p.FW + t.j0;
            atomicAdd(o, h0 * w0 * gbar);
            if (t.lj != 0.f) atomicAdd(o + t.dj, h0 * t.lj * gbar);
            if (t.li != 0.f) {
                float* o2 = o + (size_t)t.di * p.FW;
                atomicAdd(o2, t.li * w0 * gbar);
                if (t.lj != 0.f) atomicAdd(o2 + t.dj, t.li * t.lj * gbar);
            }
        }
    }
    if (a.gmu) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmu_part += __shfl_xor_sync(0xffffffffu, gmu_part, o);
        if (lane == 0) red[warp] = gmu_part;
        __syncthreads();
        if (warp == 0) {
            float v = lane < nwarp ? red[lane] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) a.gmu[b] = v;
        }
    }
}

}  // namespace ppnn
#include <cstdlib>
#include "coarse_adjoint_staged.cuh"
#include "coarse_adjoint_periodic.cuh"

using namespace ppnn;

// the reference's own meshes: compile-time period of the prolongation (coarse_adjoint_periodic.cuh)
template <int PER, int ADV>
static int launch_adjoint_periodic(const CoarseParams& p, const AdjParams& a, int B, void* stream) {
    const PhaseWeights pw = make_phase_weights(p.suh, PER);
    void (*kern)(CoarseParams, AdjParams, PhaseWeights, const float*);
    kern = a.gnn ? coarse_adjoint_periodic_kernel<PER, ADV, true, true>
         : a.add_identity ? coarse_adjoint_periodic_kernel<PER, ADV, true, false>
                          : coarse_adjoint_periodic_kernel<PER, ADV, false, false>;
    constexpr size_t smem = adjoint_periodic_smem<PER, ADV>();
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    if (a.gmu) {                                                 // the two CTAs of a sample add their halves
        e = cudaMemsetAsync(a.gmu, 0, (size_t)B * sizeof(float), (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    const float* g_end = a.g + (size_t)B * 2 * kPerF * kPerF;
    kern<<<2 * B, kAdjThreads, smem, (cudaStream_t)stream>>>(p, a, pw, g_end);
    return launch_status();
}

static int run_adjoint(const ppnn_coarse_cfg* cfg, const float* u, const float* coarse, const float* mu,
                       const float* grad_out, float* grad_u, float* grad_mu, float* grad_nn, const float* outstd,
                       int B, int add_identity, void* stream) {
    if ((!u && !coarse) || !mu || !grad_out || !grad_u) return PPNN_E_NULL;
    CoarseParams p;
    int st = fill_coarse_params(cfg, &p);
    if (st != PPNN_OK) return st;
    if (B < 0) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    AdjParams a;
    a.u = u; a.coarse = coarse; a.mu = mu; a.g = grad_out; a.gu = grad_u; a.gmu = grad_mu; a.add_identity = add_identity;
    a.gnn = grad_nn; a.os[0] = outstd ? outstd[0] : 1.f; a.os[1] = outstd ? outstd[1] : 1.f;
    a.apitch = p.FW | 1;
    // fine columns whose 4-tap window can reach one coarse column: 4 / suw (+ slack)
    int support = (int)(4.0f / p.suw) + 4;
    if (support > kMaxSupport) return PPNN_E_TOOBIG;
    if (support > p.FW) support = p.FW;
    a.support = support;
    // the fused network-output gradient needs the last fine column on the last coarse node (see the kernel)
    const float slast = p.suw * (float)(p.FW - 1);
    const bool gnn_ok = (p.FW - 1) % 32 == 0 && floorf(slast) == slast && (int)slast >= p.CW - 1;
    if (grad_nn && !add_identity) return PPNN_E_SHAPE;      // the network-output gradient only exists for the update
    if (p.FH == kPerF && p.FW == kPerF && p.CH == p.CW && grad_u != grad_out && !PPNN_ENV_OFF("PPNN_NO_PERIODIC") &&
        (!coarse || (reinterpret_cast<uintptr_t>(coarse) & 15) == 0)) {
        if (p.CH == 49) return launch_adjoint_periodic<16, 3>(p, a, B, stream);
        if (p.CH == 33) return launch_adjoint_periodic<8, 1>(p, a, B, stream);
    }
    // staged path: (sample, field) CTAs in 2-CTA clusters, TMA ring (coarse_adjoint_staged.cuh)
    {
        AdjStagedLayout lay;
        lay.wmain = ((p.FW - 1) % 32 == 0) ? p.FW - 1 : p.FW;
        lay.nchunk = (lay.wmain + 31) / 32;
        lay.apitch = a.apitch; lay.support = support;
        lay.gstage = (kAdjStageRows * p.FW + 8 + 3) & ~3;
        lay.split = p.FH;
        if (2 * lay.nchunk <= kAdjWarps && p.FH >= 4 * kAdjStageRows)
            lay.split = ((p.FH / 2 + kAdjStageRows - 1) / kAdjStageRows) * kAdjStageRows;
        int n0 = (lay.split + kAdjStageRows - 1) / kAdjStageRows;
        int n1 = (p.FH - lay.split + kAdjStageRows - 1) / kAdjStageRows;
        lay.nsteps = n0 > n1 ? n0 : n1;
        const int nn2 = p.NH * p.NW;
        size_t un = (size_t)kAdjStages * 2 * lay.gstage;
        if (un < (size_t)2 * nn2) un = (size_t)2 * nn2;
        size_t smem2 = (size_t)p.FH * sizeof(float4) + (size_t)((p.FH + 2 * lay.nsteps + p.CW + 3) & ~3) * sizeof(int) +
                       ((size_t)((p.CW * support + 3) & ~3) + ((nn2 + 3) & ~3) + (((size_t)(p.CH + 4) * lay.apitch + 3) & ~(size_t)3) + un) * sizeof(float);
        if (lay.nchunk <= kAdjWarps && (p.FW - 1) % 4 == 0 && p.CH <= p.FH && smem2 <= 112 * 1024 &&
            grad_u != grad_out && !PPNN_ENV_OFF("PPNN_NO_TMA") && (!grad_nn || gnn_ok)) {
            void (*kern)(CoarseParams, AdjParams, AdjStagedLayout, const float*);
            if (p.FW == 257) kern = grad_nn ? coarse_adjoint_staged_kernel<257, true, true>
                                  : add_identity ? coarse_adjoint_staged_kernel<257, true, false> : coarse_adjoint_staged_kernel<257, false, false>;
            else             kern = grad_nn ? coarse_adjoint_staged_kernel<0, true, true>
                                  : add_identity ? coarse_adjoint_staged_kernel<0, true, false> : coarse_adjoint_staged_kernel<0, false, false>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e != cudaSuccess) return (int)e;
            const float* g_end = grad_out + (size_t)B * 2 * p.FH * p.FW;
            if (a.gmu) {                                         // the two CTAs of a sample add their halves
                e = cudaMemsetAsync(a.gmu, 0, (size_t)B * sizeof(float), (cudaStream_t)stream);
                if (e != cudaSuccess) return (int)e;
            }
            kern<<<2 * B, kAdjThreads, smem2, (cudaStream_t)stream>>>(p, a, lay, g_end);
            return launch_status();
        }
    }
    if (grad_nn) return PPNN_E_SHAPE;          // only the staged kernel fuses the network-output gradient
    size_t smem = (size_t)4 * p.NH * p.NW * sizeof(float) + (size_t)p.FH * (sizeof(float4) + sizeof(int)) +
                  (size_t)p.CW * sizeof(int) + (size_t)p.CW * support * sizeof(float) +
                  (size_t)2 * p.CH * a.apitch * sizeof(float);
    if (smem > 220 * 1024) return PPNN_E_TOOBIG;
    cudaError_t e = cudaFuncSetAttribute(coarse_adjoint_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int nchunk = (p.FW + 31) / 32;
    int warps = 2 * nchunk;
    if (warps > 18) warps = 18;
    if (warps < 4) warps = 4;
    coarse_adjoint_kernel<<<B, warps * 32, smem, (cudaStream_t)stream>>>(p, a);
    return launch_status();
}

extern "C" int ppnn_pde_du_bwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                               const float* grad_out, float* grad_u, float* grad_mu,
                               int B, int add_identity, void* stream) {
    if (!u) return PPNN_E_NULL;
    return run_adjoint(cfg, u, nullptr, mu, grad_out, grad_u, grad_mu, nullptr, nullptr, B, add_identity, stream);
}

extern "C" int ppnn_step_update_bwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                                    const float* grad_out, const float* outstd,
                                    float* grad_u, float* grad_mu, float* grad_nn, int B, void* stream) {
    if (!u) return PPNN_E_NULL;
    if (grad_nn && !outstd) return PPNN_E_NULL;
    return run_adjoint(cfg, u, nullptr, mu, grad_out, grad_u, grad_mu, grad_nn, outstd, B, 1, stream);
}

#ifdef PPNN_TRACE
extern "C" int ppnn_trace_read_adj(unsigned long long* host, long n) {
    return (int)cudaMemcpyFromSymbol(host, ppnn::g_trace_adj, (size_t)n * sizeof(unsigned long long));
}
#endif

extern "C" int ppnn_pde_du_bwd_saved(const ppnn_coarse_cfg* cfg, const float* coarse_state, const float* mu,
                                     const float* grad_out, float* grad_u, float* grad_mu,
                                     int B, int add_identity, void* stream) {
    if (!coarse_state) return PPNN_E_NULL;
    return run_adjoint(cfg, nullptr, coarse_state, mu, grad_out, grad_u, grad_mu, nullptr, nullptr, B, add_identity, stream);
}

extern "C" int ppnn_step_update_bwd_saved(const ppnn_coarse_cfg* cfg, const float* coarse_state, const float* mu,
                                          const float* grad_out, const float* outstd,
                                          float* grad_u, float* grad_mu, float* grad_nn, int B, void* stream) {
    if (!coarse_state) return PPNN_E_NULL;
    if (grad_nn && !outstd) return PPNN_E_NULL;
    return run_adjoint(cfg, nullptr, coarse_state, mu, grad_out, grad_u, grad_mu, grad_nn, outstd, B, 1, stream);
}

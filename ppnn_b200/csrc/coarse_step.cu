// Fused coarse-grid PPNN step, forward:  pde_du (src/train2D.py:66-79) and the
// two kernels of one rollout step around the trainable network
// (src/train2D.py:230-245).  One CTA owns one sample:
//
//   phase 1  bilinear restriction gather of both fields into shared memory
//            (mesh_convertor.down + [:, :, :-1, :-1])
//   phase 2  dt * rhs on the periodic N x N coarse grid, still in shared memory
//            (padbcx/padbcy wrap, Upwind1 / Central2 stencils, rhs.py algebra)
//   phase 3  stream the fine grid once: every warp takes (field, 32-column
//            chunk, row band) strips, keeps a 4-row window of y-interpolated
//            coarse values in registers (padBC_rd + bicubic, W pass then H pass
//            like torch) and fuses the epilogue (pdeu | normalised net inputs |
//            explicit update with the network output) into the store.
//
// HBM-bound by construction: per fine point the update kernel reads u and
// nn_out once and writes u_next once (12 B); the restriction gather re-reads
// (F-1)/(C-1)-strided rows only.  No tensor cores: 1-channel 3..9-tap stencils
// and a 4x4 separable interpolation have no dense contraction to feed them.
#include "coarse_common.cuh"

namespace ppnn {

enum { MODE_PDEU = 0, MODE_INPUTS = 1, MODE_UPDATE = 2 };

struct StepIO {
    const float* u;        // [B,2,FH,FW]
    const float* mu;       // [B]
    const float* nn;       // [B,2,FH-1,FW-1] or null          (MODE_UPDATE)
    float* out;            // pdeu | u_next [B,2,FH,FW]        (may be null in MODE_INPUTS)
    float* u_norm;         // [B,2,FH-1,FW-1]                  (MODE_INPUTS)
    float* p_norm;         // [B,2,FH-1,FW-1] or null          (MODE_INPUTS)
    float m0[2], s0[2];    // inmean/instd   (INPUTS) | outmean/outstd (UPDATE)
    float m1[2], s1[2];    // pdemean/pdestd (INPUTS)
};

constexpr int kRowUnroll = 8;

template <int MODE>
__global__ void __launch_bounds__(512, 2)
coarse_step_kernel(const CoarseParams p, const StepIO io, int band_rows) {
    extern __shared__ float smem[];
    const int nn2 = p.NH * p.NW;
    float* cu = smem;
    float* cv = cu + nn2;
    float* ru = cv + nn2;
    float* rv = ru + nn2;
    float4* roww = reinterpret_cast<float4*>(rv + nn2);     // [FH] x-weights (16 B aligned: 4*nn2 floats)
    int* rowi = reinterpret_cast<int*>(roww + p.FH);        // [FH] floor(i * suh)

    const int b = blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const size_t plane = (size_t)p.FH * p.FW;
    const float* ub = io.u + (size_t)b * 2 * plane;
    const float mu = __ldg(io.mu + b);

    // ---- phase 1: restriction gather ------------------------------------
    for (int idx = tid; idx < 2 * nn2; idx += nthr) {
        int f = idx >= nn2;
        int r = idx - f * nn2;
        int I = r / p.NW, J = r - I * p.NW;
        cu[idx] = restrict_point(p, ub + f * plane, I, J);     // cv follows cu contiguously
    }
    for (int i = tid; i < p.FH; i += nthr) {
        float w[4]; int i0;
        cubic_entry(p.suh, i, &i0, w);
        rowi[i] = i0;
        roww[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();

    // ---- phase 2: dt * rhs on the periodic coarse grid --------------------
    for (int idx = tid; idx < nn2; idx += nthr) {
        int I = idx / p.NW, J = idx - I * p.NW;
        ru[idx] = p.dt * eval_rhs(p, p.rhs_u, cu, cu, cv, I, J, mu);
        rv[idx] = p.dt * eval_rhs(p, p.rhs_v, cv, cu, cv, I, J, mu);
    }
    __syncthreads();

    // ---- phase 3: prolongation + fused epilogue, one strip per warp -------
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int nchunk = (p.FW + 31) >> 5;
    const int nband = (p.FH + band_rows - 1) / band_rows;
    const int nstrip = 2 * nband * nchunk;
    const int PH = p.FH - 1, PW = p.FW - 1;                    // network grid (no endpoint)
    const size_t nplane = (size_t)PH * PW;

    for (int s = warp; s < nstrip; s += nwarp) {
        // consecutive warps take consecutive column chunks of the same band
        const int chunk = s % nchunk;
        const int t2 = s / nchunk;
        const int f = t2 & 1;
        const int band = t2 >> 1;
        const int j = chunk * 32 + lane;
        const bool active = j < p.FW;
        const int jc = active ? j : p.FW - 1;
        const float* rf = f ? rv : ru;

        float wy[4]; int j0;
        cubic_entry(p.suw, jc, &j0, wy);
        const float* rc0 = rf + ext_index(j0 - 1, p.CW);
        const float* rc1 = rf + ext_index(j0, p.CW);
        const float* rc2 = rf + ext_index(j0 + 1, p.CW);
        const float* rc3 = rf + ext_index(j0 + 2, p.CW);
        auto yrow = [&](int I) -> float {     // W-pass of row I of the extended coarse increment
            const int o = ext_index(I, p.CH) * p.NW;
            return rc0[o] * wy[0] + rc1[o] * wy[1] + rc2[o] * wy[2] + rc3[o] * wy[3];
        };

        const int r0 = band * band_rows;
        const int r1 = min(r0 + band_rows, p.FH);
        int base = rowi[r0];
        float h0 = yrow(base - 1), h1 = yrow(base), h2 = yrow(base + 1), h3 = yrow(base + 2);
        // H-pass of fine row ii (advances the window when floor(ii*suh) moves on; warp-uniform)
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
        const size_t off = ((size_t)b * 2 + f) * plane + (size_t)r0 * p.FW + jc;
        const float* up = io.u + off;                          // row pointers, bumped by one pitch per row
        float* op = io.out ? io.out + off : nullptr;

        if (MODE == MODE_PDEU) {
            for (int ii = r0; ii < r1; ++ii, op += p.FW) {
                float pde = prolong(ii);
                if (active) st_stream(op, pde);
            }
        } else if (MODE == MODE_UPDATE) {
            const bool has_nn = io.nn != nullptr;
            const int jn = jc == PW ? 0 : jc;                  // padBC_rd column fold
            const float* nbase = has_nn ? io.nn + ((size_t)b * 2 + f) * nplane + jn : nullptr;
            const float* np = has_nn ? nbase + (size_t)r0 * PW : nullptr;
            const int rmain = min(r1, PH);                     // rows whose network row is the row itself
            int ii = r0;
            for (; ii + kRowUnroll <= rmain; ii += kRowUnroll) {
                float uu[kRowUnroll], nv[kRowUnroll];
                if (active) {
#pragma unroll
                    for (int q = 0; q < kRowUnroll; ++q) uu[q] = ld_stream(up + (size_t)q * p.FW);
                    if (has_nn) {
#pragma unroll
                        for (int q = 0; q < kRowUnroll; ++q) nv[q] = ld_stream(np + (size_t)q * PW);
                    }
                }
#pragma unroll
                for (int q = 0; q < kRowUnroll; ++q) {
                    float pde = prolong(ii + q);
                    if (active) {
                        float v = has_nn ? (nv[q] * s0 + m0) + uu[q] + pde : uu[q] + pde;
                        st_stream(op + (size_t)q * p.FW, v);
                    }
                }
                up += (size_t)kRowUnroll * p.FW; op += (size_t)kRowUnroll * p.FW;
                if (has_nn) np += (size_t)kRowUnroll * PW;
            }
            for (; ii < r1; ++ii, up += p.FW, op += p.FW) {    // ragged tail and the endpoint row
                float pde = prolong(ii);
                if (active) {
                    float uu = ld_stream(up);
                    float v = uu + pde;
                    if (has_nn) v = (ld_stream(nbase + (size_t)(ii == PH ? 0 : ii) * PW) * s0 + m0) + uu + pde;
                    st_stream(op, v);
                }
            }
        } else {   // MODE_INPUTS
            const float rs0 = 1.f / s0, rs1 = 1.f / s1;
            const bool colok = active && jc < PW;
            float* unp = io.u_norm + ((size_t)b * 2 + f) * nplane + (size_t)r0 * PW + jc;
            float* pnp = io.p_norm ? io.p_norm + ((size_t)b * 2 + f) * nplane + (size_t)r0 * PW + jc : nullptr;
            const int rmain = min(r1, PH);
            int ii = r0;
            for (; ii + kRowUnroll <= rmain; ii += kRowUnroll) {
                float uu[kRowUnroll];
                if (colok) {
#pragma unroll
                    for (int q = 0; q < kRowUnroll; ++q) uu[q] = ld_stream(up + (size_t)q * p.FW);
                }
#pragma unroll
                for (int q = 0; q < kRowUnroll; ++q) {
                    float pde = prolong(ii + q);
                    if (op && active) st_stream(op + (size_t)q * p.FW, pde);
                    if (colok) {
                        st_stream(unp + (size_t)q * PW, div_by(uu[q] - m0, s0, rs0));
                        if (pnp) st_stream(pnp + (size_t)q * PW, div_by(pde - m1, s1, rs1));
                    }
                }
                up += (size_t)kRowUnroll * p.FW; unp += (size_t)kRowUnroll * PW;
                if (op) op += (size_t)kRowUnroll * p.FW;
                if (pnp) pnp += (size_t)kRowUnroll * PW;
            }
            for (; ii < r1; ++ii) {
                float pde = prolong(ii);
                if (op && active) st_stream(op, pde);
                if (colok && ii < PH) {
                    st_stream(unp, div_by(ld_stream(up) - m0, s0, rs0));
                    if (pnp) st_stream(pnp, div_by(pde - m1, s1, rs1));
                }
                up += p.FW; unp += PW;
                if (op) op += p.FW;
                if (pnp) pnp += PW;
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
    return (size_t)4 * p.NH * p.NW * sizeof(float) + (size_t)p.FH * (sizeof(float4) + sizeof(int));
}

template <int MODE>
static int launch_step(const ppnn_coarse_cfg* cfg, const StepIO& io, int B, void* stream) {
    CoarseParams p;
    int st = fill_coarse_params(cfg, &p);
    if (st != PPNN_OK) return st;
    if (B < 0) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    size_t smem = step_smem_bytes(p);
    if (smem > 200 * 1024) return PPNN_E_TOOBIG;
    auto kern = coarse_step_kernel<MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    // strips: 2 fields x ceil(FW/32) chunks x bands; aim at >= 2 strips per warp
    int nchunk = (p.FW + 31) / 32;
    int nband = (p.FH + 63) / 64;
    int band_rows = (p.FH + nband - 1) / nband;
    band_rows = (band_rows + kRowUnroll - 1) / kRowUnroll * kRowUnroll;
    int nstrip = 2 * nchunk * ((p.FH + band_rows - 1) / band_rows);
    int warps = nstrip < 16 ? nstrip : 16;
    if (warps < 4) warps = 4;
    kern<<<B, warps * 32, smem, (cudaStream_t)stream>>>(p, io, band_rows);
    return launch_status();
}

}  // namespace ppnn

using namespace ppnn;

extern "C" int ppnn_pde_du_fwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                               float* pdeu, int B, void* stream) {
    if (!u || !mu || !pdeu) return PPNN_E_NULL;
    StepIO io = {};
    io.u = u; io.mu = mu; io.out = pdeu;
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

extern "C" int ppnn_step_update(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                                const float* nn_out, const float* outmean, const float* outstd,
                                float* u_next, int B, void* stream) {
    if (!u || !mu || !u_next) return PPNN_E_NULL;
    if (nn_out && (!outmean || !outstd)) return PPNN_E_NULL;
    StepIO io = {};
    io.u = u; io.mu = mu; io.nn = nn_out; io.out = u_next;
    for (int c = 0; c < 2; ++c) { io.m0[c] = nn_out ? outmean[c] : 0.f; io.s0[c] = nn_out ? outstd[c] : 1.f; }
    return launch_step<MODE_UPDATE>(cfg, io, B, stream);
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

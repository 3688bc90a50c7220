// Tier A: mesh_convertor.down / .up (src/utility/utils.py:13-31) as stand-alone
// kernels: torch F.interpolate(mode='bilinear' | 'bicubic', size=..., align_corners)
// semantics -- source index = scale * o (align_corners) or scale*(o+.5)-.5,
// bilinear weights (1-t, t) with the +1 neighbour dropped at the border, Keys
// cubic A = -0.75 with taps clamped to the border.  Forward is
// a gather (thread per output, coalesced stores); the adjoint is a gather too
// (thread per grad_in element, fixed summation order): unlike torch's own CUDA
// backward, which scatters with atomics, it is bit-identical from run to run.
#include "common.cuh"

namespace ppnn {

struct ResampleParams {
    const float* in; float* out;
    int Hin, Win, Hout, Wout;
    long in_stride;  // floats between consecutive input planes (Hin*Win when dense; forward only)
    int align;
    float sh, sw;
    long total;     // planes * Hout * Wout
};

template <int MODE>
__global__ void __launch_bounds__(256)
resample_kernel(const ResampleParams p) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total;
         idx += (long)gridDim.x * blockDim.x) {
        int j = (int)(idx % p.Wout);
        long t = idx / p.Wout;
        int i = (int)(t % p.Hout);
        long pl = t / p.Hout;
        const float* q = p.in + (size_t)pl * p.in_stride;
        if (MODE == PPNN_BILINEAR) {
            float si = resample_src(p.sh, i, p.align, false), sj = resample_src(p.sw, j, p.align, false);
            int i0 = min((int)si, p.Hin - 1), j0 = min((int)sj, p.Win - 1);
            int di = i0 < p.Hin - 1 ? 1 : 0, dj = j0 < p.Win - 1 ? 1 : 0;
            float li = si - (float)i0, lj = sj - (float)j0;
            float h0 = 1.f - li, w0 = 1.f - lj;
            size_t o00 = (size_t)i0 * p.Win + j0;
            size_t o01 = o00 + dj, o10 = o00 + (size_t)di * p.Win, o11 = o10 + dj;
            float v = h0 * (w0 * __ldg(q + o00) + lj * __ldg(q + o01)) +
                      li * (w0 * __ldg(q + o10) + lj * __ldg(q + o11));
            st_stream(p.out + idx, v);
        } else {
            float si = resample_src(p.sh, i, p.align, true), sj = resample_src(p.sw, j, p.align, true);
            float fi = floorf(si), fj = floorf(sj);
            int i0 = (int)fi, j0 = (int)fj;
            float wi[4], wj[4];
            cubic_weights(si - fi, wi);
            cubic_weights(sj - fj, wj);
            int cj[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) cj[k] = clampi(j0 - 1 + k, 0, p.Win - 1);
            float v = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float* r = q + (size_t)clampi(i0 - 1 + a, 0, p.Hin - 1) * p.Win;
                // W pass then H pass, like torch's cubic_interp1d nesting
                v += (__ldg(r + cj[0]) * wj[0] + __ldg(r + cj[1]) * wj[1] +
                      __ldg(r + cj[2]) * wj[2] + __ldg(r + cj[3]) * wj[3]) * wi[a];
            }
            st_stream(p.out + idx, v);
        }
    }
}

// ---- adjoint as a GATHER: thread = one element of grad_in, sums the contributions of every output whose
// window reaches it in a fixed (row-major) order => bit-identical from run to run (no atomics, no zero fill).
// weight with which output index o feeds input index I along one axis (0 if it does not)
template <int MODE>
__device__ __forceinline__ float adjoint_weight(float scale, bool align, int o, int I, int nin) {
    if (MODE == PPNN_BILINEAR) {
        const float s = resample_src(scale, o, align, false);
        const int i0 = min((int)s, nin - 1);
        const int d = i0 < nin - 1 ? 1 : 0;
        const float l = s - (float)i0;
        float w = 0.f;
        if (i0 == I) w += 1.f - l;
        if (i0 + d == I) w += l;
        return w;
    }
    const float s = resample_src(scale, o, align, true);
    const float fl = floorf(s);
    const int i0 = (int)fl;
    float wv[4];
    cubic_weights(s - fl, wv);
    float w = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
        if (clampi(i0 - 1 + a, 0, nin - 1) == I) w += wv[a];
    return w;
}

// conservative range of output indices whose window can reach input index I (exact test: adjoint_weight)
template <int MODE>
__device__ __forceinline__ void adjoint_range(float scale, bool align, int I, int nin, int nout, int* lo, int* hi) {
    const float reach = MODE == PPNN_BILINEAR ? 1.f : 2.f;
    const float off = align ? 0.f : 0.5f;
    int a = 0, b = nout - 1;
    if (scale > 0.f) {
        a = (int)floorf(((float)I - reach + off) / scale - off) - 1;
        b = (int)ceilf(((float)I + reach + off) / scale - off) + 1;
    }
    if (I == 0) a = 0;                     // clamped taps below the first input pile up on it
    if (I == nin - 1) b = nout - 1;        // ... and beyond the last one
    *lo = max(a, 0);
    *hi = min(b, nout - 1);
}

template <int MODE>
__global__ void __launch_bounds__(256)
resample_adjoint_kernel(const ResampleParams p, long total_in) {
    // `in` is grad_out [planes, Hout, Wout], `out` is grad_in [planes, Hin, Win]
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_in;
         idx += (long)gridDim.x * blockDim.x) {
        const int J = (int)(idx % p.Win);
        const long t = idx / p.Win;
        const int I = (int)(t % p.Hin);
        const long pl = t / p.Hin;
        const float* g = p.in + (size_t)pl * p.Hout * p.Wout;
        int ilo, ihi, jlo, jhi;
        adjoint_range<MODE>(p.sh, p.align, I, p.Hin, p.Hout, &ilo, &ihi);
        adjoint_range<MODE>(p.sw, p.align, J, p.Win, p.Wout, &jlo, &jhi);
        float acc = 0.f;
        for (int i = ilo; i <= ihi; ++i) {
            const float wi = adjoint_weight<MODE>(p.sh, p.align, i, I, p.Hin);
            if (wi == 0.f) continue;
            float row = 0.f;
            for (int j = jlo; j <= jhi; ++j) {
                const float wj = adjoint_weight<MODE>(p.sw, p.align, j, J, p.Win);
                if (wj != 0.f) row += wj * __ldg(g + (size_t)i * p.Wout + j);
            }
            acc += wi * row;
        }
        p.out[idx] = acc;
    }
}

// ---- bicubic forward, the hot direction (mesh_convertor.up: 80 % of the reference's CPU step) ----
// thread = output column, walking a band of output rows: the W pass of an input row is done
// once per thread and kept in a 4-row register window (W pass then H pass, like torch);
// per-row H weights come from a shared table built once per CTA.
constexpr int kUpBand = 64;      // output rows per CTA

// Write-only kernels on rows of 257 floats run at the store path's rate, not the HBM's: rows start on every 4-byte
// phase, so a warp's 128-byte store touches five 32-byte sectors, two of them partially.  Ordinary (write-back)
// stores: in the whole-plane kernel the streaming hint costs 20 % (0.207 vs 0.169 ms for 2160 planes 33^2 -> 257^2).
// A bare store microbenchmark (tools/microbench/store_pattern.cu) puts this layout at 2.8 TB/s and aligned lines of the
// linear plane at 6.2 - 6.8 TB/s.  Tried and dropped (same box, profiles/r2_mesh_up_ab.txt): rows staged in shared
// memory and written as 16-byte-aligned vectors -- 0.31 ms with one barrier per 4 rows, 0.21 ms with 8 rows per band
// and full occupancy (ncu: issue-bound, 39 instructions per row and warp), 0.19 ms with two columns per thread;
// the plane as a linear array of aligned float4 with the H pass from a W-interpolated copy of the source in shared
// memory (0.27 ms: ~18 instructions per output).  The thread-per-column kernel below stays at 0.17 ms.
__device__ __forceinline__ void up_store(float* p, float v) { *p = v; }

__global__ void __launch_bounds__(128)
bicubic_up_kernel(const ResampleParams p) {
    __shared__ float4 roww[kUpBand];
    __shared__ int rowi[kUpBand];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * kUpBand;
    const int r1 = min(r0 + kUpBand, p.Hout);
    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
        const float si = resample_src(p.sh, i, p.align, true);
        const float fi = floorf(si);
        float w[4];
        cubic_weights(si - fi, w);
        rowi[i - r0] = (int)fi;
        roww[i - r0] = make_float4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    if (j >= p.Wout) return;
    const float sj = resample_src(p.sw, j, p.align, true);
    const float fj = floorf(sj);
    const int j0 = (int)fj;
    float wj[4];
    cubic_weights(sj - fj, wj);
    const int c0 = clampi(j0 - 1, 0, p.Win - 1), c1 = clampi(j0, 0, p.Win - 1),
              c2 = clampi(j0 + 1, 0, p.Win - 1), c3 = clampi(j0 + 2, 0, p.Win - 1);
    const float* in = p.in + (size_t)blockIdx.z * p.in_stride;
    float* out = p.out + ((size_t)blockIdx.z * p.Hout + r0) * p.Wout + j;
    auto wpass = [&](int I) -> float {
        const float* q = in + (size_t)clampi(I, 0, p.Hin - 1) * p.Win;
        return __ldg(q + c0) * wj[0] + __ldg(q + c1) * wj[1] + __ldg(q + c2) * wj[2] + __ldg(q + c3) * wj[3];
    };
    int base = rowi[0];
    float h0 = wpass(base - 1), h1 = wpass(base), h2 = wpass(base + 1), h3 = wpass(base + 2);
    for (int i = r0; i < r1; ++i, out += p.Wout) {
        const int i0 = rowi[i - r0];
        const float4 wx = roww[i - r0];
        if (i0 != base) {
            if (i0 == base + 1) { h0 = h1; h1 = h2; h2 = h3; h3 = wpass(i0 + 2); }
            else { h0 = wpass(i0 - 1); h1 = wpass(i0); h2 = wpass(i0 + 1); h3 = wpass(i0 + 2); }
            base = i0;
        }
        up_store(out, h0 * wx.x + h1 * wx.y + h2 * wx.z + h3 * wx.w);
    }
}

// Whole-plane variant for small sources (every coarse mesh of the reference: 33^2, 49^2): CTA = one plane, the
// source plane and the per-row tables live in shared memory, thread = output column walking ALL rows (the W pass
// reads shared memory instead of gathering from global).  Same expressions as bicubic_up_kernel: bit-identical.
__global__ void __launch_bounds__(1024)
bicubic_up_plane_kernel(const ResampleParams p) {
    extern __shared__ float4 up_smem[];
    float4* roww = up_smem;                                              // [Hout]
    int* rowi = reinterpret_cast<int*>(roww + p.Hout);                   // [Hout]
    float* sin = reinterpret_cast<float*>(rowi + ((p.Hout + 3) & ~3));   // [Hin * Win]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const float* in = p.in + (size_t)blockIdx.x * p.in_stride;
    for (int idx = tid; idx < p.Hin * p.Win; idx += nthr) sin[idx] = __ldg(in + idx);
    for (int i = tid; i < p.Hout; i += nthr) {
        const float si = resample_src(p.sh, i, p.align, true);
        const float fi = floorf(si);
        float w[4];
        cubic_weights(si - fi, w);
        rowi[i] = (int)fi;
        roww[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    for (int j = tid; j < p.Wout; j += nthr) {
        const float sj = resample_src(p.sw, j, p.align, true);
        const float fj = floorf(sj);
        const int j0 = (int)fj;
        float wj[4];
        cubic_weights(sj - fj, wj);
        const float* q0 = sin + clampi(j0 - 1, 0, p.Win - 1);
        const float* q1 = sin + clampi(j0, 0, p.Win - 1);
        const float* q2 = sin + clampi(j0 + 1, 0, p.Win - 1);
        const float* q3 = sin + clampi(j0 + 2, 0, p.Win - 1);
        auto wpass = [&](int I) -> float {
            const int o = clampi(I, 0, p.Hin - 1) * p.Win;
            return q0[o] * wj[0] + q1[o] * wj[1] + q2[o] * wj[2] + q3[o] * wj[3];
        };
        float* out = p.out + (size_t)blockIdx.x * p.Hout * p.Wout + j;
        int base = rowi[0];
        float h0 = wpass(base - 1), h1 = wpass(base), h2 = wpass(base + 1), h3 = wpass(base + 2);
        for (int i = 0; i < p.Hout; ++i, out += p.Wout) {
            const int i0 = rowi[i];
            const float4 wx = roww[i];
            if (i0 != base) {
                if (i0 == base + 1) { h0 = h1; h1 = h2; h2 = h3; h3 = wpass(i0 + 2); }
                else { h0 = wpass(i0 - 1); h1 = wpass(i0); h2 = wpass(i0 + 1); h3 = wpass(i0 + 2); }
                base = i0;
            }
            up_store(out, h0 * wx.x + h1 * wx.y + h2 * wx.z + h3 * wx.w);
        }
    }
}

static int run_resample(const float* in, float* out, long planes, long in_stride, int Hin, int Win, int Hout, int Wout,
                        int mode, int align, bool bwd, void* stream) {
    if (!in || !out) return PPNN_E_NULL;
    if (planes < 0 || Hin < 1 || Win < 1 || Hout < 1 || Wout < 1 || in_stride < (long)Hin * Win) return PPNN_E_SHAPE;
    if (mode != PPNN_BILINEAR && mode != PPNN_BICUBIC) return PPNN_E_MODE;
    if (planes == 0) return PPNN_OK;
    ResampleParams p;
    p.in = in; p.out = out; p.Hin = Hin; p.Win = Win; p.Hout = Hout; p.Wout = Wout; p.in_stride = in_stride;
    p.align = align != 0;
    p.sh = resample_scale(Hin, Hout, p.align);
    p.sw = resample_scale(Win, Wout, p.align);
    p.total = planes * Hout * Wout;
    cudaStream_t s = (cudaStream_t)stream;
    if (bwd) {
        const long total_in = planes * Hin * Win;
        long nb = (total_in + 255) / 256;
        if (nb > 148L * 32) nb = 148L * 32;
        if (mode == PPNN_BILINEAR) resample_adjoint_kernel<PPNN_BILINEAR><<<(unsigned)nb, 256, 0, s>>>(p, total_in);
        else                       resample_adjoint_kernel<PPNN_BICUBIC><<<(unsigned)nb, 256, 0, s>>>(p, total_in);
        return launch_status();
    }
    long blocks = (p.total + 255) / 256;
    if (blocks > 148L * 32) blocks = 148L * 32;
    unsigned g = (unsigned)blocks;
    const size_t up_smem = (size_t)Hout * sizeof(float4) + (size_t)((Hout + 3) & ~3) * sizeof(int) + (size_t)Hin * Win * sizeof(float);
    if (mode == PPNN_BICUBIC && !bwd && up_smem <= 96 * 1024 && planes < (1L << 31)) {
        cudaError_t e = cudaFuncSetAttribute(bicubic_up_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)up_smem);
        if (e != cudaSuccess) return (int)e;
        const int threads = Wout >= 1024 ? 1024 : (Wout + 31) & ~31;
        bicubic_up_plane_kernel<<<(unsigned)planes, threads, up_smem, s>>>(p);
        return launch_status();
    }
    if (mode == PPNN_BICUBIC && !bwd && planes <= 65535 && (Hout + kUpBand - 1) / kUpBand <= 65535) {
        dim3 grid((Wout + 127) / 128, (Hout + kUpBand - 1) / kUpBand, (unsigned)planes);
        bicubic_up_kernel<<<grid, 128, 0, s>>>(p);
        return launch_status();
    }
    if (mode == PPNN_BILINEAR) resample_kernel<PPNN_BILINEAR><<<g, 256, 0, s>>>(p);
    else                       resample_kernel<PPNN_BICUBIC><<<g, 256, 0, s>>>(p);
    return launch_status();
}

}  // namespace ppnn

using namespace ppnn;

extern "C" int ppnn_resample_fwd(const float* in, float* out, long planes, int Hin, int Win,
                                 int Hout, int Wout, int mode, int align_corners, void* stream) {
    return run_resample(in, out, planes, (long)Hin * Win, Hin, Win, Hout, Wout, mode, align_corners, false, stream);
}

extern "C" int ppnn_resample_fwd_strided(const float* in, float* out, long planes, long in_plane_stride, int Hin, int Win,
                                         int Hout, int Wout, int mode, int align_corners, void* stream) {
    return run_resample(in, out, planes, in_plane_stride, Hin, Win, Hout, Wout, mode, align_corners, false, stream);
}

extern "C" int ppnn_resample_bwd(const float* grad_out, float* grad_in, long planes, int Hin, int Win,
                                 int Hout, int Wout, int mode, int align_corners, void* stream) {
    return run_resample(grad_out, grad_in, planes, (long)Hin * Win, Hin, Win, Hout, Wout, mode, align_corners, true, stream);
}

extern "C" int ppnn_abi_version(void) { return PPNN_B200_ABI_VERSION; }

extern "C" const char* ppnn_status_string(int status) {
    switch (status) {
    case PPNN_OK: return "ok";
    case PPNN_E_NULL: return "required pointer is NULL";
    case PPNN_E_SHAPE: return "non-positive or inconsistent sizes";
    case PPNN_E_SCHEME: return "unknown scheme / accuracy combination";
    case PPNN_E_RHS: return "unknown right-hand-side id";
    case PPNN_E_TOOBIG: return "coarse grid does not fit in shared memory";
    case PPNN_E_MODE: return "unknown mode / flag value";
    case PPNN_E_HALO: return "halo does not match the stencils";
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown status";
}

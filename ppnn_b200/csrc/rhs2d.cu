// Tier A: one rhs.py function (src/rhs.py:1-47,52-56) evaluated in a single
// pass on the caller's pre-padded halos ax [B,H+2h,W] / ay [B,H,W+2h] -- what
// train2D.py:75-77 and operators.py:474-475 compute with ~30 ATen launches.
// 20 B per output point (u1, v1, ax, ay read, out written).
#include "common.cuh"
#include "tma.cuh"

namespace ppnn {

struct RhsParams {
    const float* u1; const float* v1; const float* mu; const float* ax; const float* ay;
    float* out;
    int H, W, halo, K;         // K = 2*halo + 1 taps for every operator
    int rhs_id, adv_upwind;
    float aa[kMaxTaps], ab[kMaxTaps], d[kMaxTaps];
    float dx, dy, dx2, dy2;
    float rdx, rdy, rdx2, rdy2;   // reciprocals for div_by (correctly rounded quotient without the division sequence)
    long total;
    long uv_plane;                // floats between consecutive samples of u1 / v1 (H*W when dense)
    int uv_pitch;                 // floats between consecutive rows of u1 / v1 (W when dense; plane kernel only)
};

constexpr int kRhsRows = 4;     // output rows per thread: the x-direction window is loaded once for all of them

// thread = column j, kRhsRows consecutive output rows; grid (column chunks, row groups, samples)
template <int K>
__global__ void __launch_bounds__(128)
rhs2d_kernel(const RhsParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p.W) return;
    const int i0 = blockIdx.y * kRhsRows;
    const long b = blockIdx.z;
    const int Wy = p.W + 2 * p.halo;
    const float mu = __ldg(p.mu + b);
    const float* cx = p.ax + ((size_t)b * (p.H + 2 * p.halo) + i0) * p.W + j;     // padded rows i0 .. i0+K+kRhsRows-2
    float xw[K + kRhsRows - 1];
#pragma unroll
    for (int k = 0; k < K + kRhsRows - 1; ++k)
        xw[k] = (i0 + k < p.H + 2 * p.halo) ? ld_stream(cx + (size_t)k * p.W) : 0.f;
#pragma unroll
    for (int q = 0; q < kRhsRows; ++q) {
        const int i = i0 + q;
        if (i >= p.H) break;
        const size_t o = ((size_t)b * p.H + i) * p.W + j;
        const float u1 = ld_stream(p.u1 + o), v1 = ld_stream(p.v1 + o);
        const float* cy = p.ay + ((size_t)b * p.H + i) * Wy + j;                   // padded cols j .. j+K-1
        float yw[K];
#pragma unroll
        for (int k = 0; k < K; ++k) yw[k] = __ldg(cy + k);
        float acc = 0.f;
        if (p.rhs_id & PPNN_RHS_ADV) {
            float fx = 0.f, bx = 0.f, fy = 0.f, by = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fx += p.aa[k] * xw[q + k]; fy += p.aa[k] * yw[k];
                bx += p.ab[k] * xw[q + k]; by += p.ab[k] * yw[k];
            }
            if (p.adv_upwind) {
                const float c1 = xw[q + (K - 1) / 2], c2 = yw[(K - 1) / 2];
                fx = c1 <= 0.f ? fx : (c1 > 0.f ? bx : 0.f);
                fy = c2 <= 0.f ? fy : (c2 > 0.f ? by : 0.f);
            }
            acc = (-u1 * fx) / p.dx - (v1 * fy) / p.dy;
        }
        if (p.rhs_id & PPNN_RHS_DIFF) {
            float lx = 0.f, ly = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) { lx += p.d[k] * xw[q + k]; ly += p.d[k] * yw[k]; }
            const float lap = mu * (lx / p.dx2 + ly / p.dy2);
            acc = (p.rhs_id & PPNN_RHS_ADV) ? acc + lap : lap;
        }
        st_stream(p.out + o, add_reaction(acc, p.rhs_id & ~3, u1, v1));
    }
}

// ---- tiled variant: thread = 4 consecutive columns, marching down the rows of a tile ----
// u1, v1, ax and out rows are 16-byte aligned when W % 4 == 0, so they move as float4; the
// x-direction window is a rolling set of K float4 in registers (every ax row is loaded once per
// tile).  The rows of ay have pitch W + 2*halo and start on arbitrary 4-byte phases: they are
// staged by `cp.async.bulk` (one copy per row, from the enclosing 16-byte boundary) and read back
// as four aligned float4 per row, from which the K+3 window values are picked for the row's `lead`.
#ifndef PPNN_RHS_TR
#define PPNN_RHS_TR 4
#endif
constexpr int kRTileCols = 512;
constexpr int kRPitch = kRTileCols + 16;

template <int K, int LEAD>
__device__ __forceinline__ void rtile_ywindow(const float* srow, int t, float* yw) {
    const float4* c = reinterpret_cast<const float4*>(srow) + t;
    const float4 c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    const float v[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
#pragma unroll
    for (int k = 0; k < K + 3; ++k) yw[k] = v[LEAD + k];
}

template <int K, int TR>
__global__ void __launch_bounds__(128)
rhs2d_tile_kernel(const RhsParams p, const float* ay_end) {
    // every operand of the tile is staged by TMA: [ax: TR+K-1 rows][u1: TR][v1: TR] x 512 (aligned rows),
    // [ay: TR rows x kRPitch] (arbitrary 4-byte phase per row -> lead)
    extern __shared__ __align__(16) float smem_rhs[];
    float* sax = smem_rhs;
    float* su = sax + (TR + K - 1) * kRTileCols;
    float* sv = su + TR * kRTileCols;
    float* say = sv + TR * kRTileCols;
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x;
    const int j0 = blockIdx.x * kRTileCols;
    const int i0 = blockIdx.y * TR;
    const long b = blockIdx.z;
    const int rows = min(TR, p.H - i0);
    const int Wy = p.W + 2 * p.halo;
    const int wy = min(kRTileCols + K - 1, Wy - j0);            // ay floats per row of this tile
    const int wx = min(kRTileCols, p.W - j0);                   // multiple of 4
    const float* ay = p.ay + ((size_t)b * p.H + i0) * Wy + j0;
    if (t == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        const uintptr_t end_al = reinterpret_cast<uintptr_t>(ay_end) & ~(uintptr_t)15;
        uint32_t total = (uint32_t)((rows + K - 1) + 2 * rows) * wx * 4, nb[TR]; uintptr_t a0[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) {
            nb[r] = 0;
            if (r >= rows) continue;
            const uintptr_t a = reinterpret_cast<uintptr_t>(ay + (size_t)r * Wy);
            a0[r] = a & ~(uintptr_t)15;
            nb[r] = (uint32_t)((a - a0[r]) + (size_t)wy * 4 + 15) & ~15u;
            if (a0[r] + nb[r] > end_al) {                        // last bytes of the tensor: stay inside it
                float* dst = say + r * kRPitch;
                for (const float* q = reinterpret_cast<const float*>(end_al); q < ay_end; ++q)
                    dst[(reinterpret_cast<uintptr_t>(q) - a0[r]) >> 2] = *q;
                nb[r] = (uint32_t)(end_al - a0[r]);
            }
            total += nb[r];
        }
        mbar_expect_tx(&bar, total);
        const float* gx = p.ax + ((size_t)b * (p.H + 2 * p.halo) + i0) * p.W + j0;
        for (int r = 0; r < rows + K - 1; ++r) bulk_g2s(sax + r * kRTileCols, gx + (size_t)r * p.W, (uint32_t)wx * 4, &bar);
        const size_t o0 = ((size_t)b * p.H + i0) * p.W + j0;
        for (int r = 0; r < rows; ++r) {
            bulk_g2s(su + r * kRTileCols, p.u1 + o0 + (size_t)r * p.W, (uint32_t)wx * 4, &bar);
            bulk_g2s(sv + r * kRTileCols, p.v1 + o0 + (size_t)r * p.W, (uint32_t)wx * 4, &bar);
        }
#pragma unroll
        for (int r = 0; r < TR; ++r)
            if (nb[r]) bulk_g2s(say + r * kRPitch, reinterpret_cast<const void*>(a0[r]), nb[r], &bar);
    }
    __syncthreads();
    const int jg = j0 + 4 * t;
    const float mu = __ldg(p.mu + b);
    mbar_wait(&bar, 0);
    if (jg >= p.W) return;                                      // W % 4 == 0: a quad is all in or all out
    const float4* xs = reinterpret_cast<const float4*>(sax) + t;
    float xw[K][4];
#pragma unroll
    for (int k = 0; k < K - 1; ++k) {
        const float4 q = xs[k * (kRTileCols / 4)];
        xw[k + 1][0] = q.x; xw[k + 1][1] = q.y; xw[k + 1][2] = q.z; xw[k + 1][3] = q.w;
    }
#pragma unroll
    for (int r = 0; r < TR; ++r) {
        if (r >= rows) break;
#pragma unroll
        for (int k = 0; k < K - 1; ++k) { xw[k][0] = xw[k + 1][0]; xw[k][1] = xw[k + 1][1]; xw[k][2] = xw[k + 1][2]; xw[k][3] = xw[k + 1][3]; }
        {
            const float4 q = xs[(r + K - 1) * (kRTileCols / 4)];
            xw[K - 1][0] = q.x; xw[K - 1][1] = q.y; xw[K - 1][2] = q.z; xw[K - 1][3] = q.w;
        }
        const float4 u4 = reinterpret_cast<const float4*>(su + r * kRTileCols)[t];
        const float4 v4 = reinterpret_cast<const float4*>(sv + r * kRTileCols)[t];
        const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
        float yw[K + 3];
        const float* srow = say + r * kRPitch;
        const int lead = (int)(reinterpret_cast<uintptr_t>(ay + (size_t)r * Wy) & 15) >> 2;      // warp-uniform
        switch (lead) {
        case 0: rtile_ywindow<K, 0>(srow, t, yw); break;
        case 1: rtile_ywindow<K, 1>(srow, t, yw); break;
        case 2: rtile_ywindow<K, 2>(srow, t, yw); break;
        default: rtile_ywindow<K, 3>(srow, t, yw); break;
        }
        // two column pairs per thread, each on packed fp32x2 arithmetic (FFMA2: one issue slot, two columns; the
        // kernel is issue-bound at ~120 scalar instructions per output).  Same operations in the same order per
        // element as the scalar form (sums in tap order starting from fma(a, x, 0), div_by): bit-identical results.
        float res[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = 2 * h;
            const float2 u2 = make_float2(uu[c], uu[c + 1]), v2 = make_float2(vv[c], vv[c + 1]);
            auto X = [&](int k) { return make_float2(xw[k][c], xw[k][c + 1]); };
            auto Y = [&](int k) { return make_float2(yw[c + k], yw[c + 1 + k]); };
            auto bc = [](float a) { return make_float2(a, a); };
            auto div2 = [&](float2 x, float s, float rcp) {
                const float2 q = __fmul2_rn(x, bc(rcp));
                const float2 rem = __ffma2_rn(q, bc(-s), x);
                return __ffma2_rn(rem, bc(rcp), q);
            };
            float2 acc = bc(0.f);
            if (p.rhs_id & PPNN_RHS_ADV) {
                float2 fx = bc(0.f), fy = bc(0.f);
                if (p.adv_upwind) {
                    // Upwind1: the forward filter is zero left of the centre, the backward one right of
                    // it (operators.py:211-218) -- the zero taps add nothing and are skipped
                    constexpr int C = (K - 1) / 2;
                    float2 bx = bc(0.f), by = bc(0.f);
#pragma unroll
                    for (int k = C; k < K; ++k) { fx = __ffma2_rn(bc(p.aa[k]), X(k), fx); fy = __ffma2_rn(bc(p.aa[k]), Y(k), fy); }
#pragma unroll
                    for (int k = 0; k <= C; ++k) { bx = __ffma2_rn(bc(p.ab[k]), X(k), bx); by = __ffma2_rn(bc(p.ab[k]), Y(k), by); }
                    const float2 c1 = X(C), c2 = Y(C);
                    fx.x = c1.x <= 0.f ? fx.x : (c1.x > 0.f ? bx.x : 0.f); fx.y = c1.y <= 0.f ? fx.y : (c1.y > 0.f ? bx.y : 0.f);
                    fy.x = c2.x <= 0.f ? fy.x : (c2.x > 0.f ? by.x : 0.f); fy.y = c2.y <= 0.f ? fy.y : (c2.y > 0.f ? by.y : 0.f);
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) { fx = __ffma2_rn(bc(p.aa[k]), X(k), fx); fy = __ffma2_rn(bc(p.aa[k]), Y(k), fy); }
                }
                // div_by(-u1 * fx) - div_by(v1 * fy); a - b as fma(b, -1, a): one rounding, same signed zeros
                const float2 un = make_float2(-u2.x, -u2.y);
                acc = __ffma2_rn(div2(__fmul2_rn(v2, fy), p.dy, p.rdy), bc(-1.f), div2(__fmul2_rn(un, fx), p.dx, p.rdx));
            }
            if (p.rhs_id & PPNN_RHS_DIFF) {
                float2 lx = bc(0.f), ly = bc(0.f);
#pragma unroll
                for (int k = 0; k < K; ++k) { lx = __ffma2_rn(bc(p.d[k]), X(k), lx); ly = __ffma2_rn(bc(p.d[k]), Y(k), ly); }
                const float2 lap = __fmul2_rn(bc(mu), __fadd2_rn(div2(lx, p.dx2, p.rdx2), div2(ly, p.dy2, p.rdy2)));
                acc = (p.rhs_id & PPNN_RHS_ADV) ? __fadd2_rn(acc, lap) : lap;
            }
            res[c] = add_reaction(acc.x, p.rhs_id & ~3, u2.x, v2.x);
            res[c + 1] = add_reaction(acc.y, p.rhs_id & ~3, u2.y, v2.y);
        }
        const size_t o = ((size_t)b * p.H + i0 + r) * p.W + jg;
        asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p.out + o), "f"(res[0]), "f"(res[1]), "f"(res[2]), "f"(res[3]) : "memory");
    }
}

// ---- whole-plane variant for small grids (every coarse mesh of the reference: 32^2, 48^2) ----
// The tile kernel above is laid out for rows of >= 512 columns; on a 32-column grid a quarter of one warp
// of each of its CTAs works.  Here CTA = one sample: its four operands (contiguous per sample) are staged in
// shared memory by coalesced loads, thread = column j and row i0 of a sweep of 256 / W rows.  Same
// arithmetic as the tile kernel (div_by, skipped zero taps of Upwind1).
template <int K>
__device__ __forceinline__ float rhs_point(const RhsParams& p, const float* xw, const float* yw, float u1, float v1, float mu) {
    float acc = 0.f;
    if (p.rhs_id & PPNN_RHS_ADV) {
        float fx = 0.f, fy = 0.f;
        if (p.adv_upwind) {
            constexpr int C = (K - 1) / 2;
            float bx = 0.f, by = 0.f;
#pragma unroll
            for (int k = C; k < K; ++k) { fx += p.aa[k] * xw[k]; fy += p.aa[k] * yw[k]; }
#pragma unroll
            for (int k = 0; k <= C; ++k) { bx += p.ab[k] * xw[k]; by += p.ab[k] * yw[k]; }
            const float c1 = xw[C], c2 = yw[C];
            fx = c1 <= 0.f ? fx : (c1 > 0.f ? bx : 0.f);
            fy = c2 <= 0.f ? fy : (c2 > 0.f ? by : 0.f);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) { fx += p.aa[k] * xw[k]; fy += p.aa[k] * yw[k]; }
        }
        acc = div_by(-u1 * fx, p.dx, p.rdx) - div_by(v1 * fy, p.dy, p.rdy);
    }
    if (p.rhs_id & PPNN_RHS_DIFF) {
        float lx = 0.f, ly = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { lx += p.d[k] * xw[k]; ly += p.d[k] * yw[k]; }
        const float lap = mu * (div_by(lx, p.dx2, p.rdx2) + div_by(ly, p.dy2, p.rdy2));
        acc = (p.rhs_id & PPNN_RHS_ADV) ? acc + lap : lap;
    }
    return add_reaction(acc, p.rhs_id & ~3, u1, v1);
}

constexpr int kRPlaneThreads = 256;

template <int K>
__global__ void __launch_bounds__(kRPlaneThreads)
rhs2d_plane_kernel(const RhsParams p) {
    extern __shared__ __align__(16) float smem_rhs[];
    const int Hx = p.H + 2 * p.halo, Wy = p.W + 2 * p.halo;
    const int nx = Hx * p.W, ny = p.H * Wy, no = p.H * p.W;
    float* sax = smem_rhs;
    float* say = sax + ((nx + 3) & ~3);
    float* su = say + ((ny + 3) & ~3);
    float* sv = su + ((no + 3) & ~3);
    const int t = threadIdx.x;
    const long b = blockIdx.x;
    auto stage = [&](float* dst, const float* src, int n) {
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && n % 4 == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int k = t; k < n / 4; k += kRPlaneThreads) d4[k] = __ldcs(s4 + k);
        } else {
            for (int k = t; k < n; k += kRPlaneThreads) dst[k] = ld_stream(src + k);
        }
    };
    stage(sax, p.ax + (size_t)b * nx, nx);
    stage(say, p.ay + (size_t)b * ny, ny);
    if (p.uv_pitch == p.W) {
        stage(su, p.u1 + (size_t)b * p.uv_plane, no);
        stage(sv, p.v1 + (size_t)b * p.uv_plane, no);
    } else {
        // u1 / v1 as the reference passes them: the view down(u)[:, :, :-1, :-1] of a [B,1,N+1,N+1] tensor
        // (train2D.py:68-69) -- read in place instead of through a contiguous copy
        const float* gu = p.u1 + (size_t)b * p.uv_plane;
        const float* gv = p.v1 + (size_t)b * p.uv_plane;
        for (int k = t; k < no; k += kRPlaneThreads) {
            const int i = k / p.W, j = k - i * p.W;
            su[k] = ld_stream(gu + (size_t)i * p.uv_pitch + j);
            sv[k] = ld_stream(gv + (size_t)i * p.uv_pitch + j);
        }
    }
    const float mu = __ldg(p.mu + b);
    __syncthreads();
    const int rps = kRPlaneThreads / p.W;                       // rows per sweep (W <= 128)
    const int ib = t / p.W, j = t - ib * p.W;
    if (ib >= rps) return;
    float* out = p.out + (size_t)b * no;
    for (int i = ib; i < p.H; i += rps) {
        float xw[K], yw[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { xw[k] = sax[(i + k) * p.W + j]; yw[k] = say[i * Wy + j + k]; }
        const int o = i * p.W + j;
        st_stream(out + o, rhs_point<K>(p, xw, yw, su[o], sv[o], mu));
    }
}

template <int K>
static int launch_rhs2d(const RhsParams& p, int B, cudaStream_t s) {
    const size_t plane_smem = (size_t)((((p.H + 2 * p.halo) * p.W + 3) & ~3) + ((p.H * (p.W + 2 * p.halo) + 3) & ~3) +
                                       2 * ((p.H * p.W + 3) & ~3)) * sizeof(float);
    const bool dense = p.uv_pitch == p.W && p.uv_plane == (long)p.H * p.W;
    if (!dense && !(p.W <= 128 && plane_smem <= 100 * 1024)) return PPNN_E_SHAPE;   // strided u1 / v1: whole-plane kernel only
    if (p.W <= 128 && plane_smem <= 100 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(rhs2d_plane_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_smem);
        if (e != cudaSuccess) return (int)e;
        rhs2d_plane_kernel<K><<<(unsigned)B, kRPlaneThreads, plane_smem, s>>>(p);
        return launch_status();
    }
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (K > 1 && p.W % 4 == 0 && al(p.u1) && al(p.v1) && al(p.ax) && al(p.out) &&
        p.H <= 65535) {
        constexpr int TR = PPNN_RHS_TR;
        const size_t smem = ((size_t)(TR + K - 1 + 2 * TR) * kRTileCols + (size_t)TR * kRPitch) * sizeof(float);
        cudaError_t e = cudaFuncSetAttribute(rhs2d_tile_kernel<K, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((p.W + kRTileCols - 1) / kRTileCols, (p.H + TR - 1) / TR, (unsigned)B);
        // a thread owns 4 columns: grids narrower than the 512-column tile launch only the warps that have some
        const int threads = p.W >= kRTileCols ? 128 : ((p.W / 4 + 31) & ~31);
        rhs2d_tile_kernel<K, TR><<<grid, threads, smem, s>>>(p, p.ay + (size_t)B * p.H * (p.W + 2 * p.halo));
        return launch_status();
    }
    dim3 grid((p.W + 127) / 128, (p.H + kRhsRows - 1) / kRhsRows, (unsigned)B);
    rhs2d_kernel<K><<<grid, 128, 0, s>>>(p);
    return launch_status();
}

}  // namespace ppnn

using namespace ppnn;

static int run_rhs2d(const float* u1, const float* v1, const float* mu,
                     const float* ax, const float* ay, float* out,
                     int B, int H, int W, int halo, int rhs_id,
                     int adv_scheme, int adv_accuracy, int diff_scheme, int diff_accuracy,
                     float dx, float dy, float dx2, float dy2, long uv_plane, int uv_pitch, void* stream) {
    if (!u1 || !v1 || !mu || !ax || !ay || !out) return PPNN_E_NULL;
    if (B < 0 || H < 1 || W < 1 || halo < 0 || uv_pitch < W || uv_plane < (long)(H - 1) * uv_pitch + W) return PPNN_E_SHAPE;
    if (!valid_rhs_id(rhs_id)) return PPNN_E_RHS;
    RhsParams p;
    p.K = 2 * halo + 1;
    if (p.K > kMaxTaps) return PPNN_E_HALO;
    Taps ta, tb, td;
    for (int k = 0; k < kMaxTaps; ++k) p.aa[k] = p.ab[k] = p.d[k] = 0.f;
    if (rhs_id & PPNN_RHS_ADV) {
        if (!make_taps(adv_scheme, adv_accuracy, &ta, &tb)) return PPNN_E_SCHEME;
        if (ta.n != p.K) return PPNN_E_HALO;
        for (int k = 0; k < p.K; ++k) { p.aa[k] = ta.c[k]; p.ab[k] = adv_scheme == PPNN_UPWIND1 ? tb.c[k] : 0.f; }
    }
    if (rhs_id & PPNN_RHS_DIFF) {
        if (!make_taps(diff_scheme, diff_accuracy, &td, nullptr)) return PPNN_E_SCHEME;
        if (td.n != p.K || diff_scheme == PPNN_UPWIND1) return PPNN_E_HALO;
        for (int k = 0; k < p.K; ++k) p.d[k] = td.c[k];
    }
    if (B == 0) return PPNN_OK;
    p.u1 = u1; p.v1 = v1; p.mu = mu; p.ax = ax; p.ay = ay; p.out = out;
    p.H = H; p.W = W; p.halo = halo; p.rhs_id = rhs_id; p.adv_upwind = adv_scheme == PPNN_UPWIND1;
    p.dx = dx; p.dy = dy; p.dx2 = dx2; p.dy2 = dy2;
    p.rdx = 1.f / dx; p.rdy = 1.f / dy; p.rdx2 = 1.f / dx2; p.rdy2 = 1.f / dy2;
    p.total = (long)B * H * W;
    p.uv_plane = uv_plane; p.uv_pitch = uv_pitch;
    if ((H + kRhsRows - 1) / kRhsRows > 65535) return PPNN_E_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    // samples ride on gridDim.z (<= 65535): larger batches go out as several launches
    const size_t axs = (size_t)(H + 2 * halo) * W, ays = (size_t)H * (W + 2 * halo), os = (size_t)H * W;
    for (long b0 = 0; b0 < B; b0 += 65535) {
        const int n = (int)(B - b0 < 65535 ? B - b0 : 65535);
        p.u1 = u1 + b0 * uv_plane; p.v1 = v1 + b0 * uv_plane; p.mu = mu + b0; p.ax = ax + b0 * axs; p.ay = ay + b0 * ays;
        p.out = out + b0 * os;
        p.total = (long)n * H * W;
        int st;
        switch (p.K) {
        case 1: st = launch_rhs2d<1>(p, n, s); break;
        case 3: st = launch_rhs2d<3>(p, n, s); break;
        case 5: st = launch_rhs2d<5>(p, n, s); break;
        case 7: st = launch_rhs2d<7>(p, n, s); break;
        case 9: st = launch_rhs2d<9>(p, n, s); break;
        default: return PPNN_E_HALO;
        }
        if (st != PPNN_OK) return st;
    }
    return PPNN_OK;
}

extern "C" int ppnn_rhs2d_fwd(const float* u1, const float* v1, const float* mu,
                              const float* ax, const float* ay, float* out,
                              int B, int H, int W, int halo, int rhs_id,
                              int adv_scheme, int adv_accuracy, int diff_scheme, int diff_accuracy,
                              float dx, float dy, float dx2, float dy2, void* stream) {
    return run_rhs2d(u1, v1, mu, ax, ay, out, B, H, W, halo, rhs_id, adv_scheme, adv_accuracy, diff_scheme, diff_accuracy,
                     dx, dy, dx2, dy2, (long)H * W, W, stream);
}

extern "C" int ppnn_rhs2d_fwd_strided(const float* u1, const float* v1, const float* mu,
                                      const float* ax, const float* ay, float* out,
                                      int B, int H, int W, int halo, int rhs_id,
                                      int adv_scheme, int adv_accuracy, int diff_scheme, int diff_accuracy,
                                      float dx, float dy, float dx2, float dy2,
                                      long uv_plane_stride, int uv_row_pitch, void* stream) {
    return run_rhs2d(u1, v1, mu, ax, ay, out, B, H, W, halo, rhs_id, adv_scheme, adv_accuracy, diff_scheme, diff_accuracy,
                     dx, dy, dx2, dy2, uv_plane_stride, uv_row_pitch, stream);
}

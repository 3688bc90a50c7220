// Tier B: the reference's SMALL explicit solvers, many time steps per launch with the state on chip.
//
//   * 2-D Kuramoto-Sivashinsky, classic RK4          src/operators.py:750-843 (f at :806-809, rk4 :805-815)
//   * 1-D Kuramoto-Sivashinsky, classic RK4          src/operators.py:845-918 (f at :883-884), src/rhs.py:49-50
//   * 1-D heat equation with Dirichlet / Neumann
//     ghost cells (`padder`) + the generator loop    Bv/data.py:9-28, 92-97
//   * its coarse-grid increment (`genforward`)       Bv/data.py:39-51
//
// These meshes are tiny (65^2, 129, 126 points per sample) and the reference advances them for 10^4..2*10^5
// steps with ~60-250 ATen launches per step, i.e. it is launch-bound by three orders of magnitude.  Here ONE
// CTA owns ONE sample: the RK4 stage vector ping-pongs between two shared-memory buffers, the state and the
// RK4 accumulator stay in registers, the periodic wrap (duplicated-endpoint convention of the generators:
// neighbour (i+d) taken modulo n-1 on the n-point array) is index arithmetic, and `nsteps` steps run inside
// one launch with one __syncthreads per stage (two for the 2-D case, whose mixed fourth derivative needs the
// second-derivative fields of the stage first).  HBM is touched once on entry and once on exit, so the bound is
// shared-memory issue rate, not bandwidth; there is nothing to stream and nothing for tensor cores.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace ppnn {

constexpr int kMaxHalo = 4;

// all stencils of a solver re-centred on a common halo: out[i] = sum_{d=-HALO..HALO} c[d+HALO] * f[i+d]
struct CenteredTaps { float c[2 * kMaxHalo + 1]; };

static void centre(const Taps& t, int halo, CenteredTaps* o) {
    for (int k = 0; k < 2 * kMaxHalo + 1; ++k) o->c[k] = 0.f;
    for (int k = 0; k < t.n; ++k) o->c[t.off + k + halo] = t.c[k];
}

// neighbour index on an n-point axis of period P (P = n-1: duplicated endpoint, P = n: true periodic)
__device__ __forceinline__ int wrap_axis(int j, int n, int P) { return j < 0 ? j + P : (j >= n ? j - P : j); }

// ------------------------------------------------------------------------------------------------------------
// 1-D Kuramoto-Sivashinsky
// ------------------------------------------------------------------------------------------------------------
struct Ks1dParams {
    const float* u; const float* source; float* out;
    int n, P, nsteps;
    int rhs_only;        // write f(u) instead of stepping (the reference's `f` closure)
    float dt, dx, rdx, dx2, rdx2, dx4, rdx4;
    CenteredTaps d2, d4, af, ab;
};

template <int HALO, bool WRAP>
__device__ __forceinline__ float ks1d_rhs(const Ks1dParams& p, const float* __restrict__ s, int i, float src) {
    float w[2 * HALO + 1];
#pragma unroll
    for (int d = -HALO; d <= HALO; ++d) w[d + HALO] = s[WRAP ? wrap_axis(i + d, p.n, p.P) : i + d];
    float l2 = 0.f, l4 = 0.f, fw = 0.f, bw = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * HALO + 1; ++k) {
        const float q = w[k] * w[k];
        l2 = fmaf(p.d2.c[k], w[k], l2);
        l4 = fmaf(p.d4.c[k], w[k], l4);
        fw = fmaf(p.af.c[k], q, fw);
        bw = fmaf(p.ab.c[k], q, bw);
    }
    const float c = w[HALO] * w[HALO];
    // (u_int <= 0) * forward + (u_int > 0) * backward on the differentiated field u*u (operators.py:126-127)
    const float adv = (c <= 0.f ? fw : 0.f) + (c > 0.f ? bw : 0.f);
    // - d2/dx2 - d4/dx4 - 0.5*d(u^2)/dx + source, left to right
    float r = -div_by(l2, p.dx2, p.rdx2);
    r -= div_by(l4, p.dx4, p.rdx4);
    r -= div_by(0.5f * adv, p.dx, p.rdx);
    return r + src;
}

// RK4 bookkeeping shared by both KS kernels: stage 0..3 consumes k = f(stage input) and returns the next stage
// input (or the new state after stage 3), in the reference's operation order
//   u + dt*k/2, u + dt*k/2, u + dt*k, u + dt*(k1 + 2*k2 + 2*k3 + k4)/6.
__device__ __forceinline__ float rk4_advance(int stage, float k, float& U, float& A, float dt) {
    switch (stage) {
    case 0: A = k;                 return U + (dt * k) * 0.5f;
    case 1: A = fmaf(2.f, k, A);   return U + (dt * k) * 0.5f;
    case 2: A = fmaf(2.f, k, A);   return U + dt * k;
    default: {
        const float s = dt * (A + k);
        U = U + div_by(s, 6.f, 1.f / 6.f);
        return U;
    }
    }
}

template <int HALO, int PTS>
__global__ void __launch_bounds__(1024)
ks1d_rk4_kernel(const Ks1dParams p) {
    extern __shared__ float sm[];
    float* cur = sm;
    float* nxt = sm + p.n;
    const int T = blockDim.x;
    const float* uin = p.u + (size_t)blockIdx.x * p.n;
    float U[PTS], A[PTS], SRC[PTS];
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
        const int i = threadIdx.x + q * T;
        U[q] = 0.f; A[q] = 0.f; SRC[q] = 0.f;
        if (i < p.n) {
            U[q] = uin[i];
            cur[i] = U[q];
            if (p.source) SRC[q] = __ldg(p.source + i);
        }
    }
    __syncthreads();
    if (p.rhs_only) {
        float* o = p.out + (size_t)blockIdx.x * p.n;
#pragma unroll
        for (int q = 0; q < PTS; ++q) {
            const int i = threadIdx.x + q * T;
            if (i < p.n) o[i] = ks1d_rhs<HALO, true>(p, cur, i, SRC[q]);
        }
        return;
    }
    for (int step = 0; step < p.nsteps; ++step) {
#pragma unroll
        for (int stage = 0; stage < 4; ++stage) {
#pragma unroll
            for (int q = 0; q < PTS; ++q) {
                const int i = threadIdx.x + q * T;
                if (i < p.n) {
                    const bool inner = i >= HALO && i < p.n - HALO;
                    const float k = inner ? ks1d_rhs<HALO, false>(p, cur, i, SRC[q])
                                          : ks1d_rhs<HALO, true>(p, cur, i, SRC[q]);
                    nxt[i] = rk4_advance(stage, k, U[q], A[q], p.dt);
                }
            }
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
        }
    }
    float* o = p.out + (size_t)blockIdx.x * p.n;
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
        const int i = threadIdx.x + q * T;
        if (i < p.n) o[i] = U[q];
    }
}

template <int HALO>
static int launch_ks1d(const Ks1dParams& p, int B, void* stream) {
    int pts = 1;
    while ((p.n + pts - 1) / pts > 1024) pts *= 2;
    if (pts > 8) return PPNN_E_TOOBIG;
    const int T = (((p.n + pts - 1) / pts + 31) / 32) * 32;
    const size_t smem = (size_t)2 * p.n * sizeof(float);
    void (*kern)(Ks1dParams) = pts == 1 ? ks1d_rk4_kernel<HALO, 1> : pts == 2 ? ks1d_rk4_kernel<HALO, 2>
                             : pts == 4 ? ks1d_rk4_kernel<HALO, 4> : ks1d_rk4_kernel<HALO, 8>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<B, T, smem, (cudaStream_t)stream>>>(p);
    return launch_status();
}

// ------------------------------------------------------------------------------------------------------------
// 2-D Kuramoto-Sivashinsky
// ------------------------------------------------------------------------------------------------------------
struct Ks2dParams {
    const float* u; float* out;
    int H, W, PH, PW, nsteps;
    int rhs_only;        // write f(u) instead of stepping (the reference's `f` closure)
    int FH, FW;          // padded frame: (H + 2*HALO) x (W + 2*HALO)
    float invW, invFW;
    float c2, mixed, dt;
    float dx, rdx, dy, rdy, dx2, rdx2, dy2, rdy2, dx4, rdx4, dy4, rdy4;
    CenteredTaps d2, d4, af, ab;
};

// The stage vector lives in shared memory as a PADDED FRAME: HALO ghost rows / columns around the H x W points hold
// the periodic images the reference's pads create (operators.py:761-765), so every stencil reads linearly -- no wrap
// arithmetic and, more important on a 65-point row, no warp that has to run both a wrapped and an interior code path.
// The ghosts cost no extra pass: the thread that produces a point near an edge also stores its periodic image(s).
//
// phase A: second-derivative fields of the stage input, LX = d2/dx2 on rows [0,H) x frame columns and LY = d2/dy2 on
// frame rows x columns [0,W) (their ghosts are evaluated directly, from the ghosts of S); their second derivatives
// across the other axis are the mixed term - d2udy2(padbcy(d2udx2(padbcx(u)))) - d2udx2(padbcx(d2udy2(padbcy(u)))),
// operators.py:808.
template <int HALO>
__device__ __forceinline__ void ks2d_second(const Ks2dParams& p, const float* __restrict__ s, float* __restrict__ LX,
                                            float* __restrict__ LY, int cell, int fi, int fj) {
    const bool rowok = fi >= HALO && fi < p.H + HALO, colok = fj >= HALO && fj < p.W + HALO;
    if (rowok) {
        float ax = 0.f;
#pragma unroll
        for (int d = -HALO; d <= HALO; ++d) ax = fmaf(p.d2.c[d + HALO], s[cell + d * p.FW], ax);
        LX[cell] = ax;
    }
    if (colok) {
        float ay = 0.f;
#pragma unroll
        for (int d = -HALO; d <= HALO; ++d) ay = fmaf(p.d2.c[d + HALO], s[cell + d], ay);
        LY[cell] = ay;
    }
}

// phase B: f at interior point `cell` of the frame
template <int HALO>
__device__ __forceinline__ float ks2d_rhs(const Ks2dParams& p, const float* __restrict__ s,
                                          const float* __restrict__ LX, const float* __restrict__ LY, int cell) {
    float d4x = 0.f, d4y = 0.f, fx = 0.f, bx = 0.f, fy = 0.f, by = 0.f, mxy = 0.f, myx = 0.f;
#pragma unroll
    for (int d = -HALO; d <= HALO; ++d) {
        const int ox = cell + d * p.FW, oy = cell + d;
        const float sx = s[ox], sy = s[oy];
        d4x = fmaf(p.d4.c[d + HALO], sx, d4x);
        d4y = fmaf(p.d4.c[d + HALO], sy, d4y);
        const float qx = sx * sx, qy = sy * sy;
        fx = fmaf(p.af.c[d + HALO], qx, fx);
        bx = fmaf(p.ab.c[d + HALO], qx, bx);
        fy = fmaf(p.af.c[d + HALO], qy, fy);
        by = fmaf(p.ab.c[d + HALO], qy, by);
        mxy = fmaf(p.d2.c[d + HALO], LX[oy], mxy);      // d2/dy2 of the d2/dx2 field
        myx = fmaf(p.d2.c[d + HALO], LY[ox], myx);      // d2/dx2 of the d2/dy2 field
    }
    const float c = s[cell] * s[cell];
    // (u_int <= 0) * forward + (u_int > 0) * backward on the differentiated field u*u (operators.py:257-258)
    const float advx = (c <= 0.f ? fx : 0.f) + (c > 0.f ? bx : 0.f);
    const float advy = (c <= 0.f ? fy : 0.f) + (c > 0.f ? by : 0.f);
    // - c2*Lx/dx2 - c2*Ly/dy2 - D4x/dx4 - D4y/dy4 - Lyy(Lx) - Lxx(Ly) - 0.5*d(u^2)/dx /dx - 0.5*d(u^2)/dy /dy
    float r = -div_by(p.c2 * LX[cell], p.dx2, p.rdx2);
    r -= div_by(p.c2 * LY[cell], p.dy2, p.rdy2);
    r -= div_by(d4x, p.dx4, p.rdx4);
    r -= div_by(d4y, p.dy4, p.rdy4);
    r -= p.mixed * mxy;
    r -= p.mixed * myx;
    r -= div_by(0.5f * advx, p.dx, p.rdx);
    r -= div_by(0.5f * advy, p.dy, p.rdy);
    return r;
}

// store a stage value at its point and at the periodic images of that point inside the ghost frame: along an axis of
// n points and period P the image below the first point is j - P when j in [P-HALO, P), the one above the last is
// j + P when j in [n-P, n-P+HALO)  (duplicated endpoint, P = n-1: u[P-3..P-1] | u | u[1..3], operators.py:761-765)
template <int HALO>
__device__ __forceinline__ void ks2d_store(const Ks2dParams& p, float* __restrict__ s, int i, int j, float v) {
    const int base = (i + HALO) * p.FW + (j + HALO);
    int dj = 0, di = 0;
    if (j >= p.PW - HALO && j < p.PW) dj = -p.PW; else if (j >= p.W - p.PW && j < p.W - p.PW + HALO) dj = p.PW;
    if (i >= p.PH - HALO && i < p.PH) di = -p.PH; else if (i >= p.H - p.PH && i < p.H - p.PH + HALO) di = p.PH;
    s[base] = v;
    if (dj) s[base + dj] = v;
    if (di) {
        s[base + di * p.FW] = v;
        if (dj) s[base + di * p.FW + dj] = v;
    }
    // a point can sit in BOTH image ranges of an axis only when the period is shorter than 2*HALO: rejected on the host
}

// threads per CTA: 1024 (64 registers) while a thread owns <= 2 points, 640 (96 registers) beyond
constexpr int ks2d_max_threads(int pts) { return pts <= 2 ? 1024 : 640; }

template <int HALO, int PTS>
__global__ void __launch_bounds__(ks2d_max_threads(PTS))
ks2d_rk4_kernel(const Ks2dParams p) {
    extern __shared__ float sm[];
    const int N = p.H * p.W, NF = p.FH * p.FW;
    float* cur = sm;
    float* nxt = sm + NF;
    float* LX = sm + 2 * NF;
    float* LY = sm + 3 * NF;
    const int T = blockDim.x;
    const float* uin = p.u + (size_t)blockIdx.x * N;
    float U[PTS], A[PTS];
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
        const int idx = threadIdx.x + q * T;
        U[q] = 0.f; A[q] = 0.f;
        if (idx < N) {
            const int i = __float2int_rd(((float)idx + 0.5f) * p.invW), j = idx - i * p.W;
            U[q] = uin[idx];
            ks2d_store<HALO>(p, cur, i, j, U[q]);
        }
    }
    __syncthreads();
    if (p.rhs_only) {
#pragma unroll 1
        for (int cell = threadIdx.x; cell < NF; cell += T) {
            const int fi = __float2int_rd(((float)cell + 0.5f) * p.invFW), fj = cell - fi * p.FW;
            ks2d_second<HALO>(p, cur, LX, LY, cell, fi, fj);
        }
        __syncthreads();
        float* o = p.out + (size_t)blockIdx.x * N;
#pragma unroll 1
        for (int idx = threadIdx.x; idx < N; idx += T) {
            const int i = __float2int_rd(((float)idx + 0.5f) * p.invW), j = idx - i * p.W;
            o[idx] = ks2d_rhs<HALO>(p, cur, LX, LY, (i + HALO) * p.FW + j + HALO);
        }
        return;
    }
    for (int step = 0; step < p.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
#pragma unroll 2
            for (int cell = threadIdx.x; cell < NF; cell += T) {
                const int fi = __float2int_rd(((float)cell + 0.5f) * p.invFW), fj = cell - fi * p.FW;
                ks2d_second<HALO>(p, cur, LX, LY, cell, fi, fj);
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < PTS; ++q) {
                const int idx = threadIdx.x + q * T;
                if (idx < N) {
                    const int i = __float2int_rd(((float)idx + 0.5f) * p.invW), j = idx - i * p.W;
                    const float k = ks2d_rhs<HALO>(p, cur, LX, LY, (i + HALO) * p.FW + j + HALO);
                    ks2d_store<HALO>(p, nxt, i, j, rk4_advance(stage, k, U[q], A[q], p.dt));
                }
            }
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
        }
    }
    float* o = p.out + (size_t)blockIdx.x * N;
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
        const int idx = threadIdx.x + q * T;
        if (idx < N) o[idx] = U[q];
    }
}

template <int HALO, int PTS>
static int launch_ks2d_pts(const Ks2dParams& p, int B, size_t smem, void* stream) {
    const int N = p.H * p.W;
    const int T = (((N + PTS - 1) / PTS + 31) / 32) * 32;
    void (*kern)(Ks2dParams) = ks2d_rk4_kernel<HALO, PTS>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<B, T, smem, (cudaStream_t)stream>>>(p);
    return launch_status();
}

template <int HALO>
static int launch_ks2d_single(Ks2dParams& p, int B, void* stream) {
    const int N = p.H * p.W;
    p.FH = p.H + 2 * HALO; p.FW = p.W + 2 * HALO;
    p.invFW = 1.f / (float)p.FW;
    const size_t smem = (size_t)4 * p.FH * p.FW * sizeof(float);
    if (smem > 227 * 1024) return PPNN_E_TOOBIG;
    auto fits = [N](int pts) { return (N + pts - 1) / pts <= ks2d_max_threads(pts); };
    if (fits(1)) return launch_ks2d_pts<HALO, 1>(p, B, smem, stream);
    if (fits(2)) return launch_ks2d_pts<HALO, 2>(p, B, smem, stream);
    if (fits(4)) return launch_ks2d_pts<HALO, 4>(p, B, smem, stream);
    if (fits(8)) return launch_ks2d_pts<HALO, 8>(p, B, smem, stream);
    if (fits(16)) return launch_ks2d_pts<HALO, 16>(p, B, smem, stream);
    return launch_ks2d_pts<HALO, 24>(p, B, smem, stream);
}

// ------------------------------------------------------------------------------------------------------------
// 2-D KS for FEW samples: one sample per THREAD-BLOCK CLUSTER (the reference runs a single 65^2 sample,
// operators.py:754-757, which would leave 147 of 148 SMs idle).  The CS CTAs of a cluster own row slabs of the frame;
// each keeps its slab + HALO ghost rows above and below in its own shared memory.  The thread that produces a stage
// value stores it into every frame that shows it: its own, the ghost rows of the slab above / below (distributed
// shared memory, st.shared::cluster through map_shared_rank) and -- for the rows next to the period's ends -- the
// ghost rows of the first / last CTA, which hold the periodic images.  One cluster barrier per stage orders those
// remote stores against the next stage's reads; with ping-pong stage buffers nothing else is needed.
// ------------------------------------------------------------------------------------------------------------
struct Ks2dSlab { int r0, n; };                       // first owned row, owned rows

__host__ __device__ inline Ks2dSlab ks2d_slab(int H, int CS, int r) {
    const int base = H / CS, rem = H % CS;
    Ks2dSlab s;
    s.r0 = r * base + (r < rem ? r : rem);
    s.n = base + (r < rem ? 1 : 0);
    return s;
}

template <int HALO, int CS, int PTS>
__global__ void __launch_bounds__(512)
ks2d_rk4_cluster_kernel(const Ks2dParams p) {
    extern __shared__ float sm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / CS;
    const Ks2dSlab me = ks2d_slab(p.H, CS, rank);
    const int LFHmax = (p.H + CS - 1) / CS + 2 * HALO;          // same frame pitch in every CTA (remote offsets)
    const int NFl = LFHmax * p.FW;
    float* cur = sm;
    float* nxt = sm + NFl;
    float* LX = sm + 2 * NFl;
    float* LY = sm + 3 * NFl;
    const int T = blockDim.x;
    const int N = p.H * p.W;
    const float* uin = p.u + (size_t)b * N;
    const int LFH = me.n + 2 * HALO;                            // this CTA's frame rows

    // ---- frame of the initial state: every cell straight from global memory, periodic images by index ----
    for (int cell = threadIdx.x; cell < LFH * p.FW; cell += T) {
        const int fi = __float2int_rd(((float)cell + 0.5f) * p.invFW), fj = cell - fi * p.FW;
        const int gi = wrap_axis(me.r0 - HALO + fi, p.H, p.PH), gj = wrap_axis(fj - HALO, p.W, p.PW);
        cur[cell] = uin[gi * p.W + gj];
    }

    // ---- this thread's points and where their values have to go ----
    float U[PTS], A[PTS];
    int cellq[PTS], djq[PTS];                                   // local cell; column offset of the periodic image (0: none)
    int adjq[PTS], imgq[PTS];                                   // (rank << 20 | local cell) of the copy in the adjacent slab's
                                                                // ghost rows / in the ghost rows that hold the periodic image; -1: none
    const int npts = me.n * p.W;
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
        const int idx = threadIdx.x + q * T;
        U[q] = 0.f; A[q] = 0.f; cellq[q] = 0; djq[q] = 0; adjq[q] = -1; imgq[q] = -1;
        if (idx < npts) {
            const int li = __float2int_rd(((float)idx + 0.5f) * p.invW), j = idx - li * p.W;
            const int gi = me.r0 + li;
            cellq[q] = (li + HALO) * p.FW + j + HALO;
            U[q] = uin[gi * p.W + j];
            if (j >= p.PW - HALO && j < p.PW) djq[q] = -p.PW;
            else if (j >= p.W - p.PW && j < p.W - p.PW + HALO) djq[q] = p.PW;
            // ghost rows of the neighbouring slabs (slabs hold >= 2*HALO rows: a row is near one edge at most)
            if (li < HALO && rank > 0) {
                const Ks2dSlab up = ks2d_slab(p.H, CS, rank - 1);
                adjq[q] = ((rank - 1) << 20) | ((up.n + HALO + li) * p.FW + j + HALO);
            } else if (li >= me.n - HALO && rank < CS - 1) {
                adjq[q] = ((rank + 1) << 20) | ((li - me.n + HALO) * p.FW + j + HALO);
            }
            // periodic images of the rows: gi - PH in [-HALO, -1] -> ghost rows above slab 0; gi + PH in [H, H+HALO) ->
            // ghost rows below the last slab
            if (gi >= p.PH - HALO && gi < p.PH) {
                imgq[q] = (0 << 20) | ((gi - p.PH + HALO) * p.FW + j + HALO);
            } else if (gi >= p.H - p.PH && gi < p.H - p.PH + HALO) {
                const Ks2dSlab last = ks2d_slab(p.H, CS, CS - 1);
                imgq[q] = ((CS - 1) << 20) | ((gi + p.PH - last.r0 + HALO) * p.FW + j + HALO);
            }
        }
    }
    auto store_all = [&](float* buf, int q, float v) {
        buf[cellq[q]] = v;
        if (djq[q]) buf[cellq[q] + djq[q]] = v;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int tg = t ? imgq[q] : adjq[q];
            if (tg >= 0) {
                float* rb = cluster.map_shared_rank(buf, tg >> 20);
                const int c = tg & 0xfffff;
                rb[c] = v;
                if (djq[q]) rb[c + djq[q]] = v;
            }
        }
    };
    cluster.sync();                                             // every frame initialised before anyone stores into it

    auto phase_a = [&]() {
        for (int cell = threadIdx.x; cell < LFH * p.FW; cell += T) {
            const int fi = __float2int_rd(((float)cell + 0.5f) * p.invFW), fj = cell - fi * p.FW;
            if (fi >= HALO && fi < me.n + HALO) {               // LX on the owned rows, all frame columns
                float ax = 0.f;
#pragma unroll
                for (int d = -HALO; d <= HALO; ++d) ax = fmaf(p.d2.c[d + HALO], cur[cell + d * p.FW], ax);
                LX[cell] = ax;
            }
            if (fj >= HALO && fj < p.W + HALO) {                // LY on all frame rows, real columns
                float ay = 0.f;
#pragma unroll
                for (int d = -HALO; d <= HALO; ++d) ay = fmaf(p.d2.c[d + HALO], cur[cell + d], ay);
                LY[cell] = ay;
            }
        }
    };

    if (p.rhs_only) {
        phase_a();
        __syncthreads();
        float* o = p.out + (size_t)b * N;
#pragma unroll
        for (int q = 0; q < PTS; ++q) {
            const int idx = threadIdx.x + q * T;
            if (idx < npts) o[me.r0 * p.W + idx] = ks2d_rhs<HALO>(p, cur, LX, LY, cellq[q]);
        }
        return;
    }
    for (int step = 0; step < p.nsteps; ++step) {
#pragma unroll 1
        for (int stage = 0; stage < 4; ++stage) {
            phase_a();
            __syncthreads();
#pragma unroll
            for (int q = 0; q < PTS; ++q) {
                const int idx = threadIdx.x + q * T;
                if (idx < npts) {
                    const float k = ks2d_rhs<HALO>(p, cur, LX, LY, cellq[q]);
                    store_all(nxt, q, rk4_advance(stage, k, U[q], A[q], p.dt));
                }
            }
            cluster.sync();                                     // remote ghost rows of `nxt` complete, `cur` free for reuse
            float* t = cur; cur = nxt; nxt = t;
        }
    }
    float* o = p.out + (size_t)b * N;
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
        const int idx = threadIdx.x + q * T;
        if (idx < npts) o[me.r0 * p.W + idx] = U[q];
    }
}

template <int HALO, int CS>
static int launch_ks2d_cluster(Ks2dParams& p, int B, void* stream, bool* used) {
    *used = false;
    if (p.H / CS < 2 * HALO) return PPNN_OK;                    // slabs too thin for the one-edge rule
    p.FH = p.H + 2 * HALO; p.FW = p.W + 2 * HALO;
    p.invFW = 1.f / (float)p.FW;
    const int rows = (p.H + CS - 1) / CS;
    const int npts = rows * p.W;
    if (npts > 2 * 512 || (rows + 2 * HALO) * p.FW >= (1 << 20)) return PPNN_OK;
    const size_t smem = (size_t)4 * (rows + 2 * HALO) * p.FW * sizeof(float);
    if (smem > 200 * 1024) return PPNN_OK;
    const bool two = npts > 512;
    const int per = two ? (npts + 1) / 2 : npts;
    const int T = ((per + 31) / 32) * 32;
    void (*kern)(Ks2dParams) = two ? ks2d_rk4_cluster_kernel<HALO, CS, 2> : ks2d_rk4_cluster_kernel<HALO, CS, 1>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CS), 1, 1);
    cfg.blockDim = dim3((unsigned)T, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e != cudaSuccess) return (int)e;
    *used = true;
    return launch_status();
}

// few samples: spread each over a cluster of 8 (<= 18 samples) or 4 (<= 37) SMs; otherwise one CTA per sample
template <int HALO>
static int launch_ks2d(Ks2dParams& p, int B, void* stream) {
    bool used = false;
    int st = PPNN_OK;
    if (B * 8 <= 148) st = launch_ks2d_cluster<HALO, 8>(p, B, stream, &used);
    if (st == PPNN_OK && !used && B * 4 <= 148) st = launch_ks2d_cluster<HALO, 4>(p, B, stream, &used);
    if (used || st != PPNN_OK) return st;
    return launch_ks2d_single<HALO>(p, B, stream);
}

// common host side: stencil tables of a KS solver on ONE halo
static int ks_taps(int adv_accuracy, int d2_accuracy, int d4_accuracy, int* halo,
                   CenteredTaps* d2, CenteredTaps* d4, CenteredTaps* af, CenteredTaps* ab) {
    Taps t2, t4, tf, tb;
    if (!make_taps(PPNN_CENTRAL2, d2_accuracy, &t2, nullptr) || !make_taps(PPNN_CENTRAL4, d4_accuracy, &t4, nullptr) ||
        !make_taps(PPNN_FORWARD1, adv_accuracy, &tf, nullptr) || !make_taps(PPNN_BACKWARD1, adv_accuracy, &tb, nullptr))
        return PPNN_E_SCHEME;
    int h = -t2.off;
    if (-t4.off > h) h = -t4.off;
    if (adv_accuracy > h) h = adv_accuracy;
    if (h > kMaxHalo) return PPNN_E_SCHEME;
    *halo = h;
    centre(t2, h, d2); centre(t4, h, d4); centre(tf, h, af); centre(tb, h, ab);
    return PPNN_OK;
}

// ------------------------------------------------------------------------------------------------------------
// 1-D heat equation with Dirichlet / Neumann ghost cells (Bv/data.py)
// ------------------------------------------------------------------------------------------------------------
// padder.__call__ (Bv/data.py:25-28): ghost = (cls == 0) * bv + (cls == 1) * (u_edge + bv * dx)
__device__ __forceinline__ float bc_ghost(int cls, float bv, float edge, float dx) {
    return (cls == 0 ? bv : 0.f) + (cls == 1 ? fmaf(bv, dx, edge) : 0.f);
}

struct HeatParams {
    const float* u; const int* cls; const float* bv; float* out; float* padded;
    long padded_stride;
    int n, nsteps;
    float dt, dx, dx2, rdx2;
};

template <int PTS>
__global__ void __launch_bounds__(1024)
heat1d_bc_kernel(const HeatParams p) {
    extern __shared__ float sm[];
    const int b = blockIdx.x, T = blockDim.x, n = p.n;
    float* cur = sm;               // [n + 2]: ghost, interior, ghost
    float* nxt = sm + n + 2;
    const float* uin = p.u + (size_t)b * n;
    const int c0 = p.cls[2 * b], c1 = p.cls[2 * b + 1];
    const float b0 = p.bv[2 * b], b1 = p.bv[2 * b + 1];
    float U[PTS];
#pragma unroll
    for (int q = 0; q < PTS; ++q) {
        const int i = threadIdx.x + q * T;
        U[q] = 0.f;
        if (i < n) {
            U[q] = uin[i];
            cur[i + 1] = U[q];
            if (i == 0) cur[0] = bc_ghost(c0, b0, U[q], p.dx);
            if (i == n - 1) cur[n + 1] = bc_ghost(c1, b1, U[q], p.dx);
        }
    }
    __syncthreads();
    if (p.padded) {
        float* o = p.padded + (size_t)b * p.padded_stride;
        for (int i = threadIdx.x; i < n + 2; i += T) o[i] = cur[i];
    }
    for (int step = 0; step < p.nsteps; ++step) {
#pragma unroll
        for (int q = 0; q < PTS; ++q) {
            const int i = threadIdx.x + q * T;
            if (i < n) {
                // u + diffusion1D(up)/dx2*dt (Bv/data.py:94), conv taps [1,-2,1] accumulated left to right
                const float lap = fmaf(-2.f, cur[i + 1], cur[i]) + cur[i + 2];
                U[q] = U[q] + div_by(lap, p.dx2, p.rdx2) * p.dt;
                nxt[i + 1] = U[q];
                if (i == 0) nxt[0] = bc_ghost(c0, b0, U[q], p.dx);
                if (i == n - 1) nxt[n + 1] = bc_ghost(c1, b1, U[q], p.dx);
            }
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    if (p.out) {
        float* o = p.out + (size_t)b * n;
#pragma unroll
        for (int q = 0; q < PTS; ++q) {
            const int i = threadIdx.x + q * T;
            if (i < n) o[i] = U[q];
        }
    }
}

// genforward (Bv/data.py:39-51): restrict (linear, align_corners) -> one coarse heat step on the interior ->
// coarse ghost cells -> prolong (linear) -> minus the fine state
struct HeatCoarseParams {
    const float* u; const int* cls; const float* bv; float* out;
    int nf, nc;
    float dt, cdx, cdx2, rcdx2, sdown, sup;
};

__device__ __forceinline__ float lerp_at(const float* __restrict__ f, int nin, float scale, int o) {
    // torch upsample_linear1d, align_corners=True: src = scale*o, i0 = (int)src, i1 = i0 + (i0 < nin-1)
    const float s = scale * (float)o;
    const int i0 = min((int)s, nin - 1);
    const int i1 = i0 + (i0 < nin - 1 ? 1 : 0);
    const float l1 = s - (float)i0, l0 = 1.f - l1;
    return l0 * f[i0] + l1 * f[i1];
}

__global__ void __launch_bounds__(256)
heat1d_coarse_increment_kernel(const HeatCoarseParams p) {
    extern __shared__ float sm[];
    float* uc = sm;              // [nc]
    float* u1 = sm + p.nc;       // [nc]
    const int b = blockIdx.x;
    const float* uin = p.u + (size_t)b * p.nf;
    for (int J = threadIdx.x; J < p.nc; J += blockDim.x) uc[J] = lerp_at(uin, p.nf, p.sdown, J);
    __syncthreads();
    for (int J = threadIdx.x + 1; J < p.nc - 1; J += blockDim.x) {
        const float lap = fmaf(-2.f, uc[J], uc[J - 1]) + uc[J + 1];
        u1[J] = uc[J] + div_by(lap, p.cdx2, p.rcdx2) * p.dt;          // uc[1:-1] + fn(uc)/cdx2*dt
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u1[0] = bc_ghost(p.cls[2 * b], p.bv[2 * b], u1[1], p.cdx);
        u1[p.nc - 1] = bc_ghost(p.cls[2 * b + 1], p.bv[2 * b + 1], u1[p.nc - 2], p.cdx);
    }
    __syncthreads();
    float* o = p.out + (size_t)b * p.nf;
    for (int i = threadIdx.x; i < p.nf; i += blockDim.x) o[i] = lerp_at(u1, p.nc, p.sup, i) - uin[i];
}

__global__ void __launch_bounds__(256)
heat1d_pad_kernel(const float* __restrict__ u, const int* __restrict__ cls, const float* __restrict__ bv,
                  float* __restrict__ out, long total, int n, float dx) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / (n + 2);
        const int i = (int)(idx - b * (n + 2));
        const float* row = u + b * n;
        float v;
        if (i == 0) v = bc_ghost(cls[2 * b], bv[2 * b], row[0], dx);
        else if (i == n + 1) v = bc_ghost(cls[2 * b + 1], bv[2 * b + 1], row[n - 1], dx);
        else v = row[i - 1];
        out[idx] = v;
    }
}

// rollout update of the Bv workload (Bv/train.py:162, Bv/test.py:153): padder(u[1:-1] + nn + pdeu[1:-1], cls, bv)
__global__ void __launch_bounds__(256)
heat1d_update_kernel(const float* __restrict__ u, const float* __restrict__ nn, const float* __restrict__ pdeu,
                     const int* __restrict__ cls, const float* __restrict__ bv, float* __restrict__ out,
                     long total, int n, float dx) {
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const long b = idx / (n + 2);
        const int i = (int)(idx - b * (n + 2));
        const float* ur = u + b * (n + 2);
        const float* pr = pdeu ? pdeu + b * (n + 2) : nullptr;
        const float* nr = nn + b * n;
        auto inner = [&](int k) -> float {                        // interior point k of the new state
            float v = ur[k + 1] + nr[k];
            if (pr) v += pr[k + 1];
            return v;
        };
        float v;
        if (i == 0) v = bc_ghost(cls[2 * b], bv[2 * b], inner(0), dx);
        else if (i == n + 1) v = bc_ghost(cls[2 * b + 1], bv[2 * b + 1], inner(n - 1), dx);
        else v = inner(i - 1);
        out[idx] = v;
    }
}

}  // namespace ppnn

using namespace ppnn;

static int ks1d_entry(const float* u, const float* source, float* u_out, int B, int n, int endpoint,
                      int nsteps, int rhs_only, int adv_accuracy, int d2_accuracy, int d4_accuracy,
                      float dt, float dx, float dx2, float dx4, void* stream) {
    if (!u || !u_out) return PPNN_E_NULL;
    if (B < 0 || nsteps < 0) return PPNN_E_SHAPE;
    Ks1dParams p;
    int halo;
    int st = ks_taps(adv_accuracy, d2_accuracy, d4_accuracy, &halo, &p.d2, &p.d4, &p.af, &p.ab);
    if (st != PPNN_OK) return st;
    p.n = n; p.P = endpoint ? n - 1 : n;
    if (p.P <= 2 * halo) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    p.u = u; p.source = source; p.out = u_out; p.nsteps = nsteps; p.rhs_only = rhs_only;
    p.dt = dt; p.dx = dx; p.rdx = 1.f / dx; p.dx2 = dx2; p.rdx2 = 1.f / dx2; p.dx4 = dx4; p.rdx4 = 1.f / dx4;
    switch (halo) {
    case 2: return launch_ks1d<2>(p, B, stream);
    case 3: return launch_ks1d<3>(p, B, stream);
    case 4: return launch_ks1d<4>(p, B, stream);
    }
    return PPNN_E_SCHEME;
}

extern "C" int ppnn_ks1d_rk4(const float* u, const float* source, float* u_out, int B, int n, int endpoint,
                             int nsteps, int adv_accuracy, int d2_accuracy, int d4_accuracy,
                             float dt, float dx, float dx2, float dx4, void* stream) {
    return ks1d_entry(u, source, u_out, B, n, endpoint, nsteps, 0, adv_accuracy, d2_accuracy, d4_accuracy,
                      dt, dx, dx2, dx4, stream);
}

extern "C" int ppnn_ks1d_rhs(const float* u, const float* source, float* f_out, int B, int n, int endpoint,
                             int adv_accuracy, int d2_accuracy, int d4_accuracy,
                             float dx, float dx2, float dx4, void* stream) {
    if (u == f_out) return PPNN_E_MODE;
    return ks1d_entry(u, source, f_out, B, n, endpoint, 0, 1, adv_accuracy, d2_accuracy, d4_accuracy,
                      0.f, dx, dx2, dx4, stream);
}

static int ks2d_entry(const float* u, float* u_out, int B, int H, int W, int endpoint, int nsteps, int rhs_only,
                      int adv_accuracy, int d2_accuracy, int d4_accuracy, float c2, float mixed_scale,
                      float dt, float dx, float dy, float dx2, float dy2, float dx4, float dy4,
                      void* stream) {
    if (!u || !u_out) return PPNN_E_NULL;
    if (B < 0 || nsteps < 0 || H < 2 || W < 2) return PPNN_E_SHAPE;
    Ks2dParams p;
    int halo;
    int st = ks_taps(adv_accuracy, d2_accuracy, d4_accuracy, &halo, &p.d2, &p.d4, &p.af, &p.ab);
    if (st != PPNN_OK) return st;
    p.H = H; p.W = W; p.PH = endpoint ? H - 1 : H; p.PW = endpoint ? W - 1 : W;
    if (p.PH <= 2 * halo || p.PW <= 2 * halo) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    p.u = u; p.out = u_out; p.nsteps = nsteps; p.rhs_only = rhs_only;
    p.invW = 1.f / (float)W;
    p.c2 = c2; p.mixed = mixed_scale; p.dt = dt;
    p.dx = dx; p.rdx = 1.f / dx; p.dy = dy; p.rdy = 1.f / dy;
    p.dx2 = dx2; p.rdx2 = 1.f / dx2; p.dy2 = dy2; p.rdy2 = 1.f / dy2;
    p.dx4 = dx4; p.rdx4 = 1.f / dx4; p.dy4 = dy4; p.rdy4 = 1.f / dy4;
    switch (halo) {
    case 2: return launch_ks2d<2>(p, B, stream);
    case 3: return launch_ks2d<3>(p, B, stream);
    case 4: return launch_ks2d<4>(p, B, stream);
    }
    return PPNN_E_SCHEME;
}

extern "C" int ppnn_ks2d_rk4(const float* u, float* u_out, int B, int H, int W, int endpoint, int nsteps,
                             int adv_accuracy, int d2_accuracy, int d4_accuracy, float c2, float mixed_scale,
                             float dt, float dx, float dy, float dx2, float dy2, float dx4, float dy4,
                             void* stream) {
    return ks2d_entry(u, u_out, B, H, W, endpoint, nsteps, 0, adv_accuracy, d2_accuracy, d4_accuracy, c2, mixed_scale,
                      dt, dx, dy, dx2, dy2, dx4, dy4, stream);
}

extern "C" int ppnn_ks2d_rhs(const float* u, float* f_out, int B, int H, int W, int endpoint,
                             int adv_accuracy, int d2_accuracy, int d4_accuracy, float c2, float mixed_scale,
                             float dx, float dy, float dx2, float dy2, float dx4, float dy4, void* stream) {
    if (u == f_out) return PPNN_E_MODE;
    return ks2d_entry(u, f_out, B, H, W, endpoint, 0, 1, adv_accuracy, d2_accuracy, d4_accuracy, c2, mixed_scale,
                      0.f, dx, dy, dx2, dy2, dx4, dy4, stream);
}

extern "C" int ppnn_heat1d_pad(const float* u, const int* cls, const float* bv, float* out,
                               int B, int n, float dx, void* stream) {
    if (!u || !cls || !bv || !out) return PPNN_E_NULL;
    if (B < 0 || n < 1) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    const long total = (long)B * (n + 2);
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    heat1d_pad_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(u, cls, bv, out, total, n, dx);
    return launch_status();
}

extern "C" int ppnn_heat1d_update(const float* u, const float* nn, const float* pdeu, const int* cls, const float* bv,
                                  float* out, int B, int n, float dx, void* stream) {
    if (!u || !nn || !cls || !bv || !out) return PPNN_E_NULL;
    if (B < 0 || n < 1) return PPNN_E_SHAPE;
    if (out == u || out == pdeu) return PPNN_E_MODE;
    if (B == 0) return PPNN_OK;
    const long total = (long)B * (n + 2);
    long blocks = (total + 255) / 256;
    if (blocks > 148L * 16) blocks = 148L * 16;
    heat1d_update_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(u, nn, pdeu, cls, bv, out, total, n, dx);
    return launch_status();
}

extern "C" int ppnn_heat1d_bc_steps(const float* u, const int* cls, const float* bv, float* u_out,
                                    float* padded_out, long padded_stride, int B, int n, int nsteps,
                                    float dt, float dx, float dx2, void* stream) {
    if (!u || !cls || !bv) return PPNN_E_NULL;
    if (!u_out && !padded_out) return PPNN_E_NULL;
    if (B < 0 || n < 1 || nsteps < 0) return PPNN_E_SHAPE;
    if (padded_out && padded_stride < n + 2) return PPNN_E_SHAPE;
    if (B == 0) return PPNN_OK;
    HeatParams p;
    p.u = u; p.cls = cls; p.bv = bv; p.out = u_out; p.padded = padded_out; p.padded_stride = padded_stride;
    p.n = n; p.nsteps = nsteps; p.dt = dt; p.dx = dx; p.dx2 = dx2; p.rdx2 = 1.f / dx2;
    const size_t smem = (size_t)2 * (n + 2) * sizeof(float);
    if (smem > 227 * 1024) return PPNN_E_TOOBIG;
    int pts = 1;
    while (((n + pts - 1) / pts) > 1024) pts *= 2;
    if (pts > 8) {
        if (n > 16 * 1024) return PPNN_E_TOOBIG;
        pts = 16;
    }
    const int T = (((n + pts - 1) / pts + 31) / 32) * 32;
    void (*kern)(HeatParams) = pts == 1 ? heat1d_bc_kernel<1> : pts == 2 ? heat1d_bc_kernel<2>
                             : pts == 4 ? heat1d_bc_kernel<4> : pts == 8 ? heat1d_bc_kernel<8> : heat1d_bc_kernel<16>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<B, T, smem, (cudaStream_t)stream>>>(p);
    return launch_status();
}

extern "C" int ppnn_heat1d_coarse_increment(const float* u, const int* cls, const float* bv, float* out,
                                            int B, int nf, int nc, float dt, float cdx, float cdx2,
                                            void* stream) {
    if (!u || !cls || !bv || !out) return PPNN_E_NULL;
    if (B < 0 || nf < 2 || nc < 3) return PPNN_E_SHAPE;
    if (u == out) return PPNN_E_MODE;
    if (B == 0) return PPNN_OK;
    HeatCoarseParams p;
    p.u = u; p.cls = cls; p.bv = bv; p.out = out; p.nf = nf; p.nc = nc;
    p.dt = dt; p.cdx = cdx; p.cdx2 = cdx2; p.rcdx2 = 1.f / cdx2;
    p.sdown = resample_scale(nf, nc, true);
    p.sup = resample_scale(nc, nf, true);
    const size_t smem = (size_t)2 * nc * sizeof(float);
    if (smem > 48 * 1024) return PPNN_E_TOOBIG;
    heat1d_coarse_increment_kernel<<<B, 256, smem, (cudaStream_t)stream>>>(p);
    return launch_status();
}

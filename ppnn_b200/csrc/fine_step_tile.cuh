// Column-tiled TMA variant of the fine-grid solver step for WIDE true-periodic meshes
// (256 ... 8192 columns; included by fine_step.cu).
//
// Whole rows in the shared-memory ring stop fitting at ~1088 columns.  Here the plane is cut into `ntile` column tiles of Wt <= 1024 columns and
// the unit of work is one row of one (sample, tile).  CTAs come in groups of `ntile`: the CTAs of
// a group walk the SAME contiguous range of (sample, row) pairs, one column tile each, so that at
// any moment the group streams whole rows (DRAM pages are not visited in 4 KB pieces at a 16-32 KB
// stride).  A ring row holds [4 halo | Wt | 4 halo] floats, so
//   * every copy is a 16-byte-aligned `cp.async.bulk` (W, Wt and the tile origins are
//     multiples of 4 floats): one copy per row for interior tiles, main + wrapped halo
//     chunks for the first / last tile -- no ordinary loads anywhere in the producer;
//   * the y-direction taps are plain immediates: the periodic column wrap is resolved by
//     the producer, there are no seam warps.
// Consumer side: lanes own columns t and t + Wt/2
// packed in float2 (FFMA2), x-direction taps in a register window fed from the ring.
#pragma once
#include <type_traits>
#include "tma.cuh"

namespace ppnn {

constexpr int kTmaRows = 4;            // rows per ring band = rows per compute batch
constexpr int kTmaMaxRing = 8;

constexpr int kTileHalo = 4;           // halo floats kept on each side of a ring row (>= HL, 16-byte chunk)

struct FineTilePlan {
    long total_units;      // B * H (sample, row) pairs
    long units_per_cta;    // ... per CTA group
    int ntile, Wt;         // column tiles per plane, columns per tile
    int lanes;             // Wt / 2
    int cwarps;
    int ring;
    int band_floats;       // floats per (slot, field): kTmaRows * (Wt + 8)
};

// One explicit update of a packed pair of grid points from its x-direction windows (xu, xv: 2*HL+1
// rows centred on the point) and y-direction windows (yu, yv).
// UNI: take the warp-uniform upwind fast path (it pays in the instruction-bound endpoint kernel, not in the
// memory-bound tile kernel)
template <int HL, int MODE, bool UNI = false>
__device__ __forceinline__ void fine_point_update(const FineParams& p, float mu, float2 c, float2 vcen,
                                                  const float2* xu, const float2* xv, const float2* yu, const float2* yv,
                                                  float2& unew, float2& vnew) {
    if (MODE == FINE_GENERIC) {
        float2 ru = f2(0.f), rv = f2(0.f);
        if (p.rhs_u & PPNN_RHS_ADV)
            ru = add2(mul2(mul2(neg2(c), upwind2<HL>(p, c, xu)), f2(p.rdx)),
                      neg2(mul2(mul2(vcen, upwind2<HL>(p, c, yu)), f2(p.rdy))));
        if (p.rhs_u & PPNN_RHS_DIFF) {
            float2 lap = mul2(f2(mu), add2(mul2(central2<HL>(p, xu), f2(p.rdx2)), mul2(central2<HL>(p, yu), f2(p.rdy2))));
            ru = (p.rhs_u & PPNN_RHS_ADV) ? add2(ru, lap) : lap;
        }
        ru = react2(ru, p.rhs_u & ~3, c, vcen);
        if (p.rhs_v & PPNN_RHS_ADV)
            rv = add2(mul2(mul2(neg2(c), upwind2<HL>(p, vcen, xv)), f2(p.rdx)),
                      neg2(mul2(mul2(vcen, upwind2<HL>(p, vcen, yv)), f2(p.rdy))));
        if (p.rhs_v & PPNN_RHS_DIFF) {
            float2 lap = mul2(f2(mu), add2(mul2(central2<HL>(p, xv), f2(p.rdx2)), mul2(central2<HL>(p, yv), f2(p.rdy2))));
            rv = (p.rhs_v & PPNN_RHS_ADV) ? add2(rv, lap) : lap;
        }
        rv = react2(rv, p.rhs_v & ~3, c, vcen);
        unew = fma2(p.dt, ru, c);
        vnew = fma2(p.dt, rv, vcen);
    } else {
        // burgers2Dpu then burgers2Dpv with the updated u (operators.py:474-475)
        constexpr bool VDER = MODE != FINE_AS_WRITTEN;          // quirk Q3: as written, v re-uses u's derivative fields
        float2 dxu, dyu, dxv = f2(0.f), dyv = f2(0.f);
        // (the fast path is taken in as-written mode only: with v's windows live as well the two code paths cost
        //  the intended-mode kernel its registers -- 0.285 -> 0.41 ms at 64 x 1025^2 when tried)
        if (UNI && !VDER && all_positive(c, vcen, false)) {
            dxu = upwind2_bwd<HL>(p, xu); dyu = upwind2_bwd<HL>(p, yu);
        } else {
            dxu = upwind2<HL>(p, c, xu); dyu = upwind2<HL>(p, c, yu);
            if (VDER) { dxv = upwind2<HL>(p, vcen, xv); dyv = upwind2<HL>(p, vcen, yv); }
        }
        const float2 lapu = mul2(f2(mu), add2(mul2(central2<HL>(p, xu), f2(p.rdx2)), mul2(central2<HL>(p, yu), f2(p.rdy2))));
        const float2 advy = neg2(mul2(mul2(vcen, dyu), f2(p.rdy)));
        unew = fma2(p.dt, add2(add2(mul2(mul2(neg2(c), dxu), f2(p.rdx)), advy), lapu), c);
        if (!VDER) {
            vnew = fma2(p.dt, add2(add2(mul2(mul2(neg2(unew), dxu), f2(p.rdx)), advy), lapu), vcen);
        } else {
            const float2 lapv = mul2(f2(mu), add2(mul2(central2<HL>(p, xv), f2(p.rdx2)), mul2(central2<HL>(p, yv), f2(p.rdy2))));
            vnew = fma2(p.dt, add2(add2(mul2(mul2(neg2(unew), dxv), f2(p.rdx)),
                                        neg2(mul2(mul2(vcen, dyv), f2(p.rdy)))), lapv), vcen);
        }
    }
}

template <int HL, int MODE, int WTC>
__global__ void __launch_bounds__(544)
fine_step_tile_kernel(const FineParams p, const FineTilePlan pl) {
    static_assert(HL <= 3, "band geometry assumes a halo of at most 3 rows / columns");
    constexpr int R = kTmaRows;
    constexpr bool VWIN = MODE != FINE_AS_WRITTEN;
    const int Wt = WTC ? WTC : pl.Wt;
    const int SW = Wt + 2 * kTileHalo;                           // ring row pitch
    const int W = p.W;                                           // global row pitch
    extern __shared__ float4 smem4[];
    float* ringbuf = reinterpret_cast<float*>(smem4);            // [ring][2][R][SW]
    __shared__ __align__(8) uint64_t full[kTmaMaxRing];
    __shared__ __align__(8) uint64_t empty[kTmaMaxRing];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == pl.cwarps;
    const size_t plane = (size_t)p.H * W;
    const int tile = (int)(blockIdx.x % pl.ntile);              // this CTA's column tile
    const long g0 = (long)(blockIdx.x / pl.ntile) * pl.units_per_cta;
    const long g1 = min(g0 + pl.units_per_cta, pl.total_units);

    if (tid == 0) {
        for (int s = 0; s < pl.ring; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], pl.cwarps); }
        fence_barrier_init();
    }
    __syncthreads();

    if (producer) {
        // lane = (piece, field, row of the band): the 8 (or 24, first / last tile) copies of a band are
        // issued by one instruction each; a single lane walking them serially was the bottleneck
        const int piece = lane >> 3, pf = (lane >> 2) & 1, pj = lane & 3;
        long cidx = 0;
        int pslot = 0; uint32_t pphase = 0;
        for (long g = g0; g < g1;) {
            const int b = (int)(g / p.H);
            const int r0 = (int)(g % p.H);
            const int c0 = tile * Wt;
            const int r1 = (int)min((long)p.H, r0 + (g1 - g));
            const int nb = (r1 - r0 + R - 1) / R;
            const float* fsrc = (pf ? p.v : p.u) + (size_t)b * plane;
            const bool interior = c0 > 0 && c0 + Wt < W;
            // this lane's piece of a ring row: source column, destination offset, bytes
            int scol, doff; uint32_t bytes;
            if (interior)        { scol = c0 - kTileHalo; doff = 0; bytes = piece == 0 ? (uint32_t)SW * 4 : 0u; }
            else if (piece == 0) { scol = c0; doff = kTileHalo; bytes = (uint32_t)Wt * 4; }
            else if (piece == 1) { scol = c0 > 0 ? c0 - kTileHalo : W - kTileHalo; doff = 0; bytes = kTileHalo * 4; }
            else if (piece == 2) { scol = c0 + Wt < W ? c0 + Wt : 0; doff = kTileHalo + Wt; bytes = kTileHalo * 4; }
            else                 { scol = 0; doff = 0; bytes = 0; }
            for (int c = 0; c < nb + 2; ++c, ++cidx) {
                const int slot = pslot;
                if (cidx >= pl.ring) mbar_wait(&empty[slot], pphase ^ 1);
                if (++pslot == pl.ring) { pslot = 0; pphase ^= 1; }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (lane == 0) mbar_expect_tx(&full[slot], (uint32_t)(2 * R * SW * 4));
                __syncwarp();
                if (bytes) {
                    const int k = wrap_pt(r0 - R + c * R + pj, p.H, p.PH);
                    bulk_g2s(ringbuf + ((size_t)slot * 2 + pf) * pl.band_floats + pj * SW + doff,
                             fsrc + (size_t)k * W + scol, bytes, &full[slot]);
                }
            }
            g += r1 - r0;
        }
        return;
    }

    // -------------------------------- consumer warps --------------------------------
    const bool active = tid < pl.lanes;
    const int colA = active ? tid : 0;
    const int dB = pl.lanes;                                     // column B = column A + Wt / 2
    const int slot_floats = 2 * pl.band_floats;
    const int offU = kTileHalo + colA, offV = pl.band_floats + kTileHalo + colA;
    int slot = 0; uint32_t phase = 0;
    auto next_slot = [&](int& sl, uint32_t& ph) { if (++sl == pl.ring) { sl = 0; ph ^= 1; } };

    for (long g = g0; g < g1;) {
        const int b = (int)(g / p.H);
        const int r0 = (int)(g % p.H);
        const int r1 = (int)min((long)p.H, r0 + (g1 - g));
        const int nb = (r1 - r0 + R - 1) / R;
        float* ouA = p.uo + (size_t)b * plane + (size_t)r0 * W + tile * Wt + colA;
        float* ovA = p.vo + (size_t)b * plane + (size_t)r0 * W + tile * Wt + colA;
        const float mu = __ldg(p.mu + b);

        int s0 = slot; uint32_t ph0 = phase;
        int s1 = s0; uint32_t ph1 = ph0; next_slot(s1, ph1);
        mbar_wait(&full[s0], ph0);
        mbar_wait(&full[s1], ph1);
        float2 wu[2 * HL + R], wv[VWIN ? 2 * HL + R : 1];
        {
            const float* b0 = ringbuf + (size_t)s0 * slot_floats;
            const float* b1 = ringbuf + (size_t)s1 * slot_floats;
#pragma unroll
            for (int k = 0; k < HL; ++k) {
                const float* ua = b0 + offU + (R - HL + k) * SW;
                const float* ub2 = b1 + offU + k * SW;
                wu[k] = make_float2(ua[0], ua[dB]);
                wu[HL + k] = make_float2(ub2[0], ub2[dB]);
                if (VWIN) {
                    const float* va = b0 + offV + (R - HL + k) * SW;
                    const float* vb2 = b1 + offV + k * SW;
                    wv[k] = make_float2(va[0], va[dB]);
                    wv[HL + k] = make_float2(vb2[0], vb2[dB]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s0]);
        int sc = s1; uint32_t phc = ph1;
        int sn = sc; uint32_t phn = phc; next_slot(sn, phn);

        for (int n = 0; n < nb; ++n) {
            mbar_wait(&full[sn], phn);
            const float* cb = ringbuf + (size_t)sc * slot_floats;
            const float* nbp = ringbuf + (size_t)sn * slot_floats;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int jr = q + HL;
                const float* fb = jr < R ? cb + jr * SW : nbp + (jr - R) * SW;
                wu[2 * HL + q] = make_float2(fb[offU], fb[offU + dB]);
                if (VWIN) wv[2 * HL + q] = make_float2(fb[offV], fb[offV + dB]);
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const float* ur = cb + offU + q * SW;
                const float* vr = cb + offV + q * SW;
                float2 yu[2 * HL + 1], yv[VWIN ? 2 * HL + 1 : 1];
#pragma unroll
                for (int k = 0; k <= 2 * HL; ++k) {
                    yu[k] = make_float2(ur[k - HL], ur[dB + k - HL]);
                    if (VWIN) yv[k] = make_float2(vr[k - HL], vr[dB + k - HL]);
                }
                const float2 c = wu[HL + q];
                const float2 vcen = VWIN ? wv[HL + q] : make_float2(vr[0], vr[dB]);
                float2 unew, vnew;
                fine_point_update<HL, MODE>(p, mu, c, vcen, wu + q, wv + (VWIN ? q : 0), yu, yv, unew, vnew);
                if (active && r0 + n * R + q < r1) {
                    ouA[(size_t)q * W] = unew.x; ovA[(size_t)q * W] = vnew.x;
                    ouA[(size_t)q * W + dB] = unew.y; ovA[(size_t)q * W + dB] = vnew.y;
                }
            }
#pragma unroll
            for (int k = 0; k < 2 * HL; ++k) {
                wu[k] = wu[k + R];
                if (VWIN) wv[k] = wv[k + R];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[sc]);
            sc = sn; phc = phn;
            next_slot(sn, phn);
            ouA += (size_t)R * W; ovA += (size_t)R * W;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[sc]);
        slot = sn; phase = phn;
        g += r1 - r0;
    }
}

template <int HL, int MODE>
static int launch_fine_tile(const FineParams& p, int B, void* stream, bool* used) {
    *used = false;
    if (PPNN_ENV_OFF("PPNN_NO_TMA") || PPNN_ENV_OFF("PPNN_NO_TILE")) return PPNN_OK;
    // true periodic, 16-byte aligned rows, too wide for whole rows in the ring
    if (p.PW != p.W || p.PH != p.H || p.W % 8 != 0 || p.W < 256 || p.H < 2 * kTmaRows) return PPNN_OK;
    const uintptr_t al = reinterpret_cast<uintptr_t>(p.u) | reinterpret_cast<uintptr_t>(p.v);
    if (al & 15) return PPNN_OK;
    FineTilePlan pl;
    pl.ntile = (p.W + 1023) / 1024;
    if (p.W % pl.ntile != 0) return PPNN_OK;
    pl.Wt = p.W / pl.ntile;
    if (pl.Wt % 8 != 0 || pl.Wt < 256) return PPNN_OK;
    pl.lanes = pl.Wt / 2;
    pl.cwarps = (pl.lanes + 31) / 32;
    pl.band_floats = kTmaRows * (pl.Wt + 2 * kTileHalo);
    pl.total_units = (long)B * p.H;
    void (*kern)(FineParams, FineTilePlan) = pl.Wt == 1024 ? fine_step_tile_kernel<HL, MODE, 1024> : fine_step_tile_kernel<HL, MODE, 0>;
    // narrow tiles: several persistent CTAs per SM (4 consumer warps each at 256 columns), smaller rings
    int cpsm = 1024 / pl.Wt;
    if (cpsm < 1) cpsm = 1;
    if (cpsm > 4) cpsm = 4;
    const size_t band_bytes = (size_t)2 * pl.band_floats * sizeof(float);
    int ring = (int)((200 * 1024 / cpsm) / band_bytes);
    if (ring > kTmaMaxRing) ring = kTmaMaxRing;
    if (ring < 4) return PPNN_OK;
    pl.ring = ring;
    const size_t smem = (size_t)ring * band_bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long ngroups = (long)sms * cpsm / pl.ntile;                 // persistent CTAs, whole groups only
    if (ngroups < 1) ngroups = 1;
    long units = (pl.total_units + ngroups - 1) / ngroups;
    if (units < 4 * kTmaRows) units = 4 * kTmaRows;
    ngroups = (pl.total_units + units - 1) / units;
    pl.units_per_cta = units;
    const long nctas = ngroups * pl.ntile;
    kern<<<(unsigned)nctas, (pl.cwarps + 1) * 32, smem, (cudaStream_t)stream>>>(p, pl);
    *used = true;
    return launch_status();
}

}  // namespace ppnn

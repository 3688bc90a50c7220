// Fine-grid solver step for meshes that keep the DUPLICATED PERIODIC ENDPOINT of the reference
// (257^2, 1025^2: operators.py:416-420) -- rows of 4*W bytes with W odd start on every 4-byte
// phase, so neither whole-band bulk copies with one phase nor aligned halo chunks exist
// (included by fine_step.cu after fine_step_tile.cuh).
//
// The reference pads [u[P-3..P-1] | u[0..P] | u[1..3]] (P = W-1) and reads neighbours LINEARLY in
// the padded array.  The ring rows reproduce exactly that layout, so the consumers are the seam-free
// ones of fine_step_tile_kernel:
//
//   producer warp, stage 1 (8 lanes, one instruction): one `cp.async.bulk` per (field, row) of a
//       4-row band from the enclosing 16-byte boundary of the row; the row's phase (`lead`, 0..3
//       floats) is kept in a small table next to the ring;
//   producer warp, stage 2 (ring-1 bands later, when the band has long landed): 48 lanes-iterations
//       copy the 3 + 3 halo floats of every row inside shared memory (u[P-3..P-1] in front of
//       column 0, u[1..3] behind column P), the last <= 3 floats of the tensor are fetched by
//       ordinary loads, and only then is the band published on its `ready` barrier;
//   consumer warps: lanes own columns t and t + ceil(W/2) packed in float2 (FFMA2); all taps are
//       immediates relative to the row's lead.
// Several persistent CTAs share an SM on narrow meshes (257: 5 consumer warps per CTA).
#pragma once
#include "fine_step_tile.cuh"

namespace ppnn {

struct FineEpPlan {
    long total_units;      // B * H (sample, row) pairs
    long units_per_cta;
    int lanes, cwarps, ring;
    int SW;                // ring row pitch (floats): W + 12 rounded up to 4
    int band_floats;       // per (slot, field): kTmaRows * SW
    const float* u_end;    // one past the last element of u / v
    const float* v_end;
};

// position of the producer in the sequence of bands of its row range
struct BandCursor {
    long g, g1;            // first unit of the current segment, end of the CTA's range
    int H;
    int b, r0, r1, nb, c;  // sample, row range of the segment, its bands - 2, current band
    __device__ __forceinline__ void open() {
        b = (int)(g / H); r0 = (int)(g % H);
        r1 = (int)min((long)H, r0 + (g1 - g));
        nb = (r1 - r0 + kTmaRows - 1) / kTmaRows;
        c = 0;
    }
    __device__ __forceinline__ bool done() const { return g >= g1; }
    __device__ __forceinline__ void next() {
        if (++c == nb + 2) { g += r1 - r0; if (g < g1) open(); }
    }
    __device__ __forceinline__ int first_row() const { return r0 - kTmaRows + c * kTmaRows; }   // unwrapped
};

#ifdef PPNN_TRACE
// dev-only CTA timeline (tools/trace_fine.py); never in the shipped library
__device__ unsigned long long g_trace_fine[4 * 1024];
__device__ __forceinline__ unsigned long long gtime_fine() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PPNN_FSTAMP(k) do { if (threadIdx.x == 0 && blockIdx.x < 1024) g_trace_fine[4 * blockIdx.x + (k)] = gtime_fine(); } while (0)
#else
#define PPNN_FSTAMP(k) do { } while (0)
#endif

template <int HL, int MODE, int WC>
__global__ void __launch_bounds__(576)
fine_step_ep_kernel(const FineParams p, const FineEpPlan pl) {
    static_assert(HL <= 3, "halo of at most 3 rows / columns");
    constexpr int R = kTmaRows;
    constexpr bool VWIN = MODE != FINE_AS_WRITTEN;
    constexpr int FRONT = 4;                                     // floats in front of the copied data (left halo)
    const int W = WC ? WC : p.W;
    const int SW = WC ? ((WC + 12 + 3) & ~3) : pl.SW;
    extern __shared__ float4 smem4[];
    float* ringbuf = reinterpret_cast<float*>(smem4);            // [ring][2][R][SW]
    __shared__ __align__(8) uint64_t landed[kTmaMaxRing];       // TMA bytes of a band have arrived (producer only)
    __shared__ __align__(8) uint64_t ready[kTmaMaxRing];        // halos fixed up: consumers may read
    __shared__ __align__(8) uint64_t empty[kTmaMaxRing];
    __shared__ int leads[kTmaMaxRing][2][R];
    __shared__ unsigned tailmask[kTmaMaxRing];                   // lanes whose row was clamped at the end of the tensor

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == pl.cwarps;
    PPNN_FSTAMP(0);
    const size_t plane = (size_t)p.H * W;
    const long g0 = (long)blockIdx.x * pl.units_per_cta;
    const long g1 = min(g0 + pl.units_per_cta, pl.total_units);
    const int band_floats = R * SW;

    if (tid == 0) {
        for (int s = 0; s < pl.ring; ++s) { mbar_init(&landed[s], 1); mbar_init(&ready[s], 1); mbar_init(&empty[s], pl.cwarps); }
        fence_barrier_init();
    }
    __syncthreads();
    if (g0 >= g1) return;

    if (producer) {
        const int pf = (lane >> 2) & 1, pj = lane & 3;           // lanes 0..7: (field, row of the band)
        BandCursor ci, cf;                                       // issue / fix-up cursors
        ci.g = cf.g = g0; ci.g1 = cf.g1 = g1; ci.H = cf.H = p.H;
        ci.open(); cf.open();
        long ni = 0, nf = 0;                                     // bands issued / fixed so far
        int islot = 0, fslot = 0; uint32_t iphase = 0, fphase = 0;
        // half of the ring is in flight by TMA, the other half is fixed up ahead of the consumers: neither
        // the issue (waits for a slot released long ago) nor the fix-up (band landed long ago) blocks
        const int ahead = pl.ring / 2;
        // this lane's halo elements (element e = lane, lane + 32 of the 2*R rows x 2*HL floats): offsets are constant
        int hoff[2], hdst[2], hsrc[2], hf[2], hj[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int e = lane + 32 * t;
            const int rowi = e / (2 * HL), h = e - rowi * (2 * HL);
            hf[t] = rowi / R; hj[t] = rowi - hf[t] * R;
            hoff[t] = e < 2 * R * 2 * HL ? hf[t] * band_floats + hj[t] * SW + FRONT : -1;
            hdst[t] = h < HL ? h - HL : W + (h - HL);
            hsrc[t] = h < HL ? p.PW - HL + h : (W - p.PW) + (h - HL);
            if (hoff[t] < 0) { hf[t] = 0; hj[t] = 0; }
        }

        auto issue = [&]() {
            if (ni >= pl.ring) mbar_wait(&empty[islot], iphase ^ 1);      // previous use of the slot released
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            uint32_t bytes = 0; uintptr_t a0 = 0; int lead = 0; bool clamped = false;
            if (lane < 2 * R) {
                const float* fsrc = (pf ? p.v : p.u) + (size_t)ci.b * plane;
                const float* fend = pf ? pl.v_end : pl.u_end;
                const int k = wrap_pt(ci.first_row() + pj, p.H, p.PH);
                const uintptr_t a = reinterpret_cast<uintptr_t>(fsrc + (size_t)k * W);
                a0 = a & ~(uintptr_t)15;
                lead = (int)(a - a0) >> 2;
                bytes = (uint32_t)((a - a0) + (size_t)W * 4 + 15) & ~15u;
                const uintptr_t end_al = reinterpret_cast<uintptr_t>(fend) & ~(uintptr_t)15;
                if (a0 + bytes > end_al) { bytes = (uint32_t)(end_al - a0); clamped = true; }   // last row of the tensor: tail by the fix-up
                leads[islot][pf][pj] = lead;
            }
            const unsigned cm = __ballot_sync(0xffffffffu, clamped);
            if (lane == 0) tailmask[islot] = cm;
            uint32_t total = bytes;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);   // lanes 0..7 -> lane 0
            if (lane == 0) mbar_expect_tx(&landed[islot], total);
            __syncwarp();
            if (bytes) bulk_g2s(ringbuf + ((size_t)islot * 2 + pf) * band_floats + pj * SW + FRONT,
                                reinterpret_cast<const void*>(a0), bytes, &landed[islot]);
            if (++islot == pl.ring) { islot = 0; iphase ^= 1; }
            ++ni; ci.next();
        };
        auto fixup = [&]() {
            mbar_wait(&landed[fslot], fphase);
            float* sb = ringbuf + (size_t)fslot * 2 * band_floats;
            const unsigned tm = tailmask[fslot];
            if (tm) {
                // the floats of the tensor's last row that the clamped bulk copy left out (once per launch)
                if ((tm >> lane) & 1u) {
                    const float* fsrc = (pf ? p.v : p.u) + (size_t)cf.b * plane;
                    const float* fend = pf ? pl.v_end : pl.u_end;
                    const float* src = fsrc + (size_t)wrap_pt(cf.first_row() + pj, p.H, p.PH) * W;
                    float* row = sb + pf * band_floats + pj * SW + FRONT + leads[fslot][pf][pj];
                    for (const float* q = reinterpret_cast<const float*>(reinterpret_cast<uintptr_t>(fend) & ~(uintptr_t)15); q < src + W; ++q)
                        row[q - src] = __ldg(q);
                }
                __syncwarp();
            }
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (hoff[t] >= 0) {
                    float* row = sb + hoff[t] + leads[fslot][hf[t]][hj[t]];
                    row[hdst[t]] = row[hsrc[t]];
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[fslot]);
            if (++fslot == pl.ring) { fslot = 0; fphase ^= 1; }
            ++nf; cf.next();
        };

        for (int k = 0; k < ahead && !ci.done(); ++k) issue();
        while (!cf.done()) {
            fixup();
            if (!ci.done()) issue();
        }
        return;
    }

    // -------------------------------- consumer warps --------------------------------
    const bool active = tid < pl.lanes;
    const int colA = active ? tid : 0;
    const bool hasB = active && tid + pl.lanes < W;
    const int dB = hasB ? pl.lanes : 0;                          // column B = column A + ceil(W / 2)
    const int slot_floats = 2 * band_floats;
    const int offU = FRONT + colA, offV = band_floats + FRONT + colA;
    int slot = 0; uint32_t phase = 0;
    auto next_slot = [&](int& sl, uint32_t& ph) { if (++sl == pl.ring) { sl = 0; ph ^= 1; } };
#ifdef PPNN_FINE_NOSTORE
    float sink = 0.f;
#endif

    // Odd widths (W = 2^k + 1) leave the last consumer warp with ONE live column.  Walking the rows with a
    // register window like the others would cost a full warp's instruction stream for it (1/17 of the issue
    // slots at 1025, 1/9 at 513); instead lane q of that warp takes row q of the band and reads all its taps
    // from the three resident bands, i.e. one pass per band instead of one per row.
    const bool lonely = (pl.lanes & 31) == 1 && warp == pl.cwarps - 1;

    for (long g = g0; g < g1;) {
        const int b = (int)(g / p.H);
        const int r0 = (int)(g % p.H);
        const int r1 = (int)min((long)p.H, r0 + (g1 - g));
        const int nb = (r1 - r0 + R - 1) / R;
        float* ouA = p.uo + (size_t)b * plane + (size_t)r0 * W + colA;
        float* ovA = p.vo + (size_t)b * plane + (size_t)r0 * W + colA;
        const float mu = __ldg(p.mu + b);

        if (lonely) {
            const int col = pl.lanes - 1;                          // == colA of lane 0
            float* ou = p.uo + (size_t)b * plane + (size_t)r0 * W + col;
            float* ov = p.vo + (size_t)b * plane + (size_t)r0 * W + col;
            int sp = slot; uint32_t php = phase;                   // bands n-1, n, n+1
            int sc = sp; uint32_t phc = php; next_slot(sc, phc);
            int sn = sc; uint32_t phn = phc; next_slot(sn, phn);
            mbar_wait(&ready[sp], php);
            mbar_wait(&ready[sc], phc);
            for (int n = 0; n < nb; ++n) {
                mbar_wait(&ready[sn], phn);
                if (lane < R) {
                    const int q = lane;
                    float2 xu[2 * HL + 1], xv[VWIN ? 2 * HL + 1 : 1], yu[2 * HL + 1], yv[VWIN ? 2 * HL + 1 : 1];
#pragma unroll
                    for (int k = 0; k <= 2 * HL; ++k) {
                        const int rr = q + k - HL;
                        const int sl = rr < 0 ? sp : rr < R ? sc : sn;
                        const int row = rr < 0 ? rr + R : rr < R ? rr : rr - R;
                        const float* fb = ringbuf + (size_t)sl * slot_floats + row * SW + FRONT + col;
                        xu[k] = make_float2(fb[leads[sl][0][row]], 0.f);
                        if (VWIN) xv[k] = make_float2(fb[band_floats + leads[sl][1][row]], 0.f);
                    }
                    const float* ur = ringbuf + (size_t)sc * slot_floats + q * SW + FRONT + col + leads[sc][0][q];
                    const float* vr = ringbuf + (size_t)sc * slot_floats + band_floats + q * SW + FRONT + col + leads[sc][1][q];
#pragma unroll
                    for (int k = 0; k <= 2 * HL; ++k) {
                        yu[k] = make_float2(ur[k - HL], 0.f);
                        if (VWIN) yv[k] = make_float2(vr[k - HL], 0.f);
                    }
                    const float2 c = xu[HL];
                    const float2 vcen = make_float2(vr[0], 0.f);
                    float2 unew, vnew;
                    fine_point_update<HL, MODE, true>(p, mu, c, vcen, xu, xv, yu, yv, unew, vnew);
                    if (r0 + n * R + q < r1) { ou[(size_t)q * W] = unew.x; ov[(size_t)q * W] = vnew.x; }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[sp]);
                sp = sc; php = phc; sc = sn; phc = phn;
                next_slot(sn, phn);
                ou += (size_t)R * W; ov += (size_t)R * W;
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&empty[sp]); mbar_arrive(&empty[sc]); }
            slot = sn; phase = phn;
            g += r1 - r0;
            continue;
        }

        int s0 = slot; uint32_t ph0 = phase;
        int s1 = s0; uint32_t ph1 = ph0; next_slot(s1, ph1);
        mbar_wait(&ready[s0], ph0);
        mbar_wait(&ready[s1], ph1);
        if (g == g0) PPNN_FSTAMP(1);
        float2 wu[2 * HL + R], wv[VWIN ? 2 * HL + R : 1];
#pragma unroll
        for (int k = 0; k < HL; ++k) {
            const float* ua = ringbuf + (size_t)s0 * slot_floats + offU + (R - HL + k) * SW + leads[s0][0][R - HL + k];
            const float* ub2 = ringbuf + (size_t)s1 * slot_floats + offU + k * SW + leads[s1][0][k];
            wu[k] = make_float2(ua[0], ua[dB]);
            wu[HL + k] = make_float2(ub2[0], ub2[dB]);
            if (VWIN) {
                const float* va = ringbuf + (size_t)s0 * slot_floats + offV + (R - HL + k) * SW + leads[s0][1][R - HL + k];
                const float* vb2 = ringbuf + (size_t)s1 * slot_floats + offV + k * SW + leads[s1][1][k];
                wv[k] = make_float2(va[0], va[dB]);
                wv[HL + k] = make_float2(vb2[0], vb2[dB]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s0]);
        int sc = s1; uint32_t phc = ph1;
        int sn = sc; uint32_t phn = phc; next_slot(sn, phn);

        for (int n = 0; n < nb; ++n) {
            mbar_wait(&ready[sn], phn);
            const float* cb = ringbuf + (size_t)sc * slot_floats;
            const float* nbp = ringbuf + (size_t)sn * slot_floats;
            int lcu[R], lcv[R];
#pragma unroll
            for (int q = 0; q < R; ++q) { lcu[q] = leads[sc][0][q]; lcv[q] = leads[sc][1][q]; }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const int jr = q + HL;
                if (jr < R) {
                    const float* fb = cb + jr * SW;
                    wu[2 * HL + q] = make_float2(fb[offU + lcu[jr]], fb[offU + lcu[jr] + dB]);
                    if (VWIN) wv[2 * HL + q] = make_float2(fb[offV + lcv[jr]], fb[offV + lcv[jr] + dB]);
                } else {
                    const float* fb = nbp + (jr - R) * SW;
                    const int lu = leads[sn][0][jr - R], lv = leads[sn][1][jr - R];
                    wu[2 * HL + q] = make_float2(fb[offU + lu], fb[offU + lu + dB]);
                    if (VWIN) wv[2 * HL + q] = make_float2(fb[offV + lv], fb[offV + lv + dB]);
                }
            }
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const float* ur = cb + offU + q * SW + lcu[q];
                const float* vr = cb + offV + q * SW + lcv[q];
                float2 yu[2 * HL + 1], yv[VWIN ? 2 * HL + 1 : 1];
#pragma unroll
                for (int k = 0; k <= 2 * HL; ++k) {
                    yu[k] = make_float2(ur[k - HL], ur[dB + k - HL]);
                    if (VWIN) yv[k] = make_float2(vr[k - HL], vr[dB + k - HL]);
                }
                const float2 c = wu[HL + q];
                const float2 vcen = VWIN ? wv[HL + q] : make_float2(vr[0], vr[dB]);
                float2 unew, vnew;
                fine_point_update<HL, MODE, true>(p, mu, c, vcen, wu + q, wv + (VWIN ? q : 0), yu, yv, unew, vnew);
#ifdef PPNN_FINE_NOSTORE      // dev-only timing build: the arithmetic stays alive, nothing is written
                sink += unew.x + vnew.x + unew.y + vnew.y;
#else
                if (active && r0 + n * R + q < r1) {
                    ouA[(size_t)q * W] = unew.x; ovA[(size_t)q * W] = vnew.x;
                    if (hasB) { ouA[(size_t)q * W + dB] = unew.y; ovA[(size_t)q * W + dB] = vnew.y; }
                }
#endif
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
#ifdef PPNN_FINE_NOSTORE
    if (sink == 12345.678f) p.uo[0] = sink;
#endif
    PPNN_FSTAMP(2);
}

template <int HL, int MODE, int WC>
static int launch_fine_ep(const FineParams& p, int B, void* stream, bool* used) {
    *used = false;
    if (PPNN_ENV_OFF("PPNN_NO_TMA") || PPNN_ENV_OFF("PPNN_NO_EP")) return PPNN_OK;
    // rows that the aligned tile kernel does not take; whole rows must fit a ring row.  From the reference's
    // own 257^2 up (4 CTAs of 4 + 1 consumer warps per SM) this kernel beats the register-window one
    // (257^2: 0.134 -> 0.124 ms at B = 512, intended mode 0.227 -> 0.174 ms); narrower meshes stay there.
    if (p.W < 257 || p.W > 1088 || p.H < 2 * kTmaRows) return PPNN_OK;
    if ((reinterpret_cast<uintptr_t>(p.u) | reinterpret_cast<uintptr_t>(p.v)) & 15) return PPNN_OK;   // first row: lead 0, nothing in front of it
    FineEpPlan pl;
    pl.lanes = (p.W + 1) / 2;
    pl.cwarps = (pl.lanes + 31) / 32;
    pl.SW = (p.W + 12 + 3) & ~3;
    pl.band_floats = kTmaRows * pl.SW;
    pl.total_units = (long)B * p.H;
    pl.u_end = p.u + (size_t)B * p.H * p.W;
    pl.v_end = p.v + (size_t)B * p.H * p.W;
    auto kern = fine_step_ep_kernel<HL, MODE, WC>;
    // persistent CTAs per SM: two up to 544 columns (513^2: 9 + 1 warps and a 5-band ring each; one CTA of 9
    // consumer warps left the SM at 65 % of the copy bandwidth, two reach 81 %)
    int cpsm = 1088 / p.W;
    if (cpsm < 1) cpsm = 1;
    if (cpsm > 4) cpsm = 4;
    const size_t band_bytes = (size_t)2 * pl.band_floats * sizeof(float);
    int ring = (int)(((cpsm > 1 ? 180 : 200) * 1024 / cpsm) / band_bytes);
    if (ring > kTmaMaxRing) ring = kTmaMaxRing;
    if (ring < 4) return PPNN_OK;
    pl.ring = ring;
    const size_t smem = (size_t)ring * band_bytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = (pl.cwarps + 1) * 32;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
    if (per_sm < 1) return PPNN_OK;
    if (per_sm > cpsm) per_sm = cpsm;
    long nctas = (long)sms * per_sm;
    long units = (pl.total_units + nctas - 1) / nctas;
    if (units < 4 * kTmaRows) units = 4 * kTmaRows;
    nctas = (pl.total_units + units - 1) / units;
    pl.units_per_cta = units;
    kern<<<(unsigned)nctas, threads, smem, (cudaStream_t)stream>>>(p, pl);
    *used = true;
    return launch_status();
}

}  // namespace ppnn

// Shared host/device helpers for libppnn_b200 (sm_100a).
//
// Coefficient tables follow src/operators.py:150-218 of the reference: every
// coefficient is the fp32 rounding of the Python double the reference writes
// (SURVEY.md appendix A.1), so each literal below is (float)(double expression).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ppnn_b200.h"

// Kernel-selection switches for debugging (force a fallback kernel).  The shipped library reads NO
// environment and keeps no global state: they exist only in builds with -DPPNN_DEBUG_ENV.
#ifdef PPNN_DEBUG_ENV
#include <stdlib.h>
#define PPNN_ENV_OFF(name) (getenv(name) != nullptr)
#else
#define PPNN_ENV_OFF(name) false
#endif

namespace ppnn {

constexpr int kMaxTaps = 9;

// out[i] = sum_k c[k] * f[i + off + k]  (un-padded coordinates); for a pre-padded
// 'valid' correlation the k-th tap simply reads in_pad[i + k] for k in [0, n).
struct Taps {
    int   n;
    int   off;
    float c[kMaxTaps];
};

// Host: coefficients of (scheme, accuracy).  For PPNN_UPWIND1 `a` receives the
// forward-biased filter and `b` the backward-biased one, both zero-padded to
// 2*accuracy+1 taps exactly as operators.py:211-218 writes them.
inline bool make_taps(int scheme, int acc, Taps* a, Taps* b) {
    auto set = [](Taps* t, int off, std::initializer_list<double> v) {
        t->n = (int)v.size();
        t->off = off;
        int i = 0;
        for (double d : v) t->c[i++] = (float)d;
        for (; i < kMaxTaps; ++i) t->c[i] = 0.f;
    };
    Taps dummy;
    if (!b) b = &dummy;
    switch (scheme) {
    case PPNN_CENTRAL1:
        if (acc == 2) { set(a, -1, {-.5, 0., .5}); return true; }
        if (acc == 4) { set(a, -2, {1. / 12, -2. / 3, 0., 2. / 3, -1. / 12}); return true; }
        return false;
    case PPNN_CENTRAL2:
        if (acc == 2) { set(a, -1, {1., -2., 1.}); return true; }
        if (acc == 4) { set(a, -2, {-1. / 12, 4. / 3, -5. / 2, 4. / 3, -1. / 12}); return true; }
        if (acc == 6) { set(a, -3, {1. / 90, -3. / 20, 3. / 2, -49. / 18, 3. / 2, -3. / 20, 1. / 90}); return true; }
        if (acc == 8) { set(a, -4, {-1. / 560, 8. / 315, -1. / 5, 8. / 5, -205. / 72, 8. / 5, -1. / 5, 8. / 315, -1. / 560}); return true; }
        return false;
    case PPNN_CENTRAL4:
        if (acc == 2) { set(a, -2, {1., -4., 6., -4., 1.}); return true; }
        if (acc == 4) { set(a, -3, {-1. / 6, 2., -13. / 2, 28. / 3, -13. / 2, 2., -1. / 6}); return true; }
        if (acc == 6) { set(a, -4, {7. / 240, -2. / 5, 169. / 60, -122. / 15, 91. / 8, -122. / 15, 169. / 60, -2. / 5, 7. / 240}); return true; }
        return false;
    case PPNN_FORWARD1:
        if (acc == 1) { set(a, 0, {-1., 1.}); return true; }
        if (acc == 2) { set(a, 0, {-3. / 2, 2., -1. / 2}); return true; }
        if (acc == 3) { set(a, 0, {-11. / 6, 3., -3. / 2, 1. / 3}); return true; }
        return false;
    case PPNN_BACKWARD1:
        if (acc == 1) { set(a, -1, {-1., 1.}); return true; }
        if (acc == 2) { set(a, -2, {1. / 2, -2., 3. / 2}); return true; }
        if (acc == 3) { set(a, -3, {-1. / 3, 3. / 2, -3., 11. / 6}); return true; }
        return false;
    case PPNN_UPWIND1:
        if (acc == 1) { set(a, -1, {0., -1., 1.}); set(b, -1, {-1., 1., 0.}); return true; }
        if (acc == 2) { set(a, -2, {0., 0., -3. / 2, 2., -1. / 2}); set(b, -2, {1. / 2, -2., 3. / 2, 0., 0.}); return true; }
        if (acc == 3) { set(a, -3, {0., 0., 0., -11. / 6, 3., -3. / 2, 1. / 3}); set(b, -3, {-1. / 3, 3. / 2, -3., 11. / 6, 0., 0., 0.}); return true; }
        return false;
    }
    return false;
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PPNN_OK : (int)e;
}

inline bool valid_rhs_id(int id) {
    if (id < 0) return false;
    int react = id & ~3;
    return react == PPNN_REACT_NONE || react == PPNN_REACT_BFU || react == PPNN_REACT_BFV ||
           react == PPNN_REACT_RDU || react == PPNN_REACT_RDV;
}

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__

// torch's area_pixel_compute_scale for align_corners=True / False.
__host__ __device__ inline float resample_scale(int in, int out, bool align) {
    if (align) return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    return (float)in / (float)out;
}

// source coordinate of output index o (torch area_pixel_compute_source_index).
__device__ __forceinline__ float resample_src(float scale, int o, bool align, bool cubic) {
    if (align) return scale * (float)o;
    float s = scale * ((float)o + 0.5f) - 0.5f;
    return (!cubic && s < 0.f) ? 0.f : s;
}

// Keys cubic convolution weights, A = -0.75 (torch get_cubic_upsample_coefficients).
__host__ __device__ __forceinline__ void cubic_weights(float t, float w[4]) {
    const float A = -0.75f;
    float x;
    x = t + 1.f;  w[0] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
    x = t;        w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 1.f - t;  w[2] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    x = 2.f - t;  w[3] = ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// wrap an index that is at most one period outside [0, n)
__device__ __forceinline__ int wrap1(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

// Reaction terms of rhs.py; `acc` is the value of the differential part so far,
// summed in the reference's left-to-right order.
__device__ __forceinline__ float add_reaction(float acc, int react, float u1, float v1) {
    switch (react) {
    case PPNN_REACT_BFU: { float uv = u1 * u1 + v1 * v1; return acc + (1.f - uv) * u1 + uv * v1; }
    case PPNN_REACT_BFV: { float uv = u1 * u1 + v1 * v1; return acc - uv * u1 + (1.f - uv) * v1; }
    case PPNN_REACT_RDU: return acc + u1 - u1 * u1 * u1 - v1 + 0.01f;
    case PPNN_REACT_RDV: return acc + 0.25f * (u1 - v1);
    default: return acc;
    }
}

// x / s for a divisor whose reciprocal r = 1/s is already known: one multiply and
// one Newton correction give the correctly rounded quotient (the IEEE division the
// reference performs) without the slow-path division sequence.
__device__ __forceinline__ float div_by(float x, float s, float r) {
    float q = x * r;
    float rem = fmaf(-q, s, x);
    return fmaf(rem, r, q);
}

// streaming (read-once / write-once) global accesses: keep them out of L1
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void  st_stream(float* p, float v) { __stcs(p, v); }

#endif  // __CUDACC__
}  // namespace ppnn

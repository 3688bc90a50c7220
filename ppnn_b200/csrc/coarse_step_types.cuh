// Types shared by the coarse-step kernels (coarse_step.cu, coarse_step_staged.cuh,
// coarse_dataset.cu): kernel modes, operand bundle, CTA decomposition plan.
#pragma once
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
    float* coarse;         // [B,2,CH-1,CW-1] or null: the restricted coarse state, saved for the adjoint pass
    int inner;             // coarse Euler sub-steps before the increment is evaluated (post_process/test2D2.py:70-91)
    float m0[2], s0[2];    // inmean/instd   (INPUTS) | outmean/outstd (UPDATE)
    float m1[2], s1[2];    // pdemean/pdestd (INPUTS)
};

// Keys-cubic weights of the phases of one prolongation period (kernels specialised for the reference's
// meshes, 257 -> 33 / 49: every PER fine rows the 4-row coarse window moves on by exactly ADV rows and the
// PER phases repeat).  Passed by value as a kernel parameter, so that with the period unrolled at compile
// time the weights are constant-bank operands of the FFMAs.
struct PhaseWeights { float w[16][4]; };

inline PhaseWeights make_phase_weights(float scale, int per) {
    PhaseWeights pw;
    for (int r = 0; r < 16; ++r) {
        const float s = scale * (float)(r < per ? r : 0);
        cubic_weights(s - floorf(s), pw.w[r]);
    }
    return pw;
}

struct StepPlan {
    int wmain;             // columns handled by strips (FW-1 when that is a multiple of 32, else FW)
    int nchunk, nband, band_rows;
};

#ifndef PPNN_STEP_WARPS
#define PPNN_STEP_WARPS 16
#endif
#ifndef PPNN_MINCTA
#define PPNN_MINCTA 2
#endif
#ifndef PPNN_DEPTH
#define PPNN_DEPTH 4
#endif
constexpr int kStepThreads = PPNN_STEP_WARPS * 32;
constexpr int kDepth = PPNN_DEPTH;   // rows of streamed operands in flight per lane (register path)

}  // namespace ppnn

/* ppnn_b200.h -- C ABI of libppnn_b200.so: the PDE-preserving hot path of PPNN
 * (jx-wang-s-group/ppnn) as hand-written sm_100a CUDA kernels.
 *
 * Contract:
 *   - every pointer is a DEVICE pointer to contiguous fp32 owned by the caller;
 *     the library never allocates, frees or retains memory and keeps no state;
 *   - every entry enqueues work on `stream` (a cudaStream_t passed as void*) and
 *     returns without synchronising; entries are re-entrant and graph-capturable;
 *   - return value: 0 = ok, negative = PPNN_E_* (bad argument, nothing launched),
 *     positive = the cudaError_t reported by the launch;
 *   - "x" is tensor dim 2 (H, rows) and "y" is dim 3 (W, columns), as in the
 *     reference (src/operators.py:194-198 builds x-filters by permuting y-filters).
 *
 * Each entry names the reference interface it replaces (paths relative to the
 * reference checkout).
 */
#ifndef PPNN_B200_H
#define PPNN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define PPNN_B200_ABI_VERSION 4

/* ---- status codes --------------------------------------------------------- */
#define PPNN_OK            0
#define PPNN_E_NULL       -1   /* required pointer is NULL                    */
#define PPNN_E_SHAPE      -2   /* non-positive or inconsistent sizes          */
#define PPNN_E_SCHEME     -3   /* unknown scheme / accuracy combination       */
#define PPNN_E_RHS        -4   /* unknown right-hand-side id                  */
#define PPNN_E_TOOBIG     -5   /* coarse grid does not fit in shared memory   */
#define PPNN_E_MODE       -6   /* unknown mode / flag value                   */
#define PPNN_E_HALO       -7   /* pre-padded halo does not match the stencils */

/* ---- finite-difference schemes (src/operators.py:58-61,200-218) ----------- */
enum ppnn_scheme {
    PPNN_CENTRAL1  = 0,   /* 'Central1'  d/dx,    accuracy 2,4                 */
    PPNN_CENTRAL2  = 1,   /* 'Central2'  d2/dx2,  accuracy 2,4,6,8 (1-D: 'Central')     */
    PPNN_CENTRAL4  = 2,   /* 'Central4'  d4/dx4,  accuracy 2,4,6   (1-D: 'Central_4th') */
    PPNN_FORWARD1  = 3,   /* 'Forward1'  accuracy 1,2,3                        */
    PPNN_BACKWARD1 = 4,   /* 'Backward1' accuracy 1,2,3                        */
    PPNN_UPWIND1   = 5    /* 'Upwind1'   accuracy 1,2,3 (1-D: 'Upwind'); switch on sign of the field */
};

/* ---- right-hand sides (src/rhs.py) ---------------------------------------- *
 * id = PPNN_RHS_ADV | PPNN_RHS_DIFF | reaction term.  The sixteen 2-D names of
 * rhs.py map onto these nine ids (ppnn_b200/rhs.py holds the table).           */
#define PPNN_RHS_ADV      1    /* -u1*dudx(a)/dx - v1*dudy(a)/dy   rhs.py:43-47 */
#define PPNN_RHS_DIFF     2    /* mu*(d2udx2(a)/dx2+d2udy2(a)/dy2) rhs.py:17-23 */
#define PPNN_REACT_NONE   0
#define PPNN_REACT_BFU    4    /* + (1-uv)*u1 + uv*v1              rhs.py:1-3   */
#define PPNN_REACT_BFV    8    /* - uv*u1 + (1-uv)*v1              rhs.py:5-7   */
#define PPNN_REACT_RDU   12    /* + u1 - u1^3 - v1 + 0.01          rhs.py:25-26,52-53 */
#define PPNN_REACT_RDV   16    /* + 0.25*(u1 - v1)                 rhs.py:28-29,55-56 */

/* ---- resampling modes (src/utility/utils.py:13-31) ------------------------ */
#define PPNN_BILINEAR 0        /* mesh_convertor.down (also 'linear' with H=1) */
#define PPNN_BICUBIC  1        /* mesh_convertor.up   (Keys A=-0.75, clamped)  */

/* ---- fine solver modes (src/operators.py:474-475) -------- */
#define PPNN_FINE_AS_WRITTEN 0 /* v update sees the updated u and u's halos    */
#define PPNN_FINE_INTENDED   1 /* the commented formula: v's own halos         */

int         ppnn_abi_version(void);
const char* ppnn_status_string(int status);

/* =========================== Tier A: operator API ========================== */

/* d2udx2_2D / d2udy2_2D / dudx_2D / dudy_2D.__call__ (src/operators.py:236-237,
 * 246-247, 255-260, 268-273): 'valid' cross-correlation of `planes` single-
 * channel images [Hin, Win] along `axis` (2 = x/H, 3 = y/W).  Output is
 * [Hin-K+1, Win] or [Hin, Win-K+1], K = number of taps (2a+1 for Upwind1).
 * With Hin = 1, axis = 3 this is diffusion1D / dudx_1D (operators.py:85,126).  */
int ppnn_stencil_fwd(const float* in, float* out, long planes, int Hin, int Win,
                     int axis, int scheme, int accuracy, void* stream);

/* adjoint of the above w.r.t. `in` (what autograd derives from conv2d and the
 * comparison masks).  `in` is only read for PPNN_UPWIND1 (may be NULL otherwise).
 * grad_in has the padded shape [Hin, Win] and is fully overwritten.            */
int ppnn_stencil_bwd(const float* grad_out, const float* in, float* grad_in,
                     long planes, int Hin, int Win, int axis, int scheme,
                     int accuracy, void* stream);

/* One rhs.py function on the caller's pre-padded halos (src/rhs.py:1-47,52-56):
 * u1, v1, out are [B, H, W]; ax is [B, H+2h, W], ay is [B, H, W+2h]; mu is [B].
 * adv_*: scheme/accuracy of dudx and dudy; diff_*: of d2udx2 and d2udy2; both
 * must need the same halo h (as the reference's single pad implies).           */
int ppnn_rhs2d_fwd(const float* u1, const float* v1, const float* mu,
                   const float* ax, const float* ay, float* out,
                   int B, int H, int W, int halo, int rhs_id,
                   int adv_scheme, int adv_accuracy, int diff_scheme, int diff_accuracy,
                   float dx, float dy, float dx2, float dy2, void* stream);

/* The same with u1 / v1 read in place from a strided view: rows of W floats every
 * uv_row_pitch floats, samples every uv_plane_stride floats -- the view
 * down(u)[:, :, :-1, :-1] of a [B,1,N+1,N+1] tensor that train2D.py:68-69 passes
 * (no contiguous copy).  Grids of at most 128 columns (every coarse mesh of the
 * reference); PPNN_E_SHAPE otherwise, and the caller copies.                     */
int ppnn_rhs2d_fwd_strided(const float* u1, const float* v1, const float* mu,
                           const float* ax, const float* ay, float* out,
                           int B, int H, int W, int halo, int rhs_id,
                           int adv_scheme, int adv_accuracy, int diff_scheme, int diff_accuracy,
                           float dx, float dy, float dx2, float dy2,
                           long uv_plane_stride, int uv_row_pitch, void* stream);

/* mesh_convertor.down / .up (src/utility/utils.py:24,27-31): torch
 * F.interpolate(mode='bilinear'|'bicubic', size=(Hout,Wout), align_corners).    */
int ppnn_resample_fwd(const float* in, float* out, long planes, int Hin, int Win,
                      int Hout, int Wout, int mode, int align_corners, void* stream);

/* The same on input planes that are `in_plane_stride` floats apart (each plane itself dense,
 * row-major): lets the caller pass a channel slice such as u[:, :1] of a [B,2,H,W] state, which is
 * what train2D.py:68-69 hands to mesh_convertor.down, without a contiguous copy. */
int ppnn_resample_fwd_strided(const float* in, float* out, long planes, long in_plane_stride,
                              int Hin, int Win, int Hout, int Wout, int mode, int align_corners,
                              void* stream);
/* adjoint w.r.t. `in`; grad_in [planes,Hin,Win] is fully overwritten (gather: deterministic). */
int ppnn_resample_bwd(const float* grad_out, float* grad_in, long planes, int Hin, int Win,
                      int Hout, int Wout, int mode, int align_corners, void* stream);

/* ====================== Tier B: fused coarse-grid step ===================== */

/* Geometry and physics of one YAML case (src/train2D.py:49-64).               */
typedef struct ppnn_coarse_cfg {
    int   FH, FW;            /* fine mesh incl. duplicated endpoint (finemeshsize)   */
    int   CH, CW;            /* coarse mesh incl. duplicated endpoint (coarsemeshsize) */
    int   rhs_u, rhs_v;      /* PPNN_RHS_* | PPNN_REACT_*  (rhsu / rhsv)             */
    int   adv_accuracy;      /* Upwind1 accuracy of dudx/dudy (train2D.py:55-56: 1)  */
    int   diff_accuracy;     /* Central2 accuracy of d2udx2/d2udy2 (train2D.py:53-54: 2) */
    float dt, dx, dy, dx2, dy2;
} ppnn_coarse_cfg;

/* pde_du (src/train2D.py:66-79): u [B,2,FH,FW], mu [B] -> pdeu [B,2,FH,FW].    */
int ppnn_pde_du_fwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                    float* pdeu, int B, void* stream);

/* The "iter" variant of pde_du in post_process/test2D2.py:70-91: `inner` sequential explicit sub-steps of the
 * restricted coarse state (u1 <- u1 + dt*rhsu(u1,v1); v1 <- v1 + dt*rhsv(new u1, v1) on the halos of the old v1)
 * before the increment dt*rhs is evaluated and prolonged; inner = 0 is ppnn_pde_du_fwd.  Forward only (the
 * script uses it for inference).                                                              */
int ppnn_pde_du_iter_fwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                         float* pdeu, int B, int inner, void* stream);

/* Vector-Jacobian product of pde_du: grad_u [B,2,FH,FW] (overwritten) and
 * grad_mu [B] (overwritten; may be NULL).  add_identity != 0 returns
 * grad_out + J^T grad_out, the adjoint of  u -> u + pde_du(u).                  */
int ppnn_pde_du_bwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                    const float* grad_out, float* grad_u, float* grad_mu,
                    int B, int add_identity, void* stream);

/* Backward of ppnn_step_update in one launch: grad_u = grad_out + J^T grad_out (as ppnn_pde_du_bwd with
 * add_identity), grad_mu [B] (may be NULL) and grad_nn [B,2,FH-1,FW-1] (may be NULL) = outstd * padBC_rd^T(grad_out),
 * the gradient of the network output (train2D.py:235-238).  grad_nn needs the reference mesh family (last fine
 * column on the last coarse node, FW-1 a multiple of 32); otherwise PPNN_E_SHAPE: compose it from grad_out. */
int ppnn_step_update_bwd(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                         const float* grad_out, const float* outstd,
                         float* grad_u, float* grad_mu, float* grad_nn, int B, void* stream);

/* Saved-state variants for training through unrolled rollouts (autograd over src/train2D.py:66-79,
 * 235-238).  The forward entries additionally store the restricted coarse state
 * coarse_state [B,2,CH-1,CW-1] (= mesh_convertor.down(u)[:,:,:-1,:-1] of both fields, train2D.py:67-68;
 * 18 KB per sample on the 257/49 mesh); coarse_state may be NULL (then they are the plain entries).
 * The backward entries take that state INSTEAD of u: pde_du depends on u only through it, so the
 * 528 KB state of a step need not be kept alive for the backward pass and the adjoint kernel does not
 * re-gather it from HBM.  coarse_state must be 16-byte aligned.  Same results as the plain entries. */
int ppnn_pde_du_fwd_save(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                         float* pdeu, float* coarse_state, int B, void* stream);
int ppnn_step_update_save(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                          const float* nn_out, const float* outmean, const float* outstd,
                          float* u_next, float* coarse_state, int B, void* stream);
int ppnn_pde_du_bwd_saved(const ppnn_coarse_cfg* cfg, const float* coarse_state, const float* mu,
                          const float* grad_out, float* grad_u, float* grad_mu,
                          int B, int add_identity, void* stream);
int ppnn_step_update_bwd_saved(const ppnn_coarse_cfg* cfg, const float* coarse_state, const float* mu,
                               const float* grad_out, const float* outstd,
                               float* grad_u, float* grad_mu, float* grad_nn, int B, void* stream);

/* Network inputs of one rollout step (src/train2D.py:235-237), fused with pde_du:
 *   u_norm    [B,2,FH-1,FW-1] = (u[:,:,:-1,:-1]      - inmean ) / instd
 *   pdeu_norm [B,2,FH-1,FW-1] = (pde_du(u)[:,:,:-1,:-1] - pdemean) / pdestd
 * stats are per-channel pairs {c0, c1}.  pdeu_norm may be NULL (cnn2d ignores it,
 * src/models.py:95-96); pdeu (un-normalised, full [B,2,FH,FW]) may be NULL.     */
int ppnn_step_inputs(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                     const float* inmean, const float* instd,
                     const float* pdemean, const float* pdestd,
                     float* u_norm, float* pdeu_norm, float* pdeu, int B, void* stream);

/* The explicit update of one rollout step (src/train2D.py:235-238), fused with
 * pde_du:  u_next = padBC_rd(nn_out*outstd + outmean) + u + pde_du(u, mu).
 * nn_out [B,2,FH-1,FW-1] may be NULL: u_next = u + pde_du(u)  (the PDE-only
 * rollout of post_process/test_predict_2Dp.py:116).  u_next must NOT overlap u
 * (the two field planes of a sample are processed by different CTAs).            */
int ppnn_step_update(const ppnn_coarse_cfg* cfg, const float* u, const float* mu,
                     const float* nn_out, const float* outmean, const float* outstd,
                     float* u_next, int B, void* stream);

/* Coarse-only solver increment (post_process/test_2D_csolver.py:56-66): the state
 * lives on the coarse grid, uc [B,2,CH,CW]; out = padBC_rd(dt*rhs) (add_state=0)
 * or uc + padBC_rd(dt*rhs) (add_state=1).  FH/FW of cfg are ignored.            */
int ppnn_coarse_increment(const ppnn_coarse_cfg* cfg, const float* uc, const float* mu,
                          float* out, int B, int add_state, void* stream);

/* Dataset-side precompute of the training set (src/train2D.py:88-90,109-151,
 * class `myset`), device-resident.  Sample b is the pair (u, u_next) of states one
 * step apart, both [*,2,FH,FW] fp32; traj_len = T > 0: the two arrays are views of
 * one trajectory tensor [P,T+1,2,FH,FW] (u = its base, u_next = base + one state) and
 * sample b sits at state slot b + b/T; traj_len = 0: dense arrays [B,2,FH,FW].
 * With  pdeu = pde_du(u, mu)  and  label = (u_next - u) - pdeu :
 *
 * ppnn_dataset_stats  writes partial[B][2][6] (fp64): per (sample, channel) the sums
 *   {u, u^2} over [:-1,:-1], {pdeu, pdeu^2} over the FULL plane (train2D.py:126-127
 *   takes them before the crop), {label, label^2} over [:-1,:-1]; summed over B by the
 *   caller, they give inmean/instd, pdemean/pdestd, outmean/outstd (:115-116,141-142).
 * ppnn_dataset_emit   writes u0_normd, pdeu_normd, du_normd [B,2,FH-1,FW-1]
 *   (:117,128,143); mean6 / std6 are HOST arrays {in c0,c1, pde c0,c1, out c0,c1}.
 *
 * Meshes the streamed kernel does not take (FW-1 not a multiple of 4, more than 512
 * columns) return PPNN_E_SHAPE; compose from ppnn_pde_du_fwd in that case.          */
int ppnn_dataset_stats(const ppnn_coarse_cfg* cfg, const float* u, const float* u_next, const float* mu,
                       long B, long traj_len, double* partial, void* stream);
int ppnn_dataset_emit(const ppnn_coarse_cfg* cfg, const float* u, const float* u_next, const float* mu,
                      long B, long traj_len, const float* mean6, const float* std6,
                      float* u_norm, float* pdeu_norm, float* label_norm, void* stream);

/* ===================== Tier B: fine-grid explicit solver =================== */

/* One explicit Euler step of the 2-D Burgers generator (src/operators.py:466-475)
 * on u, v [B,H,W] with mu [B]:  Upwind1(adv_accuracy) + Central2(diff_accuracy).
 * endpoint != 0: the grid carries a duplicated periodic endpoint (period H-1 /
 * W-1, operators.py:416-420); endpoint == 0: true periodic (period H / W).
 * mode: PPNN_FINE_AS_WRITTEN | PPNN_FINE_INTENDED.  Outputs must not alias inputs. */
int ppnn_fine_burgers_step(const float* u, const float* v, const float* mu,
                           float* u_out, float* v_out, int B, int H, int W,
                           int endpoint, int mode, int adv_accuracy, int diff_accuracy,
                           float dt, float dx, float dy, void* stream);

/* Random-Fourier initial condition of the generator (src/operators.py:423-445): for every sample
 * out[b,i,j] = sum_{k,l=-modes..modes} a_kl sin(k x_i + l y_j) + c_kl cos(k x_i + l y_j) on
 * x = y = linspace(0, 2*pi, n), min-max normalised per sample to [offset, offset + 1] (the reference:
 * modes = 4, offset = 0.1).  coef [B, (2*modes+1)^2, 2] holds (a, c) in the reference's (k, l) loop order
 * (the values its torch.randn_like calls draw); out [B, n, n]; minmax_scratch [B, 2] floats of device
 * scratch (the library does not allocate).  modes <= 8.                                              */
int ppnn_fourier_ic(const float* coef, float* out, float* minmax_scratch,
                    int B, int n, int modes, float offset, void* stream);

/* Generic fused explicit step  f <- f + dt*rhs(f)  for any rhs id pair on a
 * periodic fine grid (the commented generators of operators.py:494-748: RD,
 * diffusion, convection); Jacobi update (both fields from the old state).      */
int ppnn_fine_rhs_step(const float* u, const float* v, const float* mu,
                       float* u_out, float* v_out, int B, int H, int W,
                       int endpoint, int rhs_u, int rhs_v,
                       int adv_accuracy, int diff_accuracy,
                       float dt, float dx, float dy, void* stream);

/* ============ Tier B: small explicit solvers, many steps per launch ======== *
 * One CTA owns one sample; the state stays in registers / shared memory for all
 * `nsteps` steps of the launch (HBM is touched on entry and exit only).        */

/* 1-D Kuramoto-Sivashinsky generator, classic RK4 (src/operators.py:845-918; f at :883-884 is
 * rhs.ks1d, src/rhs.py:49-50, plus an optional source term):
 *   f(u) = -d2udx2(pad u)/dx2 - d4udx4(pad u)/dx4 - 0.5*dudx(pad(u*u))/dx + source
 *   u <- u + dt*(k1 + 2*k2 + 2*k3 + k4)/6
 * with diffusion1D('Central', d2_accuracy), diffusion1D('Central_4th', d4_accuracy) and
 * dudx_1D('Upwind', adv_accuracy) (switch on the sign of u*u, as operators.py:126-127 does).
 * u, u_out [B, n]; source [n] or NULL.  endpoint != 0: duplicated periodic endpoint, the pad of
 * operators.py:852-853 (neighbour (i+d) modulo n-1 on the n-point array); endpoint == 0: period n.
 * dx2, dx4 are the caller's dx**2 and dx2**2 (Python doubles in the reference), rounded to fp32.
 * n <= 8192.  u_out may alias u.                                                             */
int ppnn_ks1d_rk4(const float* u, const float* source, float* u_out, int B, int n, int endpoint,
                  int nsteps, int adv_accuracy, int d2_accuracy, int d4_accuracy,
                  float dt, float dx, float dx2, float dx4, void* stream);

/* f(u) alone (the `f` closure of operators.py:883-884): f_out [B, n], must not alias u.          */
int ppnn_ks1d_rhs(const float* u, const float* source, float* f_out, int B, int n, int endpoint,
                  int adv_accuracy, int d2_accuracy, int d4_accuracy,
                  float dx, float dx2, float dx4, void* stream);

/* 2-D Kuramoto-Sivashinsky generator, classic RK4 (src/operators.py:750-843; f at :806-809):
 *   f(u) = -c2*d2udx2(u)/dx2 - c2*d2udy2(u)/dy2 - d4udx4(u)/dx4 - d4udy4(u)/dy4
 *          - mixed_scale*d2udy2(pad(d2udx2(pad u))) - mixed_scale*d2udx2(pad(d2udy2(pad u)))
 *          - 0.5*dudx(pad(u*u))/dx - 0.5*dudy(pad(u*u))/dy
 * (reference: c2 = 0.1, mixed_scale = 1: the mixed term is NOT divided by dx2*dy2 there), Central2 /
 * Central4 / Upwind1 stencils, every pad the duplicated-endpoint wrap of operators.py:761-765.
 * u, u_out [B, H, W]; four frames of (H+2h) x (W+2h) floats (h = the widest stencil's halo) must fit in
 * shared memory ((H+2h)*(W+2h) <= 14528, e.g. 114^2 at h = 3), else PPNN_E_TOOBIG: compose the step from
 * the Tier-A operators.  u_out may alias u.                                                   */
int ppnn_ks2d_rk4(const float* u, float* u_out, int B, int H, int W, int endpoint, int nsteps,
                  int adv_accuracy, int d2_accuracy, int d4_accuracy, float c2, float mixed_scale,
                  float dt, float dx, float dy, float dx2, float dy2, float dx4, float dy4,
                  void* stream);

/* f(u) alone (the `f` closure of operators.py:806-809): f_out [B, H, W], must not alias u.      */
int ppnn_ks2d_rhs(const float* u, float* f_out, int B, int H, int W, int endpoint,
                  int adv_accuracy, int d2_accuracy, int d4_accuracy, float c2, float mixed_scale,
                  float dx, float dy, float dx2, float dy2, float dx4, float dy4, void* stream);

/* padder.__call__ (Bv/data.py:25-28): out [B, n+2] = [ghost_l | u | ghost_r] with
 * ghost = (cls == 0)*bv + (cls == 1)*(u_edge + bv*dx); cls [B,2] int32 (0 Dirichlet, 1 Neumann),
 * bv [B,2].                                                                                   */
int ppnn_heat1d_pad(const float* u, const int* cls, const float* bv, float* out,
                    int B, int n, float dx, void* stream);

/* Rollout update of the Bv workload (Bv/train.py:162, Bv/test.py:153):
 *   out = padder(u[...,1:-1] + nn + pdeu[...,1:-1], cls, bv)
 * u, pdeu, out [B, n+2] (padded states; pdeu may be NULL: the network-only model), nn [B, n].  out must not
 * alias u or pdeu.                                                                             */
int ppnn_heat1d_update(const float* u, const float* nn, const float* pdeu, const int* cls, const float* bv,
                       float* out, int B, int n, float dx, void* stream);

/* The generator loop of Bv/data.py:92-97, `nsteps` steps in one launch:
 *   up = padder(u, cls, bv);  u <- u + diffusion1D(accuracy=2)(up)/dx2*dt
 * padded_out (may be NULL) receives up = padder(u_in) BEFORE the first step (what the loop saves
 * every `saveevery` steps), sample b at padded_out + b*padded_stride, n+2 floats; u_out [B, n]
 * (may be NULL if only the padded snapshot is wanted; may alias u).                           */
int ppnn_heat1d_bc_steps(const float* u, const int* cls, const float* bv, float* u_out,
                         float* padded_out, long padded_stride, int B, int n, int nsteps,
                         float dt, float dx, float dx2, void* stream);

/* genforward(...).forward (Bv/data.py:39-51), the 1-D analogue of pde_du with ghost cells:
 *   uc = linear_down(u);  uc1 = padder_c(uc[1:-1] + diffusion1D(uc)/cdx2*dt);  out = linear_up(uc1) - u
 * u, out [B, nf] (the PADDED fine state, nf = n+2), nc coarse points (F.interpolate with
 * align_corners=True: floor(nf*scale)).  out must not alias u.                                */
int ppnn_heat1d_coarse_increment(const float* u, const int* cls, const float* bv, float* out,
                                 int B, int nf, int nc, float dt, float cdx, float cdx2,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PPNN_B200_H */

/* unidom_b200 -- C ABI of the B200-native DaXBench simulator step.
 *
 * Drop-in boundary: these entry points are what an
 * XLA-FFI / ctypes / cgo binding for the reference's
 *   SimpleMPMSimulator.step_jax   (DaXBench/daxbench/core/engine/mpm_simulator.py:61-63, 413-429)
 *   ClothSimulator.step_jax       (DaXBench/daxbench/core/engine/cloth_simulator.py:68-70, 163-180)
 * binds, at `step` granularity (one call = all `conf.steps` MPM substeps or
 * all 50 cloth substeps of one sub-action, batched over num_envs).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says host;
 *   - float = IEEE binary32, int32 indices, row-major arrays in the reference's
 *     own leaf shapes (the pytree order of MPMState / PrimitiveState / ClothState);
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*):
 *     it never synchronises, never allocates and never throws;
 *   - `workspace` is caller-owned scratch of at least *_workspace_bytes() bytes,
 *     256-byte aligned; contents are undefined between calls except where a
 *     bwd call documents that it consumes what the matching fwd call left;
 *   - return value: 0 = enqueued, <0 = UD_E_* (nothing was enqueued).
 */
#ifndef UNIDOM_B200_H
#define UNIDOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UD_MAX_PRIM 4

#define UD_OK 0
#define UD_E_INVALID (-1)   /* bad shape / parameter / null pointer   */
#define UD_E_WORKSPACE (-2) /* workspace too small or misaligned      */
#define UD_E_CUDA (-3)      /* a launch failed (cudaGetLastError)     */

#define UD_SDF_BOX 0        /* core/engine/primitives/box.py:6-18       */
#define UD_SDF_CONTAINER 1  /* core/engine/primitives/container.py:8-16 */

#define UD_P2G_ATOMIC 0        /* vector RED per touched cell (fast, order-nondeterministic)   */
#define UD_P2G_LIQUID_FAST 0x100 /* flag, OR-ed into p2g_mode: the scene holds material-0 (liquid) particles; they
                                   then skip the SVD in both passes (J = |det F1|; mpm_simulator.py:241-242 make mu 0).
                                   Without the flag liquid particles take the general SVD path: same results to
                                   rounding, only slower.  The kernels of scenes without liquid stay free of the branch. */
#define UD_P2G_DETERMINISTIC 1 /* same sorted in-CTA segment sums, combined across CTAs with 64-bit
                                  fixed-point integer REDs (associative): bit-reproducible run to run and
                                  between the forward and the adjoint's recompute pass            */

/* Scalars of the reference's per-task DefaultConf (e.g. envs/shape_elasto_plastic.py:23-54).
 * Doubles are the Python floats of the conf; the library folds derived constants in
 * double and rounds to float exactly where the reference's expressions do. */
typedef struct ud_mpm_params {
  int32_t num_envs;          /* B: batch_size (mpm_simulator.py:34)                      */
  int32_t n_particles;       /* n per env (mpm_simulator.py:154)                         */
  int32_t steps;             /* S = conf.steps: substeps per step and rows of the tables */
  int32_t res[3];            /* conf.res                                                 */
  int32_t n_grid;            /* conf.n_grid (upper wall test uses this, :312)            */
  double dt, dx, inv_dx;     /* conf.dt, conf.dx, conf.inv_dx                            */
  double p_mass, p_vol;      /* conf.p_mass, conf.p_vol                                  */
  double gravity[3];         /* conf.gravity                                             */
  int32_t n_primitive;       /* conf.n_primitive, <= UD_MAX_PRIM                         */
  int32_t sdf_kind;          /* UD_SDF_* (the reference's global set_sdf, primitives.py:26) */
  int32_t use_position_control; /* mpm_simulator.py:289                                  */
  int32_t p2g_mode;          /* UD_P2G_*                                                 */
} ud_mpm_params;

/* Float leaves of one PrimitiveState (primitives.py:9-23), batched on axis 0. */
typedef struct ud_primitive {
  float* size;          /* [B,3]   */
  float* friction;      /* [B]     */
  float* softness;      /* [B]  (the int leaf 666 as float; never differentiated) */
  float* position;      /* [B,S,3] */
  float* rotation;      /* [B,S,4] */
  float* v;             /* [B,S,3] */
  float* w;             /* [B,S,3] */
  float* action_buffer; /* [B,6]   */
  float* action_scale;  /* [B,6]   */
} ud_primitive;

/* Float leaves of MPMState (mpm_simulator.py:13-24), batched on axis 0.
 * cur_step / key are pass-through in the step and stay on the host side. */
typedef struct ud_mpm_state {
  float* x;        /* [B,n,3]   */
  float* v;        /* [B,n,3]   */
  float* C;        /* [B,n,3,3] */
  float* F;        /* [B,n,3,3] */
  float* J;        /* [B,n]     */
  float* friction; /* [B] (leaf shape (B,1)) */
  float* mu;       /* [B]       */
  float* lamda;    /* [B]       */
  ud_primitive prim[UD_MAX_PRIM];
} ud_mpm_state;

const char* ud_version(void);
/* Name of the last failing check of the calling thread (for error messages). */
const char* ud_last_error(void);

/* Bytes of workspace for one fwd call / for a fwd(save=1)+bwd pair. */
size_t ud_mpm_fwd_workspace_bytes(const ud_mpm_params* p);
size_t ud_mpm_bwd_workspace_bytes(const ud_mpm_params* p);

/* Replaces vmap(jit(step)) (mpm_simulator.py:61-63,413-429): norm_grad_state fwd
 * (nan_to_num), action clip, set_action, S x substep, copy_frame.
 * material [n] int32 and h [n] float are SimpleMPMSimulator.material / .h
 * (mpm_simulator.py:112-122), shared by all envs.  action is [B, 6*n_primitive].
 * `out` must not alias `in`. */
int ud_mpm_step_fwd(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material,
                    const float* h, const float* action, ud_mpm_state* out, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Reverse-mode of the same step (replaces substep_wrapper's VJP :332-363, copy_frame,
 * set_action, action clip, norm_grad_state/norm_grad bwd :389-408).  Recomputes the S
 * substeps from `in` (checkpoint = the step input) inside `workspace`, then reverses.
 * `gout` holds the cotangents of the step output leaves (carry + ys already summed),
 * `gin` receives the cotangents of the input leaves, `gaction` [B, 6*n_primitive].
 * Null pointers inside gout mean "zero cotangent"; null pointers inside gin mean
 * "not wanted". */
int ud_mpm_step_bwd(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material,
                    const float* h, const float* action, const ud_mpm_state* gout,
                    ud_mpm_state* gin, float* gaction, void* workspace, size_t workspace_bytes,
                    void* stream);

/* The same adjoint with K-spaced substep checkpoints INSIDE the step (north_star (e), SURVEY section 5 "save state
 * every K substeps"; replaces the store-all residuals of the lax.scan at mpm_simulator.py:425 for long steps such as
 * whip_rope's S = 70).  A checkpoint pass re-runs the forward in place from `in` and keeps the start state of every
 * `window`-th substep (96 + 48 B per particle each); the windows are then recomputed one at a time, last first, into
 * `window` state/SVD/grid slots and reversed.  Peak workspace is (ceil(S/window) * 144 + window * 192) B per particle +
 * window * 36 B per cell instead of S * 192 B per particle + S * 36 B per cell, for one extra forward sweep.
 * window >= S is ud_mpm_step_bwd itself.  Same results as ud_mpm_step_bwd (identical arithmetic; the fp32 REDs of the
 * scatter kernels are order-dependent unless p2g_mode is UD_P2G_DETERMINISTIC). */
size_t ud_mpm_bwd_windowed_workspace_bytes(const ud_mpm_params* p, int32_t window);
int ud_mpm_step_bwd_windowed(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material,
                             const float* h, const float* action, const ud_mpm_state* gout,
                             ud_mpm_state* gin, float* gaction, int32_t window, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Taped pair: the same step, but the forward keeps every substep's start state, both grids, the
 * SVD factors and the active-cell lists in `tape` (ud_mpm_tape_bytes(p) bytes, 256-byte aligned,
 * caller-owned), and ud_mpm_step_bwd_taped reverses from it without recomputing the S substeps.
 * This is what jax.grad of the reference does implicitly (XLA keeps the residuals of every substep
 * of the lax.scan, mpm_simulator.py:425); ud_mpm_step_bwd is the low-memory alternative.  The tape
 * must stay untouched between the two calls; it is only read by the reverse pass (its tail holds
 * the reverse pass's scratch), so the same tape can be reversed more than once.  Same results as
 * ud_mpm_step_fwd / ud_mpm_step_bwd: bit-identical with UD_P2G_DETERMINISTIC. */
size_t ud_mpm_tape_bytes(const ud_mpm_params* p);
int ud_mpm_step_fwd_taped(const ud_mpm_params* p, const ud_mpm_state* in, const int32_t* material,
                          const float* h, const float* action, ud_mpm_state* out, void* tape,
                          size_t tape_bytes, void* stream);
int ud_mpm_step_bwd_taped(const ud_mpm_params* p, const ud_mpm_state* in, const float* action,
                          const ud_mpm_state* gout, ud_mpm_state* gin, float* gaction, void* tape,
                          size_t tape_bytes, void* stream);

/* The per-frame binning the step uses, exposed for the bit-exactness check:
 * base = int32(x*inv_dx - 0.5) (mpm_simulator.py:233, no FMA contraction),
 * key  = 4x4x4-block-major cell key, perm = stable argsort(key) per env.
 * out_base [B,n,3], out_key [B,n], out_perm [B,n] (int32). */
int ud_mpm_sort_bins(const ud_mpm_params* p, const float* x, int32_t* out_base, int32_t* out_key,
                     int32_t* out_perm, void* workspace, size_t workspace_bytes, void* stream);

/* Number of keys per env used by ud_mpm_sort_bins (for sizing host-side checks). */
int32_t ud_mpm_num_keys(const ud_mpm_params* p);

/* ------------------------------------------------------------------------------------------------
 * Mass-spring cloth (core/engine/cloth_simulator.py).  One call = robot_step (:163-180): the 8-vector
 * sub-action is scaled, then `substeps` (50) x step (:257-337) run inside one launch.
 * Topology is passed as two tables built on the host from the cloth mask (:48-66):
 *   nbr [P,8] int32 : node index of the k-th link neighbour (links order of :48), or -1 when the
 *                     link carries no force (neighbour outside the mask, or zero rest length)
 *   L0  [P,8] float : rest length cell_size*|link| clipped to >= 1e-12 (:61-63)
 * One thread per node.  n_nodes <= 1024: one CTA per environment; up to 8192 (fold_tshirt: 3 573):
 * one thread-block cluster (<= 8 CTAs, distributed shared memory) per environment.                 */
typedef struct ud_cloth_params {
  int32_t num_envs;   /* B                                       */
  int32_t n_nodes;    /* P = cloth_mask.sum()                    */
  int32_t N;          /* conf.N (board size; informational)      */
  int32_t substeps;   /* 50 in robot_step (:176)                 */
  double dt, gravity, damping, max_v, small_num, cell_size;
  double mask_sum;    /* divisor of norm_grad's backward (:192)  */
  int32_t stiffness_is_float; /* 0: int leaf, no gradient (conf.stiffness=900); 1: float (GenDOM, apg_para) */
} ud_cloth_params;

/* Float leaves of ClothState (cloth_simulator.py:13-23), batched on axis 0. */
typedef struct ud_cloth_state {
  float* x;          /* [B,P,3] */
  float* v;          /* [B,P,3] */
  float* primitive0; /* [B,4] xyz + radius */
  float* primitive1; /* [B,4] */
  float* action0;    /* [B,4] dxyz + suction */
  float* action1;    /* [B,4] */
  float* stiffness;  /* [B] (as float) */
  float* mu;         /* [B] */
} ud_cloth_state;

size_t ud_cloth_workspace_bytes(const ud_cloth_params* p);
/* Replaces vmap(jit(robot_step_wrapper)) forward; action is [B,8]. */
int ud_cloth_step_fwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                      const float* action, ud_cloth_state* out, void* workspace, size_t workspace_bytes,
                      void* stream);
/* Reverse mode of the same call (replaces the two custom_vjp levels :107-145, :228-255 and the 8
 * norm_grad re-normalisations per substep): recomputes the substeps from `in`, then reverses. */
int ud_cloth_step_bwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                      const float* action, const ud_cloth_state* gout, ud_cloth_state* gin, float* gaction,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Fused env step (replaces jax.lax.scan(self.simulator.step_jax, state, actions) of envs/basic/cloth_env.py:211):
 * `actions` is [T,B,8] (the 40 pick-and-place sub-actions of get_pnp_actions, :136-173); the forward runs all
 * T*substeps substeps in ONE launch with the node state in registers and, when `ckpt` is non-null, leaves the state at
 * the start of every sub-action there (caller-owned, ud_cloth_multi_ckpt_bytes, 256-byte aligned, must stay untouched
 * until the matching bwd call).  The adjoint recomputes and reverses sub-action by sub-action from those checkpoints;
 * `gout` = cotangents of the FINAL state, `gin` (all of x, v, primitive0/1, stiffness, mu required) = cotangents of the
 * input state, `gactions` [T,B,8]. */
size_t ud_cloth_multi_ckpt_bytes(const ud_cloth_params* p, int32_t T);
size_t ud_cloth_multi_workspace_bytes(const ud_cloth_params* p, int32_t T);
int ud_cloth_multi_step_fwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                            const float* actions, int32_t T, ud_cloth_state* out, void* ckpt, size_t ckpt_bytes,
                            void* stream);
int ud_cloth_multi_step_bwd(const ud_cloth_params* p, const ud_cloth_state* in, const int32_t* nbr, const float* L0,
                            const float* actions, int32_t T, const void* ckpt, const ud_cloth_state* gout,
                            ud_cloth_state* gin, float* gactions, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ---- reward kernels either side of the step inside the differentiated rollout -----------------
 * calc_chamfer(x, y) (DaXBench/daxbench/core/utils/util.py:138-153, metric l2, direction bi):
 *   out[b] = mean_q min_p d(x[b,p], y[q]) + mean_p min_q d(x[b,p], y[q]),  d(a,b) = sqrt(mean_c (a_c-b_c)^2)
 * x [B,P,3] per env, y [Q,3] the goal shared by all envs (cloth_env.py:205,217), out [B].
 * `residuals` (ud_chamfer_residual_bytes, 256-byte aligned, caller-owned) receives per point the minimum squared
 * distance and the number of exact ties; ud_chamfer_bwd needs it unchanged.  The adjoint splits the cotangent evenly
 * over ties like jnp.min's VJP and returns d out / d x only (the goal is a constant of the task). */
size_t ud_chamfer_residual_bytes(int32_t B, int32_t P, int32_t Q);
int ud_chamfer_fwd(const float* x, const float* y, int32_t B, int32_t P, int32_t Q, float* out, void* residuals,
                   size_t residual_bytes, void* stream);
int ud_chamfer_bwd(const float* x, const float* y, int32_t B, int32_t P, int32_t Q, const float* gout,
                   const void* residuals, size_t residual_bytes, float* gx, void* stream);
/* calc_l2(x, y) (util.py:156-159): out[b] = mean_p sqrt(mean_c (x[b,p,c]-y[p,c])^2); y [P,3] (mpm_env.py:91-94). */
int ud_l2_fwd(const float* x, const float* y, int32_t B, int32_t P, float* out, void* stream);
int ud_l2_bwd(const float* x, const float* y, int32_t B, int32_t P, const float* gout, float* gx, void* stream);

/* APG update on the flat fp32 policy-gradient buffer (DaXBench/daxbench/algorithms/apg/apg.py:233-240, 260-267).
 * ud_apg_scrub_clip: grad <- nan_to_num(grad); norm = ||grad||; grad <- norm < max ? grad : (grad / norm) * max, in
 *   place, per rank, BEFORE the mean over ranks; *sumsq (device, 1 float) receives norm^2 (the `grad_norm` metric).
 * The caller all-reduces (SUM) grad over the ranks, then
 * ud_adam_step: g = grad / world_size; optax.adam(lr, b1, b2, eps) step number t >= 1 on params with moments m, v. */
int ud_apg_scrub_clip(float* grad, int64_t n, float max_grad_norm, float* sumsq, void* stream);
int ud_adam_step(float* params, const float* grad, float* m, float* v, int64_t n, int32_t world_size, double lr,
                 double b1, double b2, double eps, int32_t t, void* stream);

/* The same update as ONE kernel per rank, the collective included ("fused clip + all-reduce + Adam"): scrub,
 * per-rank global-norm clip, mean over the ranks, optax.adam -- identical arithmetic and order of operations to
 * ud_apg_scrub_clip -> all-reduce(sum) -> ud_adam_step(world_size).  The ranks exchange the clipped gradient through
 * peer-mapped device memory (NVLink): peer_stage[r] / peer_flags[r] are DEVICE arrays of `world` device pointers, one
 * per rank, to that rank's staging buffer (float[4 * n]: two staging slots + two slots of the reduced gradient) and flag
 * array (int32[64], zeroed once before the first call), valid in the calling process (cudaIpc / cuMem fabric handles / torch symmetric memory -- obtaining the mapping
 * is the caller's business).  t = 1, 2, ... must advance by one per call on every rank.  scratch: 8 floats.
 * Every rank sums the staged gradients in rank order, so replicas stay bit-identical.  From 4 ranks up a rank sums only
 * its slice and stores the mean into every peer's reduced slot (reduce-scatter + broadcast over peer memory, a second
 * flag barrier): 2 n instead of world * n elements over NVLink per rank, same bits.  world == 1 degenerates to
 * scrub + clip + Adam in one launch. */
int ud_apg_fused_update(float* params, const float* grad, float* m, float* v, int64_t n, float max_grad_norm,
                        double lr, double b1, double b2, double eps, int32_t t, int32_t rank, int32_t world,
                        const uint64_t* peer_stage, const uint64_t* peer_flags, float* scratch, void* stream);

/* ---- instrumentation (bench.py): launch counting and per-kernel-class CUDA-event timing ------
 * ud_launch_count: kernels + memsets enqueued by this library since the last reset (host counter).
 * ud_timing_enable(1): subsequent calls bracket every kernel class with cudaEvents on the call's
 * stream; ud_timing_collect synchronises those events and returns, per class, the summed device
 * milliseconds and number of launches since the last collect.  Class names: ud_timing_class_name. */
uint64_t ud_launch_count(int reset);
/* A/B switches for the measurements DESIGN.md quotes (process-global, not thread-safe, default = the fast path):
 *   "svd_warm"  1: warm-start the per-particle Jacobi SVD from the previous substep's V (default) / 0: cold start
 *   "sort"      1: per-frame binning by grid block (default) / 0: particles stay in input order (ud_mpm_sort_bins unaffected)
 *   "stage"     1: P2G staged in shared memory, one vector RED per (cell segment, node) (default) /
 *               0: 27 vector REDs per particle straight to the grid in HBM
 *   "mark"      2: P2G marks the 4x4x4 grid blocks it scatters into from the corner nodes of every cell segment
 *               (default) / 1: from all 27 nodes / 0: no marks (timing experiments only: the grid update then
 *               visits nothing and the results are wrong)
 *   "cloth_cta_nodes"  cloth nodes per CTA (default and maximum 1024, multiple of 32); an env with more nodes runs
 *               on a thread-block cluster of ceil(n_nodes / value) <= 8 CTAs (tests force small cloths onto it)
 * Returns the previous value, or -1 for an unknown name. */
int ud_tuning_set(const char* name, int value);
void ud_timing_enable(int on);
int ud_timing_num_classes(void);
const char* ud_timing_class_name(int cls);
int ud_timing_collect(double* ms_by_class, int64_t* launches_by_class, int n_classes);

#ifdef __cplusplus
}
#endif
#endif /* UNIDOM_B200_H */

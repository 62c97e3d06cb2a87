/* pstl.h — C ABI of libpstl_b200.so (sm_100a).
 *
 * The reference (mengyuest/pSTL-diffusion-policy) is pure Python/PyTorch and has no FFI or
 * operator registry; its boundary for the hot path is the Python API.  The
 * entry points below are what a maintainer binds with ctypes from that API;
 * each one names the reference code it replaces (paths relative to the reference checkout).
 *
 * Conventions: every function returns 0 on success, <0 on error (pstl_last_error() gives the
 * thread-local message).  All tensor arguments are caller-owned DEVICE pointers (fp32 unless
 * noted, row-major, contiguous); the library never frees or keeps them past the call.  Every
 * launch goes to the given stream (cudaStream_t passed as void*); there are no hidden
 * synchronisations and no allocations in steady state except inside handle creation.
 * Handles are immutable after creation.
 */
#ifndef PSTL_H_
#define PSTL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSTL_OK 0
#define PSTL_ERR_ARG -1
#define PSTL_ERR_CUDA -2
#define PSTL_ERR_UNSUPPORTED -3

typedef void* pstl_stream_t; /* cudaStream_t */
typedef struct pstl_program* pstl_program_t;
typedef struct pstl_denoiser* pstl_denoiser_t;

const char* pstl_last_error(void);
/* library / device facts: returns sm count, writes compute capability major*10+minor */
int pstl_device_info(int* sm_count, int* cc);
int pstl_version(void);
/* kernels launched by this library since load (bench.py reports it as gpu_launches) */
unsigned long long pstl_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * STL formula programs (replaces the recursive node evaluation of stl_d_lib.py:70-203).
 * A formula tree is flattened by the host into postfix ops; every op consumes the top
 * trace(s) of a stack of (T)-long traces per trajectory and pushes one trace.
 * ------------------------------------------------------------------------------------- */
enum pstl_opcode {
  PSTL_OP_SIGNAL = 0,      /* a0 = signal id p: push sig[:,p,:]                 (AP leaf, stl_d_lib.py:70-82) */
  PSTL_OP_PRED = 1,        /* fused driving predicate, see pstl_pred_* below     (nusc_train.py:98-134)      */
  PSTL_OP_NEG = 2,         /* Not                                                (stl_d_lib.py:126-131)      */
  PSTL_OP_SMIN2 = 3,       /* And: soft-min of the two top traces                (stl_d_lib.py:87-95)        */
  PSTL_OP_SMAX2 = 4,       /* Or  (Imply = NEG on lhs + SMAX2)                   (stl_d_lib.py:114-142)      */
  PSTL_OP_SMIN_K = 5,      /* a0 = k: ListAnd over the k top traces              (stl_d_lib.py:97-112)       */
  PSTL_OP_WIN_SMIN = 6,    /* a0 = ts, a1 = te: Always, window [t+ts,t+te)∩[0,T) (stl_d_lib.py:157-169)      */
  PSTL_OP_WIN_SMAX = 7,    /* a0 = ts, a1 = te: Eventually / Once                (stl_d_lib.py:144-155,171-180) */
  PSTL_OP_PREFIX_SMIN = 8, /* -logcumsumexp(-x*tau)/tau (always soft)            (stl_d_lib.py:189)          */
  PSTL_OP_SUFFIX_SMAX = 9  /* flip(logcumsumexp(flip(x)*tau))/tau (always soft)  (stl_d_lib.py:191)          */
};

typedef struct {
  int32_t op;
  int32_t a0;
  int32_t a1;
} pstl_op;

/* PSTL_OP_PRED encoding: value[t] = (ss*base[sid][t] + sp*stlp[pid]) / den
 *   a0 = sid | (ss_negative << 8)           base signals: see enum pstl_base_signal
 *   a1 = pid | (sp_negative << 8) | (den << 16)   den: see enum pstl_denominator          */
enum pstl_base_signal {
  PSTL_SIG_V = 0,        /* ego speed                                                   */
  PSTL_SIG_D_CURR = 1,   /* signed distance / heading error to the current, left, right */
  PSTL_SIG_TH_CURR = 2,  /* lane (nusc_api.py:685-739)                                  */
  PSTL_SIG_D_LEFT = 3,
  PSTL_SIG_TH_LEFT = 4,
  PSTL_SIG_D_RIGHT = 5,
  PSTL_SIG_TH_RIGHT = 6,
  PSTL_SIG_NEI = 7,      /* min neighbour clearance (utils.py:465-526, nusc_train.py:142-148) */
  PSTL_N_BASE_SIGNALS = 8
};
enum pstl_denominator {
  PSTL_DEN_ONE = 0,
  PSTL_DEN_THMAX = 1,    /* stlp[5]                                       (nusc_train.py:132-134) */
  PSTL_DEN_VFACTOR = 2,  /* clip(vmax-vmin, .3)        --norm_stl         (nusc_train.py:88-91)   */
  PSTL_DEN_DFACTOR = 3,  /* clip(5*(dmax-dmin), .3)                                               */
  PSTL_DEN_SFACTOR = 4   /* clip(dsafe, .3)                                                       */
};

/* n_signals: number of generic signals the program reads (0 for pure PRED programs).
 * need_t: how many leading time steps of the top-level trace the caller will read
 *         (T for the node API, 1 for scoring which keeps only [:,0], nusc_train.py:321). */
int pstl_program_create(const pstl_op* postfix, int n_ops, int n_signals, int T, int need_t, pstl_program_t* out);
int pstl_program_destroy(pstl_program_t prog);
/* floats of scratch per trajectory the program needs (values only / values + adjoints) */
int pstl_program_tape_floats(pstl_program_t prog, int with_grad);

/* Generic node evaluation: node(x, tau, d) -> (N,T)   (stl_d_lib.py node.__call__).
 * sig (N,P,T); out_trace (N,need_t) or NULL; out_t0 (N) or NULL.
 * workspace: device scratch of >= pstl_stl_workspace_bytes(...) bytes, or NULL when that is 0. */
size_t pstl_stl_workspace_bytes(pstl_program_t prog, int N, int with_grad);
int pstl_stl_eval_signals(pstl_program_t prog, const float* sig, int N, int P, int T, float tau, int hard,
                          float* out_trace, float* out_t0, void* workspace, pstl_stream_t stream);
/* Reverse mode of the above: grad_trace (N,need_t) -> grad_sig (N,P,T) (overwritten). */
int pstl_stl_eval_signals_bwd(pstl_program_t prog, const float* sig, const float* grad_trace, int N, int P, int T,
                              float tau, int hard, float* grad_sig, void* workspace, pstl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused rollout + predicates + STL scoring (+ best-of-K), replacing
 *   generate_trajs (nusc_train.py:39-49), prep_stl_cache (:74-93), compute_stl_dense (:318-345),
 *   get_stl_scores (:150-151) and the best-of-K block (:992-1013).
 * Scene tensors are indexed, not replicated: row n uses scene n / rows_per_scene
 * (rows_per_scene = 1 reproduces the reference's dense per-row layout).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  const float* neighbors; /* (n_scenes, Knei, T, 7) = [valid,x,y,th,v,L,W]                */
  const float* lanes[3];  /* curr, left, right: (n_scenes, nseg, 3) = [x,y,th]            */
  int n_scenes, Knei, nseg, T;
  int rows_per_scene;
} pstl_scene_view;

typedef struct {
  float dt, tau;
  float ego_L, ego_W;
  float w_scale, a_scale; /* controls = in * scale (1 for physical controls, (w_max,a_max) for mu) */
  int clip_controls;      /* clip to +-scale after scaling (normalize_diff, nusc_train.py:647-655) */
  int clip_dist;          /* lane flag word: bit 0 --clip_dist, lane distance clipped to +-5 (nusc_api.py:732-733);
                             bit 1 --inline, end-cap distances before / after the polyline (nusc_api.py:716-724)   */
  int hard;
} pstl_spec_params;

/* progs[3]: programs of the three formulas [curr,left,right] (need_t = 1, PRED leaves);
 * mode (N) float in {0,1,2,3}: formula selector (3 = outlier, score 1.0);
 * state0 (N,4); controls (C,N,T,2) candidate-major, or NULL when ego_traj is given;
 * ego_traj (N,T,4+) pre-rolled states with row stride ego_stride floats, or NULL;
 * stlp (N,6).  Outputs (each may be NULL): scores_all (C,N), best_score (N), best_idx (N, int32),
 * best_controls (N,T,2) physical (scaled/clipped) controls of the arg-max candidate (first max wins),
 * traj_out (N,T+1,4) rollout of the best candidate. */
size_t pstl_score_workspace_bytes(pstl_program_t const* progs, int N, int T, int with_grad);
int pstl_score_fused(pstl_program_t const* progs, const pstl_scene_view* scenes, const pstl_spec_params* sp,
                     const float* mode, const float* state0, const float* controls, int C,
                     const float* ego_traj, int ego_stride, const float* stlp, int N,
                     float* scores_all, float* best_score, int32_t* best_idx, float* best_controls,
                     float* traj_out, void* workspace, pstl_stream_t stream);

/* Reverse mode for guidance / training (nusc_train.py:600-623): gradient of
 *   sum_n grad_score[n] * score[n]   w.r.t. controls_in (N,T,2) (pre-scale), or w.r.t. ego_traj.
 * grad_controls (N,T,2) or NULL; grad_ego (N,T,4) or NULL; scores (N) optional output. */
int pstl_score_fused_bwd(pstl_program_t const* progs, const pstl_scene_view* scenes, const pstl_spec_params* sp,
                         const float* mode, const float* state0, const float* controls,
                         const float* ego_traj, int ego_stride, const float* stlp, int N,
                         const float* grad_score, float* scores, float* grad_controls, float* grad_ego,
                         void* workspace, pstl_stream_t stream);

/* One guidance iteration on mu (N,T,2) in place (nusc_train.py:599-627, niters handled by the
 * caller's loop):  loss = mean(relu(thres-score)*valid)/clip(mean(valid),1e-2); Adam step with
 * state (m,v) (N,T,2) zero-initialised by the caller at iter 0; iter >= 1 additionally applies
 * mu = mu_anchor + clip(|mu - mu_anchor|, -beta_t, beta_t) as upstream does (see DESIGN.md on the
 * aliasing quirk that makes iter 0 a plain Adam step).  inv_norm = 1/(N_total*clip(mean(valid),1e-2)). */
int pstl_guidance_step(pstl_program_t const* progs, const pstl_scene_view* scenes, const pstl_spec_params* sp,
                       const float* mode, const float* state0, const float* stlp, const float* valid, int N,
                       float thres, float inv_norm, const float* inv_norm_dev /* device word overriding inv_norm, or NULL */,
                       float lr, float beta_t, int iter,
                       float* mu, float* adam_m, float* adam_v, float* mu_anchor,
                       void* workspace, pstl_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Denoiser / sampler / RefineNet, replacing Net.forward (nusc_model.py:97-180),
 * diffusion_rollout + normalize_diff (nusc_train.py:557-655) and Net.rect_forward
 * (nusc_model.py:182-235).
 * ------------------------------------------------------------------------------------- */
#define PSTL_MAX_STEPS 1024   /* most diffusion steps a sampler call takes (workspace's per-step table) */
#define PSTL_PRECISION_FP32 0 /* SIMT fp32, 1e-5 parity mode                       */
#define PSTL_PRECISION_BF16 1 /* tcgen05 bf16 operands, fp32 accumulate (2e-2)     */
#define PSTL_PRECISION_F16 3 /* the bf16 engine on fp16 operands: 11 mantissa bits instead of 8 at the same tensor-core rate
                              * (measured 8x closer to the fp32 chain); |weights|, |activations| < 65,504 (saturating)  */
#define PSTL_PRECISION_F16X3 2 /* tcgen05, every operand as two fp16 pieces (22 mantissa bits) and every product as three
                                * MMAs (hi.hi + lo.hi + hi.lo), fp32 accumulate: the accuracy of fp32 arithmetic (inside the
                                * 1e-5 bound of PSTL_PRECISION_FP32) at tensor-core speed; |weights|, |activations| < 65,504.
                                * The sampler and the RefineNet head use it; the training entry points of such a handle
                                * run the fp32 SIMT kernels.                                                      */

typedef struct {
  /* device pointers to state_dict tensors (row-major (out,in) like nn.Linear) */
  const float *p0_w, *p0_b, *p2_w, *p2_b, *p4_w, *p4_b; /* policy_net 303->H->H->2T  */
  const float *m0_w, *m0_b, *m2_w, *m2_b, *m4_w, *m4_b; /* merge_net  2T->32->32->2T (NULL if absent) */
  const float *r0_w, *r0_b, *r2_w, *r2_b, *r4_w, *r4_b; /* rect_net   271->H->H->2T  (NULL if absent) */
  int hidden, rect_hidden, merge_hidden, feat_dim, time_dim, T;
} pstl_weights;

/* PSTL_ERR_UNSUPPORTED when 2*T+7 exceeds the packed input row (T <= 20). */
int pstl_denoiser_create(const pstl_weights* w, int precision, pstl_denoiser_t* out);
int pstl_denoiser_destroy(pstl_denoiser_t d);
/* The caller updated the weight tensors IN PLACE (an optimiser step on the same storage): re-derive the handle's own
 * copies (hoisted first-layer blocks, bf16 operand images) on `stream`.  No allocation, no synchronisation. */
int pstl_denoiser_refresh(pstl_denoiser_t d, pstl_stream_t stream);

/* Which tcgen05 engine PSTL_PRECISION_BF16 launches: 0 = automatic (default: the CTA-pair engine, two 256-row tiles
 * in flight per SM pair, once the batch exceeds one wave of 128-row tiles; the one-SM engine below that), 1 = one-SM
 * engine, 2 = CTA-pair engine whenever its tile fits rows_per_scene.  Both compute the same function with the same
 * noise stream; the reference has one code path (nusc_train.py:557-645), this only picks the kernel. */
int pstl_denoiser_set_engine(pstl_denoiser_t d, int engine);

/* Philox offset word in DEVICE memory (or NULL to clear): the sampler adds *device_counter to the
 * `offset` argument of pstl_denoiser_sample when it draws noise.  A CUDA graph that captured the
 * sampler bumps this word in-graph so every replay draws fresh normals (upstream: randn_like per
 * step, nusc_train.py:627). */
int pstl_denoiser_set_noise_counter(pstl_denoiser_t d, const uint64_t* device_counter);

typedef struct {
  const float* valid;    /* (N) lane validity mask of the guidance loss                       */
  const float* state0;   /* (N,4)                                                             */
  pstl_program_t const* progs;
  const pstl_scene_view* scenes;
  const pstl_spec_params* sp; /* w_scale/a_scale = (w_max,a_max), clip_controls = 0          */
  int before;            /* guide reverse steps i <= before (when step_mask is NULL)          */
  const unsigned char* step_mask; /* HOST pointer, `steps` entries, or NULL: step_mask[i] != 0 guides reverse
                            step i (--guidance_sets / --guidance_freq / --guidance_reverse, nusc_train.py:589-598) */
  int niters;
  float lr, thres;
  float inv_norm;        /* 1/(N_total*clip(mean(valid),1e-2)); N_total spans all shards if the
                            caller wants single-batch semantics across GPUs                   */
  const float* inv_norm_dev; /* DEVICE word holding inv_norm, or NULL: read by the kernels at run time, so a CUDA
                            graph that captured the sampler follows the batch's own lane-validity mean           */
} pstl_guidance_cfg;

/* Scratch the sampler needs for N chains (activations, Adam state, STL tape). */
size_t pstl_denoiser_workspace_bytes(pstl_denoiser_t d, int N, int n_scenes, const pstl_guidance_cfg* g);

/* The whole reverse loop i = steps-1 .. 1 (t == i), fused:
 *   eps = policy_net([feat, x, temb(t), hl, stlp]) + x ; mu = (x-(1-a)/sqrt(1-abar)*eps)/sqrt(a);
 *   [guidance]; x = mu + sqrt(beta)*z.
 * scene_feat (n_scenes,feat_dim); rows_per_scene as in pstl_scene_view; hl (N); stlp (N,6);
 * sched: HOST pointer, (3,steps) = beta, alpha, alpha_hat (get_diffusion_coeffs, nusc_train.py:528-537);
 * temb (steps,time_dim) device copy of the sinusoid table (pos_encoding, nusc_model.py:48-53);
 * x_init (N,2T): x_T, or NULL to draw it from the same Philox stream (step word `steps`);
 * noise (steps-2,N,2T) injected z for i = steps-1..2, or NULL to draw
 * Philox normals (seed, offset); keep_last_k iterates are written, scaled by (w_max,a_max) and
 * clipped when clip != 0, to iterates_out (keep_last_k,N,T,2) in chronological order (last = x_0). */
int pstl_denoiser_sample(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                         const float* hl, const float* stlp, int N, const float* sched, const float* temb,
                         int steps, const float* x_init, const float* noise, uint64_t seed, uint64_t offset,
                         float w_max, float a_max, int clip, int keep_last_k, const pstl_guidance_cfg* guidance,
                         float* iterates_out, float* x_final, void* workspace, pstl_stream_t stream);

/* One eps evaluation (Net.forward with prev_feature, nusc_model.py:118-162): eps_out (N,2T). */
int pstl_denoiser_eps(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                      const float* hl, const float* stlp, const float* x, int N, const float* temb_row,
                      float* eps_out, void* workspace, pstl_stream_t stream);

/* RefineNet (nusc_model.py:182-235; diverse_loss, fuse=add, interval):
 * u0 (N,T,2) physical controls, scores (N); out (N,T,2).  group = n_randoms/n_shards samples of
 * one (scene, mode) are max-pooled; flat row index n = (scene*n_randoms + r)*3 + m. */
int pstl_refine(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                const float* hl, const float* stlp, const float* u0, const float* scores, int N,
                int n_randoms, int n_shards, float w_max, float a_max, int clip_rect,
                float* out, void* workspace, pstl_stream_t stream);

/* Standalone rollout (generate_trajs, nusc_train.py:39-49) and its adjoint.
 * state0 (N,4), controls (N,T,2) -> traj (N,T+1,4).  Backward: grad_traj (N,T+1,4) ->
 * grad_state0 (N,4) (may be NULL), grad_controls (N,T,2). */
int pstl_rollout(const float* state0, const float* controls, int N, int T, float dt, float* traj,
                 pstl_stream_t stream);
int pstl_rollout_bwd(const float* traj, const float* grad_traj, int N, int T, float dt, float* grad_state0,
                     float* grad_controls, pstl_stream_t stream);

/* Predicate signals of prep_stl_cache (nusc_train.py:74-93) for custom formulas:
 * ego (N,T,>=3) with row stride ego_stride -> sig (N,7,T) = [d_curr, th_curr, d_left, th_left,
 * d_right, th_right, min_nei_d]; part (N,12,T) (may be NULL) = per lane (dd/dx, dd/dy, dth/dtheta),
 * then d min_nei_d/d(x, y, theta): what autograd needs to chain into ego. */
int pstl_predicates(const pstl_scene_view* scenes, float ego_L, float ego_W, int clip_dist /* lane flag word */, const float* ego,
                    int ego_stride, int N, float* sig, float* part, pstl_stream_t stream);

/* Car-to-car anchor distances for any (--refined_nL, --refined_nW) grid, and the extra signals --collision_loss adds to
 * prep_stl_cache (dist_between_two_cars with full=True, utils.py:465-526; nusc_train.py:81-83, 142-148):
 * min_dist (N,Knei,T) = min over the (nL*nW)^2 anchor pairs of the centre distance, rad_sum (N,Knei,T) = r_ego + r_nei;
 * part (N,Knei,T,3) (may be NULL) = d min_dist / d (ego x, y, theta).  The callers compose
 *   car_dist = min_dist - rad_sum,  min_nei_d = min_k(clip(car_dist,-5,20)*valid + (1-valid)*100),
 *   min_centroid_d = min_dist*valid + (1-valid)*100   exactly as upstream.  nL*nW <= 64. */
int pstl_car_distances(const pstl_scene_view* scenes, float ego_L, float ego_W, int nL, int nW, const float* ego,
                       int ego_stride, int N, float* min_dist, float* rad_sum, float* part, pstl_stream_t stream);

/* Plain fused linear layer used by the scene encoders (nusc_model.py:82-91):
 * y (M,Nout) = act(x (M,K) @ w(Nout,K)^T + b); act: 0 none, 1 relu. */
int pstl_linear(const float* x, const float* w, const float* b, int M, int K, int Nout, int act, float* y,
                pstl_stream_t stream);

/* One iteration of the trajectory optimisation that produces the training targets (nusc_train.py:287-316, 1303-1325):
 *   loss = sum_n relu(thres - score_n) valid_n * inv_norm + reg * (mean relu(w^2 - w_max^2) + mean relu(a^2 - a_max^2))
 *   params <- Adam_step(params, d loss / d params)        (torch.optim.Adam defaults, bias correction of step iter+1)
 * params (N,T,2) physical controls (sp->w_scale = a_scale = 1, clip_controls = 0), updated in place; adam_m / adam_v (N,T,2)
 * zeroed by the caller before iter 0; scores (N, may be NULL) receives the robustness of the params BEFORE the step.
 * workspace: N*T*2 floats + pstl_score_workspace_bytes(progs, N, T, 1). */
int pstl_trajopt_step(pstl_program_t const* progs, const pstl_scene_view* scenes, const pstl_spec_params* sp,
                      const float* mode, const float* state0, const float* stlp, const float* valid, int N, float thres,
                      float inv_norm, float reg, float w_max, float a_max, float lr, int iter, float* params,
                      float* adam_m, float* adam_v, float* scores, void* workspace, pstl_stream_t stream);

/* Diversity metrics of the sampling test (measure_diversity, nusc_api.py:817-877), per (scene, lane mode):
 * std_out (n_scenes,3) = mean over the 2*nt way-point features of the population std over the samples with score > 0
 * (0 when none); vol_out (n_scenes,3) = sum over the steps of the convex-hull area of those samples' (x,y) (0 when the
 * lane is invalid, fewer than three samples are accepted or they are collinear).
 * trajs (n_scenes,m,3,2*nt) [x0,y0,x1,y1,...], scores / valids (n_scenes,m,3), m <= 128. */
int pstl_diversity(const float* trajs, const float* scores, const float* valids, int n_scenes, int m, int nt,
                   float* std_out, float* vol_out, pstl_stream_t stream);

/* Denoiser training step (diffusion_prep + net(...) + loss_diffusion, nusc_train.py:539-555, 1352-1356, 432-436).
 * pstl_denoiser_eps_rows: eps = policy_net([feature, x, temb(t_row), hl, stlp]) + x with ONE TIMESTEP PER ROW;
 * temb_rows (N, time_dim) is pos_encoding of each row's timestep (nusc_model.py:48-53).  fp32.
 * pstl_denoiser_eps_backward: from d_eps (N, 2*nt) to the gradients of policy_net.{0,2,4}.{weight,bias} (reference
 * shapes, row-major (out, in)) and, when d_scene_feat (n_scenes, feat_dim) is non-NULL, of the per-scene feature (the
 * encoders' backward continues from there).  Stateless (activations are recomputed); deterministic reductions.
 * Workspace: pstl_refine_backward_workspace_bytes for the backward, pstl_denoiser_workspace_bytes for the forward.
 * reuse_activations != 0 (both backward entries): the workspace is the very buffer the matching forward call
 * (pstl_denoiser_eps_rows / pstl_refine on a PSTL_PRECISION_FP32 handle; sized for the backward) wrote, with the same
 * arguments and weights, untouched since — the recompute is skipped. */
int pstl_denoiser_eps_rows(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene, const float* hl,
                           const float* stlp, const float* x, int N, const float* temb_rows, float* eps_out,
                           void* workspace, pstl_stream_t stream);
int pstl_denoiser_eps_backward(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene,
                               const float* hl, const float* stlp, const float* x, int N, const float* temb_rows,
                               const float* d_eps, float* g_p0_w, float* g_p0_b, float* g_p2_w, float* g_p2_b,
                               float* g_p4_w, float* g_p4_b, float* d_scene_feat, int reuse_activations, void* workspace,
                               pstl_stream_t stream);

/* RefineNet backward for the --rect_head training step (autograd over Net.rect_forward, nusc_model.py:182-235; the
 * optimiser upstream holds net.rect_net.parameters() only, nusc_train.py:1228-1233): from d_out = d loss / d rect_controls
 * (N, 2*nt) to the gradients of rect_net.{0,2,4}.{weight,bias} in the reference's own shapes (row-major (out, in):
 * g_r0_w (H, feat+7+2*nt), g_r2_w (H, H), g_r4_w (2*nt, H)).  Same row / scene arguments as pstl_refine.  The call is
 * stateless unless reuse_activations is set (see pstl_denoiser_eps_backward): it recomputes the fp32 activations in its
 * workspace (whatever the handle's precision), so it pairs with a forward made on a PSTL_PRECISION_FP32 handle.  Reductions over the rows run as split-K tiles summed in a fixed
 * order: results are deterministic.  merge_net and the scene encoders receive no gradient (not in upstream's optimiser
 * without --joint). */
size_t pstl_refine_backward_workspace_bytes(pstl_denoiser_t d, int N, int n_scenes);
int pstl_refine_backward(pstl_denoiser_t d, const float* scene_feat, int n_scenes, int rows_per_scene, const float* hl,
                         const float* stlp, const float* u0, const float* scores, int N, int n_randoms, int n_shards,
                         float w_max, float a_max, int clip_rect, const float* d_out, float* g_r0_w, float* g_r0_b,
                         float* g_r2_w, float* g_r2_b, float* g_r4_w, float* g_r4_b, int reuse_activations,
                         void* workspace, pstl_stream_t stream);

/* RefineNet training losses, value and gradient (compute_policy_loss, nusc_train.py:411 loss_stl, :439-466 the
 * --diverse_loss branch, :468-478 the plain branch).  Rows n = (scene*S + sample)*3 + mode as everywhere.
 *   loss_stl = stl_weight * mask_mean(relu(stl_nn_thres - scores), valid)
 *   diverse_loss:  groups of G = S/n_shards samples of one (scene, mode) (reshape/permute of :443, G <= 32);
 *                  sim = exp(-diversity_scale * ||s_i - s_j||) on controls / [w_max, a_max];
 *                  q = exp(score)[score>0]  (diverse_detach: [score>0], no gradient);
 *                  loss_diversity = -diversity_weight * mean_g tr(I - (Q sim Q + I)^-1);
 *                  loss_reg = mask_mean((rect - nn)^2, score >= 0)   (reported unweighted, as upstream);
 *                  loss = loss_stl + rect_reg_loss * loss_reg + loss_diversity
 *   otherwise:     loss_reg = rect_reg_loss * (mean(((rect-nn)[...,0]/w_max)^2) + mean(((rect-nn)[...,1]/a_max)^2));
 *                  extra_loss_reg = extra_rect_reg * (mean(relu((rect[...,0]/w_max)^2 - 1)) + the same for a_max);
 *                  loss = loss_stl + loss_reg + extra_loss_reg
 * losses[PSTL_LOSS_N]: see the enum.  d_rect (N, 2*nt) / d_scores (N) receive d loss / d rect_controls (with scores
 * held fixed) and d loss / d scores — chain d_scores through pstl_score_fused_bwd for the full gradient; both NULL for
 * the value only.  nn_controls is a constant (upstream detaches it).  Sums are reduced in a fixed order in fp64. */
typedef struct {
  int n_scenes, S, nt, n_shards;
  int diverse_loss, diverse_detach;
  float w_max, a_max;
  float stl_nn_thres, stl_weight;
  float diversity_scale, diversity_weight;
  float rect_reg_loss, extra_rect_reg;
} pstl_loss_cfg;
enum { PSTL_LOSS_TOTAL = 0, PSTL_LOSS_STL = 1, PSTL_LOSS_REG = 2, PSTL_LOSS_DIVERSITY = 3, PSTL_LOSS_EXTRA_REG = 4,
       PSTL_LOSS_MEAN_VALID = 5, PSTL_LOSS_MEAN_NONNEG = 6, PSTL_LOSS_N = 8 };
size_t pstl_refine_losses_workspace_bytes(const pstl_loss_cfg* cfg);
int pstl_refine_losses(const pstl_loss_cfg* cfg, const float* rect_controls, const float* nn_controls,
                       const float* scores, const float* valid, float* losses, float* d_rect, float* d_scores,
                       void* workspace, pstl_stream_t stream);

/* acc and scene_acc of the sampling test (mask_mean, nusc_train.py:23-27, 336-343) for scores / valid (n_scenes*S*3) with
 * rows n = (scene*S + sample)*3 + mode: out[0] = mean((score>0)*valid)/clip(mean(valid),1e-2); out[1] = the same for
 * (max over the samples of a (scene, mode)) > 0 with the scene's lane validity.  partial: n_scenes*4 floats of scratch. */
int pstl_accuracy(const float* scores, const float* valid, int n_scenes, int S, float* partial, float* out,
                  pstl_stream_t stream);

/* Three-layer ReLU MLP of the scene encoders in one launch (nusc_model.py:82-91, hidden width 256):
 * y (M,out) = W4 relu(W2 relu(W0 x + b0) + b2) + b4, weights (out_features, in_features) row-major as in nn.Linear.
 * Same summation order as three pstl_linear calls. */
int pstl_mlp3(const float* x, int M, int in_dim, const float* w0, const float* b0, const float* w2, const float* b2,
              const float* w4, const float* b4, int hidden, int out_dim, float* y, pstl_stream_t stream);
/* Several independent MLPs of that shape in ONE grid (the ego / neighbour / lane encoders side by side). */
#define PSTL_MLP3_MAX 4
typedef struct {
  const float *x, *w0, *b0, *w2, *b2, *w4, *b4;
  float* y;
  int M, in_dim, hidden, out_dim;
} pstl_mlp3_problem;
int pstl_mlp3_batch(const pstl_mlp3_problem* problems, int n, pstl_stream_t stream);

/* Scene-encoder glue (Net.encode_feat, nusc_model.py:55-95).
 * pstl_encoder_inputs: ego (n_scenes rows of >= 6 floats [x,y,th,v,L,W], row stride ego_row_stride), neighbors
 * (n_scenes,Knei,7), three lanes (n_scenes,nseg,3) and their validity ids (n_scenes) -> the inputs of the three encoder
 * MLPs after the ego-frame transform normalize_xyth (nusc_model.py:238-263): ego_in (n_scenes,6), nei_in
 * (n_scenes*Knei,7), lane_in (n_scenes*3, nseg*3) (first point, then differences of consecutive points).
 * pstl_encoder_pool: feature (n_scenes,7F) = [ego | min_k nei | mean_k nei | max_k nei | lanes] from the MLP outputs. */
int pstl_encoder_inputs(const float* ego, int ego_row_stride, const float* neighbors, const float* lane_c,
                        const float* lane_l, const float* lane_r, const float* id_c, const float* id_l,
                        const float* id_r, int n_scenes, int Knei, int nseg, float* ego_in, float* nei_in,
                        float* lane_in, pstl_stream_t stream);
int pstl_encoder_pool(const float* ego_feat, const float* nei_feat, const float* lane_feat, int n_scenes, int Knei,
                      int F, float* feature, pstl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PSTL_H_ */

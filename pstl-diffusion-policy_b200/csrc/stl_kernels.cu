// stl_kernels.cu — STL program interpreter kernels and the fused rollout+predicate+STL scoring
// kernels (forward, reverse, guidance).  Compiled with -fmad=false (see drive_core.cuh).
//
// Thread mapping: one thread per trajectory row.  The per-row traces ("tape") live in shared
// memory with stride blockDim+1 (bank-conflict free) or, when they do not fit, in a caller
// workspace with stride N (coalesced).  In the scene-indexed layout all rows of a block belong
// to one scene, whose lanes and pre-computed neighbour circles are staged once in shared memory
// and read as broadcasts.
#include <new>

#include "common.cuh"
#include "drive_eval.cuh"

struct pstl_program {
  PstlProgView h;        // host copy
  PstlProgView* d;       // device copy
  PstlPlan plan;         // closed form for the streaming scorer (valid == 0: interpreter only)
};

static const int kSmemBudget = 200 * 1024;

// --------------------------------------------------------------------------------------
// program handles
// --------------------------------------------------------------------------------------
extern "C" int pstl_program_create(const pstl_op* postfix, int n_ops, int n_signals, int T, int need_t,
                                   pstl_program_t* out) {
  PSTL_CHECK_ARG(postfix && out, "null argument");
  pstl_program* p = new (std::nothrow) pstl_program();
  PSTL_CHECK_ARG(p, "out of host memory");
  char err[256];
  if (pstl_resolve_program(postfix, n_ops, n_signals, T, need_t, &p->h, err, sizeof(err)) != 0) {
    delete p;
    pstl_set_error("pstl_program_create: %s", err);
    return PSTL_ERR_ARG;
  }
  p->d = nullptr;
  pstl_make_plan(p->h, &p->plan);
  cudaError_t e = cudaMalloc(&p->d, sizeof(PstlProgView));
  if (e == cudaSuccess) e = cudaMemcpy(p->d, &p->h, sizeof(PstlProgView), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (p->d) cudaFree(p->d);
    delete p;
    pstl_set_error("pstl_program_create: %s", cudaGetErrorString(e));
    return PSTL_ERR_CUDA;
  }
  *out = p;
  return PSTL_OK;
}

extern "C" int pstl_program_destroy(pstl_program_t prog) {
  if (!prog) return PSTL_OK;
  cudaFree(prog->d);
  delete prog;
  return PSTL_OK;
}

extern "C" int pstl_program_tape_floats(pstl_program_t prog, int with_grad) {
  if (!prog) return PSTL_ERR_ARG;
  return with_grad ? prog->h.grad_floats : prog->h.val_floats;
}

// pick block size / tape placement for F floats per row (+ extra shared bytes)
struct TapePlan {
  int block;
  int smem_tape;  // 1: shared, 0: workspace
  size_t smem_bytes;
};

static TapePlan plan_tape(int F, size_t extra, int prefer_block) {
  TapePlan p;
  for (int b = prefer_block; b >= 32; b >>= 1) {
    size_t bytes = extra + (size_t)F * (b + 1) * sizeof(float);
    if (bytes <= (size_t)kSmemBudget) {
      p.block = b;
      p.smem_tape = 1;
      p.smem_bytes = bytes;
      return p;
    }
  }
  p.block = prefer_block;
  p.smem_tape = 0;
  p.smem_bytes = extra;
  return p;
}

// --------------------------------------------------------------------------------------
// generic signal programs:  node(x, tau, d) of stl_d_lib.py on pre-evaluated AP signals
// --------------------------------------------------------------------------------------
struct LeafNone {
  __device__ float signal(int, int) const { return 0.f; }
  __device__ void signal(int, int, float) const {}
  __device__ PstlIn pred_in(int, int) const { return PstlIn{nullptr, 1.f, 0.f, 1.f, 0}; }
  __device__ PstlOut pred_out(int, int) const { return PstlOut{nullptr, 0.f}; }
};

template <bool BWD>
__global__ void k_stl_signals(const PstlProgView* __restrict__ dprog, const float* __restrict__ sig,
                              const float* __restrict__ grad_trace, int N, int P, int T, float tau, int hard,
                              float* __restrict__ out_trace, float* __restrict__ out_t0, float* __restrict__ grad_sig,
                              float* __restrict__ ws, int smem_tape) {
  extern __shared__ float sm[];
  __shared__ PstlProgView prog;
  {
    const int* src = reinterpret_cast<const int*>(dprog);
    int* dst = reinterpret_cast<int*>(&prog);
    for (int i = threadIdx.x; i < (int)(sizeof(PstlProgView) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int B = blockDim.x;
  const int n0 = blockIdx.x * B;
  const int cnt = min(B, N - n0);
  const int n = n0 + threadIdx.x;
  const int F = prog.val_floats;
  float *vt, *gt = nullptr;
  int stride;
  if (smem_tape) {
    stride = B + 1;
    vt = sm + threadIdx.x;
    if (BWD) gt = sm + (size_t)F * stride + threadIdx.x;
  } else {
    stride = N;
    vt = ws + min(n, N - 1);
    if (BWD) gt = ws + (size_t)F * N + min(n, N - 1);
  }
  // stage the (cnt,P,T) signal chunk transposed into the tape: coalesced global reads
  const int PT_ = P * T;
  const float* chunk = sig + (size_t)n0 * PT_;
  for (int e = threadIdx.x; e < cnt * PT_; e += B) {
    const int r = e / PT_, q = e - r * PT_;
    const float v = chunk[e];
    if (smem_tape) sm[(size_t)q * stride + r] = v; else ws[(size_t)q * N + n0 + r] = v;
  }
  __syncthreads();
  LeafNone leaf;
  const PstlROp top = prog.ops[prog.n_ops - 1];
  if (n < N) {
    pstl_interp_fwd(prog, vt, stride, tau, hard, leaf);
    if (!BWD) {
      if (out_trace)
        for (int t = 0; t < prog.need_t; ++t) out_trace[(size_t)n * prog.need_t + t] = vt[(size_t)(top.out_off + t) * stride];
      if (out_t0) out_t0[n] = vt[(size_t)top.out_off * stride];
    } else {
      for (int i = 0; i < F; ++i) gt[(size_t)i * stride] = 0.f;
      for (int t = 0; t < prog.need_t; ++t) gt[(size_t)(top.out_off + t) * stride] += grad_trace[(size_t)n * prog.need_t + t];
      pstl_interp_bwd(prog, vt, gt, stride, tau, hard, leaf, leaf);
    }
  }
  if (BWD) {
    __syncthreads();
    float* gchunk = grad_sig + (size_t)n0 * PT_;
    for (int e = threadIdx.x; e < cnt * PT_; e += B) {
      const int r = e / PT_, q = e - r * PT_;
      gchunk[e] = smem_tape ? sm[(size_t)F * stride + (size_t)q * stride + r] : ws[(size_t)F * N + (size_t)q * N + n0 + r];
    }
  }
}

extern "C" size_t pstl_stl_workspace_bytes(pstl_program_t prog, int N, int with_grad) {
  if (!prog) return 0;
  const int F = prog->h.val_floats * (with_grad ? 2 : 1);
  TapePlan p = plan_tape(F, 0, 128);
  return p.smem_tape ? 0 : (size_t)F * N * sizeof(float);
}

static int launch_signals(bool bwd, pstl_program_t prog, const float* sig, const float* grad_trace, int N, int P,
                          int T, float tau, int hard, float* out_trace, float* out_t0, float* grad_sig, void* ws,
                          pstl_stream_t stream) {
  PSTL_CHECK_ARG(prog && sig, "null argument");
  PSTL_CHECK_ARG(P == prog->h.n_signals && T == prog->h.T, "signal shape does not match the program");
  if (N <= 0) return PSTL_OK;
  const int F = prog->h.val_floats * (bwd ? 2 : 1);
  TapePlan p = plan_tape(F, 0, 128);
  PSTL_CHECK_ARG(p.smem_tape || ws, "workspace required (see pstl_stl_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = pstl_ceil_div(N, p.block);
  if (bwd) {
    PSTL_CUDA(cudaFuncSetAttribute(k_stl_signals<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    k_stl_signals<true><<<grid, p.block, p.smem_bytes, st>>>(prog->d, sig, grad_trace, N, P, T, tau, hard, nullptr,
                                                             nullptr, grad_sig, (float*)ws, p.smem_tape);
  } else {
    PSTL_CUDA(cudaFuncSetAttribute(k_stl_signals<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    k_stl_signals<false><<<grid, p.block, p.smem_bytes, st>>>(prog->d, sig, nullptr, N, P, T, tau, hard, out_trace,
                                                              out_t0, nullptr, (float*)ws, p.smem_tape);
  }
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

extern "C" int pstl_stl_eval_signals(pstl_program_t prog, const float* sig, int N, int P, int T, float tau, int hard,
                                     float* out_trace, float* out_t0, void* workspace, pstl_stream_t stream) {
  return launch_signals(false, prog, sig, nullptr, N, P, T, tau, hard, out_trace, out_t0, nullptr, workspace, stream);
}

extern "C" int pstl_stl_eval_signals_bwd(pstl_program_t prog, const float* sig, const float* grad_trace, int N, int P,
                                         int T, float tau, int hard, float* grad_sig, void* workspace,
                                         pstl_stream_t stream) {
  PSTL_CHECK_ARG(grad_trace && grad_sig, "null gradient buffers");
  return launch_signals(true, prog, sig, grad_trace, N, P, T, tau, hard, nullptr, nullptr, grad_sig, workspace, stream);
}

// --------------------------------------------------------------------------------------
// fused scoring
// --------------------------------------------------------------------------------------
#define PSTL_NEI_W 16  // floats per staged (neighbour, step): cx[4], cy[4], r, valid, centre x, y, L/2, pad
struct SceneSmem {
  const float* circ;  // (K,T,PSTL_NEI_W)
  const float* ln;    // (3,nseg,3)
  int K, T, nseg;
  __device__ float lane(int l, int j, int f) const { return ln[(l * nseg + j) * 3 + f]; }
  __device__ void nei_meta(int k, int t, float& cx, float& cy, float& reach, float& valid) const {
    const float4 m = *reinterpret_cast<const float4*>(circ + (k * T + t) * PSTL_NEI_W + 8);
    const float4 m2 = *reinterpret_cast<const float4*>(circ + (k * T + t) * PSTL_NEI_W + 12);
    valid = m.y; cx = m.z; cy = m.w; reach = m2.x;
  }
  __device__ void nei(int k, int t, PstlNei& out) const {
    const float* p = circ + (k * T + t) * PSTL_NEI_W;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    out.cx[0] = a.x; out.cx[1] = a.y; out.cx[2] = a.z; out.cx[3] = a.w;
    out.cy[0] = b.x; out.cy[1] = b.y; out.cy[2] = b.z; out.cy[3] = b.w;
    out.r = p[8];
    out.valid = p[9];
  }
};

struct ScoreArgs {
  const PstlProgView* progs[3];
  const float* neighbors;
  const float* lanes[3];
  int n_scenes, rows_per_scene;
  PstlEvalCfg cfg;
  const float *mode, *state0, *controls, *ego, *stlp;
  int ego_stride, N, C;
  float *scores_all, *best_score, *best_controls, *traj_out;
  int32_t* best_idx;
  // reverse mode
  const float *grad_score, *valid;
  float thres, inv_norm;
  const float* inv_norm_dev;  // when set, the loss normaliser is read from device memory (graph-capturable guidance)
  float *scores, *grad_controls, *grad_ego;
  float* ws;
  int smem_tape, F;
  int tc;           // streaming scorer: steps of the horizon held by the shared-memory scene tile
  int tape_global;  // streaming scorer: X(t) columns in the workspace (stride N) instead of shared memory
};

__device__ __forceinline__ size_t scene_tile_floats(const PstlEvalCfg& c) {
  return (size_t)c.K * c.T * PSTL_NEI_W + (size_t)3 * c.nseg * 3;
}

// block-cooperative staging of one scene into shared memory
__device__ void stage_scene(const ScoreArgs& a, int scene, float* tile) {
  const PstlEvalCfg& c = a.cfg;
  const int W = PSTL_NEI_W;
  float* circ = tile;
  float* ln = tile + (size_t)c.K * c.T * W;
  const float* nb = a.neighbors + (size_t)scene * c.K * c.T * 7;
  for (int e = threadIdx.x; e < c.K * c.T; e += blockDim.x) {
    const float* p = nb + (size_t)e * 7;
    PstlCircles cc;
    pstl_car_circles(p[1], p[2], cosf(p[3]), sinf(p[3]), p[5], p[6], cc);
    float* o = circ + (size_t)e * W;
    for (int i = 0; i < PSTL_NL; ++i) { o[i] = cc.cx[i]; o[PSTL_NL + i] = cc.cy[i]; }
    o[8] = cc.r; o[9] = p[0]; o[10] = p[1]; o[11] = p[2]; o[12] = pstl_car_reach(p[5], p[6]);
    o[13] = o[14] = o[15] = 0.f;
  }
  for (int l = 0; l < 3; ++l) {
    const float* src = a.lanes[l] + (size_t)scene * c.nseg * 3;
    for (int e = threadIdx.x; e < c.nseg * 3; e += blockDim.x) ln[l * c.nseg * 3 + e] = src[e];
  }
}

template <bool SMEM_SCENE, bool BWD>
__global__ void __launch_bounds__(256) k_score(ScoreArgs a) {
  extern __shared__ float sm[];
  __shared__ PstlProgView progs[3];
  for (int k = 0; k < 3; ++k) {
    const int* src = reinterpret_cast<const int*>(a.progs[k]);
    int* dst = reinterpret_cast<int*>(&progs[k]);
    for (int i = threadIdx.x; i < (int)(sizeof(PstlProgView) / 4); i += blockDim.x) dst[i] = src[i];
  }
  const PstlEvalCfg c = a.cfg;
  const int B = blockDim.x;
  const int n0 = blockIdx.x * B;
  // rows of the pipeline cycle through the three formulas (n % 3): deal them to warps so that a warp
  // interprets ONE program (purely a divergence optimisation; every thread still reads its own mode)
  int li = threadIdx.x;
  if (B % 96 == 0) {
    const int g = li / 96, w = li - g * 96;
    li = g * 96 + (w & 31) * 3 + (w >> 5);
  }
  const int n = n0 + li;
  float* tile = sm;
  float* tape0 = sm;
  if (SMEM_SCENE) {
    stage_scene(a, n0 / a.rows_per_scene, tile);
    tape0 = sm + ((scene_tile_floats(c) + 3) & ~(size_t)3);
  }
  __syncthreads();
  if (n >= a.N) return;
  const int T = c.T;
  float *vt, *gt = nullptr, *pt = nullptr;
  int stride;
  if (a.smem_tape) {
    stride = B + 1;
    vt = tape0 + threadIdx.x;
  } else {
    stride = a.N;
    vt = a.ws + n;
  }
  const float md = a.mode[n];
  const int m = (md == 0.f) ? 0 : (md == 1.f) ? 1 : (md == 2.f) ? 2 : (md == 3.f) ? 3 : 4;
  const PstlProgView& P = progs[m < 3 ? m : 0];
  if (BWD) {
    gt = vt + (size_t)P.val_floats * stride;
    pt = vt + (size_t)P.part_off * stride;
  }
  PstlPose s0{0.f, 0.f, 0.f, 0.f};
  if (a.state0) { s0.x = a.state0[n * 4 + 0]; s0.y = a.state0[n * 4 + 1]; s0.th = a.state0[n * 4 + 2]; s0.v = a.state0[n * 4 + 3]; }
  const float* stlp = a.stlp + (size_t)n * 6;
  const float* ego = a.ego ? a.ego + (size_t)n * T * a.ego_stride : nullptr;
  const int scene = n / a.rows_per_scene;

  SceneSmem ss{tile, tile + (size_t)c.K * c.T * PSTL_NEI_W, c.K, c.T, c.nseg};
  PstlSceneGlobal sg;
  sg.neib = a.neighbors + (size_t)scene * c.K * c.T * 7;
  for (int l = 0; l < 3; ++l) sg.ln[l] = a.lanes[l] + (size_t)scene * c.nseg * 3;
  sg.K = c.K; sg.T = c.T;

  if (!BWD) {
    float best = -INFINITY;
    int bi = 0;
    for (int cand = 0; cand < a.C; ++cand) {
      const float* u = a.controls ? a.controls + ((size_t)cand * a.N + n) * T * 2 : nullptr;
      float sc;
      if (m < 3) {
        sc = SMEM_SCENE ? pstl_eval_traj<SceneSmem, false, true>(P, ss, c, s0, u, ego, a.ego_stride, stlp, vt, nullptr, stride)
                        : pstl_eval_traj<PstlSceneGlobal, false, true>(P, sg, c, s0, u, ego, a.ego_stride, stlp, vt, nullptr, stride);
      } else {
        sc = (m == 3) ? 1.0f : 0.0f;  // nusc_train.py:322 outlier score; unknown mode selects nothing (:150-151)
      }
      if (a.scores_all) a.scores_all[(size_t)cand * a.N + n] = sc;
      if (cand == 0 || sc > best) { best = sc; bi = cand; }  // torch.max(dim=0): first maximum
    }
    if (a.best_score) a.best_score[n] = best;
    if (a.best_idx) a.best_idx[n] = bi;
    if ((a.best_controls || a.traj_out) && a.controls) {
      const float* u = a.controls + ((size_t)bi * a.N + n) * T * 2;
      PstlPose s = s0;
      for (int t = 0; t < T; ++t) {
        float w, ac;
        pstl_scaled_control(u, t, c, w, ac);
        if (a.best_controls) { a.best_controls[((size_t)n * T + t) * 2] = w; a.best_controls[((size_t)n * T + t) * 2 + 1] = ac; }
        if (a.traj_out) {
          float* o = a.traj_out + ((size_t)n * (T + 1) + t) * 4;
          o[0] = s.x; o[1] = s.y; o[2] = s.th; o[3] = s.v;
          s = pstl_unicycle_step(s, w, ac, c.dt, cosf(s.th), sinf(s.th));
        }
      }
      if (a.traj_out) {
        float* o = a.traj_out + ((size_t)n * (T + 1) + T) * 4;
        o[0] = s.x; o[1] = s.y; o[2] = s.th; o[3] = s.v;
      }
    }
  } else {
    const float* u = a.controls ? a.controls + (size_t)n * T * 2 : nullptr;
    float* gu = a.grad_controls ? a.grad_controls + (size_t)n * T * 2 : nullptr;
    float* ge = a.grad_ego ? a.grad_ego + (size_t)n * T * 4 : nullptr;
    if (m >= 3) {
      if (a.scores) a.scores[n] = (m == 3) ? 1.0f : 0.0f;
      if (gu) for (int i = 0; i < T * 2; ++i) gu[i] = 0.f;
      if (ge) for (int i = 0; i < T * 4; ++i) ge[i] = 0.f;
      return;
    }
    const float sc = SMEM_SCENE ? pstl_eval_traj<SceneSmem, true, true>(P, ss, c, s0, u, ego, a.ego_stride, stlp, vt, pt, stride)
                                : pstl_eval_traj<PstlSceneGlobal, true, true>(P, sg, c, s0, u, ego, a.ego_stride, stlp, vt, pt, stride);
    if (a.scores) a.scores[n] = sc;
    float g;
    if (a.grad_score) {
      g = a.grad_score[n];
    } else {  // guidance loss (nusc_train.py:616-619): mean(relu(thres-score)*valid)/clip(mean(valid),1e-2)
      g = (a.thres - sc > 0.f) ? -a.valid[n] * (a.inv_norm_dev ? __ldg(a.inv_norm_dev) : a.inv_norm) : 0.f;
    }
    pstl_eval_traj_bwd<true>(P, c, u, stlp, g, vt, gt, pt, stride, gu, ge);
  }
}

#include "score_warp.cuh"
#include "score_stream.cuh"
#include "score_dense.cuh"

// PSTL_SCORE_KERNEL=stream|warp|thread pins the forward scoring kernel (tests compare the three)
static bool score_kernel_forced(const char* name) {
  const char* e = getenv("PSTL_SCORE_KERNEL");
  return e && strcmp(e, name) == 0;
}

static int max3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

// forward scoring on the warp-per-trajectory kernel (T <= 31); returns 1 if it took the launch
static int launch_score_warp(ScoreArgs& a, pstl_program_t const* progs, cudaStream_t st, int* took) {
  *took = 0;
  const PstlEvalCfg& c = a.cfg;
  if (c.T > 31 || score_kernel_forced("thread")) return PSTL_OK;
  a.F = max3(progs[0]->h.n_ops, progs[1]->h.n_ops, progs[2]->h.n_ops);
  const size_t stack_bytes = (size_t)8 * a.F * 32 * sizeof(float);
  const size_t tile_bytes = ((((size_t)c.K * PSTL_SOA_F * c.T + (size_t)9 * c.nseg + 3) & ~(size_t)3)) * sizeof(float);
  const bool smem_scene = a.rows_per_scene % PSTL_WARP_ROWS == 0 && tile_bytes + stack_bytes <= 64 * 1024;
  const int grid = pstl_ceil_div(a.N, PSTL_WARP_ROWS);
  WarpProgs wp;
  for (int k = 0; k < 3; ++k) wp.p[k] = progs[k]->h;
  if (smem_scene) {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    k_score_warp<true><<<grid, 256, tile_bytes + stack_bytes, st>>>(a, wp);
  } else {
    k_score_warp<false><<<grid, 256, stack_bytes, st>>>(a, wp);
  }
  PSTL_LAUNCH_CHECK();
  *took = 1;
  return PSTL_OK;
}

// forward scoring of the reference's dense per-row layout on the time-parallel kernel (score_dense.cuh);
// sets *took when it owned the launch
static int launch_score_dense_tp(ScoreArgs& a, pstl_program_t const* progs, cudaStream_t st, int* took) {
  *took = 0;
  const PstlEvalCfg& c = a.cfg;
  const bool forced = score_kernel_forced("dense");
  if (c.hard || score_kernel_forced("warp") || score_kernel_forced("thread") || score_kernel_forced("stream")) return PSTL_OK;
  if (a.rows_per_scene != 1 || !a.ego || a.controls || a.C != 1 || a.traj_out || a.best_controls) return PSTL_OK;
  if (c.T > PSTL_DENSE_MAX_THREADS || c.T < 1 || c.K < 1) return PSTL_OK;
  if (((size_t)c.K * c.T) % 4 != 0 || ((uintptr_t)a.neighbors & 15) != 0) return PSTL_OK;  // 16-byte bulk copies
  if (a.ego_stride == 4 && ((uintptr_t)a.ego & 15) != 0) return PSTL_OK;
  if (!forced && a.N < 1024) return PSTL_OK;  // tiny calls stay on the thread-per-row kernel (one wave either way)
  StreamPlans sp;
  for (int k = 0; k < 3; ++k) {
    if (!progs[k]->plan.valid) return PSTL_OK;
    sp.p[k] = progs[k]->plan;
  }
  DenseTpCfg d;
  const char* er = getenv("PSTL_DENSE_ROWS");
  int R = er ? atoi(er) : (c.T <= 21 ? 12 : 8);  // measured at T=20, K=8: 12 rows (240 threads) 2.39 ms per 1M rows, 8 rows 2.50
  if (R < 1) R = 1;
  if (R > PSTL_DENSE_MAX_THREADS / c.T) R = PSTL_DENSE_MAX_THREADS / c.T;
  if (R > 32) R = 32;
  const char* eb = getenv("PSTL_DENSE_CHUNK_KB");
  const size_t chunk_budget = (size_t)(eb ? atoi(eb) : 32) * 1024;
  int KC = 0;
  for (int kc = c.K; kc >= 1; --kc)
    if ((size_t)R * kc * c.T * 28 <= chunk_budget && ((size_t)kc * c.T * 7) % 4 == 0) { KC = kc; break; }
  if (!eb && (size_t)R * c.K * c.T * 28 <= 64 * 1024) KC = c.K;  // the whole block in one go when it is small enough
  if (!KC) return PSTL_OK;
  const char* enb = getenv("PSTL_DENSE_NBUF");
  d.R = R; d.KC = KC; d.nbuf = (KC < c.K) ? ((enb && atoi(enb) == 1) ? 1 : 2) : 1; d.chunk_floats = KC * c.T * 7;
  const size_t smem = dtp_smem_layout(d, c.T, c.nseg).total * sizeof(float);
  if (smem > 160 * 1024 || c.nseg * 3 > 64 || R * c.T > 65535) return PSTL_OK;
  int threads = ((R * c.T + 31) / 32) * 32;
  if (threads < 64) threads = 64;  // the lane staging deals 64 threads per row
  if (threads < ((R * PSTL_MAX_TERMS + 31) & ~31)) threads = (R * PSTL_MAX_TERMS + 31) & ~31;
  const int per = c.nseg * 3;
  d.bulk_small = (a.ego_stride == 4 && (R * per) % 4 == 0 && R % 4 == 0 && ((uintptr_t)a.ego & 15) == 0 &&
                  ((uintptr_t)a.lanes[0] & 15) == 0 && ((uintptr_t)a.lanes[1] & 15) == 0 && ((uintptr_t)a.lanes[2] & 15) == 0 &&
                  ((uintptr_t)a.stlp & 15) == 0 && ((uintptr_t)a.mode & 15) == 0) ? 1 : 0;
  if (c.T == 20 && c.nseg == 15) {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_dense_tp<20, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    k_score_dense_tp<20, 15><<<pstl_ceil_div(a.N, R), threads, smem, st>>>(a, sp, d);
  } else {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_dense_tp<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    k_score_dense_tp<0, 0><<<pstl_ceil_div(a.N, R), threads, smem, st>>>(a, sp, d);
  }
  PSTL_LAUNCH_CHECK();
  *took = 1;
  return PSTL_OK;
}

// forward scoring on the streaming kernel (all three programs have a plan, soft semantics);
// sets *took when it owned the launch
static int launch_score_stream(ScoreArgs& a, pstl_program_t const* progs, cudaStream_t st, int* took) {
  *took = 0;
  const PstlEvalCfg& c = a.cfg;
  if (c.hard || score_kernel_forced("warp") || score_kernel_forced("thread")) return PSTL_OK;
  StreamPlans sp;
  int n_tapes = 0;
  for (int k = 0; k < 3; ++k) {
    if (!progs[k]->plan.valid) return PSTL_OK;
    sp.p[k] = progs[k]->plan;
    n_tapes = progs[k]->plan.n_tapes > n_tapes ? progs[k]->plan.n_tapes : n_tapes;
  }
  // scene tile: as many steps of the horizon as fit in ~48 KB (all of them at the pipeline's shape)
  const size_t budget = 100 * 1024, tile_budget = 48 * 1024;
  int tc = c.T;
  while (tc > 1 && stream_tile_f4(c.K, tc, c.nseg) * sizeof(float4) > tile_budget) tc = (tc + 1) / 2;
  const size_t tile_bytes = stream_tile_f4(c.K, tc, c.nseg) * sizeof(float4);
  int block = 0;
  bool smem_scene = false, tape_global = false;
  const size_t tape_row = (size_t)n_tapes * c.T * sizeof(float);
  if (a.rows_per_scene % 32 == 0 && tile_bytes <= tile_budget && (tc == c.T || tc >= 4)) {
    // one scene per block, staged in shared memory (whole horizon, or chunk by chunk)
    static const int cands[] = {192, 96, 128, 64, 32};
    int first = 0;  // largest block that divides the scene's rows
    for (int b : cands)
      if (!first && a.rows_per_scene % b == 0) first = b;
    if (!first) first = 32;
    if (tile_bytes + (size_t)first * tape_row <= budget) {
      block = first;
      smem_scene = true;
    } else if (a.ws) {  // long horizon: keep the block large, the X(t) columns go to the workspace (coalesced)
      block = first;
      smem_scene = true;
      tape_global = true;
    } else {
      for (int b : cands)
        if (a.rows_per_scene % b == 0 && tile_bytes + (size_t)b * tape_row <= budget) { block = b; smem_scene = true; break; }
    }
  }
  if (!block) {
    static const int cands[] = {128, 64, 32};
    for (int b : cands)
      if ((size_t)b * tape_row <= budget / 2) {
        block = b;
        break;
      }
    if (!block && a.ws) { block = 128; tape_global = true; }
  }
  if (!block) return PSTL_OK;
  a.tc = smem_scene ? tc : c.T;
  a.tape_global = tape_global ? 1 : 0;
  const size_t smem = (smem_scene ? tile_bytes : 0) + (tape_global ? 0 : (size_t)block * tape_row);
  const int grid = pstl_ceil_div(a.N, block);
  if (smem_scene && tc < c.T) {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_stream<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    k_score_stream<true, true><<<grid, block, smem, st>>>(a, sp);
  } else if (smem_scene) {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_stream<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    k_score_stream<true, false><<<grid, block, smem, st>>>(a, sp);
  } else {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_stream<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    k_score_stream<false, false><<<grid, block, smem, st>>>(a, sp);
  }
  PSTL_LAUNCH_CHECK();
  *took = 1;
  return PSTL_OK;
}

// reverse-mode scoring on the streaming kernel; sets *took when it owned the launch
static size_t stream_bwd_ws_floats(pstl_program_t const* progs, int N, int T) {
  int n_tapes = 0;
  for (int k = 0; k < 3; ++k) {
    if (!progs[k]->plan.valid) return 0;
    n_tapes = progs[k]->plan.n_tapes > n_tapes ? progs[k]->plan.n_tapes : n_tapes;
  }
  return (size_t)pstl_stream_grad_floats(n_tapes, T) * N;
}

static int launch_score_stream_bwd(ScoreArgs& a, pstl_program_t const* progs, cudaStream_t st, int* took) {
  *took = 0;
  const PstlEvalCfg& c = a.cfg;
  if (c.hard || score_kernel_forced("warp") || score_kernel_forced("thread") || !a.ws) return PSTL_OK;
  if (!stream_bwd_ws_floats(progs, a.N, c.T)) return PSTL_OK;
  StreamPlans sp;
  for (int k = 0; k < 3; ++k) sp.p[k] = progs[k]->plan;
  const size_t budget = 100 * 1024, tile_budget = 48 * 1024;
  int tc = c.T;
  while (tc > 1 && stream_tile_f4(c.K, tc, c.nseg) * sizeof(float4) > tile_budget) tc = (tc + 1) / 2;
  const size_t tile_bytes = stream_tile_f4(c.K, tc, c.nseg) * sizeof(float4);
  int block = 128;
  bool smem_scene = false;
  if (a.rows_per_scene % 32 == 0 && tile_bytes <= tile_budget && (tc == c.T || tc >= 4)) {
    static const int cands[] = {192, 96, 128, 64, 32};
    for (int b : cands)
      if (a.rows_per_scene % b == 0) { block = b; smem_scene = true; break; }
  }
  a.tc = smem_scene ? tc : c.T;
  const int grid = pstl_ceil_div(a.N, block);
  if (smem_scene && tc < c.T) {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_stream_bwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    k_score_stream_bwd<true, true><<<grid, block, tile_bytes, st>>>(a, sp);
  } else if (smem_scene) {
    PSTL_CUDA(cudaFuncSetAttribute(k_score_stream_bwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    k_score_stream_bwd<true, false><<<grid, block, tile_bytes, st>>>(a, sp);
  } else {
    k_score_stream_bwd<false, false><<<grid, block, 0, st>>>(a, sp);
  }
  PSTL_LAUNCH_CHECK();
  *took = 1;
  return PSTL_OK;
}

static int fill_cfg(const pstl_scene_view* sv, const pstl_spec_params* sp, PstlEvalCfg* c) {
  c->dt = sp->dt; c->tau = sp->tau; c->ego_L = sp->ego_L; c->ego_W = sp->ego_W;
  c->w_scale = sp->w_scale; c->a_scale = sp->a_scale;
  c->clip_controls = sp->clip_controls; c->clip_dist = sp->clip_dist; c->hard = sp->hard;
  c->nseg = sv->nseg; c->K = sv->Knei; c->T = sv->T;
  return 0;
}

struct ScorePlan {
  TapePlan tp;
  int smem_scene;
  int F;
};

static ScorePlan plan_score(pstl_program_t const* progs, const pstl_scene_view* sv, int N, int with_grad) {
  ScorePlan sp;
  sp.F = with_grad ? max3(progs[0]->h.grad_floats, progs[1]->h.grad_floats, progs[2]->h.grad_floats)
                   : max3(progs[0]->h.val_floats, progs[1]->h.val_floats, progs[2]->h.val_floats);
  const size_t tile = (((size_t)sv->Knei * sv->T * PSTL_NEI_W + (size_t)9 * sv->nseg + 3) & ~(size_t)3) * sizeof(float);
  // one scene per block: the block size must divide rows_per_scene
  sp.smem_scene = 0;
  const int prefer = with_grad ? 64 : 128;
  if (sv->rows_per_scene >= 32 && tile <= 96 * 1024) {
    static const int cands[] = {192, 96, 128, 64, 32};
    for (int b : cands) {
      if (sv->rows_per_scene % b != 0) continue;
      const size_t bytes = tile + (size_t)sp.F * (b + 1) * sizeof(float);
      if (bytes > (size_t)kSmemBudget / 2 && b > 32) continue;  // keep at least two blocks per SM
      if (bytes > (size_t)kSmemBudget) continue;
      sp.tp.block = b; sp.tp.smem_tape = 1; sp.tp.smem_bytes = bytes;
      sp.smem_scene = 1;
      return sp;
    }
    for (int b : cands)
      if (sv->rows_per_scene % b == 0) {  // tape too large for shared memory: workspace tape, scene tile stays
        sp.tp.block = b; sp.tp.smem_tape = 0; sp.tp.smem_bytes = tile;
        sp.smem_scene = 1;
        return sp;
      }
  }
  sp.tp = plan_tape(sp.F, 0, prefer);
  return sp;
}

extern "C" size_t pstl_score_workspace_bytes(pstl_program_t const* progs, int N, int T, int with_grad) {
  if (!progs || !progs[0] || !progs[1] || !progs[2]) return 0;
  // conservative: size for the workspace tape; the launch uses shared memory when it fits
  const int F = with_grad ? max3(progs[0]->h.grad_floats, progs[1]->h.grad_floats, progs[2]->h.grad_floats)
                          : max3(progs[0]->h.val_floats, progs[1]->h.val_floats, progs[2]->h.val_floats);
  TapePlan t = plan_tape(F, 96 * 1024, 32);
  const size_t interp = t.smem_tape ? 0 : (size_t)F * N * sizeof(float);
  size_t stream = with_grad ? stream_bwd_ws_floats(progs, N, T) * sizeof(float) : 0;
  if (!with_grad && stream_bwd_ws_floats(progs, N, T)) {  // forward: X(t) columns when they do not fit in shared memory
    int n_tapes = 0;
    for (int k = 0; k < 3; ++k) n_tapes = progs[k]->plan.n_tapes > n_tapes ? progs[k]->plan.n_tapes : n_tapes;
    stream = (size_t)n_tapes * T * N * sizeof(float);
  }
  return interp > stream ? interp : stream;
}

template <bool BWD>
static int launch_score(ScoreArgs& a, const ScorePlan& sp, cudaStream_t st) {
  const int grid = pstl_ceil_div(a.N, sp.tp.block);
  a.smem_tape = sp.tp.smem_tape;
  a.F = sp.F;
  if (sp.smem_scene) {
    PSTL_CUDA(cudaFuncSetAttribute(k_score<true, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    k_score<true, BWD><<<grid, sp.tp.block, sp.tp.smem_bytes, st>>>(a);
  } else {
    PSTL_CUDA(cudaFuncSetAttribute(k_score<false, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    k_score<false, BWD><<<grid, sp.tp.block, sp.tp.smem_bytes, st>>>(a);
  }
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

static int check_score_args(pstl_program_t const* progs, const pstl_scene_view* sv, const pstl_spec_params* sp,
                            const float* mode, const float* stlp, int N) {
  PSTL_CHECK_ARG(progs && progs[0] && progs[1] && progs[2], "three programs required");
  PSTL_CHECK_ARG(sv && sp && mode && stlp, "null argument");
  PSTL_CHECK_ARG(sv->rows_per_scene >= 1 && sv->Knei >= 0 && sv->nseg >= 2, "bad scene view");
  for (int k = 0; k < 3; ++k) PSTL_CHECK_ARG(progs[k]->h.T == sv->T && progs[k]->h.n_signals == 0, "program/scene horizon mismatch");
  PSTL_CHECK_ARG((long long)sv->n_scenes * sv->rows_per_scene >= N, "fewer scene rows than trajectories");
  return PSTL_OK;
}

static void base_args(ScoreArgs& a, pstl_program_t const* progs, const pstl_scene_view* sv,
                      const pstl_spec_params* sp) {
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < 3; ++k) { a.progs[k] = progs[k]->d; a.lanes[k] = sv->lanes[k]; }
  a.neighbors = sv->neighbors;
  a.n_scenes = sv->n_scenes;
  a.rows_per_scene = sv->rows_per_scene;
  fill_cfg(sv, sp, &a.cfg);
}

extern "C" int pstl_score_fused(pstl_program_t const* progs, const pstl_scene_view* scenes, const pstl_spec_params* sp,
                                const float* mode, const float* state0, const float* controls, int C,
                                const float* ego_traj, int ego_stride, const float* stlp, int N, float* scores_all,
                                float* best_score, int32_t* best_idx, float* best_controls, float* traj_out,
                                void* workspace, pstl_stream_t stream) {
  int rc = check_score_args(progs, scenes, sp, mode, stlp, N);
  if (rc) return rc;
  PSTL_CHECK_ARG((controls && state0 && C >= 1) || (ego_traj && ego_stride >= 4), "need controls+state0 or ego_traj");
  if (N <= 0) return PSTL_OK;
  ScoreArgs a;
  base_args(a, progs, scenes, sp);
  a.mode = mode; a.state0 = state0; a.controls = ego_traj ? nullptr : controls; a.ego = ego_traj;
  a.ego_stride = ego_stride; a.stlp = stlp; a.N = N; a.C = ego_traj ? 1 : C;
  a.scores_all = scores_all; a.best_score = best_score; a.best_idx = best_idx;
  a.best_controls = best_controls; a.traj_out = traj_out; a.ws = (float*)workspace;
  int took = 0;
  rc = launch_score_dense_tp(a, progs, (cudaStream_t)stream, &took);
  if (rc || took) return rc;
  rc = launch_score_stream(a, progs, (cudaStream_t)stream, &took);
  if (rc || took) return rc;
  rc = launch_score_warp(a, progs, (cudaStream_t)stream, &took);
  if (rc || took) return rc;
  ScorePlan plan = plan_score(progs, scenes, N, 0);
  PSTL_CHECK_ARG(plan.tp.smem_tape || workspace, "workspace required (see pstl_score_workspace_bytes)");
  return launch_score<false>(a, plan, (cudaStream_t)stream);
}

extern "C" int pstl_score_fused_bwd(pstl_program_t const* progs, const pstl_scene_view* scenes,
                                    const pstl_spec_params* sp, const float* mode, const float* state0,
                                    const float* controls, const float* ego_traj, int ego_stride, const float* stlp,
                                    int N, const float* grad_score, float* scores, float* grad_controls,
                                    float* grad_ego, void* workspace, pstl_stream_t stream) {
  int rc = check_score_args(progs, scenes, sp, mode, stlp, N);
  if (rc) return rc;
  PSTL_CHECK_ARG(grad_score, "grad_score required");
  PSTL_CHECK_ARG((controls && state0 && grad_controls && !ego_traj) || (ego_traj && ego_stride >= 4 && grad_ego),
                 "need (controls,state0,grad_controls) or (ego_traj,grad_ego)");
  if (N <= 0) return PSTL_OK;
  ScoreArgs a;
  base_args(a, progs, scenes, sp);
  a.mode = mode; a.state0 = state0; a.controls = ego_traj ? nullptr : controls; a.ego = ego_traj;
  a.ego_stride = ego_stride; a.stlp = stlp; a.N = N; a.C = 1;
  a.grad_score = grad_score; a.scores = scores;
  a.grad_controls = ego_traj ? nullptr : grad_controls; a.grad_ego = ego_traj ? grad_ego : nullptr;
  a.ws = (float*)workspace;
  int took = 0;
  rc = launch_score_stream_bwd(a, progs, (cudaStream_t)stream, &took);
  if (rc || took) return rc;
  ScorePlan plan = plan_score(progs, scenes, N, 1);
  PSTL_CHECK_ARG(plan.tp.smem_tape || workspace, "workspace required (see pstl_score_workspace_bytes)");
  return launch_score<true>(a, plan, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------------------
// guidance: gradient (fused above) + Adam / clip epilogue (nusc_train.py:606-626)
// --------------------------------------------------------------------------------------
__global__ void k_guidance_apply(const float* __restrict__ g, float* __restrict__ mu, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ anchor, size_t n, float step_size,
                                 float bc2_sqrt, float beta_t, int iter) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, single-tensor path
  const float b2 = 0.999f, eps = 1e-8f;
  const float w1 = (float)(1.0 - 0.9), w2 = (float)(1.0 - 0.999);
  const float gi = g[i];
  const float mi = m[i] + w1 * (gi - m[i]);                   // exp_avg.lerp_(grad, 1-beta1), weight < 0.5 branch
  const float vi = v[i] * b2 + (w2 * gi) * gi;                // exp_avg_sq.mul_(b2).addcmul_(g, g, value=1-b2)
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;             // (exp_avg_sq.sqrt()/bias_correction2_sqrt).add_(eps)
  float p = mu[i] + ((-step_size) * mi) / denom;              // param.addcdiv_(exp_avg, denom, value=-step_size)
  if (iter == 0) {
    // upstream aliases mu_init with the parameter: the first "clip" sees delta == 0 and the
    // once-stepped value becomes the anchor of every later clip
    anchor[i] = p;
  } else {
    const float a0 = anchor[i];
    const float d = fminf(fmaxf(fabsf(p - a0), -beta_t), beta_t);
    p = a0 + d;
  }
  mu[i] = p;
}

extern "C" int pstl_guidance_step(pstl_program_t const* progs, const pstl_scene_view* scenes,
                                  const pstl_spec_params* sp, const float* mode, const float* state0,
                                  const float* stlp, const float* valid, int N, float thres, float inv_norm,
                                  const float* inv_norm_dev, float lr, float beta_t, int iter, float* mu, float* adam_m,
                                  float* adam_v, float* mu_anchor, void* workspace, pstl_stream_t stream) {
  int rc = check_score_args(progs, scenes, sp, mode, stlp, N);
  if (rc) return rc;
  PSTL_CHECK_ARG(state0 && valid && mu && adam_m && adam_v && mu_anchor && workspace, "null argument");
  if (N <= 0) return PSTL_OK;
  const int T = scenes->T;
  // workspace layout: [grad (N,T,2)] [tape]
  float* grad = (float*)workspace;
  float* tape = grad + (size_t)N * T * 2;
  ScoreArgs a;
  base_args(a, progs, scenes, sp);
  a.mode = mode; a.state0 = state0; a.controls = mu; a.stlp = stlp; a.N = N; a.C = 1;
  a.valid = valid; a.thres = thres; a.inv_norm = inv_norm; a.inv_norm_dev = inv_norm_dev; a.grad_controls = grad;
  a.ws = tape;
  int took = 0;
  rc = launch_score_stream_bwd(a, progs, (cudaStream_t)stream, &took);
  if (rc) return rc;
  if (!took) {
    ScorePlan plan = plan_score(progs, scenes, N, 1);
    rc = launch_score<true>(a, plan, (cudaStream_t)stream);
    if (rc) return rc;
  }
  const size_t n = (size_t)N * T * 2;
  // torch computes the bias corrections in Python doubles, then applies them to fp32 tensors
  const double bc1 = 1.0 - pow(0.9, (double)(iter + 1)), bc2 = 1.0 - pow(0.999, (double)(iter + 1));
  k_guidance_apply<<<pstl_ceil_div((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(
      grad, mu, adam_m, adam_v, mu_anchor, n, (float)((double)lr / bc1), (float)sqrt(bc2), beta_t, iter);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

// --------------------------------------------------------------------------------------
// trajectory optimisation (nusc_train.py:287-316, 1303-1325): loss gradient (fused reverse-mode scorer) + control
// regulariser + one torch.optim.Adam step (single-tensor arithmetic, as k_guidance_apply) on the stored controls
// --------------------------------------------------------------------------------------
__global__ void k_trajopt_adam(const float* __restrict__ g, float* __restrict__ p, float* __restrict__ m,
                               float* __restrict__ v, size_t n, float step_size, float bc2_sqrt, float reg_over_numel,
                               float w_max2, float a_max2) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float b2 = 0.999f, eps = 1e-8f;
  const float w1 = (float)(1.0 - 0.9), w2 = (float)(1.0 - 0.999);
  const float x = p[i];
  // d/dx reg * mean(relu(x^2 - lim^2)): MeanBackward (grad / numel), relu mask, PowBackward (grad * (2 x))
  const float lim2 = (i & 1) ? a_max2 : w_max2;
  float gi = g[i];
  if (x * x - lim2 > 0.f) gi += reg_over_numel * (2.f * x);
  const float mi = m[i] + w1 * (gi - m[i]);
  const float vi = v[i] * b2 + (w2 * gi) * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = x + ((-step_size) * mi) / denom;
}

extern "C" int pstl_trajopt_step(pstl_program_t const* progs, const pstl_scene_view* scenes, const pstl_spec_params* sp,
                                 const float* mode, const float* state0, const float* stlp, const float* valid, int N,
                                 float thres, float inv_norm, float reg, float w_max, float a_max, float lr, int iter,
                                 float* params, float* adam_m, float* adam_v, float* scores, void* workspace,
                                 pstl_stream_t stream) {
  int rc = check_score_args(progs, scenes, sp, mode, stlp, N);
  if (rc) return rc;
  PSTL_CHECK_ARG(state0 && valid && params && adam_m && adam_v && workspace, "null argument");
  if (N <= 0) return PSTL_OK;
  const int T = scenes->T;
  float* grad = (float*)workspace;            // workspace: [grad (N,T,2)] [tape]
  float* tape = grad + (size_t)N * T * 2;
  ScoreArgs a;
  base_args(a, progs, scenes, sp);
  a.mode = mode; a.state0 = state0; a.controls = params; a.stlp = stlp; a.N = N; a.C = 1;
  a.valid = valid; a.thres = thres; a.inv_norm = inv_norm; a.grad_controls = grad; a.scores = scores; a.ws = tape;
  int took = 0;
  rc = launch_score_stream_bwd(a, progs, (cudaStream_t)stream, &took);
  if (rc) return rc;
  if (!took) {
    ScorePlan plan = plan_score(progs, scenes, N, 1);
    rc = launch_score<true>(a, plan, (cudaStream_t)stream);
    if (rc) return rc;
  }
  const size_t n = (size_t)N * T * 2;
  const double bc1 = 1.0 - pow(0.9, (double)(iter + 1)), bc2 = 1.0 - pow(0.999, (double)(iter + 1));
  const float numel = (float)((size_t)N * T);
  k_trajopt_adam<<<pstl_ceil_div((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(
      grad, params, adam_m, adam_v, n, (float)((double)lr / bc1), (float)sqrt(bc2), reg / numel, w_max * w_max, a_max * a_max);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}


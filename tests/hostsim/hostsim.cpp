// hostsim.cpp — TEST INFRASTRUCTURE ONLY.  Compiles the per-trajectory device math of the CUDA
// kernels (stl_core.cuh / drive_eval.cuh are __host__ __device__) with g++ so the CPU test tier can
// check it against the golden fixtures where no GPU exists.  The product never loads this file.
#include <vector>
#include <cstring>
#include "../../pstl-diffusion-policy_b200/csrc/drive_eval.cuh"
#include "../../pstl-diffusion-policy_b200/csrc/score_stream.cuh"

struct LeafHost {
  float signal(int, int) const { return 0.f; }
  void signal(int, int, float) const {}
  PstlIn pred_in(int, int) const { return PstlIn{nullptr, 1.f, 0.f, 1.f, 0}; }
  PstlOut pred_out(int, int) const { return PstlOut{nullptr, 0.f}; }
};

extern "C" int hs_stl_signals(const pstl_op* ops, int n_ops, int P, int T, int need_t, const float* sig, int N,
                              float tau, int hard, float* out_trace, const float* grad_trace, float* grad_sig) {
  PstlProgView pv;
  char err[256];
  if (pstl_resolve_program(ops, n_ops, P, T, need_t, &pv, err, sizeof(err))) { fprintf(stderr, "%s\n", err); return -1; }
  std::vector<float> vt(pv.val_floats), gt(pv.val_floats);
  LeafHost leaf;
  for (int n = 0; n < N; ++n) {
    for (int q = 0; q < P * T; ++q) vt[q] = sig[(size_t)n * P * T + q];
    pstl_interp_fwd(pv, vt.data(), 1, tau, hard, leaf);
    const int top = pv.ops[pv.n_ops - 1].out_off;
    if (out_trace) for (int t = 0; t < need_t; ++t) out_trace[(size_t)n * need_t + t] = vt[top + t];
    if (grad_trace && grad_sig) {
      std::fill(gt.begin(), gt.end(), 0.f);
      for (int t = 0; t < need_t; ++t) gt[top + t] += grad_trace[(size_t)n * need_t + t];
      pstl_interp_bwd(pv, vt.data(), gt.data(), 1, tau, hard, leaf, leaf);
      for (int q = 0; q < P * T; ++q) grad_sig[(size_t)n * P * T + q] = gt[q];
    }
  }
  return 0;
}

// dense rows (rows_per_scene = 1): neighbors (N,K,T,7), lanes 3 x (N,nseg,3)
extern "C" int hs_score(const pstl_op* const* ops3, const int* n_ops3, int T, int K, int nseg, const float* neighbors,
                        const float* l0, const float* l1, const float* l2, int rows_per_scene, const float* mode,
                        const float* state0, const float* controls, const float* ego, int es, const float* stlp, int N,
                        float dt, float tau, float w_scale, float a_scale, int clip_controls, int hard,
                        const float* grad_score, float* scores, float* grad_controls, float* grad_ego) {
  PstlProgView pv[3];
  char err[256];
  for (int k = 0; k < 3; ++k)
    if (pstl_resolve_program(ops3[k], n_ops3[k], 0, T, 1, &pv[k], err, sizeof(err))) { fprintf(stderr, "%s\n", err); return -1; }
  PstlEvalCfg c{dt, tau, 4.084f, 1.730f, w_scale, a_scale, clip_controls, 0, hard, nseg, K, T};
  const float* ln[3] = {l0, l1, l2};
  for (int n = 0; n < N; ++n) {
    const int m = (int)mode[n];
    if (m < 0 || m > 2) { scores[n] = (m == 3) ? 1.f : 0.f; continue; }
    const PstlProgView& P = pv[m];
    std::vector<float> tape(P.grad_floats);
    PstlSceneGlobal sg;
    const int scene = n / rows_per_scene;
    sg.neib = neighbors + (size_t)scene * K * T * 7;
    for (int l = 0; l < 3; ++l) sg.ln[l] = ln[l] + (size_t)scene * nseg * 3;
    sg.K = K; sg.T = T;
    PstlPose s0{0, 0, 0, 0};
    if (state0) s0 = PstlPose{state0[n * 4], state0[n * 4 + 1], state0[n * 4 + 2], state0[n * 4 + 3]};
    const float* u = controls ? controls + (size_t)n * T * 2 : nullptr;
    const float* e = ego ? ego + (size_t)n * T * es : nullptr;
    float* vt = tape.data();
    float* gt = vt + P.val_floats;
    float* pt = vt + P.part_off;
    scores[n] = pstl_eval_traj<PstlSceneGlobal, true, false>(P, sg, c, s0, u, e, es, stlp + (size_t)n * 6, vt, pt, 1);
    if (grad_score)
      pstl_eval_traj_bwd<false>(P, c, u, stlp + (size_t)n * 6, grad_score[n], vt, gt, pt, 1,
                         grad_controls ? grad_controls + (size_t)n * T * 2 : nullptr,
                         grad_ego ? grad_ego + (size_t)n * T * 4 : nullptr);
  }
  return 0;
}

// streaming (plan) scorer on the same inputs as hs_score (forward only).  Returns -2 if a program has no plan.
extern "C" int hs_score_stream(const pstl_op* const* ops3, const int* n_ops3, int T, int K, int nseg,
                               const float* neighbors, const float* l0, const float* l1, const float* l2,
                               int rows_per_scene, const float* mode, const float* state0, const float* controls,
                               const float* ego, int es, const float* stlp, int N, float dt, float tau, float w_scale,
                               float a_scale, int clip_controls, float* scores) {
  PstlProgView pv[3];
  PstlPlan pl[3];
  char err[256];
  for (int k = 0; k < 3; ++k) {
    if (pstl_resolve_program(ops3[k], n_ops3[k], 0, T, 1, &pv[k], err, sizeof(err))) { fprintf(stderr, "%s\n", err); return -1; }
    pstl_make_plan(pv[k], &pl[k]);
    if (!pl[k].valid) return -2;
  }
  PstlEvalCfg c{dt, tau, 4.084f, 1.730f, w_scale, a_scale, clip_controls, 0, 0, nseg, K, T};
  const float* ln[3] = {l0, l1, l2};
  std::vector<float> tape((size_t)PSTL_MAX_TAPES * T);
  for (int n = 0; n < N; ++n) {
    const int m = (int)mode[n];
    if (m < 0 || m > 2) { scores[n] = (m == 3) ? 1.f : 0.f; continue; }
    PstlStreamSceneGlobal sg;
    const int scene = n / rows_per_scene;
    sg.neib = neighbors + (size_t)scene * K * T * 7;
    for (int l = 0; l < 3; ++l) sg.ln[l] = ln[l] + (size_t)scene * nseg * 3;
    sg.K = K; sg.T = T; sg.ego_half = pstl_car_reach(c.ego_L, c.ego_W);
    PstlPose s0{0, 0, 0, 0};
    if (state0) s0 = PstlPose{state0[n * 4], state0[n * 4 + 1], state0[n * 4 + 2], state0[n * 4 + 3]};
    const float* u = controls ? controls + (size_t)n * T * 2 : nullptr;
    const float* e = ego ? ego + (size_t)n * T * es : nullptr;
    scores[n] = pstl_stream_eval(pl[m], sg, c, s0, u, e, es, stlp + (size_t)n * 6, tape.data(), 1);
  }
  return 0;
}

// plan recognition (stl_program.h): out = {valid, n_terms, listand, lane, n_tapes, need_pose, need_lane, need_nei, nei_term}
extern "C" int hs_plan_info(const pstl_op* ops, int n_ops, int T, int* out) {
  PstlProgView pv;
  PstlPlan pl;
  char err[256];
  if (pstl_resolve_program(ops, n_ops, 0, T, 1, &pv, err, sizeof(err))) { fprintf(stderr, "%s\n", err); return -1; }
  pstl_make_plan(pv, &pl);
  const int v[9] = {pl.valid, pl.n_terms, pl.listand, pl.lane, pl.n_tapes, pl.need_pose, pl.need_lane, pl.need_nei, pl.nei_term};
  for (int i = 0; i < 9; ++i) out[i] = v[i];
  return 0;
}

// streaming forward + reverse sweep (score_stream.cuh) on the same inputs as hs_score
extern "C" int hs_score_stream_grad(const pstl_op* const* ops3, const int* n_ops3, int T, int K, int nseg,
                                    const float* neighbors, const float* l0, const float* l1, const float* l2,
                                    int rows_per_scene, const float* mode, const float* state0, const float* controls,
                                    const float* ego, int es, const float* stlp, int N, float dt, float tau,
                                    float w_scale, float a_scale, int clip_controls, const float* grad_score,
                                    float* scores, float* grad_controls, float* grad_ego) {
  PstlProgView pv[3];
  PstlPlan pl[3];
  char err[256];
  for (int k = 0; k < 3; ++k) {
    if (pstl_resolve_program(ops3[k], n_ops3[k], 0, T, 1, &pv[k], err, sizeof(err))) { fprintf(stderr, "%s\n", err); return -1; }
    pstl_make_plan(pv[k], &pl[k]);
    if (!pl[k].valid) return -2;
  }
  PstlEvalCfg c{dt, tau, 4.084f, 1.730f, w_scale, a_scale, clip_controls, 0, 0, nseg, K, T};
  const float* ln[3] = {l0, l1, l2};
  std::vector<float> tape((size_t)pstl_stream_grad_floats(PSTL_MAX_TAPES, T));
  for (int n = 0; n < N; ++n) {
    const int m = (int)mode[n];
    float* gu = grad_controls ? grad_controls + (size_t)n * T * 2 : nullptr;
    float* ge = grad_ego ? grad_ego + (size_t)n * T * 4 : nullptr;
    if (m < 0 || m > 2) {
      scores[n] = (m == 3) ? 1.f : 0.f;
      if (gu) std::fill(gu, gu + 2 * T, 0.f);
      if (ge) std::fill(ge, ge + 4 * T, 0.f);
      continue;
    }
    PstlStreamSceneGlobal sg;
    const int scene = n / rows_per_scene;
    sg.neib = neighbors + (size_t)scene * K * T * 7;
    for (int l = 0; l < 3; ++l) sg.ln[l] = ln[l] + (size_t)scene * nseg * 3;
    sg.K = K; sg.T = T; sg.ego_half = pstl_car_reach(c.ego_L, c.ego_W);
    PstlPose s0{0, 0, 0, 0};
    if (state0) s0 = PstlPose{state0[n * 4], state0[n * 4 + 1], state0[n * 4 + 2], state0[n * 4 + 3]};
    const float* u = controls ? controls + (size_t)n * T * 2 : nullptr;
    const float* e = ego ? ego + (size_t)n * T * es : nullptr;
    PstlStreamAcc A;
    const float sc = pstl_stream_fwd<true>(pl[m], sg, c, s0, u, e, es, stlp + (size_t)n * 6, tape.data(), 1, A);
    scores[n] = sc;
    pstl_stream_bwd(pl[m], c, u, stlp + (size_t)n * 6, grad_score[n], sc != -INFINITY, A, tape.data(), 1, gu, ge);
  }
  return 0;
}

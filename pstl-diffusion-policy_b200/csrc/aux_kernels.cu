// aux_kernels.cu — standalone rollout (+adjoint) and predicate-signal kernels behind the
// reference's generate_trajs / prep_stl_cache entry points.  Compiled with -fmad=false.
#include "common.cuh"
#include "drive_eval.cuh"

__global__ void k_rollout(const float* __restrict__ s0, const float* __restrict__ u, int N, int T, float dt,
                          float* __restrict__ traj) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  PstlPose s{s0[n * 4], s0[n * 4 + 1], s0[n * 4 + 2], s0[n * 4 + 3]};
  float* o = traj + (size_t)n * (T + 1) * 4;
  const float* un = u + (size_t)n * T * 2;
  for (int t = 0; t < T; ++t) {
    o[t * 4] = s.x; o[t * 4 + 1] = s.y; o[t * 4 + 2] = s.th; o[t * 4 + 3] = s.v;
    s = pstl_unicycle_step(s, un[2 * t], un[2 * t + 1], dt, cosf(s.th), sinf(s.th));
  }
  o[T * 4] = s.x; o[T * 4 + 1] = s.y; o[T * 4 + 2] = s.th; o[T * 4 + 3] = s.v;
}

__global__ void k_rollout_bwd(const float* __restrict__ traj, const float* __restrict__ g, int N, int T, float dt,
                              float* __restrict__ gs0, float* __restrict__ gu) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* tr = traj + (size_t)n * (T + 1) * 4;
  const float* gg = g + (size_t)n * (T + 1) * 4;
  float ax = gg[T * 4], ay = gg[T * 4 + 1], ath = gg[T * 4 + 2], av = gg[T * 4 + 3];
  for (int t = T - 1; t >= 0; --t) {
    gu[((size_t)n * T + t) * 2] = ath * dt;
    gu[((size_t)n * T + t) * 2 + 1] = av * dt;
    const float th = tr[t * 4 + 2], v = tr[t * 4 + 3];
    const float cs = cosf(th), sn = sinf(th);
    const float nth = ath + ax * (-(v * sn) * dt) + ay * ((v * cs) * dt);
    const float nv = av + ax * (cs * dt) + ay * (sn * dt);
    ax += gg[t * 4]; ay += gg[t * 4 + 1]; ath = nth + gg[t * 4 + 2]; av = nv + gg[t * 4 + 3];
  }
  if (gs0) { gs0[n * 4] = ax; gs0[n * 4 + 1] = ay; gs0[n * 4 + 2] = ath; gs0[n * 4 + 3] = av; }
}

extern "C" int pstl_rollout(const float* state0, const float* controls, int N, int T, float dt, float* traj,
                            pstl_stream_t stream) {
  PSTL_CHECK_ARG(state0 && controls && traj && T > 0, "bad argument");
  if (N <= 0) return PSTL_OK;
  k_rollout<<<pstl_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(state0, controls, N, T, dt, traj);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

extern "C" int pstl_rollout_bwd(const float* traj, const float* grad_traj, int N, int T, float dt, float* grad_state0,
                                float* grad_controls, pstl_stream_t stream) {
  PSTL_CHECK_ARG(traj && grad_traj && grad_controls && T > 0, "bad argument");
  if (N <= 0) return PSTL_OK;
  k_rollout_bwd<<<pstl_ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(traj, grad_traj, N, T, dt, grad_state0,
                                                                        grad_controls);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

// one thread per (row, t)
__global__ void k_predicates(const float* __restrict__ nei, const float* __restrict__ l0, const float* __restrict__ l1,
                             const float* __restrict__ l2, int K, int nseg, int T, int rows_per_scene, float ego_L,
                             float ego_W, int clip_dist, const float* __restrict__ ego, int es, int N,
                             float* __restrict__ sig, float* __restrict__ part) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * T) return;
  const int n = (int)(i / T), t = (int)(i - (long long)n * T);
  const int scene = n / rows_per_scene;
  const float* e = ego + ((size_t)n * T + t) * es;
  const float px = e[0], py = e[1], pth = e[2];
  const float* ln[3] = {l0 + (size_t)scene * nseg * 3, l1 + (size_t)scene * nseg * 3, l2 + (size_t)scene * nseg * 3};
  for (int l = 0; l < 3; ++l) {
    float d, a, p[3];
    PstlLaneAcc acc{ln[l]};
    pstl_lane_pred(px, py, pth, acc, nseg, clip_dist, d, a, p);
    sig[((size_t)n * 7 + 2 * l) * T + t] = d;
    sig[((size_t)n * 7 + 2 * l + 1) * T + t] = a;
    if (part)
      for (int q = 0; q < 3; ++q) part[((size_t)n * 12 + 3 * l + q) * T + t] = p[q];
  }
  const float cs = cosf(pth), sn = sinf(pth);
  PstlCircles ec;
  pstl_car_circles(px, py, cs, sn, ego_L, ego_W, ec);
  PstlSceneGlobal sg;
  sg.neib = nei + (size_t)scene * K * T * 7;
  sg.K = K; sg.T = T;
  float best = INFINITY, bg[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < K; ++k) {
    PstlNei nb;
    float g[3];
    sg.nei(k, t, nb);
    const float term = pstl_pair_clearance(ec, cs, sn, nb, g);
    if (term < best) { best = term; bg[0] = g[0]; bg[1] = g[1]; bg[2] = g[2]; }
  }
  sig[((size_t)n * 7 + 6) * T + t] = best;
  if (part)
    for (int q = 0; q < 3; ++q) part[((size_t)n * 12 + 9 + q) * T + t] = bg[q];
}

extern "C" int pstl_predicates(const pstl_scene_view* sv, float ego_L, float ego_W, int clip_dist, const float* ego,
                               int ego_stride, int N, float* sig, float* part, pstl_stream_t stream) {
  PSTL_CHECK_ARG(sv && ego && sig && ego_stride >= 3, "bad argument");
  PSTL_CHECK_ARG(sv->rows_per_scene >= 1 && (long long)sv->n_scenes * sv->rows_per_scene >= N, "bad scene view");
  if (N <= 0) return PSTL_OK;
  const long long tot = (long long)N * sv->T;
  k_predicates<<<pstl_ceil_div(tot, 128), 128, 0, (cudaStream_t)stream>>>(
      sv->neighbors, sv->lanes[0], sv->lanes[1], sv->lanes[2], sv->Knei, sv->nseg, sv->T, sv->rows_per_scene, ego_L,
      ego_W, clip_dist, ego, ego_stride, N, sig, part);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

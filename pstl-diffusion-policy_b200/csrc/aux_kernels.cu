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

// Car-to-car distances for any (--refined_nL, --refined_nW) anchor grid and the extra signals of --collision_loss
// (utils.py:465-526 with full=True; nusc_train.py:81-83, 142-148).  One thread per (row, neighbour, step):
//   min_dist = min over the (nL*nW)^2 anchor pairs of the centre distance (first minimum, as torch.min)
//   rad_sum  = r_ego + r_nei,  r = min(max(L/nL/2, W/nW/2), W/2)
// part (N,K,T,3), optional: d min_dist / d (ego x, y, theta) through the arg-min pair (torch.norm's backward).
#define PSTL_MAX_ANCHORS 64
__device__ __forceinline__ void car_anchor(float x, float y, float c, float sn, float L, float W, int nL, int nW, int i,
                                           float r, float& ax, float& ay, float& qx, float& qy) {
  const int il = i / nW, iw = i - il * nW;
  const float al = pstl_linspace01(il, nL), be = pstl_linspace01(iw, nW);
  qx = (-L / 2.f + r) * (1.f - al) + (L / 2.f - r) * al;
  qy = (-(W / 2.f) + r) * (1.f - be) + (W / 2.f - r) * be;
  ax = qx * c - qy * sn + x;
  ay = qx * sn + qy * c + y;
}

__global__ void k_car_distances(const float* __restrict__ nei, int K, int T, int rows_per_scene, float ego_L, float ego_W,
                                int nL, int nW, const float* __restrict__ ego, int es, int N, float* __restrict__ min_dist,
                                float* __restrict__ rad_sum, float* __restrict__ part) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)N * K * T) return;
  const int t = (int)(i % T), k = (int)((i / T) % K), n = (int)(i / ((long long)T * K));
  const float* e = ego + ((size_t)n * T + t) * es;
  const float* p = nei + (((size_t)(n / rows_per_scene) * K + k) * T + t) * 7;
  const float ec = cosf(e[2]), esn = sinf(e[2]), nc = cosf(p[3]), nsn = sinf(p[3]);
  const float r1 = fminf(fmaxf(ego_L / (float)nL / 2.f, ego_W / (float)nW / 2.f), ego_W / 2.f);
  const float r2 = fminf(fmaxf(p[5] / (float)nL / 2.f, p[6] / (float)nW / 2.f), p[6] / 2.f);
  const int na = nL * nW;
  float best = INFINITY, bdx = 0.f, bdy = 0.f, bqx = 0.f, bqy = 0.f;
  for (int a = 0; a < na; ++a) {
    float ax, ay, qx, qy;
    car_anchor(e[0], e[1], ec, esn, ego_L, ego_W, nL, nW, a, r1, ax, ay, qx, qy);
    for (int b = 0; b < na; ++b) {
      float bx, by, ux, uy;
      car_anchor(p[1], p[2], nc, nsn, p[5], p[6], nL, nW, b, r2, bx, by, ux, uy);
      const float dx = ax - bx, dy = ay - by;
      const float d2 = dx * dx + dy * dy;
      if (d2 < best) { best = d2; bdx = dx; bdy = dy; bqx = qx; bqy = qy; }
    }
  }
  const float md = sqrtf(best);
  min_dist[i] = md;
  rad_sum[i] = r1 + r2;
  if (part) {
    const float ux = bdx / md, uy = bdy / md;  // NaN when the anchors coincide, as torch.norm's backward
    part[i * 3] = ux;
    part[i * 3 + 1] = uy;
    part[i * 3 + 2] = ux * (-bqx * esn - bqy * ec) + uy * (bqx * ec - bqy * esn);
  }
}

extern "C" int pstl_car_distances(const pstl_scene_view* sv, float ego_L, float ego_W, int nL, int nW, const float* ego,
                                  int ego_stride, int N, float* min_dist, float* rad_sum, float* part,
                                  pstl_stream_t stream) {
  PSTL_CHECK_ARG(sv && ego && min_dist && rad_sum && ego_stride >= 3, "bad argument");
  PSTL_CHECK_ARG(nL >= 1 && nW >= 1 && nL * nW <= PSTL_MAX_ANCHORS, "bad anchor grid");
  PSTL_CHECK_ARG(sv->rows_per_scene >= 1 && (long long)sv->n_scenes * sv->rows_per_scene >= N, "bad scene view");
  if (N <= 0) return PSTL_OK;
  const long long tot = (long long)N * sv->Knei * sv->T;
  k_car_distances<<<pstl_ceil_div(tot, 128), 128, 0, (cudaStream_t)stream>>>(
      sv->neighbors, sv->Knei, sv->T, sv->rows_per_scene, ego_L, ego_W, nL, nW, ego, ego_stride, N, min_dist, rad_sum,
      part);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

// --------------------------------------------------------------------------------------
// scene-encoder glue (Net.encode_feat, reference nusc_model.py:55-95): the ego-frame transform of every
// neighbour / lane point and the input layouts of the three encoder MLPs in ONE launch (upstream: ~90
// elementwise launches), and the neighbour pooling + feature concatenation in another.  Products and sums
// round separately, as the PyTorch expressions do (-fmad=false).
// --------------------------------------------------------------------------------------
// normalize_xyth (reference nusc_model.py:238-263); valid < 0 means "no valid factor"
__device__ __forceinline__ void enc_normalize(float x, float y, float th, float bx, float by, float bth, float valid,
                                              bool has_valid, float& xr, float& yr, float& tr) {
  const float xt = has_valid ? x - bx * valid : x - bx;
  const float yt = has_valid ? y - by * valid : y - by;
  const float c = cosf(bth), s = sinf(bth);
  xr = xt * c + yt * s;
  yr = (-xt) * s + yt * c;
  tr = has_valid ? th - bth * valid : th - bth;
}

// items per scene: [0] ego, [1, 1+K) neighbours, [1+K, 1+K+3*nseg) lane points
__global__ void k_encoder_inputs(const float* __restrict__ ego, int ego_stride, const float* __restrict__ nei,
                                 const float* __restrict__ l0, const float* __restrict__ l1, const float* __restrict__ l2,
                                 const float* __restrict__ id0, const float* __restrict__ id1,
                                 const float* __restrict__ id2, int n_scenes, int K, int nseg,
                                 float* __restrict__ ego_in, float* __restrict__ nei_in, float* __restrict__ lane_in) {
  const int per = 1 + K + 3 * nseg;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_scenes * per) return;
  const int b = (int)(i / per), it = (int)(i - (long long)b * per);
  const float* e = ego + (size_t)b * ego_stride;
  const float bx = e[0], by = e[1], bth = e[2];
  if (it == 0) {
    float xr, yr, tr;
    enc_normalize(e[0], e[1], e[2], bx, by, bth, 0.f, false, xr, yr, tr);
    float* o = ego_in + (size_t)b * 6;
    o[0] = xr; o[1] = yr; o[2] = tr; o[3] = e[3]; o[4] = e[4]; o[5] = e[5];
  } else if (it < 1 + K) {
    const int k = it - 1;
    const float* p = nei + ((size_t)b * K + k) * 7;
    float xr, yr, tr;
    enc_normalize(p[1], p[2], p[3], bx, by, bth, p[0], true, xr, yr, tr);
    float* o = nei_in + ((size_t)b * K + k) * 7;
    o[0] = p[0]; o[1] = xr; o[2] = yr; o[3] = tr; o[4] = p[4]; o[5] = p[5]; o[6] = p[6];
  } else {
    const int q = it - 1 - K, l = q / nseg, j = q - l * nseg;
    const float* lane = (l == 0 ? l0 : (l == 1 ? l1 : l2)) + (size_t)b * nseg * 3;
    const float valid = (l == 0 ? id0 : (l == 1 ? id1 : id2))[b];
    float xr, yr, tr;
    enc_normalize(lane[j * 3], lane[j * 3 + 1], lane[j * 3 + 2], bx, by, bth, valid, true, xr, yr, tr);
    if (j > 0) {  // points after the first enter as differences of the NORMALISED points
      float px, py, pt;
      enc_normalize(lane[(j - 1) * 3], lane[(j - 1) * 3 + 1], lane[(j - 1) * 3 + 2], bx, by, bth, valid, true, px, py, pt);
      xr = xr - px; yr = yr - py; tr = tr - pt;
    }
    float* o = lane_in + (((size_t)b * 3 + l) * nseg + j) * 3;
    o[0] = xr; o[1] = yr; o[2] = tr;
  }
}

extern "C" int pstl_encoder_inputs(const float* ego, int ego_row_stride, const float* neighbors, const float* lane_c,
                                   const float* lane_l, const float* lane_r, const float* id_c, const float* id_l,
                                   const float* id_r, int n_scenes, int Knei, int nseg, float* ego_in, float* nei_in,
                                   float* lane_in, pstl_stream_t stream) {
  PSTL_CHECK_ARG(ego && neighbors && lane_c && lane_l && lane_r && id_c && id_l && id_r && ego_in && nei_in && lane_in,
                 "null argument");
  PSTL_CHECK_ARG(ego_row_stride >= 6 && Knei >= 0 && nseg >= 1, "bad shape");
  if (n_scenes <= 0) return PSTL_OK;
  const long long tot = (long long)n_scenes * (1 + Knei + 3 * nseg);
  k_encoder_inputs<<<pstl_ceil_div(tot, 128), 128, 0, (cudaStream_t)stream>>>(
      ego, ego_row_stride, neighbors, lane_c, lane_l, lane_r, id_c, id_l, id_r, n_scenes, Knei, nseg, ego_in, nei_in, lane_in);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

// feature (bs, 7F) = [ego | min_k nei | mean_k nei | max_k nei | lane curr | left | right]  (nusc_model.py:88-93)
__global__ void k_encoder_pool(const float* __restrict__ ego_f, const float* __restrict__ nei_f,
                               const float* __restrict__ lane_f, int n_scenes, int K, int F, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_scenes * F) return;
  const int b = (int)(i / F), f = (int)(i - (long long)b * F);
  float mn = INFINITY, mx = -INFINITY, sum = 0.f;
  for (int k = 0; k < K; ++k) {
    const float v = nei_f[((size_t)b * K + k) * F + f];
    mn = fminf(mn, v); mx = fmaxf(mx, v); sum += v;
  }
  float* o = out + (size_t)b * 7 * F;
  o[f] = ego_f[(size_t)b * F + f];
  o[F + f] = mn;
  o[2 * F + f] = sum * (1.0f / (float)K);
  o[3 * F + f] = mx;
  for (int l = 0; l < 3; ++l) o[(4 + l) * F + f] = lane_f[((size_t)b * 3 + l) * F + f];
}

extern "C" int pstl_encoder_pool(const float* ego_feat, const float* nei_feat, const float* lane_feat, int n_scenes,
                                 int Knei, int F, float* feature, pstl_stream_t stream) {
  PSTL_CHECK_ARG(ego_feat && nei_feat && lane_feat && feature && Knei >= 1 && F >= 1, "bad argument");
  if (n_scenes <= 0) return PSTL_OK;
  k_encoder_pool<<<pstl_ceil_div((long long)n_scenes * F, 128), 128, 0, (cudaStream_t)stream>>>(
      ego_feat, nei_feat, lane_feat, n_scenes, Knei, F, feature);
  PSTL_LAUNCH_CHECK();
  return PSTL_OK;
}

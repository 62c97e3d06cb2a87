// score_warp.cuh — forward scoring with one WARP per trajectory and one lane per time step (T <= 32).
//
// The per-step predicates (3 lane searches, neighbour clearance) are independent across t and run one
// per lane; the Euler rollout is re-associated as prefix sums that each lane accumulates in the
// reference's own left-to-right order (so states are bit-identical to the sequential scan); the
// formula's temporal operators are warp reductions / Hillis-Steele scans of (max, sum-exp) pairs over
// shuffles.  No per-trajectory tape: a warp keeps one 32-lane slot per program op in shared memory,
// so occupancy is bounded by registers, not by shared memory (the thread-per-trajectory kernel k_score
// is kept for the reverse mode and for T > 32).
#pragma once
#include "drive_eval.cuh"

#define PSTL_WARP_ROWS 64   // trajectories per block (8 warps x 8)
#define PSTL_SOA_F 13       // staged floats per (neighbour, step): cx[4], cy[4], r, valid, centre x, y, L/2

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// (max, sum exp(x-max)) pairs: the associative form of logsumexp
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float M = fmaxf(m, m2);
  if (M == -INFINITY) { m = M; s = 0.f; return; }
  s = s * __expf(m - M) + s2 * __expf(m2 - M);
  m = M;
}

struct WBase {
  float v, d[3], th[3], nei;
};

__device__ __forceinline__ float wbase_sel(const WBase& b, int sid) {
  switch (sid) {
    case PSTL_SIG_V: return b.v;
    case PSTL_SIG_D_CURR: return b.d[0];
    case PSTL_SIG_TH_CURR: return b.th[0];
    case PSTL_SIG_D_LEFT: return b.d[1];
    case PSTL_SIG_TH_LEFT: return b.th[1];
    case PSTL_SIG_D_RIGHT: return b.d[2];
    case PSTL_SIG_TH_RIGHT: return b.th[2];
    default: return b.nei;
  }
}

// value of a typed predicate leaf at this lane's time step (called from ONE place: the kernel is
// instruction-cache bound if every consumer carries its own copy of the decode + IEEE division)
__device__ __forceinline__ float wpred(const WBase& b, int a0, int a1, const float* p) {
  const int sid = a0 & 0xff, pid = a1 & 0xff, den = (a1 >> 16) & 0xff;
  float x = wbase_sel(b, sid);
  if ((a0 >> 8) & 1) x = -x;
  float q = p[pid];
  if ((a1 >> 8) & 1) q = -q;
  float v = x + q;
  if (den != PSTL_DEN_ONE) v = v / pstl_pred_den(den, p);
  return v;
}

// value of op `idx` at this lane's time step
__device__ __forceinline__ float wget(const PstlProgView& P, int idx, const float* st, int lane, const WBase& b,
                                      const float* p) {
  return st[idx * 32 + lane];  // leaves were materialised by warp_interp's first pass
}

// inclusive scan of (max, sum) pairs towards lane 0 (dir = +1: suffix) or towards lane 31 (dir = -1: prefix)
__device__ __noinline__ void wscan(float& m, float& s, int lane, int dir, int hard) {
#pragma unroll 1
  for (int off = 1; off < 32; off <<= 1) {
    const int src = lane + dir * off;
    const float m2 = __shfl_sync(0xffffffffu, m, src & 31), s2 = __shfl_sync(0xffffffffu, s, src & 31);
    if (src >= 0 && src < 32) {
      if (hard) m = fmaxf(m, m2);
      else lse_merge(m, s, m2, s2);
    }
  }
}

// Interpret the program for one trajectory; returns the top-level robustness at t = 0 (all lanes).
__device__ __noinline__ float warp_interp(const PstlProgView& P, float* st, int lane, int T, const WBase& b,
                                          const float* p, float tau, int hard) {
  const float ts = hard ? 1.f : tau;  // scale applied before reductions
#pragma unroll 1
  for (int i = 0; i < P.n_ops; ++i) {
    const PstlROp o = P.ops[i];
    if (o.op == PSTL_OP_PRED) {  // a leaf is one word per lane here: materialise it
      st[i * 32 + lane] = wpred(b, o.a0, o.a1, p);
      continue;
    }
    const bool is_min = (o.op == PSTL_OP_SMIN2 || o.op == PSTL_OP_SMIN_K || o.op == PSTL_OP_WIN_SMIN ||
                         o.op == PSTL_OP_PREFIX_SMIN);
    const float sg = is_min ? -1.f : 1.f;
    float out;
    if (o.op == PSTL_OP_NEG) {
      out = -wget(P, o.in0, st, lane, b, p);
    } else if (o.op == PSTL_OP_SMIN2 || o.op == PSTL_OP_SMAX2 || o.op == PSTL_OP_SMIN_K) {
      // per-lane soft-min/max over k stacked children, accumulated as (max, sum) pairs
      const bool two = o.op != PSTL_OP_SMIN_K;
      const int k = two ? 2 : o.a0;
      float m = -INFINITY, s = 0.f;
#pragma unroll 1
      for (int j = 0; j < k; ++j) {
        const int src = two ? (j == 0 ? o.in0 : o.in1) : P.klist[o.a1 + j];
        const float x = sg * wget(P, src, st, lane, b, p) * ts;
        if (hard) m = fmaxf(m, x);
        else lse_merge(m, s, x, 1.f);
      }
      out = sg * (hard ? m : (__logf(s) + m) / tau);
    } else {
      // temporal operators: this lane's scaled input, then a reduction / scan across lanes
      const float mine = (lane < T) ? sg * wget(P, o.in0, st, lane, b, p) * ts : -INFINITY;
      float m = mine, s = (lane < T) ? 1.f : 0.f;
      bool empty = false;
      if (o.op == PSTL_OP_PREFIX_SMIN) {
        wscan(m, s, lane, -1, 0);
      } else if (o.op == PSTL_OP_SUFFIX_SMAX) {
        wscan(m, s, lane, +1, 0);
      } else if (o.n_out == 1) {
        // only t = 0 is consumed: one masked warp reduction over the window [clip(ts), clip(te))
        const int lo = pstl_clipi(o.a0, 0, T), hi = pstl_clipi(o.a1, 0, T);
        const bool in = lane >= lo && lane < hi;
        empty = hi <= lo;
        const float M = wmax(in ? mine : -INFINITY);
        const float mm = isinf(M) ? 0.f : M;
        s = hard ? 0.f : wsum(in ? __expf(mine - mm) : 0.f);
        m = hard ? M : mm;
      } else if (o.a0 == 0 && o.a1 >= T) {
        wscan(m, s, lane, +1, hard);  // suffix windows [t, T)
      } else {
        // general window [t+ts, t+te) clipped to [0,T): walk the window through shuffles
        const int lo = pstl_clipi(lane + o.a0, 0, T), hi = pstl_clipi(lane + o.a1, 0, T);
        int len = (lane < T) ? hi - lo : 0;
        empty = hi <= lo;
#pragma unroll
        for (int of = 16; of > 0; of >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, of));
        m = -INFINITY; s = 0.f;
#pragma unroll 1
        for (int d = 0; d < len; ++d) {
          const float x = __shfl_sync(0xffffffffu, mine, (lo + d) & 31);
          if (lo + d < hi) {
            if (hard) m = fmaxf(m, x);
            else lse_merge(m, s, x, 1.f);
          }
        }
      }
      // empty window: -inf for soft-min and soft-max alike (stl_d_lib.py:7-8,16-17)
      const bool scan_only = (o.op == PSTL_OP_PREFIX_SMIN || o.op == PSTL_OP_SUFFIX_SMAX);
      const float r = (hard && !scan_only) ? m : (__logf(s) + m) / tau;
      out = empty ? -INFINITY : sg * r;
    }
    st[i * 32 + lane] = out;
  }
  return __shfl_sync(0xffffffffu, st[(P.n_ops - 1) * 32 + lane], 0);
}

// scene accessors (SoA tile in shared memory, or this row's raw tensors in global memory)
struct WSceneSmem {
  const float* soa;  // [K][PSTL_SOA_F][T]
  const float* ln;   // [3][nseg][3]
  int K, T, nseg;
  __device__ float lane(int l, int j, int f) const { return ln[(l * nseg + j) * 3 + f]; }
  __device__ float f(int k, int fi, int t) const { return soa[(k * PSTL_SOA_F + fi) * T + t]; }
  __device__ void nei_meta(int k, int t, float& cx, float& cy, float& reach, float& valid) const {
    valid = f(k, 9, t); cx = f(k, 10, t); cy = f(k, 11, t); reach = f(k, 12, t);
  }
  __device__ void nei(int k, int t, PstlNei& o) const {
#pragma unroll
    for (int i = 0; i < 4; ++i) { o.cx[i] = f(k, i, t); o.cy[i] = f(k, 4 + i, t); }
    o.r = f(k, 8, t);
    o.valid = f(k, 9, t);
  }
};

__device__ void stage_scene_soa(const ScoreArgs& a, int scene, float* tile) {
  const PstlEvalCfg& c = a.cfg;
  float* ln = tile + (size_t)c.K * PSTL_SOA_F * c.T;
  const float* nb = a.neighbors + (size_t)scene * c.K * c.T * 7;
  for (int e = threadIdx.x; e < c.K * c.T; e += blockDim.x) {
    const int k = e / c.T, t = e - k * c.T;
    const float* p = nb + (size_t)e * 7;
    PstlCircles cc;
    pstl_car_circles(p[1], p[2], cosf(p[3]), sinf(p[3]), p[5], p[6], cc);
    float* o = tile + (size_t)k * PSTL_SOA_F * c.T + t;
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i * c.T] = cc.cx[i]; o[(4 + i) * c.T] = cc.cy[i]; }
    o[8 * c.T] = cc.r; o[9 * c.T] = p[0]; o[10 * c.T] = p[1]; o[11 * c.T] = p[2]; o[12 * c.T] = pstl_car_reach(p[5], p[6]);
  }
  for (int l = 0; l < 3; ++l) {
    const float* src = a.lanes[l] + (size_t)scene * c.nseg * 3;
    for (int e = threadIdx.x; e < c.nseg * 3; e += blockDim.x) ln[l * c.nseg * 3 + e] = src[e];
  }
}

template <class Scene>
__device__ __noinline__ void warp_predicates(const PstlProgView& P, const Scene& sc, const PstlEvalCfg& c,
                                                const PstlPose& s, float cs, float sn, int t, WBase& b) {
  b.v = s.v;
  b.nei = 0.f;
#pragma unroll
  for (int l = 0; l < 3; ++l) { b.d[l] = 0.f; b.th[l] = 0.f; }
#pragma unroll 1
  for (int l = 0; l < 3; ++l) {  // one copy of the lane search in the instruction stream
    const int sd = PSTL_SIG_D_CURR + 2 * l;
    if (t < P.base_need[sd] || t < P.base_need[sd + 1]) {
      struct L {
        const Scene* s; int l;
        __device__ float operator()(int j, int f) const { return s->lane(l, j, f); }
      } lacc{&sc, l};
      float d, th;
      pstl_lane_pred(s.x, s.y, s.th, lacc, c.nseg, c.clip_dist, d, th, nullptr);
      if (l == 0) { b.d[0] = d; b.th[0] = th; }
      else if (l == 1) { b.d[1] = d; b.th[1] = th; }
      else { b.d[2] = d; b.th[2] = th; }
    }
  }
  if (t < P.base_need[PSTL_SIG_NEI]) {
    PstlCircles e;
    pstl_car_circles(s.x, s.y, cs, sn, c.ego_L, c.ego_W, e);
    const float ego_half = pstl_car_reach(c.ego_L, c.ego_W);
    float best = INFINITY;
    for (int k = 0; k < c.K; ++k) {
      float ncx, ncy, reach, valid;
      sc.nei_meta(k, t, ncx, ncy, reach, valid);
      if (valid == 0.f) { best = fminf(best, 100.f); continue; }
      if (valid == 1.f && pstl_cull_neighbour(s.x - ncx, s.y - ncy, ego_half, reach, best)) {
        best = fminf(best, 20.f);
        continue;
      }
      PstlNei nb;
      sc.nei(k, t, nb);
      best = fminf(best, pstl_pair_clearance(e, cs, sn, nb, nullptr));
    }
    b.nei = best;
  }
}

// The three resolved programs travel as kernel parameters (constant bank): the interpreter's decode and
// branching then run on uniform loads / the uniform datapath instead of per-thread shared-memory reads.
struct WarpProgs {
  PstlProgView p[3];
};

template <bool SMEM_SCENE>
__global__ void __launch_bounds__(256, 3) k_score_warp(const __grid_constant__ ScoreArgs a, const __grid_constant__ WarpProgs wp) {
  extern __shared__ float sm[];
  const PstlProgView* progs = wp.p;
  const PstlEvalCfg c = a.cfg;
  const int T = c.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * PSTL_WARP_ROWS;
  float* tile = sm;
  size_t tile_f = 0;
  if (SMEM_SCENE) {
    stage_scene_soa(a, row0 / a.rows_per_scene, tile);
    tile_f = ((size_t)c.K * PSTL_SOA_F * T + (size_t)9 * c.nseg + 3) & ~(size_t)3;
  }
  float* st = sm + tile_f + (size_t)warp * a.F * 32;  // a.F = slots per warp (max n_ops)
  __syncthreads();
  WSceneSmem ss{tile, tile + (size_t)c.K * PSTL_SOA_F * T, c.K, T, c.nseg};

  for (int r = warp; r < PSTL_WARP_ROWS; r += 8) {
    const int n = row0 + r;
    if (n >= a.N) break;
    const float md = a.mode[n];
    const int m = (md == 0.f) ? 0 : (md == 1.f) ? 1 : (md == 2.f) ? 2 : (md == 3.f) ? 3 : 4;
    const PstlProgView& P = progs[m < 3 ? m : 0];
    float p[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) p[i] = a.stlp[(size_t)n * 6 + i];
    const int scene = n / a.rows_per_scene;
    PstlSceneGlobal sg;
    sg.neib = a.neighbors + (size_t)scene * c.K * T * 7;
    for (int l = 0; l < 3; ++l) sg.ln[l] = a.lanes[l] + (size_t)scene * c.nseg * 3;
    sg.K = c.K; sg.T = T;
    PstlPose s0{0.f, 0.f, 0.f, 0.f};
    if (a.state0) { s0.x = a.state0[n * 4]; s0.y = a.state0[n * 4 + 1]; s0.th = a.state0[n * 4 + 2]; s0.v = a.state0[n * 4 + 3]; }

    float best = -INFINITY;
    int bi = 0;
    float bw = 0.f, ba = 0.f;
    PstlPose bs = s0;
    for (int cand = 0; cand < a.C; ++cand) {
      PstlPose s = s0;
      float w = 0.f, ac = 0.f, cs, sn;
      if (a.ego) {
        if (lane < T) {
          const float* e = a.ego + ((size_t)n * T + lane) * a.ego_stride;
          s.x = e[0]; s.y = e[1]; s.th = e[2]; s.v = e[3];
        }
        sincosf(s.th, &sn, &cs);
      } else {
        // lane t holds control t; states are prefix sums accumulated in the reference's order
        if (lane < T) {
          const float2 u = *reinterpret_cast<const float2*>(a.controls + (((size_t)cand * a.N + n) * T + lane) * 2);
          w = u.x * c.w_scale;
          ac = u.y * c.a_scale;
          if (c.clip_controls) {
            w = fminf(fmaxf(w, -c.w_scale), c.w_scale);
            ac = fminf(fmaxf(ac, -c.a_scale), c.a_scale);
          }
        }
        const float ith = w * c.dt, iv = ac * c.dt;
        for (int j = 0; j < T; ++j) {
          const float a_th = __shfl_sync(0xffffffffu, ith, j), a_v = __shfl_sync(0xffffffffu, iv, j);
          if (j < lane) { s.th = s.th + a_th; s.v = s.v + a_v; }
        }
        sincosf(s.th, &sn, &cs);
        const float ix = (s.v * cs) * c.dt, iy = (s.v * sn) * c.dt;
        for (int j = 0; j < T; ++j) {
          const float a_x = __shfl_sync(0xffffffffu, ix, j), a_y = __shfl_sync(0xffffffffu, iy, j);
          if (j < lane) { s.x = s.x + a_x; s.y = s.y + a_y; }
        }
      }
      float sc;
      if (m < 3) {
        WBase b;
        if (lane < T) {
          if (SMEM_SCENE) warp_predicates(P, ss, c, s, cs, sn, lane, b);
          else warp_predicates(P, sg, c, s, cs, sn, lane, b);
        } else {
          b.v = 0.f; b.nei = 0.f;
          for (int l = 0; l < 3; ++l) { b.d[l] = 0.f; b.th[l] = 0.f; }
        }
        sc = warp_interp(P, st, lane, T, b, p, c.tau, c.hard);
      } else {
        sc = (m == 3) ? 1.0f : 0.0f;
      }
      if (lane == 0 && a.scores_all) a.scores_all[(size_t)cand * a.N + n] = sc;
      if (cand == 0 || sc > best) {  // torch.max(dim=0): first maximum
        best = sc; bi = cand; bw = w; ba = ac; bs = s;
      }
    }
    if (lane == 0) {
      if (a.best_score) a.best_score[n] = best;
      if (a.best_idx) a.best_idx[n] = bi;
    }
    if (a.controls) {
      if (a.best_controls && lane < T)
        *reinterpret_cast<float2*>(a.best_controls + ((size_t)n * T + lane) * 2) = make_float2(bw, ba);
      if (a.traj_out && lane <= T && lane < 32) {
        float* o = a.traj_out + ((size_t)n * (T + 1) + lane) * 4;
        *reinterpret_cast<float4*>(o) = make_float4(bs.x, bs.y, bs.th, bs.v);
      }
    }
  }
}

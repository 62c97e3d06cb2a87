// drive_eval.cuh — one trajectory through rollout -> predicates -> STL program (and back).
// Replaces, per row: generate_trajs (nusc_train.py:39-49), prep_stl_cache (:74-93),
// the formula evaluation of compute_stl_dense (:318-323) and its autograd graph.
#pragma once
#include "drive_core.cuh"
#include "stl_program.h"

// Scene accessor concept:
//   float lane(int l, int j, int f)                      lane l point j component f
//   void  nei_meta(int k, int t, float& cx, float& cy, float& reach, float& valid)   car centre, L/2
//   void  nei(int k, int t, PstlNei& out)                circle centres, radius, valid
struct PstlSceneGlobal {
  const float* neib;  // (K,T,7) of this row's scene
  const float* ln[3];
  int K, T;
  PSTL_HD float lane(int l, int j, int f) const { return ln[l][j * 3 + f]; }
  PSTL_HD void nei_meta(int k, int t, float& cx, float& cy, float& reach, float& valid) const {
    const float* p = neib + ((size_t)k * T + t) * 7;
    valid = p[0]; cx = p[1]; cy = p[2]; reach = pstl_car_reach(p[5], p[6]);
  }
  PSTL_HD void nei(int k, int t, PstlNei& out) const {
    const float* p = neib + ((size_t)k * T + t) * 7;
    PstlCircles c;
    pstl_car_circles(p[1], p[2], cosf(p[3]), sinf(p[3]), p[5], p[6], c);
#pragma unroll
    for (int i = 0; i < PSTL_NL; ++i) { out.cx[i] = c.cx[i]; out.cy[i] = c.cy[i]; }
    out.r = c.r;
    out.valid = p[0];
  }
};

struct PstlLaneAcc {
  const float* p;
  PSTL_HD float operator()(int j, int f) const { return p[j * 3 + f]; }
};

struct PstlEvalCfg {
  float dt, tau, ego_L, ego_W, w_scale, a_scale;
  int clip_controls, clip_dist, hard, nseg, K, T;
};

struct PstlLeafFused {
  const PstlProgView* P;
  const float* vt;
  int stride;
  float p[6];  // this row's pSTL parameters, read once
  PSTL_HD PstlLeafFused(const PstlProgView* P_, const float* vt_, int stride_, const float* stlp)
      : P(P_), vt(vt_), stride(stride_) {
    for (int i = 0; i < 6; ++i) p[i] = stlp[i];
  }
  PSTL_HD float signal(int, int) const { return 0.f; }
  // value[t] = (+-base[sid][t] +- stlp[pid]) / den   (include/pstl.h, PSTL_OP_PRED)
  PSTL_HD PstlIn pred_in(int a0, int a1) const {
    const int sid = a0 & 0xff, pid = a1 & 0xff, den = (a1 >> 16) & 0xff;
    PstlIn r;
    r.p = vt + (size_t)P->base_off[sid] * stride;
    r.sb = ((a0 >> 8) & 1) ? -1.f : 1.f;
    r.pp = ((a1 >> 8) & 1) ? -p[pid] : p[pid];
    r.den = (den == PSTL_DEN_ONE) ? 1.f : pstl_pred_den(den, p);
    r.pred = 1;
    return r;
  }
};

struct PstlLeafFusedGrad {
  const PstlProgView* P;
  float* gt;
  int stride;
  const float* p;
  PSTL_HD void signal(int, int, float) const {}
  PSTL_HD PstlOut pred_out(int a0, int a1) const {
    const int sid = a0 & 0xff, den = (a1 >> 16) & 0xff;
    PstlOut r;
    r.g = gt + (size_t)P->base_off[sid] * stride;
    r.w = ((a0 >> 8) & 1) ? -1.f : 1.f;
    if (den != PSTL_DEN_ONE) r.w = r.w / pstl_pred_den(den, p);
    return r;
  }
};

// Controls source: either candidate controls (pre-scale) or nothing (ego_traj given).
PSTL_HD void pstl_scaled_control(const float* u, int t, const PstlEvalCfg& c, float& w, float& a) {
  w = u[2 * t] * c.w_scale;
  a = u[2 * t + 1] * c.a_scale;
  if (c.clip_controls) {
    w = fminf(fmaxf(w, -c.w_scale), c.w_scale);
    a = fminf(fmaxf(a, -c.a_scale), c.a_scale);
  }
}

PSTL_HD int pstl_need_pose(const PstlProgView& P) {
  int n = 0;
  for (int b = 0; b < PSTL_N_BASE_SIGNALS; ++b) n = P.base_need[b] > n ? P.base_need[b] : n;
  return n;
}

// Forward: fills the value tape (and the partial block pt when GRAD) and returns the score.
// u: controls of this row (T*2 floats, pre-scale) or nullptr; ego: pre-rolled states (stride es) or nullptr.
template <class Scene, bool GRAD, bool FAST>
PSTL_HD float pstl_eval_traj(const PstlProgView& P, const Scene& sc, const PstlEvalCfg& c, PstlPose s,
                             const float* u, const float* ego, int es, const float* stlp, float* vt, float* pt,
                             int stride) {
  const int T = c.T;
  const int need_pose = pstl_need_pose(P);
  const float ego_half = pstl_car_reach(c.ego_L, c.ego_W);
#define VT(off) vt[(size_t)(off) * stride]
#define PT(row, t) pt[(size_t)((row) * T + (t)) * stride]
  for (int t = 0; t < need_pose; ++t) {
    if (ego) {
      s.x = ego[t * es + 0]; s.y = ego[t * es + 1]; s.th = ego[t * es + 2]; s.v = ego[t * es + 3];
    }
    const float cs = cosf(s.th), sn = sinf(s.th);
    if (GRAD) { PT(0, t) = cs; PT(1, t) = sn; PT(2, t) = s.v; }
    if (t < P.base_need[PSTL_SIG_V]) VT(P.base_off[PSTL_SIG_V] + t) = s.v;
    for (int l = 0; l < 3; ++l) {
      const int sd = PSTL_SIG_D_CURR + 2 * l, sa = sd + 1;
      if (t < P.base_need[sd] || t < P.base_need[sa]) {
        float d, a, part[3];
        struct L {
          const Scene* s; int l;
          PSTL_HD float operator()(int j, int f) const { return s->lane(l, j, f); }
        } lacc{&sc, l};
        pstl_lane_pred(s.x, s.y, s.th, lacc, c.nseg, c.clip_dist, d, a, GRAD ? part : nullptr);
        if (t < P.base_need[sd]) VT(P.base_off[sd] + t) = d;
        if (t < P.base_need[sa]) VT(P.base_off[sa] + t) = a;
        if (GRAD) { PT(3 + 3 * l, t) = part[0]; PT(4 + 3 * l, t) = part[1]; PT(5 + 3 * l, t) = part[2]; }
      }
    }
    if (t < P.base_need[PSTL_SIG_NEI]) {
      PstlCircles e;
      pstl_car_circles(s.x, s.y, cs, sn, c.ego_L, c.ego_W, e);
      float best = INFINITY, bg0 = 0.f, bg1 = 0.f, bg2 = 0.f;
      for (int k = 0; k < c.K; ++k) {
        float ncx, ncy, reach, valid;
        sc.nei_meta(k, t, ncx, ncy, reach, valid);
        if (valid == 0.f) {  // clip(d)*0 + (1-0)*100 (zero-padded rows, nusc_api.py:615,639)
          if (100.f < best) { best = 100.f; bg0 = bg1 = bg2 = 0.f; }
          continue;
        }
        if (valid == 1.f && pstl_cull_neighbour(s.x - ncx, s.y - ncy, ego_half, reach, best)) {
          if (20.f < best) { best = 20.f; bg0 = bg1 = bg2 = 0.f; }  // clipped at 20, zero gradient
          continue;
        }
        PstlNei nb;
        sc.nei(k, t, nb);
        float g[3];
        const float term = pstl_pair_clearance(e, cs, sn, nb, GRAD ? g : nullptr);
        if (term < best) {  // torch.min(dim=1): first minimal index
          best = term;
          if (GRAD) { bg0 = g[0]; bg1 = g[1]; bg2 = g[2]; }
        }
      }
      VT(P.base_off[PSTL_SIG_NEI] + t) = best;
      if (GRAD) { PT(12, t) = bg0; PT(13, t) = bg1; PT(14, t) = bg2; }
    }
    if (!ego && t + 1 < need_pose) {
      float w, a;
      pstl_scaled_control(u, t, c, w, a);
      s = pstl_unicycle_step(s, w, a, c.dt, cs, sn);
    }
  }
  PstlLeafFused leaf(&P, vt, stride, stlp);
  pstl_interp_fwd<FAST>(P, vt, stride, c.tau, c.hard, leaf);
  return VT(P.ops[P.n_ops - 1].out_off);
#undef VT
#undef PT
}

// Reverse: after pstl_eval_traj<.., true, ..>.  gscore = d loss / d score.  Writes d loss / d controls
// (pre-scale, T*2 floats, row stride 1) when gu != nullptr, d loss / d ego (T*4) when ge != nullptr.
template <bool FAST>
PSTL_HD void pstl_eval_traj_bwd(const PstlProgView& P, const PstlEvalCfg& c, const float* u, const float* stlp,
                                float gscore, const float* vt, float* gt, const float* pt, int stride, float* gu,
                                float* ge) {
  const int T = c.T;
#define GT(off) gt[(size_t)(off) * stride]
#define PT(row, t) pt[(size_t)((row) * T + (t)) * stride]
  for (int i = 0; i < P.val_floats; ++i) GT(i) = 0.f;
  GT(P.ops[P.n_ops - 1].out_off) = gscore;
  PstlLeafFused leaf(&P, vt, stride, stlp);
  PstlLeafFusedGrad lg{&P, gt, stride, leaf.p};
  pstl_interp_bwd<FAST>(P, vt, gt, stride, c.tau, c.hard, leaf, lg);
  // adjoint of the pose at every step from the base-signal adjoints
  float ax = 0.f, ay = 0.f, ath = 0.f, av = 0.f;  // adjoint of s_{t+1} accumulated so far
  const int need_pose = pstl_need_pose(P);
  for (int t = need_pose; t < T; ++t) {  // poses the formula never reads
    if (ge) { ge[t * 4 + 0] = 0.f; ge[t * 4 + 1] = 0.f; ge[t * 4 + 2] = 0.f; ge[t * 4 + 3] = 0.f; }
    if (gu) { gu[2 * t] = 0.f; gu[2 * t + 1] = 0.f; }
  }
  for (int t = need_pose - 1; t >= 0; --t) {
    float lx = 0.f, ly = 0.f, lth = 0.f, lv = 0.f;
    if (t < P.base_need[PSTL_SIG_V]) lv += GT(P.base_off[PSTL_SIG_V] + t);
    for (int l = 0; l < 3; ++l) {
      const int sd = PSTL_SIG_D_CURR + 2 * l, sa = sd + 1;
      if (t < P.base_need[sd]) {
        const float g = GT(P.base_off[sd] + t);
        lx += g * PT(3 + 3 * l, t);
        ly += g * PT(4 + 3 * l, t);
      }
      if (t < P.base_need[sa]) lth += GT(P.base_off[sa] + t) * PT(5 + 3 * l, t);
    }
    if (t < P.base_need[PSTL_SIG_NEI]) {
      const float g = GT(P.base_off[PSTL_SIG_NEI] + t);
      if (g != 0.f) { lx += g * PT(12, t); ly += g * PT(13, t); lth += g * PT(14, t); }
    }
    if (ge) { ge[t * 4 + 0] = lx; ge[t * 4 + 1] = ly; ge[t * 4 + 2] = lth; ge[t * 4 + 3] = lv; }
    if (gu) {
      // controls at step t move s_{t+1}: d th_{t+1}/d w_t = dt, d v_{t+1}/d a_t = dt
      float gw = ath * c.dt * c.w_scale, ga = av * c.dt * c.a_scale;
      if (c.clip_controls) {
        const float w = u[2 * t] * c.w_scale, a = u[2 * t + 1] * c.a_scale;
        if (w < -c.w_scale || w > c.w_scale) gw = 0.f;
        if (a < -c.a_scale || a > c.a_scale) ga = 0.f;
      }
      gu[2 * t] = gw;
      gu[2 * t + 1] = ga;
      // A_t = lambda_t + J_t^T A_{t+1}
      const float cs = PT(0, t), sn = PT(1, t), v = PT(2, t);
      const float nth = ath + ax * (-(v * sn) * c.dt) + ay * ((v * cs) * c.dt);
      const float nv = av + ax * (cs * c.dt) + ay * (sn * c.dt);
      ax += lx; ay += ly; ath = nth + lth; av = nv + lv;
    }
  }
#undef GT
#undef PT
}

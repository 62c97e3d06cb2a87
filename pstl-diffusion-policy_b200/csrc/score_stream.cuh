// score_stream.cuh — streaming forward scorer: ONE thread per trajectory, ONE pass over the rollout.
//
// For programs with a PstlPlan (stl_program.h: [ListAnd of] R1_[lo,hi) [R2_suffix] X) the robustness at
// t = 0 needs no trace tape: while the Euler rollout advances, every term keeps an online log-sum-exp
// accumulator (max, sum) in two registers; only the terms with an inner suffix operator park X(t) in a
// shared-memory column and are folded by one backward sweep after the rollout.  All threads of a block
// are at the same time step of the same scene, so every scene read (lane points, neighbour circles) is a
// warp-uniform shared-memory broadcast, and warps are dealt rows of one formula (rows cycle through the
// three modes), so term decoding is uniform too.
//
// Replaces, per row: generate_trajs (nusc_train.py:39-49), prep_stl_cache (:74-93), the three formula
// calls of compute_stl_dense (:318-323) and the best-of-K max/gather (:992-1007).
//
// Soft reductions run in the base-2 domain: x2 = x * (+-tau * log2 e), e = ex2(-|x2 - m|), result
// (lg2(s) + m) * ln2 / tau.  Mathematically torch.logsumexp; rounding differs by a few ulp of the result.
#pragma once
#include "drive_eval.cuh"

#define PSTL_STREAM_NEI_F4 4  // float4 per staged (neighbour, step): cx[4] | cy[4] | valid, centre x, y, L/2 | r

PSTL_HD float pstl_ex2(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return exp2f(x);
#endif
}
PSTL_HD float pstl_lg2(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return log2f(x);
#endif
}

#define PSTL_LSE2_INIT (-3.402823466e+38f)

// online log-sum-exp in base 2: exactly one of the two rescale factors is 1, so one ex2 per element
PSTL_HD void pstl_lse2_add(float& m, float& s, float x) {
  const float d = x - m;
  const float e = pstl_ex2(-fabsf(d));
  s = (d > 0.f) ? s * e + 1.f : s + e;
  m = fmaxf(m, x);
}

struct PstlF4 {
  float x, y, z, w;
};

// scene accessors of the streaming scorer
//   lane_pt(l, j)                         lane l point j as (x, y, theta, -)
//   nei_begin(t, count, init)             neighbours to visit at step t and the initial running minimum
//                                         (100 when a zero-valid neighbour was dropped from the list)
//   nei_meta(slot, t, valid, cx, cy, rsum) car centre and  reach_ego + reach_nei + 1e-3  (cull radius less min(best,20))
//   nei(slot, t, out)                     circle centres, radius, valid
struct PstlStreamSceneGlobal {  // raw tensors of this row's scene
  const float* neib;            // (K,T,7)
  const float* ln[3];           // (nseg,3)
  int K, T;
  float ego_half;
  PSTL_HD PstlF4 lane_pt(int l, int j) const {
    const float* p = ln[l] + j * 3;
    return PstlF4{p[0], p[1], p[2], 0.f};
  }
  PSTL_HD PstlF4 lane_xy(int l, int j) const { return lane_pt(l, j); }  // (x, y) for the segment search
  PSTL_HD void nei_begin(int, int& count, float& init) const { count = K; init = INFINITY; }
  PSTL_HD void nei_meta(int k, int t, float& valid, float& cx, float& cy, float& rsum) const {
    const float* p = neib + ((size_t)k * T + t) * 7;
    valid = p[0]; cx = p[1]; cy = p[2]; rsum = ego_half + pstl_car_reach(p[5], p[6]) + 1e-3f;
  }
  PSTL_HD void nei(int k, int t, PstlNei& out) const {
    const float* p = neib + ((size_t)k * T + t) * 7;
    PstlCircles c;
    pstl_car_circles(p[1], p[2], cosf(p[3]), sinf(p[3]), p[5], p[6], c);
    for (int i = 0; i < PSTL_NL; ++i) { out.cx[i] = c.cx[i]; out.cy[i] = c.cy[i]; }
    out.r = c.r;
    out.valid = p[0];
  }
};

// Exact cull (see pstl_cull_neighbour): rsum = reach_ego + reach_nei + margin (pstl_car_reach)
PSTL_HD bool pstl_cull_neighbour_r(float dx, float dy, float rsum, float best) {
  const float R = fminf(best, 20.f) + rsum;
  return R > 0.f && fmaf(dx, dx, dy * dy) >= R * R;
}

// value of a typed leaf (include/pstl.h, PSTL_OP_PRED): (sb*base + sp*stlp[pid]) / den
PSTL_HD float pstl_plan_leaf(const PstlLeafC& l, float v, float d, float th, float nei, const float* p) {
  const float base = (l.c == 0) ? v : (l.c == 1) ? d : (l.c == 2) ? th : nei;
  float x = l.sb * base + l.sp * p[l.pid];  // signs are +-1: exact
  if (l.den != PSTL_DEN_ONE) x = x / pstl_pred_den(l.den, p);
  return x;
}

// One trajectory through rollout -> predicates -> plan.  u: this row's controls (T*2, pre-scale) or null;
// ego: pre-rolled states (stride es) or null; p: this row's six pSTL parameters; tape: this row's column
// base, element (col, t) at tape[(col * T + t) * tstride].
//
// A single-leaf term  R_t (sb*b_t + q)/den  is shift-invariant:  = q/den + R_t (sb*b_t/den), so inside the
// time loop such a term costs one multiply by its per-row factor g = +-tau*log2(e)/den and one online
// accumulator update; q/den is added once after the loop.  Two-leaf (And/Or) terms evaluate X(t) in full.
//
// GRAD: the pass also records, per step, what the reverse sweep (pstl_stream_bwd) needs — cos, sin, v, the three
// other base signals and the partial derivatives of the lane / clearance predicates — in tape columns 0..11; the
// X(t) column of an inner-operator term moves to 12 + 2*tm.tape and the next column receives
// L_t = lse2_{t' >= t}(inner * X2(t')) for t in [lo, hi).
#define PSTL_STREAM_GRAD_COLS 12
PSTL_HD int pstl_stream_xcol(const PstlTerm& tm, bool grad) { return grad ? PSTL_STREAM_GRAD_COLS + 2 * tm.tape : tm.tape; }

struct PstlStreamAcc {
  float am[PSTL_MAX_TERMS], as[PSTL_MAX_TERMS], g2[PSTL_MAX_TERMS];
  float top_m, top_s;
  float rg_nei;  // 1 / g2 of the clearance term (negative), see the value-aware bound
};

#define PSTL_TAPE(col, t) tape[(size_t)((col) * T + (t)) * tstride]

// The pass is split in three so a kernel can walk the horizon in chunks (scene tile re-staged per chunk):
// init -> steps(t0, t1) ... -> finish.  pstl_stream_fwd below is the one-shot composition.
PSTL_HD void pstl_stream_init(const PstlPlan& pl, const PstlEvalCfg& c, const float* p, PstlStreamAcc& A) {
  const float k2 = c.tau * 1.4426950408889634f;   // tau * log2(e)
  float (&am)[PSTL_MAX_TERMS] = A.am;
  float (&as)[PSTL_MAX_TERMS] = A.as;
  float (&g2)[PSTL_MAX_TERMS] = A.g2;
#pragma unroll
  for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
    am[k] = PSTL_LSE2_INIT; as[k] = 0.f; g2[k] = 0.f;
    if (k < pl.n_terms) {
      const PstlTerm& tm = pl.terms[k];
      if (tm.pair == 0) {
        // single leaf: accumulate (or tape) b_t * g2 ; the operator sign joins g2 only when there is no inner operator
        float g = tm.a.sb * k2;
        if (tm.a.den != PSTL_DEN_ONE) g = g / pstl_pred_den(tm.a.den, p);
        g2[k] = (tm.inner == 0) ? g * (float)tm.outer : g;
      }
    }
  }

  A.rg_nei = (pl.nei_term == 0) ? 1.f / g2[0] : 0.f;
}

// steps t0 <= t < t1 (t1 <= pl.need_pose); the accessor sc serves exactly those steps; s is the pose at t0 on entry and
// at t1 on exit
template <bool GRAD, class Scene>
PSTL_HD void pstl_stream_steps(const PstlPlan& pl, const Scene& sc, const PstlEvalCfg& c, PstlPose& s, const float* u,
                               const float* ego, int es, const float* p, float* tape, int tstride, PstlStreamAcc& A,
                               int t0, int t1) {
  const int T = c.T;
  const float k2 = c.tau * 1.4426950408889634f;   // tau * log2(e)
  float (&am)[PSTL_MAX_TERMS] = A.am;
  float (&as)[PSTL_MAX_TERMS] = A.as;
  float (&g2)[PSTL_MAX_TERMS] = A.g2;
  const float rg_nei = A.rg_nei;
#pragma unroll 1
  for (int t = t0; t < t1; ++t) {
    if (ego) {
      const float* e = ego + (size_t)t * es;
      s.x = e[0]; s.y = e[1]; s.th = e[2]; s.v = e[3];
    }
    // the controls that take the pose to t+1 are fetched now and used at the bottom of the step: the load (one sector per
    // lane, rows are 8T bytes apart) then has the whole step to land (it was 5.5 % of the launch's stall samples)
    const bool advance = !ego && t + 1 < pl.need_pose;
    float ut[2] = {0.f, 0.f};
    if (advance) { ut[0] = u[2 * t]; ut[1] = u[2 * t + 1]; }
    float sn, cs;
#if defined(__CUDA_ARCH__)
    sincosf(s.th, &sn, &cs);
#else
    sn = sinf(s.th); cs = cosf(s.th);
#endif
    float d = 0.f, th = 0.f, nei = 0.f;
    if (t < pl.need_lane) {
      // nusc_api.py:693-712: closest segment = first arg-min of d_j + d_{j+1}
      const int l = pl.lane;
      PstlF4 q = sc.lane_xy(l, 0);
      float dx = s.x - q.x, dy = s.y - q.y;
      float prev = pstl_sqrt_search(fmaf(dx, dx, dy * dy));
      float bestv = INFINITY;
      int bi = 0;
      for (int j = 1; j < c.nseg; ++j) {
        q = sc.lane_xy(l, j);
        dx = s.x - q.x; dy = s.y - q.y;
        const float dj = pstl_sqrt_search(fmaf(dx, dx, dy * dy));
        const float sum = prev + dj;
        if (sum < bestv) { bestv = sum; bi = j - 1; }
        prev = dj;
      }
      const PstlF4 p2 = sc.lane_pt(l, bi), p3 = sc.lane_pt(l, bi + 1);
      float part[3];
      pstl_lane_finish(s.x, s.y, s.th, p2.x, p2.y, p2.z, p3.x, p3.y, c.clip_dist, bi == 0, bi == c.nseg - 2, d, th,
                       GRAD ? part : nullptr);
      if (GRAD) { PSTL_TAPE(6, t) = part[0]; PSTL_TAPE(7, t) = part[1]; PSTL_TAPE(8, t) = part[2]; }
    }
    if (t < pl.need_nei) {
      // utils.py:465-526 + nusc_train.py:142-148 with exact culling (drive_core.cuh).
      // Value-aware bound: when the clearance only feeds soft-min terms  G_t(nei_t - q)  (pl.nei_term >= 0), a
      // step whose clearance exceeds that term's running minimum by more than 36 / (tau log2 e) adds less than
      // 2^-36 to a sum >= 1 — nothing in fp32 — so neighbours that cannot come closer than thr are skipped.
      float thr = 20.f;
#pragma unroll
      for (int k = 0; k < 2; ++k)  // pstl_make_plan keeps the clearance term in slot 0 (the two-slot scan only steers ptxas
        if (k == pl.nei_term) thr = fminf(thr, (am[k] - 36.f) * rg_nei);  // away from a spilling allocation)
      int cnt;
      float best, bg0 = 0.f, bg1 = 0.f, bg2 = 0.f;  // partials of the minimal term (first minimum, as torch.min)
      sc.nei_begin(t, cnt, best);
      // pass 1 (branch-free): which neighbours can be closer than thr at all
      unsigned cand = 0u;
      for (int k = 0; k < cnt; ++k) {
        float valid, ncx, ncy, rsum;
        sc.nei_meta(k, t, valid, ncx, ncy, rsum);
        const float dx = s.x - ncx, dy = s.y - ncy;
        const float R = thr + rsum;
        const bool far = (valid == 1.f) && (R > 0.f) && (fmaf(dx, dx, dy * dy) >= R * R);
        if (valid == 0.f || far) {  // 100, or a clipped clearance >= thr (== 20 when thr == 20): zero gradient
          const float ph = (valid == 0.f) ? 100.f : thr;
          if (ph < best) { best = ph; bg0 = bg1 = bg2 = 0.f; }
        } else {
          cand |= 1u << (k & 31);
        }
        if ((k & 31) == 31 || k == cnt - 1) {
          // pass 2: the survivors, nearest first, re-tested against the running minimum
          if (cand) {
            PstlCircles e;
            pstl_car_circles(s.x, s.y, cs, sn, c.ego_L, c.ego_W, e);
            const int k0 = k & ~31;
            while (cand) {
#if defined(__CUDA_ARCH__)
              const int b = __ffs(cand) - 1;
#else
              const int b = __builtin_ctz(cand);
#endif
              cand &= cand - 1u;
              sc.nei_meta(k0 + b, t, valid, ncx, ncy, rsum);
              if (valid == 1.f && pstl_cull_neighbour_r(s.x - ncx, s.y - ncy, rsum, best)) continue;
              PstlNei nb;
              sc.nei(k0 + b, t, nb);
              float g[3];
              const float term = pstl_pair_clearance(e, cs, sn, nb, GRAD ? g : nullptr);
              if (term < best) {
                best = term;
                if (GRAD) { bg0 = g[0]; bg1 = g[1]; bg2 = g[2]; }
              }
            }
          }
        }
      }
      nei = best;
      if (GRAD) { PSTL_TAPE(9, t) = bg0; PSTL_TAPE(10, t) = bg1; PSTL_TAPE(11, t) = bg2; }
    }
    if (GRAD) {
      PSTL_TAPE(0, t) = cs; PSTL_TAPE(1, t) = sn; PSTL_TAPE(2, t) = s.v;
      PSTL_TAPE(3, t) = d; PSTL_TAPE(4, t) = th; PSTL_TAPE(5, t) = nei;
    }
#pragma unroll
    for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
      if (k < pl.n_terms) {
        const PstlTerm& tm = pl.terms[k];
        if (tm.inner != 0 || (t >= tm.lo && t < tm.hi)) {
          float x2;  // X(t) in tau*log2(e) units (single leaf: without its constant q/den)
          if (tm.pair == 0) {
            const float base = (tm.a.c == 0) ? s.v : (tm.a.c == 1) ? d : (tm.a.c == 2) ? th : nei;
            x2 = base * g2[k];
          } else {  // soft-min / soft-max of two leaves (stl_d_lib.py:21-26)
            const float g = (float)tm.pair * k2;
            const float xa = pstl_plan_leaf(tm.a, s.v, d, th, nei, p) * g, xb = pstl_plan_leaf(tm.b, s.v, d, th, nei, p) * g;
            x2 = (pstl_lg2(1.f + pstl_ex2(-fabsf(xa - xb))) + fmaxf(xa, xb)) * (float)tm.pair;
            if (tm.inner == 0) x2 = x2 * (float)tm.outer;
          }
          if (tm.inner != 0) PSTL_TAPE(pstl_stream_xcol(tm, GRAD), t) = x2;
          else pstl_lse2_add(am[k], as[k], x2);
        }
      }
    }
    if (advance) {
      float w, a;
      pstl_scaled_control(ut, 0, c, w, a);
      s = pstl_unicycle_step(s, w, a, c.dt, cs, sn);
    }
  }
}

PSTL_HD float pstl_stream_top(const PstlPlan& pl, const PstlEvalCfg& c, const float* p, PstlStreamAcc& A);

template <bool GRAD>
PSTL_HD float pstl_stream_finish(const PstlPlan& pl, const PstlEvalCfg& c, const float* p, float* tape, int tstride,
                                 PstlStreamAcc& A) {
  const int T = c.T;
  float (&am)[PSTL_MAX_TERMS] = A.am;
  float (&as)[PSTL_MAX_TERMS] = A.as;
  // terms with an inner suffix operator: y(t) = R2_{t' >= t} X(t') by one backward sweep, folded into R1
#pragma unroll
  for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
    if (k < pl.n_terms && pl.terms[k].inner != 0) {
      const PstlTerm& tm = pl.terms[k];
      const float gi = (float)tm.inner;
      const float go = (float)(tm.inner * tm.outer);  // y2 = (lg2 s + m) * inner ; outer argument = y2 * outer
      float m = PSTL_LSE2_INIT, sm = 0.f;
      const int xc = pstl_stream_xcol(tm, GRAD);
#pragma unroll 1
      for (int t = T - 1; t >= tm.lo; --t) {
        pstl_lse2_add(m, sm, PSTL_TAPE(xc, t) * gi);
        if (t < tm.hi) {
          const float Lt = pstl_lg2(sm) + m;
          if (GRAD) PSTL_TAPE(xc + 1, t) = Lt;
          pstl_lse2_add(am[k], as[k], Lt * go);
        }
      }
    }
  }

  return pstl_stream_top(pl, c, p, A);
}

// top level (stl_d_lib.py:97-112) from the per-term accumulators: ListAnd = soft-min over the terms;
// empty window -> -inf (:7-8,16-17)
PSTL_HD float pstl_stream_top(const PstlPlan& pl, const PstlEvalCfg& c, const float* p, PstlStreamAcc& A) {
  const float k2 = c.tau * 1.4426950408889634f;   // tau * log2(e)
  const float back = 0.6931471805599453f / c.tau;  // ln2 / tau
  float (&am)[PSTL_MAX_TERMS] = A.am;
  float (&as)[PSTL_MAX_TERMS] = A.as;
  float top_m = PSTL_LSE2_INIT, top_s = 0.f, single = 0.f;
  bool empty = false;
#pragma unroll
  for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
    if (k < pl.n_terms) {
      const PstlTerm& tm = pl.terms[k];
      if (tm.hi <= tm.lo) empty = true;
      float y2 = (pstl_lg2(as[k]) + am[k]) * (float)tm.outer;  // term value * tau * log2 e
      if (tm.pair == 0) {  // the leaf's constant, q/den
        float q = tm.a.sp * p[tm.a.pid];
        if (tm.a.den != PSTL_DEN_ONE) q = q / pstl_pred_den(tm.a.den, p);
        y2 = y2 + q * k2;
      }
      single = y2;
      pstl_lse2_add(top_m, top_s, -y2);
    }
  }
  A.top_m = top_m; A.top_s = top_s;
  if (empty) return -INFINITY;
  if (!pl.listand) return single * back;
  return -((pstl_lg2(top_s) + top_m) * back);
}

template <bool GRAD, class Scene>
PSTL_HD float pstl_stream_fwd(const PstlPlan& pl, const Scene& sc, const PstlEvalCfg& c, PstlPose s, const float* u,
                              const float* ego, int es, const float* p, float* tape, int tstride, PstlStreamAcc& A) {
  pstl_stream_init(pl, c, p, A);
  pstl_stream_steps<GRAD>(pl, sc, c, s, u, ego, es, p, tape, tstride, A, 0, pl.need_pose);
  return pstl_stream_finish<GRAD>(pl, c, p, tape, tstride, A);
}

template <class Scene>
PSTL_HD float pstl_stream_eval(const PstlPlan& pl, const Scene& sc, const PstlEvalCfg& c, PstlPose s, const float* u,
                               const float* ego, int es, const float* p, float* tape, int tstride) {
  PstlStreamAcc A;
  return pstl_stream_fwd<false>(pl, sc, c, s, u, ego, es, p, tape, tstride, A);
}

// floats of tape per trajectory for the forward+reverse pair
PSTL_HD int pstl_stream_grad_floats(int n_tapes, int T) { return (PSTL_STREAM_GRAD_COLS + 2 * n_tapes) * T; }

// Reverse sweep after pstl_stream_fwd<true>: d loss / d controls (gu, T*2, pre-scale; may be null) and
// d loss / d ego states (ge, T*4; may be null) for gscore = d loss / d score.
//   score = -lse2_k(-y2_k)/k2                       W_k  = softmax_k(-y2_k)
//   y2_k  = outer * lse2_t(outer * X2_k(t))          p(t) = 2^(outer X2(t) - m_k) / s_k
//   inner terms: d y2_k / d X2(t') = 2^(z(t')) * sum_{t <= t'} 2^((go - 1) L_t - m_k - lg2 s_k),  z = inner * X2
// (one prefix pass in the log domain), then the adjoint of the Euler rollout exactly as pstl_eval_traj_bwd.
PSTL_HD void pstl_stream_bwd(const PstlPlan& pl, const PstlEvalCfg& c, const float* u, const float* p, float gscore,
                             bool finite, const PstlStreamAcc& A, float* tape, int tstride, float* gu, float* ge) {
  const int T = c.T;
  const float k2 = c.tau * 1.4426950408889634f;
  const float back = 0.6931471805599453f / c.tau;
  float W[PSTL_MAX_TERMS];
#pragma unroll
  for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
    W[k] = 0.f;
    if (k < pl.n_terms && finite) {
      const PstlTerm& tm = pl.terms[k];
      float y2 = (pstl_lg2(A.as[k]) + A.am[k]) * (float)tm.outer;
      if (tm.pair == 0) {
        float q = tm.a.sp * p[tm.a.pid];
        if (tm.a.den != PSTL_DEN_ONE) q = q / pstl_pred_den(tm.a.den, p);
        y2 = y2 + q * k2;
      }
      W[k] = pl.listand ? gscore * pstl_ex2(-y2 - A.top_m) / A.top_s : gscore;
    }
  }
  // inner-operator terms: prefix pass, X column + 1 becomes d y2_k / d X2(t)
#pragma unroll
  for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
    if (k < pl.n_terms && pl.terms[k].inner != 0) {
      const PstlTerm& tm = pl.terms[k];
      const int xc = pstl_stream_xcol(tm, true);
      const float gi = (float)tm.inner, go1 = (float)(tm.inner * tm.outer) - 1.f;
      const float norm = A.am[k] + pstl_lg2(A.as[k]);
      float cm = PSTL_LSE2_INIT, cs = 0.f;
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        float g = 0.f;
        if (t >= tm.lo) {
          if (t < tm.hi) pstl_lse2_add(cm, cs, go1 * PSTL_TAPE(xc + 1, t) - norm);
          g = cs * pstl_ex2(gi * PSTL_TAPE(xc, t) + cm);
        }
        PSTL_TAPE(xc + 1, t) = g;
      }
    }
  }
  const int need_pose = pl.need_pose;
  for (int t = need_pose; t < T; ++t) {  // poses the formula never reads
    if (ge) { ge[t * 4 + 0] = 0.f; ge[t * 4 + 1] = 0.f; ge[t * 4 + 2] = 0.f; ge[t * 4 + 3] = 0.f; }
    if (gu) { gu[2 * t] = 0.f; gu[2 * t + 1] = 0.f; }
  }
  float ax = 0.f, ay = 0.f, ath = 0.f, av = 0.f;  // adjoint of s_{t+1} accumulated so far
#pragma unroll 1
  for (int t = need_pose - 1; t >= 0; --t) {
    const float v = PSTL_TAPE(2, t), d = PSTL_TAPE(3, t), th = PSTL_TAPE(4, t), nei = PSTL_TAPE(5, t);
    float G[4] = {0.f, 0.f, 0.f, 0.f};  // d loss / d (v, lane distance, heading error, clearance) at step t
#pragma unroll
    for (int k = 0; k < PSTL_MAX_TERMS; ++k) {
      if (k < pl.n_terms && W[k] != 0.f) {
        const PstlTerm& tm = pl.terms[k];
        const bool in = tm.inner != 0 ? (t >= tm.lo) : (t >= tm.lo && t < tm.hi);
        if (in) {
          const int xc = pstl_stream_xcol(tm, true);
          if (tm.pair == 0) {
            const float base = (tm.a.c == 0) ? v : (tm.a.c == 1) ? d : (tm.a.c == 2) ? th : nei;
            // d X / d base = sb / den = g2 / k2 (times outer where g2 carries it)
            float w, dxb;
            if (tm.inner != 0) { w = PSTL_TAPE(xc + 1, t); dxb = A.g2[k] * back; }
            else { w = pstl_ex2(base * A.g2[k] - A.am[k]) / A.as[k]; dxb = A.g2[k] * back * (float)tm.outer; }
            const float gb = W[k] * w * dxb;
#pragma unroll
            for (int q = 0; q < 4; ++q) G[q] += (tm.a.c == q) ? gb : 0.f;
          } else {
            const float g = (float)tm.pair * k2;
            const float xa = pstl_plan_leaf(tm.a, v, d, th, nei, p) * g, xb = pstl_plan_leaf(tm.b, v, d, th, nei, p) * g;
            const float l = pstl_lg2(1.f + pstl_ex2(-fabsf(xa - xb))) + fmaxf(xa, xb);
            float w;
            if (tm.inner != 0) w = PSTL_TAPE(xc + 1, t);
            else w = pstl_ex2(l * (float)(tm.pair * tm.outer) - A.am[k]) / A.as[k];
            const float gx = W[k] * w;  // d loss / d X(t)
            float da = tm.a.sb, db = tm.b.sb;
            if (tm.a.den != PSTL_DEN_ONE) da = da / pstl_pred_den(tm.a.den, p);
            if (tm.b.den != PSTL_DEN_ONE) db = db / pstl_pred_den(tm.b.den, p);
            const float ga = gx * pstl_ex2(xa - l) * da, gb = gx * pstl_ex2(xb - l) * db;
#pragma unroll
            for (int q = 0; q < 4; ++q) G[q] += ((tm.a.c == q) ? ga : 0.f) + ((tm.b.c == q) ? gb : 0.f);
          }
        }
      }
    }
    float lx = 0.f, ly = 0.f, lth = 0.f;
    const float lv = G[0];
    if (t < pl.need_lane) {
      lx += G[1] * PSTL_TAPE(6, t);
      ly += G[1] * PSTL_TAPE(7, t);
      lth += G[2] * PSTL_TAPE(8, t);
    }
    if (t < pl.need_nei && G[3] != 0.f) {
      lx += G[3] * PSTL_TAPE(9, t); ly += G[3] * PSTL_TAPE(10, t); lth += G[3] * PSTL_TAPE(11, t);
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
      const float cs = PSTL_TAPE(0, t), sn = PSTL_TAPE(1, t);
      const float nth = ath + ax * (-(v * sn) * c.dt) + ay * ((v * cs) * c.dt);
      const float nv = av + ax * (cs * c.dt) + ay * (sn * c.dt);
      ax += lx; ay += ly; ath = nth + lth; av = nv + lv;
    }
  }
}
#undef PSTL_TAPE


#if defined(__CUDACC__)
// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
struct StreamSceneSmem {
  const float4* nb;   // [TC][K slots][PSTL_STREAM_NEI_F4], valid neighbours only, nearest (to the scene's ego) first
  const float2* hdr;  // [TC] (count, initial minimum)
  const float4* ln;   // [3][nseg] (x, y, theta, 0)
  int K, t0, nseg;    // the tile holds steps t0 .. t0+TC-1 of the horizon
  __device__ __forceinline__ PstlF4 lane_pt(int l, int j) const {
    const float4 q = ln[l * nseg + j];
    return PstlF4{q.x, q.y, q.z, q.w};
  }
  __device__ __forceinline__ PstlF4 lane_xy(int l, int j) const { return lane_pt(l, j); }
  __device__ __forceinline__ void nei_begin(int t, int& count, float& init) const {
    const float2 h = hdr[t - t0];
    count = __float_as_int(h.x);
    init = h.y;
  }
  __device__ __forceinline__ void nei_meta(int k, int t, float& valid, float& cx, float& cy, float& rsum) const {
    const float4 q = nb[((t - t0) * K + k) * PSTL_STREAM_NEI_F4 + 2];
    valid = q.x; cx = q.y; cy = q.z; rsum = q.w;
  }
  __device__ __forceinline__ void nei(int k, int t, PstlNei& o) const {
    const float4* q = nb + ((t - t0) * K + k) * PSTL_STREAM_NEI_F4;
    const float4 a = q[0], b = q[1];
    o.cx[0] = a.x; o.cx[1] = a.y; o.cx[2] = a.z; o.cx[3] = a.w;
    o.cy[0] = b.x; o.cy[1] = b.y; o.cy[2] = b.z; o.cy[3] = b.w;
    o.valid = q[2].x;
    o.r = q[3].x;
  }
};

// shared-memory tile of TC steps of one scene, in float4 units: neighbours | lanes | per-step headers | sort keys
__host__ __device__ __forceinline__ size_t stream_tile_f4(int K, int TC, int nseg) {
  return (size_t)K * TC * PSTL_STREAM_NEI_F4 + (size_t)3 * nseg + (size_t)(TC + 1) / 2 + (size_t)(K * TC + 3) / 4;
}

// Block-cooperative staging.  Per step the valid neighbours are compacted and ordered nearest-first with
// respect to the constant-velocity prediction of the block's first row, so the running minimum tightens early
// and the exact cull rejects most of the rest (ordering and compaction never change the minimum itself).
// Stages steps [t0, t0 + tc) of the horizon (tc <= a.tc, the tile's capacity).
__device__ void stream_stage_scene(const ScoreArgs& a, int scene, int n0, float4* tile, int t0, int tc) {
  const PstlEvalCfg& c = a.cfg;
  const int K = c.K, T = c.T, TC = a.tc;
  float4* ln = tile + (size_t)K * TC * PSTL_STREAM_NEI_F4;
  float2* hdr = reinterpret_cast<float2*>(ln + 3 * c.nseg);
  float* keys = reinterpret_cast<float*>(ln + 3 * c.nseg + (TC + 1) / 2);
  const float* nb = a.neighbors + (size_t)scene * K * T * 7;
  // reference point of the ordering heuristic
  float rx = 0.f, ry = 0.f, rvx = 0.f, rvy = 0.f;
  if (a.state0) {
    const float* s0 = a.state0 + (size_t)n0 * 4;
    rx = s0[0]; ry = s0[1];
    rvx = s0[3] * cosf(s0[2]) * c.dt; rvy = s0[3] * sinf(s0[2]) * c.dt;
  }
  for (int e = threadIdx.x; e < K * tc; e += blockDim.x) {
    const int k = e / tc, tl = e - k * tc, t = t0 + tl;
    const float* p = nb + ((size_t)k * T + t) * 7;
    float px = rx + rvx * (float)t, py = ry + rvy * (float)t;
    if (a.ego) { const float* q = a.ego + ((size_t)n0 * T + t) * a.ego_stride; px = q[0]; py = q[1]; }
    const float dx = p[1] - px, dy = p[2] - py;
    keys[tl * K + k] = (p[0] != 0.f) ? dx * dx + dy * dy : INFINITY;
  }
  for (int e = threadIdx.x; e < 3 * c.nseg; e += blockDim.x) {
    const int l = e / c.nseg, j = e - l * c.nseg;
    const float* src = a.lanes[l] + ((size_t)scene * c.nseg + j) * 3;
    ln[e] = make_float4(src[0], src[1], src[2], 0.f);
  }
  __syncthreads();
  const float ego_half = pstl_car_reach(c.ego_L, c.ego_W);
  for (int e = threadIdx.x; e < K * tc; e += blockDim.x) {
    const int k = e / tc, tl = e - k * tc, t = t0 + tl;
    const float* kt = keys + tl * K;
    const float key = kt[k];
    int rank = 0, n_valid = 0;
    for (int j = 0; j < K; ++j) {
      const float kj = kt[j];
      n_valid += (kj != INFINITY) ? 1 : 0;
      rank += (kj < key || (kj == key && j < k)) ? 1 : 0;
    }
    if (k == 0) hdr[tl] = make_float2(__int_as_float(n_valid), n_valid < K ? 100.f : INFINITY);
    if (key == INFINITY) continue;  // valid == 0: clip(d)*0 + (1-0)*100, folded into the initial minimum
    const float* p = nb + ((size_t)k * T + t) * 7;
    PstlCircles cc;
    pstl_car_circles(p[1], p[2], cosf(p[3]), sinf(p[3]), p[5], p[6], cc);
    float4* o = tile + (size_t)(tl * K + rank) * PSTL_STREAM_NEI_F4;
    o[0] = make_float4(cc.cx[0], cc.cx[1], cc.cx[2], cc.cx[3]);
    o[1] = make_float4(cc.cy[0], cc.cy[1], cc.cy[2], cc.cy[3]);
    o[2] = make_float4(p[0], p[1], p[2], ego_half + pstl_car_reach(p[5], p[6]) + 1e-3f);
    o[3] = make_float4(cc.r, 0.f, 0.f, 0.f);
  }
}

struct StreamPlans {
  PstlPlan p[3];
};

// Dense per-row inputs come with arbitrary modes: a block sorts its rows by mode (stable counting sort, one thread —
// ~1k cycles against ~100k of scoring) so that most warps evaluate ONE plan instead of serialising three.
__device__ __forceinline__ int stream_sort_rows_by_mode(const float* __restrict__ mode, int n0, int N, int* s_perm) {
  __shared__ unsigned char s_cls[256];
  {
    int v = 4;
    if (n0 + (int)threadIdx.x < N) {
      const float md = mode[n0 + threadIdx.x];
      v = (md == 0.f) ? 0 : (md == 1.f) ? 1 : (md == 2.f) ? 2 : 3;
    }
    s_cls[threadIdx.x] = (unsigned char)v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int B = blockDim.x;
    int cnt[5] = {0, 0, 0, 0, 0};
    for (int r = 0; r < B; ++r) ++cnt[s_cls[r]];
    int off[5], acc = 0;
    for (int v = 0; v < 5; ++v) { off[v] = acc; acc += cnt[v]; }
    for (int r = 0; r < B; ++r) s_perm[off[s_cls[r]]++] = r;
  }
  __syncthreads();
  return s_perm[threadIdx.x];
}

#define PSTL_STREAM_BLOCK_MAX 192

// CHUNKED (horizon walked in tiles of a.tc steps) is a template parameter: compiled into the same kernel, the two walks
// made it 11,200 instructions (180 KB, more than the instruction cache: 11 % of the pipeline launch's stall samples were
// "no instruction"); the whole-horizon instance the pipeline launches carries only its own walk.
template <bool SMEM_SCENE, bool CHUNKED>
__global__ void __launch_bounds__(PSTL_STREAM_BLOCK_MAX, 4)
k_score_stream(const __grid_constant__ ScoreArgs a, const __grid_constant__ StreamPlans sp) {
  extern __shared__ float4 sm4[];
  const PstlEvalCfg c = a.cfg;
  const int B = blockDim.x, T = c.T;
  const int n0 = blockIdx.x * B;
  // rows of the pipeline cycle through the three formulas (n % 3): deal them to warps so that a warp
  // evaluates ONE plan (a divergence optimisation only; every thread still reads its own mode)
  __shared__ int s_perm[PSTL_STREAM_BLOCK_MAX];
  int li = threadIdx.x;
  if (!SMEM_SCENE) {
    li = stream_sort_rows_by_mode(a.mode, n0, a.N, s_perm);
  } else if (B % 96 == 0) {
    const int g = li / 96, w = li - g * 96;
    li = g * 96 + (w & 31) * 3 + (w >> 5);
  }
  const int n = n0 + li;
  // the scene tile holds a.tc steps: the whole horizon when it fits (staged once), else the horizon is walked in
  // chunks and the tile re-staged per chunk (long horizons / many neighbours, BASELINE config 5)
  const int TC = SMEM_SCENE ? a.tc : T;
  constexpr bool chunked = SMEM_SCENE && CHUNKED;
  float4* tile = sm4;
  float* tape = reinterpret_cast<float*>(sm4);
  int tstride = B;
  if (SMEM_SCENE) {
    if (!chunked) {
      stream_stage_scene(a, n0 / a.rows_per_scene, n0, tile, 0, T);
      __syncthreads();
    }
    tape = reinterpret_cast<float*>(sm4 + stream_tile_f4(c.K, TC, c.nseg));
  }
  tape += threadIdx.x;
  const bool live = n < a.N;
  int m = 4;
  if (live) {
    const float md = a.mode[n];
    m = (md == 0.f) ? 0 : (md == 1.f) ? 1 : (md == 2.f) ? 2 : (md == 3.f) ? 3 : 4;
  }
  const int nn = live ? n : 0;
  PstlPose s0{0.f, 0.f, 0.f, 0.f};
  if (a.state0) {
    const float4 q = *reinterpret_cast<const float4*>(a.state0 + (size_t)nn * 4);
    s0 = PstlPose{q.x, q.y, q.z, q.w};
  }
  const float* stlp = a.stlp + (size_t)nn * 6;
  const float* ego = a.ego ? a.ego + (size_t)nn * T * a.ego_stride : nullptr;
  const int scene = nn / a.rows_per_scene;
  if (a.tape_global) { tape = a.ws + nn; tstride = a.N; }  // the X(t) columns do not fit next to the tile
  const float4* tile_ln = tile + (size_t)c.K * TC * PSTL_STREAM_NEI_F4;
  StreamSceneSmem ss{tile, reinterpret_cast<const float2*>(tile_ln + 3 * c.nseg), tile_ln, c.K, 0, c.nseg};
  PstlStreamSceneGlobal sg;
  sg.neib = a.neighbors + (size_t)scene * c.K * T * 7;
  for (int l = 0; l < 3; ++l) sg.ln[l] = a.lanes[l] + (size_t)scene * c.nseg * 3;
  sg.K = c.K; sg.T = T; sg.ego_half = pstl_car_reach(c.ego_L, c.ego_W);

  float best = -INFINITY;
  int bi = 0;
#pragma unroll 1
  for (int cand = 0; cand < a.C; ++cand) {
    const float* u = a.controls ? a.controls + ((size_t)cand * a.N + nn) * T * 2 : nullptr;
    float sc = (m == 3) ? 1.0f : 0.0f;  // nusc_train.py:322 outlier score; unknown mode selects nothing (:150-151)
    if (!chunked) {
#pragma unroll 1
      for (int mm = 0; mm < 3; ++mm) {
        if (!__any_sync(0xffffffffu, m == mm)) continue;
        if (m == mm) {
          sc = SMEM_SCENE ? pstl_stream_eval(sp.p[mm], ss, c, s0, u, ego, a.ego_stride, stlp, tape, tstride)
                          : pstl_stream_eval(sp.p[mm], sg, c, s0, u, ego, a.ego_stride, stlp, tape, tstride);
        }
      }
    } else {
      PstlStreamAcc A;
      PstlPose s = s0;
      if (m < 3) pstl_stream_init(sp.p[m], c, stlp, A);
#pragma unroll 1
      for (int t0 = 0; t0 < T; t0 += TC) {
        const int tc = (T - t0 < TC) ? T - t0 : TC;
        __syncthreads();  // every thread is done with the previous chunk
        stream_stage_scene(a, n0 / a.rows_per_scene, n0, tile, t0, tc);
        __syncthreads();
        ss.t0 = t0;
#pragma unroll 1
        for (int mm = 0; mm < 3; ++mm) {
          if (!__any_sync(0xffffffffu, m == mm)) continue;
          if (m == mm) {
            const int t1 = (t0 + tc < sp.p[mm].need_pose) ? t0 + tc : sp.p[mm].need_pose;
            pstl_stream_steps<false>(sp.p[mm], ss, c, s, u, ego, a.ego_stride, stlp, tape, tstride, A, t0, t1);
          }
        }
      }
      if (m < 3) sc = pstl_stream_finish<false>(sp.p[m], c, stlp, tape, tstride, A);
    }
    if (live && a.scores_all) a.scores_all[(size_t)cand * a.N + n] = sc;
    if (cand == 0 || sc > best) { best = sc; bi = cand; }  // torch.max(dim=0): first maximum
  }
  if (!live) return;
  if (a.best_score) a.best_score[n] = best;
  if (a.best_idx) a.best_idx[n] = bi;
  if ((a.best_controls || a.traj_out) && a.controls) {
    const float* u = a.controls + ((size_t)bi * a.N + n) * T * 2;
    PstlPose s = s0;
    for (int t = 0; t < T; ++t) {
      float w, ac;
      pstl_scaled_control(u, t, c, w, ac);
      if (a.best_controls) *reinterpret_cast<float2*>(a.best_controls + ((size_t)n * T + t) * 2) = make_float2(w, ac);
      if (a.traj_out) {
        *reinterpret_cast<float4*>(a.traj_out + ((size_t)n * (T + 1) + t) * 4) = make_float4(s.x, s.y, s.th, s.v);
        float sn, cs;
        sincosf(s.th, &sn, &cs);
        s = pstl_unicycle_step(s, w, ac, c.dt, cs, sn);
      }
    }
    if (a.traj_out) *reinterpret_cast<float4*>(a.traj_out + ((size_t)n * (T + 1) + T) * 4) = make_float4(s.x, s.y, s.th, s.v);
  }
}

// Reverse mode (guidance, autograd of compute_stl_dense): forward with the per-step record in a global tape
// (element-major, stride N: coalesced), then pstl_stream_bwd.  C == 1.
template <bool SMEM_SCENE, bool CHUNKED>
__global__ void __launch_bounds__(PSTL_STREAM_BLOCK_MAX, 3)
k_score_stream_bwd(const __grid_constant__ ScoreArgs a, const __grid_constant__ StreamPlans sp) {
  extern __shared__ float4 sm4[];
  const PstlEvalCfg c = a.cfg;
  const int B = blockDim.x, T = c.T;
  const int n0 = blockIdx.x * B;
  __shared__ int s_perm[PSTL_STREAM_BLOCK_MAX];
  int li = threadIdx.x;
  if (!SMEM_SCENE) {
    li = stream_sort_rows_by_mode(a.mode, n0, a.N, s_perm);
  } else if (B % 96 == 0) {
    const int g = li / 96, w = li - g * 96;
    li = g * 96 + (w & 31) * 3 + (w >> 5);
  }
  const int n = n0 + li;
  const int TC = SMEM_SCENE ? a.tc : T;
  constexpr bool chunked = SMEM_SCENE && CHUNKED;
  float4* tile = sm4;
  if (SMEM_SCENE && !chunked) {
    stream_stage_scene(a, n0 / a.rows_per_scene, n0, tile, 0, T);
    __syncthreads();
  }
  const bool live = n < a.N;
  int m = 4;
  if (live) {
    const float md = a.mode[n];
    m = (md == 0.f) ? 0 : (md == 1.f) ? 1 : (md == 2.f) ? 2 : (md == 3.f) ? 3 : 4;
  }
  const int nn = live ? n : 0;
  PstlPose s0{0.f, 0.f, 0.f, 0.f};
  if (a.state0) {
    const float4 q = *reinterpret_cast<const float4*>(a.state0 + (size_t)nn * 4);
    s0 = PstlPose{q.x, q.y, q.z, q.w};
  }
  const float* stlp = a.stlp + (size_t)nn * 6;
  const float* ego = a.ego ? a.ego + (size_t)nn * T * a.ego_stride : nullptr;
  const float* u = a.controls ? a.controls + (size_t)nn * T * 2 : nullptr;
  float* gu = a.grad_controls ? a.grad_controls + (size_t)nn * T * 2 : nullptr;
  float* ge = a.grad_ego ? a.grad_ego + (size_t)nn * T * 4 : nullptr;
  float* tape = a.ws + nn;
  const int scene = nn / a.rows_per_scene;
  const float4* tile_ln = tile + (size_t)c.K * TC * PSTL_STREAM_NEI_F4;
  StreamSceneSmem ss{tile, reinterpret_cast<const float2*>(tile_ln + 3 * c.nseg), tile_ln, c.K, 0, c.nseg};
  PstlStreamSceneGlobal sg;
  sg.neib = a.neighbors + (size_t)scene * c.K * T * 7;
  for (int l = 0; l < 3; ++l) sg.ln[l] = a.lanes[l] + (size_t)scene * c.nseg * 3;
  sg.K = c.K; sg.T = T; sg.ego_half = pstl_car_reach(c.ego_L, c.ego_W);

  if (live && m >= 3) {
    if (a.scores) a.scores[n] = (m == 3) ? 1.0f : 0.0f;
    if (gu) for (int i = 0; i < T * 2; ++i) gu[i] = 0.f;
    if (ge) for (int i = 0; i < T * 4; ++i) ge[i] = 0.f;
  }
  PstlStreamAcc A;
  float sc = 0.f;
  if (!chunked) {
#pragma unroll 1
    for (int mm = 0; mm < 3; ++mm) {
      if (!__any_sync(0xffffffffu, m == mm)) continue;
      if (m == mm)
        sc = SMEM_SCENE ? pstl_stream_fwd<true>(sp.p[mm], ss, c, s0, u, ego, a.ego_stride, stlp, tape, a.N, A)
                        : pstl_stream_fwd<true>(sp.p[mm], sg, c, s0, u, ego, a.ego_stride, stlp, tape, a.N, A);
    }
  } else {
    PstlPose s = s0;
    if (m < 3) pstl_stream_init(sp.p[m], c, stlp, A);
#pragma unroll 1
    for (int t0 = 0; t0 < T; t0 += TC) {
      const int tc = (T - t0 < TC) ? T - t0 : TC;
      __syncthreads();
      stream_stage_scene(a, n0 / a.rows_per_scene, n0, tile, t0, tc);
      __syncthreads();
      ss.t0 = t0;
#pragma unroll 1
      for (int mm = 0; mm < 3; ++mm) {
        if (!__any_sync(0xffffffffu, m == mm)) continue;
        if (m == mm) {
          const int t1 = (t0 + tc < sp.p[mm].need_pose) ? t0 + tc : sp.p[mm].need_pose;
          pstl_stream_steps<true>(sp.p[mm], ss, c, s, u, ego, a.ego_stride, stlp, tape, a.N, A, t0, t1);
        }
      }
    }
    if (m < 3) sc = pstl_stream_finish<true>(sp.p[m], c, stlp, tape, a.N, A);
  }
  if (m < 3) {  // reverse sweep: scene-independent
    if (a.scores) a.scores[n] = sc;
    float g;
    if (a.grad_score) g = a.grad_score[n];
    else g = (a.thres - sc > 0.f) ? -a.valid[n] * (a.inv_norm_dev ? __ldg(a.inv_norm_dev) : a.inv_norm) : 0.f;  // guidance loss, nusc_train.py:616-619
    pstl_stream_bwd(sp.p[m], c, u, stlp, g, sc != -INFINITY, A, tape, a.N, gu, ge);
  }
}
#endif  // __CUDACC__

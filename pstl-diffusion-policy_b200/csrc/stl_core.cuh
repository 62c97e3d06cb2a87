// stl_core.cuh — postfix STL program: resolved form + per-trajectory interpreter (forward and
// reverse mode).  One thread owns one trajectory; its traces live in a "tape" addressed as
// tape[offset * stride] so the same code runs on a shared-memory tile (stride = block+1) or on
// a global workspace (stride = N, coalesced across threads).
//
// Semantics follow the reference's stl_d_lib.py (file:line in each case below).
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/pstl.h"

#if defined(__CUDACC__)
#define PSTL_HD __host__ __device__ __forceinline__
#else
#define PSTL_HD inline
#endif

#define PSTL_MAX_OPS 48
#define PSTL_NEG_INF (-INFINITY)

struct PstlROp {
  int op, a0, a1;
  int in0, in1;      // producing op index of the inputs (SMIN_K: a1 = base into klist, a0 = k)
  int out_off;       // tape offset (floats) of the output trace
  int n_out;         // outputs t in [0,n_out) are needed downstream
};

struct PstlProgView {
  int n_ops, n_signals, T, need_t;
  int val_floats;             // floats of value tape per trajectory (base signals + op outputs)
  int base_off[PSTL_N_BASE_SIGNALS];   // tape offset of each fused base signal, -1 if unused
  int base_need[PSTL_N_BASE_SIGNALS];  // how many leading steps of it are read
  int part_off;               // offset of the partial-derivative block (grad tapes only)
  int grad_floats;            // total floats with adjoints + partials
  PstlROp ops[PSTL_MAX_OPS];
  int klist[PSTL_MAX_OPS];
};

// ----------------------------------------------------------------------------------------
// soft reductions, written to mirror torch.logsumexp (amax, exp(x-max) summed in order, log, +max)
// ----------------------------------------------------------------------------------------
struct PstlLse {
  float m, s;
  PSTL_HD void init() { m = PSTL_NEG_INF; s = 0.f; }
};

// FAST selects the SFU intrinsics (ex2/lg2 based, ~2 ulp) inside the fused scoring kernels, where
// arguments are tau-scaled and the result is divided by tau again; the node API keeps expf/logf.
template <bool FAST>
PSTL_HD float pstl_exp(float x) {
#if defined(__CUDA_ARCH__)
  if (FAST) return __expf(x);
#endif
  return expf(x);
}
template <bool FAST>
PSTL_HD float pstl_log(float x) {
#if defined(__CUDA_ARCH__)
  if (FAST) return __logf(x);
#endif
  return logf(x);
}

template <bool FAST>
PSTL_HD float pstl_lse_finish(float m, float s) {
  // torch: maxes_squeezed = where(|max|==inf, 0, max); log(sum(exp(x-maxes_squeezed))) + maxes_squeezed
  return pstl_log<FAST>(s) + m;
}

// two-pass reduction over n values fetched by functor f(j) (already scaled by +-tau)
template <bool FAST, class F>
PSTL_HD float pstl_lse_n(int n, F f) {
  float m = PSTL_NEG_INF;
  for (int j = 0; j < n; ++j) m = fmaxf(m, f(j));
  if (isinf(m)) m = 0.f;
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += pstl_exp<FAST>(f(j) - m);
  return pstl_lse_finish<FAST>(m, s);
}

PSTL_HD float pstl_logaddexp(float a, float b) {
  // torch logcumsumexp helper: max + log1p(exp(min-max)); both -inf -> -inf
  float mn = fminf(a, b), mx = fmaxf(a, b);
  if (mn != mx || isfinite(mn)) return log1pf(expf(mn - mx)) + mx;
  return a;
}

PSTL_HD int pstl_clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ----------------------------------------------------------------------------------------
// Hoisted accessors: an op decodes each input ONCE (plain trace, or typed predicate evaluated from
// a base signal as (sb*x + pp)/den) and its inner loops only do get()/add().
// ----------------------------------------------------------------------------------------
struct PstlIn {
  const float* p;
  float sb, pp, den;
  int pred;
  PSTL_HD float get(int t, int stride) const {
    float v = p[(size_t)t * stride];
    if (pred) {
      v = sb * v + pp;  // sb = +-1: exact, so this is the reference's single rounding of x - p / -x + p
      if (den != 1.f) v = v / den;
    }
    return v;
  }
};
struct PstlOut {
  float* g;
  float w;
  PSTL_HD void add(int t, int stride, float grad) const { g[(size_t)t * stride] += grad * w; }
};

// ----------------------------------------------------------------------------------------
// forward.  Leaf must provide: float signal(int p, int t); PstlIn pred_in(int a0, int a1).
// ----------------------------------------------------------------------------------------
template <bool FAST = false, class Leaf>
PSTL_HD void pstl_interp_fwd(const PstlProgView& P, float* tape, int stride, float tau, int hard, Leaf& leaf) {
  const int T = P.T;
#define TP(off) tape[(size_t)(off) * stride]
  auto IN = [&](int idx) -> PstlIn {
    const PstlROp& p = P.ops[idx];
    if (p.op == PSTL_OP_PRED) return leaf.pred_in(p.a0, p.a1);
    PstlIn r;
    r.p = tape + (size_t)p.out_off * stride;
    r.sb = 1.f; r.pp = 0.f; r.den = 1.f; r.pred = 0;
    return r;
  };
  for (int i = 0; i < P.n_ops; ++i) {
    const PstlROp o = P.ops[i];
    const int oo = o.out_off;
    switch (o.op) {
      case PSTL_OP_SIGNAL:
        for (int t = 0; t < o.n_out; ++t) TP(oo + t) = leaf.signal(o.a0, t);
        break;
      case PSTL_OP_PRED:  // typed leaves are evaluated by their consumer; only a bare-leaf formula stores
        if (oo >= 0) {
          const PstlIn a = leaf.pred_in(o.a0, o.a1);
          for (int t = 0; t < o.n_out; ++t) TP(oo + t) = a.get(t, stride);
        }
        break;
      case PSTL_OP_NEG: {  // stl_d_lib.py:130-131
        const PstlIn a = IN(o.in0);
        for (int t = 0; t < o.n_out; ++t) TP(oo + t) = -a.get(t, stride);
      } break;
      case PSTL_OP_SMIN2:
      case PSTL_OP_SMAX2: {  // stl_d_lib.py:21-26 (stack dim=1, logsumexp)
        const float sg = (o.op == PSTL_OP_SMIN2) ? -1.f : 1.f;
        const PstlIn ia = IN(o.in0), ib = IN(o.in1);
        for (int t = 0; t < o.n_out; ++t) {
          const float a = sg * ia.get(t, stride), b = sg * ib.get(t, stride);
          float r;
          if (hard) {
            r = fmaxf(a, b);
          } else {
            const float xa = a * tau, xb = b * tau;
            float m = fmaxf(xa, xb);
            if (isinf(m)) m = 0.f;
            r = pstl_lse_finish<FAST>(m, pstl_exp<FAST>(xa - m) + pstl_exp<FAST>(xb - m)) / tau;
          }
          TP(oo + t) = sg * r;
        }
      } break;
      case PSTL_OP_SMIN_K: {  // stl_d_lib.py:101-107 (soft-min across the k stacked children)
        const int k = o.a0, kb = o.a1;
        for (int t = 0; t < o.n_out; ++t) {
          float m = PSTL_NEG_INF;
          for (int j = 0; j < k; ++j) m = fmaxf(m, -IN(P.klist[kb + j]).get(t, stride) * (hard ? 1.f : tau));
          float r = m;
          if (!hard) {
            if (isinf(m)) m = 0.f;
            float s = 0.f;
            for (int j = 0; j < k; ++j) s += pstl_exp<FAST>(-IN(P.klist[kb + j]).get(t, stride) * tau - m);
            r = pstl_lse_finish<FAST>(m, s) / tau;
          }
          TP(oo + t) = -r;
        }
      } break;
      case PSTL_OP_WIN_SMIN:
      case PSTL_OP_WIN_SMAX: {  // stl_d_lib.py:151,164 window [t+ts, t+te) clipped to [0,T); empty -> -inf (:7-8,:16-17)
        const float sg = (o.op == PSTL_OP_WIN_SMIN) ? -1.f : 1.f;
        const PstlIn a = IN(o.in0);
        for (int t = 0; t < o.n_out; ++t) {
          const int lo = pstl_clipi(t + o.a0, 0, T), hi = pstl_clipi(t + o.a1, 0, T);
          if (hi <= lo) {
            TP(oo + t) = PSTL_NEG_INF;
            continue;
          }
          float r;
          if (hard) {
            r = PSTL_NEG_INF;
            for (int j = lo; j < hi; ++j) r = fmaxf(r, sg * a.get(j, stride));
          } else {
            r = pstl_lse_n<FAST>(hi - lo, [&](int j) { return sg * a.get(lo + j, stride) * tau; }) / tau;
          }
          TP(oo + t) = sg * r;
        }
      } break;
      case PSTL_OP_PREFIX_SMIN: {  // stl_d_lib.py:189
        const PstlIn a = IN(o.in0);
        float acc = PSTL_NEG_INF;
        for (int t = 0; t < o.n_out; ++t) {
          acc = pstl_logaddexp(acc, -a.get(t, stride) * tau);
          TP(oo + t) = -acc / tau;
        }
      } break;
      case PSTL_OP_SUFFIX_SMAX: {  // stl_d_lib.py:191
        const PstlIn a = IN(o.in0);
        float acc = PSTL_NEG_INF;
        for (int t = T - 1; t >= 0; --t) {
          acc = pstl_logaddexp(acc, a.get(t, stride) * tau);
          if (t < o.n_out) TP(oo + t) = acc / tau;
        }
      } break;
      default:
        break;
    }
  }
#undef TP
}

// ----------------------------------------------------------------------------------------
// reverse mode.  vt = value tape (filled by pstl_interp_fwd), gt = adjoint tape with the same
// offsets; the caller zeroes gt and seeds the top op's adjoint.  LeafGrad must provide
// void signal(int p,int t,float g); PstlOut pred_out(int a0,int a1).
// Soft-max weights are recomputed exactly as torch's logsumexp backward: exp(x*tau - lse).
// ----------------------------------------------------------------------------------------
template <bool FAST = false, class Leaf, class LeafGrad>
PSTL_HD void pstl_interp_bwd(const PstlProgView& P, const float* vt, float* gt, int stride, float tau, int hard,
                             Leaf& leaf, LeafGrad& lg) {
  const int T = P.T;
#define VT(off) vt[(size_t)(off) * stride]
#define GT(off) gt[(size_t)(off) * stride]
  auto IN = [&](int idx) -> PstlIn {
    const PstlROp& p = P.ops[idx];
    if (p.op == PSTL_OP_PRED) return leaf.pred_in(p.a0, p.a1);
    PstlIn r;
    r.p = vt + (size_t)p.out_off * stride;
    r.sb = 1.f; r.pp = 0.f; r.den = 1.f; r.pred = 0;
    return r;
  };
  auto OUT = [&](int idx) -> PstlOut {
    const PstlROp& p = P.ops[idx];
    if (p.op == PSTL_OP_PRED) return lg.pred_out(p.a0, p.a1);
    PstlOut r;
    r.g = gt + (size_t)p.out_off * stride;
    r.w = 1.f;
    return r;
  };
  for (int i = P.n_ops - 1; i >= 0; --i) {
    const PstlROp o = P.ops[i];
    const int oo = o.out_off;
    switch (o.op) {
      case PSTL_OP_SIGNAL:
        for (int t = 0; t < o.n_out; ++t) lg.signal(o.a0, t, GT(oo + t));
        break;
      case PSTL_OP_PRED:
        if (oo >= 0) {
          const PstlOut w = lg.pred_out(o.a0, o.a1);
          for (int t = 0; t < o.n_out; ++t) w.add(t, stride, GT(oo + t));
        }
        break;
      case PSTL_OP_NEG: {
        const PstlOut w = OUT(o.in0);
        for (int t = 0; t < o.n_out; ++t) w.add(t, stride, -GT(oo + t));
      } break;
      case PSTL_OP_SMIN2:
      case PSTL_OP_SMAX2: {
        const float sg = (o.op == PSTL_OP_SMIN2) ? -1.f : 1.f;
        const PstlIn ia = IN(o.in0), ib = IN(o.in1);
        const PstlOut wa = OUT(o.in0), wb = OUT(o.in1);
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const float a = sg * ia.get(t, stride), b = sg * ib.get(t, stride);
          if (hard) {  // torch.max(dim) routes to the first maximal index
            if (a >= b) wa.add(t, stride, g); else wb.add(t, stride, g);
          } else {
            const float xa = a * tau, xb = b * tau;
            float m = fmaxf(xa, xb);
            if (isinf(m)) m = 0.f;
            const float lse = pstl_lse_finish<FAST>(m, pstl_exp<FAST>(xa - m) + pstl_exp<FAST>(xb - m));
            // d out/d in = sg * (1/tau) * softmax * tau * sg = softmax weight
            wa.add(t, stride, g * pstl_exp<FAST>(xa - lse));
            wb.add(t, stride, g * pstl_exp<FAST>(xb - lse));
          }
        }
      } break;
      case PSTL_OP_SMIN_K: {
        const int k = o.a0, kb = o.a1;
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          if (hard) {
            int bj = 0;
            float bv = -IN(P.klist[kb]).get(t, stride);
            for (int j = 1; j < k; ++j) {
              const float v = -IN(P.klist[kb + j]).get(t, stride);
              if (v > bv) { bv = v; bj = j; }
            }
            OUT(P.klist[kb + bj]).add(t, stride, g);
          } else {
            const float lse = pstl_lse_n<FAST>(k, [&](int j) { return -IN(P.klist[kb + j]).get(t, stride) * tau; });
            for (int j = 0; j < k; ++j)
              OUT(P.klist[kb + j]).add(t, stride, g * pstl_exp<FAST>(-IN(P.klist[kb + j]).get(t, stride) * tau - lse));
          }
        }
      } break;
      case PSTL_OP_WIN_SMIN:
      case PSTL_OP_WIN_SMAX: {
        const float sg = (o.op == PSTL_OP_WIN_SMIN) ? -1.f : 1.f;
        const PstlIn a = IN(o.in0);
        const PstlOut w = OUT(o.in0);
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const int lo = pstl_clipi(t + o.a0, 0, T), hi = pstl_clipi(t + o.a1, 0, T);
          if (hi <= lo) continue;
          if (hard) {
            int bj = lo;
            float bv = sg * a.get(lo, stride);
            for (int j = lo + 1; j < hi; ++j) {
              const float v = sg * a.get(j, stride);
              if (v > bv) { bv = v; bj = j; }
            }
            w.add(bj, stride, g);
          } else {
            const float lse = pstl_lse_n<FAST>(hi - lo, [&](int j) { return sg * a.get(lo + j, stride) * tau; });
            for (int j = lo; j < hi; ++j) w.add(j, stride, g * pstl_exp<FAST>(sg * a.get(j, stride) * tau - lse));
          }
        }
      } break;
      case PSTL_OP_PREFIX_SMIN: {
        // out[t] = -LSE_{j<=t}(-x_j tau)/tau  ->  d out[t]/d x_j = exp(-x_j tau - lse_t)
        const PstlIn a = IN(o.in0);
        const PstlOut w = OUT(o.in0);
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const float lse = -VT(oo + t) * tau;
          for (int j = 0; j <= t; ++j) w.add(j, stride, g * expf(-a.get(j, stride) * tau - lse));
        }
      } break;
      case PSTL_OP_SUFFIX_SMAX: {
        const PstlIn a = IN(o.in0);
        const PstlOut w = OUT(o.in0);
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const float lse = VT(oo + t) * tau;
          for (int j = t; j < T; ++j) w.add(j, stride, g * expf(a.get(j, stride) * tau - lse));
        }
      } break;
      default:
        break;
    }
  }
#undef VT
#undef GT
}

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

PSTL_HD float pstl_lse_finish(float m, float s) {
  // torch: maxes_squeezed = where(|max|==inf, 0, max); log(sum(exp(x-maxes_squeezed))) + maxes_squeezed
  return logf(s) + m;
}

// two-pass reduction over n values fetched by functor f(j) (already scaled by +-tau)
template <class F>
PSTL_HD float pstl_lse_n(int n, F f) {
  float m = PSTL_NEG_INF;
  for (int j = 0; j < n; ++j) m = fmaxf(m, f(j));
  if (isinf(m)) m = 0.f;
  float s = 0.f;
  for (int j = 0; j < n; ++j) s += expf(f(j) - m);
  return pstl_lse_finish(m, s);
}

PSTL_HD float pstl_logaddexp(float a, float b) {
  // torch logcumsumexp helper: max + log1p(exp(min-max)); both -inf -> -inf
  float mn = fminf(a, b), mx = fmaxf(a, b);
  if (mn != mx || isfinite(mn)) return log1pf(expf(mn - mx)) + mx;
  return a;
}

PSTL_HD int pstl_clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ----------------------------------------------------------------------------------------
// forward.  Leaf must provide: float signal(int p, int t); float pred(int a0, int a1, int t).
// ----------------------------------------------------------------------------------------
template <class Leaf>
PSTL_HD void pstl_interp_fwd(const PstlProgView& P, float* tape, int stride, float tau, int hard, Leaf& leaf) {
  const int T = P.T;
#define TP(off) tape[(size_t)(off) * stride]
  for (int i = 0; i < P.n_ops; ++i) {
    const PstlROp o = P.ops[i];
    const int oo = o.out_off;
    switch (o.op) {
      case PSTL_OP_SIGNAL:
        for (int t = 0; t < o.n_out; ++t) TP(oo + t) = leaf.signal(o.a0, t);
        break;
      case PSTL_OP_PRED:
        for (int t = 0; t < o.n_out; ++t) TP(oo + t) = leaf.pred(o.a0, o.a1, t);
        break;
      case PSTL_OP_NEG: {  // stl_d_lib.py:130-131
        const int io = P.ops[o.in0].out_off;
        for (int t = 0; t < o.n_out; ++t) TP(oo + t) = -TP(io + t);
      } break;
      case PSTL_OP_SMIN2:
      case PSTL_OP_SMAX2: {  // stl_d_lib.py:21-26 (stack dim=1, logsumexp)
        const int ia = P.ops[o.in0].out_off, ib = P.ops[o.in1].out_off;
        const float sg = (o.op == PSTL_OP_SMIN2) ? -1.f : 1.f;
        for (int t = 0; t < o.n_out; ++t) {
          const float a = sg * TP(ia + t), b = sg * TP(ib + t);
          float r;
          if (hard) {
            r = fmaxf(a, b);
          } else {
            const float xa = a * tau, xb = b * tau;
            float m = fmaxf(xa, xb);
            if (isinf(m)) m = 0.f;
            r = pstl_lse_finish(m, expf(xa - m) + expf(xb - m)) / tau;
          }
          TP(oo + t) = sg * r;
        }
      } break;
      case PSTL_OP_SMIN_K: {  // stl_d_lib.py:101-107 (soft-min across the k stacked children)
        const int k = o.a0, kb = o.a1;
        for (int t = 0; t < o.n_out; ++t) {
          float r;
          if (hard) {
            r = PSTL_NEG_INF;
            for (int j = 0; j < k; ++j) r = fmaxf(r, -TP(P.ops[P.klist[kb + j]].out_off + t));
          } else {
            r = pstl_lse_n(k, [&](int j) { return -TP(P.ops[P.klist[kb + j]].out_off + t) * tau; }) / tau;
          }
          TP(oo + t) = -r;
        }
      } break;
      case PSTL_OP_WIN_SMIN:
      case PSTL_OP_WIN_SMAX: {  // stl_d_lib.py:151,164 window [t+ts, t+te) clipped to [0,T); empty -> -inf (:7-8,:16-17)
        const int io = P.ops[o.in0].out_off;
        const float sg = (o.op == PSTL_OP_WIN_SMIN) ? -1.f : 1.f;
        for (int t = 0; t < o.n_out; ++t) {
          const int lo = pstl_clipi(t + o.a0, 0, T), hi = pstl_clipi(t + o.a1, 0, T);
          if (hi <= lo) {
            TP(oo + t) = PSTL_NEG_INF;
            continue;
          }
          float r;
          if (hard) {
            r = PSTL_NEG_INF;
            for (int j = lo; j < hi; ++j) r = fmaxf(r, sg * TP(io + j));
          } else {
            r = pstl_lse_n(hi - lo, [&](int j) { return sg * TP(io + lo + j) * tau; }) / tau;
          }
          TP(oo + t) = sg * r;
        }
      } break;
      case PSTL_OP_PREFIX_SMIN: {  // stl_d_lib.py:189
        const int io = P.ops[o.in0].out_off;
        float acc = PSTL_NEG_INF;
        for (int t = 0; t < o.n_out; ++t) {
          acc = pstl_logaddexp(acc, -TP(io + t) * tau);
          TP(oo + t) = -acc / tau;
        }
      } break;
      case PSTL_OP_SUFFIX_SMAX: {  // stl_d_lib.py:191
        const int io = P.ops[o.in0].out_off;
        float acc = PSTL_NEG_INF;
        for (int t = T - 1; t >= 0; --t) {
          acc = pstl_logaddexp(acc, TP(io + t) * tau);
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
// void signal(int p,int t,float g); void pred(int a0,int a1,int t,float g).
// Soft-max weights are recomputed exactly as torch's logsumexp backward: exp(x*tau - lse).
// ----------------------------------------------------------------------------------------
template <class LeafGrad>
PSTL_HD void pstl_interp_bwd(const PstlProgView& P, const float* vt, float* gt, int stride, float tau, int hard,
                             LeafGrad& lg) {
  const int T = P.T;
#define VT(off) vt[(size_t)(off) * stride]
#define GT(off) gt[(size_t)(off) * stride]
  for (int i = P.n_ops - 1; i >= 0; --i) {
    const PstlROp o = P.ops[i];
    const int oo = o.out_off;
    switch (o.op) {
      case PSTL_OP_SIGNAL:
        for (int t = 0; t < o.n_out; ++t) lg.signal(o.a0, t, GT(oo + t));
        break;
      case PSTL_OP_PRED:
        for (int t = 0; t < o.n_out; ++t) lg.pred(o.a0, o.a1, t, GT(oo + t));
        break;
      case PSTL_OP_NEG: {
        const int io = P.ops[o.in0].out_off;
        for (int t = 0; t < o.n_out; ++t) GT(io + t) -= GT(oo + t);
      } break;
      case PSTL_OP_SMIN2:
      case PSTL_OP_SMAX2: {
        const int ia = P.ops[o.in0].out_off, ib = P.ops[o.in1].out_off;
        const float sg = (o.op == PSTL_OP_SMIN2) ? -1.f : 1.f;
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const float a = sg * VT(ia + t), b = sg * VT(ib + t);
          if (hard) {  // torch.max(dim) routes to the first maximal index
            if (a >= b) GT(ia + t) += g; else GT(ib + t) += g;
          } else {
            const float xa = a * tau, xb = b * tau;
            float m = fmaxf(xa, xb);
            if (isinf(m)) m = 0.f;
            const float lse = pstl_lse_finish(m, expf(xa - m) + expf(xb - m));
            // d out/d in = sg * (1/tau) * softmax * tau * sg = softmax weight
            GT(ia + t) += g * expf(xa - lse);
            GT(ib + t) += g * expf(xb - lse);
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
            float bv = -VT(P.ops[P.klist[kb]].out_off + t);
            for (int j = 1; j < k; ++j) {
              const float v = -VT(P.ops[P.klist[kb + j]].out_off + t);
              if (v > bv) { bv = v; bj = j; }
            }
            GT(P.ops[P.klist[kb + bj]].out_off + t) += g;
          } else {
            const float lse = pstl_lse_n(k, [&](int j) { return -VT(P.ops[P.klist[kb + j]].out_off + t) * tau; });
            for (int j = 0; j < k; ++j) {
              const int io = P.ops[P.klist[kb + j]].out_off;
              GT(io + t) += g * expf(-VT(io + t) * tau - lse);
            }
          }
        }
      } break;
      case PSTL_OP_WIN_SMIN:
      case PSTL_OP_WIN_SMAX: {
        const int io = P.ops[o.in0].out_off;
        const float sg = (o.op == PSTL_OP_WIN_SMIN) ? -1.f : 1.f;
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const int lo = pstl_clipi(t + o.a0, 0, T), hi = pstl_clipi(t + o.a1, 0, T);
          if (hi <= lo) continue;
          if (hard) {
            int bj = lo;
            float bv = sg * VT(io + lo);
            for (int j = lo + 1; j < hi; ++j) {
              const float v = sg * VT(io + j);
              if (v > bv) { bv = v; bj = j; }
            }
            GT(io + bj) += g;
          } else {
            const float lse = pstl_lse_n(hi - lo, [&](int j) { return sg * VT(io + lo + j) * tau; });
            for (int j = lo; j < hi; ++j) GT(io + j) += g * expf(sg * VT(io + j) * tau - lse);
          }
        }
      } break;
      case PSTL_OP_PREFIX_SMIN: {
        // out[t] = -LSE_{j<=t}(-x_j tau)/tau  ->  d out[t]/d x_j = exp(-x_j tau - lse_t)
        const int io = P.ops[o.in0].out_off;
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const float lse = -VT(oo + t) * tau;
          for (int j = 0; j <= t; ++j) GT(io + j) += g * expf(-VT(io + j) * tau - lse);
        }
      } break;
      case PSTL_OP_SUFFIX_SMAX: {
        const int io = P.ops[o.in0].out_off;
        for (int t = 0; t < o.n_out; ++t) {
          const float g = GT(oo + t);
          if (g == 0.f) continue;
          const float lse = VT(oo + t) * tau;
          for (int j = t; j < T; ++j) GT(io + j) += g * expf(VT(io + j) * tau - lse);
        }
      } break;
      default:
        break;
    }
  }
#undef VT
#undef GT
}

// stl_program.h — host-side resolution of a postfix STL program: stack simulation, demand
// analysis (how many leading time steps of every trace are actually read) and tape layout.
#pragma once
#include <stdio.h>
#include <string.h>

#include "stl_core.cuh"

// Partial-derivative block of the grad tape (floats, per trajectory), all T-long rows:
//   state: cos, sin, v            -> 3T at offset 0
//   lane l in {0,1,2}: dd/dx, dd/dy, dang/dth -> 3T at offset 3T*(1+l)
//   neighbour: d/dx, d/dy, d/dth  -> 3T at offset 12T
#define PSTL_PART_ROWS 15

static inline int pstl_resolve_program(const pstl_op* ops, int n_ops, int n_signals, int T, int need_t,
                                       PstlProgView* P, char* err, size_t errlen) {
#define FAIL(...)                        \
  do {                                   \
    snprintf(err, errlen, __VA_ARGS__);  \
    return -1;                           \
  } while (0)
  if (n_ops <= 0 || n_ops > PSTL_MAX_OPS) FAIL("program has %d ops (max %d)", n_ops, PSTL_MAX_OPS);
  if (T <= 0 || need_t <= 0 || need_t > T) FAIL("bad T=%d need_t=%d", T, need_t);
  memset(P, 0, sizeof(*P));
  P->n_ops = n_ops;
  P->n_signals = n_signals;
  P->T = T;
  P->need_t = need_t;
  int stack[PSTL_MAX_OPS];
  int sp = 0, kcur = 0;
  for (int i = 0; i < n_ops; ++i) {
    PstlROp& o = P->ops[i];
    o.op = ops[i].op;
    o.a0 = ops[i].a0;
    o.a1 = ops[i].a1;
    o.in0 = o.in1 = -1;
    switch (o.op) {
      case PSTL_OP_SIGNAL:
        if (o.a0 < 0 || o.a0 >= n_signals) FAIL("op %d: signal id %d out of range", i, o.a0);
        break;
      case PSTL_OP_PRED:
        if ((o.a0 & 0xff) >= PSTL_N_BASE_SIGNALS || (o.a1 & 0xff) >= 6) FAIL("op %d: bad predicate", i);
        break;
      case PSTL_OP_NEG:
      case PSTL_OP_WIN_SMIN:
      case PSTL_OP_WIN_SMAX:
      case PSTL_OP_PREFIX_SMIN:
      case PSTL_OP_SUFFIX_SMAX:
        if (sp < 1) FAIL("op %d: stack underflow", i);
        o.in0 = stack[--sp];
        break;
      case PSTL_OP_SMIN2:
      case PSTL_OP_SMAX2:
        if (sp < 2) FAIL("op %d: stack underflow", i);
        o.in1 = stack[--sp];
        o.in0 = stack[--sp];
        break;
      case PSTL_OP_SMIN_K: {
        const int k = o.a0;
        if (k < 1 || sp < k || kcur + k > PSTL_MAX_OPS) FAIL("op %d: bad ListAnd arity %d", i, k);
        o.a1 = kcur;
        for (int j = 0; j < k; ++j) P->klist[kcur + j] = stack[sp - k + j];
        sp -= k;
        kcur += k;
      } break;
      default:
        FAIL("op %d: unknown opcode %d", i, o.op);
    }
    stack[sp++] = i;
  }
  if (sp != 1) FAIL("program leaves %d traces on the stack", sp);
  // demand analysis, top-down
  P->ops[n_ops - 1].n_out = need_t;
  for (int s = 0; s < PSTL_N_BASE_SIGNALS; ++s) {
    P->base_off[s] = -1;
    P->base_need[s] = 0;
  }
  auto want = [&](int idx, int n) {
    if (idx >= 0 && P->ops[idx].n_out < n) P->ops[idx].n_out = n;
  };
  for (int i = n_ops - 1; i >= 0; --i) {
    PstlROp& o = P->ops[i];
    const int n = o.n_out;
    if (n == 0) continue;
    switch (o.op) {
      case PSTL_OP_NEG:
      case PSTL_OP_PREFIX_SMIN:
        want(o.in0, n);
        break;
      case PSTL_OP_SUFFIX_SMAX:
        want(o.in0, T);
        break;
      case PSTL_OP_SMIN2:
      case PSTL_OP_SMAX2:
        want(o.in0, n);
        want(o.in1, n);
        break;
      case PSTL_OP_SMIN_K:
        for (int j = 0; j < o.a0; ++j) want(P->klist[o.a1 + j], n);
        break;
      case PSTL_OP_WIN_SMIN:
      case PSTL_OP_WIN_SMAX:
        want(o.in0, pstl_clipi(n - 1 + o.a1, 0, T));
        break;
      case PSTL_OP_PRED: {
        const int sid = o.a0 & 0xff;
        if (P->base_need[sid] < n) P->base_need[sid] = n;
      } break;
      default:
        break;
    }
  }
  // tape layout: [generic signals P*T][fused base signals][op outputs]
  int cur = n_signals * T;
  for (int s = 0; s < PSTL_N_BASE_SIGNALS; ++s)
    if (P->base_need[s] > 0) {
      P->base_off[s] = cur;
      cur += P->base_need[s];
    }
  for (int i = 0; i < n_ops; ++i) {
    PstlROp& o = P->ops[i];
    if (o.op == PSTL_OP_SIGNAL) {  // staged signals are read in place
      o.out_off = o.a0 * T;
      o.n_out = 0;
    } else if (o.op == PSTL_OP_PRED && i != n_ops - 1) {  // evaluated on the fly by the consumer
      o.out_off = -1;
    } else {
      o.out_off = cur;
      cur += o.n_out;
    }
  }
  P->val_floats = cur;
  P->part_off = 2 * cur;
  P->grad_floats = 2 * cur + PSTL_PART_ROWS * T;
  return 0;
#undef FAIL
}

// ----------------------------------------------------------------------------------------
// Plan: closed form of a resolved driving program of the shape
//     [ListAnd of]  R1_{t in [lo,hi)} [ R2_{t' in [t,T)} ]  X(t)        (only t = 0 consumed)
// with R1, R2 in {soft-min (Always), soft-max (Eventually)} and X a typed predicate leaf or a
// soft-min/max of two typed leaves — the whole spec of build_stl_cache (nusc_train.py:95-140).
// The streaming scorer (score_stream.cuh) evaluates such a program in ONE pass over the rollout with
// two registers per term; anything else runs on the postfix interpreter kernels.
// ----------------------------------------------------------------------------------------
#define PSTL_MAX_TERMS 8
#define PSTL_MAX_TAPES 4
struct PstlLeafC {
  int c;      // canonical base signal: 0 speed, 1 lane distance, 2 lane heading error, 3 neighbour clearance
  int pid;    // stlp column
  int den;    // PSTL_DEN_*
  float sb, sp;  // +-1
};
struct PstlTerm {
  PstlLeafC a, b;
  int pair;    // 0: X = a ; -1: soft-min(a,b) ; +1: soft-max(a,b)
  int inner;   // 0: none ; -1 / +1: soft-min / soft-max over the suffix [t, T)
  int outer;   // -1 / +1: soft-min / soft-max over [lo, hi) evaluated at t = 0
  int lo, hi;  // clipped to [0, T]
  int tape;    // tape column of X(t) for terms with an inner operator, else -1
};
struct PstlPlan {
  int valid, n_terms, listand;
  int lane;       // lane (0 curr, 1 left, 2 right) the distance / heading leaves refer to, -1: none
  int n_tapes;
  int need_pose, need_lane, need_nei;  // leading steps for which the pose / lane search / clearance are read
  int nei_term;   // 0: the clearance signal feeds exactly term 0 and it is  Always_[lo,hi) (nei - q)/den
                  // (soft-min, positive sign, no And/Or, no inner operator; pstl_make_plan moves it to slot 0); -1 otherwise
  PstlTerm terms[PSTL_MAX_TERMS];
};

static inline bool pstl_plan_leaf(const PstlProgView& P, int idx, PstlLeafC* out, int* lane) {
  const PstlROp& o = P.ops[idx];
  if (o.op != PSTL_OP_PRED) return false;
  const int sid = o.a0 & 0xff;
  if (sid == PSTL_SIG_V) out->c = 0;
  else if (sid == PSTL_SIG_NEI) out->c = 3;
  else {
    const int l = (sid - PSTL_SIG_D_CURR) / 2;
    if (*lane >= 0 && *lane != l) return false;  // one lane search per program
    *lane = l;
    out->c = 1 + (sid - PSTL_SIG_D_CURR) % 2;
  }
  out->sb = ((o.a0 >> 8) & 1) ? -1.f : 1.f;
  out->pid = o.a1 & 0xff;
  out->sp = ((o.a1 >> 8) & 1) ? -1.f : 1.f;
  out->den = (o.a1 >> 16) & 0xff;
  return true;
}

static inline bool pstl_plan_x(const PstlProgView& P, int idx, PstlTerm* t, int* lane) {
  const PstlROp& o = P.ops[idx];
  if (o.op == PSTL_OP_SMIN2 || o.op == PSTL_OP_SMAX2) {
    t->pair = (o.op == PSTL_OP_SMIN2) ? -1 : 1;
    return pstl_plan_leaf(P, o.in0, &t->a, lane) && pstl_plan_leaf(P, o.in1, &t->b, lane);
  }
  t->pair = 0;
  t->b = PstlLeafC{0, 0, 0, 1.f, 1.f};
  return pstl_plan_leaf(P, idx, &t->a, lane);
}

static inline bool pstl_plan_term(const PstlProgView& P, int idx, PstlTerm* t, int* lane, int* n_tapes) {
  const PstlROp& o = P.ops[idx];
  if ((o.op != PSTL_OP_WIN_SMIN && o.op != PSTL_OP_WIN_SMAX) || o.n_out != 1) return false;
  t->outer = (o.op == PSTL_OP_WIN_SMIN) ? -1 : 1;
  t->lo = pstl_clipi(o.a0, 0, P.T);
  t->hi = pstl_clipi(o.a1, 0, P.T);
  t->tape = -1;
  const PstlROp& c = P.ops[o.in0];
  if ((c.op == PSTL_OP_WIN_SMIN || c.op == PSTL_OP_WIN_SMAX) && c.a0 == 0 && c.a1 >= P.T) {
    t->inner = (c.op == PSTL_OP_WIN_SMIN) ? -1 : 1;
    if (*n_tapes >= PSTL_MAX_TAPES) return false;
    t->tape = (*n_tapes)++;
    return pstl_plan_x(P, c.in0, t, lane);
  }
  t->inner = 0;
  return pstl_plan_x(P, o.in0, t, lane);
}

static inline void pstl_make_plan(const PstlProgView& P, PstlPlan* pl) {
  memset(pl, 0, sizeof(*pl));
  pl->lane = -1;
  if (P.need_t != 1 || P.n_signals != 0) return;
  const PstlROp& top = P.ops[P.n_ops - 1];
  int lane = -1, n_tapes = 0;
  if (top.op == PSTL_OP_SMIN_K) {
    if (top.a0 > PSTL_MAX_TERMS) return;
    for (int j = 0; j < top.a0; ++j)
      if (!pstl_plan_term(P, P.klist[top.a1 + j], &pl->terms[j], &lane, &n_tapes)) return;
    pl->n_terms = top.a0;
    pl->listand = 1;
  } else {
    if (!pstl_plan_term(P, P.n_ops - 1, &pl->terms[0], &lane, &n_tapes)) return;
    pl->n_terms = 1;
  }
  pl->lane = lane;
  pl->n_tapes = n_tapes;
  pl->need_lane = 0;
  if (lane >= 0) {
    const int sd = PSTL_SIG_D_CURR + 2 * lane;
    pl->need_lane = P.base_need[sd] > P.base_need[sd + 1] ? P.base_need[sd] : P.base_need[sd + 1];
  }
  pl->need_nei = P.base_need[PSTL_SIG_NEI];
  pl->need_pose = 0;
  for (int b = 0; b < PSTL_N_BASE_SIGNALS; ++b) pl->need_pose = P.base_need[b] > pl->need_pose ? P.base_need[b] : pl->need_pose;
  // the value-aware neighbour bound of the streaming scorer applies only to a lone  G(nei - q)  term
  int uses = 0, idx = -1;
  for (int k = 0; k < pl->n_terms; ++k) {
    const PstlTerm& t = pl->terms[k];
    const bool a_nei = t.a.c == 3, b_nei = t.pair != 0 && t.b.c == 3;
    if (a_nei || b_nei) {
      ++uses;
      if (a_nei && t.pair == 0 && t.inner == 0 && t.outer == -1 && t.a.sb == 1.f && t.lo == 0) idx = k;
    }
  }
  pl->nei_term = (uses == 1) ? idx : -1;
  if (pl->nei_term > 0) {  // keep it in slot 0 (ListAnd is symmetric): the scorer reads its accumulator at a fixed index
    const PstlTerm t0 = pl->terms[0];
    pl->terms[0] = pl->terms[pl->nei_term];
    pl->terms[pl->nei_term] = t0;
    pl->nei_term = 0;
  }
  pl->valid = 1;
}

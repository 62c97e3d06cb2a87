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

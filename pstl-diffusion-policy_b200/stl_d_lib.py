"""Drop-in for the reference's ``stl_d_lib`` (reference stl_d_lib.py:1-203).

Same node classes and call convention — ``node(x, tau, d=None) -> (N,T)`` robustness trace,
differentiable w.r.t. the tensors the AP leaves read — but a formula tree is not evaluated
node by node: it is flattened once into a postfix op program (``compile_formula``) that one
CUDA kernel interprets per trajectory (csrc/stl_core.cuh), with a hand-written reverse pass.

AP leaves stay arbitrary Python callables: they are evaluated with PyTorch into a (N,P,T)
signal tensor and only the temporal / boolean program runs in the kernel.  Leaves built with
``AP.predicate`` additionally carry a typed description of a driving predicate so that
``nusc_train.compute_stl_dense`` can run rollout, predicates and formula in one fused kernel.
"""
import torch

from . import native as _nv


def clip(x, a, b):
    return max(min(x, b), a)


class STLFormula:
    def __init__(self, ts=None, te=None, node=None, lhs=None, rhs=None, lists=None, operator=None):
        self.ts, self.te = ts, te
        self.node, self.lhs, self.rhs, self.lists = node, lhs, rhs, lists
        self.operator = operator
        self.format = "symbol"

    def __call__(self, x, tau, d=None):
        out = evaluate(self, x, tau, d)
        if d is not None and "idx" in d:
            print(type(self).__name__, "output", out[d["idx"]])
        return out

    def __str__(self):
        ops = self.operator[self.format]
        if self.ts is not None:
            ops = "%s[%d:%d]" % (ops, self.ts, self.te + 1)
        if self.node is not None:
            return "%s (%s)" % (ops, self.node)
        if self.lhs is not None:
            return "(%s) %s (%s)" % (self.lhs, ops, self.rhs)
        if self.lists is not None:
            return "%s {%s}" % (ops, ",".join("|%s|" % c for c in self.lists))
        raise NotImplementedError

    def children(self):
        return [self.node] if self.node is not None else [self.lhs, self.rhs]

    def update_format(self, format):
        self.format = format
        for child in self.children():
            if hasattr(child, "update_format"):
                child.update_format(format)

    def build(self, s):
        raise NotImplementedError


class AP:
    n_aps = 0

    def __init__(self, expression, comment=None):
        self.expression = expression
        self.comment = comment
        self.apid = AP.n_aps
        AP.n_aps += 1
        self.pred = None  # (a0, a1) of PSTL_OP_PRED for typed driving predicates

    @classmethod
    def predicate(cls, expression, signal, signal_neg, param, param_neg, den=_nv.DEN_ONE, comment=None):
        """AP that also says what it is: (±base[signal] ± stlp[param]) / den (include/pstl.h)."""
        ap = cls(expression, comment)
        ap.pred = (int(signal) | (int(bool(signal_neg)) << 8),
                   int(param) | (int(bool(param_neg)) << 8) | (int(den) << 16))
        return ap

    def __call__(self, x, tau, d=None):
        s = self.expression(x)
        if d is not None and "idx" in d:
            print(self.__str__(), "out", s[d["idx"]])
        return s

    def __str__(self):
        return "AP%d" % self.apid if self.comment is None else self.comment


class And(STLFormula):
    def __init__(self, lhs, rhs):
        super().__init__(lhs=lhs, rhs=rhs, operator={"symbol": "&", "word": "AND"})


class ListAnd(STLFormula):
    def __init__(self, lists):
        super().__init__(lists=lists, operator={"symbol": "&", "word": "AND"})

    def __call__(self, x, tau, d=None, full=False):
        if not full:
            return evaluate(self, x, tau, d)
        # full=True also returns the stacked child traces (reference stl_d_lib.py:109-110)
        v = torch.stack([c(x, tau, d) for c in self.lists], dim=1)
        probe = ListAnd([AP((lambda k: (lambda s: s[:, k]))(k)) for k in range(len(self.lists))])
        return evaluate(probe, v, tau, d), v

    def children(self):
        return list(self.lists)


class Or(STLFormula):
    def __init__(self, lhs, rhs):
        super().__init__(lhs=lhs, rhs=rhs, operator={"symbol": "|", "word": "OR"})


class Not(STLFormula):
    def __init__(self, node):
        super().__init__(node=node, operator={"symbol": "¬", "word": "NOT"})


class Imply(STLFormula):
    def __init__(self, lhs, rhs):
        super().__init__(lhs=lhs, rhs=rhs, operator={"symbol": "->", "word": "IMPLY"})
        self.eval = Or(Not(self.lhs), self.rhs)


class Eventually(STLFormula):
    def __init__(self, ts, te, node):
        super().__init__(ts=ts, te=te, node=node, operator={"symbol": "♢", "word": "EVENTUALLY"})


class Always(STLFormula):
    def __init__(self, ts, te, node):
        super().__init__(ts=ts, te=te, node=node, operator={"symbol": "◻", "word": "ALWAYS"})


class Once(STLFormula):
    def __init__(self, ts, te, node):
        super().__init__(ts=ts, te=te, node=node, operator={"symbol": "O", "word": "ONCE"})
        assert ts < 0 and te >= ts and te <= 0


class UntimedUntil(STLFormula):
    def __init__(self, lhs, rhs):
        super().__init__(lhs=lhs, rhs=rhs, operator={"symbol": "U", "word": "UNTIL"})


class Until(STLFormula):
    def __init__(self, ts, te, lhs, rhs):
        super().__init__(ts=ts, te=te, lhs=lhs, rhs=rhs, operator={"symbol": "U", "word": "UNTIL"})
        if ts == 0:
            self.eval = UntimedUntil(lhs, rhs)
        else:
            self.eval = And(Eventually(ts, te, rhs), Always(0, ts, UntimedUntil(lhs, rhs)))


# ---------------------------------------------------------------------------------------
# tree -> postfix program
# ---------------------------------------------------------------------------------------

def compile_formula(node, fused=False):
    """Return (ops, leaves): ops = list of (opcode, a0, a1); leaves = AP objects in signal-id order.
    With ``fused=True`` every leaf must be a typed predicate and ops contain PSTL_OP_PRED."""
    ops, leaves = [], []

    def emit(n):
        if isinstance(n, AP):
            if fused:
                if n.pred is None:
                    raise ValueError("formula has an untyped AP leaf: cannot fuse")
                ops.append((_nv.OP_PRED, n.pred[0], n.pred[1]))
            else:
                for i, l in enumerate(leaves):
                    if l is n:
                        ops.append((_nv.OP_SIGNAL, i, 0))
                        return
                leaves.append(n)
                ops.append((_nv.OP_SIGNAL, len(leaves) - 1, 0))
        elif isinstance(n, Not):
            emit(n.node)
            ops.append((_nv.OP_NEG, 0, 0))
        elif isinstance(n, (Imply, Until)):
            emit(n.eval)
        elif isinstance(n, And):
            emit(n.lhs)
            emit(n.rhs)
            ops.append((_nv.OP_SMIN2, 0, 0))
        elif isinstance(n, Or):
            emit(n.lhs)
            emit(n.rhs)
            ops.append((_nv.OP_SMAX2, 0, 0))
        elif isinstance(n, ListAnd):
            for c in n.lists:
                emit(c)
            ops.append((_nv.OP_SMIN_K, len(n.lists), 0))
        elif isinstance(n, Always):
            emit(n.node)
            ops.append((_nv.OP_WIN_SMIN, n.ts, n.te))
        elif isinstance(n, (Eventually, Once)):
            emit(n.node)
            ops.append((_nv.OP_WIN_SMAX, n.ts, n.te))
        elif isinstance(n, UntimedUntil):
            # stack([rs, inf_ls]) -> soft-min -> suffix soft-max (reference stl_d_lib.py:187-191)
            emit(n.rhs)
            emit(n.lhs)
            ops.append((_nv.OP_PREFIX_SMIN, 0, 0))
            ops.append((_nv.OP_SMIN2, 0, 0))
            ops.append((_nv.OP_SUFFIX_SMAX, 0, 0))
        else:
            raise TypeError("not an STL node: %r" % (n,))

    emit(node)
    return ops, leaves


_prog_cache = {}


def get_program(ops, n_signals, T, need_t):
    key = (tuple(ops), n_signals, T, need_t, torch.cuda.current_device())
    p = _prog_cache.get(key)
    if p is None:
        p = _nv.Program(ops, n_signals, T, need_t)
        _prog_cache[key] = p
    return p


class _StlSignals(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sig, prog, tau, hard):
        N, P, T = sig.shape
        out = torch.empty((N, prog.need_t), dtype=torch.float32, device=sig.device)
        L = _nv.lib()
        ws = _nv.workspace(L.pstl_stl_workspace_bytes(prog.h, N, 0), sig.device, "stl")
        _nv.check(L.pstl_stl_eval_signals(prog.h, _nv.fptr(sig), N, P, T, _nv.C.c_float(tau), int(hard),
                                          _nv.fptr(out), None, _nv.ptr(ws), _nv.stream()), "pstl_stl_eval_signals")
        ctx.save_for_backward(sig)
        ctx.prog, ctx.tau, ctx.hard = prog, tau, hard
        return out

    @staticmethod
    def backward(ctx, gout):
        (sig,) = ctx.saved_tensors
        N, P, T = sig.shape
        gsig = torch.empty_like(sig)
        L = _nv.lib()
        ws = _nv.workspace(L.pstl_stl_workspace_bytes(ctx.prog.h, N, 1), sig.device, "stl")
        _nv.check(L.pstl_stl_eval_signals_bwd(ctx.prog.h, _nv.fptr(sig), _nv.fptr(_nv.f32(gout)), N, P, T,
                                              _nv.C.c_float(ctx.tau), int(ctx.hard), _nv.fptr(gsig), _nv.ptr(ws),
                                              _nv.stream()), "pstl_stl_eval_signals_bwd")
        return gsig, None, None, None


def evaluate(node, x, tau, d=None, need_t=None):
    """Robustness trace of ``node`` on ``x``: AP leaves in PyTorch, the rest in one kernel."""
    ops, leaves = compile_formula(node)
    vals = [leaf(x, tau, d) for leaf in leaves]
    for v in vals:
        _nv.require_cuda(v, "AP output")
    shape = torch.broadcast_shapes(*[v.shape for v in vals])
    if len(shape) != 2:
        raise ValueError("AP leaves must produce (N,T) signals, got %s" % (tuple(shape),))
    sig = torch.stack([v.to(torch.float32).expand(shape) for v in vals], dim=1).contiguous()
    T = shape[1]
    hard = bool(d is not None and d.get("hard"))
    prog = get_program(ops, len(leaves), T, T if need_t is None else need_t)
    return _StlSignals.apply(sig, prog, float(tau), hard)


# the reference's module-level soft reductions, kept for callers that import them
def softmax(x, tau, d=None, dim=1):
    if x.shape[1] == 0:
        return torch.ones(x.shape[0], 1, device=x.device) * -float("inf")
    if d is not None and d.get("hard"):
        return torch.max(x, dim=dim, keepdim=True)[0]
    return torch.logsumexp(x * tau, dim=dim, keepdim=True) / tau


def softmin(x, tau, d=None, dim=1):
    if x.shape[1] == 0:
        return torch.ones(x.shape[0], 1, device=x.device) * -float("inf")
    return -softmax(-x, tau, d, dim)


def softmax_pairs(x, y, tau, d=None):
    return softmax(torch.stack([x, y], dim=1), tau, d).squeeze(1)


def softmin_pairs(x, y, tau, d=None):
    return -softmax_pairs(-x, -y, tau, d)

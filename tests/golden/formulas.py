"""Formula recipes shared by the golden generator (reference nodes), the oracle (tuples) and the
product (pstl_b200.stl_d_lib nodes).  ``ns`` is any namespace exposing the node constructors
AP, And, Or, Not, Imply, ListAnd, Eventually, Always, Once, UntimedUntil, Until."""


class TupleNS:
    """Constructor namespace that builds the oracle's nested-tuple formulas."""
    AP = staticmethod(lambda fn, comment=None: ("ap", fn))
    And = staticmethod(lambda l, r: ("and", l, r))
    Or = staticmethod(lambda l, r: ("or", l, r))
    Not = staticmethod(lambda n: ("not", n))
    Imply = staticmethod(lambda l, r: ("imply", l, r))
    ListAnd = staticmethod(lambda lst: ("listand", list(lst)))
    Eventually = staticmethod(lambda ts, te, n: ("eventually", ts, te, n))
    Always = staticmethod(lambda ts, te, n: ("always", ts, te, n))
    Once = staticmethod(lambda ts, te, n: ("once", ts, te, n))
    UntimedUntil = staticmethod(lambda l, r: ("untimed_until", l, r))
    Until = staticmethod(lambda ts, te, l, r: ("until", ts, te, l, r))


def recipes(ns):
    """name -> formula over signals x['a'], x['b'], x['c'] (each (N,T))."""
    a = lambda: ns.AP(lambda x: x["a"], comment="a")
    b = lambda: ns.AP(lambda x: x["b"], comment="b")
    c = lambda: ns.AP(lambda x: x["c"], comment="c")
    return {
        "ap": a(),
        "not": ns.Not(a()),
        "and": ns.And(a(), b()),
        "or": ns.Or(a(), b()),
        "imply": ns.Imply(a(), b()),
        "listand3": ns.ListAnd([a(), b(), c()]),
        "always_0_3": ns.Always(0, 3, a()),
        "always_0_T": ns.Always(0, 64, a()),
        "always_2_5": ns.Always(2, 5, a()),
        "eventually_1_4": ns.Eventually(1, 4, a()),
        "eventually_0_T": ns.Eventually(0, 64, b()),
        "once_m3_0": ns.Once(-3, 0, a()),
        "untimed_until": ns.UntimedUntil(a(), b()),
        "until_0_8": ns.Until(0, 8, a(), b()),
        "until_2_5": ns.Until(2, 5, a(), b()),
        "ev_alw_and": ns.Eventually(0, 4, ns.Always(0, 8, ns.And(a(), b()))),
        "alw_ev": ns.Always(0, 6, ns.Eventually(0, 3, c())),
        "nested_mix": ns.ListAnd([ns.Always(0, 64, ns.Or(a(), ns.Not(b()))),
                                  ns.Eventually(0, 5, ns.Always(0, 64, ns.And(b(), c()))),
                                  ns.Imply(ns.Always(1, 3, a()), ns.Eventually(0, 2, c()))]),
        "empty_window": ns.Always(0, 0, a()),
    }

"""Formula.eval on the device (lk_formula_eval; Formula.scala:32-69, ASTUtils.scala:50-64,87-89) against the oracle's
restatement (`formula_eval`, `constant_expr_eval`, `base_expr_eval`): per timestamp, both sides' rows become maps
groupKey -> (ts, value, tags) in row order (a later row replaces an earlier one), the oracle combines the maps, and the device
must produce the same (timestamp, group key) -> (value bits, tags) -- add / sub / mul / div, missing sides, zero divisors,
constants on either side, no group-bys ("default"), a group-by that does not exist."""
import math
import struct
from collections import defaultdict

import numpy as np
import pytest

import helpers as H
import lakeside_oracle as lo
from lakeside_b200 import synth

pytestmark = pytest.mark.gpu
STEP = 10000


def _expr(svcs, group_bys, names=None):
    be = synth.c2_base_expr()
    f = {"k": synth.TAG_SERVICE, "v": list(svcs), "op": "in", "dataType": "string", "extracted": False, "computed": False}
    if names:
        f = {"q1": f, "q2": {"k": synth.NAME, "v": list(names), "op": "in", "dataType": "string", "extracted": False, "computed": False}, "op": "and"}
    be["filter"] = f
    be["chart"]["groupBys"] = list(group_bys)
    return be


class _Side:
    """One BaseExpr side evaluated on the GPU: the query stays alive (its reduced rows are the formula's input)."""

    def __init__(self, be, paths, idx, agg, chart="line", metric="gauge", aggs=synth.C2_AGGREGATES):
        from lakeside_b200 import api

        self.be, self.agg, self.chart, self.metric = be, agg, chart, metric
        self.q = api.Query(H.request_json(be, idx, STEP), aggregates=aggs)
        for p in paths:
            self.q.add_segment_file(p)
        self.q.prepare()
        self.q.execute()
        self.res = self.q.finalize()
        self.values = self.q.eval(self.res.num_rows, agg, chart, metric)
        self.dps = self.res.to_data_points({})
        self.group_bys = list(be["chart"]["groupBys"])

    def spec(self):
        return (self.q, {"aggregation": self.agg, "chartType": self.chart, "metricType": self.metric, "groupBys": self.group_bys})

    def maps(self):
        """{ts: {groupKey: (ts, value, tags)}} in row order: BaseExpr.eval's map per SketchGroup."""
        keys = set(self.group_bys)
        out = defaultdict(dict)
        for d, v in zip(self.dps, self.values.tolist()):
            out[d.timestamp]["default" if not keys else lo.to_group_by_key(keys, d.tags)] = (d.timestamp, float(v), d.tags)
        return out

    def close(self):
        self.res.close()
        self.q.close()


def _bits(v):
    return "nan" if v != v else struct.pack("<d", v)


def _check(op, e1, e2, final_keys):
    from lakeside_b200 import api

    sides = [x for x in (e1, e2) if isinstance(x, _Side)]
    m = [x.maps() if isinstance(x, _Side) else None for x in (e1, e2)]
    all_ts = sorted(set().union(*[set(mm) for mm in m if mm is not None]))
    want = {}
    for ts in all_ts:
        maps = []
        for x, mm in zip((e1, e2), m):
            if mm is not None:
                maps.append(mm.get(ts, {}))
            else:  # ConstantExpr: keyed by the final grouping over every sketch input of the group (the other side's rows, in row order)
                other = sides[0]
                inputs = [lo.SketchInput(ts, d.tags, {}) for d in other.dps if d.timestamp == ts] if final_keys else []
                maps.append(lo.constant_expr_eval(float(x), final_keys, ts, inputs))
        for k, (t, v, tags) in lo.formula_eval(op, maps[0], maps[1]).items():
            want[(t, k)] = (_bits(v), tuple(sorted(tags.items())))
    ts, val, side, row = api.formula_eval(op, e1.spec() if isinstance(e1, _Side) else e1, e2.spec() if isinstance(e2, _Side) else e2)
    assert list(ts) == sorted(ts), "results come in timestamp order"
    got = {}
    for t, v, s, r in zip(ts.tolist(), val.tolist(), side.tolist(), row.tolist()):
        src = e1 if s == 1 else e2
        assert isinstance(src, _Side)
        d = src.dps[r]
        assert d.timestamp == t
        keys = set(final_keys) if not isinstance(e1, _Side) or not isinstance(e2, _Side) else set(src.group_bys)
        k = "default" if not keys else lo.to_group_by_key(keys, d.tags)
        assert (t, k) not in got
        tags = {} if s == 0 else d.tags  # side 0: a constant e1 without group-bys has no tags (ASTUtils.scala:51-55)
        got[(t, k)] = (_bits(v), tuple(sorted(tags.items())))
    assert len(got) == len(want), (op, len(got), len(want))
    assert got == want
    return len(want)


@pytest.fixture(scope="module")
def sides():
    from lakeside_b200 import api

    api.init()
    spec = synth.SynthSpec(dataset="metrics", rows=60000, n_names=3, cards=(8, 4, 4, 2))
    _, paths = H.dataset("eval_metrics", spec, 2)
    svc = synth.tag_values(spec.prefixes[0], spec.cards[0])
    g = [synth.TAG_NAMESPACE, synth.TAG_ZONE]
    out = {
        # overlapping but different key sets: some (ts, key) exist on one side only
        "a": _Side(_expr(svc[:4], g, names=["metric_000"]), paths, [0, 1], "sum"),
        "b": _Side(_expr(svc[2:6], g, names=["metric_000"]), paths, [0, 1], "count"),
        # all names: several rows share one (ts, group key) -> the later row replaces the earlier one
        "dup": _Side(_expr(svc[:3], g), paths, [0, 1], "max"),
        "nogroup1": _Side(_expr(svc[:2], []), paths, [0, 1], "sum"),
        "nogroup2": _Side(_expr(svc[1:3], []), paths, [0, 1], "avg", chart="rate", metric="count"),
        # one group-by does not exist in the files: it reads as "" in the key
        "ghost": _Side(_expr(svc[:4], [synth.TAG_NAMESPACE, "no.such.tag"], names=["metric_000"]), paths, [0, 1], "sum"),
        "zone_only": _Side(_expr(svc[:4], [synth.TAG_ZONE, "zz.absent"], names=["metric_000"]), paths, [0, 1], "min"),
    }
    yield out
    for s in out.values():
        s.close()


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_formula_two_base_exprs(sides, op):
    n = _check(op, sides["a"], sides["b"], None)
    assert n > 100
    _check(op, sides["b"], sides["a"], None)
    _check(op, sides["dup"], sides["b"], None)
    _check(op, sides["nogroup1"], sides["nogroup2"], None)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_formula_constants(sides, op):
    g = sides["a"].group_bys
    for c in (2.5, 0.0, -1.0):
        _check(op, sides["a"], c, g)
        _check(op, c, sides["a"], g)
        _check(op, sides["dup"], c, sides["dup"].group_bys)
    _check(op, sides["nogroup1"], 4.0, [])
    _check(op, 4.0, sides["nogroup1"], [])


def test_formula_missing_group_by_and_zero_divisor(sides):
    # "no.such.tag" reads as "" in one side's key: it meets the other side only in rows whose zone is NULL (also ""), so
    # `add` mostly keeps both sides' rows apart and `mul` combines just those few
    n_add = _check("add", sides["ghost"], sides["a"], None)
    n_mul = _check("mul", sides["ghost"], sides["a"], None)
    assert 0 < n_mul < n_add
    # both sides have an absent group-by, at different positions of the sorted key
    _check("add", sides["ghost"], sides["zone_only"], None)
    # count side holds zeros nowhere, but min(rollup_min) may: division drops exactly those
    _check("div", sides["a"], sides["zone_only"], None)


def test_formula_errors(sides):
    from lakeside_b200 import api

    with pytest.raises(api.LakesideError):
        api.formula_eval("pow", sides["a"].spec(), sides["b"].spec())
    with pytest.raises(api.LakesideError):  # different numbers of group-bys
        api.formula_eval("add", sides["a"].spec(), sides["nogroup1"].spec())
    with pytest.raises(api.LakesideError):
        api.formula_eval("add", 1.0, 2.0)

"""BaseExpr.eval on the device (lk_query_eval) against the oracle's restatement of BaseExpr.scala:665-695 / :47-95 and
ASTUtils.scala:190-219 (`base_expr_eval`, `get_from_sketch`, `transformer`): one value per reduced row, bit-exact --
the arithmetic is one IEEE division and/or one multiplication per row on both sides."""
import math
import struct

import numpy as np
import pytest

import helpers as H
import lakeside_oracle as lo
from lakeside_b200 import synth

pytestmark = pytest.mark.gpu

AGGS = ["sum", "count", "min", "max"]  # synth.C2_AGGREGATES: sum(rollup_sum) sum(rollup_count) min(rollup_min) max(rollup_max)


def _same_bits(a: float, b: float) -> bool:
    if a != a and b != b:
        return True
    return struct.pack("<d", a) == struct.pack("<d", b)


def _rows_and_eval(rq, paths, aggs, combos, path="auto"):
    from lakeside_b200 import api

    api.init()
    with api.Query(rq, aggregates=aggs, path=path) as q:
        for p in paths:
            q.add_segment_file(p)
        q.prepare()
        q.execute()
        res = q.finalize()
        try:
            vals = [np.array(v, copy=True) for v in res.values]
            n = len(res.ts)
            got = {c: q.eval(n, *c) for c in combos}
        finally:
            res.close()
    return n, vals, got


@pytest.mark.parametrize("path", ["dense", "hash", "records"])
def test_eval_metrics_all_transforms(path):
    spec = synth.SynthSpec(dataset="metrics", rows=60000, n_names=3, cards=(8, 4, 4, 2))
    _, paths = H.dataset("eval_metrics", spec, 2)
    step = 10000
    rq = H.request_json(synth.c2_base_expr(), [0, 1], step)
    combos = [(agg, chart, metric) for agg in AGGS + ["avg", "p99"] for chart, metric in
              [("line", "gauge"), ("count", "rate"), ("rate", "count"), ("rate", "rate")]]
    # map-sketch keys of the four aggregates: sum, count (= sum of rollup_count on pre-rolled metrics), min, max
    n, vals, got = _rows_and_eval(rq, paths, synth.C2_AGGREGATES, combos, path)
    assert n > 100
    for (agg, chart, metric), g in got.items():
        fn = lo.transformer(chart, metric, lo.METRICS, step)
        assert len(g) == n
        for i in range(n):
            sketch = {a: float(vals[k][i]) for k, a in enumerate(AGGS)}
            want = fn(lo.get_from_sketch(sketch, agg))
            assert _same_bits(float(g[i]), want), (agg, chart, metric, i, g[i], want)


def test_eval_events_rate_and_zero_step_seconds():
    spec = synth.SynthSpec(dataset="logs", rows=50000)
    _, paths = H.dataset("eval_logs", spec, 1)
    for step in (60000, 500):  # 500 ms: step / 1000 == 0 -> division by zero follows IEEE (inf / nan), as on the JVM
        rq = H.request_json(synth.c1_base_expr(), [0], step)
        # c1: chart aggregation "sum" over _cardinalhq.value -> the map sketch has the single key "sum"
        n, vals, got = _rows_and_eval(rq, paths, None, [("sum", "rate", "gauge"), ("sum", "count", "gauge"), ("avg", "rate", "gauge"),
                                                        ("count", "rate", "gauge")])
        assert n > 0
        fn_rate = lo.transformer("rate", "gauge", "logs", step)
        for i in range(n):
            sketch = {"sum": float(vals[0][i])}
            assert _same_bits(float(got[("sum", "rate", "gauge")][i]), fn_rate(lo.get_from_sketch(sketch, "sum")))
            assert _same_bits(float(got[("sum", "count", "gauge")][i]), float(vals[0][i]))
            assert math.isnan(float(got[("avg", "rate", "gauge")][i]))    # no `count` in the sketch: sum / NaN
            assert math.isnan(float(got[("count", "rate", "gauge")][i]))  # key absent from the sketch

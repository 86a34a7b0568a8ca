"""lk_result_to_sse (native serialisation of result rows into the reference's stream elements) checked
through the oracle's restatement of the CONSUMER (SegmentSequencer.decode): what the library writes must decode to exactly
the SketchInputs the reference's own pipeline (toDataPoint -> PushDownAggregatorStage -> dataPointResponseToSSE) yields."""
import math
import struct

import pytest

import helpers as H
import lakeside_oracle as lo
from lakeside_b200 import synth

pytestmark = pytest.mark.gpu


def _bits(x):
    return struct.pack("<d", x)


def _key(s):
    return (s.timestamp, tuple(sorted(s.tags.items())))


def test_sse_single_aggregate_matches_reference_pipeline():
    from lakeside_b200 import api

    spec = synth.SynthSpec(dataset="metrics", rows=50000, n_names=3, cards=(8, 4, 4, 2), null_frac=0.2)
    _, paths = H.dataset("sse_metrics", spec, 2)
    be = synth.c2_base_expr()
    rq = H.request_json(be, [0, 1], 10000)
    req = lo.push_down_request_from_json(rq)
    query_tags = {"_cardinalhq.name": "fallback"}
    want = lo.push_down_aggregator_stage(req, lo.to_data_points(lo.evaluate_glob(req, paths), query_tags))
    api.init()
    res = api.eval_glob(rq, list(paths))
    try:
        n = res.num_rows
        wire = res.to_sse([req.globalAgg], query_tags)
        # the whole result in one call equals the concatenation of row ranges
        assert wire == res.to_sse([req.globalAgg], query_tags, 0, n // 2) + res.to_sse([req.globalAgg], query_tags, n // 2, n)
    finally:
        res.close()
    assert wire.count(b"\r\n\r\n") == n == len(want)
    got = lo.sse_decode(wire)
    assert [g.timestamp for g in got] == sorted(w.timestamp for w in want)  # ORDER BY timestamp survives
    wmap = {}
    for w in want:
        wmap.setdefault(_key(w), []).append(w)
    for g in got:
        (w,) = wmap[_key(g)]  # same (timestamp, tags) cell, exactly once
        assert set(g.sketch) == set(w.sketch) == {req.globalAgg}
        a, b = g.sketch[req.globalAgg], w.sketch[req.globalAgg]
        assert (math.isnan(a) and math.isnan(b)) or abs(a - b) <= H.SUM_RTOL * max(abs(a), abs(b)), (a, b)
        assert g.sketchType == "map"


def test_sse_multi_aggregate_values_round_trip_bit_exact():
    from lakeside_b200 import api

    spec = synth.SynthSpec(dataset="metrics", rows=40000, n_names=2, cards=(4, 4, 2, 2), extra_nan_inf=True)
    _, paths = H.dataset("sse_naninf", spec, 1)
    rq = H.request_json(synth.c2_base_expr(), [0], 10000)
    api.init()
    with api.Query(rq, aggregates=synth.C2_AGGREGATES) as q:
        q.add_segment_file(paths[0])
        q.prepare()
        q.execute()
        res = q.finalize()
        try:
            vals = [list(map(float, v)) for v in res.values]
            ts = list(map(int, res.ts))
            wire = res.to_sse(["sum", "count", "min", "max"])
        finally:
            res.close()
    got = lo.sse_decode(wire)
    assert len(got) == len(ts) > 0
    seen_nonfinite = False
    for i, g in enumerate(got):
        assert g.timestamp == ts[i]
        for k, name in enumerate(["sum", "count", "min", "max"]):
            a, b = g.sketch[name], vals[k][i]
            seen_nonfinite |= not math.isfinite(b)
            assert (math.isnan(a) and math.isnan(b)) or _bits(a) == _bits(b), (i, name, a, b)
    assert seen_nonfinite  # the dataset carries NaN / Inf values: they travel as "NaN" / "Infinity" strings

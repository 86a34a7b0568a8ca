"""exact_sums: the fixed-order (segment order, then row order) summation option gives double sums that are BIT-identical
to the oracle's sequential left fold (north-star: 'a fixed-order reduction option that gives bit-exact sums')."""
import struct

import pytest

import helpers as H
from lakeside_b200 import synth

pytestmark = pytest.mark.gpu


def _bits(x):
    return struct.pack("<d", x)


def _run(rq, paths, aggs, path, exact):
    from lakeside_b200 import api

    api.init()
    with api.Query(rq, aggregates=aggs, path=path, exact_sums=exact) as q:
        for p in paths:
            q.add_segment_file(p)
        q.prepare()
        q.execute()
        res = q.finalize()
        out = H.canon_from_gpu(res)
        res.close()
        return out


@pytest.mark.parametrize("path", ["dense", "hash"])
def test_exact_sums_bit_identical_small_groups(path):
    # few groups => thousands of lognormal addends per cell: atomics would differ in the last bits
    spec = synth.SynthSpec(dataset="metrics", rows=120000, n_names=2, cards=(16, 2, 2, 2))
    _, paths = H.dataset("exact_small", spec, 3)
    be = synth.c2_base_expr()
    be["filter"] = {"k": synth.TAG_SERVICE, "v": ["svc-0[0-7]"], "op": "regex"}
    rq = H.request_json(be, [0, 1, 2], 10000)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    got = _run(rq, paths, synth.C2_AGGREGATES, path, True)
    assert set(got["rows"]) == set(want["rows"])
    for k, w in want["rows"].items():
        g = got["rows"][k]
        assert _bits(g[0]) == _bits(w[0]), (k, g[0], w[0])  # sum(rollup_sum): lognormal doubles
        assert _bits(g[1]) == _bits(w[1])                    # sum(rollup_count)
        assert g[2] == w[2] and g[3] == w[3]
    # and the default (atomic) mode really is order-dependent on this data, i.e. the option does something
    loose = _run(rq, paths, synth.C2_AGGREGATES, path, False)
    H.assert_same(loose, want, ["sum", "sum", "min", "max"], "default mode within 1e-12")
    n_diff = sum(_bits(loose["rows"][k][0]) != _bits(w[0]) for k, w in want["rows"].items())
    assert n_diff > 0, "atomic-order sums happened to equal the sequential fold in every cell: the test data no longer shows what exact_sums is for"


def test_exact_sums_c1_events():
    spec = synth.SynthSpec(dataset="logs", rows=400000)
    _, paths = H.dataset("exact_c1", spec, 2)
    rq = H.request_json(synth.c1_base_expr(), [0, 1], 60000)
    want = H.oracle_single(rq, paths)
    from lakeside_b200 import api

    api.init()
    with api.Query(rq, exact_sums=True) as q:
        for p in paths:
            q.add_segment_file(p)
        q.prepare()
        q.execute()
        res = q.finalize()
        got = H.canon_from_gpu(res)
        res.close()
    assert set(got["rows"]) == set(want["rows"])
    for k, w in want["rows"].items():
        assert _bits(got["rows"][k][0]) == _bits(w[0]), (k, got["rows"][k][0], w[0])

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


@pytest.mark.parametrize("path", ["dense", "hash", "records"])
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


def test_exact_sums_record_path_is_what_the_planner_picks():
    """A large group space + exact_sums stays on the record path (round 1 silently fell back to the hash table)."""
    spec = synth.SynthSpec(dataset="metrics", rows=100000)
    _, paths = H.dataset("exact_rec_auto", spec, 2)
    rq = H.request_json(synth.c2_base_expr(), [0, 1], 10000)
    from lakeside_b200 import api

    api.init()
    with api.Query(rq, aggregates=synth.C2_AGGREGATES, exact_sums=True) as q:
        for p in paths:
            q.add_segment_file(p)
        q.prepare()
        assert q.info["path"] == "records"
        q.execute()
        res = q.finalize()
        got = H.canon_from_gpu(res)
        res.close()
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    assert set(got["rows"]) == set(want["rows"])
    for k, w in want["rows"].items():
        assert all(_bits(a) == _bits(b) for a, b in zip(got["rows"][k], w)), (k, got["rows"][k], w)
    assert got["ts_order"] == sorted(got["ts_order"])


@pytest.mark.parametrize("world", [2, 3])
def test_exact_sums_sharded_fixed_order(world):
    """Sharded fixed-order mode: every record carries its global row sequence (shards hold contiguous blocks of
    the request's segments; seq_offset = rows of the shards before), the owner rank folds in (cell, sequence) order: the sums of
    `world` ranks are bit-identical to the oracle's sequential fold over all segments -- whichever rank scanned what."""
    import json

    from lakeside_b200 import api

    api.init()
    spec = synth.SynthSpec(dataset="metrics", rows=90000, n_names=2, cards=(16, 3, 2, 2))  # few groups: thousands of addends per cell
    per = 2
    _, paths = H.dataset("exact_sharded", spec, per * world)
    be = synth.c2_base_expr()
    be["filter"] = {"k": synth.TAG_SERVICE, "v": ["svc-0[0-7]"], "op": "regex"}
    full = synth.push_down_request(be, list(range(per * world)), 10000)
    want = H.oracle_multi(json.dumps(full), paths, synth.C2_AGGREGATES)
    comms = [api.Comm(r, world, 400000, max_aggs=len(synth.C2_AGGREGATES) + 1) for r in range(world)]
    handles = [c.handle() for c in comms]
    for c in comms:
        c.connect(handles)
    qs = []
    for rank in range(world):
        sub = dict(full, segmentRequests=full["segmentRequests"][per * rank:per * (rank + 1)])
        q = api.Query(json.dumps(sub), aggregates=synth.C2_AGGREGATES, path="records", exact_sums=True, seq_offset=rank * per * spec.rows)
        for p in paths[per * rank:per * (rank + 1)]:
            q.add_segment_file(p)
        q.plan()
        qs.append(q)
    blob = api.union_dictionaries([q.export_dictionaries() for q in qs])
    for q, c in zip(qs, comms):
        q.import_dictionaries(blob)
        q.set_comm(c)
        q.prepare()
    for q in qs:
        q.execute()
    for q in qs:
        q.sync()
    got_rows = {}
    for q in qs:
        res = q.finalize()
        g = H.canon_from_gpu(res)
        res.close()
        assert not (set(g["rows"]) & set(got_rows))
        assert g["ts_order"] == sorted(g["ts_order"])
        got_rows.update(g["rows"])
    assert set(got_rows) == set(want["rows"])
    for k, w in want["rows"].items():
        assert all(_bits(a) == _bits(b) for a, b in zip(got_rows[k], w)), (k, got_rows[k], w)
    for q in qs:
        q.close()
    for c in comms:
        c.close()

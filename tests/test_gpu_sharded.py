"""Sharded evaluation on ONE GPU: two rank-shards as two queries, exchanged the way bench.py exchanges
them over NCCL (dense: reduce of the planes; hash: partitioned all-to-all of the occupied cells), against the oracle
over all segments."""
import json

import numpy as np
import pytest

import helpers as H
from lakeside_b200 import synth

pytestmark = pytest.mark.gpu


def _shards(name, spec_a, spec_b, n_each, be, aggs, path):
    from lakeside_b200 import api

    api.init()
    _, pa = H.dataset(name + "_a", spec_a, n_each)
    _, pb = H.dataset(name + "_b", spec_b, n_each, first_index=100)
    paths = pa + pb
    full = synth.push_down_request(be, list(range(n_each)) + list(range(100, 100 + n_each)), 10000)
    qs = []
    for rank in range(2):
        sub, idx = api.shard_request(full, rank, 2)
        q = api.Query(json.dumps(sub), aggregates=aggs, path=path)
        for i in idx:
            q.add_segment_file(paths[i])
        q.plan()
        qs.append(q)
    blob = api.union_dictionaries([q.export_dictionaries() for q in qs])
    for q in qs:
        q.import_dictionaries(blob)
        q.prepare()
    return qs, json.dumps(full), paths


def test_dense_partials_reduce_like_nccl():
    import torch

    be = synth.c2_base_expr()
    # the two shards see different tag dictionaries (pods 0..5 vs 0..9): only the imported union makes cells line up
    sa = synth.SynthSpec(dataset="metrics", rows=60000, n_names=3, cards=(16, 4, 6, 2))
    sb = synth.SynthSpec(dataset="metrics", rows=60000, n_names=5, cards=(16, 4, 10, 3))
    qs, rq, paths = _shards("shard_dense", sa, sb, 2, be, synth.C2_AGGREGATES, "dense")
    assert qs[0].info["n_groups"] == qs[1].info["n_groups"]
    for q in qs:
        q.execute()
    for q in qs:
        q.sync()
    parts = [q.partial_dense() for q in qs]
    n_cells = parts[0][0]
    assert parts[1][0] == n_cells

    def tensor(ptr, dtype):
        class A:
            pass
        a = A()
        a.__cuda_array_interface__ = {"shape": (n_cells,), "typestr": "<f8" if dtype == torch.float64 else "<i8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(a, device="cuda")

    for (pa, op), (pb, _) in zip(parts[0][1], parts[1][1]):
        if op == 0:
            tensor(pa, torch.float64).add_(tensor(pb, torch.float64))
        elif op == 1:
            tensor(pa, torch.int64).add_(tensor(pb, torch.int64))
        else:  # unsigned max on order-preserving keys: flip the sign bit, signed max, flip back
            a, b = tensor(pa, torch.int64), tensor(pb, torch.int64)
            a.bitwise_xor_(torch.iinfo(torch.int64).min)
            b.bitwise_xor_(torch.iinfo(torch.int64).min)
            torch.maximum(a, b, out=a)
            a.bitwise_xor_(torch.iinfo(torch.int64).min)
    torch.cuda.synchronize()
    res = qs[0].finalize()
    got = H.canon_from_gpu(res)
    res.close()
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "sharded/dense")
    for q in qs:
        q.close()


def test_dense_reduce_phase_from_the_other_rank():
    """sharded dense path with an unaligned startTs.  Rank 0 -- the rank that finalizes the reduced planes -- keeps
    no row itself (its segment has no svc-03), so the phase must come from rank 1 (lk_query_phase -> host MIN / MAX ->
    lk_query_set_phase); without it the rows would be stamped with phase 0.  Ranks that saw different phases are an error."""
    import torch
    from lakeside_b200 import api

    api.init()
    be = synth.c2_base_expr()
    sa = synth.SynthSpec(dataset="metrics", rows=40000, n_names=3, cards=(1, 4, 6, 2))   # one service value: no svc-03
    sb = synth.SynthSpec(dataset="metrics", rows=40000, n_names=3, cards=(16, 4, 6, 2))
    pa = H.dataset("dphase_a", sa, 1)[1]
    pb = H.dataset("dphase_b", sb, 1, first_index=100)[1]
    paths = pa + pb
    start = synth.T0 - 3000
    full = synth.push_down_request(be, [0, 100], 10000, start_ts=start, end_ts=synth.T0 + synth.HOUR_MS)
    qs = []
    for rank in range(2):
        q = api.Query(json.dumps(dict(full, segmentRequests=[full["segmentRequests"][rank]])), aggregates=synth.C2_AGGREGATES, path="dense")
        q.add_segment_file(paths[rank])
        q.plan()
        qs.append(q)
    blob = api.union_dictionaries([q.export_dictionaries() for q in qs])
    for q in qs:
        q.import_dictionaries(blob)
        q.prepare()
        q.execute()
    phases = [q.phase() for q in qs]
    assert phases[0][0] == 0xFFFFFFFF and phases[1] == (3000, 3000)
    assert qs[0].survivors == 0
    parts = [q.partial_dense() for q in qs]
    n_cells = parts[0][0]

    def tensor(ptr, f64):
        class A:
            pass
        a = A()
        a.__cuda_array_interface__ = {"shape": (n_cells,), "typestr": "<f8" if f64 else "<i8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(a, device="cuda")

    for (p0, op), (p1, _) in zip(parts[0][1], parts[1][1]):
        if op == 0:
            tensor(p0, True).add_(tensor(p1, True))
        elif op == 1:
            tensor(p0, False).add_(tensor(p1, False))
        else:
            a, b = tensor(p0, False), tensor(p1, False)
            a.bitwise_xor_(torch.iinfo(torch.int64).min)
            b.bitwise_xor_(torch.iinfo(torch.int64).min)
            torch.maximum(a, b, out=a)
            a.bitwise_xor_(torch.iinfo(torch.int64).min)
    torch.cuda.synchronize()
    qs[0].set_phase(min(p[0] for p in phases), max(p[1] for p in phases))
    res = qs[0].finalize()
    got = H.canon_from_gpu(res)
    res.close()
    want = H.oracle_multi(json.dumps(full), paths, synth.C2_AGGREGATES)
    assert len(want["rows"]) > 100
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "sharded/dense/phase")
    assert all((ts - start) % 10000 == 3000 for ts in got["ts_order"])
    # ranks that disagree on the phase: distinct raw timestamps would fold into one row
    qs[1].set_phase(1000, 3000)
    with pytest.raises(api.LakesideUnsupported):
        qs[1].finalize()
    for q in qs:
        q.close()


def test_hash_partials_partitioned_exchange():
    be = synth.c2_base_expr()
    sa = synth.SynthSpec(dataset="metrics", rows=80000, cards=(16, 20, 32, 16))
    sb = synth.SynthSpec(dataset="metrics", rows=80000, cards=(16, 32, 24, 16))
    qs, rq, paths = _shards("shard_sparse", sa, sb, 2, be, synth.C2_AGGREGATES, "hash")
    for q in qs:
        q.execute()
    parts = [q.partial_sparse(2) for q in qs]
    stride = parts[0][2]
    for owner, q in enumerate(qs):
        for ptr, counts, _ in parts:
            off = sum(counts[:owner])
            q.merge_sparse(ptr + off * stride, counts[owner])
    got_rows, order = {}, []
    for q in qs:
        res = q.finalize()
        g = H.canon_from_gpu(res)
        res.close()
        assert not (set(g["rows"]) & set(got_rows)), "partitions overlap"
        got_rows.update(g["rows"])
        assert g["ts_order"] == sorted(g["ts_order"])
        order += g["ts_order"]
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    got = {"cols": want["cols"], "rows": got_rows, "ts_order": sorted(order)}
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "sharded/hash")
    assert all(sum(c) > 0 for _, c, _ in parts) and all(parts[0][1][p] + parts[1][1][p] > 0 for p in range(2))
    for q in qs:
        q.close()


def _comm_ranks(world, pool_records, max_aggs=4):
    """`world` ranks of one communicator inside this process, all on cuda:0 (the ranks' blocks are addressed directly: same
    protocol, same kernels as the one-process-per-GPU layout, where the blocks are mapped with CUDA IPC)."""
    from lakeside_b200 import api

    comms = [api.Comm(r, world, pool_records, max_aggs) for r in range(world)]
    handles = [c.handle() for c in comms]
    for c in comms:
        c.connect(handles)
    return comms


def _run_sharded_records(qs, comms, steps=1):
    """All ranks execute (scan + exchange: records land in the owners' pools), THEN all finalize -- on one GPU a rank's wait
    kernel must not be launched before its sources' scans (kernels that wait on one another cannot share a GPU)."""
    out = None
    for _ in range(steps):
        for q in qs:
            q.execute()
        for q in qs:
            q.sync()
        got_rows, order, parts = {}, [], []
        for q in qs:
            res = q.finalize()
            g = H.canon_from_gpu(res)
            res.close()
            assert not (set(g["rows"]) & set(got_rows)), "partitions overlap"
            got_rows.update(g["rows"])
            assert g["ts_order"] == sorted(g["ts_order"])
            order += g["ts_order"]
            parts.append(len(g["rows"]))
        out = (got_rows, order, parts)
    return out


@pytest.mark.parametrize("world", [1, 2, 3])
def test_records_exchange_through_comm(world):
    """The record path's exchange inside the scan: `world` rank-shards with different tag dictionaries, cells hash-partitioned
    over the ranks' receive pools; the union of the ranks' rows equals the oracle over all segments.  Two steps: the pools
    alternate between epochs."""
    from lakeside_b200 import api

    api.init()
    be = synth.c2_base_expr()
    specs = [synth.SynthSpec(dataset="metrics", rows=80000, cards=(16, 20 + 4 * r, 32 - 3 * r, 16)) for r in range(world)]
    paths = []
    for r, sp in enumerate(specs):
        paths += H.dataset(f"shard_comm_{r}", sp, 2, first_index=100 * r)[1]
    idx = [100 * r + i for r in range(world) for i in range(2)]
    full = synth.push_down_request(be, idx, 10000)
    comms = _comm_ranks(world, pool_records=200000)
    qs = []
    for rank in range(world):
        sub = dict(full, segmentRequests=full["segmentRequests"][2 * rank:2 * rank + 2])
        q = api.Query(json.dumps(sub), aggregates=synth.C2_AGGREGATES, path="records")
        for p in paths[2 * rank:2 * rank + 2]:
            q.add_segment_file(p)
        q.plan()
        qs.append(q)
    blob = api.union_dictionaries([q.export_dictionaries() for q in qs])
    for q, c in zip(qs, comms):
        q.import_dictionaries(blob)
        q.set_comm(c)
        q.prepare()
    got_rows, order, parts = _run_sharded_records(qs, comms, steps=3)
    want = H.oracle_multi(json.dumps(full), paths, synth.C2_AGGREGATES)
    got = {"cols": want["cols"], "rows": got_rows, "ts_order": sorted(order)}
    H.assert_same(got, want, ["sum", "sum", "min", "max"], f"sharded/records/world{world}")
    assert all(n > 0 for n in parts), parts  # every rank owns part of the cells
    if world > 1:
        assert max(parts) < 0.75 * sum(parts), parts
    for q in qs:
        q.close()
    for c in comms:
        c.close()


def test_records_exchange_phase_and_empty_source():
    """an unaligned startTs gives the metric timestamps a phase; a rank whose filter keeps nothing must still
    stamp the rows it OWNS with the phase its peers saw (phase min / max travel with the exchange), and ranks that saw
    different phases are detected."""
    from lakeside_b200 import api

    api.init()
    be = synth.c2_base_expr()
    sa = synth.SynthSpec(dataset="metrics", rows=50000, cards=(16, 8, 8, 4))
    # rank 1's segments have no svc-03 at all (one service value): its scan finds no survivor
    sb = synth.SynthSpec(dataset="metrics", rows=50000, cards=(1, 8, 8, 4))
    pa = H.dataset("phase_a", sa, 1)[1]
    pb = H.dataset("phase_b", sb, 1, first_index=100)[1]
    paths = pa + pb
    start = synth.T0 - 3000  # timestamps sit at phase 3000 of the 10 s grid counted from startTs
    full = synth.push_down_request(be, [0, 100], 10000, start_ts=start, end_ts=synth.T0 + synth.HOUR_MS)
    comms = _comm_ranks(2, pool_records=100000)
    qs = []
    for rank in range(2):
        sub = dict(full, segmentRequests=[full["segmentRequests"][rank]])
        q = api.Query(json.dumps(sub), aggregates=synth.C2_AGGREGATES, path="records")
        q.add_segment_file(paths[rank])
        q.plan()
        qs.append(q)
    blob = api.union_dictionaries([q.export_dictionaries() for q in qs])
    for q, c in zip(qs, comms):
        q.import_dictionaries(blob)
        q.set_comm(c)
        q.prepare()
    got_rows, order, parts = _run_sharded_records(qs, comms)
    assert qs[1].survivors == 0 and parts[1] > 0  # rank 1 scanned nothing useful but owns half of the cells
    want = H.oracle_multi(json.dumps(full), paths, synth.C2_AGGREGATES)
    got = {"cols": want["cols"], "rows": got_rows, "ts_order": sorted(order)}
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "sharded/records/phase")
    assert all((ts - start) % 10000 == 3000 for ts in order)
    for q in qs:
        q.close()
    for c in comms:
        c.close()

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


@pytest.mark.parametrize("table", ["hash", "records"])
def test_sparse_partials_partitioned_exchange(table):
    be = synth.c2_base_expr()
    sa = synth.SynthSpec(dataset="metrics", rows=80000, cards=(16, 20, 32, 16))
    sb = synth.SynthSpec(dataset="metrics", rows=80000, cards=(16, 32, 24, 16))
    qs, rq, paths = _shards("shard_sparse", sa, sb, 2, be, synth.C2_AGGREGATES, table)
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
    assert all(len(q.export_dictionaries()) > 0 for q in qs)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    got = {"cols": want["cols"], "rows": got_rows, "ts_order": sorted(order)}
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "sharded/sparse")
    # both partitions are populated
    assert all(sum(c) > 0 for _, c, _ in parts) and all(parts[0][1][p] + parts[1][1][p] > 0 for p in range(2))
    for q in qs:
        q.close()

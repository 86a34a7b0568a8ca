"""GPU parity tests (run on the B200 box: pytest -m gpu): the CUDA path through the C ABI against the oracle.

Bit-exact: row selection, group keys, counts, min, max.  Double sums: relative 1e-12 (helpers.SUM_RTOL)."""
import json

import numpy as np
import pytest

import helpers as H
from lakeside_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    from lakeside_b200 import api

    api.init()


def test_c1_single_segment_name_filter_sum_60s():
    # BASELINE.json configs[0]: one sealed segment, name filter, sum by 60 s step
    spec = synth.SynthSpec(dataset="logs", rows=1 << 20)
    _, paths = H.dataset("c1_logs_1m", spec, 1)
    rq = H.request_json(synth.c1_base_expr(), [0], 60000)
    got = H.gpu_eval_single(rq, paths)
    want = H.oracle_single(rq, paths)
    assert 0 < len(want["rows"]) <= 60
    H.assert_same(got, want, ["sum"], "c1")


def test_c1_int_values_bit_exact_sum():
    spec = synth.SynthSpec(dataset="logs", rows=300000, int_values=True)
    _, paths = H.dataset("c1_logs_int", spec, 1)
    rq = H.request_json(synth.c1_base_expr(), [0], 60000)
    got = H.gpu_eval_single(rq, paths)
    want = H.oracle_single(rq, paths)
    for k, v in want["rows"].items():
        assert got["rows"][k][0] == v[0]  # integer-valued doubles: exact in any order


@pytest.mark.parametrize("path", ["hash", "records"])
def test_c2_four_aggregates(path):
    spec = synth.SynthSpec(dataset="metrics", rows=200000)
    _, paths = H.dataset("c2_m200k", spec, 3)
    rq = H.request_json(synth.c2_base_expr(), [0, 1, 2], 10000)
    got = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES, path=path)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    assert got["info"]["path"] == path
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "c2/" + path)
    assert got["survivors"] >= len(want["rows"])


@pytest.mark.parametrize("slices", [1, 3])
def test_c2_records_partitioned_by_bucket_first(monkeypatch, slices):
    # the record finalize with its partition pass (what sharded queries and long record lists take by
    # default); slices > 1: several partitions per time bucket (by a hash of the group id), as for C4's 139 k records per bucket
    monkeypatch.setenv("LK_REC_SCATTER", "1")
    monkeypatch.setenv("LK_REC_SLICES", str(slices))
    spec = synth.SynthSpec(dataset="metrics", rows=200000)
    _, paths = H.dataset("c2_m200k", spec, 3)
    rq = H.request_json(synth.c2_base_expr(), [0, 1, 2], 10000)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    for _ in range(2):  # the second run re-uses the scratch (cleared on the side stream)
        got = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES, path="records")
        H.assert_same(got, want, ["sum", "sum", "min", "max"], "c2/records+scatter")
    # three aggregates: the generic (not 32-byte) record row
    got = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES[:3], path="records")
    H.assert_same(got, H.oracle_multi(rq, paths, synth.C2_AGGREGATES[:3]), ["sum", "sum", "min"], "c2/records+scatter/3 aggs")


def test_c2_snappy_segments():
    # SNAPPY-compressed metric segments (1 MiB pages: long literals for the PLAIN doubles, dense copy elements for the tag
    # indices, a compressed numeric dictionary for the timestamps), inflated on the device; same rows, same answers
    spec = synth.SynthSpec(dataset="metrics", rows=200000, compression="SNAPPY")
    _, paths = H.dataset("c2_m200k_snappy", spec, 2)
    rq = H.request_json(synth.c2_base_expr(), [0, 1], 10000)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    for tag in ("cold", "warm"):  # warm: inflated pages come from the segment cache
        got = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES)
        H.assert_same(got, want, ["sum", "sum", "min", "max"], "c2/snappy/" + tag)
    plain = H.dataset("c2_m200k", synth.SynthSpec(dataset="metrics", rows=200000), 2)[1]
    assert H.oracle_multi(H.request_json(synth.c2_base_expr(), [0, 1], 10000), plain, synth.C2_AGGREGATES)["rows"] == want["rows"]


@pytest.mark.parametrize("path", ["dense", "hash", "records"])
def test_small_group_space_dense_and_hash(path):
    spec = synth.SynthSpec(dataset="metrics", rows=150000, n_names=4, cards=(16, 4, 4, 2))
    _, paths = H.dataset("small_groups", spec, 2)
    rq = H.request_json(synth.c2_base_expr(), [0, 1], 10000)
    got = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES, path=path)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    assert got["info"]["path"] == path
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "small/" + path)


@pytest.mark.parametrize("agg,rollup", [("sum", "sum"), ("count", "sum"), ("min", "min"), ("max", "max")])
def test_single_aggregate_requests(agg, rollup):
    spec = synth.SynthSpec(dataset="metrics", rows=100000, n_names=8, cards=(16, 8, 8, 4))
    _, paths = H.dataset("single_agg", spec, 2)
    rq = H.request_json(synth.c2_base_expr(agg, rollup), [0, 1], 10000)
    got = H.gpu_eval_single(rq, paths)
    want = H.oracle_single(rq, paths)
    H.assert_same(got, want, [agg], f"{agg}({rollup})")


@pytest.mark.parametrize("null_frac", [2e-5, 1e-3, 0.5, 0.97])
def test_definition_level_shapes(null_frac):
    """def_expand_kernel over very different run structures: a handful of NULLs per chunk (RLE runs of 10^5 rows: written
    to global memory directly, whole words queued for the CTA), one NULL per ~1000 rows (runs longer than a thread's fill
    budget but inside the window), every other row NULL (bit-packed runs of hundreds of rows), almost only NULLs."""
    spec = synth.SynthSpec(dataset="metrics", rows=300000, null_frac=null_frac, cards=(16, 8, 8, 4), n_names=4)
    _, paths = H.dataset(f"defshape_{null_frac}", spec, 2)
    rq = H.request_json(synth.c2_base_expr(), [0, 1], 10000)
    got = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    H.assert_same(got, want, ["sum", "sum", "min", "max"], f"defshape/{null_frac}")

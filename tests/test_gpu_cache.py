"""HBM-resident segment cache (SURVEY §8b "Ownership"; analogue of WorkerApi.scala:53-64): cold and warm evaluations of the
same segment files give the oracle's rows, a replaced file is noticed, eviction and capacity 0 work, threads may share it."""
import os
import shutil
import threading

import pytest

import cases as C
import helpers as H
from lakeside_b200 import synth

pytestmark = pytest.mark.gpu
OPS = ["sum", "sum", "min", "max"]


@pytest.fixture(autouse=True)
def _fresh_cache():
    from lakeside_b200 import api

    api.init()
    cap = api.cache_stats()["capacity_bytes"]
    api.cache_clear()
    yield
    api.cache_configure(cap)
    api.cache_clear()


def _c2(paths, idx):
    rq = H.request_json(synth.c2_base_expr(), idx, 10000)
    return rq, H.oracle_multi(rq, paths, synth.C2_AGGREGATES)


def test_cold_then_warm_same_rows_no_new_misses():
    from lakeside_b200 import api

    _, paths = H.dataset("c2_m200k", synth.SynthSpec(dataset="metrics", rows=200000), 3)
    rq, want = _c2(paths, [0, 1, 2])
    s0 = api.cache_stats()
    assert s0["capacity_bytes"] > 0 and s0["resident_bytes"] == 0
    cold = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES)
    s1 = api.cache_stats()
    assert s1["segments"] == 3 and s1["column_hits"] == s0["column_hits"] and s1["resident_bytes"] > 0
    warm = H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES)
    s2 = api.cache_stats()
    assert s2["column_misses"] == s1["column_misses"], "the warm query must find every column chunk in HBM"
    assert s2["column_hits"] - s1["column_hits"] == 3 * 10  # ts, name, filter tag, 3 group tags, 4 value columns per segment
    assert s2["resident_bytes"] == s1["resident_bytes"]
    H.assert_same(cold, want, OPS, "cache/cold")
    H.assert_same(warm, want, OPS, "cache/warm")
    # the hash-table layout over the same cached chunks
    H.assert_same(H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES, path="hash"), want, OPS, "cache/warm-hash")


def test_other_query_reuses_shared_columns_and_adds_its_own():
    from lakeside_b200 import api

    _, paths = H.dataset("c2_m200k", synth.SynthSpec(dataset="metrics", rows=200000), 3)
    rq, want = _c2(paths, [0, 1, 2])
    H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES)
    s1 = api.cache_stats()
    # a single-aggregate query through lk_eval with another filter: shares ts / name / tags / rollup_sum with the first one
    be = synth.c2_base_expr()
    spec = synth.SynthSpec(dataset="metrics", rows=200000)
    be["filter"] = {"k": synth.GROUP_TAGS[0], "v": synth.tag_values(spec.prefixes[1], spec.cards[1])[:2], "op": "in", "dataType": "string", "extracted": False, "computed": False}
    rq2 = H.request_json(be, [0, 1, 2], 10000)
    got = H.gpu_eval_single(rq2, paths)
    s2 = api.cache_stats()
    assert s2["column_hits"] > s1["column_hits"] and s2["column_misses"] == s1["column_misses"]
    H.assert_same(got, H.oracle_single(rq2, paths), ["sum"], "cache/second-query")


def test_replaced_file_is_noticed(tmp_path):
    from lakeside_b200 import api

    spec_a = synth.SynthSpec(dataset="metrics", rows=60000)
    root_a = str(tmp_path / "a")
    pa = synth.write_dataset(root_a, spec_a, 1)
    rq, want_a = _c2(pa, [0])
    H.assert_same(H.gpu_eval_multi(rq, pa, synth.C2_AGGREGATES), want_a, OPS, "cache/file-a")
    # same path, other content (another seed -> other rows): size or mtime differ, the cached chunks must not be used
    pb = synth.write_dataset(str(tmp_path / "b"), spec_a, 1, first_index=5)
    shutil.copyfile(pb[0], pa[0])
    os.utime(pa[0], ns=(os.stat(pa[0]).st_atime_ns, os.stat(pa[0]).st_mtime_ns + 1_000_000))
    want_b = H.oracle_multi(rq, pa, synth.C2_AGGREGATES)
    assert want_b["rows"] != want_a["rows"]
    H.assert_same(H.gpu_eval_multi(rq, pa, synth.C2_AGGREGATES), want_b, OPS, "cache/file-b")
    assert api.cache_stats()["segments"] == 1


def test_eviction_and_capacity_zero():
    from lakeside_b200 import api

    _, paths = H.dataset("c2_m200k", synth.SynthSpec(dataset="metrics", rows=200000), 3)
    rq, want = _c2(paths, [0, 1, 2])
    H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES)
    per_seg = api.cache_stats()["resident_bytes"] // 3
    api.cache_clear()
    api.cache_configure(int(per_seg * 1.5))  # room for one segment only
    H.assert_same(H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES), want, OPS, "cache/small")
    s = api.cache_stats()
    assert s["resident_bytes"] <= s["capacity_bytes"] and s["evicted_segments"] > 0 and s["segments"] >= 1
    H.assert_same(H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES), want, OPS, "cache/small-again")
    api.cache_configure(0)
    assert api.cache_stats()["resident_bytes"] == 0
    H.assert_same(H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES), want, OPS, "cache/off")
    assert api.cache_stats()["segments"] == 0


def test_threads_share_the_cache():
    _, paths = H.dataset("c2_m200k", synth.SynthSpec(dataset="metrics", rows=200000), 3)
    rq, want = _c2(paths, [0, 1, 2])
    errs = []

    def work():
        try:
            for _ in range(3):
                H.assert_same(H.gpu_eval_multi(rq, paths, synth.C2_AGGREGATES), want, OPS, "cache/threads")
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


def test_layout_cases_warm():
    # every file-layout case (PLAIN string pages re-encoded behind their chunk, several row groups, DataPage V2, NULL runs ...)
    # evaluated twice: the second time from cached chunks and cached page / dictionary indexes
    n = 0
    for name, paths, rq, ops in C.all_cases() + C.gpu_only_cases():
        if not name.startswith("layout/"):
            continue
        want = H.oracle_single(rq, paths)
        H.assert_same(H.gpu_eval_single(rq, paths), want, ops, name + "/cold")
        H.assert_same(H.gpu_eval_single(rq, paths), want, ops, name + "/warm")
        n += 1
    assert n > 5

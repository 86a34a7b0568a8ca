"""CPU tests: the host planner + decode primitives (run through the sequential emulator of the scan kernel)
against the oracle, over the BASELINE.json configurations at reduced sizes and the edge cases of SURVEY §8c."""
import json

import pytest

import helpers as H
from lakeside_b200 import synth


def test_c2_shape_four_aggregates():
    spec = synth.SynthSpec(dataset="metrics", rows=60000)
    _, paths = H.dataset("m60k", spec, 2)
    rq = H.request_json(synth.c2_base_expr(), [0, 1], 10000)
    got, info = H.emul_eval(rq, paths, aggs=synth.C2_AGGREGATES)
    want = H.oracle_multi(rq, paths, synth.C2_AGGREGATES)
    H.assert_same(got, want, ["sum", "sum", "min", "max"], "c2")
    assert info["path"] == "records"  # group space too large for dense planes: the planner aggregates by sorting records
    # four nullable tag columns (filter + three group-bys) x two row groups get a definition bitmap (one bit per row + slack), expanded on the device
    assert info["def_chunks"] == 8 and info["def_bitmap_bytes"] >= 8 * (60000 // 8)
    assert len(got["rows"]) > 5000

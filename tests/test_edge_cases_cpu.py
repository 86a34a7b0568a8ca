"""CPU: literal row-wise oracle == vectorised oracle == host planner + decode primitives (scan-kernel emulator)
on every edge case of tests/cases.py."""
import json

import pytest

import cases
import helpers as H
import lakeside_oracle as lo

CASES = cases.all_cases()
ERRORS = cases.error_cases()


@pytest.mark.parametrize("cid,paths,rq,ops", CASES, ids=[c[0] for c in CASES])
def test_oracles_and_emulator_agree(cid, paths, rq, ops):
    req = lo.push_down_request_from_json(rq)
    row = H.canon_from_glob_result(lo.evaluate_glob_rowwise(req, paths))
    vec = H.canon_from_glob_result(lo.evaluate_glob(req, paths))
    H.assert_same(vec, row, ops, cid + " oracle(vectorised) vs oracle(row-wise)")
    emu, info = H.emul_eval(rq, paths)
    H.assert_same(emu, vec, ops, cid + " emulator vs oracle")
    assert emu["cols"][0] == vec["cols"][0] and emu["cols"][2:] == vec["cols"][2:]  # JDBC column names toDataPoint sees
    if "small_tiles" not in cid:
        emu2, _ = H.emul_eval(rq, paths, path="hash", tile_rows=64)
        H.assert_same(emu2, vec, ops, cid + " emulator(64-row tiles) vs oracle")


@pytest.mark.parametrize("cid,paths,rq,kind", ERRORS, ids=[c[0] for c in ERRORS])
def test_error_mapping(cid, paths, rq, kind):
    code = {"invalid": 1, "unsupported": 2, "query": 5}[kind]
    with pytest.raises(H.EmulError) as e:
        H.emul_eval(rq, paths)
    assert e.value.code == code, str(e.value)
    # the oracle classifies the same way
    if kind == "unsupported" and cid not in ("err/compressed_pages", "err/backreference"):
        with pytest.raises(lo.OracleUnsupported):
            lo.evaluate_glob(lo.push_down_request_from_json(rq), paths)
    if kind == "query":
        with pytest.raises(lo.OracleQueryError):
            lo.evaluate_glob(lo.push_down_request_from_json(rq), paths)

"""Shared test helpers: dataset cache, oracle/emulator/GPU result canonicalisation and comparison."""
from __future__ import annotations

import ctypes
import json
import math
import os
import subprocess
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "_data")
CSRC = os.path.join(ROOT, "lakeside_b200", "csrc")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")

import lakeside_oracle as lo  # noqa: E402  (tests/conftest.py puts oracle/ on sys.path)
from lakeside_b200 import synth  # noqa: E402

SUM_RTOL = 1e-12  # north-star tolerance for double sums; everything else is bit-exact


def dataset(name: str, spec: synth.SynthSpec, n_segments: int, first_index: int = 0) -> Tuple[str, List[str]]:
    root = os.path.join(DATA, name)
    paths = synth.write_dataset(root, spec, n_segments, first_index=first_index, workers=min(4, n_segments))
    return root, paths


def request_json(base_expr: dict, indices: Sequence[int], step: int, **kw) -> str:
    return json.dumps(synth.push_down_request(base_expr, indices, step, **kw))


# ---------------------------------------------------------------- emulator (test infrastructure)
_emul = None


def emul_lib():
    global _emul
    if _emul is not None:
        return _emul
    out = os.path.join(EMUL_DIR, "_build", "liblk_emul.so")
    srcs = [os.path.join(EMUL_DIR, "lk_emul.cpp")] + [os.path.join(CSRC, f) for f in
                                                     ("lk_regex.cpp", "lk_expr.cpp", "lk_parquet.cpp", "lk_plan.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-I" + CSRC, "-o", out] + srcs + ["-lpthread"])
    lib = ctypes.CDLL(out)
    lib.lk_emul_last_error.restype = ctypes.c_char_p
    lib.lk_emul_eval.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int,
                                 ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p)]
    for f, rt in (("lk_emul_rows", ctypes.c_int64), ("lk_emul_survivors", ctypes.c_int64), ("lk_emul_nvalues", ctypes.c_int),
                  ("lk_emul_ntags", ctypes.c_int), ("lk_emul_ts", ctypes.POINTER(ctypes.c_int64)), ("lk_emul_info", ctypes.c_char_p)):
        getattr(lib, f).restype = rt
        getattr(lib, f).argtypes = [ctypes.c_void_p]
    lib.lk_emul_value.restype = ctypes.POINTER(ctypes.c_double)
    lib.lk_emul_null.restype = ctypes.POINTER(ctypes.c_uint8)
    lib.lk_emul_codes.restype = ctypes.POINTER(ctypes.c_int32)
    lib.lk_emul_dict_size.restype = ctypes.c_int
    lib.lk_emul_dict.restype = ctypes.c_char_p
    lib.lk_emul_col_name.restype = ctypes.c_char_p
    for f in ("lk_emul_value", "lk_emul_null", "lk_emul_codes", "lk_emul_dict_size", "lk_emul_col_name"):
        getattr(lib, f).argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.lk_emul_dict.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.lk_emul_free.argtypes = [ctypes.c_void_p]
    _emul = lib
    return lib


class EmulError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def emul_eval(req_json: str, paths: Sequence[str], aggs: Optional[Sequence[Tuple[str, str]]] = None, path: str = "", tile_rows: int = 0):
    """Runs the CPU emulation of the scan kernel; returns the canonical dict (see canon_*)."""
    lib = emul_lib()
    blobs = [open(p, "rb").read() for p in paths]
    bufs = (ctypes.c_void_p * len(blobs))(*[ctypes.cast(ctypes.c_char_p(b), ctypes.c_void_p) for b in blobs])
    lens = (ctypes.c_size_t * len(blobs))(*[len(b) for b in blobs])
    aggs_json = json.dumps([{"aggregation": a, "rollup": r} for a, r in aggs]).encode() if aggs else b""
    h = ctypes.c_void_p()
    rc = lib.lk_emul_eval(req_json.encode(), aggs_json, path.encode(), tile_rows, len(blobs), bufs, lens, ctypes.byref(h))
    if rc != 0:
        raise EmulError(rc, lib.lk_emul_last_error().decode())
    try:
        n = lib.lk_emul_rows(h)
        na, nk = lib.lk_emul_nvalues(h), lib.lk_emul_ntags(h)
        ts = np.ctypeslib.as_array(lib.lk_emul_ts(h), (n,)).copy() if n else np.zeros(0, np.int64)
        vals = [np.ctypeslib.as_array(lib.lk_emul_value(h, a), (n,)).copy() if n else np.zeros(0) for a in range(na)]
        nulls = [np.ctypeslib.as_array(lib.lk_emul_null(h, a), (n,)).copy() if n else np.zeros(0, np.uint8) for a in range(na)]
        codes = [np.ctypeslib.as_array(lib.lk_emul_codes(h, k), (n,)).copy() if n else np.zeros(0, np.int32) for k in range(nk)]
        dicts = [[lib.lk_emul_dict(h, k, i).decode() for i in range(lib.lk_emul_dict_size(h, k))] for k in range(nk)]
        cols = [lib.lk_emul_col_name(h, i).decode() for i in range(1 + na + nk)]
        info = json.loads(lib.lk_emul_info(h).decode())
        return canon_from_arrays(ts, vals, nulls, codes, dicts, cols), info
    finally:
        lib.lk_emul_free(h)


# ---------------------------------------------------------------- canonical form
def canon_from_arrays(ts, vals, nulls, codes, dicts, cols) -> dict:
    rows: Dict[tuple, tuple] = {}
    n = len(ts)
    tagcols = []
    for k, c in enumerate(codes):
        d = np.array(dicts[k] + [None], dtype=object)
        tagcols.append(d[np.where(c < 0, len(dicts[k]), c)] if n else [])
    for i in range(n):
        key = (int(ts[i]),) + tuple(t[i] for t in tagcols)
        assert key not in rows, f"duplicate output row {key}"
        rows[key] = tuple(None if nulls[a][i] else float(vals[a][i]) for a in range(len(vals)))
    return {"cols": cols, "rows": rows, "ts_order": [int(t) for t in ts]}


def canon_from_oracle_multi(res: dict) -> dict:
    n = len(res["ts"])
    dicts = [res["key_dicts"][c] for c in res["key_cols"]]
    cols = [res["plan"].ts_col_name] + ["value"] * len(res["values"]) + ["name"] + list(res["plan"].group_cols)
    nulls = [np.asarray(m, np.uint8) for m in res["nulls"]]
    return canon_from_arrays(res["ts"], res["values"], nulls, [np.asarray(c) for c in res["key_codes"]], dicts, cols) if n or True else {}


def canon_from_glob_result(res: "lo.GlobResult") -> dict:
    rows = {}
    for r in res.rows:
        rows[(r.ts,) + tuple(r.tags)] = (r.value,)
    return {"cols": list(res.columns), "rows": rows, "ts_order": [r.ts for r in res.rows]}


def assert_same(got: dict, want: dict, ops: Sequence[str], what: str = "", null_sum_is_zero: bool = True):
    """Bit-exact keys / counts / min / max; sums within SUM_RTOL.  SQL-NULL sums read as 0.0 through
    ResultSet.getDouble (Commons.scala:427) and the GPU path does not track them separately."""
    gk, wk = set(got["rows"]), set(want["rows"])
    assert gk == wk, f"{what}: group keys differ: missing {sorted(wk - gk, key=str)[:5]} extra {sorted(gk - wk, key=str)[:5]} ({len(gk)} vs {len(wk)})"
    assert got["ts_order"] == sorted(got["ts_order"]), f"{what}: rows are not sorted by timestamp"
    for key, wv in want["rows"].items():
        gv = got["rows"][key]
        assert len(gv) == len(wv) == len(ops)
        for a, op in enumerate(ops):
            g, w = gv[a], wv[a]
            if op == "sum":
                g = 0.0 if g is None else g
                w = 0.0 if w is None else w
                if math.isnan(w) or math.isinf(w):
                    assert (math.isnan(g) and math.isnan(w)) or g == w, f"{what}: sum {key}: {g} vs {w}"
                else:
                    assert abs(g - w) <= SUM_RTOL * max(abs(w), abs(g)), f"{what}: sum {key}: {g!r} vs {w!r}"
            else:
                if w is None or g is None:
                    assert g is None and w is None, f"{what}: {op} {key}: {g} vs {w}"
                elif math.isnan(w):
                    assert math.isnan(g), f"{what}: {op} {key}: {g} vs NaN"
                else:
                    assert g == w and math.copysign(1, g) == math.copysign(1, w) or (g == w == 0.0), f"{what}: {op} {key}: {g!r} vs {w!r}"


def oracle_multi(req_json: str, paths, aggs: Sequence[Tuple[str, str]]) -> dict:
    req = lo.push_down_request_from_json(req_json)
    pref = "rollup_" if req.baseExpr.dataset == "metrics" else None
    res = lo.evaluate_glob(req, paths, aggs=[(a, (pref + r) if pref else lo.VALUE) for a, r in aggs])
    return canon_from_oracle_multi(res)


def oracle_single(req_json: str, paths) -> dict:
    return canon_from_glob_result(lo.evaluate_glob(lo.push_down_request_from_json(req_json), paths))


# ---------------------------------------------------------------- GPU (through the C ABI)
def canon_from_gpu(res) -> dict:
    return canon_from_arrays(res.ts, res.values, res.value_nulls, res.tag_codes, res.tag_dicts, res.columns)


def gpu_eval_single(req_json: str, paths) -> dict:
    from lakeside_b200 import api

    res = api.eval_glob(req_json, list(paths))
    try:
        return canon_from_gpu(res)
    finally:
        res.close()


def gpu_eval_multi(req_json: str, paths, aggs, path: str = "auto", buffers: bool = False) -> dict:
    from lakeside_b200 import api

    with api.Query(req_json, aggregates=aggs, path=path) as q:
        for p in paths:
            if buffers:
                q.add_segment_bytes(open(p, "rb").read())
            else:
                q.add_segment_file(p)
        q.prepare()
        q.execute()
        res = q.finalize()
        try:
            out = canon_from_gpu(res)
            out["info"] = q.info
            out["survivors"] = q.survivors
            return out
        finally:
            res.close()

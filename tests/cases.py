"""Edge-case segments and DataExprs shared by the CPU (oracle x oracle x emulator) and GPU parity tests (SURVEY §8c:
every operator, NULL tags, missing columns, ts on bucket edges S / E-1 / E, multi-file globs with different schemas,
NaN/inf/-0.0, required columns, PLAIN vs dictionary pages, several row groups / pages, DataPage V2, int/float values)."""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, List, Tuple

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

import helpers as H
from lakeside_b200 import synth

T0 = synth.T0
TS, NAME, VALUE = synth.TIMESTAMP, synth.NAME, synth.VALUE


def _write(name: str, idx: int, table: pa.Table, dataset: str, **kw) -> str:
    path = synth.segment_path(os.path.join(H.DATA, "edge_" + name), dataset, synth.segment_id_for(idx))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    opts = dict(compression="NONE", use_dictionary=True, data_page_version="1.0", write_statistics=True)
    opts.update(kw)
    pq.write_table(table, path, **opts)
    return path


def _req(be: dict, n: int, step: int, start=T0, end=T0 + 3600000) -> str:
    return json.dumps(synth.push_down_request(be, list(range(n)), step, start_ts=start, end_ts=end))


def _logs_be(filt, agg="sum", group_bys=(), **chart):
    return {"id": "e", "dataset": "logs", "filter": filt, "chart": dict({"aggregation": agg, "groupBys": list(group_bys), "type": "count"}, **chart)}


def _metrics_be(filt, agg="sum", rollup="sum", group_bys=()):
    return {"id": "e", "dataset": "metrics", "filter": filt, "chart": {"aggregation": agg, "rollup": rollup, "groupBys": list(group_bys), "type": "count"}}


def F(k, op, *v, **kw):
    return dict({"k": k, "v": list(v), "op": op, "dataType": "string", "extracted": False, "computed": False}, **kw)


def _small_logs(n=4000, seed=1, null_value=False, nan=False, tags=True):
    rng = np.random.default_rng(seed)
    ts = np.sort(T0 + rng.integers(-2000, 3600000 + 2000, n)).astype(np.int64)
    cols = {TS: pa.array(ts), NAME: pa.array(rng.choice(["alpha", "beta", "Gamma", ""], n), mask=rng.random(n) < 0.03)}
    if tags:
        cols["resource.service.name"] = pa.array(rng.choice(["svc-a", "svc-b", "SVC-C", "null", "x.y"], n), mask=rng.random(n) < 0.2)
        cols["level"] = pa.array(rng.choice(["info", "warn", "error"], n), mask=rng.random(n) < 0.1)
    v = rng.normal(0, 100, n)
    if nan:
        sp = rng.random(n)
        v[sp < 0.05] = np.nan
        v[(sp >= 0.05) & (sp < 0.08)] = np.inf
        v[(sp >= 0.08) & (sp < 0.11)] = -np.inf
        v[(sp >= 0.11) & (sp < 0.2)] = -0.0
        v[(sp >= 0.2) & (sp < 0.25)] = 0.0
    cols[VALUE] = pa.array(v, mask=(rng.random(n) < 0.15) if null_value else None)
    return pa.table(cols)


def case_operators() -> List[Tuple[str, List[str], str, List[str]]]:
    """One small logs segment, one request per operator / logic shape.  Returns (id, paths, request, ops)."""
    p = [_write("ops", 0, _small_logs(), "logs")]
    svc = "resource.service.name"
    filters = {
        "eq": F(svc, "eq", "svc-a"),
        "ne": F(svc, "!=", "svc-a"),
        "in": F(svc, "in", "svc-a", "SVC-C", "nope"),
        "not_in": F(svc, "not_in", "svc-a", "x.y"),
        "regex_ci": F(svc, "regex", "^svc-[ab]$"),
        "regex_partial": F(svc, "regex", "c"),
        "contains": F(svc, "contains", "VC-"),
        "has": F(svc, "has", ""),
        "exists": F("level", "exists"),
        "eq_literal_null_string": F(svc, "eq", "null"),
        "not_eq": {"not": F(svc, "eq", "svc-a")},
        "not_exists": {"not": F(svc, "exists")},
        "and": {"q1": F(svc, "eq", "svc-b"), "q2": F("level", "!=", "info"), "op": "and"},
        "or_with_null": {"q1": F(svc, "eq", "svc-b"), "q2": F("level", "eq", "error"), "op": "or"},
        "nary_or": {"q1": F(svc, "eq", "svc-a"), "q2": F(svc, "eq", "svc-b"), "q3": F("level", "eq", "warn"), "op": "or"},
        "not_and_or": {"not": {"q1": {"q1": F(svc, "regex", "svc"), "q2": F("level", "in", "info", "warn"), "op": "and"},
                               "q2": F(NAME, "eq", "beta"), "op": "or"}},
        "missing_col_is_false_or": {"q1": F("no.such.column", "eq", "x"), "q2": F(svc, "eq", "svc-a"), "op": "or"},
        "missing_col_is_false_and": {"q1": F("no.such.column", "exists"), "q2": F(svc, "eq", "svc-a"), "op": "and"},
        "not_missing_col_in_fieldset": {"q1": {"not": F("no.such.column", "eq", "x")}, "q2": F("no.such.column", "exists"), "op": "or"},
        "value_gt": F(VALUE, "gt", "25.5", dataType="number"),
        "value_range": {"q1": F(VALUE, "ge", "-50", dataType="number"), "q2": F(VALUE, "lt", "50", dataType="number"), "op": "and"},
        "value_le_or_tag": {"q1": F(VALUE, "le", "-100", dataType="number"), "q2": F(svc, "eq", "x.y"), "op": "or"},
        "name_empty_string": F(NAME, "eq", ""),
    }
    out = []
    for fid, flt in filters.items():
        out.append((f"ops/{fid}", p, _req(_logs_be(flt, "sum", ["level"]), 1, 60000), ["sum"]))
    for agg in ("count", "min", "max"):
        out.append((f"ops/agg_{agg}", p, _req(_logs_be(filters["regex_ci"], agg, [svc, "missing.group"]), 1, 300000), [agg]))
    return out


def case_time_edges():
    # rows exactly at S-1, S, bucket edges, E-1, E; S not aligned to the step (events bucket by absolute ts - ts % step)
    S, E, step = T0 + 12345, T0 + 12345 + 10 * 7000, 7000
    ts = np.array(sorted([S - 1, S, S + 1, E - 1, E, E + 1] + [S - S % step + k * step + d for k in range(0, 12) for d in (-1, 0, 1)]), np.int64)
    n = len(ts)
    t = pa.table({TS: pa.array(ts), NAME: pa.array(["m"] * n), "k": pa.array(["a", "b"] * (n // 2) + ["a"] * (n % 2)), VALUE: pa.array(np.arange(n, dtype=np.float64))})
    p = [_write("edges", 0, t, "logs")]
    out = [("edges/events", p, _req(_logs_be(F(NAME, "eq", "m"), "sum", ["k"]), 1, step, S, E), ["sum"]),
           ("edges/events_count", p, _req(_logs_be(F(NAME, "eq", "m"), "count", []), 1, step, S, E), ["count"]),
           ("edges/empty_range", p, _req(_logs_be(F(NAME, "eq", "m"), "sum", []), 1, step, S, S), ["sum"])]
    # metrics: GROUP BY raw timestamp, all timestamps on one grid offset from startTs
    g = np.sort(np.concatenate([S + 500 + step * np.arange(10), S + 500 + step * np.arange(0, 10, 3)])).astype(np.int64)
    tm = pa.table({TS: pa.array(g), NAME: pa.array(["m"] * len(g)), "rollup_sum": pa.array(np.arange(len(g), dtype=np.float64)),
                   "rollup_max": pa.array(np.arange(len(g), dtype=np.float64) * 2)})
    pm = [_write("edges_m", 0, tm, "metrics")]
    out.append(("edges/metrics_phase", pm, _req(_metrics_be(F(NAME, "eq", "m"), "sum", "sum"), 1, step, S, E), ["sum"]))
    out.append(("edges/metrics_max", pm, _req(_metrics_be(F(NAME, "eq", "m"), "max", "max"), 1, step, S, E), ["max"]))
    return out


def case_nulls_nan():
    p = [_write("nan", 0, _small_logs(3000, 5, null_value=True, nan=True), "logs")]
    out = []
    for agg in ("sum", "count", "min", "max"):
        out.append((f"nan/{agg}", p, _req(_logs_be(F("level", "exists"), agg, ["level"]), 1, 600000), [agg]))
    return out


def case_schema_drift():
    a = _small_logs(1500, 11)
    b = _small_logs(1500, 12).drop(["level"])                      # group-by / filter column absent in one file
    c = _small_logs(1500, 13).drop(["resource.service.name"])      # filter column absent in another
    paths = [_write("drift", 0, a, "logs"), _write("drift", 1, b, "logs"), _write("drift", 2, c, "logs")]
    svc = "resource.service.name"
    return [("drift/group_col_missing_in_one", paths, _req(_logs_be(F(svc, "!=", "svc-a"), "sum", ["level"]), 3, 600000), ["sum"]),
            ("drift/filter_exists", paths, _req(_logs_be({"not": F(svc, "exists")}, "count", ["level", svc]), 3, 600000), ["count"]),
            ("drift/or", paths, _req(_logs_be({"q1": F(svc, "eq", "svc-b"), "q2": F("level", "eq", "warn"), "op": "or"}, "max", []), 3, 600000), ["max"])]


def case_physical_layouts():
    """Same logical data written in different physical shapes must give identical answers."""
    base = _small_logs(20000, 21, null_value=True)
    be = _logs_be({"q1": F("resource.service.name", "regex", "svc-[ab]"), "q2": F("level", "!=", "warn"), "op": "and"}, "sum", ["level"])
    rq = _req(be, 1, 60000)
    variants = {
        "default": {},
        "small_pages": {"data_page_size": 512, "write_batch_size": 64},
        "row_groups": {"row_group_size": 3000},
        "no_dictionary": {"use_dictionary": ["resource.service.name", "level", NAME]},   # ts/value PLAIN, tags dictionary
        "v2_pages": {"data_page_version": "2.0"},
        "row_groups_small_pages": {"row_group_size": 4500, "data_page_size": 300, "write_batch_size": 32},
    }
    out = []
    for i, (vid, kw) in enumerate(variants.items()):
        out.append((f"layout/{vid}", [_write("layout_" + vid, 0, base, "logs", **kw)], rq, ["sum"]))
    # required (non-nullable) columns: no definition levels at all
    n = 5000
    rng = np.random.default_rng(3)
    schema = pa.schema([pa.field(TS, pa.int64(), nullable=False), pa.field(NAME, pa.string(), nullable=False),
                        pa.field("level", pa.string(), nullable=False), pa.field(VALUE, pa.float64(), nullable=False)])
    t = pa.table({TS: np.sort(T0 + rng.integers(0, 3600000, n)), NAME: rng.choice(["a", "b"], n), "level": rng.choice(["i", "w", "e"], n),
                  VALUE: rng.random(n)}, schema=schema)
    out.append(("layout/required_columns", [_write("required", 0, t, "logs")], _req(_logs_be(F("level", "in", "i", "e"), "min", [NAME]), 1, 60000), ["min"]))
    # int64 / float32 / int32 value columns, wide dictionary (12-bit codes), chart field variant
    wide = np.array([f"pod-{i:04d}" for i in range(3000)], dtype=object)
    t2 = pa.table({TS: np.sort(T0 + rng.integers(0, 3600000, n)).astype(np.int64), NAME: rng.choice(["a", "b"], n),
                   "pod": pa.array(wide[rng.integers(0, 3000, n)], mask=rng.random(n) < 0.05),
                   VALUE: pa.array(rng.integers(-1000, 1000, n).astype(np.int64)),
                   "latency$number": pa.array(rng.random(n).astype(np.float32), mask=rng.random(n) < 0.3),
                   "bytes$datasize": pa.array(rng.integers(0, 1 << 20, n).astype(np.int32))})
    p2 = [_write("types", 0, t2, "logs")]
    out.append(("layout/int64_value_wide_dict", p2, _req(_logs_be(F("pod", "regex", "^pod-1"), "sum", ["pod"]), 1, 600000), ["sum"]))
    out.append(("layout/field_float32_notnull", p2, _req(_logs_be(F(NAME, "eq", "a"), "max", [], fieldName="latency", fieldType="number"), 1, 600000), ["max"]))
    # PLAIN (non-dictionary) string pages -- what a writer falls back to when a chunk's dictionary outgrows its page limit
    # (high-cardinality tags such as pod names): (a) the whole column written without a dictionary, (b) a dictionary that
    # overflows in the middle of the chunk (dictionary pages first, PLAIN pages after), as filter and as group-by column
    hc = np.array([f"pod-{i:05d}-{'x' * (i % 7)}" for i in range(2500)], dtype=object)
    t3 = pa.table({TS: np.sort(T0 + rng.integers(0, 3600000, n)).astype(np.int64), NAME: rng.choice(["a", "b"], n),
                   "pod": pa.array(hc[rng.integers(0, 2500, n)], mask=rng.random(n) < 0.07),
                   "level": pa.array(rng.choice(["info", "warn", "error"], n), mask=rng.random(n) < 0.1),
                   VALUE: pa.array(rng.normal(0, 10, n))})
    p3a = [_write("plain_str_all", 0, t3, "logs", use_dictionary=[NAME, "level"], data_page_size=4096)]
    p3b = [_write("plain_str_fallback", 0, t3, "logs", dictionary_pagesize_limit=2048, data_page_size=4096, row_group_size=3000)]
    for vid, pp in (("all_plain", p3a), ("dict_then_plain", p3b)):
        out.append((f"layout/plain_strings_{vid}_group_by", pp, _req(_logs_be(F("level", "!=", "warn"), "sum", ["pod"]), 1, 600000), ["sum"]))
        out.append((f"layout/plain_strings_{vid}_filter", pp, _req(_logs_be(F("pod", "regex", "^pod-0[0-3]"), "count", ["level"]), 1, 600000), ["count"]))
    out.append(("layout/field_datasize_int32", p2, _req(_logs_be(F(NAME, "eq", "b"), "sum", [], fieldName="bytes", fieldType="datasize"), 1, 600000), ["sum"]))
    return out


def case_compressed():
    """SNAPPY pages (inflated on the device; the CPU emulator of the tests has no inflater, so these are GPU-only cases):
    V1 and V2 data pages, many small pages and row groups, numeric dictionary pages, metrics segments with NULL tags."""
    base = _small_logs(20000, 21, null_value=True)
    be = _logs_be({"q1": F("resource.service.name", "regex", "svc-[ab]"), "q2": F("level", "!=", "warn"), "op": "and"}, "sum", ["level"])
    rq = _req(be, 1, 60000)
    variants = {
        "snappy": {"compression": "SNAPPY"},
        "snappy_v2": {"compression": "SNAPPY", "data_page_version": "2.0"},
        "snappy_small_pages_row_groups": {"compression": "SNAPPY", "row_group_size": 4500, "data_page_size": 300, "write_batch_size": 32},
        "snappy_no_dictionary": {"compression": "SNAPPY", "use_dictionary": ["resource.service.name", "level", NAME]},
        "snappy_mixed_codecs": {"compression": {TS: "SNAPPY", NAME: "NONE", "resource.service.name": "SNAPPY", "level": "NONE", VALUE: "SNAPPY"}},
    }
    out = []
    for vid, kw in variants.items():
        out.append((f"layout/{vid}", [_write("layout_" + vid, 0, base, "logs", **kw)], rq, ["sum"]))
    n = 5000
    rng = np.random.default_rng(3)
    wide = np.array([f"pod-{i:04d}" for i in range(3000)], dtype=object)
    t2 = pa.table({TS: np.sort(T0 + rng.integers(0, 3600000, n)).astype(np.int64), NAME: rng.choice(["a", "b"], n),
                   "pod": pa.array(wide[rng.integers(0, 3000, n)], mask=rng.random(n) < 0.05),
                   VALUE: pa.array(rng.integers(-1000, 1000, n).astype(np.int64)),
                   "latency$number": pa.array(rng.random(n).astype(np.float32), mask=rng.random(n) < 0.3)})
    p2 = [_write("types_snappy", 0, t2, "logs", compression="SNAPPY")]
    out.append(("layout/snappy_int64_value_wide_dict", p2, _req(_logs_be(F("pod", "regex", "^pod-1"), "sum", ["pod"]), 1, 600000), ["sum"]))
    out.append(("layout/snappy_field_float32_notnull", p2, _req(_logs_be(F(NAME, "eq", "a"), "max", [], fieldName="latency", fieldType="number"), 1, 600000), ["max"]))
    return out


def case_empty():
    e = _small_logs(0, 1)
    nz = _small_logs(500, 2)
    pe = _write("empty", 0, e, "logs")
    pn = _write("empty", 1, nz, "logs")
    return [("empty/zero_rows_only", [pe], _req(_logs_be(F(NAME, "eq", "alpha"), "sum", []), 1, 60000), ["sum"]),
            ("empty/zero_rows_plus_data", [pe, pn], _req(_logs_be(F(NAME, "eq", "alpha"), "sum", []), 2, 60000), ["sum"]),
            ("empty/nothing_passes", [pn], _req(_logs_be(F(NAME, "eq", "no-such-name"), "sum", ["level"]), 1, 60000), ["sum"])]


def gpu_only_cases():
    return case_compressed()


def all_cases():
    out = []
    for fn in (case_operators, case_time_edges, case_nulls_nan, case_schema_drift, case_physical_layouts, case_empty):
        out += fn()
    return out


# (id, builder of (paths, request), expected error kind): "unsupported" | "query" | "invalid"
def error_cases():
    p = [_write("ops", 0, _small_logs(), "logs")]
    svc = "resource.service.name"
    snappy = [_write("zstd", 0, _small_logs(), "logs", compression="ZSTD")]
    tagq = json.loads(_req(_logs_be(F(svc, "eq", "svc-a")), 1, 60000))
    tagq["isTagQuery"] = True
    noch = json.loads(_req(_logs_be(F(svc, "eq", "svc-a")), 1, 60000))
    del noch["baseExpr"]["chart"]
    ext = json.loads(_req(_logs_be(F(svc, "eq", "svc-a")), 1, 60000))
    ext["baseExpr"]["extract"] = {"regex": "(a)", "fields": [{"name": "x", "type": "string"}]}
    return [
        ("err/percentile", p, _req(_logs_be(F(svc, "eq", "svc-a"), "p95"), 1, 60000), "unsupported"),
        ("err/ces", p, _req(_logs_be(F(svc, "eq", "svc-a"), "ces"), 1, 60000), "unsupported"),
        ("err/avg", p, _req(_logs_be(F(svc, "eq", "svc-a"), "avg"), 1, 60000), "unsupported"),
        ("err/tag_query", p, json.dumps(tagq), "unsupported"),
        ("err/exemplar", p, json.dumps(noch), "unsupported"),
        ("err/extract", p, json.dumps(ext), "unsupported"),
        ("err/compressed_pages", snappy, _req(_logs_be(F(svc, "eq", "svc-a")), 1, 60000), "unsupported"),
        ("err/backreference", p, _req(_logs_be(F(svc, "regex", r"(a)\1")), 1, 60000), "unsupported"),
        ("err/missing_value_column", p, _req(_metrics_be(F(svc, "eq", "svc-a")), 1, 60000), "query"),
        ("err/missing_col_under_not", p, _req(_logs_be({"not": F("no.such.column", "eq", "x")}), 1, 60000), "query"),
        ("err/unknown_aggregate", p, _req(_logs_be(F(svc, "eq", "svc-a"), "median"), 1, 60000), "query"),
        ("err/bad_operator", p, _req(_logs_be(F(svc, "like", "x")), 1, 60000), "invalid"),
        ("err/no_filter", p, json.dumps({"baseExpr": {"dataset": "logs", "chart": {"aggregation": "sum"}}, "segmentRequests": json.loads(_req(_logs_be(F(svc, "eq", "a")), 1, 1))["segmentRequests"],
                                        "reverseSort": False, "isTagQuery": False}), "invalid"),
    ]

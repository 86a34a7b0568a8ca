"""Independent pin of the oracle's numeric evaluation: the SQL TEXT the oracle generates (golden-pinned to the reference's own
strings in test_oracle_sql_golden.py) is executed by SQLite -- an SQL engine that shares no code with the oracle's
Arrow/NumPy evaluator -- over the same Parquet rows, and the two results must agree.

DuckDB 1.3.2, the engine the reference actually calls (Commons.scala:240), is not available in this image; SQLite agrees
with it on everything these queries use (three-valued WHERE, NULL groups, sum/count/min/max over doubles, integer `%`,
ORDER BY) as long as the data holds no NaN (SQLite stores NaN as NULL), so the datasets here are NaN-free.  DuckDB-only
functions are registered as SQLite user functions: regexp_matches (RE2 partial match, used for regex and contains)."""
import math
import re
import sqlite3

import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import helpers as H
import lakeside_oracle as lo
from lakeside_b200 import synth


def _load(paths):
    """read_parquet([...], union_by_name = true) into an in-memory SQLite table t."""
    tables = [pq.read_table(p) for p in paths]
    names = []
    for t in tables:
        for n in t.schema.names:
            if n not in names:
                names.append(n)
    con = sqlite3.connect(":memory:")
    def regexp_matches(s, p, flags=""):  # DuckDB: RE2 partial match; 'i' = case-insensitive (BaseExpr.scala:485-486, 500-501)
        return None if s is None else int(re.search(p, s, re.IGNORECASE if "i" in flags else 0) is not None)

    con.create_function("regexp_matches", -1, regexp_matches, deterministic=True)
    cols = ", ".join(f'"{n}"' for n in names)
    con.execute(f"CREATE TABLE t ({cols})")
    for t in tables:
        data = {}
        for n in names:
            if n in t.schema.names:
                c = t.column(n)
                if pa.types.is_dictionary(c.type):
                    c = c.cast(c.type.value_type)
                data[n] = c.to_pylist()
            else:
                data[n] = [None] * t.num_rows
        con.executemany(f"INSERT INTO t VALUES ({', '.join('?' * len(names))})", zip(*[data[n] for n in names]))
    return con, set(names)


def _oracle_rows(req, paths):
    res = lo.evaluate_glob(req, paths)
    return res.columns, [(r.ts, r.value) + tuple(r.tags) for r in res.rows]


def _sqlite_rows(req, paths):
    con, columns = _load(paths)
    plan = lo.plan_glob(req, columns)
    sql = lo.generate_sql(req.baseExpr, plan.start_ts, plan.end_ts, step_in_millis=plan.step, global_agg=req.globalAgg,
                          non_existent_fields=plan.non_existent).replace("{tableName}", "t")
    cur = con.execute(sql)
    names = [d[0] for d in cur.description]
    return names, [tuple(r) for r in cur.fetchall()]


def _check(rq, paths, agg):
    req = lo.push_down_request_from_json(rq)
    ocols, orows = _oracle_rows(req, paths)
    scols, srows = _sqlite_rows(req, paths)
    assert list(ocols) == list(scols)  # column names and order of the JDBC result (Commons.toDataPoint branches on them)
    assert [r[0] for r in srows] == sorted(r[0] for r in srows)  # ORDER BY timestamp
    assert len(orows) == len(srows) > 0
    omap = {(r[0],) + r[2:]: r[1] for r in orows}
    smap = {(r[0],) + r[2:]: r[1] for r in srows}
    assert len(omap) == len(orows) and len(smap) == len(srows)
    assert set(omap) == set(smap)  # same (timestamp, name, group-by...) cells, NULL groups included
    for k, sv in smap.items():
        ov = omap[k]
        if sv is None:  # SQL NULL aggregate (e.g. min over all-NULL values): getDouble reads it as 0.0 (Commons.scala:426)
            assert ov == 0.0 or ov is None or (isinstance(ov, float) and math.isnan(ov))
            continue
        if agg in ("min", "max", "count"):
            assert float(sv) == ov, (k, sv, ov)
        else:
            assert abs(float(sv) - ov) <= H.SUM_RTOL * max(abs(float(sv)), abs(ov)) + 1e-300, (k, sv, ov)


@pytest.mark.parametrize("agg,rollup", [("sum", "sum"), ("sum", "count"), ("min", "min"), ("max", "max")])
def test_metrics_group_by_tags(agg, rollup):
    spec = synth.SynthSpec(dataset="metrics", rows=6000, n_names=3, cards=(6, 4, 3, 2), null_frac=0.15)
    _, paths = H.dataset("sqlite_metrics", spec, 2)
    rq = H.request_json(synth.c2_base_expr(agg, rollup), [0, 1], 10000)
    _check(rq, paths, agg)


@pytest.mark.parametrize("flt", [
    {"k": synth.TAG_SERVICE, "v": ["svc-0[0-2]"], "op": "regex"},
    {"k": synth.TAG_SERVICE, "v": ["svc-01", "svc-04"], "op": "in"},
    {"k": synth.TAG_SERVICE, "v": ["svc-01", "svc-04"], "op": "not_in"},
    {"q1": {"k": synth.TAG_SERVICE, "v": ["svc-03"], "op": "eq"}, "q2": {"k": synth.NAME, "v": ["metric_001"], "op": "eq"}, "op": "or"},
    {"q1": {"k": synth.TAG_SERVICE, "v": ["svc-0"], "op": "contains"}, "q2": {"k": "no.such.column", "v": ["x"], "op": "eq"}, "op": "or"},
])
def test_metrics_filter_shapes(flt):
    spec = synth.SynthSpec(dataset="metrics", rows=5000, n_names=3, cards=(6, 4, 3, 2), null_frac=0.15)
    _, paths = H.dataset("sqlite_metrics_filters", spec, 2)
    be = synth.c2_base_expr()
    for leaf in (flt, flt.get("q1"), flt.get("q2")):
        if leaf and "k" in leaf:
            leaf.update({"dataType": "string", "extracted": False, "computed": False})
    be["filter"] = flt
    rq = H.request_json(be, [0, 1], 10000)
    _check(rq, paths, "sum")


def test_events_step_buckets():
    spec = synth.SynthSpec(dataset="logs", rows=8000)
    _, paths = H.dataset("sqlite_logs", spec, 2)
    rq = H.request_json(synth.c1_base_expr(), [0, 1], 60000)
    _check(rq, paths, "sum")

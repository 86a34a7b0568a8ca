"""Tag queries (BaseExpr.scala:127-143): ``SELECT "tag" as "tag", COUNT(*) AS count FROM T WHERE <filter> AND
<ts range> GROUP BY "tag"`` -- the tag as the only key column, one time bucket, COUNT(*).  CPU: host planner + scan-kernel
emulator against the oracle; GPU: lk_eval through the C ABI, including the two-column JDBC view Commons.toDataPoint reads."""
import json

import pytest

import cases as C
import helpers as H
import lakeside_oracle as lo
from lakeside_b200 import synth


def _tagq(be: dict, tag: str, n: int, step: int = 60000, chart: bool = False, **kw) -> str:
    rq = json.loads(C._req(be, n, step, **kw))
    rq["isTagQuery"] = True
    rq["tagDataType"] = {"tagName": tag, "dataType": "string"}
    if not chart:
        rq["baseExpr"].pop("chart", None)  # QueryEngineV2.scala:429-466 sends tag queries without chart options
    return json.dumps(rq)


def _cases():
    svc = "resource.service.name"
    p = [C._write("ops", 0, C._small_logs(), "logs")]
    drift = [C._write("tagq_drift", 0, C._small_logs(3000, 5), "logs"), C._write("tagq_drift", 1, C._small_logs(3000, 6, tags=False), "logs")]
    m = H.dataset("m60k", synth.SynthSpec(dataset="metrics", rows=60000), 2)[1]
    return [
        ("tagq/level_by_service", p, _tagq(C._logs_be(C.F(svc, "eq", "svc-a")), "level", 1)),
        ("tagq/service_regex_or_null", p, _tagq(C._logs_be({"q1": C.F("level", "regex", "^(info|warn)$"), "q2": C.F(svc, "exists"), "op": "or"}), svc, 1)),
        ("tagq/with_chart_options_kept", p, _tagq(C._logs_be(C.F(svc, "!=", "svc-b"), "sum", ["level"]), "level", 1, chart=True)),
        ("tagq/narrow_time_range", p, _tagq(C._logs_be(C.F(svc, "in", "svc-a", "svc-b")), "level", 1, start=synth.T0 + 600000, end=synth.T0 + 1234567)),
        ("tagq/missing_filter_field_is_false", p, _tagq(C._logs_be({"q1": C.F("no.such.field", "eq", "x"), "q2": C.F(svc, "eq", "svc-a"), "op": "or"}), "level", 1)),
        ("tagq/tag_absent_in_one_file", drift, _tagq(C._logs_be(C.F(synth.NAME, "in", "alpha", "beta")), "level", 2)),
        ("tagq/metrics_pods_of_a_service", m, _tagq(synth.c2_base_expr(), synth.TAG_POD, 2, step=10000)),
    ]


CASES = _cases()


def _emul_counts(rq, paths):
    got, _ = H.emul_eval(rq, paths)
    return {k[1]: int(v[0]) for k, v in got["rows"].items()}


@pytest.mark.parametrize("cid,paths,rq", CASES, ids=[c[0] for c in CASES])
def test_tag_query_emulator_matches_oracle(cid, paths, rq):
    want = lo.evaluate_tag_query(lo.push_down_request_from_json(rq), paths)
    assert sum(want.values()) > 0
    assert _emul_counts(rq, paths) == want
    # the SQL text the oracle pins (same golden shape as ASTUtilsBaseExprTest.scala:73)
    req = lo.push_down_request_from_json(rq)
    sql = lo.generate_sql(req.baseExpr, 0, 1, is_tag_query=True, tag_data_type=req.tagDataType, non_existent_fields=set())
    tag = req.tagDataType["tagName"]
    assert sql.startswith(f'SELECT "{tag}" as "{tag}", COUNT(*) AS count FROM {{tableName}} WHERE ') and sql.endswith(f'GROUP BY "{tag}"')


def test_tag_query_errors_cpu():
    p = [C._write("ops", 0, C._small_logs(), "logs")]
    with pytest.raises(H.EmulError) as e:  # the tag is not a column of any file: DuckDB's binder fails -> stream nothing
        H.emul_eval(_tagq(C._logs_be(C.F("level", "eq", "info")), "no.such.tag", 1), p)
    assert e.value.code == 5
    with pytest.raises(lo.OracleQueryError):
        lo.evaluate_tag_query(lo.push_down_request_from_json(_tagq(C._logs_be(C.F("level", "eq", "info")), "no.such.tag", 1)), p)
    rq = json.loads(_tagq(C._logs_be(C.F("level", "eq", "info")), "level", 1))
    del rq["tagDataType"]  # SELECT * FROM (...) WHERE filter: whole rows, not an aggregate
    with pytest.raises(H.EmulError) as e:
        H.emul_eval(json.dumps(rq), p)
    assert e.value.code == 2


@pytest.mark.gpu
@pytest.mark.parametrize("cid,paths,rq", CASES, ids=[c[0] for c in CASES])
def test_tag_query_gpu_matches_oracle(cid, paths, rq):
    from lakeside_b200 import api

    api.init()
    want = lo.evaluate_tag_query(lo.push_down_request_from_json(rq), paths)
    for _ in range(2):  # cold, then from the segment cache
        res = api.eval_glob(rq, list(paths))
        try:
            tag = json.loads(rq)["tagDataType"]["tagName"]
            assert res.columns == [tag, "count"]
            assert res.tag_counts() == want
            # the JDBC view: both columns through getString (Commons.scala:407-416)
            for row in range(res.num_rows):
                v = res.get_string(row, 1)
                assert res.get_string(row, 2) == str(want[v]) and res.get_long(row, 2) == want[v]
        finally:
            res.close()

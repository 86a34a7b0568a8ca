"""Pins oracle.generate_sql against the reference's own golden SQL strings.

Vectors: query-api/src/test/scala/com/cardinal/queryapi/utils/ASTUtilsBaseExprTest.scala
  :29-75 (tag query, expected :73), :79-216 (logs chart :210 + exemplar :214), :218-289 (group-by on
  extracted field, expected :286, whitespace-normalised compare :291-305).
"""
import json
import os
import re

import lakeside_oracle as lo

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sql_golden.json")))


def _normalize(s: str) -> str:
    # ASTUtilsBaseExprTest.normalizeSql (:291-305)
    s = re.sub(r"(?m)--.*?$", "", s)
    s = re.sub(r"(?s)/\*.*?\*/", "", s)
    s = re.sub(r"\s+", " ", s)
    s = re.sub(r"\s*([(),=+\-*/%])\s*", r" \1 ", s)
    s = re.sub(r"\s+,\s*", ", ", s)
    return s.strip().lower()


def test_tag_api_should_not_do_a_select_star():
    g = GOLDEN["tag_query"]
    b = lo.to_ast_input(json.dumps(g["payload"]))["A"]
    sql = lo.generate_sql(b, 1, 1, global_agg="sum", is_tag_query=True,
                          tag_data_type={"tagName": "resource.container.name", "dataType": "string"})
    assert sql == g["expected"]


def test_query_api_payload_with_extract():
    g = GOLDEN["extract_chart"]
    b = lo.to_ast_input(json.dumps(g["payload"]))["A"]
    ts = 1694635527646
    assert lo.generate_sql(b, ts, ts, global_agg="sum") == g["expected_chart"]
    import dataclasses
    nb = dataclasses.replace(b, chartOpts=None)
    assert lo.generate_sql(nb, ts, ts, global_agg="sum") == g["expected_exemplar"]


def test_group_by_on_extracted_field():
    g = GOLDEN["groupby_extracted"]
    b = lo.to_ast_input(json.dumps(g["payload"]))["a"]
    ts = 1694635527646
    sql = lo.generate_sql(b, ts, ts, is_tag_query=False, tag_data_type=None, step_in_millis=10000, global_agg="sum")
    assert _normalize(sql) == _normalize(g["expected"])


def test_metrics_sql_shape():
    # BaseExpr.scala:390-394 -- no golden string upstream; shape check only
    b = lo.to_base_expr({"dataset": "metrics", "filter": {"k": "_cardinalhq.name", "v": ["m"], "op": "eq"},
                         "chart": {"aggregation": "max", "rollup": "max", "groupBys": ["resource.a", "gone"]}})
    sql = lo.generate_sql(b, 10, 20, global_agg="max", non_existent_fields={"gone"})
    assert sql == ('SELECT "_cardinalhq.timestamp", max(rollup_max) as value, "_cardinalhq.name" as name  , "resource.a" '
                   'FROM (SELECT * FROM {tableName} WHERE "_cardinalhq.timestamp" >= 10 AND "_cardinalhq.timestamp" < 20)  '
                   'WHERE true AND "_cardinalhq.name" = \'m\' GROUP BY "_cardinalhq.timestamp" , "resource.a", name  '
                   'ORDER BY "_cardinalhq.timestamp" ASC')


def test_missing_filter_column_is_literal_false():
    b = lo.to_base_expr({"dataset": "logs", "filter": {"q1": {"k": "a.b", "v": ["x"], "op": "eq"},
                                                        "q2": {"k": "nope", "v": ["5"], "op": "gt", "dataType": "number"},
                                                        "op": "or"},
                         "chart": {"aggregation": "count", "groupBys": []}})
    sql = lo.generate_sql(b, 0, 1, global_agg="count", non_existent_fields={"nope"})
    assert "(\"a.b\" = 'x' or false)" in sql
    sql = lo.generate_sql(b, 0, 1, global_agg="count")
    assert "(\"a.b\" = 'x' or nope > 5.0)" in sql

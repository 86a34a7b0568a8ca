"""CPU oracle for lakeside's per-segment DataExpr evaluation path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU *restatement* of the reference's algorithm for the hot path
(scan -> filter -> step-bucket -> group-by aggregate -> K-way merge -> time-grouped merge).
It must never be imported by the product package ``lakeside_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference`` legs use it.

PARITY STATUS: *partially pinned*.
  * The SQL text produced by ``generate_sql`` is pinned against the reference's own golden
    strings (query-api/src/test/scala/com/cardinal/queryapi/utils/ASTUtilsBaseExprTest.scala:73,
    :210, :214, :286) -- see tests/test_oracle_sql_golden.py.
  * The *numeric* result of executing that SQL is produced in the reference by a third-party
    engine, DuckDB 1.3.2 (org.duckdb:duckdb_jdbc:1.3.2.0, ext.gradle:24), which is neither in
    /root/reference nor installed here, and the reference has no test that pins a numeric
    aggregate.  The evaluation semantics below therefore restate documented SQL/DuckDB behaviour
    (3-valued logic, NULL grouping, NaN ordering ...): **parity unpinned** for numeric results.
    Two independent evaluators live here (a literal row-at-a-time one and a vectorised one that
    decodes Parquet with Arrow C++) and are cross-checked against each other and against
    hand-computed fixtures under tests/golden/.  The generated SQL text is additionally executed by SQLite
    (tests/test_oracle_vs_sqlite.py) over the same rows: a third evaluator that shares no code with this
    file and interprets the reference's SQL directly (NaN-free data only: SQLite has no NaN).

Reference citations are ``path:line`` relative to /root/reference.
"""
from __future__ import annotations

import json
import math
import re
from dataclasses import dataclass, field
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# --- constants: core/src/main/scala/com/cardinal/utils/Commons.scala:42-60, logs/LogCommons.scala:21-44
TIMESTAMP = "_cardinalhq.timestamp"
NAME = "_cardinalhq.name"
VALUE = "_cardinalhq.value"
MESSAGE = "_cardinalhq.message"
STEP_TS = "step_ts"
SPAN_NAME = "span.name"
SPAN_KIND = "span.kind"
LOGS, TRACES, METRICS = "logs", "traces", "metrics"
DESCENDING = "DESC"
EQ, HAS, EXISTS, NOT_EQUALS, REGEX, IN, NOT_IN, CONTAINS = "eq", "has", "exists", "!=", "regex", "in", "not_in", "contains"
GT, GE, LT, LE = "gt", "ge", "lt", "le"
SUM, COUNT, MIN, MAX, AVG = "sum", "count", "min", "max", "avg"
STRING_TYPE, NUMBER_TYPE, DURATION_TYPE, DATA_SIZE_TYPE = "string", "number", "duration", "datasize"
CES = "ces"  # BaseExpr.scala:40


class OracleUnsupported(Exception):
    """Query shape outside the hot path (extract / compute / percentile / ces / exemplar / tag queries)."""


class OracleQueryError(Exception):
    """The reference's generated SQL would fail to bind/execute in DuckDB (Commons.scala:249-253 swallows
    the exception and the glob streams nothing)."""


# ----------------------------------------------------------------------------------------------
# a1: DataExpr types + JSON decode (ASTUtils.scala:124-137, 222-229, 276-417)
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Filter:
    k: str
    v: Tuple[str, ...]
    op: str
    extracted: bool = False
    computed: bool = False
    dataType: str = STRING_TYPE


@dataclass(frozen=True)
class BinaryClause:
    q1: Any
    q2: Any
    op: str


@dataclass(frozen=True)
class NotClause:
    not_: Any


@dataclass(frozen=True)
class ChartOptions:
    aggregation: str = SUM
    groupBys: Tuple[str, ...] = ()
    type: str = "count"
    rollup: Optional[str] = None
    fieldName: Optional[str] = None
    fieldType: Optional[str] = None


@dataclass(frozen=True)
class ExtractedField:
    name: str
    type: str


@dataclass(frozen=True)
class Extractor:
    regex: str
    fields: Tuple[ExtractedField, ...]
    inputField: str = MESSAGE  # model/query/pipeline/Extractor.scala:22


@dataclass(frozen=True)
class BaseExpr:
    id: str
    dataset: str
    filter: Any
    extractor: Optional[Extractor] = None
    compute: Optional[Any] = None
    chartOpts: Optional[ChartOptions] = None
    limit: Optional[int] = 1000
    order: Optional[str] = DESCENDING
    metricType: str = "gauge"
    returnResults: bool = True


def _text(node):
    """Jackson ``JsonNode.textValue()``: the string for textual nodes, else null."""
    return node if isinstance(node, str) else None


def _to_basic_filter(node: dict) -> Filter:
    # ASTUtils.scala:276-288
    key = _text(node.get("k"))
    if key is None:
        raise ValueError("No `k` provided in filter!")
    op = _text(node.get("op"))
    if op is None:
        raise ValueError("No op provided for filter!")
    values = tuple(_text(x) for x in node["v"]) if isinstance(node.get("v"), list) else ()
    if not values and op != EXISTS:
        raise ValueError(f"No value for key = {key} provided in filter!")
    return Filter(
        k=key,
        v=values,
        op=op,
        extracted=node.get("extracted") is True,
        computed=node.get("computed") is True,
        dataType=_text(node.get("dataType")) or STRING_TYPE,
    )


def _to_binary_clause(node: dict):
    # ASTUtils.scala:379-404 -- n-ary node folded left into binary clauses, textual members skipped
    op = node.get("op")
    if op is None:
        raise ValueError("No `op` provided in binary query clause!")
    clauses = [handle_filter(el) for el in node.values() if not isinstance(el, str)]
    if len(clauses) < 2:
        raise ValueError("Atleast two clauses required in a binary clause!")
    out = None
    for c in clauses:
        out = c if out is None else BinaryClause(out, c, op)
    return out


def handle_filter(node: dict):
    # ASTUtils.scala:406-417
    if "not" in node and node["not"] is not None:
        return NotClause(handle_filter(node["not"]))
    if node.get("k") is not None:
        return _to_basic_filter(node)
    return _to_binary_clause(node)


def to_base_expr(payload, id_: Optional[str] = None) -> BaseExpr:
    """ASTUtils.toBaseExpr (ASTUtils.scala:290-377)."""
    node = json.loads(payload) if isinstance(payload, (str, bytes)) else payload
    if id_ is None:
        id_ = _text(node.get("id")) or "_"
    dataset = _text(node.get("dataset")) if "dataset" in node else METRICS
    metric_type = _text(node.get("metricType")) if "metricType" in node else "gauge"
    extractor = None
    ext = node.get("extract")
    if ext is not None:
        extractor = Extractor(
            regex=ext["regex"], fields=tuple(ExtractedField(f["name"], f["type"]) for f in ext["fields"])
        )
    compute = node.get("compute")
    chart = None
    if "chart" in node and node["chart"] is not None:
        c = node["chart"]
        gb = c.get("groupBys")
        chart = ChartOptions(
            aggregation=c["aggregation"] if isinstance(c.get("aggregation"), str) else SUM,
            groupBys=tuple(_text(g) for g in gb) if isinstance(gb, list) else (),
            type=_text(c.get("type")) if "type" in c else "count",
            rollup=_text(c.get("rollup")) if "rollup" in c else None,
            fieldName=_text(c.get("fieldName")) if "fieldName" in c else None,
            fieldType=_text(c.get("fieldType")) if "fieldType" in c else None,
        )
    if node.get("filter") is None:
        raise ValueError("No filter provided!")
    return BaseExpr(
        id=id_,
        dataset=dataset,
        filter=handle_filter(node["filter"]),
        extractor=extractor,
        compute=compute,
        chartOpts=chart,
        order=_text(node.get("order")) if "order" in node else DESCENDING,
        limit=int(node["limit"]) if "limit" in node else 1000,
        metricType=metric_type,
        returnResults=bool(node.get("returnResults", True)),
    )


def to_ast_input(payload: str) -> Dict[str, BaseExpr]:
    """ASTUtils.toASTInput: ``{"baseExpressions": {id: expr}, "formulae": [...]}`` (ASTUtils.scala:165-187)."""
    node = json.loads(payload)
    return {k: to_base_expr(v, k) for k, v in node.get("baseExpressions", {}).items()}


def filter_field_set(q) -> set:
    # BaseExpr.scala:652-663 -- NB: does NOT descend into NotClause (``case _ =>``)
    if isinstance(q, Filter):
        return {q.k}
    if isinstance(q, BinaryClause):
        return filter_field_set(q.q1) | filter_field_set(q.q2)
    return set()


def field_set(b: BaseExpr) -> set:
    # BaseExpr.scala:648-650
    return filter_field_set(b.filter) | (set(b.chartOpts.groupBys) if b.chartOpts else set())


def query_tags(b: BaseExpr) -> Dict[str, Any]:
    # BaseExpr.scala:623-646
    def exact(q):
        out: Dict[str, Any] = {}
        if isinstance(q, Filter):
            if q.op == EQ:
                out[q.k] = q.v[0]
            elif q.op == IN:
                out[q.k] = list(q.v)
        elif isinstance(q, BinaryClause) and q.op == "and":
            out.update(exact(q.q1))
            out.update(exact(q.q2))
        return out

    return exact(b.filter)


# ----------------------------------------------------------------------------------------------
# a2: PushDownRequest / SegmentRequest (model/SegmentRequest.scala:30-98)
# ----------------------------------------------------------------------------------------------
@dataclass
class SegmentRequest:
    hour: str
    dateInt: str
    segmentId: str
    sealedStatus: bool
    dataset: str
    queryTags: Dict[str, Any]
    stepInMillis: int
    customerId: str
    collectorId: str
    bucketName: str
    cName: str
    startTs: int
    endTs: int


@dataclass
class PushDownRequest:
    baseExpr: BaseExpr
    segmentRequests: List[SegmentRequest]
    processor: Optional[dict] = None
    reverseSort: bool = False
    isTagQuery: bool = False
    tagDataType: Optional[dict] = None

    @property
    def globalAgg(self) -> Optional[str]:  # SegmentRequest.scala:79
        return self.baseExpr.chartOpts.aggregation if self.baseExpr.chartOpts else None

    @property
    def rollupAgg(self) -> Optional[str]:  # SegmentRequest.scala:81
        return self.baseExpr.chartOpts.rollup if self.baseExpr.chartOpts else None

    @property
    def groupBys(self) -> Tuple[str, ...]:  # SegmentRequest.scala:71-77
        return self.baseExpr.chartOpts.groupBys if self.globalAgg is not None else ()


def push_down_request_from_json(s: str) -> PushDownRequest:
    # SegmentRequest.scala:45-60
    p = json.loads(s)
    return PushDownRequest(
        baseExpr=to_base_expr(p["baseExpr"]),
        segmentRequests=[SegmentRequest(**sr) for sr in p["segmentRequests"]],
        processor=p.get("processor"),
        reverseSort=bool(p["reverseSort"]),
        isTagQuery=bool(p["isTagQuery"]),
        tagDataType=p.get("tagDataType"),
    )


def to_parquet_file_path(sr: SegmentRequest, db_root: str = "./db") -> str:
    # Commons.scala:160-177, 256-278 (local layout)
    return f"{db_root}/{sr.customerId}/{sr.collectorId}/{sr.dateInt}/{sr.dataset}/{sr.hour}/{sr.segmentId}.parquet"


# ----------------------------------------------------------------------------------------------
# a5: BaseExpr.generateSql (BaseExpr.scala:108-513) -- the text is pinned by the reference's tests
# ----------------------------------------------------------------------------------------------
def _scala_double(d: float) -> str:
    """``java.lang.Double.toString`` (s"$labelName > ${normalizedValue()}", BaseExpr.scala:488-498)."""
    from decimal import Decimal

    if d != d:
        return "NaN"
    if d in (math.inf, -math.inf):
        return "Infinity" if d > 0 else "-Infinity"
    if d == 0:
        return "-0.0" if math.copysign(1, d) < 0 else "0.0"
    r = repr(d)
    if 1e-3 <= abs(d) < 1e7:
        return r
    t = Decimal(r).as_tuple()
    digits = "".join(map(str, t.digits))
    exp10 = len(digits) + t.exponent - 1
    digits = digits.rstrip("0") or "0"
    return ("-" if t.sign else "") + digits[0] + "." + (digits[1:] or "0") + "E" + str(exp10)


_BINARY_OPS = {"and", "or"}
_NORMALIZED_TYPES = {DURATION_TYPE, DATA_SIZE_TYPE, NUMBER_TYPE}


def _normalized_value(f: Filter) -> float:
    # BaseExpr.scala:454-459 (duration / datasize need QuantityParser -> outside the hot path)
    if f.dataType == NUMBER_TYPE:
        return float(f.v[0])
    if f.dataType in (DURATION_TYPE, DATA_SIZE_TYPE):
        raise OracleUnsupported("duration/datasize quantities (QuantityParser) are outside the hot path")
    return math.nan


def filter_sql(q, extracted: list, computed: list, top: list, non_existent: set) -> str:
    """BaseExpr.filterSqlAndAccumulateFields (BaseExpr.scala:433-513)."""
    if isinstance(q, Filter):
        acc = extracted if q.extracted else computed if q.computed else top
        if q not in acc:
            acc.append(q)
        if q.dataType in _NORMALIZED_TYPES and len(q.v) != 1:
            raise ValueError(f"filter value is a list of values for dataType: {q.dataType}")
        label = q.k
        if label in non_existent and not q.extracted and not q.computed:
            return "false"
        if "." in label:
            label = f'"{label}"'
        op = q.op
        if op in (HAS, EXISTS):
            return f"{label} IS NOT NULL"
        if op == EQ:
            return f"{label} = '{q.v[0]}'"
        if op == NOT_EQUALS:
            return f"{label} != '{q.v[0]}'"
        if op == IN:
            return f"{label} IN ({', '.join(repr_sql(v) for v in q.v)})"
        if op == NOT_IN:
            return f"{label} NOT IN ({', '.join(repr_sql(v) for v in q.v)})"
        if op == REGEX:
            return f"regexp_matches({label}, '{q.v[0]}','i')"
        if op in (GT, GE, LT, LE):
            sym = {GT: ">", GE: ">=", LT: "<", LE: "<="}[op]
            return f"{label} {sym} {_scala_double(_normalized_value(q))}"
        if op == CONTAINS:
            return f"regexp_matches({label}, '.*{q.v[0]}.*','i')"
        raise ValueError(f"Invalid operator {op}")
    if isinstance(q, BinaryClause):
        if q.op not in _BINARY_OPS:
            raise ValueError(f"unknown binary op {q.op}")
        a = filter_sql(q.q1, extracted, computed, top, non_existent)
        b = filter_sql(q.q2, extracted, computed, top, non_existent)
        return f"({a} {q.op} {b})"
    if isinstance(q, NotClause):
        return f"NOT ({filter_sql(q.not_, extracted, computed, top, non_existent)})"
    raise TypeError(q)


def repr_sql(v: str) -> str:
    return f"'{v}'"


def _timestamp_filter(start: int, end: int) -> str:
    return f'"{TIMESTAMP}" >= {start} AND "{TIMESTAMP}" < {end}'  # BaseExpr.scala:159-161


def _step_ts_sql(step: int) -> str:
    return f'("{TIMESTAMP}" - ("{TIMESTAMP}" % {step}.0)) as {STEP_TS}'  # BaseExpr.scala:163-165


def _extract_map(extracted: list) -> Dict[str, Tuple[str, str]]:
    # BaseExpr.scala:291-304 (Scala immutable Map keeps insertion order up to 4 entries; the golden strings use <= 2)
    out: Dict[str, Tuple[str, str]] = {}
    for f in extracted:
        out[f.k] = (f"nlp_struct['{f.k}'] as {f.k}", f"{f.k} IS NOT NULL")
    return out


def _extract_sql(ex: Extractor, extracted: list, sub: str) -> str:
    # BaseExpr.scala:244-265
    rx = ex.regex.replace("'", "")
    names = ", ".join(f"'{f.name}'" for f in ex.fields)
    regex_extract = f" regexp_extract(replace(\"{ex.inputField}\", '''', ''), '{rx}', [{names}]) as nlp_struct"
    regex_matches = f" regexp_matches(replace(\"{ex.inputField}\", '''', ''), '{rx}')"
    proj = ", ".join(p for p, _ in _extract_map(extracted).values())
    return f"SELECT {proj}, * FROM (SELECT {regex_extract}, * FROM ({sub}) WHERE {regex_matches})"


def _chart_field_filter(chart: ChartOptions, extracted_names: set, computed_names: set) -> str:
    # BaseExpr.scala:407-426
    if chart.fieldName is not None:
        if chart.fieldName in extracted_names or chart.fieldName in computed_names:
            return f"{chart.fieldName} IS NOT NULL"
        return f"{chart.fieldName}${chart.fieldType} IS NOT NULL"
    return "true"


def _chart_sql(chart, step, extracted, computed, fsql, sub, global_agg, dataset, non_existent) -> str:
    # BaseExpr.scala:319-405
    ext_names = set(_extract_map(extracted).keys())
    comp_names = {f.k for f in computed}

    def synthetic(g):
        return g in ext_names or g in comp_names

    existing = [g for g in chart.groupBys if synthetic(g) or g not in non_existent]
    gb = (", " + ", ".join(f'"{g}"' for g in existing)) if chart.groupBys and existing else ""
    agg = global_agg if global_agg is not None else chart.aggregation
    agg_value = chart.fieldName if chart.fieldName is not None else VALUE
    if agg_value == VALUE:
        calc = f'{agg}("{VALUE}")'
    else:
        if chart.fieldType is None:
            raise ValueError("Required property: fieldType when chartType = `field`")
        col = chart.fieldName if synthetic(chart.fieldName) else f"{chart.fieldName}${chart.fieldType}"
        calc = f"{agg}(try_cast({col} as double))"
        if chart.fieldType == DURATION_TYPE:
            calc = f"{calc} / 1000000"
        elif chart.fieldType == DATA_SIZE_TYPE:
            calc = f"{calc} / 1000"
    cfp = f", {chart.fieldName}" if (chart.fieldName is not None and chart.fieldType is not None) else ""
    cff = _chart_field_filter(chart, ext_names, comp_names)
    if dataset == METRICS:
        rollup = chart.rollup if chart.rollup is not None else SUM
        if agg.startswith("p"):
            return (
                f'SELECT "{TIMESTAMP}", {MAX}(rollup_{rollup}) as value, "{NAME}" as name {cfp} {gb} FROM ({sub}) '
                f' WHERE {cff} AND {fsql} GROUP BY "{TIMESTAMP}" {gb}, name ORDER BY "{TIMESTAMP}" ASC'
            )
        if agg == CES:
            return (
                f'SELECT "{TIMESTAMP}", 1.0 as value, "{NAME}" as name {cfp} {gb} FROM ({sub}) '
                f' WHERE {cff} AND {fsql} ORDER BY "{TIMESTAMP}" ASC'
            )
        return (
            f'SELECT "{TIMESTAMP}", {agg}(rollup_{rollup}) as value, "{NAME}" as name {cfp} {gb} FROM ({sub}) '
            f' WHERE {cff} AND {fsql} GROUP BY "{TIMESTAMP}" {gb}, name  ORDER BY "{TIMESTAMP}" ASC'
        )
    if agg.startswith("p") or CES in agg:
        return (
            f'SELECT "{TIMESTAMP}", "{VALUE}", "{NAME}" as name {cfp} {gb} FROM ({sub}) '
            f" WHERE {cff} AND {fsql} ORDER BY {TIMESTAMP} ASC"
        )
    return (
        f'SELECT {_step_ts_sql(step)}, {calc}, "{NAME}" as name {gb} FROM ({sub}) '
        f" WHERE {cff} AND {fsql} GROUP BY {STEP_TS} {gb}, name ORDER BY {STEP_TS} ASC"
    )


_PROJECTIONS = {
    LOGS: ", ".join(f'"{c}"' for c in (TIMESTAMP, VALUE, NAME, MESSAGE)),
    METRICS: ", ".join(f'"{c}"' for c in (TIMESTAMP, NAME)),
    TRACES: ", ".join(f'"{c}"' for c in (TIMESTAMP, VALUE, SPAN_NAME, SPAN_KIND)),
}


def generate_sql(
    base_expr: BaseExpr,
    start_ts: int,
    end_ts: int,
    is_tag_query: bool = False,
    tag_data_type: Optional[dict] = None,
    step_in_millis: int = 10000,
    global_agg: Optional[str] = None,
    non_existent_fields: Iterable[str] = (),
) -> str:
    """BaseExpr.generateSql -> getBaseQuery (BaseExpr.scala:108-242)."""
    non_existent = set(non_existent_fields)
    extracted: list = []
    computed: list = []
    top: list = []
    fsql = filter_sql(base_expr.filter, extracted, computed, top, non_existent)
    ts_sql = _timestamp_filter(start_ts, end_ts)
    if base_expr.dataset not in _PROJECTIONS:
        raise ValueError(f"Invalid dataset: {base_expr.dataset}")
    sql = f"SELECT * FROM {{tableName}} WHERE {ts_sql}".strip()
    if base_expr.extractor is not None:
        sql = _extract_sql(base_expr.extractor, extracted, sql)
    if base_expr.compute is not None:
        raise OracleUnsupported("compute sub-queries are outside the hot path (SURVEY §8f rank 4)")
    if base_expr.chartOpts is not None:
        sql = _chart_sql(
            base_expr.chartOpts, step_in_millis, extracted, computed, fsql, sql, global_agg, base_expr.dataset, non_existent
        )
    elif is_tag_query and tag_data_type is None:
        sql = f"SELECT * FROM ({sql}) WHERE {fsql}"
    else:
        order = base_expr.order if base_expr.order is not None else DESCENDING
        limit = base_expr.limit if base_expr.limit is not None else 1000
        sql = (
            f'SELECT {_PROJECTIONS[base_expr.dataset]}, * FROM ({sql}) WHERE {fsql} '
            f'ORDER BY "{TIMESTAMP}" {order} LIMIT {limit}'
        )
    if is_tag_query and tag_data_type is not None:
        tag = tag_data_type["tagName"]
        col = f'"{tag}"'
        # isTagSynthetic (BaseExpr.scala:146-157): extracted and computed share one accumulator
        synthetic = any(f.k == tag for f in extracted + computed)
        if synthetic:
            return f'SELECT {col} as "{tag}", COUNT(*) AS count FROM ({sql}) GROUP BY {col}'
        return f'SELECT {col} as "{tag}", COUNT(*) AS count FROM {{tableName}} WHERE {fsql} AND {ts_sql} GROUP BY {col}'
    return sql


# ----------------------------------------------------------------------------------------------
# a4-a6: evaluation of the aggregate SQL over a glob of Parquet segments
# ----------------------------------------------------------------------------------------------
T, F, N = 1, 0, 2  # Kleene truth values


def k_and(a, b):
    if a == F or b == F:
        return F
    if a == N or b == N:
        return N
    return T


def k_or(a, b):
    if a == T or b == T:
        return T
    if a == N or b == N:
        return N
    return F


def k_not(a):
    return N if a == N else (F if a == T else T)


def _compile_regex(pat: str):
    # regexp_matches(s, p, 'i'): RE2 partial match, case-insensitive (BaseExpr.scala:485-486, 500-501)
    return re.compile(pat, re.IGNORECASE)


def leaf_truth_string(f: Filter, s: Optional[str]) -> int:
    """Truth value of one filter leaf on a VARCHAR cell (None == SQL NULL)."""
    op = f.op
    if op in (HAS, EXISTS):
        return T if s is not None else F
    if s is None:
        return N
    if op == EQ:
        return T if s == f.v[0] else F
    if op == NOT_EQUALS:
        return T if s != f.v[0] else F
    if op == IN:
        return T if s in f.v else F
    if op == NOT_IN:
        return T if s not in f.v else F
    if op == REGEX:
        return T if _compile_regex(f.v[0]).search(s) else F
    if op == CONTAINS:
        return T if _compile_regex(f".*{f.v[0]}.*").search(s) else F
    raise OracleUnsupported(f"operator {op} on a string column")


def leaf_truth_number(f: Filter, x: Optional[float]) -> int:
    op = f.op
    if op in (HAS, EXISTS):
        return T if x is not None else F
    if x is None:
        return N
    c = _normalized_value(f)
    # DuckDB orders NaN as the greatest double
    def cmp(a, b):
        an, bn = a != a, b != b
        if an or bn:
            return (1 if an else 0) - (1 if bn else 0)
        return (a > b) - (a < b)

    r = cmp(float(x), c)
    if op == GT:
        return T if r > 0 else F
    if op == GE:
        return T if r >= 0 else F
    if op == LT:
        return T if r < 0 else F
    if op == LE:
        return T if r <= 0 else F
    raise OracleUnsupported(f"operator {op} on a numeric column")


def _is_numeric_op(f: Filter) -> bool:
    return f.op in (GT, GE, LT, LE)


@dataclass
class GlobPlan:
    """What Commons.toGlobResultSet decides before running SQL (Commons.scala:200-254)."""

    start_ts: int
    end_ts: int
    step: int
    non_existent: set
    group_cols: List[str]  # groupBys that exist, in chart order (BaseExpr.scala:338-346)
    value_col: str
    agg: str
    is_metrics: bool
    ts_col_name: str  # "_cardinalhq.timestamp" (metrics) or "step_ts" (events)
    divisor: float = 1.0
    needs_value_not_null: bool = False  # "fieldName$type IS NOT NULL" chart-field filter


def plan_glob(req: PushDownRequest, columns_that_exist: set) -> GlobPlan:
    b = req.baseExpr
    if req.isTagQuery:
        raise OracleUnsupported("tag queries")
    if b.chartOpts is None:
        raise OracleUnsupported("exemplar queries (no chart)")
    if b.extractor is not None or b.compute is not None:
        raise OracleUnsupported("extract/compute sub-queries")
    chart = b.chartOpts
    agg = chart.aggregation  # globalAgg = chartOpts.aggregation (Commons.scala:235)
    if agg.startswith("p") or CES in agg:
        raise OracleUnsupported("percentile / cardinality sketches")
    if agg not in (SUM, COUNT, MIN, MAX, AVG):
        raise OracleQueryError(f"unknown aggregate function {agg}")
    srs = req.segmentRequests
    non_existent = field_set(b) - set(columns_that_exist)
    group_cols = [g for g in chart.groupBys if g not in non_existent]
    is_metrics = b.dataset == METRICS
    divisor = 1.0
    needs_nn = False
    if is_metrics:
        value_col = f"rollup_{chart.rollup if chart.rollup is not None else SUM}"
        if chart.fieldName is not None:
            raise OracleUnsupported("chart field on metrics")
    elif chart.fieldName is not None and chart.fieldName != VALUE:
        if chart.fieldType is None:
            raise ValueError("Required property: fieldType when chartType = `field`")
        value_col = f"{chart.fieldName}${chart.fieldType}"
        needs_nn = True
        if chart.fieldType == DURATION_TYPE:
            divisor = 1000000.0
        elif chart.fieldType == DATA_SIZE_TYPE:
            divisor = 1000.0
    else:
        value_col = VALUE
    return GlobPlan(
        start_ts=min(s.startTs for s in srs),
        end_ts=max(s.endTs for s in srs),
        step=srs[0].stepInMillis,
        non_existent=non_existent,
        group_cols=group_cols,
        value_col=value_col,
        agg=agg,
        is_metrics=is_metrics,
        ts_col_name=TIMESTAMP if is_metrics else STEP_TS,
        divisor=divisor,
        needs_value_not_null=needs_nn,
    )


def _referenced_columns(q) -> set:
    if isinstance(q, Filter):
        return {q.k}
    if isinstance(q, BinaryClause):
        return _referenced_columns(q.q1) | _referenced_columns(q.q2)
    if isinstance(q, NotClause):
        return _referenced_columns(q.not_)
    return set()


def _bind_check(req: PushDownRequest, plan: GlobPlan, columns: set):
    """Columns the generated SQL references must exist or DuckDB's binder fails (-> empty stream)."""
    need = {TIMESTAMP, NAME, plan.value_col}
    # leaves replaced by the literal `false` do not reference their column; leaves under NOT are not in
    # fieldSet() (BaseExpr.scala:652-663), so a missing column there is a bind error.
    need |= {c for c in _referenced_columns(req.baseExpr.filter) if c not in plan.non_existent}
    missing = need - columns
    if missing:
        raise OracleQueryError(f"Binder Error: column(s) {sorted(missing)} not found")


# ---- DuckDB double ordering helpers (NaN greatest) ----
def _dmin(a: float, b: float) -> float:
    if a != a:
        return b
    if b != b:
        return a
    return b if b < a else a


def _dmax(a: float, b: float) -> float:
    if a != a or b != b:
        return a if a != a else b
    return b if b > a else a


@dataclass
class ResultRow:
    ts: int
    value: Optional[float]  # None == SQL NULL (getDouble -> 0.0)
    tags: Tuple[Optional[str], ...]  # name, then group columns


@dataclass
class GlobResult:
    columns: List[str]  # JDBC column names: ts col, value col, "name", group cols...
    rows: List[ResultRow]


def _value_sql_name(plan: GlobPlan) -> str:
    return "value" if plan.is_metrics else f"{plan.agg}(\"{plan.value_col}\")"


def _undict(t):
    """Arrow may hand dictionary-typed columns back (ARROW:schema metadata); SQL sees plain VARCHAR/DOUBLE."""
    import pyarrow as pa

    for i, f_ in enumerate(t.schema):
        if pa.types.is_dictionary(f_.type):
            t = t.set_column(i, f_.name, t.column(i).cast(f_.type.value_type))
    return t


def _read_tables(paths: Sequence[str], columns=None):
    import pyarrow.parquet as pq

    out = []
    for p in paths:
        cols = None if columns is None else [c for c in columns if c in pq.read_schema(p).names]
        out.append(_undict(pq.read_table(p, columns=cols)))
    return out


def _eval_filter_row(q, row: Dict[str, Any], plan: GlobPlan, col_is_numeric: Dict[str, bool]) -> int:
    if isinstance(q, Filter):
        if q.k in plan.non_existent and not q.extracted and not q.computed:
            return F  # literal `false` (BaseExpr.scala:462-464)
        cell = row.get(q.k)
        if col_is_numeric.get(q.k, False):
            if q.op in (HAS, EXISTS) or _is_numeric_op(q):
                return leaf_truth_number(q, cell)
            raise OracleUnsupported(f"string operator {q.op} on numeric column {q.k}")
        if _is_numeric_op(q):
            raise OracleUnsupported(f"numeric operator {q.op} on string column {q.k}")
        return leaf_truth_string(q, cell)
    if isinstance(q, BinaryClause):
        a = _eval_filter_row(q.q1, row, plan, col_is_numeric)
        b = _eval_filter_row(q.q2, row, plan, col_is_numeric)
        return k_and(a, b) if q.op == "and" else k_or(a, b)
    if isinstance(q, NotClause):
        return k_not(_eval_filter_row(q.not_, row, plan, col_is_numeric))
    raise TypeError(q)


def _finish(plan: GlobPlan, acc: Dict[tuple, list]) -> GlobResult:
    rows = []
    for key, st in acc.items():
        v = st[0]
        if v is not None and plan.agg == COUNT:
            v = float(v)
        if v is not None and plan.divisor != 1.0:
            v = v / plan.divisor
        rows.append(ResultRow(ts=key[0], value=v, tags=tuple(key[1:])))
    rows.sort(key=lambda r: (r.ts, tuple("" if t is None else "\x01" + t for t in r.tags)))
    cols = [plan.ts_col_name, _value_sql_name(plan), "name"] + list(plan.group_cols)
    return GlobResult(columns=cols, rows=rows)


def evaluate_glob_rowwise(req: PushDownRequest, paths: Sequence[str]) -> GlobResult:
    """Literal row-at-a-time execution of the SQL ``generate_sql`` produces.  Small inputs only."""
    import pyarrow as pa

    tables = _read_tables(paths)
    columns = set()
    for t in tables:
        columns |= set(t.column_names)
    plan = plan_glob(req, columns)
    _bind_check(req, plan, columns)
    col_is_numeric: Dict[str, bool] = {}
    for t in tables:
        for f_ in t.schema:
            col_is_numeric[f_.name] = not (pa.types.is_string(f_.type) or pa.types.is_large_string(f_.type))
    agg = plan.agg
    acc: Dict[tuple, list] = {}
    for t in tables:  # union_by_name: missing columns are NULL
        d = t.to_pydict()
        n = t.num_rows
        for i in range(n):
            row = {c: d[c][i] for c in d}
            ts = row.get(TIMESTAMP)
            if ts is None or not (plan.start_ts <= ts < plan.end_ts):
                continue
            if _eval_filter_row(req.baseExpr.filter, row, plan, col_is_numeric) != T:
                continue
            x = row.get(plan.value_col)
            if plan.needs_value_not_null and x is None:
                continue
            bucket = ts if plan.is_metrics else ts - ts % plan.step
            key = (bucket, row.get(NAME)) + tuple(row.get(g) for g in plan.group_cols)
            st = acc.setdefault(key, [0 if agg == COUNT else None])
            if x is None:
                continue
            x = float(x)
            if agg == COUNT:
                st[0] += 1
            elif agg in (SUM, AVG):
                st[0] = x if st[0] is None else st[0] + x
            elif agg == MIN:
                st[0] = x if st[0] is None else _dmin(st[0], x)
            elif agg == MAX:
                st[0] = x if st[0] is None else _dmax(st[0], x)
    if agg == AVG:
        raise OracleUnsupported("avg is decomposed into sum+count by QueryEngineV2.scala:280-283")
    return _finish(plan, acc)


def _select_rows(t, base_filter, start_ts: int, end_ts: int, non_existent) -> Tuple[Optional[np.ndarray], Optional[np.ndarray]]:
    """Timestamp range + WHERE clause of one file (3-valued logic; BaseExpr.scala:159-161, 433-513): returns (ts, keep mask),
    or (None, None) for a file without a timestamp column."""
    import pyarrow as pa

    n = t.num_rows
    if TIMESTAMP not in t.column_names:
        return None, None
    ts_col = t.column(TIMESTAMP).combine_chunks()
    ts = ts_col.fill_null(0).to_numpy(zero_copy_only=False).astype(np.int64)
    sel = (ts >= start_ts) & (ts < end_ts)
    if ts_col.null_count:
        sel &= ~ts_col.is_null().to_numpy(zero_copy_only=False)

    def leaf_vec(f: Filter) -> np.ndarray:
        if f.k in non_existent and not f.extracted and not f.computed:
            return np.full(n, F, np.uint8)
        if f.k not in t.column_names:  # union_by_name NULL column
            return np.full(n, F if f.op in (HAS, EXISTS) else N, np.uint8)
        col = t.column(f.k).combine_chunks()
        if pa.types.is_string(col.type) or pa.types.is_large_string(col.type):
            if _is_numeric_op(f):
                raise OracleUnsupported(f"numeric operator {f.op} on string column {f.k}")
            de = col.dictionary_encode()
            dvals = de.dictionary.to_pylist()
            lut = np.array([leaf_truth_string(f, s) for s in dvals] + [leaf_truth_string(f, None)], np.uint8)
            idx = de.indices.fill_null(len(dvals)).to_numpy(zero_copy_only=False).astype(np.int64)
            return lut[idx]
        if not (f.op in (HAS, EXISTS) or _is_numeric_op(f)):
            raise OracleUnsupported(f"string operator {f.op} on numeric column {f.k}")
        isnull = col.is_null().to_numpy(zero_copy_only=False)
        x = col.fill_null(0).to_numpy(zero_copy_only=False).astype(np.float64)
        out = np.empty(n, np.uint8)
        if f.op in (HAS, EXISTS):
            out[:] = T
            out[isnull] = F
            return out
        c = _normalized_value(f)
        xn = np.isnan(x)
        cn = c != c
        with np.errstate(invalid="ignore"):
            if f.op == GT:
                r = (x > c) | (xn & (not cn))
            elif f.op == GE:
                r = (x >= c) | xn
            elif f.op == LT:
                r = (x < c) | ((~xn) & cn)
            else:
                r = (x <= c) | cn
            if f.op in (GT, LT):
                r = r & ~(xn & cn)
        out[:] = np.where(r, T, F)
        out[isnull] = N
        return out

    def tree(q) -> np.ndarray:
        if isinstance(q, Filter):
            return leaf_vec(q)
        if isinstance(q, BinaryClause):
            a, b = tree(q.q1), tree(q.q2)
            if q.op == "and":
                return np.where((a == F) | (b == F), F, np.where((a == N) | (b == N), N, T)).astype(np.uint8)
            return np.where((a == T) | (b == T), T, np.where((a == N) | (b == N), N, F)).astype(np.uint8)
        if isinstance(q, NotClause):
            a = tree(q.not_)
            return np.where(a == N, N, np.where(a == T, F, T)).astype(np.uint8)
        raise TypeError(q)

    sel &= tree(base_filter) == T
    return ts, sel


def evaluate_tag_query(req: PushDownRequest, paths: Sequence[str]) -> Dict[Optional[str], int]:
    """Tag query with a tagDataType (BaseExpr.scala:127-143, the non-synthetic branch):
    ``SELECT "tag" as "tag", COUNT(*) AS count FROM T WHERE <filter> AND ts >= S AND ts < E GROUP BY "tag"`` over the glob
    (Commons.scala:200-254: start = min, end = max over the segment requests).  Returns {tag value (None = the NULL group):
    count}; the reference reads both columns through ``getString`` (Commons.scala:407-416)."""
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.parquet as pq

    if not req.isTagQuery or not req.tagDataType:
        raise OracleUnsupported("tag query without a tagDataType (SELECT *)")
    b = req.baseExpr
    if b.extractor is not None or b.compute is not None:
        raise OracleUnsupported("extract/compute sub-queries")
    tag = req.tagDataType["tagName"]
    columns = set()
    for p_ in paths:
        columns |= set(pq.read_schema(p_).names)
    if tag not in columns:
        raise OracleQueryError(f"Binder Error: column {tag} not found")
    referenced = _referenced_columns(b.filter)
    non_existent = {c for c in referenced if c not in columns}  # only the filter's fields: a tag query has no group-bys to drop
    srs = req.segmentRequests
    start_ts, end_ts = min(s.startTs for s in srs), max(s.endTs for s in srs)
    want = sorted({TIMESTAMP, tag, *(c for c in referenced if c in columns)})
    out: Dict[Optional[str], int] = {}
    for t in _read_tables(paths, want):
        ts, sel = _select_rows(t, b.filter, start_ts, end_ts, non_existent)
        if ts is None:
            raise OracleQueryError(f"Binder Error: column {TIMESTAMP} not found")
        idx = np.nonzero(sel)[0]
        if tag in t.column_names:
            col = t.column(tag).combine_chunks()
            if not (pa.types.is_string(col.type) or pa.types.is_large_string(col.type)):
                raise OracleUnsupported(f"tag query on non-string column {tag}")
            vals = col.take(pa.array(idx)).to_pylist()
        else:
            vals = [None] * len(idx)  # union_by_name: NULL for this file
        for v in vals:
            out[v] = out.get(v, 0) + 1
    return out


def evaluate_glob(req: PushDownRequest, paths: Sequence[str], aggs: Optional[Sequence[Tuple[str, str]]] = None):
    """Vectorised evaluation (Arrow C++ Parquet decode + NumPy), same semantics as ``evaluate_glob_rowwise``.

    ``aggs``: optional list of (aggregation, value column) evaluated in ONE pass sharing filter and
    grouping (the fused multi-aggregate extension); returns a dict with key arrays and one value array
    per aggregate instead of a GlobResult.
    """
    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.parquet as pq

    schemas = [pq.read_schema(p) for p in paths]
    columns = set()
    for s in schemas:
        columns |= set(s.names)
    plan = plan_glob(req, columns)
    _bind_check(req, plan, columns)
    if plan.agg == AVG:
        raise OracleUnsupported("avg is decomposed into sum+count by QueryEngineV2.scala:280-283")
    multi = aggs is not None
    agg_list = list(aggs) if multi else [(plan.agg, plan.value_col)]
    for _, vc in agg_list:
        if vc not in columns:
            raise OracleQueryError(f"Binder Error: column {vc} not found")
    filt_cols = sorted(c for c in _referenced_columns(req.baseExpr.filter) if c not in plan.non_existent)
    key_cols = [NAME] + plan.group_cols
    want = sorted({TIMESTAMP, *key_cols, *filt_cols, *(vc for _, vc in agg_list)})

    # global dictionaries (sorted) for key columns so group codes agree across files
    per_file = _read_tables(paths, want)

    def string_col(t, c):
        if c not in t.column_names:
            return None
        col = t.column(c).combine_chunks()
        if not (pa.types.is_string(col.type) or pa.types.is_large_string(col.type)):
            return None
        return col

    key_dicts: Dict[str, List[str]] = {}
    for c in key_cols:
        vals = set()
        for t in per_file:
            col = string_col(t, c)
            if col is None:
                if c in t.column_names:
                    raise OracleUnsupported(f"group-by on non-string column {c}")
                continue
            vals |= set(pc.unique(col).drop_null().to_pylist())
        key_dicts[c] = sorted(vals)

    # walk the filter tree once per file on dictionary entries
    ts_all, keycode_all, val_all = [], [[] for _ in key_cols], [[] for _ in agg_list]
    for t in per_file:
        n = t.num_rows
        ts, sel = _select_rows(t, req.baseExpr.filter, plan.start_ts, plan.end_ts, plan.non_existent)
        if ts is None:
            continue
        vals = []
        for _, vc in agg_list:
            if vc in t.column_names:
                vcol = t.column(vc).combine_chunks()
                vnull = vcol.is_null().to_numpy(zero_copy_only=False)
                v = vcol.fill_null(0).to_numpy(zero_copy_only=False).astype(np.float64)
            else:
                vnull = np.ones(n, bool)
                v = np.zeros(n)
            vals.append((v, vnull))
        if plan.needs_value_not_null:
            sel &= ~vals[0][1]
        idx = np.nonzero(sel)[0]
        ts_all.append(ts[idx])
        for j, c in enumerate(key_cols):
            col = string_col(t, c)
            if col is None:
                keycode_all[j].append(np.full(len(idx), -1, np.int64))
            else:
                gd = pa.array(key_dicts[c], type=col.type)
                codes = pc.index_in(col.take(pa.array(idx)), value_set=gd).fill_null(-1)
                keycode_all[j].append(codes.to_numpy(zero_copy_only=False).astype(np.int64))
        for j, (v, vnull) in enumerate(vals):
            val_all[j].append((v[idx], vnull[idx]))

    if ts_all:
        ts = np.concatenate(ts_all)
        kc = [np.concatenate(k) for k in keycode_all]
        vs = [(np.concatenate([a for a, _ in v]), np.concatenate([m for _, m in v])) for v in val_all]
    else:
        ts = np.zeros(0, np.int64)
        kc = [np.zeros(0, np.int64) for _ in key_cols]
        vs = [(np.zeros(0), np.zeros(0, bool)) for _ in agg_list]
    bucket = ts if plan.is_metrics else ts - ts % plan.step

    # group: lexicographic unique over (bucket, codes...) ; codes -1 == NULL group
    keys = np.stack([bucket] + kc, axis=1) if len(ts) else np.zeros((0, 1 + len(kc)), np.int64)
    if len(ts):
        uniq, inv = np.unique(keys, axis=0, return_inverse=True)
        inv = inv.reshape(-1)
    else:
        uniq, inv = keys, np.zeros(0, np.int64)
    ng = len(uniq)
    out_vals, out_null = [], []
    for (agg, _), (v, vnull) in zip(agg_list, vs):
        nn = ~vnull
        cnt = np.bincount(inv[nn], minlength=ng)
        if agg == COUNT:
            out_vals.append(cnt.astype(np.float64))
            out_null.append(np.zeros(ng, bool))
            continue
        if agg == SUM:
            r = np.zeros(ng)
            np.add.at(r, inv[nn], v[nn])  # sequential, row order: the oracle's fixed summation order
        elif agg in (MIN, MAX):
            x = v[nn]
            g = inv[nn]
            isn = np.isnan(x)
            r = np.full(ng, np.inf if agg == MIN else -np.inf)
            if agg == MIN:
                np.minimum.at(r, g[~isn], x[~isn])
                only_nan = (np.bincount(g[~isn], minlength=ng) == 0) & (cnt > 0)
                r[only_nan] = np.nan  # NaN is the greatest double: min is NaN only if every value is NaN
            else:
                np.maximum.at(r, g[~isn], x[~isn])
                has_nan = np.bincount(g[isn], minlength=ng) > 0
                r[has_nan] = np.nan
        else:
            raise OracleQueryError(f"unknown aggregate function {agg}")
        isnull = cnt == 0
        r[isnull] = 0.0
        if plan.divisor != 1.0:
            r = r / plan.divisor
        out_vals.append(r)
        out_null.append(isnull)

    if multi:
        return {
            "plan": plan,
            "ts": uniq[:, 0] if ng else np.zeros(0, np.int64),
            "key_cols": key_cols,
            "key_dicts": key_dicts,
            "key_codes": [uniq[:, 1 + j] for j in range(len(key_cols))] if ng else [np.zeros(0, np.int64) for _ in key_cols],
            "values": out_vals,
            "nulls": out_null,
        }
    rows = []
    for i in range(ng):
        tags = tuple(None if uniq[i, 1 + j] < 0 else key_dicts[c][uniq[i, 1 + j]] for j, c in enumerate(key_cols))
        rows.append(ResultRow(ts=int(uniq[i, 0]), value=None if out_null[0][i] else float(out_vals[0][i]), tags=tags))
    rows.sort(key=lambda r: (r.ts, tuple("" if t is None else "\x01" + t for t in r.tags)))
    cols = [plan.ts_col_name, _value_sql_name(plan), "name"] + list(plan.group_cols)
    return GlobResult(columns=cols, rows=rows)


# ----------------------------------------------------------------------------------------------
# a7-a8: ResultSet row -> DataPoint -> map-sketch SketchInput
# ----------------------------------------------------------------------------------------------
@dataclass
class DataPoint:  # model/DataPoint.scala:19
    timestamp: int
    value: float
    tags: Dict[str, str]


@dataclass
class SketchInput:  # utils/ast/SketchInput.scala:73 (map-sketch subset)
    timestamp: int
    tags: Dict[str, str]
    sketch: Dict[str, float]
    sketchType: str = "map"


def to_data_points(res: GlobResult, query_tags: Dict[str, Any]) -> List[DataPoint]:
    """Commons.toDataPoint, aggregate branch (Commons.scala:424-461)."""
    out = []
    for r in res.rows:
        tags = {}
        for cname, tv in zip(res.columns[2:], r.tags):
            if tv is not None and tv != "null" and tv != "":
                tags[cname] = tv
        if not tags:
            tags.update({k: v for k, v in query_tags.items()})
        out.append(DataPoint(timestamp=r.ts, value=0.0 if r.value is None else r.value, tags=tags))
    return out


def push_down_aggregator_stage(req: PushDownRequest, dps: List[DataPoint]) -> List[SketchInput]:
    """PushDownAggregatorStage map-sketch branch (PushDownAggregatorStage.scala:95-106)."""
    agg = req.globalAgg
    return [SketchInput(timestamp=d.timestamp, tags=d.tags, sketch={agg: d.value}) for d in dps]


# ----------------------------------------------------------------------------------------------
# a9: wire format of a per-segment stream element.
#   producer  Commons.dataPointResponseToSSE (Commons.scala:474-502) -> GenericSSEPayload(id="_", type="data", message)
#             .toChunkStreamPart = "data: " + toJson + "\r\n\r\n" (SSEMessage.scala:23-34); JSON by Jackson
#             (atlas-json defaults: non-finite doubles are quoted, "NaN" / "Infinity" / "-Infinity")
#   consumer  EventStreamParser -> Json.decode(sse.data)("message") (QueryEngineV2.scala:133-140)
#             -> SegmentSequencer.decode (SegmentSequencer.scala:35-101)
# ----------------------------------------------------------------------------------------------
def _json_double(v: float):
    if v != v:
        return "NaN"
    if v in (math.inf, -math.inf):
        return "Infinity" if v > 0 else "-Infinity"
    return v


def to_sse(elements: List[Any]) -> bytes:
    """dataPointResponseToSSE over DataPoints (Left: "exemplar") and map-sketch SketchInputs (Right: "sketch")."""
    import json as _json

    out = []
    for e in elements:
        if isinstance(e, DataPoint):
            msg = {"timestamp": int(e.timestamp), "value": _json_double(float(e.value)), "tags": dict(e.tags), "type": "exemplar"}
        else:
            msg = {"timestamp": int(e.timestamp), "tags": dict(e.tags), "type": "sketch", "sketchType": e.sketchType,
                   "sketch": {k: _json_double(float(v)) for k, v in e.sketch.items()}}
        out.append("data: " + _json.dumps({"id": "_", "type": "data", "message": msg}, separators=(",", ":"), ensure_ascii=False) + "\r\n\r\n")
    return "".join(out).encode("utf-8")


def _as_double(v) -> float:  # SegmentSequencer.scala:35-45
    if isinstance(v, bool):
        return math.nan
    if isinstance(v, (int, float)):
        return float(v)
    if isinstance(v, str):
        if v in ("NaN", "nan"):
            return math.nan
        if v in ("Infinity", "+Infinity"):
            return math.inf
        if v == "-Infinity":
            return -math.inf
        try:
            return float(v)
        except ValueError:
            return math.nan
    return math.nan


def _as_long(v) -> int:  # SegmentSequencer.scala:47-51
    if isinstance(v, bool):
        return 0
    if isinstance(v, (int, float)):
        return int(v)
    if isinstance(v, str):
        try:
            return int(v)
        except ValueError:
            return 0
    return 0


def sse_decode(stream: bytes) -> List[Any]:
    """The consumer side: split the event stream, keep events whose JSON has a "message", decode each message as
    SegmentSequencer.decode does (tags must be JSON strings; map-sketch values tolerate "NaN" / "Infinity")."""
    import json as _json

    out = []
    for event in stream.decode("utf-8").split("\r\n\r\n"):
        if not event.strip():
            continue
        assert event.startswith("data: "), event[:40]
        obj = _json.loads(event[len("data: "):], parse_int=float)  # ujson: every JSON number is a Double ("-0" is -0.0)
        if "message" not in obj:
            continue  # heartbeat / done
        m = obj["message"]
        tags = {}
        for k, v in m["tags"].items():
            if not isinstance(v, str):
                raise OracleQueryError("tag value is not a JSON string")  # ujson `.str` throws (SegmentSequencer.scala:68)
            tags[k] = v
        if m["type"] == "exemplar":
            out.append(DataPoint(timestamp=_as_long(m["timestamp"]), value=_as_double(m["value"]), tags=tags))
        elif m["type"] == "sketch":
            if m["sketchType"] != "map":
                raise OracleUnsupported("only map sketches are restated")
            out.append(SketchInput(timestamp=_as_long(m["timestamp"]), tags=tags,
                                   sketch={k: _as_double(v) for k, v in m["sketch"].items()}, sketchType="map"))
    return out


# ----------------------------------------------------------------------------------------------
# a10: K-way merge.  ``sources.fold(Source.empty)((s1, s2) => s1.mergeSorted(s2))``
#      (Commons.scala:391-392, WorkerApi.scala:173, QueryEngineV2.scala:96).
# akka-stream 2.6.20 MergeSorted (third-party, not in tree) emits the LEFT head only if left < right,
# so on equal timestamps the RIGHT (later-folded) source is drained first.  Left-deep fold => for equal
# ts, elements come out by source index DESCENDING, original order within a source.
# ----------------------------------------------------------------------------------------------
def merge_sorted_pair(left: list, right: list, key=lambda e: e.timestamp, reverse: bool = False) -> list:
    out = []
    i = j = 0
    while i < len(left) and j < len(right):
        kl, kr = key(left[i]), key(right[j])
        less = kl > kr if reverse else kl < kr
        if less:
            out.append(left[i])
            i += 1
        else:
            out.append(right[j])
            j += 1
    out.extend(left[i:])
    out.extend(right[j:])
    return out


def merge_sorted_source(sources: List[list], key=lambda e: e.timestamp, reverse: bool = False) -> list:
    merged: list = []
    for s in sources:
        merged = merge_sorted_pair(merged, s, key, reverse)
    return merged


def merge_sorted_arrays(ts_list: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised statement of the same order: returns (source index, position) of every output element.

    Order = (ts asc, source index desc, position asc)."""
    src = np.concatenate([np.full(len(t), i, np.int64) for i, t in enumerate(ts_list)]) if ts_list else np.zeros(0, np.int64)
    pos = np.concatenate([np.arange(len(t), dtype=np.int64) for t in ts_list]) if ts_list else np.zeros(0, np.int64)
    ts = np.concatenate(ts_list) if ts_list else np.zeros(0, np.int64)
    order = np.lexsort((pos, -src, ts))
    return src[order], pos[order]


# ----------------------------------------------------------------------------------------------
# a11: TimeGroupedSketchAggregator, map-sketch merge (TimeGroupedSketchAggregator.scala:57-114, 200-253)
# ----------------------------------------------------------------------------------------------
DOUBLE_MAX = 1.7976931348623157e308


def merge_map_sketch(existing: Dict[str, float], incoming: Dict[str, float]) -> Dict[str, float]:
    # SimpleSketchMerger.mergeSketch (:63-93)
    out = {}
    for key in list(existing.keys()) + [k for k in incoming if k not in existing]:
        if key in (SUM, COUNT):
            out[key] = existing.get(key, 0.0) + incoming.get(key, 0.0)
        elif key == MIN:
            out[key] = _java_min(existing.get(key, DOUBLE_MAX), incoming.get(key, DOUBLE_MAX))
        elif key == MAX:
            out[key] = _java_max(existing.get(key, -DOUBLE_MAX), incoming.get(key, -DOUBLE_MAX))
        else:
            raise ValueError(key)  # scala.MatchError
    return out


def _java_min(a: float, b: float) -> float:
    if a != a or b != b:
        return math.nan
    if a == 0.0 and b == 0.0:
        return a if math.copysign(1, a) < 0 else b
    return a if a <= b else b


def _java_max(a: float, b: float) -> float:
    if a != a or b != b:
        return math.nan
    if a == 0.0 and b == 0.0:
        return a if math.copysign(1, a) > 0 else b
    return a if a >= b else b


def time_grouped_aggregate(stream: List[SketchInput], num_buffers: int = 4, now_ms: Optional[int] = None):
    """Returns [(timestamp, [SketchInput merged per tags map])...] in emission order.

    Restates the ring-buffer behaviour: ``findBuffer`` / ``flush`` / cutoffTime dropping
    (TimeGroupedSketchAggregator.scala:141-150, 187-193, 200-228, 237-253); numBuffers=4 as wired by
    EvalUtils.astEvalFlow (EvalUtils.scala:27-37)."""
    import time as _time

    now = int(_time.time() * 1000) if now_ms is None else now_ms
    buf: List[Dict[tuple, SketchInput]] = [dict() for _ in range(num_buffers)]
    order: List[List[tuple]] = [[] for _ in range(num_buffers)]
    timestamps = [0] * num_buffers
    cutoff = 0
    out = []

    def find_buffer(t):
        mn = 0
        for i in range(num_buffers):
            if timestamps[i] == t:
                return i
            if i > 0 and timestamps[i] < timestamps[i - 1]:
                mn = i
        return -mn - 1

    def group(i):
        return [buf[i][k] for k in order[i]]

    def aggregate(i, v: SketchInput):
        k = tuple(sorted(v.tags.items()))
        cur = buf[i].get(k)
        if cur is None:
            buf[i][k] = SketchInput(v.timestamp, dict(v.tags), dict(v.sketch), v.sketchType)
            order[i].append(k)
        else:
            cur.sketch = merge_map_sketch(cur.sketch, v.sketch)

    for v in stream:
        t = v.timestamp
        if t > now or t <= cutoff:
            continue  # droppedRecords
        i = find_buffer(t)
        if i >= 0:
            aggregate(i, v)
        else:
            pos = -i - 1
            if timestamps[pos] > 0:
                out.append((timestamps[pos], group(pos)))
            cutoff = timestamps[pos]
            buf[pos], order[pos] = {}, []
            aggregate(pos, v)
            timestamps[pos] = t
    pending = [(timestamps[i], group(i)) for i in range(num_buffers) if timestamps[i] > 0]
    pending.sort(key=lambda g: g[0])
    out.extend(pending)
    return out


# ----------------------------------------------------------------------------------------------
# a12: BaseExpr.eval / getFromSketch / getTransformerFunc (BaseExpr.scala:47-95, 665-695; ASTUtils.scala:87-89, 190-219)
# ----------------------------------------------------------------------------------------------
def get_from_sketch(sketch: Dict[str, float], aggregation: str) -> float:
    if aggregation == AVG:
        s = sketch.get(SUM, math.nan)
        c = sketch.get(COUNT, math.nan)
        if c == 0:
            return math.nan if (s == 0 or s != s) else math.copysign(math.inf, s)
        return s / c
    return sketch.get(aggregation, math.nan)


def transformer(chart_type: str, metric_type: str, dataset: str, step_ms: int):
    secs = step_ms // 1000  # integer division first (ASTUtils.scala:201, 204, 213)
    if dataset == METRICS:
        if chart_type == "count" and metric_type == "rate":
            return lambda v: v * secs
        if chart_type == "rate" and metric_type == "count":
            return lambda v: v / secs if secs else (math.nan if v == 0 or v != v else math.copysign(math.inf, v))
        return lambda v: v
    if chart_type == "rate":
        return lambda v: v / secs if secs else (math.nan if v == 0 or v != v else math.copysign(math.inf, v))
    return lambda v: v


def to_group_by_key(group_by_keys: Iterable[str], tags: Dict[str, Any]) -> str:
    return ":".join(str(tags.get(k, "")) for k in sorted(set(group_by_keys)))


def base_expr_eval(b: BaseExpr, sketches: List[SketchInput], step_ms: int, aggregation: Optional[str] = None):
    """BaseExpr.eval over one SketchGroup: {groupKey -> (ts, value, tags)}."""
    out = {}
    if b.chartOpts is None:
        return out
    fn = transformer(b.chartOpts.type, b.metricType, b.dataset, step_ms)
    agg = aggregation or b.chartOpts.aggregation
    keys = set(b.chartOpts.groupBys)
    for s in sketches:
        v = fn(get_from_sketch(s.sketch, agg))
        out["default" if not keys else to_group_by_key(keys, s.tags)] = (s.timestamp, v, s.tags)
    return out


def constant_expr_eval(value: float, group_by_keys, ts: int, sketches: List[SketchInput]):
    """ASTUtils.eval, ConstantExpr branch (ASTUtils.scala:50-64): one "default" result without group-bys, else one result per
    sketch input of the group, keyed by the formula's final grouping (a later input of the same key replaces an earlier one)."""
    keys = set(group_by_keys)
    if not keys:
        return {"default": (ts, float(value), {})}
    return {to_group_by_key(keys, s.tags): (ts, float(value), s.tags) for s in sketches}


def formula_eval(op: str, e1_map: Dict[str, tuple], e2_map: Dict[str, tuple]) -> Dict[str, tuple]:
    """Formula.eval (Formula.scala:32-69) for one SketchGroup: both sides already evaluated into {groupKey: (ts, value, tags)}.
    Equal keys are combined; `add` takes a missing side as 0 (with the other side's timestamp and tags), the other operators
    give no result; a zero divisor gives no result.  The result carries e1's timestamp and tags."""
    out = {}
    for k in list(e1_map.keys()) + [k for k in e2_map if k not in e1_map]:
        r1, r2 = e1_map.get(k), e2_map.get(k)
        if r1 is None and r2 is None:
            continue
        if r2 is None:
            if op != "add":
                continue
            r2 = (r1[0], 0.0, r1[2])
        elif r1 is None:
            if op != "add":
                continue
            r1 = (r2[0], 0.0, r2[2])
        a, b = float(r1[1]), float(r2[1])
        if op == "add":
            v = a + b
        elif op == "sub":
            v = a - b
        elif op == "mul":
            v = a * b
        elif op == "div":
            if not (b != 0):  # e2Result.value != 0: NaN passes, +-0.0 does not
                continue
            v = a / b
        else:
            raise ValueError(op)
        out[k] = (r1[0], v, r1[2])
    return out


# ----------------------------------------------------------------------------------------------
# a3: evaluatePushDownRequest -- globs of 10 (local) / 5 (S3), merged by timestamp (Commons.scala:343-397)
# ----------------------------------------------------------------------------------------------
def evaluate_push_down_request(req: PushDownRequest, db_root: str, local_parquet: bool = True, evaluator=evaluate_glob):
    import os

    glob_size = 10 if local_parquet else 5
    srs = req.segmentRequests
    if not srs:
        return [DataPoint(timestamp=-1, value=-1.0, tags={})]  # Commons.scala:393-396 sentinel
    sources = []
    for g in range(0, len(srs), glob_size):
        group = srs[g : g + glob_size]
        sub = PushDownRequest(req.baseExpr, group, req.processor, req.reverseSort, req.isTagQuery, req.tagDataType)
        paths = [to_parquet_file_path(s, db_root) for s in group]
        try:
            if not all(os.path.exists(p) for p in paths):
                raise OracleQueryError("IO Error: No files found")
            res = evaluator(sub, paths)
            dps = to_data_points(res, group[0].queryTags)
            sources.append(push_down_aggregator_stage(sub, dps))
        except OracleQueryError:
            sources.append([])  # Commons.scala:249-253, 338-340: log and stream nothing
    return merge_sorted_source(sources)

"""Synthetic observability segments (schema, writer options, seeds, distributions).

Column names are the ones the reference hard-codes (core/src/main/scala/com/cardinal/utils/Commons.scala:45-68,
BaseExpr.scala:377-394); file layout follows Commons.getDbPath / toParquetFilePath (Commons.scala:160-177, 256-278):
``<root>/<customerId>/<collectorId>/<dateInt>/<dataset>/<hour>/<segmentId>.parquet``.

Writer is pinned (encoded size depends on it): pyarrow ``write_table(compression="NONE", use_dictionary=True,
data_page_version="1.0", data_page_size=1 MiB, row_group_size=1 Mi rows, write_statistics=True)``, rows sorted by
timestamp inside a segment.  Seeds: ``PCG64(20240 + segment_index)``.
"""
from __future__ import annotations

import os
from concurrent.futures import ProcessPoolExecutor
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

T0 = 1699999200000  # 2023-11-14T22:00:00Z, hour aligned
HOUR_MS = 3600000
TIMESTAMP = "_cardinalhq.timestamp"
NAME = "_cardinalhq.name"
VALUE = "_cardinalhq.value"
TELEMETRY_TYPE = "_cardinalhq.telemetry_type"
TAG_SERVICE = "resource.service.name"
TAG_NAMESPACE = "resource.k8s.namespace.name"
TAG_POD = "resource.k8s.pod.name"
TAG_ZONE = "resource.cloud.availability_zone"
GROUP_TAGS = (TAG_NAMESPACE, TAG_POD, TAG_ZONE)

CUSTOMER_ID = "org-synth"
COLLECTOR_ID = "collector-0"
DATE_INT = "20231114"
HOUR = "22"


@dataclass
class SynthSpec:
    """One dataset flavour.  ``cards`` = distinct values of (service, namespace, pod, zone)."""

    dataset: str = "metrics"  # "metrics" | "logs"
    rows: int = 1 << 20
    step: int = 10000
    n_names: int = 64
    cards: Tuple[int, int, int, int] = (16, 32, 32, 16)
    null_frac: float = 0.05
    int_values: bool = False  # twin dataset: U{0..2^20} as double => sums exact in any order
    row_group_size: int = 1 << 20
    seed_base: int = 20240
    prefixes: Tuple[str, str, str, str] = ("svc", "ns", "pod", "az")
    extra_nan_inf: bool = False  # sprinkle NaN / +-inf / -0.0 into value columns (edge-case tests)
    drop_columns: Tuple[str, ...] = ()  # schema drift between segments (union_by_name tests)
    compression: str = "NONE"  # the benchmark's writer is pinned to NONE; "SNAPPY" for the format-breadth tests


def tag_values(prefix: str, k: int) -> List[str]:
    w = max(2, len(str(k - 1)))
    return [f"{prefix}-{i:0{w}d}" for i in range(k)]


def name_values(n: int) -> List[str]:
    return [f"metric_{i:03d}" for i in range(n)]


def segment_path(root: str, dataset: str, segment_id: str) -> str:
    return os.path.join(root, CUSTOMER_ID, COLLECTOR_ID, DATE_INT, dataset, HOUR, f"{segment_id}.parquet")


def segment_id_for(index: int) -> str:
    return f"tbl_{index:06d}"


def make_table(spec: SynthSpec, index: int):
    import pyarrow as pa

    n = spec.rows
    rng = np.random.Generator(np.random.PCG64(spec.seed_base + index))
    if spec.dataset == "metrics":
        ts = T0 + spec.step * rng.integers(0, HOUR_MS // spec.step, n)
    else:
        ts = T0 + rng.integers(0, HOUR_MS, n)
    ts = np.sort(ts).astype(np.int64)

    def dict_col(values: Sequence[str], codes: np.ndarray, null_frac: float):
        mask = rng.random(n) < null_frac if null_frac > 0 else None
        idx = pa.array(codes.astype(np.int32), mask=mask)
        return pa.DictionaryArray.from_arrays(idx, pa.array(list(values), type=pa.string()))

    w = 1.0 / np.arange(1, spec.n_names + 1) ** 1.1
    w /= w.sum()
    cols: Dict[str, object] = {}
    cols[TIMESTAMP] = pa.array(ts)
    cols[NAME] = dict_col(name_values(spec.n_names), rng.choice(spec.n_names, n, p=w), 0.0)
    for tag, pre, k in zip((TAG_SERVICE,) + GROUP_TAGS, spec.prefixes, spec.cards):
        cols[tag] = dict_col(tag_values(pre, k), rng.integers(0, k, n), spec.null_frac)

    def values():
        if spec.int_values:
            v = rng.integers(0, 1 << 20, n).astype(np.float64)
        else:
            v = rng.lognormal(0.0, 2.0, n)
        if spec.extra_nan_inf:
            sp = rng.random(n)
            v[sp < 0.01] = np.nan
            v[(sp >= 0.01) & (sp < 0.015)] = np.inf
            v[(sp >= 0.015) & (sp < 0.02)] = -np.inf
            v[(sp >= 0.02) & (sp < 0.03)] = -0.0
        return v

    def value_col(v, null_frac=0.0):
        mask = rng.random(n) < null_frac if null_frac > 0 else None
        return pa.array(v, mask=mask)

    vnull = spec.null_frac if spec.extra_nan_inf else 0.0
    if spec.dataset == "metrics":
        cols["rollup_sum"] = value_col(values(), vnull)
        cols["rollup_count"] = value_col(rng.integers(1, 20, n).astype(np.float64), vnull)
        cols["rollup_min"] = value_col(values(), vnull)
        cols["rollup_max"] = value_col(values(), vnull)
        cols[TELEMETRY_TYPE] = dict_col(["metrics"], np.zeros(n, np.int64), 0.0)
    else:
        cols[VALUE] = value_col(values(), vnull)
        cols[TELEMETRY_TYPE] = dict_col([spec.dataset], np.zeros(n, np.int64), 0.0)
    for c in spec.drop_columns:
        cols.pop(c, None)
    return pa.table(cols)


def write_segment(path: str, spec: SynthSpec, index: int) -> str:
    import pyarrow.parquet as pq

    os.makedirs(os.path.dirname(path), exist_ok=True)
    tbl = make_table(spec, index)
    pq.write_table(
        tbl,
        path,
        compression=spec.compression,
        use_dictionary=True,
        data_page_version="1.0",
        data_page_size=1 << 20,
        row_group_size=spec.row_group_size,
        write_statistics=True,
    )
    return path


def _write_one(args):
    path, spec, index = args
    if not os.path.exists(path):
        tmp = path + f".tmp{os.getpid()}"
        write_segment(tmp, spec, index)
        os.replace(tmp, path)
    return path


def write_dataset(root: str, spec: SynthSpec, n_segments: int, first_index: int = 0, workers: Optional[int] = None) -> List[str]:
    """Writes ``n_segments`` segments (skipping files that already exist); returns their paths."""
    jobs = [(segment_path(root, spec.dataset, segment_id_for(first_index + i)), spec, first_index + i) for i in range(n_segments)]
    workers = workers or min(len(jobs), os.cpu_count() or 1)
    if workers <= 1 or len(jobs) == 1:
        return [_write_one(j) for j in jobs]
    with ProcessPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(_write_one, jobs))


def segment_request(index: int, dataset: str, step: int, query_tags: Optional[dict] = None,
                    start_ts: int = T0, end_ts: int = T0 + HOUR_MS) -> dict:
    """A SegmentRequest JSON object (core/src/main/scala/com/cardinal/model/SegmentRequest.scala:84-98)."""
    return {
        "hour": HOUR,
        "dateInt": DATE_INT,
        "segmentId": segment_id_for(index),
        "sealedStatus": True,
        "dataset": dataset,
        "queryTags": query_tags or {},
        "stepInMillis": step,
        "customerId": CUSTOMER_ID,
        "collectorId": COLLECTOR_ID,
        "bucketName": "synthetic-bucket",
        "cName": "",
        "startTs": start_ts,
        "endTs": end_ts,
    }


def push_down_request(base_expr: dict, indices: Sequence[int], step: int, start_ts: int = T0,
                      end_ts: int = T0 + HOUR_MS) -> dict:
    """A PushDownRequest JSON object (SegmentRequest.scala:30-43)."""
    dataset = base_expr.get("dataset", "metrics")
    return {
        "baseExpr": base_expr,
        "segmentRequests": [segment_request(i, dataset, step, start_ts=start_ts, end_ts=end_ts) for i in indices],
        "reverseSort": False,
        "isTagQuery": False,
    }


# ---- the BASELINE.json configurations as concrete DataExprs ----
def c1_base_expr() -> dict:
    return {
        "id": "c1", "dataset": "logs",
        "filter": {"k": NAME, "v": ["metric_007"], "op": "eq", "dataType": "string", "extracted": False, "computed": False},
        "chart": {"aggregation": "sum", "groupBys": [], "type": "count"},
    }


def c2_base_expr(aggregation: str = "sum", rollup: str = "sum") -> dict:
    return {
        "id": "c2", "dataset": "metrics",
        "filter": {"k": TAG_SERVICE, "v": ["svc-03"], "op": "eq", "dataType": "string", "extracted": False, "computed": False},
        "chart": {"aggregation": aggregation, "rollup": rollup, "groupBys": list(GROUP_TAGS), "type": "count"},
    }


C2_AGGREGATES = (("sum", "sum"), ("sum", "count"), ("min", "min"), ("max", "max"))  # (aggregation, rollup)


def c4_base_expr(aggregation: str = "sum", rollup: str = "sum") -> dict:
    return {
        "id": "c4", "dataset": "metrics",
        "filter": {"k": TAG_POD, "v": ["^pod-[0-4].*"], "op": "regex", "dataType": "string", "extracted": False, "computed": False},
        "chart": {"aggregation": aggregation, "rollup": rollup, "groupBys": list(GROUP_TAGS), "type": "count"},
    }


def c4_spec(rows: int) -> SynthSpec:
    return SynthSpec(dataset="metrics", rows=rows, cards=(16, 100, 100, 100))

"""Host-side mirror of the reference's segment-evaluator interface, over the C ABI.

Names follow the reference: ``PushDownRequest`` / ``SegmentRequest`` (core/.../model/SegmentRequest.scala:62-98),
``evaluate_push_down_request`` = ``Commons.evaluatePushDownRequest`` (core/.../utils/Commons.scala:343-397),
``DataPoint`` (model/DataPoint.scala:19), the map-sketch ``SketchInput`` (PushDownAggregatorStage.scala:95-106).
Python is only the caller here (the reference's host is Scala; no JVM exists in this image): every row-level
operation happens inside liblakeside_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes
import json
import os
from dataclasses import dataclass
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import LakesideError, LakesideQueryError, LakesideUnsupported  # noqa: F401


def init(**options) -> None:
    _lib.check(_lib.load().lk_init(json.dumps(options).encode() if options else None))


def device_count() -> int:
    return int(_lib.load().lk_device_count())


def cache_stats() -> Dict[str, int]:
    """HBM-resident segment cache (lk_cache_stats): column chunks of segment files kept in device memory across queries."""
    out = (ctypes.c_int64 * 6)()
    _lib.check(_lib.load().lk_cache_stats(out))
    return dict(zip(["capacity_bytes", "resident_bytes", "segments", "column_hits", "column_misses", "evicted_segments"], [int(x) for x in out]))


def cache_configure(capacity_bytes: int) -> None:
    _lib.check(_lib.load().lk_cache_configure(int(capacity_bytes)))


def cache_clear() -> None:
    _lib.load().lk_cache_clear()


@dataclass
class DataPoint:
    timestamp: int
    value: float
    tags: Dict[str, str]


@dataclass
class SketchInput:
    timestamp: int
    tags: Dict[str, str]
    sketch: Dict[str, float]
    sketchType: str = "map"


class GlobResult:
    """Column view of an ``lk_result`` (what ``Commons.toDataPoint`` reads from the JDBC ResultSet)."""

    def __init__(self, handle: int):
        lib = _lib.load()
        self._h = ctypes.c_void_p(handle)
        n = self.num_rows = int(lib.lk_result_num_rows(self._h))
        self.num_values = int(lib.lk_result_num_values(self._h))
        self.num_tags = int(lib.lk_result_num_tags(self._h))
        self.columns = [lib.lk_result_col_name(self._h, i).decode() for i in range(lib.lk_result_num_cols(self._h))]

        def arr(ptr, dtype):
            # zero-copy views into the result's pinned host block: valid until close()
            return np.ctypeslib.as_array(ptr, (n,)).view(dtype) if n else np.zeros(0, dtype)

        self.ts = arr(lib.lk_result_ts(self._h), np.int64)
        self.values = [arr(lib.lk_result_value(self._h, a), np.float64) for a in range(self.num_values)]
        self.value_nulls = [arr(lib.lk_result_value_null(self._h, a), np.uint8) for a in range(self.num_values)]
        self.tag_codes = [arr(lib.lk_result_tag_codes(self._h, t), np.int32) for t in range(self.num_tags)]
        self.tag_dicts: List[List[str]] = []
        for t in range(self.num_tags):
            cnt = ctypes.c_int32()
            strs = ctypes.POINTER(ctypes.c_char_p)()
            _lib.check(lib.lk_result_tag_dict(self._h, t, ctypes.byref(cnt), ctypes.byref(strs)))
            self.tag_dicts.append([strs[i].decode("utf-8", "replace") for i in range(cnt.value)])

    # JDBC-style row access (1-based columns), used by the ResultSet shim test
    def get_long(self, row: int, col: int) -> int:
        return int(_lib.load().lk_result_get_long(self._h, row, col))

    def get_double(self, row: int, col: int) -> float:
        return float(_lib.load().lk_result_get_double(self._h, row, col))

    def get_string(self, row: int, col: int) -> Optional[str]:
        s = _lib.load().lk_result_get_string(self._h, row, col)
        return None if s is None else s.decode("utf-8", "replace")

    def close(self):
        if self._h:
            _lib.load().lk_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def to_sse(self, sketch_keys: Sequence[str], query_tags: Optional[Dict[str, str]] = None, row0: int = 0,
               row1: Optional[int] = None) -> bytes:
        """Rows [row0, row1) as the reference's per-segment stream elements (``Commons.dataPointResponseToSSE``,
        Commons.scala:474-502, over the map sketches of PushDownAggregatorStage.scala:95-106), serialised natively
        (``lk_result_to_sse``): one ``data: {...}\\r\\n\\r\\n`` event per row; value column v goes under sketch_keys[v]."""
        lib = _lib.load()
        row1 = self.num_rows if row1 is None else row1
        keys = (ctypes.c_char_p * len(sketch_keys))(*[k.encode() for k in sketch_keys])
        flat = [x.encode() for kv in (query_tags or {}).items() for x in kv]
        fb = (ctypes.c_char_p * max(1, len(flat)))(*flat) if flat else None
        need = int(lib.lk_result_to_sse(self._h, row0, row1, keys, len(sketch_keys), fb, len(flat) // 2, None, 0))
        if need < 0:
            _lib.check(_lib.LK_ERR_INVALID)
        buf = ctypes.create_string_buffer(max(1, need))
        got = int(lib.lk_result_to_sse(self._h, row0, row1, keys, len(sketch_keys), fb, len(flat) // 2, buf, need))
        assert got == need
        return buf.raw[:need]

    def tag_counts(self) -> Dict[Optional[str], int]:
        """Tag query result (BaseExpr.scala:127-143: SELECT "tag", COUNT(*) ... GROUP BY "tag"): {tag value (None = NULL group): rows}."""
        assert self.columns[1:] == ["count"], "not a tag-query result"
        d = self.tag_dicts[0]
        return {(None if c < 0 else d[c]): int(v) for c, v in zip(self.tag_codes[0].tolist(), self.values[0].tolist())}

    def to_data_points(self, query_tags: Optional[Dict[str, Any]] = None, value_index: int = 0) -> List[DataPoint]:
        """``Commons.toDataPoint`` aggregate branch (Commons.scala:424-461): null / "" / "null" tags are dropped and
        an empty tag map falls back to the segment's queryTags."""
        names = self.columns[1 + self.num_values:]
        tagcols = []
        for t in range(self.num_tags):
            d = np.array(self.tag_dicts[t] + [None], dtype=object)
            tagcols.append(d[np.where(self.tag_codes[t] < 0, len(self.tag_dicts[t]), self.tag_codes[t])])
        out = []
        vals = self.values[value_index]
        for i in range(self.num_rows):
            tags = {}
            for nme, col in zip(names, tagcols):
                v = col[i]
                if v is not None and v != "null" and v != "":
                    tags[nme] = v
            if not tags and query_tags:
                tags.update(query_tags)
            out.append(DataPoint(int(self.ts[i]), float(vals[i]), tags))
        return out


class Comm:
    """Communicator of the sharded record path (``lk_comm``): this rank's receive pools and its peers', one GPU per rank.
    ``handle()`` is all-gathered by the host (any channel), ``connect(handles)`` maps the peers' pools (CUDA IPC, or plain
    pointers inside one process).  With a Comm attached (``Query.set_comm``) survivor records travel to the rank that owns
    their (group x bucket) cell DURING the scan, over NVLink; ``finalize`` waits on the device for all sources."""

    def __init__(self, rank: int, world: int, pool_records: int, max_aggs: int = 4):
        self._h = ctypes.c_void_p()
        self.rank, self.world = rank, world
        _lib.check(_lib.load().lk_comm_create(rank, world, int(pool_records), max_aggs, ctypes.byref(self._h)))

    def handle(self) -> bytes:
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _lib.check(_lib.load().lk_comm_handle(self._h, ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p.value, n.value)

    def connect(self, handles: Sequence[bytes]):
        assert len(handles) == self.world and len({len(h) for h in handles}) == 1
        _lib.check(_lib.load().lk_comm_connect(self._h, b"".join(handles), len(handles[0])))

    def close(self):
        if self._h:
            _lib.load().lk_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Query:
    """Staged evaluation of one glob: create -> add segments -> prepare (HBM resident) -> execute -> finalize."""

    def __init__(self, push_down_request_json: str, aggregates: Optional[Sequence[Tuple[str, str]]] = None,
                 path: str = "auto", exact_sums: bool = False, seq_offset: int = 0):
        lib = _lib.load()
        opts: Dict[str, Any] = {"path": path, "exact_sums": bool(exact_sums)}
        if seq_offset:
            opts["seq_offset"] = int(seq_offset)  # exact_sums, sharded: global sequence number of this shard's first row
        if aggregates:
            opts["aggregates"] = [{"aggregation": a, "rollup": r} for a, r in aggregates]
        self._h = ctypes.c_void_p()
        self._keep: list = []
        _lib.check(lib.lk_query_create(push_down_request_json.encode(), json.dumps(opts).encode(), ctypes.byref(self._h)))

    def add_segment_file(self, path: str):
        _lib.check(_lib.load().lk_query_add_segment_file(self._h, path.encode()))

    def add_segment_buffer(self, ptr: int, length: int, keepalive=None):
        if keepalive is not None:
            self._keep.append(keepalive)
        _lib.check(_lib.load().lk_query_add_segment_buffer(self._h, ctypes.c_void_p(ptr), length))

    def add_segment_bytes(self, data: bytes):
        buf = ctypes.create_string_buffer(data, len(data))
        self.add_segment_buffer(ctypes.addressof(buf), len(data), keepalive=buf)

    def set_comm(self, comm: Optional["Comm"]):
        """Sharded evaluation: attach the communicator before the first execute (record path)."""
        _lib.check(_lib.load().lk_query_set_comm(self._h, comm._h if comm is not None else None))
        self._comm = comm  # keep it alive

    def plan(self):
        """Host half of prepare (no CUDA): usable on a CPU-only box for the sharding/dictionary logic."""
        _lib.check(_lib.load().lk_query_plan(self._h))

    def prepare(self):
        _lib.check(_lib.load().lk_query_prepare(self._h))
        self._keep.clear()

    def partial_sparse(self, nparts: int):
        """-> (device pointer, [entries per partition], stride bytes); the table is left empty."""
        ptr, stride = ctypes.c_void_p(), ctypes.c_int()
        counts = (ctypes.c_int64 * nparts)()
        _lib.check(_lib.load().lk_query_partial_sparse(self._h, nparts, ctypes.byref(ptr), counts, ctypes.byref(stride)))
        return ptr.value or 0, list(counts), stride.value

    def merge_sparse(self, device_ptr: int, n: int):
        _lib.check(_lib.load().lk_query_merge_sparse(self._h, ctypes.c_void_p(device_ptr), n))

    def execute(self):
        _lib.check(_lib.load().lk_query_execute(self._h))

    def phase(self) -> Tuple[int, int]:
        """(min, max) timestamp phase of this rank's last scan (lk_query_phase); min = 0xffffffff: no surviving row."""
        a, b = ctypes.c_uint32(), ctypes.c_uint32()
        _lib.check(_lib.load().lk_query_phase(self._h, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def set_phase(self, phase_min: int, phase_max: int):
        """The phase range reduced over all ranks (MIN, MAX), before finalizing reduced / foreign cells (lk_query_set_phase)."""
        _lib.check(_lib.load().lk_query_set_phase(self._h, int(phase_min), int(phase_max)))

    def sync(self):
        _lib.check(_lib.load().lk_query_sync(self._h))

    def finalize_device(self):
        _lib.check(_lib.load().lk_query_finalize_device(self._h))

    result_rows = 0  # rows of the last finalize()

    def finalize(self) -> GlobResult:
        r = ctypes.c_void_p()
        _lib.check(_lib.load().lk_query_finalize(self._h, ctypes.byref(r)))
        res = GlobResult(r.value)
        self.result_rows = res.num_rows
        return res

    def export_dictionaries(self) -> bytes:
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        _lib.check(_lib.load().lk_query_export_dictionaries(self._h, ctypes.byref(p), ctypes.byref(n)))
        return ctypes.string_at(p.value, n.value)

    def import_dictionaries(self, blob: bytes):
        _lib.check(_lib.load().lk_query_import_dictionaries(self._h, blob, len(blob)))

    def partial_dense(self):
        n, k = ctypes.c_int64(), ctypes.c_int()
        ptrs = (ctypes.c_void_p * 8)()
        ops = (ctypes.c_int * 8)()
        _lib.check(_lib.load().lk_query_partial_dense(self._h, ctypes.byref(n), ctypes.byref(k), ptrs, ops))
        return n.value, [(ptrs[i], ops[i]) for i in range(k.value)]

    @property
    def stream(self) -> int:
        s = ctypes.c_void_p()
        _lib.check(_lib.load().lk_query_stream(self._h, ctypes.byref(s)))
        return s.value or 0

    @property
    def info(self) -> dict:
        s = ctypes.c_char_p()
        _lib.check(_lib.load().lk_query_info_json(self._h, ctypes.byref(s)))
        return json.loads(s.value.decode())

    @property
    def timings(self) -> Dict[str, float]:
        ms = (ctypes.c_double * 8)()
        _lib.check(_lib.load().lk_query_timings(self._h, ms))
        return {"h2d_ms": ms[0], "scan_ms": ms[1], "finalize_ms": ms[2], "d2h_ms": ms[3], "plan_ms": ms[4], "def_expand_ms": ms[5], "exchange_wait_ms": ms[6]}

    @property
    def touched_bytes(self) -> int:
        return int(_lib.load().lk_query_touched_bytes(self._h))

    @property
    def total_rows(self) -> int:
        return int(_lib.load().lk_query_total_rows(self._h))

    @property
    def survivors(self) -> int:
        return int(_lib.load().lk_query_survivors(self._h))

    def eval(self, n_rows: int, aggregation: str, chart_type: str = "line", metric_type: str = "gauge") -> np.ndarray:
        """BaseExpr.eval on the reduced rows of the last finalize (BaseExpr.scala:665-695, :47-95; ASTUtils.scala:190-219):
        one value per result row -- the map-sketch entry `aggregation` (`avg` = sum / count), then the chart/metric-type
        transform.  Computed on the device from the result columns still in HBM."""
        out = np.empty(int(n_rows), dtype=np.float64)
        n = int(_lib.load().lk_query_eval(self._h, aggregation.encode(), chart_type.encode(), metric_type.encode(),
                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), int(n_rows)))
        if n < 0:
            _lib.check(_lib.LK_ERR_INVALID)  # the message comes from lk_last_error
        return out[:n]

    def eval_results(self, res: "GlobResult", aggregation: str, chart_type: str, metric_type: str, group_bys: Sequence[str],
                     query_tags: Optional[Dict[str, Any]] = None) -> List[Tuple[int, float, Dict[str, str], str]]:
        """``BaseExpr.eval`` per reduced row: (timestamp, value, tags, group key).  The value comes from the device
        (``lk_query_eval``); the group key is ``sorted(groupBys).map(tags.getOrElse(_, "")).mkString(":")``
        (ASTUtils.scala:87-89), "default" without group-bys (BaseExpr.scala:665-695)."""
        vals = self.eval(res.num_rows, aggregation, chart_type, metric_type)
        keys = sorted(set(group_bys))
        out = []
        for d, v in zip(res.to_data_points(query_tags), vals):
            out.append((d.timestamp, float(v), d.tags, ":".join(str(d.tags.get(k, "")) for k in keys) if keys else "default"))
        return out

    def close(self):
        if self._h:
            _lib.load().lk_query_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def formula_eval(op: str, e1, e2, cap: Optional[int] = None):
    """``Formula.eval`` (Formula.scala:32-69) on the device over the reduced rows of two finalized queries.
    ``e1`` / ``e2``: ``(Query, spec dict)`` for a BaseExpr side (spec: aggregation, chartType, metricType, groupBys) or a
    number for a ConstantExpr.  Returns (timestamps, values, side, row): ``side[i]`` says whose tags result i carries (1 = the
    e1 query's result row ``row[i]``, 2 = the e2 query's)."""
    def side(x):
        if isinstance(x, (int, float)):
            return None, {"constant": float(x)}
        return x[0], dict(x[1])

    q1, s1 = side(e1)
    q2, s2 = side(e2)
    spec = json.dumps({"op": op, "e1": s1, "e2": s2}).encode()
    if cap is None:
        cap = 1
        for q in (q1, q2):
            if q is not None:
                cap += q.result_rows
    ts, val = np.empty(cap, np.int64), np.empty(cap, np.float64)
    sd, row = np.empty(cap, np.int32), np.empty(cap, np.int64)
    n_out = ctypes.c_int64()
    _lib.check(_lib.load().lk_formula_eval(q1._h if q1 else None, q2._h if q2 else None, spec, cap,
                                           ts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), val.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                           sd.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), row.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                           ctypes.byref(n_out)))
    n = int(n_out.value)
    return ts[:n], val[:n], sd[:n], row[:n]


def eval_glob(push_down_request_json: str, parquet_paths: Sequence[str]) -> GlobResult:
    """``lk_eval``: drop-in for ``Commons.toGlobResultSet`` (Commons.scala:200-254)."""
    lib = _lib.load()
    arr = (ctypes.c_char_p * len(parquet_paths))(*[p.encode() for p in parquet_paths])
    r = ctypes.c_void_p()
    _lib.check(lib.lk_eval(push_down_request_json.encode(), arr, len(parquet_paths), ctypes.byref(r)))
    return GlobResult(r.value)


def to_parquet_file_path(sr: dict, db_root: str = "./db") -> str:
    """``Commons.toParquetFilePath`` for local segments (Commons.scala:160-177, 256-278)."""
    return f"{db_root}/{sr['customerId']}/{sr['collectorId']}/{sr['dateInt']}/{sr['dataset']}/{sr['hour']}/{sr['segmentId']}.parquet"


def merge_sorted_source(sources: List[list], reverse_sort: bool = False) -> list:
    """``mergeSortedSource`` (QueryEngineV2.scala:76-97): K-way merge by timestamp on the GPU (lk_merge_streams).
    Elements need a ``timestamp`` attribute; ties come out by source index descending (left-deep mergeSorted fold)."""
    lens = [len(s) for s in sources]
    total = sum(lens)
    if total == 0:
        return []
    ts = [np.fromiter((e.timestamp for e in s), np.int64, len(s)) for s in sources]
    src, pos = merge_streams_index(ts, reverse_sort)
    return [sources[s][p] for s, p in zip(src.tolist(), pos.tolist())]


def merge_streams_index(ts_list: Sequence[np.ndarray], reverse: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Merges K sorted int64 timestamp streams; returns (source index, position) of every output element."""
    lib = _lib.load()
    k = len(ts_list)
    ts_list = [np.ascontiguousarray(t, np.int64) for t in ts_list]
    gids = [np.arange(len(t), dtype=np.int32) for t in ts_list]  # carry the in-stream position as payload
    vals = [np.zeros(len(t), np.float64) for t in ts_list]
    lens = (ctypes.c_int64 * k)(*[len(t) for t in ts_list])
    total = int(sum(len(t) for t in ts_list))
    P = ctypes.c_void_p
    tsp = (P * k)(*[t.ctypes.data for t in ts_list])
    gp = (P * k)(*[g.ctypes.data for g in gids])
    vp = (P * k)(*[v.ctypes.data for v in vals])
    out_ts = np.empty(total, np.int64)
    out_gid = np.empty(total, np.int32)
    out_val = np.empty(total, np.float64)
    out_src = np.empty(total, np.int32)
    _lib.check(lib.lk_merge_streams(k, tsp, gp, vp, lens, 1 if reverse else 0, out_ts.ctypes.data, out_gid.ctypes.data,
                                    out_val.ctypes.data, out_src.ctypes.data))
    return out_src, out_gid


def parse_dictionary_blob(blob: bytes) -> List[List[bytes]]:
    """Blob of lk_query_export_dictionaries: u32 n_keys; per key: u32 n; per string: u32 len, bytes."""
    import struct

    p = 0
    (nk,) = struct.unpack_from("<I", blob, p)
    p += 4
    out = []
    for _ in range(nk):
        (n,) = struct.unpack_from("<I", blob, p)
        p += 4
        d = []
        for _ in range(n):
            (ln,) = struct.unpack_from("<I", blob, p)
            p += 4
            d.append(bytes(blob[p:p + ln]))
            p += ln
        out.append(d)
    return out


def union_dictionaries(blobs: Sequence[bytes]) -> bytes:
    """Union (byte-wise sorted) of the per-rank group-by dictionaries: the one global code space all ranks import."""
    import struct

    parsed = [parse_dictionary_blob(b) for b in blobs]
    nk = len(parsed[0])
    assert all(len(p) == nk for p in parsed), "ranks disagree on the number of key columns"
    out = [struct.pack("<I", nk)]
    for k in range(nk):
        vals = sorted(set().union(*[set(p[k]) for p in parsed]))
        out.append(struct.pack("<I", len(vals)))
        for v in vals:
            out.append(struct.pack("<I", len(v)))
            out.append(v)
    return b"".join(out)


def shard_request(push_down_request: dict, rank: int, world: int) -> Tuple[dict, List[int]]:
    """Segment sharding of a PushDownRequest over `world` GPUs (the reference shards segments over workers by
    ``floorMod(segmentId.hashCode, n)``, WorkerManager.scala:150-156; here: round-robin by position).  Every shard keeps the
    request's global [startTs, endTs) so that all ranks index the same buckets."""
    srs = push_down_request["segmentRequests"]
    lo = min(s["startTs"] for s in srs)
    hi = max(s["endTs"] for s in srs)
    idx = [i for i in range(len(srs)) if i % world == rank]
    mine = [dict(srs[i], startTs=lo, endTs=hi, stepInMillis=srs[0]["stepInMillis"]) for i in idx]
    return dict(push_down_request, segmentRequests=mine), idx


def evaluate_push_down_request(query_id: str, local_parquet: bool, push_down_request: dict, db_root: str = "./db"):
    """``Commons.evaluatePushDownRequest`` (Commons.scala:343-397): globs of 10 (local) / 5 (S3) segments, one GPU
    evaluation per glob, glob streams merged by timestamp; an empty request yields the ts = -1 sentinel; a failing
    glob streams nothing (Commons.scala:249-253)."""
    srs = push_down_request["segmentRequests"]
    if not srs:
        return [DataPoint(-1, -1.0, {})]
    glob_size = 10 if local_parquet else 5
    agg = push_down_request["baseExpr"].get("chart", {}).get("aggregation", "sum")
    sources = []
    for g in range(0, len(srs), glob_size):
        group = srs[g:g + glob_size]
        sub = dict(push_down_request, segmentRequests=group)
        paths = [to_parquet_file_path(s, db_root) for s in group]
        try:
            res = eval_glob(json.dumps(sub), paths)
        except (LakesideQueryError, LakesideUnsupported):
            sources.append([])
            continue
        except LakesideError as e:
            if e.code == _lib.LK_ERR_IO:
                sources.append([])
                continue
            raise
        dps = res.to_data_points(group[0].get("queryTags") or {})
        sources.append([SketchInput(d.timestamp, d.tags, {agg: d.value}) for d in dps])
    if len(sources) == 1:
        return sources[0]
    return merge_sorted_source(sources, False)


def merge_and_reduce(ts_list: Sequence[np.ndarray], gid_list: Sequence[np.ndarray], val_list: Sequence[np.ndarray],
                     aggregation: str = "sum", reverse: bool = False):
    """K-way merge of the per-segment streams followed by the map-sketch merge of TimeGroupedSketchAggregator
    (core/.../eval/TimeGroupedSketchAggregator.scala:63-93): equal (timestamp, group) elements are folded in merged
    (arrival) order with + (sum, count), Math.min or Math.max.  Returns (ts, gid, value) with one element per
    (timestamp, group), sorted by timestamp then group."""
    lib = _lib.load()
    k = len(ts_list)
    ts_list = [np.ascontiguousarray(t, np.int64) for t in ts_list]
    gid_list = [np.ascontiguousarray(g, np.int32) for g in gid_list]
    val_list = [np.ascontiguousarray(v, np.float64) for v in val_list]
    lens = (ctypes.c_int64 * k)(*[len(t) for t in ts_list])
    total = int(sum(len(t) for t in ts_list))
    P = ctypes.c_void_p
    h = ctypes.c_void_p()
    _lib.check(lib.lk_merge_create(k, (P * k)(*[t.ctypes.data for t in ts_list]), (P * k)(*[g.ctypes.data for g in gid_list]),
                                   (P * k)(*[v.ctypes.data for v in val_list]), lens, 1 if reverse else 0, ctypes.byref(h)))
    try:
        _lib.check(lib.lk_merge_run(h))
        op = {"sum": 0, "count": 1, "min": 2, "max": 3}[aggregation]
        o_ts, o_gid, o_val = np.empty(total, np.int64), np.empty(total, np.int32), np.empty(total, np.float64)
        n = ctypes.c_int64()
        _lib.check(lib.lk_merge_reduce(h, op, ctypes.byref(n), o_ts.ctypes.data, o_gid.ctypes.data, o_val.ctypes.data))
        return o_ts[:n.value].copy(), o_gid[:n.value].copy(), o_val[:n.value].copy()
    finally:
        lib.lk_merge_destroy(h)

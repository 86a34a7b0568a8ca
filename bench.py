#!/usr/bin/env python
"""bench.py -- scan-filter-aggregate throughput of the per-segment DataExpr path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic sealed segments.  Workload at N = 1 is
BASELINE.json configs[1] (C2): 100 metric segments x 1 Mi rows, filter `resource.service.name eq svc-03`, group by
3 tags (+ name, as the reference's SQL always does), count/sum/min/max = sum(rollup_sum), sum(rollup_count),
min(rollup_min), max(rollup_max) fused in one pass, 10 s step.  At N > 1 every rank holds its own 100 segments (weak
scaling, different seeds), evaluates them with no data-path collective and the partial aggregates meet in one NCCL
exchange (dense tables: reduce; sparse tables: gather of the occupied cells) -- SURVEY.md §8e.

value : rows/s, inputs (encoded Parquet column chunks + seek index) already resident in HBM; the timed region is
        scan kernel + on-device compaction of the aggregate table into result rows.
e2e   : the same metric through the C ABI with HOST buffers: footer/page/run indexing on the host, H2D of the touched
        column chunks from pinned memory, kernels, D2H of the result rows -- every step.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

METRIC = "rows/sec scan-filter-agg (C2: tag-equality filter + group-by 3 tags, count/sum/min/max, 10s step)"
UNIT = "rows/s"
N_SEGMENTS = int(os.environ.get("LK_BENCH_SEGMENTS", "100"))
ROWS = int(os.environ.get("LK_BENCH_ROWS", str(1 << 20)))
DATA_ROOT = os.environ.get("LK_BENCH_DATA", "/tmp/lakeside_b200_bench")
STEP_MS = 10000


def workload_name(n_gpus: int) -> str:
    return (f"C2: {N_SEGMENTS} synthetic metric segments x {ROWS} rows per GPU ({N_SEGMENTS * ROWS / 1e6:.1f}M rows/GPU), "
            "filter resource.service.name eq svc-03, group by name + 3 tags, sum(rollup_sum)/sum(rollup_count)/min(rollup_min)/max(rollup_max), "
            "step 10 s")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        """Median SM clock / throttle reasons of the samples that arrived inside [t0, t1] (the timed region).  The sampler
        is started before the warm-up so that nvidia-smi is already streaming; a timed region shorter than the 20 ms
        sampling period can still miss every sample: then all samples taken under the same load (warm-up .. per-kernel
        loop) are used and `window` says so."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.03]
        window = "timed region"
        if not inside:
            inside = [ln for (_, ln) in self.lines]
            window = "warm-up + timed region + per-kernel loop (timed region shorter than the sampling period)"
        sm, mx, reasons = [], [], set()
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def gen_dataset(rank: int):
    from lakeside_b200 import synth

    spec = synth.SynthSpec(dataset="metrics", rows=ROWS)
    first = rank * N_SEGMENTS
    root = os.path.join(DATA_ROOT, f"c2_{ROWS}")
    paths = synth.write_dataset(root, spec, N_SEGMENTS, first_index=first)
    rq = json.dumps(synth.push_down_request(synth.c2_base_expr(), list(range(first, first + N_SEGMENTS)), STEP_MS))
    return paths, rq, synth.C2_AGGREGATES


def run_reference(args):
    """--impl reference: the reference's evaluator is Scala + DuckDB 1.3.2 over JDBC; neither a JVM nor DuckDB exists in this
    image, so this arm times the oracle port of the same path (Arrow C++ Parquet decode + NumPy
    filter / group-by, all host threads Arrow wants) on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import lakeside_oracle as lo
    from lakeside_b200 import synth

    sample = max(1, min(N_SEGMENTS, int(os.environ.get("LK_BENCH_REF_SEGMENTS", "4"))))
    spec = synth.SynthSpec(dataset="metrics", rows=ROWS)
    paths = synth.write_dataset(os.path.join(DATA_ROOT, f"c2_{ROWS}"), spec, sample)
    rq = lo.push_down_request_from_json(json.dumps(synth.push_down_request(synth.c2_base_expr(), list(range(sample)), STEP_MS)))
    aggs = [(a, "rollup_" + r) for a, r in synth.C2_AGGREGATES]
    for _ in range(max(1, args.warmup // 3)):
        lo.evaluate_glob(rq, paths, aggs=aggs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lo.evaluate_glob(rq, paths, aggs=aggs)
    dt = (time.perf_counter() - t0) / args.steps
    rows = sample * ROWS
    import pyarrow as pa

    cores = pa.cpu_count()
    v = rows / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(1), "sample": f"{sample} of {N_SEGMENTS} segments per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} segments x {ROWS} rows, oracle port (Arrow C++ decode + NumPy), page cache warm"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(paths, rq_json, aggs):
    sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    import lakeside_oracle as lo
    import pyarrow as pa

    sample = max(1, min(len(paths), int(os.environ.get("LK_BENCH_REF_SEGMENTS", "4"))))
    rq = lo.push_down_request_from_json(rq_json)
    rq.segmentRequests = rq.segmentRequests[:sample]
    ag = [(a, "rollup_" + r) for a, r in aggs]
    lo.evaluate_glob(rq, paths[:sample], aggs=ag)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        lo.evaluate_glob(rq, paths[:sample], aggs=ag)
        best = min(best, time.perf_counter() - t0)
    return {"value": sample * ROWS / best, "unit": UNIT, "cores": pa.cpu_count(), "kind": "port",
            "sample": f"{sample} of {len(paths)} segments x {ROWS} rows; oracle port (Arrow C++ Parquet decode + NumPy filter/group-by), "
                      "page cache warm, best of 2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from lakeside_b200 import _lib, api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lakeside_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    api.init(device=local_rank)
    lib = _lib.load()

    paths, rq, aggs = gen_dataset(rank)
    # segment bytes in pinned host memory (what a worker's segment cache would hold)
    host = []
    for p in paths:
        n = os.path.getsize(p)
        ptr = lib.lk_host_alloc(n)
        if not ptr:
            raise SystemExit("lk_host_alloc failed: " + lib.lk_last_error().decode())
        with open(p, "rb") as f:
            f.readinto((ctypes.c_char * n).from_address(ptr))
        host.append((ptr, n))

    # the planner chooses (records for this selective, high-cardinality query); sharded, the appended records (or the hash
    # table's occupied cells) are what the sparse exchange partitions; LK_BENCH_PATH overrides (diagnosis)
    table_path = os.environ.get("LK_BENCH_PATH", "auto")

    def new_query():
        q = api.Query(rq, aggregates=aggs, path=table_path)
        for ptr, n in host:
            q.add_segment_buffer(ptr, n)
        return q

    def _as_tensor(ptr, n, typestr, dtype):
        class _A:
            pass
        a = _A()
        a.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}
        return torch.as_tensor(a, device=f"cuda:{local_rank}").view(dtype)

    def agree_on_dictionaries(qq):
        """All ranks must index ONE (group x bucket) space: all-gather the per-rank group-by dictionaries (host strings,
        a few KB), import their union everywhere."""
        if world == 1:
            return
        blobs = [None] * world
        dist.all_gather_object(blobs, qq.export_dictionaries())
        qq.import_dictionaries(api.union_dictionaries(blobs))

    SIGN = torch.iinfo(torch.int64).min
    exchange_bytes = [0]

    trace = [] if os.environ.get("LK_BENCH_TRACE") else None  # per-phase wall clock of the exchange (adds syncs: diagnosis only)

    def exchange(qq, path):
        """Host-mediated exchange of the paths that still need one (the record path -- what the planner picks for this
        workload -- exchanges INSIDE the scan through the lk_comm attached to the query: nothing to do here).
        dense : NCCL reduce of every (group x bucket) plane to rank 0 (sum f64 / sum u64 / max on order-preserving keys)
        hash  : cells are hash-partitioned over the ranks; NCCL all-to-all of the occupied 32/64-byte entries, each rank
                merges and finalises its own partition."""
        if world == 1 or path == "records":
            return
        if path == "dense":
            n_cells, planes = qq.partial_dense()
            qq.sync()
            for ptr, op in planes:
                if op == 0:
                    dist.reduce(_as_tensor(ptr, n_cells, "<f8", torch.float64), 0, op=dist.ReduceOp.SUM)
                elif op == 1:
                    dist.reduce(_as_tensor(ptr, n_cells, "<i8", torch.int64), 0, op=dist.ReduceOp.SUM)
                else:  # unsigned max: flip the sign bit so that NCCL's signed max orders the keys, flip back
                    t = _as_tensor(ptr, n_cells, "<i8", torch.int64)
                    t.bitwise_xor_(SIGN)
                    dist.reduce(t, 0, op=dist.ReduceOp.MAX)
                    t.bitwise_xor_(SIGN)
            exchange_bytes[0] = n_cells * 8 * len(planes)
            torch.cuda.synchronize()
        else:
            ptr, counts, stride = qq.partial_sparse(world)
            send_counts = torch.tensor(counts, dtype=torch.int64, device="cuda")
            recv_counts = torch.empty_like(send_counts)
            dist.all_to_all_single(recv_counts, send_counts)
            rc = recv_counts.tolist()
            n_send = sum(counts)
            send = _as_tensor(ptr, n_send * stride, "|u1", torch.uint8) if n_send else torch.empty(0, dtype=torch.uint8, device="cuda")
            recv = torch.empty(sum(rc) * stride, dtype=torch.uint8, device="cuda")
            dist.all_to_all_single(recv, send, [c * stride for c in rc], [c * stride for c in counts])
            torch.cuda.synchronize()
            qq.merge_sparse(recv.data_ptr(), sum(rc))
            qq._keep.append(recv)
            exchange_bytes[0] = (n_send - counts[rank]) * stride

    # the communicator of the record path: every rank's receive pools, mapped by its peers (CUDA IPC); the handles travel
    # over torch.distributed once, at setup -- the data path never touches NCCL or the host
    comm = None
    if world > 1 and table_path in ("auto", "records"):
        pool_records = int(os.environ.get("LK_BENCH_POOL_RECORDS", str(max(1 << 20, N_SEGMENTS * ROWS // 8))))
        comm = api.Comm(rank, world, pool_records, max_aggs=len(aggs))
        handles = [None] * world
        dist.all_gather_object(handles, comm.handle())
        comm.connect(handles)
        dist.barrier()

    # ---------------- resident ("kernel-only") arm ----------------
    sampler = ClockSampler(local_rank)  # started here: nvidia-smi needs a few 100 ms before its first sample
    sampler.start()
    q = new_query()
    if world > 1:
        q.plan()
        agree_on_dictionaries(q)
    if comm is not None:
        q.set_comm(comm)
    q.prepare()
    info = q.info
    if world > 1:  # every rank must have chosen the same aggregate layout: the exchange depends on it
        paths_all = [None] * world
        dist.all_gather_object(paths_all, info["path"])
        assert len(set(paths_all)) == 1, f"ranks disagree on the aggregate layout: {paths_all}"
        assert comm is None or info["path"] == "records", info["path"]
    rows_per_rank = q.total_rows
    touched = q.touched_bytes

    def step():
        q.execute()
        exchange(q, info["path"])
        q.finalize_device()
        q._keep.clear()

    for _ in range(max(3, args.warmup)):
        step()
    q.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ext = torch.cuda.ExternalStream(q.stream, device=torch.device("cuda", local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_region0 = time.perf_counter()
    e0.record(ext)
    for _ in range(args.steps):
        step()
    e1.record(ext)
    q.sync()
    torch.cuda.synchronize()
    t_region1 = time.perf_counter()
    dev_ms = e0.elapsed_time(e1)
    # per-kernel duration of the dominant kernel (CUDA events recorded by the library around each scan launch, on its
    # launching stream); measured in a separate loop so that reading them never serialises the timed region
    per_scan, per_defx = [], []
    for _ in range(args.steps):
        q.execute()
        exchange(q, info["path"])
        q.finalize_device()
        tm = q.timings
        per_scan.append(tm["scan_ms"])
        per_defx.append(tm["def_expand_ms"])
    clocks = sampler.stop(t_region0, t_region1)
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = rows_per_rank * world / (ms_per_step / 1e3)
    survivors = q.survivors
    result_rows = None

    # ---------------- end-to-end arm: host buffers -> result rows on the host, every step ----------------
    q.close()
    e2e_steps = max(2, min(args.steps, 5))

    e2e_trace = []

    def e2e_step():
        t = [time.perf_counter()]
        qq = new_query()
        t.append(time.perf_counter())
        if world > 1:  # the column chunks start moving during plan(); they overlap the index build and the agreement
            qq.plan()
            agree_on_dictionaries(qq)
        if comm is not None:
            qq.set_comm(comm)
        qq.prepare()
        t.append(time.perf_counter())
        qq.execute()
        exchange(qq, info["path"])
        res = qq.finalize()
        t.append(time.perf_counter())
        n = res.num_rows
        d2h = n * (8 + 8 * res.num_values + 4 * res.num_tags + res.num_values)
        h2d = qq.touched_bytes
        res.close()
        t.append(time.perf_counter())
        qq.close()
        t.append(time.perf_counter())
        e2e_trace.append([round((b - a) * 1e3, 2) for a, b in zip(t, t[1:])])
        return n, h2d, d2h

    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    import gc
    gc.collect()
    gc.disable()  # as timeit does: a generational collection over torch's module graph costs milliseconds at a random point
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        result_rows, h2d_b, d2h_b = e2e_step()
    torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    gc.enable()
    if world > 1:
        t = torch.tensor([e2e_dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = rows_per_rank * world / e2e_dt
    if trace is not None and rank == 0:
        print("e2e trace (ms) [create, plan+prepare, execute+exchange+finalize, result close, query close]:", e2e_trace[-e2e_steps:], file=sys.stderr)

    if rank == 0:
        peak, peak_src = measured_peaks()
        k_ms = sum(per_scan) / len(per_scan)
        achieved = touched / (k_ms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("segments") == N_SEGMENTS and tj.get("rows") == ROWS:
                traffic = tj.get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(world), "segments_per_gpu": N_SEGMENTS, "rows_per_segment": ROWS,
                       "aggregate_table": info["path"], "n_groups": info["n_groups"], "n_buckets": info["n_buckets"],
                       "survivor_rows_per_gpu": survivors, "result_rows": result_rows,
                       "l2": f"inputs ({touched / 1e9:.2f} GB of encoded column chunks per GPU) are larger than the 126 MB L2; no explicit flush",
                       "timed_region": "definition-level expansion + scan kernel + on-device aggregation/compaction of the result rows (inputs HBM-resident)",
                       "def_expand_ms": sum(per_defx) / len(per_defx),
                       "GBps_algorithmic": touched * world / (ms_per_step / 1e3) / 1e9,
                       "GBps_logical_36B_per_row": 36.0 * rows_per_rank * world / (ms_per_step / 1e3) / 1e9},
            "roofline": {"bound": "hbm", "kernel": "lk::scan_kernel (fused decode+filter+bucket+group-by aggregate)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": traffic,
                         "algorithmic_bytes_per_launch": touched, "kernel_ms": k_ms,
                         "note": "algorithmic bytes = sum of ColumnMetaData.total_compressed_size of the touched column chunks (SURVEY §8d)"},
            "cpu_baseline": cpu_baseline(paths, rq, aggs) if world == 1 else None,  # timed on rank 0 at N = 1 only
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_b), "d2h_bytes_per_step": int(d2h_b),
                    "ms_per_step": e2e_dt * 1e3, "steps": e2e_steps,
                    "what": "lk_query_create + add_segment_buffer(pinned host bytes) + prepare (host index + H2D) + execute + finalize (D2H)"},
            # this library's own kernels per step (CUB's radix-sort kernels of the record path are library code and not counted):
            # def_expand (when a touched column has NULLs) +
            # records: scan, rec_count, exclusive_scan, rec_emit; hash: scan, hist, exclusive_scan, scatter, emit
            # (+ sparse_hist, sparse_scatter, sparse_merge / rec_part_hist, rec_part_scatter, rec_unpack when sharded); dense: scan, count, exclusive_scan, emit
            # records: def_expand, scan, rec_bhist, rec_regions, rec_group, rec_rowscan, rec_emit (+ comm_begin, comm_seal_publish, comm_wait
            # when sharded); hash: scan, hist, exclusive_scan, scatter, emit (+ sparse_hist, sparse_scatter, sparse_merge); dense: scan, count, exclusive_scan, emit
            "gpu_launches": args.steps * ((1 if info.get("def_chunks", 1) else 0) + {"records": 6, "hash": 5, "dense": 4}[info["path"]] + (3 if world > 1 and info["path"] != "dense" else 0)),
            "exchange": None if world == 1 else {
                "kind": ("NCCL reduce of dense planes" if info["path"] == "dense" else
                         "survivor records stored into the owner rank's receive pool over NVLink during the scan (lk_comm: CUDA IPC peer pools, "
                         "one remote atomic per 256 records, device-side completion flags; no NCCL call or host round trip on the data path)"
                         if info["path"] == "records" else "NCCL all-to-all of hash-partitioned occupied cells"),
                "bytes_sent_per_rank_per_step": (int(survivors * (world - 1) / world) * 8 * (1 + len(aggs)) if info["path"] == "records" else exchange_bytes[0])},
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

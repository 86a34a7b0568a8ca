"""Wall-clock breakdown of the end-to-end arm of bench.py (not the benchmark): pinned host buffers -> create -> prepare
(host index + H2D) -> execute -> finalize (D2H) -> close, per stage, plus the library's own event timings."""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

from lakeside_b200 import _lib, api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=100)
ap.add_argument("--rows", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--data", default="/tmp/lk_probe")
ap.add_argument("--files", action="store_true", help="segment files through the HBM-resident segment cache instead of caller buffers")
ap.add_argument("--torch", action="store_true", help="initialise torch.cuda first, as bench.py does")
a = ap.parse_args()

if a.torch:
    import torch

    torch.cuda.init()
    torch.zeros(1, device="cuda")
api.init(**json.loads(os.environ.get("LK_INIT", "{}")))
lib = _lib.load()
spec = synth.SynthSpec(dataset="metrics", rows=a.rows)
paths = synth.write_dataset(os.path.join(a.data, f"c2_{a.rows}"), spec, a.segments)
rq = json.dumps(synth.push_down_request(synth.c2_base_expr(), list(range(a.segments)), 10000))
bufs = []
for p in paths:  # segment bytes in pinned host memory, as the worker's segment cache would hold them
    data = open(p, "rb").read()
    ptr = lib.lk_host_alloc(len(data))
    ctypes.memmove(ptr, data, len(data))
    bufs.append((ptr, len(data)))
out = []
for _ in range(a.steps):
    t = [time.perf_counter()]
    q = api.Query(rq, aggregates=synth.C2_AGGREGATES)
    if a.files:
        for p in paths:
            q.add_segment_file(p)
    else:
        for ptr, n in bufs:
            q.add_segment_buffer(ptr, n)
    t.append(time.perf_counter())
    q.prepare(); t.append(time.perf_counter())
    q.execute(); q.sync(); t.append(time.perf_counter())
    res = q.finalize(); t.append(time.perf_counter())
    tm = q.timings
    n_rows = res.num_rows
    res.close(); t.append(time.perf_counter())
    q.close(); t.append(time.perf_counter())
    row = {k: round((y - x) * 1e3, 2) for k, x, y in zip(["create", "prepare", "execute", "finalize", "res_close", "q_close"], t, t[1:])}
    row["total"] = round((t[-1] - t[0]) * 1e3, 2)
    row["lib"] = {k: round(v, 2) for k, v in tm.items()}
    row["rows"] = n_rows
    out.append(row)
print(json.dumps(out))

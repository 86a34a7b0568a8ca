"""Single-GPU timing of the sparse exchange's device side (not the benchmark): execute -> partition into `nparts`
hash partitions -> merge every partition back -> finalize.  Wall-clock per stage with a sync on both sides; run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel times."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]

from lakeside_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=100)
ap.add_argument("--rows", type=int, default=1 << 20)
ap.add_argument("--nparts", type=int, default=2)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--data", default="/tmp/lk_probe")
a = ap.parse_args()

api.init(**json.loads(os.environ.get("LK_INIT", "{}")))
spec = synth.SynthSpec(dataset="metrics", rows=a.rows)
paths = synth.write_dataset(os.path.join(a.data, f"c2_{a.rows}"), spec, a.segments)
rq = json.dumps(synth.push_down_request(synth.c2_base_expr(), list(range(a.segments)), 10000))
q = api.Query(rq, aggregates=synth.C2_AGGREGATES, path="hash")
for p in paths:
    q.add_segment_file(p)
q.prepare()
out = []
for _ in range(a.steps):
    t = [time.perf_counter()]
    q.execute(); q.sync(); t.append(time.perf_counter())
    ptr, counts, stride = q.partial_sparse(a.nparts); q.sync(); t.append(time.perf_counter())
    q.merge_sparse(ptr, sum(counts)); q.sync(); t.append(time.perf_counter())
    q.finalize_device(); q.sync(); t.append(time.perf_counter())
    out.append({k: round((y - x) * 1e3, 3) for k, x, y in zip(["execute", "partition", "merge", "finalize"], t, t[1:])})
print(json.dumps({"nparts": a.nparts, "entries": sum(counts), "stride": stride, "passes_ms": out}))

"""Quick device-side timing probe (not the benchmark): N synthetic segments resident in HBM, K timed passes."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

from lakeside_b200 import api, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--segments", type=int, default=8)
ap.add_argument("--rows", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--path", default="auto")
ap.add_argument("--config", default="c2")
ap.add_argument("--data", default="/tmp/lk_probe")
a = ap.parse_args()

api.init(**json.loads(os.environ.get("LK_INIT", "{}")))
if a.config == "c1":
    spec = synth.SynthSpec(dataset="logs", rows=a.rows)
    be, step, aggs = synth.c1_base_expr(), 60000, None
elif a.config == "c2name":
    spec = synth.SynthSpec(dataset="metrics", rows=a.rows)
    be = synth.c2_base_expr()
    be["filter"] = {"q1": be["filter"], "q2": {"k": synth.NAME, "v": ["metric_000"], "op": "eq"}, "op": "and"}
    step, aggs = 10000, synth.C2_AGGREGATES
else:
    spec = synth.SynthSpec(dataset="metrics", rows=a.rows)
    be, step, aggs = synth.c2_base_expr(), 10000, synth.C2_AGGREGATES
t0 = time.time()
paths = synth.write_dataset(os.path.join(a.data, f"{a.config}_{a.rows}"), spec, a.segments)
t_gen = time.time() - t0
rq = json.dumps(synth.push_down_request(be, list(range(a.segments)), step))
q = api.Query(rq, aggregates=aggs, path=a.path)
t0 = time.time()
for p in paths:
    q.add_segment_file(p)
t_read = time.time() - t0
t0 = time.time()
q.prepare()
t_prep = time.time() - t0
out = {"gen_s": t_gen, "read_s": t_read, "prepare_s": t_prep, "info": q.info, "passes": []}
for i in range(a.steps):
    t0 = time.time()
    q.execute()
    q.finalize_device()
    q.sync()
    wall = (time.time() - t0) * 1e3
    tm = q.timings
    out["passes"].append({"wall_ms": wall, **tm})
rows, byts = q.total_rows, q.touched_bytes
best = min(p["scan_ms"] for p in out["passes"])
out["rows"] = rows
out["touched_bytes"] = byts
out["survivors"] = q.survivors
out["best_scan_ms"] = best
out["scan_GBps"] = byts / best / 1e6
out["scan_Grows_s"] = rows / best / 1e6
res = q.finalize()
out["result_rows"] = res.num_rows
print(json.dumps(out))

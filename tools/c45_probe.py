"""Device-side timing probe for the two other BASELINE.json configurations (not the benchmark line):
C4 = regex-on-dictionary predicate + 10^6 tag combinations (100 segments x 1 Mi rows), C5 = K-way merge of 256 sorted
streams x 65536 elements (algorithmic bytes 2 x 20 B per element)."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

from lakeside_b200 import _lib, api, synth  # noqa: E402

api.init()
out = {}

# ---- C5 ----
K, M, T0 = 256, 65536, 1699999200000
rng = np.random.Generator(np.random.PCG64(20240))
ts = [np.sort(T0 + 10000 * rng.integers(0, 360, M)).astype(np.int64) for _ in range(K)]
gid = [rng.integers(0, 16384, M).astype(np.int32) for _ in range(K)]
val = [rng.standard_normal(M) for _ in range(K)]
lib = _lib.load()
P = ctypes.c_void_p
lens = (ctypes.c_int64 * K)(*[M] * K)
h = P()
_lib.check(lib.lk_merge_create(K, (P * K)(*[t.ctypes.data for t in ts]), (P * K)(*[g.ctypes.data for g in gid]),
                               (P * K)(*[v.ctypes.data for v in val]), lens, 0, ctypes.byref(h)))
times = []
for _ in range(8):
    _lib.check(lib.lk_merge_run(h))
    _lib.check(lib.lk_merge_sync(h))
    ms = (ctypes.c_double * 4)()
    _lib.check(lib.lk_merge_timings(h, ms))
    times.append(ms[1])
n = K * M
o_ts = np.empty(n, np.int64)
_lib.check(lib.lk_merge_download(h, o_ts.ctypes.data, None, None, None))
assert (np.diff(o_ts) >= 0).all()
lib.lk_merge_destroy(h)
best = min(times[2:])
out["c5_merge"] = {"streams": K, "elements": n, "merge_ms": best, "all_ms": times, "GBps_algorithmic_40B_per_element": 40.0 * n / (best / 1e3) / 1e9,
                   "elements_per_s": n / (best / 1e3)}

# ---- C4 ----
segs = int(os.environ.get("C4_SEGMENTS", "100"))
spec = synth.c4_spec(1 << 20)
paths = synth.write_dataset("/tmp/lk_probe/c4_1048576", spec, segs)
rq = json.dumps(synth.push_down_request(synth.c4_base_expr(), list(range(segs)), 10000))
q = api.Query(rq, aggregates=synth.C2_AGGREGATES, path=os.environ.get("C4_PATH", "auto"))
for p in paths:
    q.add_segment_file(p)
q.prepare()
passes = []
for _ in range(6):
    t0 = time.time()
    q.execute()
    q.finalize_device()
    q.sync()
    passes.append({"wall_ms": (time.time() - t0) * 1e3, **q.timings})
best = min(p["scan_ms"] for p in passes[1:])
out["c4_regex_high_cardinality"] = {"segments": segs, "rows": q.total_rows, "touched_bytes": q.touched_bytes, "survivors": q.survivors,
                                    "path": q.info["path"], "n_groups": q.info["n_groups"], "best_scan_ms": best,
                                    "scan_GBps_algorithmic": q.touched_bytes / (best / 1e3) / 1e9,
                                    "step_ms": min(p["wall_ms"] for p in passes[1:]), "passes": passes[1:]}
q.close()
print(json.dumps(out))

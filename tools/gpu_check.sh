#!/bin/bash
# One GPU-box session: the gpu-marked tests, then the device-side timing probe on the C2 workload (auto and hash tables).
tag=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/perf_probe.py --segments 100 --steps 6 > gpurun_out/probe_${tag}.json 2> gpurun_out/probe_${tag}.err
python tools/perf_probe.py --segments 100 --steps 6 --path hash > gpurun_out/probe_${tag}_hash.json 2>> gpurun_out/probe_${tag}.err
python - <<PY
import json
for f in ["gpurun_out/probe_${tag}.json", "gpurun_out/probe_${tag}_hash.json"]:
    d = json.load(open(f))
    print(d["info"]["path"], [(round(p["scan_ms"], 3), round(p.get("def_expand_ms", 0), 3), round(p["finalize_ms"], 3), round(p["wall_ms"], 3)) for p in d["passes"]])
PY

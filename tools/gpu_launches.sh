#!/bin/bash
# Timing probe on the C2 workload, then the launch list of the same command (ncu, per-launch device times).
tag=${1:-x}
python tools/perf_probe.py --segments 100 --steps 6 > gpurun_out/probe_${tag}.json 2> gpurun_out/probe_${tag}.err || { tail -5 gpurun_out/probe_${tag}.err; exit 1; }
python - <<PY
import json
d = json.load(open("gpurun_out/probe_${tag}.json"))
print(d["info"]["path"], d["result_rows"], [(round(p["scan_ms"], 3), round(p.get("def_expand_ms", 0), 3), round(p["finalize_ms"], 3), round(p["wall_ms"], 3)) for p in d["passes"]])
PY
python tools/perf_probe.py --segments 100 --steps 2 > gpurun_out/plain_${tag}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${tag}.csv python tools/perf_probe.py --segments 100 --steps 2 > gpurun_out/ncu_${tag}.log 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/launches_${tag}.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:60], r[-1])
PY

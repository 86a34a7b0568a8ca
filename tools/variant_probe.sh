#!/bin/bash
# Tuning aid: times the C2 scan with every kernel variant built into lakeside_b200/variants/ (make OUT=... EXTRA=-D...).
for so in lakeside_b200/liblakeside_b200.so lakeside_b200/variants/*.so; do
  [ -f "$so" ] || continue
  for path in ${PATHS:-auto}; do
   for l2 in ${L2S:-0}; do
    if [ "$l2" != 0 ]; then export LK_L2_FETCH=$l2; else unset LK_L2_FETCH; fi
    LK_LIB=$PWD/$so python tools/perf_probe.py --segments 100 --steps 8 --path $path > /tmp/vp.json 2> /tmp/vp.err || { echo "$so FAILED"; tail -3 /tmp/vp.err; continue; }
    python - "$so" "$path" "$l2" <<'PY'
import json, sys
d = json.load(open("/tmp/vp.json"))
ps = d["passes"][1:]
f = lambda k: round(min(p.get(k, 0) for p in ps), 4)
print(sys.argv[1].split("/")[-1], sys.argv[2], "l2fetch", sys.argv[3], d["info"]["path"], "scan", f("scan_ms"), "def", f("def_expand_ms"), "fin", f("finalize_ms"), "wall", f("wall_ms"), "surv", d["survivors"], "rows", d["result_rows"])
PY
   done
  done
done

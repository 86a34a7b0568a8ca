"""Hot SASS instructions (stall samples) of one kernel of an ncu report: python tools/ncu_hot.py report.ncu-rep kernel_regex [min_pct]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows if len(r) > si and r[si].isdigit()]
tot = sum(int(r[si]) for r in body)
print("kernel", kern, "samples", tot, "warp-inst", sum(int(r[ii]) for r in body))
agg = {}
for r in body:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for n, r in enumerate(body):
    if int(r[si]) > tot * minpct / 100:
        top = max(stalls, key=lambda i: int(r[i] or 0))
        print(f"{n:5d} {100 * int(r[si]) / tot:5.1f}% inst={r[ii]:>9} {hdr[top]:<16} {r[1].strip()[:100]}")

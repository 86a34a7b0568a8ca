"""Aggregates an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = None
out = []
hdr = None
for r in rows:
    if r and r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    def g(name):
        try:
            return int(r[hdr.index(name)] or 0)
        except (ValueError, IndexError):  # "-" cells, or a source line whose quotes confused the CSV export
            return 0
    out.append((g("# Samples"), g("stall_long_sb"), g("stall_wait"), g("stall_short_sb"), g("stall_no_inst"), g("Instructions Executed"), f, r[0], r[1].strip()[:100]))
tot = sum(o[0] for o in out); toti = sum(o[5] for o in out)
print("samples", tot, "warp-inst", toti)
print("  samp    %  longsb  wait shortsb noinst       inst   %i  where")
for o in sorted(out, reverse=True)[:top]:
    print(f"{o[0]:6d} {100*o[0]/tot:4.1f} {o[1]:7d} {o[2]:5d} {o[3]:7d} {o[4]:6d} {o[5]:10d} {100*o[5]/toti:4.1f}  {o[6]}:{o[7]} {o[8]}")

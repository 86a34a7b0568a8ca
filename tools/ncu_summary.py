"""Key metrics of one kernel from an .ncu-rep (ncu -i ... --page raw --csv), as one JSON object."""
import csv, json, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
def f(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return None
keys = {
 "duration_us": "gpu__time_duration.sum", "dram_read_GB": "dram__bytes_read.sum", "dram_write_MB": "dram__bytes_write.sum",
 "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "warp_instructions": "smsp__inst_executed.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
 "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
 "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
 "warp_latency_per_inst": "smsp__average_warp_latency_per_inst_issued.ratio",
 "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
 "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
 "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "stall_no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
 "stall_branch": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
 "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
 "stall_lg_throttle": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
 "stall_mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
 "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
}
out = {"kernel": d.get("Kernel Name"), "report": rep.split("/")[-1]}
for k, m in keys.items():
    out[k] = f(m)
    if k == "dram_read_GB" and units[hdr.index(m)] == "Mbyte": out[k] = out[k] / 1e3
    if k == "dram_write_MB" and units[hdr.index(m)] == "Gbyte": out[k] = out[k] * 1e3
    if k == "duration_us" and units[hdr.index(m)] == "ms": out[k] = out[k] * 1e3
print(json.dumps(out))

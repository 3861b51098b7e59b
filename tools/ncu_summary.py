"""Per-kernel summary (json) of an ncu --set full report: python tools/ncu_summary.py <report.ncu-rep> <out.json> "<note>" """
import csv, json, subprocess, sys
rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "smsp__inst_executed.sum": "warp_inst", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "launch__registers_per_thread": "regs",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts", "launch__grid_size": "grid",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier"}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3}
kn = hdr.index("Kernel Name")
agg = {}
for r in rows[2:]:
    name = r[kn].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, {"launches": 0})
    a["launches"] += 1
    for w, n in want.items():
        if w not in hdr:
            continue
        i = hdr.index(w)
        try:
            v = float(r[i].replace(",", "")) * mult.get(units[i], 1)
        except ValueError:
            continue
        a[n] = a.get(n, 0.0) + v
res = {}
for name, a in agg.items():
    n = a.pop("launches")
    d = {k: v / n for k, v in a.items()}
    res[name] = {"launches": n, "dram_bytes_per_launch": d.get("dram_read", 0) + d.get("dram_write", 0),
                 "ncu_us_per_launch": d.get("us"), "warp_inst_per_launch": d.get("warp_inst"),
                 **{k: d[k] for k in ("warps_active_pct", "issue_active_pct", "regs", "tensor_pct", "dram_pct",
                                      "smem_wavefronts", "grid", "stall_long_scoreboard", "stall_barrier") if k in d}}
json.dump({"source": note, "kernels": res}, open(out, "w"), indent=1)
for k, v in res.items():
    print("%-60s x%d %8.1f us  dram %7.1f MB  issue %5.1f%%  warps %5.1f%%" % (k[:60], v["launches"], v["ncu_us_per_launch"], v["dram_bytes_per_launch"] / 1e6, v.get("issue_active_pct", 0), v.get("warps_active_pct", 0)))

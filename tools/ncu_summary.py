"""Summaries for profiles/: (1) per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv),
(2) selected metrics of an ncu --set full report (ncu -i X.ncu-rep --page raw --csv > raw.csv).
usage: python tools/ncu_summary.py launches launches.csv | python tools/ncu_summary.py raw raw.csv"""
import collections, csv, sys
mode, path = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
if mode == "launches":
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID"); H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi: continue
        v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0]
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("kernel,launches,total_us,avg_us,share")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k},{len(v)},{sum(v):.1f},{sum(v)/len(v):.1f},{100*sum(v)/tot:.1f}%")
else:
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID"); H = rows[h]; U = rows[h + 1]
    want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.per_cycle_active", "sm__inst_executed.avg.per_cycle_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed.sum"]
    want += [c for c in H if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("per_issue_active.ratio")]
    data = rows[h + 2:]
    print("metric,unit," + ",".join("launch%d" % i for i in range(len(data))))
    for m in want:
        if m in H:
            c = H.index(m); print(f"{m},{U[c]}," + ",".join(r[c].replace(",", "") for r in data))

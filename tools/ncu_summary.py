"""Condenses `ncu --page raw --csv` exports (one file per kernel family, gpurun_out/raw_<kernel>_r01.csv) into
a summary JSON: per captured launch the duration, DRAM bytes, and the pipe / issue utilisation figures DESIGN.md and
bench.py's roofline.traffic quote.   python tools/ncu_summary.py [--out profiles/ncu_summary_r02.json] gpurun_out/raw_*.csv"""
import csv, json, os, sys

KEYS = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active": "pipe_fmaheavy_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active": "pipe_fmaheavy_cycles_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__inst_executed.sum": "inst_executed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle_per_issue",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle_per_issue",
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return x


def main():
    out = {}
    args = sys.argv[1:]
    dest = "profiles/ncu_summary_r01.json"
    if args and args[0] == "--out":
        dest, args = args[1], args[2:]
    for path in args:
        rows = list(csv.reader(open(path)))
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        names, units = rows[hdr], rows[hdr + 1]
        fam = os.path.basename(path).replace("raw_", "").replace("_r01.csv", "").replace(".csv", "")
        out[fam] = []
        for r in rows[hdr + 2:]:
            if len(r) != len(names):
                continue
            d = {"kernel": r[names.index("Kernel Name")][:90]}
            for i, n in enumerate(names):
                if n in KEYS:
                    v = num(r[i])
                    if KEYS[n] == "time":
                        d["time_ms"] = v / 1e6 if units[i] in ("ns", "nsecond") else (v / 1e3 if units[i].startswith("u") else v)
                        d["time_unit_raw"] = units[i]
                    elif KEYS[n] in ("dram_read", "dram_write"):
                        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
                        d[KEYS[n] + "_bytes"] = v * scale
                    else:
                        d[KEYS[n]] = v
            out[fam].append(d)
    json.dump(out, open(dest, "w"), indent=1)
    for fam, ls in out.items():
        for d in ls:
            print(fam, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k not in ("kernel", "time_unit_raw")})


main()

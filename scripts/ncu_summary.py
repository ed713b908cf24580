"""Summarise an .ncu-rep (read here, no GPU needed) into a small tracked CSV/JSON under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name [algorithmic_bytes_per_launch]
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
    "smsp__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    alg = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_col = hdr.index("Kernel Name")
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}:{r[name_col][:40]}" for i, r in enumerate(data)])
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                w.writerow([k, units[i]] + [r[i] for r in data])
    summary = {"report": rep, "kernels": [r[name_col] for r in data]}
    if "dram__bytes_read.sum" in hdr:
        ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        per = []
        for r in data:
            rd = float(r[ir]) * UNIT_SCALE.get(units[ir], 1.0)
            wr = float(r[iw]) * UNIT_SCALE.get(units[iw], 1.0)
            tus = float(r[it]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[it], 1.0)
            per.append({"dram_read_bytes": rd, "dram_write_bytes": wr, "duration_us": tus,
                        "dram_GBps_under_ncu": (rd + wr) / tus / 1e3})
        summary["per_launch"] = per
        if alg:
            summary["algorithmic_bytes_per_launch"] = alg
            summary["dram_bytes_per_input_byte"] = sum(p["dram_read_bytes"] + p["dram_write_bytes"] for p in per) / len(per) / alg
    summary["source"] = f"ncu --set full --clock-control none ({rep})"
    json.dump(summary, open(out + ".json", "w"), indent=1)
    print(json.dumps(summary)[:600])


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Condense an ncu report (.ncu-rep, `--set full`) into the small metric,unit,value CSV kept under profiles/.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep profiles/X.csv [--traffic STAGE FRAMES]

--traffic also records dram__bytes_read.sum + dram__bytes_write.sum of the (first) captured launch under
profiles/ncu_traffic.json[STAGE], which bench.py reports as roofline.traffic.
"""
import json
import os
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "sm__maximum_warps_per_active_cycle_pct")


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
            f.write(f"# kernel,{name}\nmetric,unit,value\n")
            for h, u, v in zip(hdr, units, r):
                if h in KEEP:
                    f.write(f"{h},{u},{v}\n")


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value.replace(",", "")) * scale


def traffic(csv_path, stage, frames):
    rd = wr = None
    name = ""
    for line in open(csv_path):
        if line.startswith("# kernel,") and not name:
            name = line.strip().split(",", 1)[1]
        f = line.strip().split(",")
        if f[0] == "dram__bytes_read.sum" and rd is None:
            rd = to_bytes(f[2], f[1])
        if f[0] == "dram__bytes_write.sum" and wr is None:
            wr = to_bytes(f[2], f[1])
    out = os.path.join(os.path.dirname(os.path.abspath(csv_path)), "ncu_traffic.json")
    rec = json.load(open(out)) if os.path.exists(out) else {}
    rec[stage] = {"kernel": name, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
                  "frames": int(frames), "source": os.path.basename(csv_path)}
    json.dump(rec, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
    if len(sys.argv) >= 6 and sys.argv[3] == "--traffic":
        traffic(sys.argv[2], sys.argv[4], sys.argv[5])

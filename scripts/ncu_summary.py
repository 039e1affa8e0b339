"""Summarise an .ncu-rep (ncu --set full) into profiles/: per-kernel key metrics (CSV) and the
measured DRAM traffic per launch that bench.py reports as roofline.traffic.

    python scripts/ncu_summary.py gpurun_out/prof_r01_v5.ncu-rep r01_v5 C5
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "launch__block_size", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, tag, workload = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    path = os.path.join(ROOT, "profiles", f"{tag}_ncu_full.csv")
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in data:
            w.writerow([r[c][:160] for c in cols])
    traffic = {}
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    for r in data:
        name = r[ik].split("(")[0].replace("void ", "").split("<")[0].strip()
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        traffic.setdefault(name, []).append(b)
    traffic = {k: sum(v) / len(v) for k, v in traffic.items()}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
    allt[workload] = dict(traffic, source=os.path.basename(path))
    with open(tpath, "w") as f:
        json.dump(allt, f, indent=1)
    print(path, traffic)


if __name__ == "__main__":
    main()

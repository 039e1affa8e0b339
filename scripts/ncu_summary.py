"""Summarise an .ncu-rep (ncu --set full) into profiles/: per-kernel key metrics (CSV) and the
per-PARTICLE counters that bench.py turns into roofline.traffic / issue_frac / fp32_pipe_frac with
the live kernel time (profiles/r02_counters.json: DRAM bytes, warp instructions, FMA-pipe busy cycles).

    python scripts/ncu_summary.py gpurun_out/prof_r02.ncu-rep r02_v9_C5 C5/reference 16000000
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
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_elapsed.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, tag, key, particles = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
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
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                                     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_elapsed.sum",
                                     "gpu__time_duration.sum")}
    acc = {}
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].strip().split("::")[-1]
        byt = sum(float(r[col[k]]) * UNIT[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        inst = float(r[col["smsp__inst_executed.sum"]])
        fma = float(r[col["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"]]) / 100.0 * \
            float(r[col["smsp__cycles_elapsed.sum"]])           # busy FMA-pipe cycles, summed over the 592 sub-partitions
        acc.setdefault(name, []).append((byt, inst, fma, float(r[col["gpu__time_duration.sum"]])))
    counters = {}
    for name, v in acc.items():
        n = len(v)
        counters[name] = {"dram_bytes_per_particle": sum(x[0] for x in v) / n / particles,
                          "warp_inst_per_particle": sum(x[1] for x in v) / n / particles,
                          "fma_pipe_cycles_per_particle": sum(x[2] for x in v) / n / particles,
                          "ncu_ms_per_launch": sum(x[3] for x in v) / n, "launches_captured": n}
    cpath = os.path.join(ROOT, "profiles", "r02_counters.json")
    allc = json.load(open(cpath)) if os.path.exists(cpath) else {}
    allc[key] = dict(counters, source=os.path.basename(path), particles=particles)
    with open(cpath, "w") as f:
        json.dump(allc, f, indent=1)
    print(path, json.dumps(counters, indent=1))


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Turn the `ncu --set full` reports of tools/final_capture.sh into the committed summaries under profiles/:
   <tag>_ncu_full_raw_selected_C2.csv  selected raw metrics, one row per kernel
   <tag>_ncu_full_details_*.txt        the details pages
   <tag>_traffic.json                  DRAM bytes (read + write) per launch; bench.py reports them as roofline.traffic
usage: python tools/ncu_extract.py [tag] (default r02; reads gpurun_out/<tag>_full_C2.ncu-rep, <tag>_full_C4_encode.ncu-rep)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True, check=True).stdout


def main():
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_full_C2.ncu-rep")
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    keep = ["Kernel Name", "Block Size", "Grid Size"] + METRICS
    out = [keep, [units[col[k]] for k in keep]]
    traffic = {}
    for r in rows[2:]:
        out.append([r[col[k]] for k in keep])
        name = r[col["Kernel Name"]].split("(")[0].split("::")[-1]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        b = sum(float(r[col[m]]) * scale[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic["encode_frames_kernel" if name.startswith("encode_frames") else name] = b
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_raw_selected_C2.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out)
    with open(os.path.join(ROOT, "profiles", f"{tag}_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_details_C2.txt"), "w") as f:
        f.write(ncu("-i", rep, "--page", "details"))
    rep4 = os.path.join(ROOT, "gpurun_out", f"{tag}_full_C4_encode.ncu-rep")
    if os.path.exists(rep4):
        with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_details_C4_encode.txt"), "w") as f:
            f.write(ncu("-i", rep4, "--page", "details"))
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()

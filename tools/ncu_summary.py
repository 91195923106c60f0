#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [substring ...]
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_global_ld.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print("==", name)
        for h, u, v in zip(hdr, units, r):
            if h in KEYS or "issue_stalled" in h and h.endswith("per_issue_active.ratio") or any(e in h for e in extra):
                print(f"  {h:88s} {v:>18s} {u}")


if __name__ == "__main__":
    main()

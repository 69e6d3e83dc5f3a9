#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into the few lines the design decisions rest on.
usage: tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for k, row in enumerate(rows[2:]):
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print("== launch %d: %s  grid %s block %s" % (k, d.get("Kernel Name"), d.get("Grid Size"),
                                                      d.get("Block Size")))
        for key in KEYS:
            if key in d and d[key] != "":
                print("  %-68s %s %s" % (key, d[key], u[key]))
        stalls = []
        pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        for h in hdr:
            if h.startswith(pre) and h.endswith(suf) and "not_issued" not in h:
                try:
                    stalls.append((float(d[h]), h[len(pre):-len(suf)]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  warp stall reasons (avg warps stalled per issue-active cycle):")
        for v, n in stalls[:9]:
            print("    %-28s %.3f" % (n, v))


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python
"""Key metrics of an ncu report in the `name unit value` form of profiles/*_key_metrics.txt.

usage: python tools/ncu_summary.py REPORT.ncu-rep [kernel-name-substring] > profiles/NAME_key_metrics.txt

Reads `ncu -i REPORT --page raw --csv` (ncu is in the CUDA toolkit of this image; no GPU needed) and prints, per
launch whose kernel name contains the substring, the metrics bench.py and DESIGN.md quote: duration, launch
configuration, DRAM bytes, unit utilisations, shared / global wavefronts and the stall breakdown."""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_ldgsts.sum", "sm__cycles_elapsed.avg",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "sm__sass_thread_inst_executed_op_dadd_pred_on.sum",
]


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    col = {}
    for i, n in enumerate(names):
        col.setdefault(n.split(".", 2)[-1] if n.count(".") >= 2 and n.split(".")[1][:1].isupper() else n, i)
        col.setdefault(n, i)
    k_col = names.index("Kernel Name")
    stall = [n for n in names if "issue_stalled" in n and n.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        if want not in r[k_col]:
            continue
        print(f"== {r[k_col]}  (launch id {r[0]})")
        for m in KEEP + stall:
            i = col.get(m)
            if i is None:
                i = next((j for j, n in enumerate(names) if n.endswith(m)), None)
            if i is None or r[i] == "":
                continue
            print(m.split("TriageCompute.")[-1], units[i], r[i])
        print()


if __name__ == "__main__":
    main()

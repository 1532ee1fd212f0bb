#!/usr/bin/env python
"""Summarise an `ncu --set full` report into one JSON line per kernel launch (the files kept under profiles/).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<round>_kernels_ncu_summary.jsonl
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_requests_op_red.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    for r in rows[2:]:
        d = {"Kernel Name": r[col["Kernel Name"]].replace("ps::", "")}
        for k in KEEP:
            if k in col and r[col[k]] != "":
                u = units[col[k]]
                d[f"{k} [{u}]" if u else k] = r[col[k]]
        print(json.dumps(d))


if __name__ == "__main__":
    main()

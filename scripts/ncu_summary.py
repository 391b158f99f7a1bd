#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into a small JSON list, one record per captured launch.

    python scripts/ncu_summary.py gpurun_out/prof_gemm.ncu-rep profiles/r01_ncu_full_gemm.json
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "smsp__sass_inst_executed_op_tmem_ldt.sum", "launch__grid_size", "launch__block_size",
    "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        rec = {}
        for h, u, v in zip(hdr, units, r):
            if h in ("Kernel Name", "Block Size", "Grid Size", "ID"):
                rec[h] = v
            elif h in KEYS:
                rec[f"{h} [{u}]" if u else h] = v
        recs.append(rec)
    with open(out, "w") as f:
        json.dump(recs, f, indent=1)
    print(f"{len(recs)} launches -> {out}")


if __name__ == "__main__":
    main()

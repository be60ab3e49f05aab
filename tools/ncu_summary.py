#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/: duration, DRAM bytes, pipe
utilisation, issue activity, registers, L2 traffic, top stall reasons.  Usage: ncu_summary.py in.ncu-rep out.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__inst_executed.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# summary of {rep} (ncu --set full --clock-control none); one block per captured launch\n")
        for r in rows[2:]:
            vals = dict(zip(hdr, r))
            if vals.get("gpu__time_duration.sum", "nan") in ("nan", "-nan", ""):
                continue
            if vals.get("dram__bytes_read.sum", "nan") in ("nan", "-nan", ""):
                continue
            f.write(f"\nkernel: {vals.get('Kernel Name')}\n")
            u = dict(zip(hdr, units))
            for k in KEYS:
                if k in vals and vals[k] not in ("", "nan", "-nan"):
                    f.write(f"  {k:90s} {vals[k]:>16s} {u.get(k, '')}\n")
            stalls = [(float(v), h) for h, v in vals.items()
                      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
                      and v not in ("", "nan", "-nan")]
            for v, h in sorted(stalls, reverse=True)[:6]:
                name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
                f.write(f"  stall {name:84s} {v:16.3f} warps/issue\n")
    print(open(out).read())


if __name__ == "__main__":
    main()

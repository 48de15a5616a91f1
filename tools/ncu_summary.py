#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of counters the roofline uses.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__cycles_elapsed.avg", "SM cycles elapsed (avg)"),
    ("smsp__cycles_active.avg", "SMSP cycles active (avg)"),
    ("sm__inst_executed_pipe_tensor", "tensor-pipe instructions"),
    ("sm__pipe_tensor_cycles_active", "tensor pipe active"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active", "tensor hmma subpipe active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts (LSU)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_xu", "XU (MUFU) instructions"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# ncu summary of `{rep}` ({len(data)} launches captured, `--set full --clock-control none`)\n")
    for li, r in enumerate(data):
        print(f"## launch {li}: `{r[name_i][:90]}`\n")
        print("| counter | value | unit |\n|---|---|---|")
        for i, h in enumerate(hdr):
            for key, label in WANT:
                if key in h and "peak_sustained" not in h.replace("pct_of_peak_sustained", "") and ".max" not in h and ".min" not in h:
                    print(f"| `{h}` | {r[i]} | {units[i]} |")
                    break
        print()


if __name__ == "__main__":
    main()

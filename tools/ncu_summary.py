#!/usr/bin/env python
"""Prints a compact per-launch summary of an .ncu-rep (raw page): duration, DRAM bytes, L2/SM throughput,
tensor-pipe activity, occupancy, top warp-stall reasons.  Usage: tools/ncu_summary.py file.ncu-rep [more]"""
import csv, subprocess, sys, io

WANT = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
]

def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(h, r))
            u = dict(zip(h, units))
            print("==", d.get("Kernel Name", "?")[:90], "id", d.get("ID"))
            for k, label in WANT:
                if k in d:
                    print(f"   {label:12s} {d[k]:>16s} {u[k]}")
            stalls = [(float(v.replace(',', '') or 0), k) for k, v in d.items()
                      if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v]
            for v, k in sorted(stalls, reverse=True)[:6]:
                print(f"   stall {k.split('stalled_')[1].split('_per_issue')[0]:28s} {v:8.2f}")

if __name__ == "__main__":
    main()

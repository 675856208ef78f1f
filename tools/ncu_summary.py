"""Text summary of an .ncu-rep (ncu --set full): per kernel the duration, DRAM bytes, hit rates, throughputs, occupancy,
instruction counts and the top stall reasons.   python tools/ncu_summary.py report.ncu-rep [title] > profiles/x.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__cycles_elapsed.max"]
print(f"# {title}")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"== {name[:110]}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:62s} {r[i]:>16s} {units[i]}")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, n in sorted(stalls, reverse=True)[:5]:
        print(f"   stall {n:40s} {v:8.2f} warps per issue")

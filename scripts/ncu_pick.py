"""Print selected metrics from an .ncu-rep (raw page):  python scripts/ncu_pick.py rep [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"^(gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__cycles_elapsed.max|sm__cycles_active.avg|launch__(registers_per_thread|occupancy_limit|waves|grid_size|block_size|shared_mem_per_block_dynamic)|sm__warps_active.avg.pct|sm__throughput.avg.pct|smsp__issue_active.avg.pct|sm__inst_executed_pipe_(xu|fma|alu|fmaheavy|uniform|lsu|tmem|cbu).*pct_of_peak_sustained_active|sm__pipe_tensor.*pct|gpu__dram_throughput.avg.pct|lts__throughput.avg.pct|l1tex__throughput.avg.pct|smsp__average_warp.*_per_issue_active|smsp__inst_executed.sum$|sm__inst_executed.avg.per_cycle_active|smsp__warp_issue_stalled.*_per_warp_active)")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:100], r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h):
            print(f"  {h:95s} {v} {u}")

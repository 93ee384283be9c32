"""Print the event timeline written by the TRACE instance of self_attn_tc3.cu (DADD_ATTN_TRACE=<file>).

    DADD_ATTN_TRACE=gpurun_out/attn_trace.txt python scripts/kbench.py --kernel self_attn --batch 26 --iters 1 --no-flush --dtype fp16
    python scripts/attn_trace.py gpurun_out/attn_trace.txt

Rows = (warp, step); softmax warps (0-15: group = warp / 8, column half = (warp / 4) & 1): 0 loop top, 1 S ready, 2 S in
registers, 3 local max done, 4 partner max seen, 5 exp2 phase starts, 6 P stored + arrive.  Warp 17 = MMA issuer, rows are
TILES: 0 loop top, 1 PV(i) issued, 2 QK(i + NBUF) issued.  Clocks are relative to the earliest event of the file.
"""
import sys

STEPS, EVENTS = 24, 8
rows = [list(map(int, l.split())) for l in open(sys.argv[1]) if l.strip()]
t0 = min(v for r in rows for v in r if v > 0)
first, last = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (6, 12)
for warp in (0, 4, 8, 12, 1, 17):
    for step in range(first, last):
        r = rows[warp * STEPS + step]
        print(f"warp {warp:2d} step {step:2d}: " + " ".join(f"{(v - t0) if v else -1:7d}" for v in r[:7]))
    print()

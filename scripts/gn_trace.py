"""Timeline of CTA 0 of the streaming GroupNorm kernel (DADD_GN_TRACE): clocks relative to the first event, per role and item."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r2/gn_trace.txt"
B, C, R = int(os.environ.get("GN_B", 104)), int(os.environ.get("GN_C", 320)), int(os.environ.get("GN_R", 32))
os.environ["DADD_GN_TRACE"] = path
from progressive_stable_diffusion_b200 import ops
x = torch.randn(B, C, R, R, device="cuda", dtype=torch.float16).contiguous(memory_format=torch.channels_last)
g, b, add = torch.randn(C, device="cuda"), torch.randn(C, device="cuda"), torch.randn(B, C, device="cuda")
for _ in range(3):
    ops.group_norm(x, g, b, 32, 1e-5, True, add)
torch.cuda.synchronize()
lines = open(path).read().strip().split("\n")
print(lines[0])
rows = [[int(v) for v in l.split()] for l in lines[1:]]
t0 = min(v for r in rows for v in r if v > 0)
names = ["consumer: start | full+chs_free | stats | sync | ready(k-1) | apply(k-1)", "publisher: sync | partials stored | fence | release",
         "combiner: start | flags seen | partials added | ready", "producer: start | stage free"]
for role in range(4):
    print(names[role])
    for it in range(16):
        r = rows[role * 16 + it]
        if any(r):
            print(f"  item {it:2d}: " + " ".join(f"{(v - t0) if v else 0:7d}" for v in r[:6]))

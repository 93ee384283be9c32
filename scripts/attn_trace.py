"""Dump and print CTA 0's event timeline of the tcgen05 self-attention kernel (debug aid; DADD_ATTN_TRACE)."""
import os, sys
os.environ["DADD_ATTN_TRACE"] = "/tmp/attn_trace.txt"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_stable_diffusion_b200 import ops
B, N, C = 26, 1024, 320
qkv = torch.randn(B, N, 3 * C, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.self_attention(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], 8)
rows = [list(map(int, l.split())) for l in open("/tmp/attn_trace.txt")]
t0 = min(v for r in rows for v in r if v > 0)
names = {0: "WG0", 1: "WG1", 2: "MMA"}
for step in range(0, 26):
    for role in range(3):
        r = rows[role * 64 + step]
        print(f"step {step:2d} {names[role]}: " + " ".join(f"{(v - t0) if v else -1:7d}" for v in r[:13]))

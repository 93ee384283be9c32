"""Micro-benchmark of the hot-path kernels at the bench workload's shapes (B = patients x 13 samples, 256x256).

    python scripts/kbench.py [--kernel self_attn,gn,...] [--batch 26] [--iters 20] [--res 32]

Every kernel is timed with CUDA events on the launching stream, L2 flushed (a 256 MB memset) before every timed launch,
and reported against the measured peaks in MEASURED_PEAKS.json.  Under ncu use ``--iters 1 --no-flush`` and ``-k regex:``.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from progressive_stable_diffusion_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kernel", default="all")
ap.add_argument("--batch", type=int, default=26)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--res", type=int, default=32, help="latent side (32 = 256x256 images, 64 = 512x512)")
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--no-flush", action="store_true")
args = ap.parse_args()

dev = torch.device("cuda", 0)
dt = torch.bfloat16 if args.dtype == "bf16" else torch.float16
peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
try:
    peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
except OSError:
    pass
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B = args.batch
R = args.res
torch.manual_seed(0)


def _graph_time(body, reps):
    """Mean device time of ``body()`` captured ``reps`` times into one CUDA graph (no host launch overhead in the number)."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        body()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                body()
        g.replay()
        torch.cuda.synchronize()
        best = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) * 1e3 / reps)
    return min(best)


_flush_us = [None]


def timeit(fn):
    """(cold, hot) microseconds per launch: cold = L2 flushed before every launch (flush time subtracted), hot = back to
    back launches with the operands resident in L2 (how the kernel runs inside the UNet step)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if args.no_flush:
        fn()
        torch.cuda.synchronize()
        return 0.0, 0.0
    if _flush_us[0] is None:
        _flush_us[0] = _graph_time(lambda: flush_buf.zero_(), args.iters)

    def cold():
        flush_buf.zero_()
        fn()
    return _graph_time(cold, args.iters) - _flush_us[0], _graph_time(fn, args.iters)


def report(name, us, us_hot, flops=None, bytes_=None, **kw):
    d = {"kernel": name, "us_cold": round(us, 2), "us_hot_l2": round(us_hot, 2)}
    if us <= 0:
        print(json.dumps(d), flush=True)
        return
    if flops is not None:
        d["tflops"] = round(flops / us / 1e6, 1)
        d["frac_tensor_peak"] = round(flops / us / 1e6 / peaks["bf16_tflops"], 4)
        d["tflops_hot"] = round(flops / us_hot / 1e6, 1)
    if bytes_ is not None:
        d["gbs"] = round(bytes_ / us / 1e3, 1)
        d["frac_hbm_peak"] = round(bytes_ / us / 1e3 / peaks["hbm_gbs"], 4)
        d["gbs_hot_l2"] = round(bytes_ / us_hot / 1e3, 1)
    d.update(kw)
    print(json.dumps(d), flush=True)


want = set(args.kernel.split(","))


def on(k):
    return "all" in want or k in want


sites = [(320, R * R), (640, (R // 2) ** 2), (1280, (R // 4) ** 2), (1280, (R // 8) ** 2)]

if on("self_attn"):
    for C, N in sites:
        qkv = torch.randn(B, N, 3 * C, device=dev, dtype=dt)
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        us, mn = timeit(lambda: ops.self_attention(q, k, v, 8))
        report(f"self_attn N={N} C={C} d={C // 8} B={B}", us, mn, flops=4.0 * N * N * C * B)

if on("cross"):
    gates = torch.tensor([0.9, 0.1, 3.0], device=dev)
    for C, N in sites:
        q = torch.randn(B, N, C, device=dev, dtype=dt)
        kc = torch.randn(B, 8, 48, C // 8, device=dev, dtype=dt)
        vc = torch.randn(B, 8, 48, C // 8, device=dev, dtype=dt)
        impl = os.environ.get("KBENCH_XATTN_IMPL", "auto")
        us, mn = timeit(lambda: ops.cross_attention(q, kc, vc, gates, 8, 16, 3, impl=impl))
        report(f"cross_attn[{impl}] N={N} C={C} B={B}", us, mn, bytes_=4.0 * N * C * B + 2 * 2.0 * B * 48 * C)

if on("gn"):
    shapes = [(320, R), (640, R), (960, R), (640, R // 2), (1280, R // 2), (1920, R // 2), (1280, R // 4), (2560, R // 4),
              (1280, R // 8), (2560, R // 8)]
    for C, S in shapes:
        x = torch.randn(B, C, S, S, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
        g = torch.randn(C, device=dev)
        bta = torch.randn(C, device=dev)
        add = torch.randn(B, C, device=dev)
        y = torch.empty_like(x)
        us, mn = timeit(lambda: ops.group_norm(x, g, bta, 32, 1e-5, True, add, out=y))
        report(f"gn+silu NHWC C={C} HW={S}x{S} B={B}", us, mn, bytes_=4.0 * x.numel())

if on("ln"):
    for C, N in sites:
        x = torch.randn(B, N, C, device=dev, dtype=dt)
        g = torch.randn(C, device=dev)
        bta = torch.randn(C, device=dev)
        us, mn = timeit(lambda: ops.layer_norm(x, g, bta, 1e-5))
        report(f"layernorm N={N} C={C} B={B}", us, mn, bytes_=4.0 * x.numel())

if on("geglu"):
    for C, N in sites:
        x = torch.randn(B, N, 8 * C, device=dev, dtype=dt)
        us, mn = timeit(lambda: ops.geglu(x))
        report(f"geglu N={N} C={C} B={B}", us, mn, bytes_=2.0 * x.numel() * 1.5)

if on("ff1"):
    import torch.nn.functional as F
    for C, N in sites:
        x = torch.randn(B, N, C, device=dev, dtype=dt)
        w = torch.randn(8 * C, C, device=dev, dtype=dt) * (C ** -0.5)
        b32 = torch.randn(8 * C, device=dev)
        b16 = b32.to(dt)
        flops = 2.0 * B * N * C * 8 * C
        us, mn = timeit(lambda: ops.geglu(F.linear(x, w, b16)))
        report(f"ff1 library GEMM + geglu N={N} C={C} B={B}", us, mn, flops=flops)
        us, mn = timeit(lambda: ops.ff_geglu(x, w, b32))
        report(f"ff1 fused tcgen05 N={N} C={C} B={B}", us, mn, flops=flops)

if on("lin"):
    import torch.nn.functional as F
    for C, N in sites:
        x = torch.randn(B, N, C, device=dev, dtype=dt)
        h = torch.randn(B, N, 4 * C, device=dev, dtype=dt)
        r = torch.randn(B, N, C, device=dev, dtype=dt)
        for name, inp, nout, kin, use_res in (("qkv", x, 3 * C, C, False), ("proj", x, C, C, False), ("ff2+res", h, C, 4 * C, True)):
            w = torch.randn(nout, kin, device=dev, dtype=dt) * (kin ** -0.5)
            b32 = torch.randn(nout, device=dev)
            b16 = b32.to(dt)
            flops = 2.0 * B * N * kin * nout
            if use_res:
                us, mn = timeit(lambda: ops.bias_residual(F.linear(inp, w, b16), r))
            else:
                us, mn = timeit(lambda: F.linear(inp, w, b16))
            report(f"{name} library N={N} C={C} B={B}", us, mn, flops=flops)
            us, mn = timeit(lambda: ops.linear(inp, w, b32, r if use_res else None, impl="tc"))
            report(f"{name} tcgen05 N={N} C={C} B={B}", us, mn, flops=flops)
            if use_res:      # the form the transformer blocks use: no bias (it rides on the stored residual), beta = 1 accumulate
                yo = torch.empty_like(r)
                us, mn = timeit(lambda: torch.addmm(r.view(-1, nout), inp.view(-1, kin), w.t(), out=yo.view(-1, nout)))
                report(f"{name}(no bias) library beta=1 N={N} C={C} B={B}", us, mn, flops=flops)
                us, mn = timeit(lambda: ops.linear(inp, w, None, r, impl="tc"))
                report(f"{name}(no bias) tcgen05 N={N} C={C} B={B}", us, mn, flops=flops)

if on("add_ln"):
    for C, N in sites:
        x = torch.randn(B, N, C, device=dev, dtype=dt)
        r = torch.randn(B, N, C, device=dev, dtype=dt)
        g = torch.randn(C, device=dev)
        bta = torch.randn(C, device=dev)
        us, mn = timeit(lambda: ops.add_layer_norm(x, r, g, bta, 1e-5))
        report(f"add+layernorm N={N} C={C} B={B}", us, mn, bytes_=8.0 * x.numel())

if on("bias_res"):
    for C, S in [(320, R), (640, R // 2), (1280, R // 4), (1280, R // 8)]:
        a = torch.randn(B, C, S, S, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
        r = torch.randn_like(a)
        bias = torch.randn(C, device=dev)
        y = torch.empty_like(a)
        us, mn = timeit(lambda: ops.bias_residual(a, r, bias, out=y))
        report(f"bias+residual C={C} HW={S}x{S} B={B}", us, mn, bytes_=6.0 * a.numel())

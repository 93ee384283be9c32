"""Top CUDA kernels of one training step (config 3: batch 8, bf16) by device time.  python scripts/profile_train.py"""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_stable_diffusion_b200 as P
from progressive_stable_diffusion_b200 import training as T
dev = torch.device("cuda", 0)
P.set_compute_dtype(torch.float16)
torch.manual_seed(0)
module = P.DiffusionModuleWithIP(P.default_config(), build_image_encoder=True, build_vae_encoder=True).to(dev)
tb = 8
gt = torch.Generator(device=dev).manual_seed(300)
imgs = torch.rand(tb, 3, 256, 256, device=dev, generator=gt) * 2 - 1
struct = torch.randn(tb, 3, 224, 224, device=dev, generator=gt)
labels = torch.randint(0, 4, (tb,), device=dev, generator=gt).float()
trainer = T.DataParallelTrainer(module, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0)
step = lambda: trainer.step(lambda: T.training_step(module, (imgs, labels, struct), generator=gt, compute_dtype=torch.bfloat16))
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
print(f"step: {e0.elapsed_time(e1) / 2:.1f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
tot = collections.Counter(); cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name[:90]
        tot[name] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        cnt[name] += 1
total = sum(tot.values())
print(f"sum of kernel time: {total / 1e3:.1f} ms over {sum(cnt.values())} kernels")
for name, t in tot.most_common(40):
    print(f"{t / 1e3:8.2f} ms {100 * t / total:5.1f} %  x{cnt[name]:4d}  {name}")

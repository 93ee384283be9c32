"""Two (or more) ranks: eager data-parallel training steps, then DataParallelTrainer.capture and replays; prints ms per step.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/ddp_capture_check.py"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_stable_diffusion_b200 as P
from progressive_stable_diffusion_b200 import parallel, training as T
rank, local, world = parallel.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
P.set_compute_dtype(torch.float16)
torch.manual_seed(0)
module = P.DiffusionModuleWithIP(P.default_config(), build_image_encoder=True, build_vae_encoder=True).to(dev)
tb = 8
gt = torch.Generator(device=dev).manual_seed(300 + rank)
imgs = torch.rand(tb, 3, 256, 256, device=dev, generator=gt) * 2 - 1
struct = torch.randn(tb, 3, 224, 224, device=dev, generator=gt)
labels = torch.randint(0, 4, (tb,), device=dev, generator=gt).float()
trainer = T.DataParallelTrainer(module, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0)
fn = lambda: T.training_step(module, (imgs, labels, struct), generator=gt, compute_dtype=torch.bfloat16)


def timed(f, n=3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = f()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n * 1e3, out


for _ in range(2):
    trainer.step(fn)
ms_eager, l0 = timed(lambda: trainer.step(fn))
if rank == 0:
    print(f"world {world}: eager {ms_eager:.1f} ms per step, loss {l0.item():.4f}", flush=True)
replay = trainer.capture(fn, generators=(gt,))
if rank == 0:
    print("captured", flush=True)
replay()
ms_graph, l1 = timed(replay)
w = dict(module.named_parameters())["unet.unet.conv_in.weight"].detach().float()
chk = [torch.zeros_like(w) for _ in range(world)]
dist.all_gather(chk, w)
same = all(torch.equal(chk[0], c) for c in chk)
trainer._skip_allreduce = True
ms_off, _ = timed(replay)
trainer._skip_allreduce = False
if rank == 0:
    print(f"world {world}: graphed {ms_graph:.1f} ms per step (without the all-reduce {ms_off:.1f}), loss {l1.item():.4f}, replicas identical after the graphed steps: {same}", flush=True)
dist.destroy_process_group()

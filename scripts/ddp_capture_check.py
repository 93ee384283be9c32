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


def snapshot():
    return [(bk.flat_p.clone(), bk.m.clone(), bk.v.clone()) for bk in trainer.buckets], trainer.dev_state.clone(), trainer.steps


def restore(snap):
    for bk, (p_, m_, v_) in zip(trainer.buckets, snap[0]):
        bk.flat_p.copy_(p_); bk.m.copy_(m_); bk.v.copy_(v_)
    trainer.dev_state.copy_(snap[1]); trainer.steps = snap[2]


probe = [trainer.buckets[i] for i in (0, len(trainer.buckets) // 3, 2 * len(trainer.buckets) // 3, len(trainer.buckets) - 1)]
grads = []


def run(n):      # fixed generator state -> the same draws in both modes; the reduced gradients of the first step are kept
    gt.manual_seed(1234 + rank)
    out = []
    for i in range(n):
        out.append(replay().item())
        if i == 0:
            grads.append([bk.flat_g.clone() for bk in probe])
    return out


snap = snapshot()
trainer.overlap_reduce = False
la0 = run(3)                      # control: the same mode twice (cuDNN wgrad / reductions are not bit-reproducible run to run)
restore(snap)
la = run(3)
wa = trainer.buckets[0].flat_p.clone()
restore(snap)
trainer.overlap_reduce = True
lb = run(3)
wb = trainer.buckets[0].flat_p.clone()
same_modes = la == lb and torch.equal(wa, wb)
ms_overlap, l1 = timed(replay)
trainer.overlap_reduce = False
ms_after, _ = timed(replay)
trainer.overlap_reduce = True
w = dict(module.named_parameters())["unet.unet.conv_in.weight"].detach().float()
chk = [torch.zeros_like(w) for _ in range(world)]
dist.all_gather(chk, w)
same = all(torch.equal(chk[0], c) for c in chk)
trainer._skip_allreduce = True
ms_off, _ = timed(replay)
trainer._skip_allreduce = False
def gdiff(a, b):
    return max(((x - y).abs().max() / x.abs().max().clamp_min(1e-30)).item() for x, y in zip(a, b))


if rank == 0:
    print(f"reduced gradients of step 1, max |diff| / max |g| over 4 probe buckets: after-graph vs after-graph again {gdiff(grads[0], grads[1]):.3e}, "
          f"after-graph vs behind-events {gdiff(grads[1], grads[2]):.3e}", flush=True)
    print(f"world {world}: graphed {ms_overlap:.1f} ms per step with the all-reduces behind their bucket events, {ms_after:.1f} with the "
          f"all-reduces after the graph, {ms_off:.1f} without them; losses after-graph {la0} / again {la} / behind events {lb}: bit-identical {same_modes}; "
          f"replicas identical {same}", flush=True)
dist.destroy_process_group()

"""Training step on the GPU: backward kernels vs PyTorch autograd of the same op in fp32, the fused loss /
clip / AdamW kernels vs torch, the frozen VAE encoder and the differentiable UNet forward vs the oracle, gradient parity of the
whole loss vs autograd through the ORACLE graph, and a few optimizer steps end to end."""

import math
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

from oracle import conditioning, unet as ounet, weights  # noqa: E402

DEV = "cuda:0"
DT = [torch.bfloat16, torch.float16]


def rel(a, ref):
    return ((a.double().cpu() - ref.double().cpu()).abs().max() / (ref.double().abs().max() + 1e-30)).item()


def _ops():
    from progressive_stable_diffusion_b200 import ops
    return ops


# ------------------------------------------------------------------------------------------------ backward kernels
@pytest.mark.parametrize("dtype", DT, ids=["bf16", "fp16"])
@pytest.mark.parametrize("rows,c", [(2048, 320), (512, 640), (130, 1280), (64, 768), (7, 1024), (1, 8), (300, 2048)])
def test_layernorm_backward(rows, c, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(rows + c)
    x = (torch.randn(rows, c, generator=g) * 1.5 + 0.3).to(dtype)
    dy = torch.randn(rows, c, generator=g).to(dtype)
    gamma, beta = torch.randn(c, generator=g) * 0.5 + 1.0, torch.randn(c, generator=g) * 0.1
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (c,), gr, br, 1e-5).backward(dy.float())
    dx, dg, db = ops.layer_norm_bwd(x.to(DEV), dy.to(DEV), gamma.to(DEV), 1e-5)
    tol = 1.5e-2 if dtype == torch.bfloat16 else 2e-3
    assert rel(dx, xr.grad) <= tol, rel(dx, xr.grad)
    assert rel(dg, gr.grad) <= 1e-4 and rel(db, br.grad) <= 1e-4, (rel(dg, gr.grad), rel(db, br.grad))


@pytest.mark.parametrize("dtype", DT, ids=["bf16", "fp16"])
@pytest.mark.parametrize("m,inner", [(2048, 1280), (300, 2560), (17, 5120), (1, 8)])
def test_geglu_backward(m, inner, dtype):
    ops = _ops()
    g = torch.Generator().manual_seed(m + inner)
    proj = (torch.randn(m, 2 * inner, generator=g) * 1.3).to(dtype)
    dy = torch.randn(m, inner, generator=g).to(dtype)
    pr = proj.float().requires_grad_(True)
    a, gate = pr.chunk(2, dim=-1)
    (a * F.gelu(gate)).backward(dy.float())
    got = ops.geglu_bwd(proj.to(DEV), dy.to(DEV))
    assert rel(got, pr.grad) <= (1.5e-2 if dtype == torch.bfloat16 else 2e-3), rel(got, pr.grad)


GN_SHAPES = [(2, 320, 32, 32), (2, 640, 16, 16), (2, 960, 16, 16), (3, 1280, 8, 8), (2, 2560, 4, 4), (1, 1920, 16, 16),
             (2, 128, 24, 24), (1, 512, 8, 8), (2, 320, 9, 7)]


@pytest.mark.parametrize("dtype", DT, ids=["bf16", "fp16"])
@pytest.mark.parametrize("silu,with_add", [(True, True), (True, False), (False, False)])
@pytest.mark.parametrize("shape", GN_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_groupnorm_backward(shape, silu, with_add, dtype):
    ops = _ops()
    b, c, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.randn(shape, generator=g) * 1.7 + 0.5).to(dtype)
    dy = torch.randn(shape, generator=g).to(dtype)
    gamma, beta = torch.randn(c, generator=g) * 0.5 + 1.0, torch.randn(c, generator=g) * 0.2
    add = torch.randn(b, c, generator=g) * 0.7 if with_add else None
    xr, gr, br = x.float().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ar = add.clone().requires_grad_(True) if with_add else None
    z = F.group_norm(xr + (ar[:, :, None, None] if with_add else 0.0), 32, gr, br, 1e-5)
    (F.silu(z) if silu else z).backward(dy.float())
    cl = lambda t: t.to(DEV).contiguous(memory_format=torch.channels_last)
    dx, dg, db, dadd = ops.group_norm_bwd(cl(x), cl(dy), gamma.to(DEV), beta.to(DEV), 32, 1e-5, silu,
                                          None if add is None else add.to(DEV), need_dchan_add=with_add)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert rel(dx, xr.grad) <= tol, rel(dx, xr.grad)
    assert rel(dg, gr.grad) <= 2e-4 and rel(db, br.grad) <= 2e-4, (rel(dg, gr.grad), rel(db, br.grad))
    if with_add:
        # dchan_add sums dx over the pixels: a difference of large terms, compared on the scale of the summands
        scale = xr.grad.abs().sum(dim=(2, 3)).max().item()
        assert (dadd.cpu() - ar.grad).abs().max().item() <= 2e-4 * scale + 1e-6


@pytest.mark.parametrize("dtype", DT, ids=["bf16", "fp16"])
@pytest.mark.parametrize("n,d,b,nseg", [(1024, 40, 2, 2), (256, 80, 2, 3), (64, 160, 3, 2), (16, 160, 2, 3), (100, 40, 1, 2), (333, 80, 1, 1)])
def test_cross_attention_backward(n, d, b, nseg, dtype):
    """dadd_cross_attn_bwd vs fp32 autograd of the reference arithmetic (routing_gates.py:148-178) on the same 16-bit inputs."""
    ops = _ops()
    h, seg = 8, 16
    c, l = h * d, nseg * seg
    g = torch.Generator().manual_seed(n + d + nseg)
    q = (torch.randn(b, n, c, generator=g) * 1.2).to(dtype)
    k, v = (torch.randn(b, h, l, d, generator=g) * 1.1).to(dtype), torch.randn(b, h, l, d, generator=g).to(dtype)
    do = torch.randn(b, n, c, generator=g).to(dtype)
    gates = torch.tensor([0.9, 0.1, 3.0])[:nseg]
    qq, kk, vv = q.float().requires_grad_(True), k.float().requires_grad_(True), v.float().requires_grad_(True)
    qh = qq.view(b, n, h, d).transpose(1, 2)
    out = sum(gates[s] * torch.softmax(qh @ kk[:, :, s * seg:(s + 1) * seg].transpose(-1, -2) * d ** -0.5, -1) @ vv[:, :, s * seg:(s + 1) * seg]
              for s in range(nseg))
    out.transpose(1, 2).reshape(b, n, c).backward(do.float())
    dq, dk, dv = ops.cross_attention_bwd(q.to(DEV), k.to(DEV), v.to(DEV), gates.to(DEV), do.to(DEV), h, seg, nseg)
    tol = 1.5e-2 if dtype == torch.bfloat16 else 2e-3
    assert rel(dq, qq.grad) <= tol, rel(dq, qq.grad)
    assert rel(dk, kk.grad) <= 1e-3 and rel(dv, vv.grad) <= 1e-3, (rel(dk, kk.grad), rel(dv, vv.grad))
    again = ops.cross_attention_bwd(q.to(DEV), k.to(DEV), v.to(DEV), gates.to(DEV), do.to(DEV), h, seg, nseg)
    assert all(torch.equal(a, b_) for a, b_ in zip((dq, dk, dv), again))


def test_backward_kernels_are_bit_reproducible():
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 640, 16, 16, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(3, 640, 16, 16, generator=g).to(torch.bfloat16).to(DEV).contiguous(memory_format=torch.channels_last)
    gamma, beta, add = torch.randn(640, generator=g).to(DEV), torch.randn(640, generator=g).to(DEV), torch.randn(3, 640, generator=g).to(DEV)
    a = ops.group_norm_bwd(x, dy, gamma, beta, 32, 1e-5, True, add, need_dchan_add=True)
    b = ops.group_norm_bwd(x, dy, gamma, beta, 32, 1e-5, True, add, need_dchan_add=True)
    assert all(torch.equal(p, q) for p, q in zip(a, b))
    t, dt = x.permute(0, 2, 3, 1).reshape(-1, 640), dy.permute(0, 2, 3, 1).reshape(-1, 640)
    a, b = ops.layer_norm_bwd(t, dt, gamma, 1e-5), ops.layer_norm_bwd(t, dt, gamma, 1e-5)
    assert all(torch.equal(p, q) for p, q in zip(a, b))


# ------------------------------------------------------------------------------------------------ loss, clip, AdamW
def test_minsnr_mse_loss_and_gradient():
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    pred, target, w = torch.randn(8, 4, 32, 32, generator=g), torch.randn(8, 4, 32, 32, generator=g), torch.rand(8, generator=g)
    pr = pred.clone().requires_grad_(True)
    want = (w * F.mse_loss(pr, target, reduction="none").mean(dim=(1, 2, 3))).mean()      # diffusion_module_ip.py:449-452
    want.backward()
    loss, grad = ops.minsnr_mse(pred.to(DEV), target.to(DEV), w.to(DEV))
    assert loss.item() == pytest.approx(want.item(), rel=1e-5)
    assert rel(grad, pr.grad) <= 1e-5


def test_fused_clip_and_adamw_match_torch():
    from progressive_stable_diffusion_b200 import training as T
    from tests.test_training_cpu import _Tiny
    torch.manual_seed(0)
    m, ref = _Tiny().to(DEV), _Tiny().to(DEV)
    ref.load_state_dict(m.state_dict())
    tr = T.DataParallelTrainer(m, lr=1e-2, weight_decay=0.05, max_grad_norm=0.5, bucket_bytes=64)       # fused kernels
    used = [p for n, p in ref.named_parameters() if n not in T.UNUSED_PARAMETERS]
    groups = [{"params": [p for p in g["params"] if any(p is q for q in used)], "lr": g["lr"]} for g in T.parameter_groups(ref, 1e-2)]
    opt = torch.optim.AdamW(groups, weight_decay=0.05)
    x, y = torch.randn(16, 6, device=DEV), torch.randn(16, 4, device=DEV)
    for _ in range(5):
        tr.step(lambda: ((m(x) - y) ** 2).mean())
        opt.zero_grad()
        ((ref(x) - y) ** 2).mean().backward()
        norm = torch.nn.utils.clip_grad_norm_(used, 0.5)
        opt.step()
        assert tr.grad_norm.item() == pytest.approx(norm.item(), rel=1e-4)
    for (n, p), (_, q) in zip(m.named_parameters(), ref.named_parameters()):
        torch.testing.assert_close(p, q, rtol=1e-4, atol=1e-6, msg=n)


# ------------------------------------------------------------------------------------------------ model level
@pytest.fixture(scope="module")
def setup():
    import progressive_stable_diffusion_b200 as P
    state = weights.make_module_state(seed=0)
    state.update({"vae.vae." + k: v for k, v in weights.make_vae_encoder_state().items()})
    module = P.DiffusionModuleWithIP(P.default_config(), build_vae_encoder=True)
    module.load_state_dict(state, strict=True)
    module.to(DEV).eval()
    return module, state


def test_ema_update_kernel_matches_torch_ema_avg_fn():
    """dadd_ema_update on a flat bucket vs torch.optim.swa_utils.get_ema_avg_fn (the avg_fn of the reference's EMAWeightAveraging,
    src/callbacks/ema_callback.py:414-436); odd length (scalar tail), first-update copy, and the trainer driving it on the
    callback's schedule."""
    from torch.optim.swa_utils import get_ema_avg_fn
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    n = 4 * 100003 + 3
    avg, p = torch.randn(n, generator=g).to(DEV), torch.randn(n, generator=g).to(DEV)
    want = get_ema_avg_fn(0.999)(avg.clone(), p, 7)
    got = avg.clone()
    ops.ema_update_(got, p, 0.999)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)      # (fused multiply-add vs torch's separate multiply and add)
    ops.ema_update_(got, p, 0.999, first=True)
    assert torch.equal(got, p)


def test_vae_encoder_matches_oracle(setup):
    module, state = setup
    vw = weights.sub_state(state, "vae.vae.")
    g = torch.Generator().manual_seed(4)
    img = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    with torch.no_grad():
        want = ounet.vae_encode_moments(vw, img)
        post = module.vae.encode(img.to(DEV)).latent_dist
    got = torch.cat([post.mean, post.logvar], dim=1)
    e = rel(got, torch.cat([want[:, :4], want[:, 4:].clamp(-30, 20)], 1))
    assert e <= 1e-2, e
    s1 = post.sample(torch.Generator(device=DEV).manual_seed(1))
    assert s1.shape == (2, 4, 8, 8) and torch.isfinite(s1).all()


@pytest.mark.parametrize("dtype", DT, ids=["bf16", "fp16"])
def test_unet_train_forward_and_gradients_match_oracle_autograd(setup, dtype):
    """eps and dLoss/dW of the autograd program over our kernels vs PyTorch autograd through the ORACLE graph (fp32 CPU)."""
    from progressive_stable_diffusion_b200 import training as T
    module, state = setup
    uw = {k: v.clone() for k, v in weights.sub_state(state, "unet.unet.").items()}
    probe = ["conv_in.weight", "time_embedding.linear_1.weight", "down_blocks.0.resnets.0.norm2.weight",
             "down_blocks.0.resnets.0.time_emb_proj.weight", "down_blocks.0.attentions.0.norm.bias",
             "down_blocks.0.attentions.0.transformer_blocks.0.norm1.weight", "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight",
             "down_blocks.1.attentions.1.transformer_blocks.0.attn2.to_k.weight",
             "down_blocks.1.attentions.1.transformer_blocks.0.attn2.processor.to_v_dis.weight",
             "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj.weight", "mid_block.resnets.1.conv2.bias",
             "up_blocks.1.attentions.2.transformer_blocks.0.attn2.to_q.weight", "up_blocks.3.resnets.2.conv_shortcut.weight",
             "up_blocks.3.attentions.2.transformer_blocks.0.norm3.bias", "conv_norm_out.weight", "conv_out.bias"]
    for k in probe:
        uw[k].requires_grad_(True)
    g = torch.Generator().manual_seed(9)
    b = 2
    x, noise = torch.randn(b, 4, 32, 32, generator=g), torch.randn(b, 4, 32, 32, generator=g)
    t = torch.tensor([700, 40])
    cond = torch.randn(b, 48, 768, generator=g) * 0.5
    cond[:, 32:] = 0.0                                                      # training: delta segment is zero, lambda = 0
    wgt = torch.tensor([0.6, 1.0])
    eps_ref = ounet.unet_forward(uw, x, t, cond, ounet.CrossCfg(True, 0.0))
    loss_ref = (wgt * F.mse_loss(eps_ref, noise, reduction="none").mean(dim=(1, 2, 3))).mean()
    loss_ref.backward()

    from progressive_stable_diffusion_b200.inference_pipeline_ip import _set_delta_scale_on_processors
    _set_delta_scale_on_processors(module, 0.0)
    unet = module.unet.unet
    params = dict(unet.named_parameters())
    for p in unet.parameters():
        p.grad = None
    eps = T.unet_forward_train(unet, x.to(DEV), t.to(DEV), cond.to(DEV), dtype)
    scale = 1.0 if dtype == torch.bfloat16 else 8192.0              # fp16 backward needs a loss scale (GradScaler in the reference)
    loss, grad = _ops().minsnr_mse(eps, noise.to(DEV), wgt.to(DEV), upstream=scale)
    eps.backward(grad)
    e = rel(eps, eps_ref.detach())
    print(f"train forward eps rel err {dtype}: {e:.4g}; loss {loss.item():.6f} vs {loss_ref.item():.6f}")
    assert e <= (2e-2 if dtype == torch.bfloat16 else 4e-3), e
    assert loss.item() == pytest.approx(loss_ref.item(), rel=2e-2 if dtype == torch.bfloat16 else 3e-3)
    worst = {}
    for k in probe:
        assert params[k].grad is not None, k
        got, want = params[k].grad / scale, uw[k].grad
        # cosine + norm ratio: per-element errors of a 16-bit backward are noise on small entries, direction and size are not
        cos = F.cosine_similarity(got.flatten().double().cpu(), want.flatten().double(), dim=0).item()
        ratio = (got.double().norm().item() + 1e-30) / (want.double().norm().item() + 1e-30)
        worst[k] = (cos, ratio)
        lo = 0.99 if dtype == torch.bfloat16 else 0.999
        assert cos >= lo and abs(ratio - 1.0) <= (0.05 if dtype == torch.bfloat16 else 0.01), (k, cos, ratio)
    print("gradient parity (cos, norm ratio):", {k.split(".")[-2] + "." + k.split(".")[-1]: (round(c, 5), round(r, 4)) for k, (c, r) in worst.items()})
    # every UNet parameter took part
    assert all(p.grad is not None for p in unet.parameters())


def test_training_loss_end_to_end_and_optimizer_steps(setup):
    """training_loss (conditioning front end + UNet + Min-SNR loss) against the oracle composition, then three optimizer steps on a
    fixed batch: the loss goes down and the three never-used AOE tensors stay untouched."""
    from progressive_stable_diffusion_b200 import training as T
    module, state = setup
    uw, aw, pw = (weights.sub_state(state, p) for p in ("unet.unet.", "ordinal_embedder.", "feature_purifier."))
    g = torch.Generator().manual_seed(21)
    b = 2
    lat, noise = torch.randn(b, 4, 32, 32, generator=g) * 0.18215 * 4, torch.randn(b, 4, 32, 32, generator=g)
    labels, t = torch.tensor([1.0, 2.5]), torch.tensor([850, 120])
    tok = torch.randn(b, 16, 768, generator=g)
    drop = torch.tensor([False, True])
    with torch.no_grad():
        aoe = conditioning.aoe_forward(aw, labels)
        img = conditioning.purifier_forward(pw, tok, aoe)
        img = torch.where(drop.view(-1, 1, 1), torch.zeros_like(img), img)
        cond = torch.cat([aoe, img, torch.zeros_like(aoe)], dim=1)
        ac = module.alphas_cumprod.cpu()[t]
        noisy = ac.sqrt().view(-1, 1, 1, 1) * lat + (1 - ac).sqrt().view(-1, 1, 1, 1) * noise
        eps_ref = ounet.unet_forward(uw, noisy, t, cond, ounet.CrossCfg(True, 0.0))
        snr = ac / (1 - ac + 1e-8)
        wgt = torch.minimum(snr, torch.tensor(1.0)) / (snr + 1e-8)
        want = (wgt * F.mse_loss(eps_ref, noise, reduction="none").mean(dim=(1, 2, 3))).mean().item()
    kw = dict(noise=noise.to(DEV), timesteps=t.to(DEV), aoe_noise_std=0.0, drop_mask=drop.to(DEV), compute_dtype=torch.bfloat16)
    loss, aux = T.training_loss(module, lat.to(DEV), labels.to(DEV), tok.to(DEV), **kw)
    assert loss.item() == pytest.approx(want, rel=2e-2), (loss.item(), want)
    assert aux["cfg_drop_rate"].item() == 0.5

    keep = {n: dict(module.named_parameters())[n].detach().clone() for n in T.UNUSED_PARAMETERS}
    probe = "unet.unet.conv_in.weight"
    trainer = T.DataParallelTrainer(module, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0, ema_decay=0.9, ema_update_every_n_steps=1,
                                    ema_update_starting_at_step=1)
    assert sorted(trainer.unused) == sorted(T.UNUSED_PARAMETERS)
    w0 = dict(module.named_parameters())[probe].detach().clone()
    assert torch.equal(trainer.ema_state_dict()[probe], w0)               # the average model starts as a copy of the module
    losses, seen = [], []
    for _ in range(3):
        losses.append(trainer.step(lambda: T.training_loss(module, lat.to(DEV), labels.to(DEV), tok.to(DEV), **kw)[0]).item())
        seen.append(dict(module.named_parameters())[probe].detach().clone())
    # EMA on the reference callback's schedule (from step index 1, every step): first update copies, second lerps
    want_avg = seen[1] + (seen[2] - seen[1]) * (1 - 0.9)
    torch.testing.assert_close(trainer.ema_state_dict()[probe], want_avg, rtol=1e-6, atol=1e-6)
    assert trainer.ema_updates == 2
    print("losses over 3 steps:", losses, "grad norm", trainer.grad_norm.item())
    assert losses[0] == pytest.approx(want, rel=2e-2) and losses[2] < losses[0] and all(math.isfinite(v) for v in losses)
    for n, v in keep.items():
        assert torch.equal(dict(module.named_parameters())[n], v), n
    # the trained module still samples (inference caches were dropped by the optimizer step)
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _ddim_sample_ip
    out = _ddim_sample_ip(module, torch.tensor([0.0, 3.0]), torch.zeros(2), tok.to(DEV), 2, DEV, steer_scale=3.0)
    assert torch.isfinite(out).all()


def test_graphed_training_step_equals_eager(setup):
    """DataParallelTrainer.capture: one CUDA graph of bucket zeroing + forward + backward + clip + AdamW (step number and learning-
    rate scale on the device) against the eager step on a copy of the module, same injected noise / timesteps / drop mask."""
    import progressive_stable_diffusion_b200 as P
    from progressive_stable_diffusion_b200 import training as T
    _, state = setup

    def fresh():
        m = P.DiffusionModuleWithIP(P.default_config(), build_vae_encoder=True)
        m.load_state_dict(state, strict=True)
        return m.to(DEV).eval()

    g = torch.Generator().manual_seed(33)
    b = 2
    lat, noise = (torch.randn(b, 4, 32, 32, generator=g) * 0.18215 * 4).to(DEV), torch.randn(b, 4, 32, 32, generator=g).to(DEV)
    labels, t = torch.tensor([0.5, 3.0], device=DEV), torch.tensor([700, 60], device=DEV)
    tok = torch.randn(b, 16, 768, generator=g).to(DEV)
    kw = dict(noise=noise, timesteps=t, aoe_noise_std=0.0, drop_mask=torch.tensor([False, False], device=DEV), compute_dtype=torch.bfloat16)
    ma, mb = fresh(), fresh()
    init = {n: p.detach().clone() for n, p in ma.named_parameters()}
    ta = T.DataParallelTrainer(ma, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0, ema_decay=0.9, ema_update_starting_at_step=2)
    tb = T.DataParallelTrainer(mb, lr=1e-5, weight_decay=0.01, max_grad_norm=1.0, ema_decay=0.9, ema_update_starting_at_step=2)
    la = [ta.step(lambda: T.training_loss(ma, lat, labels, tok, **kw)[0]).item() for _ in range(5)]
    replay = tb.capture(lambda: T.training_loss(mb, lat, labels, tok, **kw)[0], warmup=2)
    assert tb.steps == 2
    lb = [replay().item() for _ in range(3)]
    print("eager", la, "graphed (after 2 eager warm-up steps)", lb)
    assert tb.steps == 5 and tb.dev_state[1].item() == 5.0 and tb.ema_updates == ta.ema_updates == 3
    for x, y in zip(la[2:], lb):
        assert y == pytest.approx(x, rel=5e-3), (la, lb)
    assert lb[-1] < la[0]
    pa, pb = dict(ma.named_parameters()), dict(mb.named_parameters())
    for n in ("unet.unet.conv_in.weight", "unet.unet.mid_block.attentions.0.transformer_blocks.0.attn2.to_q.weight", "feature_purifier.gate.0.weight"):
        ua, ub = (pa[n] - init[n]).flatten().double(), (pb[n] - init[n]).flatten().double()      # what five steps did to the tensor
        cos = (ua @ ub / (ua.norm() * ub.norm())).item()
        assert cos >= 0.99, (n, cos)          # (Adam's normalised update flips sign on elements whose bf16 gradient is noise)
    ea, eb = ta.ema_state_dict(), tb.ema_state_dict()
    torch.testing.assert_close(ea["unet.unet.conv_in.weight"], eb["unet.unet.conv_in.weight"], rtol=0, atol=1e-4)

"""Whole-path parity on the GPU: UNet noise prediction, the graphed DDIM progression and decoded-image PSNR vs the oracle.

Gates: eps max relative error <= 2e-2 (16-bit kernels vs fp32 oracle), final decoded images
PSNR >= 40 dB, gate / token indexing bit-exact.

The package's DEFAULT compute dtype - the one bench.py measures - is fp16 (the reference's own mixed-precision dtype,
evaluation_pipeline.py:943) and is held to EVERY gate: eps <= 2e-2 (asserted at 4e-3), PSNR >= 40 dB after 8 and after 50
steps, CFG trajectory <= 3e-2.  bf16 is an opt-in alternative (``set_compute_dtype(torch.bfloat16)``): it meets the eps gate
(the only one BASELINE.md states for bf16 operands) but, with random-init non-contractive weights, the 50-step trajectory
amplifies its 8x larger per-step rounding to 26-30 dB for ANY bf16-operand implementation, stock PyTorch autocast included
(profiles/r01_precision_experiment.txt).  Its trajectory tests therefore check graph == eager and the measured bf16 floor,
and are NOT a claim of the 40 dB gate.
"""

import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import conditioning, sampler, unet as ounet, weights  # noqa: E402

DEV = "cuda:0"
DEFAULT_DTYPE = torch.float16          # == progressive_stable_diffusion_b200.DEFAULT_COMPUTE_DTYPE (asserted below)
TOL_EPS = {torch.bfloat16: 2e-2, torch.float16: 4e-3}
PSNR_MIN = 40.0


def test_default_dtype_is_the_one_held_to_every_gate():
    import progressive_stable_diffusion_b200 as P
    assert P.DEFAULT_COMPUTE_DTYPE == DEFAULT_DTYPE == P.compute_dtype()


def rel_err(a, ref):
    return ((a.double().cpu() - ref.double().cpu()).abs().max() / ref.double().abs().max()).item()


def psnr(a, ref):
    mse = ((a.double().cpu() - ref.double().cpu()) ** 2).mean().item()
    return 99.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


@pytest.fixture(scope="module", params=[1.0, math.sqrt(3.0)], ids=["gain1", "gain_sqrt3"])
def models(request):
    import progressive_stable_diffusion_b200 as P
    gain = request.param
    state = weights.make_module_state(seed=0, gain=gain)
    module = P.DiffusionModuleWithIP(P.default_config())
    module.load_state_dict(state, strict=True)
    module.to(DEV).eval()
    return module, state, gain, {}


@pytest.fixture(params=[torch.bfloat16, torch.float16], ids=["bf16", "fp16"])
def compute(request):
    import progressive_stable_diffusion_b200 as P
    P.set_compute_dtype(request.param)
    yield request.param
    P.set_compute_dtype(P.DEFAULT_COMPUTE_DTYPE)


def _split(state):
    return (weights.sub_state(state, "unet.unet."), weights.sub_state(state, "ordinal_embedder."),
            weights.sub_state(state, "feature_purifier."), weights.sub_state(state, "vae.vae."))


def _inputs(n, seed=1):
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(1, 4, 32, 32, generator=g)
    img_tokens = torch.randn(1, 16, 768, generator=g).expand(n, -1, -1).contiguous()
    return noise, img_tokens


def _memo(cache, key, fn):
    if key not in cache:
        cache[key] = fn()
    return cache[key]


@pytest.mark.parametrize("steer", [3.0, 0.0])
def test_unet_noise_prediction(models, compute, steer):
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _prepare_conditioning, _set_delta_scale_on_processors
    module, state, gain, cache = models
    uw, aw, pw, _ = _split(state)
    n = 3
    noise, img = _inputs(n)
    target = torch.tensor([0.0, 1.25, 3.0])
    source = torch.full((n,), 2.0)
    x = noise.repeat(n, 1, 1, 1) * torch.tensor([1.0, 0.7, 1.3]).view(n, 1, 1, 1)
    t = torch.tensor([999, 500, 20])
    with torch.no_grad():
        cond_ref = conditioning.prepare_conditioning(aw, pw, target, source, img)
        eps_ref = _memo(cache, ("eps", steer), lambda: ounet.unet_forward(uw, x, t, cond_ref, ounet.CrossCfg(True, steer)))
        cond = _prepare_conditioning(module, target.to(DEV), source.to(DEV), img.to(DEV))
        torch.testing.assert_close(cond.cpu(), cond_ref, atol=2e-4, rtol=1e-4)
        _set_delta_scale_on_processors(module, steer)
        eps = module(x.to(DEV), t.to(DEV), cond)
    assert eps.dtype == torch.float32 and eps.shape == eps_ref.shape
    e = rel_err(eps, eps_ref)
    print(f"eps rel err gain={gain:.2f} steer={steer} {compute}: {e:.4g}")
    assert e <= TOL_EPS[compute], e


def test_progression_matches_oracle_and_psnr(models, compute):
    """3 levels x 8 DDIM steps: final latents and decoded images vs the oracle loop (same noise, same weights)."""
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _ddim_sample_ip, _latents_to_images
    module, state, gain, cache = models
    uw, aw, pw, vw = _split(state)
    n, steps = 3, 8
    noise, img = _inputs(n, seed=2)
    target = torch.tensor([0.0, 1.5, 3.0])
    source = torch.zeros(n)
    with torch.no_grad():
        lat_ref, img_ref = _memo(cache, "short", lambda: (lambda l: (l, ounet.latents_to_images(vw, l)))(
            sampler.ddim_sample(uw, aw, pw, target, source, img, noise, sampling_steps=steps, steer_scale=3.0)))
        lat = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, steer_scale=3.0, init_latents=noise)
        lat_eager = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, steer_scale=3.0, init_latents=noise,
                                    use_graph=False)
        p_lat = psnr(ounet.latents_to_images(vw, lat.cpu()), img_ref)       # oracle decoder: isolates the denoising path
        p_full = psnr(_latents_to_images(module, lat), img_ref)            # product decoder: the user-visible images
    assert torch.equal(lat, lat_eager), "graph replay and eager stepping must agree bit for bit"
    print(f"8 steps gain={gain:.2f} {compute}: latent rel err {rel_err(lat, lat_ref):.4g}, PSNR(oracle decoder) {p_lat:.1f} dB, "
          f"PSNR(product decoder) {p_full:.1f} dB")
    assert p_lat >= (PSNR_MIN if compute == DEFAULT_DTYPE else 36.0), p_lat
    assert p_full >= (PSNR_MIN if compute == DEFAULT_DTYPE else 33.0), p_full


def test_full_50_step_progression_psnr(models, compute):
    """The headline schedule (50 DDIM steps, lambda = 3) on 2 levels: decoded-image PSNR vs the oracle."""
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _ddim_sample_ip, _latents_to_images
    module, state, gain, cache = models
    if gain != 1.0:
        pytest.skip("one weight set is enough for the long run")
    uw, aw, pw, vw = _split(state)
    n = 2
    noise, img = _inputs(n, seed=3)
    target = torch.tensor([0.75, 3.0])
    source = torch.ones(n)
    with torch.no_grad():
        lat_ref, img_ref = _memo(cache, "long", lambda: (lambda l: (l, ounet.latents_to_images(vw, l)))(
            sampler.ddim_sample(uw, aw, pw, target, source, img, noise, sampling_steps=50, steer_scale=3.0)))
        lat = _ddim_sample_ip(module, target, source, img.to(DEV), 50, DEV, steer_scale=3.0, init_latents=noise)
        p_lat = psnr(ounet.latents_to_images(vw, lat.cpu()), img_ref)
        p_full = psnr(_latents_to_images(module, lat), img_ref)
    print(f"50 steps {compute}: latent rel err {rel_err(lat, lat_ref):.4g}, PSNR(oracle decoder) {p_lat:.1f} dB, "
          f"PSNR(product decoder) {p_full:.1f} dB")
    if compute == DEFAULT_DTYPE:
        assert p_lat >= PSNR_MIN, p_lat      # the north-star gate, on the dtype the benchmark runs
        assert p_full >= PSNR_MIN, p_full
    else:
        assert p_lat >= 24.0, p_lat          # opt-in bf16: the floor every bf16-operand implementation hits (module docstring)


def test_baseline_mode_with_cfg(compute):
    """use_routing_gates=False: OrdinalIPAttnProcessor2_0 + two UNet passes + fused CFG/DDIM kernel."""
    import progressive_stable_diffusion_b200 as P
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _ddim_sample_ip
    state = weights.make_module_state(seed=4, routing=False, with_vae=False)
    module = P.DiffusionModuleWithIP(P.default_config(use_routing_gates=False), build_vae=False)
    module.load_state_dict(state, strict=True)
    module.to(DEV).eval()
    uw, aw, pw = (weights.sub_state(state, p) for p in ("unet.unet.", "ordinal_embedder.", "feature_purifier."))
    n, steps = 2, 4
    noise, img = _inputs(n, seed=5)
    target, source = torch.tensor([0.5, 2.0]), torch.tensor([1.0, 1.0])
    with torch.no_grad():
        lat_ref = sampler.ddim_sample(uw, aw, pw, target, source, img, noise, sampling_steps=steps, guidance_scale=2.0,
                                      use_routing_gates=False)
        lat = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, guidance_scale=2.0, init_latents=noise)
    e = rel_err(lat, lat_ref)
    print(f"baseline+CFG latent rel err {compute}: {e:.4g}")
    # default dtype: a parity bound; opt-in bf16: 4 CFG steps (guidance 2 doubles the eps error) of 1.2e-2-per-step rounding
    # (the opt-in bf16 bound is not a parity claim: 0.13-0.15 measured, box to box, with cuDNN's benchmark-mode algorithm choice)
    assert e <= (0.03 if compute == DEFAULT_DTYPE else 0.2), e


def test_eta_sampling_uses_reference_rng_order(models):
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _ddim_sample_ip
    module, _, _, _ = models
    n, steps = 2, 3
    _, img = _inputs(n, seed=6)
    target, source = torch.tensor([1.0, 2.0]), torch.zeros(n)
    torch.manual_seed(7)
    a = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, eta=0.5, steer_scale=1.0)
    torch.manual_seed(7)
    b = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, eta=0.5, steer_scale=1.0)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    torch.manual_seed(8)
    c = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, eta=0.5, steer_scale=1.0)
    assert not torch.equal(a, c)


def test_unet_noise_prediction_512(models, compute):
    """BASELINE config 5 (512x512 -> 64x64 latents): N = 4096 self-attention, tcgen05 cross-attention at N = 4096 / 1024, the
    flat GroupNorm passes (samples above the cluster kernel's shared-memory budget) and their two-source form on the up path."""
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _prepare_conditioning, _set_delta_scale_on_processors
    module, state, gain, cache = models
    if gain != 1.0:
        pytest.skip("one weight set is enough for the large-latent run")
    uw, aw, pw, _ = _split(state)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, 4, 64, 64, generator=g)
    img = torch.randn(1, 16, 768, generator=g)
    target, source, t = torch.tensor([2.25]), torch.tensor([0.0]), torch.tensor([700])
    with torch.no_grad():
        cond_ref = conditioning.prepare_conditioning(aw, pw, target, source, img)
        eps_ref = _memo(cache, "eps512", lambda: ounet.unet_forward(uw, x, t, cond_ref, ounet.CrossCfg(True, 3.0)))
        cond = _prepare_conditioning(module, target.to(DEV), source.to(DEV), img.to(DEV))
        _set_delta_scale_on_processors(module, 3.0)
        eps = module(x.to(DEV), t.to(DEV), cond)
    assert eps.shape == (1, 4, 64, 64)
    e = rel_err(eps, eps_ref)
    print(f"eps rel err 512x512 {compute}: {e:.4g}")
    assert e <= TOL_EPS[compute], e


def test_evaluation_sweep_sharded_equals_single_process(models):
    """BASELINE config 4 in miniature: a sorted job list run as one process and as two ranks' shares gives the same latents
    job by job (all initial noise is drawn up front in job order), and padding the last batch does not leak."""
    from progressive_stable_diffusion_b200.evaluation_pipeline import generate_all
    module, _, gain, _ = models
    if gain != 1.0:
        pytest.skip("one weight set is enough")
    g = torch.Generator().manual_seed(12)
    tokens = torch.randn(3, 16, 768, generator=g)
    jobs = sorted((s, float(s % 4), float(tl)) for s in range(3) for tl in (0, 1, 3))[:7]      # 7 jobs, batches of 4: ragged tail
    kw = dict(batch_size=4, sampling_steps=3, steer_scale=2.0, seed=5, decode=False)
    whole = generate_all(module, jobs, tokens, DEV, **kw)
    parts = {}
    for r in range(2):
        parts.update(generate_all(module, jobs, tokens, DEV, rank=r, world_size=2, **kw))
    assert sorted(whole) == sorted(parts) == list(range(len(jobs)))
    for i in whole:
        assert torch.equal(whole[i], parts[i]), i


def test_sample_progressions_from_pixels_encodes_each_patient_once():
    """Public API from CLIP-preprocessed pixels: the front end runs once per patient (the reference encodes 13 identical copies,
    inference_pipeline_ip.py:282-283) and the result equals the call on the tokens it produces."""
    import progressive_stable_diffusion_b200 as P
    from progressive_stable_diffusion_b200.inference_pipeline_ip import sample_progressions
    torch.manual_seed(3)
    module = P.DiffusionModuleWithIP(P.default_config(), build_vae=False, build_image_encoder=True).to(DEV).eval()
    g = torch.Generator().manual_seed(13)
    pixels = torch.randn(2, 3, 224, 224, generator=g)
    noise = torch.randn(2, 4, 32, 32, generator=g)
    src = torch.tensor([0.0, 2.0])
    with torch.no_grad():
        tokens = module._get_image_embeds(pixels.to(DEV))
        a = sample_progressions(module, pixels, src, 3, 2, DEV, init_latents=noise, decode=False)
        b = sample_progressions(module, tokens, src, 3, 2, DEV, init_latents=noise, decode=False)
    assert a.shape == (6, 4, 32, 32) and torch.equal(a, b) and torch.isfinite(a).all()


def test_progression_512_end_to_end():
    """BASELINE config 5 through the public API: 512x512 (64x64 latents) progression with the graph-replayed sampler and the
    VAE decode (GroupNorm flat passes with many chunks, 4096-token self-attention); finite images of the right shape, and the
    graph replay equals eager stepping bit for bit."""
    import progressive_stable_diffusion_b200 as P
    from progressive_stable_diffusion_b200.inference_pipeline_ip import sample_progressions
    cfg = P.default_config()
    cfg.dataset.image_size = 512
    torch.manual_seed(5)
    module = P.DiffusionModuleWithIP(cfg).to(DEV).eval()
    g = torch.Generator().manual_seed(14)
    tokens = torch.randn(1, 16, 768, generator=g)
    noise = torch.randn(1, 4, 64, 64, generator=g)
    src = torch.tensor([1.0])
    with torch.no_grad():
        imgs = sample_progressions(module, tokens, src, 2, 2, DEV, init_latents=noise)
        lat_g = sample_progressions(module, tokens, src, 2, 2, DEV, init_latents=noise, decode=False)
        lat_e = sample_progressions(module, tokens, src, 2, 2, DEV, init_latents=noise, decode=False, use_graph=False)
    assert imgs.shape == (2, 3, 512, 512) and torch.isfinite(imgs).all() and 0.0 <= imgs.min() and imgs.max() <= 1.0
    assert torch.equal(lat_g, lat_e)


def test_steer_scale_sweep_reuses_one_engine(models, compute):
    """evaluation sweeps call the sampler with several steer scales (evaluation_pipeline.py:1274).  All non-zero
    scales share one captured engine; the scale reaches the device through each processor's gate vector, which must be
    refreshed before every replay - the second scale used to produce the first scale's images."""
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _ddim_sample_ip
    module, _, gain, _ = models
    if gain != 1.0:
        pytest.skip("one weight set is enough")
    n, steps = 2, 3
    noise, img = _inputs(n, seed=21)
    target, source = torch.tensor([0.5, 3.0]), torch.zeros(n)
    out = {}
    for steer in (3.0, 1.0, 3.0):
        g = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, steer_scale=steer, init_latents=noise)
        e = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, steer_scale=steer, init_latents=noise, use_graph=False)
        assert torch.equal(g, e), steer
        out.setdefault(steer, g)
        assert torch.equal(out[steer], g)
    assert not torch.equal(out[3.0], out[1.0])
    # an in-place change of the routing gates reaches the replay as well
    proc = next(p for p in module.unet.unet.attn_processors.values() if hasattr(p, "anat_gate"))
    old = proc.anat_gate.clone()
    proc.anat_gate.fill_(0.37)
    try:
        g = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, steer_scale=3.0, init_latents=noise)
        e = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, steer_scale=3.0, init_latents=noise, use_graph=False)
        assert torch.equal(g, e) and not torch.equal(g, out[3.0])
    finally:
        proc.anat_gate.copy_(old)


def test_weight_update_reaches_a_captured_engine(compute):
    """load_state_dict after the first capture (an EMA swap, a new checkpoint) must not leave the graph
    replaying stale derived weights / time-embedding rows."""
    import progressive_stable_diffusion_b200 as P
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _ddim_sample_ip
    sa = weights.make_module_state(seed=7, with_vae=False)
    sb = weights.make_module_state(seed=8, with_vae=False)
    module = P.DiffusionModuleWithIP(P.default_config(), build_vae=False)
    module.load_state_dict(sa, strict=True)
    module.to(DEV).eval()
    n, steps = 2, 3
    noise, img = _inputs(n, seed=22)
    target, source = torch.tensor([0.0, 2.0]), torch.ones(n)
    kw = dict(steer_scale=3.0, init_latents=noise)
    a_graph = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, **kw)
    module.load_state_dict(sb, strict=True)
    b_graph = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, **kw)
    b_eager = _ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, use_graph=False, **kw)
    assert torch.equal(b_graph, b_eager) and not torch.equal(a_graph, b_graph)
    module.load_state_dict(sa, strict=True)
    assert torch.equal(_ddim_sample_ip(module, target, source, img.to(DEV), steps, DEV, **kw), a_graph)


def test_eager_calls_with_recycled_conditioning_addresses(models, compute):
    """two different conditionings of the same shape, the first freed before the second is built (the caching
    allocator then reuses its address), must each get their own K/V projections on the eager path."""
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _prepare_conditioning, _set_delta_scale_on_processors
    module, _, gain, _ = models
    if gain != 1.0:
        pytest.skip("one weight set is enough")
    _, img = _inputs(1, seed=23)
    x = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(24)).to(DEV)
    t = torch.tensor([500], device=DEV)
    _set_delta_scale_on_processors(module, 3.0)
    src = torch.zeros(1, device=DEV)

    def eps_for(label):
        cond = _prepare_conditioning(module, torch.tensor([label], device=DEV), src, img.to(DEV))
        return module(x, t, cond), cond.data_ptr()

    with torch.no_grad():
        e0, p0 = eps_for(0.0)
        e3, p3 = eps_for(3.0)          # `cond` of the first call is dead: same shape, typically the same address
        e0b, _ = eps_for(0.0)
    assert torch.equal(e0, e0b) and not torch.equal(e0, e3), (p0, p3)

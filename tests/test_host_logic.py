"""CPU tests of the host-side mirror of the reference interface: names, role maps, token layout, schedule tables,
state-dict keys, error behaviour.  (Invariants I3-I9 of SURVEY.md section 4.)"""

import os

import numpy as np
import pytest
import torch

import progressive_stable_diffusion_b200 as P
from oracle import sampler, weights
from oracle import processors as oproc
from progressive_stable_diffusion_b200 import parallel
from progressive_stable_diffusion_b200.inference_pipeline_ip import _build_labels, ddim_schedule
from tests.golden import cases

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_modules.npz"))

B1_TIMESTEPS = [999, 978, 958, 937, 917, 897, 876, 856, 835, 815, 795, 774, 754, 733, 713, 693, 672, 652, 632, 611, 591, 570,
                550, 530, 509, 489, 468, 448, 428, 407, 387, 366, 346, 326, 305, 285, 265, 244, 224, 203, 183, 163, 142, 122,
                101, 81, 61, 40, 20, 0]


@pytest.fixture(scope="module")
def module():
    return P.DiffusionModuleWithIP(P.default_config(), build_vae=False)


def test_role_and_frequency_maps_bit_exact_vs_reference():
    names = cases.cross_attention_processor_names()
    assert [P.get_block_type(n) for n in names] == list(GOLD["roles"])
    assert [P.get_frequency_mode_for_block(n) for n in names] == list(GOLD["freq_modes"])
    assert P.get_block_type("conv_in") == "both" and P.get_frequency_mode_for_block("down_blocks.x") == "both"


def test_processor_installation(module):
    unet = module.unet.unet
    procs = unet.attn_processors
    assert len(procs) == 32
    names = list(procs)
    assert names[0] == "down_blocks.0.attentions.0.transformer_blocks.0.attn1.processor"
    for name, proc in procs.items():
        if name.endswith("attn1.processor"):
            assert isinstance(proc, P.AttnProcessor2_0)
            continue
        assert isinstance(proc, P.SplitInjectionAttentionProcessor)
        role = P.get_block_type(name)
        want = {"anatomy": (0.1, 0.9), "disease": (0.9, 0.1)}[role]
        assert proc.block_type == role
        assert (round(proc.anat_gate.item(), 6), round(proc.dis_gate.item(), 6)) == want
        assert (proc.num_aoe_tokens, proc.num_image_tokens, proc.num_delta_tokens) == (16, 16, 16)
    # I5: disease K/V warm-started from the text K/V
    for mod in unet.modules():
        proc = getattr(mod, "processor", None)
        if isinstance(proc, P.SplitInjectionAttentionProcessor):
            assert torch.equal(proc.to_k_dis.weight, mod.to_k.weight) and torch.equal(proc.to_v_dis.weight, mod.to_v.weight)
    with pytest.raises(ValueError):
        unet.set_attn_processor({"x": P.AttnProcessor2_0()})


def test_state_dict_keys_equal_lightning_checkpoint_layout():
    m = P.DiffusionModuleWithIP(P.default_config())
    ref = weights.make_module_state(0)
    sd = m.state_dict()
    assert set(sd) == set(ref)
    assert all(sd[k].shape == ref[k].shape for k in sd)
    assert "alphas_cumprod" not in sd                                     # schedule buffers are non-persistent
    assert "unet.unet.mid_block.attentions.0.transformer_blocks.0.attn2.processor.anat_gate" in sd   # gates persistent


def test_schedule_and_labels(module):
    _, ac = sampler.build_noise_schedule()
    assert torch.equal(module.alphas_cumprod, ac)
    assert abs(ac[0].item() - 0.99915) < 1e-6 and abs(ac[999].item() - 0.0015790) < 1e-6 and abs(ac[978].item() - 0.0020298) < 1e-6
    ts, table = ddim_schedule(ac, 1000, 50, 0.0)
    assert ts.tolist() == B1_TIMESTEPS == sampler.ddim_timesteps().tolist()          # I6
    assert torch.equal(_build_labels(13), torch.arange(13) * 0.25)                    # I7
    for eta in (0.0, 0.5):
        ts, table = ddim_schedule(ac, 1000, 50, eta)
        coefs = sampler.ddim_coefficients(ac, ts, eta)
        for i, c in enumerate(coefs):
            assert table[i, 0].item() == c["sqrt_ab"] and table[i, 1].item() == c["sqrt_1mab"]
            assert bool(table[i, 5].item()) == c["last"]
            if not c["last"]:
                assert table[i, 2].item() == c["sqrt_abp"] and table[i, 3].item() == c["eps_coef"] and table[i, 4].item() == c["sigma"]
    with pytest.raises(ValueError):
        _build_labels(0)


def test_unet_wrapper_errors(module):
    with pytest.raises(ValueError):
        module.unet(torch.zeros(1, 4, 8, 8), torch.zeros(1, dtype=torch.long), torch.zeros(1, 2, 3, 768))
    with pytest.raises(RuntimeError, match="build_image_encoder"):      # built without the CLIP front end
        module._get_image_embeds(torch.zeros(1, 3, 224, 224))
    with pytest.raises(NotImplementedError):
        bad = P.default_config()
        bad.diffusion.noise_schedule = "cosine"
        P.DiffusionModuleWithIP(bad, build_vae=False)


def test_set_delta_scale(module):
    from progressive_stable_diffusion_b200.inference_pipeline_ip import _set_delta_scale_on_processors
    _set_delta_scale_on_processors(module, 3.0)
    vals = [p.delta_scale for p in module.unet.unet.attn_processors.values() if hasattr(p, "delta_scale")]
    assert vals == [3.0] * 16


def test_no_cpu_fallback(module):
    from progressive_stable_diffusion_b200._lib import DaddError
    with pytest.raises(DaddError):
        module.ordinal_embedder(torch.tensor([1.0]))


def test_shard_indices_cover_and_balance():
    for n in (0, 1, 13, 104, 600):
        for w in (1, 2, 4, 8):
            parts = [parallel.shard_indices(n, r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    with pytest.raises(ValueError):
        parallel.shard_indices(4, 4, 4)


def test_oracle_role_map_matches_product():
    for n in cases.cross_attention_processor_names():
        assert weights.role_of(n) == P.get_block_type(n)
        assert oproc.frequency_mode_of(n) == P.get_frequency_mode_for_block(n)


def test_front_end_state_dict_keys_equal_transformers_and_reference_layout():
    """ImageEncoder's keys are transformers' CLIPVisionModelWithProjection keys (minus the non-persistent position_ids buffer
    of older versions), ImageProjectionPlus's are the reference module's (tests/golden/cases.py states load strictly)."""
    from tests.golden import cases
    enc = P.ImageEncoder("openai/clip-vit-large-patch14")
    assert enc.hidden_size == 1024 and enc.projection_dim == 768
    want = weights.make_clip_vision_state(seed=1, layers=24)
    assert set(enc.image_encoder.state_dict()) == set(want)
    assert all(enc.image_encoder.state_dict()[k].shape == v.shape for k, v in want.items())
    plus = P.ImageProjectionPlus(clip_hidden_dim=1024, cross_attention_dim=768, num_tokens=16)
    plus.load_state_dict(cases.projection_plus_inputs(), strict=True)
    with pytest.raises(ValueError):
        P.ImageEncoder("openai/some-other-tower")


def test_load_from_checkpoint_reads_the_ema_layout(tmp_path, module):
    """A Lightning checkpoint written under the reference's EMA callback (src/callbacks/ema_callback.py:316-329) holds the
    AVERAGED weights in ``state_dict`` and the raw ones in ``current_model_state``: the loader must take the former, the way
    inference_pipeline_ip.py:587-592 does (strict=False), and leave absent keys at their initial values."""
    key_u = "unet.unet.mid_block.attentions.0.transformer_blocks.0.attn2.processor.to_k_dis.weight"
    key_p = "feature_purifier.norm_out.weight"
    ref = module.state_dict()
    ema = {key_u: torch.full_like(ref[key_u], 0.25), key_p: torch.full_like(ref[key_p], 1.5)}
    raw = {key_u: torch.zeros_like(ref[key_u]), key_p: torch.zeros_like(ref[key_p])}
    path = tmp_path / "last.ckpt"
    torch.save({"state_dict": ema, "current_model_state": raw, "averaging_state": {"n_averaged": torch.tensor(7)},
                "epoch": 3, "global_step": 1234}, path)
    loaded = P.DiffusionModuleWithIP.load_from_checkpoint(str(path), cfg=P.default_config(), weights_only=False, strict=False,
                                                         build_vae=False)
    sd = loaded.state_dict()
    assert torch.equal(sd[key_u], ema[key_u]) and torch.equal(sd[key_p], ema[key_p])
    assert loaded.image_encoder is None            # no image_encoder.* keys in this checkpoint -> front end not built
    assert set(sd) == set(module.state_dict())


def test_q_sample_and_min_snr_weight(module):
    """Host-side pieces of the training path (reference diffusion_module_ip.py:289-313) against their closed forms."""
    g = torch.Generator().manual_seed(21)
    x0, noise = torch.randn(3, 4, 8, 8, generator=g), torch.randn(3, 4, 8, 8, generator=g)
    t = torch.tensor([0, 500, 999])
    ab = module.alphas_cumprod[t].view(-1, 1, 1, 1)
    torch.testing.assert_close(module._q_sample(x0, t, noise), ab.sqrt() * x0 + (1 - ab).sqrt() * noise)
    snr = module.alphas_cumprod[t] / (1 - module.alphas_cumprod[t] + 1e-8)
    torch.testing.assert_close(module._min_snr_weight(t), torch.clamp(snr, max=module.diff_cfg.min_snr_gamma) / (snr + 1e-8))
    assert module._min_snr_weight(t)[0] < 1e-2 and abs(module._min_snr_weight(t)[2].item() - 1.0) < 1e-4   # high SNR clipped, low SNR -> snr / (snr + 1e-8) ~ 1
    ts = module._sample_timesteps(64)
    assert ts.dtype == torch.long and ts.shape == (64,) and 0 <= int(ts.min()) and int(ts.max()) < 1000


def test_cond_cache_is_tied_to_the_tensor_object_not_its_address():
    """an address-keyed K/V cache served a previous conditioning's projections to a new tensor that the allocator
    placed at the same address.  Entries now follow the tensor object: they die with it, rebuild IN PLACE when it (or a
    weight) is modified, transient conditionings are bounded by a small LRU and pinned (engine) ones are never evicted."""
    import gc
    from progressive_stable_diffusion_b200 import wcache
    cache = wcache.CondCache(max_transient=2)
    w = torch.ones(4, 4)
    calls = []

    def build_for(c):
        def build():
            calls.append(1)
            return (c @ w, c * 2.0)
        return build

    a = torch.full((2, 4), 1.0)
    k1, v1 = cache.get(a, (True,), (w,), build_for(a))
    k1b, _ = cache.get(a, (True,), (w,), build_for(a))
    assert k1b is k1 and len(calls) == 1                       # hit
    a.add_(1.0)                                                # in-place update of the conditioning -> same storage, new values
    k2, v2 = cache.get(a, (True,), (w,), build_for(a))
    assert k2 is k1 and len(calls) == 2 and torch.equal(k2, a @ w)
    w.mul_(2.0)                                                # weight update -> rebuilt in place too
    k3, _ = cache.get(a, (True,), (w,), build_for(a))
    assert k3 is k1 and len(calls) == 3 and torch.equal(k3, a @ w)
    ptr = a.data_ptr()
    del a, k1, k1b, k2, k3, v1, v2
    gc.collect()
    assert len(cache.entries) == 0                             # the entry died with its tensor
    b = torch.full((2, 4), 5.0)                                # may or may not land on `ptr`: either way it must be a miss
    kb, _ = cache.get(b, (True,), (w,), build_for(b))
    assert len(calls) == 4 and torch.equal(kb, b @ w), ptr
    # LRU of transient tensors; a pinned tensor's entry survives any number of them
    pinned = torch.zeros(2, 4)
    wcache.pin(pinned)
    kp, _ = cache.get(pinned, (True,), (w,), build_for(pinned))
    keep = [torch.full((2, 4), float(i)) for i in range(5)]
    for t in keep:
        cache.get(t, (True,), (w,), build_for(t))
    alive = {k[0] for k in cache.entries}
    assert id(pinned) in alive and len(cache.entries) == 3 and {id(keep[-1]), id(keep[-2])} <= alive
    assert cache.get(pinned, (True,), (w,), build_for(pinned))[0] is kp


def test_binding_refuses_a_library_of_another_abi(monkeypatch):
    from progressive_stable_diffusion_b200 import _lib
    _lib.load()
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "ABI_VERSION", _lib.ABI_VERSION + 1)
    with pytest.raises(_lib.DaddError, match="ABI version"):
        _lib.load()


def test_i12_progression_shares_one_noise_tensor_and_the_sweep_draws_independent_noise(module, monkeypatch):
    """I12: all MES levels of a progression start from ONE noise tensor (inference_pipeline_ip.py:377-385), the evaluation sweep
    draws independent noise per job from one seeded stream (evaluation_pipeline.py:506,909) - and a sharded sweep hands every job
    the noise the single-process run gives it."""
    from progressive_stable_diffusion_b200 import evaluation_pipeline as ev, inference_pipeline_ip as ip
    seen = {}

    def fake_sample(mod, target, source, images, latents, *a, **k):
        seen["latents"] = latents
        return latents

    monkeypatch.setattr(ip, "_sample", fake_sample)
    noise = torch.randn(1, 4, 32, 32)
    tgt = ip._build_labels(13, 0.0, 3.0, "cpu")
    ip._ddim_sample_ip(module, tgt, torch.zeros(13), torch.zeros(13, 16, 768), 50, "cpu", init_latents=noise)
    lat = seen["latents"]
    assert lat.shape == (13, 4, 32, 32) and all(torch.equal(lat[i], noise[0]) for i in range(13))

    monkeypatch.setattr(ev, "_ddim_sample_batched", lambda mod, tgt, src, tok, steps, dev, eta, isc, ssc, gsc, init_latents=None: init_latents)
    jobs = [(i % 5, float(i % 4), float((i + 1) % 4)) for i in range(29)]
    tokens = torch.zeros(5, 16, 768)
    single = ev.generate_all(module, jobs, tokens, "cpu", batch_size=4, sampling_steps=3, seed=7, decode=False)
    assert sorted(single) == list(range(29))
    assert len({single[i].flatten()[:8].numpy().tobytes() for i in single}) == 29          # independent draws
    merged = {}
    for rank in range(3):
        merged.update(ev.generate_all(module, jobs, tokens, "cpu", batch_size=4, sampling_steps=3, seed=7, rank=rank, world_size=3,
                                      decode=False))
    assert sorted(merged) == list(range(29)) and all(torch.equal(merged[i], single[i]) for i in single)


def test_evaluation_sweep_jobs_follow_the_reference_order(tmp_path):
    """``_collect_jobs`` / ``sweep_order`` / ``group_by_target``: the reference's job list (evaluation_pipeline.py:843-864), its
    walk order (:897-903: sorted by (str(path), target)) and its per-class result layout (:955-975)."""
    from progressive_stable_diffusion_b200 import evaluation_pipeline as ev
    roots = [tmp_path / "a", tmp_path / "b"]
    names = {"a": {0: ["z.png", "b.PNG", "notes.txt"], 2: ["m.jpg"]}, "b": {1: ["k.tif", "c.bmp", "d.png"], 3: []}}
    for r, classes in names.items():
        for cls, files in classes.items():
            (tmp_path / r / str(cls)).mkdir(parents=True)
            for f in files:
                (tmp_path / r / str(cls) / f).write_bytes(b"")
    jobs = ev._collect_jobs(roots, max_per_class=2)
    # root a: class 0 -> b.PNG, z.png (sorted, the .txt is skipped); class 2 -> m.jpg; root b: class 1 -> c.bmp, d.png (k.tif cut by the cap)
    assert [(j.source_path.name, j.source_label, j.target_label) for j in jobs] == [
        ("b.PNG", 0, 1), ("b.PNG", 0, 2), ("b.PNG", 0, 3), ("z.png", 0, 1), ("z.png", 0, 2), ("z.png", 0, 3),
        ("m.jpg", 2, 0), ("m.jpg", 2, 1), ("m.jpg", 2, 3), ("c.bmp", 1, 0), ("c.bmp", 1, 2), ("c.bmp", 1, 3),
        ("d.png", 1, 0), ("d.png", 1, 2), ("d.png", 1, 3)]
    jobs_sorted, sources = ev.sweep_order(list(reversed(jobs)))
    assert jobs_sorted == sorted(jobs, key=lambda j: (str(j.source_path), j.target_label))
    assert [q.name for q in sources] == ["b.PNG", "z.png", "m.jpg", "c.bmp", "d.png"] and len(set(sources)) == 5
    idx = ev.as_index_jobs(jobs_sorted, sources)
    assert idx[:4] == [(0, 0.0, 1.0), (0, 0.0, 2.0), (0, 0.0, 3.0), (1, 0.0, 1.0)] and len(idx) == 15
    images = {i: torch.full((3, 8, 8), float(i)) for i in range(15)}
    by_cls = ev.group_by_target(images, jobs_sorted, 8)
    assert [by_cls[c].shape[0] for c in ev.ALL_MES_CLASSES] == [3, 3, 4, 5]
    assert by_cls[0][:, 0, 0, 0].tolist() == [float(i) for i, j in enumerate(jobs_sorted) if j.target_label == 0]
    assert ev.group_by_target({}, jobs_sorted, 8)[2].shape == (0, 3, 8, 8)

"""Invariants the reference code states or implies, checked on the oracle (CPU)."""

import torch

from oracle import conditioning, processors, sampler, unet as ounet, weights
from tests.golden import cases


def test_i2_delta_zero_equals_two_pathways():
    case = cases.PROCESSOR_CASES[0]
    w, x, ehs = cases.processor_inputs(case)
    with torch.no_grad():
        full = processors.split_injection_attention(w, x, ehs, 0.0)
        ehs2 = ehs.clone()
        ehs2[:, -16:] = 123.0
        assert torch.equal(full, processors.split_injection_attention(w, x, ehs2, 0.0))


def test_i11_fused_form_equals_three_pathways():
    """sum_i g_i softmax(Q K_i^T) V_i == [g_i P_i]_i @ V_cat: the identity the fused kernel relies on."""
    g = torch.Generator().manual_seed(0)
    q = torch.randn(2, 8, 64, 40, generator=g)
    ks = [torch.randn(2, 8, 16, 40, generator=g) for _ in range(3)]
    vs = [torch.randn(2, 8, 16, 40, generator=g) for _ in range(3)]
    gates = [0.9, 0.1, 3.0]
    ref = sum(gt * torch.softmax(q @ k.transpose(-1, -2) / 40 ** 0.5, -1) @ v for gt, k, v in zip(gates, ks, vs))
    p = torch.cat([gt * torch.softmax(q @ k.transpose(-1, -2) / 40 ** 0.5, -1) for gt, k in zip(gates, ks)], -1)
    torch.testing.assert_close(p @ torch.cat(vs, 2), ref, atol=2e-6, rtol=1e-5)


def test_i10_baseline_frequency_reweighting_is_plain_softmax_attention():
    """I10 (attention_processor_base.py:29-37,103-116): every frequency mode multiplies the probabilities by all-ones scales and
    renormalises - the processor equals plain softmax attention up to that one extra rounding."""
    for case in (c for c in cases.PROCESSOR_CASES if c["kind"] == "base"):
        w, x, ehs = cases.processor_inputs(case)
        with torch.no_grad():
            both = processors.ordinal_ip_attention(w, x, ehs, "both")
            for mode in ("aoe_dominant", "image_dominant", "both"):
                got = processors.ordinal_ip_attention(w, x, ehs, mode)
                torch.testing.assert_close(got, both, atol=2e-6 * both.abs().max().item(), rtol=0)


def test_i8_i9_label_handling():
    w = weights.make_aoe_state(seed=23)
    lab = torch.tensor([-1.0, 0.0, 1.0, 2.5, 3.0, 7.0])
    e = conditioning.aoe_interp(w, lab)
    t = conditioning.aoe_class_table(w)
    assert torch.equal(e[0], t[0]) and torch.equal(e[1], t[0]) and torch.equal(e[2], t[1])
    assert torch.equal(e[4], t[3]) and torch.equal(e[5], t[3])
    torch.testing.assert_close(e[3], 0.5 * t[2] + 0.5 * t[3])
    neg = conditioning.aoe_negative(w, torch.tensor([0.0, 0.25, 1.0, 3.0]))
    pos = conditioning.aoe_forward(w, torch.tensor([1.0, 0.75, 0.0, 0.0]))
    assert torch.equal(neg, pos)


def test_conditioning_token_layout():
    aw, pw = weights.make_aoe_state(3), weights.make_purifier_state(2)
    g = torch.Generator().manual_seed(1)
    img = torch.randn(2, 16, 768, generator=g)
    tgt, src = torch.tensor([0.5, 3.0]), torch.tensor([1.0, 1.0])
    with torch.no_grad():
        c = conditioning.prepare_conditioning(aw, pw, tgt, src, img)
        assert c.shape == (2, 48, 768)
        assert torch.equal(c[:, :16], conditioning.aoe_forward(aw, src))                                  # I3 dis
        assert torch.equal(c[:, 16:32], conditioning.purifier_forward(pw, img, conditioning.aoe_forward(aw, src)))
        assert torch.equal(c[:, -16:], conditioning.aoe_delta(aw, src, tgt))
        c2 = conditioning.prepare_conditioning(aw, pw, tgt, src, img, use_routing_gates=False, zero_aoe=True)
        assert c2.shape == (2, 32, 768) and torch.equal(c2[:, :16], conditioning.aoe_negative(aw, tgt))


def test_sampler_last_step_returns_clamped_x0_and_shapes():
    _, ac = sampler.build_noise_schedule()
    x = torch.randn(2, 4, 8, 8) * 10
    eps = torch.randn(2, 4, 8, 8)
    out = sampler.ddim_update(x, eps, ac, 0, None)
    assert out.abs().max() <= 4.0
    shapes = ounet.attention_shapes(32)
    assert [(c, n, d) for _, c, n, d in shapes].count((320, 1024, 40)) == 5
    assert [(c, n, d) for _, c, n, d in shapes].count((1280, 16, 160)) == 1


# Published facts about the un-vendored dependency's checkpoint (CompVis/stable-diffusion-v1-4, diffusers layout) that any
# restatement of its graph must reproduce exactly: the SD-1.x UNet2DConditionModel has 859 520 964 parameters in 686 tensors
# (SURVEY.md section 2.1 quotes 859.5 M), the AutoencoderKL decoder 49 490 179 in 138 (+ post_quant_conv: 4*4 + 4 = 20 in 2).
SD1X_UNET_PARAMS, SD1X_UNET_TENSORS = 859_520_964, 686
SD1X_VAE_DECODER_PARAMS, SD1X_VAE_DECODER_TENSORS = 49_490_179, 138
DADD_PROCESSOR_PARAMS = 2 * 768 * (5 * 320 + 5 * 640 + 6 * 1280)      # to_k_dis + to_v_dis of the 16 cross-attention sites


def _count(state, prefix, exclude=()):
    items = [(k, v) for k, v in state.items() if k.startswith(prefix) and not any(e in k for e in exclude)]
    return sum(v.numel() for _, v in items), len(items)


def test_unpinned_oracle_graphs_have_the_public_sd1x_shape():
    """The UNet / VAE-decoder oracle restates an absent dependency (parity unpinned, oracle/unet.py header).  This pins its
    SHAPE to the public checkpoint: exact parameter and tensor counts, the A.6 key families, and the same for the product
    module (whose state dict must be loadable from a diffusers-layout checkpoint key for key)."""
    import progressive_stable_diffusion_b200 as P
    state = weights.make_module_state(seed=0)
    assert _count(state, "unet.unet.", exclude=("processor",)) == (SD1X_UNET_PARAMS, SD1X_UNET_TENSORS)
    proc_params, proc_tensors = _count({k: v for k, v in state.items() if "processor" in k}, "unet.unet.")
    assert proc_tensors == 16 * 4 and proc_params == DADD_PROCESSOR_PARAMS + 16 * 2        # + anat_gate / dis_gate scalars
    assert _count(state, "vae.vae.decoder.") == (SD1X_VAE_DECODER_PARAMS, SD1X_VAE_DECODER_TENSORS)
    assert _count(state, "vae.vae.post_quant_conv.") == (20, 2)
    module = P.DiffusionModuleWithIP(P.default_config())
    sd = module.state_dict()
    for prefix in ("unet.unet.", "vae.vae.", "ordinal_embedder.", "feature_purifier."):
        want = {k: tuple(v.shape) for k, v in state.items() if k.startswith(prefix)}
        got = {k: tuple(v.shape) for k, v in sd.items() if k.startswith(prefix)}
        assert got == want, prefix
    # key families of SURVEY.md A.6 and the per-site shapes of Appendix B.3
    unet_keys = [k[len("unet.unet."):] for k in state if k.startswith("unet.unet.")]
    fam = lambda s: sum(1 for k in unet_keys if s in k)
    assert fam("time_emb_proj.weight") == 22 and fam("conv_shortcut.weight") == 14 and fam(".attn1.to_q.weight") == 16
    assert fam("ff.net.0.proj.weight") == 16 and fam("downsamplers.0.conv.weight") == 3 and fam("upsamplers.0.conv.weight") == 3
    for site, c in (("down_blocks.0.attentions.0", 320), ("down_blocks.1.attentions.1", 640), ("down_blocks.2.attentions.0", 1280),
                    ("mid_block.attentions.0", 1280), ("up_blocks.1.attentions.2", 1280), ("up_blocks.2.attentions.0", 640),
                    ("up_blocks.3.attentions.2", 320)):
        base = f"unet.unet.{site}.transformer_blocks.0."
        assert state[base + "attn2.to_k.weight"].shape == (c, 768) == state[base + "attn2.processor.to_v_dis.weight"].shape
        assert state[base + "ff.net.0.proj.weight"].shape == (8 * c, c) and state[base + "ff.net.2.weight"].shape == (c, 4 * c)

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

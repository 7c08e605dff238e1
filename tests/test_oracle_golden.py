"""Pins oracle/studiogan_oracle.py to the reference: every value here was produced by the real StudioGAN code
(tests/golden/make_golden.py, run in the build container) and must be reproduced by the oracle on CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import studiogan_oracle as O

CASES = [("deep32_c8", dict(img_size=32, conv_dim=8, depth=1, attn=False)),
         ("deep32_c16_attn_d2", dict(img_size=32, conv_dim=16, depth=2, attn=True)),
         ("deep32_c8_b16", dict(img_size=32, conv_dim=8, depth=1, attn=False)),
         ("deepsg32_c8", dict(img_size=32, conv_dim=8, depth=1, attn=False, studiogan=True))]


def load_sd(npz, prefix, grad=False):
    """State dict ``prefix``; the G1/ and D1/ snapshots hold buffers only and take their parameters from G0/ and D0/."""
    sd = {}
    if prefix in ("G1/", "D1/"):
        sd = load_sd(npz, prefix[0] + "0/", grad)
    for k in npz.files:
        if k.startswith(prefix):
            t = torch.from_numpy(npz[k].copy())
            if grad and t.is_floating_point() and (k.endswith("weight_orig") or k.endswith("weight") or k.endswith("bias")
                                                   or k.endswith("sigma")) and "running" not in k:
                t.requires_grad_(True)
            sd[k[len(prefix):]] = t
    return sd


@pytest.mark.parametrize("tag,cfg", CASES)
def test_deep_d_phase_and_g_phase(golden_dir, tag, cfg):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    z, yf = torch.from_numpy(g["z"]), torch.from_numpy(g["y_fake"])
    real, yr = torch.from_numpy(g["real"]), torch.from_numpy(g["y_real"])
    kw_g = dict(img_size=cfg["img_size"], g_conv_dim=cfg["conv_dim"], g_depth=cfg["depth"], attn_g_loc=(2,), apply_attn=cfg["attn"])
    kw_d = dict(img_size=cfg["img_size"], d_conv_dim=cfg["conv_dim"], d_depth=cfg["depth"], attn_d_loc=(1,), apply_attn=cfg["attn"],
                studiogan=cfg.get("studiogan", False))

    # discriminator phase
    sdG = load_sd(g, "G0/")
    sdD = load_sd(g, "D0/", grad=True)
    with torch.no_grad():
        fake = O.deep_generator(sdG, z, yf, **kw_g)
    np.testing.assert_allclose(fake.numpy(), g["fake"], rtol=1e-4, atol=2e-5)
    adv_r, h_r = O.deep_discriminator(sdD, real, yr, **kw_d)
    adv_f, _ = O.deep_discriminator(sdD, fake, yf, **kw_d)
    np.testing.assert_allclose(adv_r.detach().numpy(), g["adv_real"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(adv_f.detach().numpy(), g["adv_fake"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(h_r.detach().numpy(), g["h_real"], rtol=1e-4, atol=1e-4)
    loss = O.d_hinge(adv_r, adv_f)
    np.testing.assert_allclose(loss.item(), g["d_loss"], rtol=1e-5)
    loss.backward()
    for k in g.files:
        if k.startswith("Dgrad/"):
            name = k[len("Dgrad/"):]
            ref = g[k]
            got = sdD[name].grad.numpy()
            assert np.abs(got - ref).max() <= 2e-3 * (1e-3 + np.abs(ref).max()), name  # ReLU-boundary flips dominate the fp32 noise
    # buffers after the forwards (u, v, running statistics)
    for k in g.files:
        if k.startswith("G1/") and ("weight_u" in k or "running_" in k):
            np.testing.assert_allclose(sdG[k[3:]].numpy(), g[k], rtol=1e-4, atol=1e-5, err_msg=k)
        if k.startswith("D1/") and "weight_u" in k:
            np.testing.assert_allclose(sdD[k[3:]].numpy(), g[k], rtol=1e-4, atol=1e-5, err_msg=k)

    # generator phase (continues from the buffers left by the discriminator phase)
    sdG2 = load_sd(g, "G1/", grad=True)
    sdD2 = load_sd(g, "D1/")
    fake2 = O.deep_generator(sdG2, z, yf, **kw_g)
    np.testing.assert_allclose(fake2.detach().numpy(), g["fake2"], rtol=1e-4, atol=2e-5)
    adv2, _ = O.deep_discriminator(sdD2, fake2, yf, **kw_d)
    gl = O.g_hinge(adv2)
    np.testing.assert_allclose(gl.item(), g["g_loss"], rtol=1e-5)
    gl.backward()
    for k in g.files:
        if k.startswith("Ggrad/"):
            name = k[len("Ggrad/"):]
            ref = g[k]
            got = sdG2[name].grad.numpy()
            assert np.abs(got - ref).max() <= 2e-3 * (1e-3 + np.abs(ref).max()), name  # ReLU-boundary flips dominate the fp32 noise


def test_metrics_and_losses(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    assert np.array_equal(O.quantize_images(torch.from_numpy(g["q_in"])), g["q_out"])            # integer path: bit exact
    rs = O.resize_legacy(g["q_out"], 19).numpy()
    np.testing.assert_allclose(rs, g["resize_legacy_19"].transpose(0, 3, 1, 2), rtol=0, atol=1e-4)
    # "friendly" = PIL 'F'-mode bilinear (reference golden): up-scaling 8 -> 19 and anti-aliased down-scaling 20 -> 13.
    # float64 sums rounded to float32: the restatement may differ from Pillow's C loop by the summation order only (1 ulp)
    fr = O.resize_friendly(g["q_out"], 19).numpy()
    np.testing.assert_allclose(fr, g["resize_friendly_19"].transpose(0, 3, 1, 2), rtol=2e-7, atol=2e-5)
    q20 = O.quantize_images(torch.from_numpy(g["q20_in"]))
    fr13 = O.resize_friendly(q20, 13).numpy()
    np.testing.assert_allclose(fr13, g["resize_friendly_20to13"].transpose(0, 3, 1, 2), rtol=2e-7, atol=2e-5)
    # and directly against Pillow at the evaluation size (32 -> 299, 300 -> 299)
    from PIL import Image
    rs_ = np.random.RandomState(9)
    for S in (32, 300):
        ch = rs_.randint(0, 256, size=(S, S)).astype(np.float32)
        ref = np.asarray(Image.fromarray(ch, mode="F").resize((299, 299), resample=Image.BILINEAR))
        got = O.resize_friendly(ch[None, None], 299).numpy()[0, 0]
        np.testing.assert_allclose(got, ref, rtol=2e-7, atol=2e-5)
    mu1, s1 = O.moments(g["feat_fake"])
    mu2, s2 = O.moments(g["feat_real"])
    np.testing.assert_allclose(O.frechet_distance(mu1, s1, mu2, s2), g["fid"], rtol=1e-9)
    m, s = O.calculate_kl_div(g["probs"], 1)
    np.testing.assert_allclose(m, g["is_1"][0], rtol=1e-6)
    assert np.isnan(g["is_1"][1])                                                               # reference quirk: splits=1 -> NaN std
    m4, s4 = O.calculate_kl_div(g["probs"], 4)
    np.testing.assert_allclose([m4, s4], g["is_4"], rtol=1e-5)
    pr = O.prdc(g["feat_real"][:200], g["feat_fake"][:180].astype(np.float64), 5)
    np.testing.assert_allclose([pr["precision"], pr["recall"], pr["density"], pr["coverage"]], g["prdc"], rtol=1e-12)
    a, b = torch.from_numpy(g["loss_in_real"]), torch.from_numpy(g["loss_in_fake"])
    got = [O.d_hinge(a, b), O.g_hinge(b), O.d_wasserstein(a, b), O.g_wasserstein(b), O.d_vanilla(a, b), O.g_vanilla(b)]
    np.testing.assert_allclose([float(x) for x in got], g["losses"], rtol=1e-6)


def test_ema(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    keys = [k[len("ema_src/"):] for k in g.files if k.startswith("ema_src/")]
    src0 = {k: torch.from_numpy(g["ema_src/" + k].copy()) for k in keys}
    # Ema.__init__ copies source -> target (src/utils/ema.py:21-25)
    for k in keys:
        if "num_batches" not in k:
            np.testing.assert_allclose(g["ema_init/" + k], src0[k].numpy())
    src2 = {k: torch.from_numpy(g["ema_src2/" + k].copy()) for k in keys}
    tgt = {k: torch.from_numpy(g["ema_init/" + k].copy()) for k in keys}
    O.ema_update(src2, tgt, decay=0.9, step=5, start_iter=2)
    for k in keys:
        np.testing.assert_allclose(tgt[k].numpy(), g["ema_after/" + k], rtol=1e-6, atol=1e-7, err_msg=k)


def test_dcgan_config1_oracle_matches_reference(golden_dir):
    """BASELINE config 1 (DCGAN, CIFAR10-shaped, unconditional, vanilla loss; the reference's CPU-only case): the oracle's
    restatement of src/models/deep_conv.py reproduces the reference's D phase and G phase.  Weights are regenerated from
    the seed on both sides (O.seeded_state); compared: images, logits, losses, running statistics, every parameter's
    gradient norm and the full gradients of the 1-D parameters."""
    import json
    g = np.load(os.path.join(golden_dir, "dcgan32.npz"))
    ks_g, ks_d = json.loads(str(g["keys_g"])), json.loads(str(g["keys_d"]))
    sdG, sdD = O.seeded_state(ks_g, 101), O.seeded_state(ks_d, 202)
    for sd in (sdD,):
        for k, v in sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    z, real = torch.from_numpy(g["z"]), torch.from_numpy(g["real"])
    with torch.no_grad():
        fake = O.dcgan_generator(sdG, z)
    np.testing.assert_allclose(fake.numpy(), g["fake"], rtol=1e-4, atol=2e-5)
    a, _ = O.dcgan_discriminator(sdD, real)
    b, _ = O.dcgan_discriminator(sdD, fake)
    np.testing.assert_allclose(a.detach().numpy(), g["adv_real"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(b.detach().numpy(), g["adv_fake"], rtol=2e-4, atol=2e-4)
    loss = O.d_vanilla(a, b)
    np.testing.assert_allclose(loss.item(), g["d_loss"], rtol=1e-4)
    loss.backward()
    # floor: conv biases feeding a BatchNorm have analytically zero gradients (pure round-off, ~1e-6 of the largest norm)
    floor_d = 1e-4 * max(float(g[k]) for k in g.files if k.startswith("Dgnorm/"))
    floor_g = 1e-4 * max(float(g[k]) for k in g.files if k.startswith("Ggnorm/"))
    for k in g.files:
        if k.startswith("Dgnorm/"):
            np.testing.assert_allclose(float(sdD[k[7:]].grad.norm()), float(g[k]), rtol=2e-3, atol=floor_d, err_msg=k)
        if k.startswith("Dgrad/"):
            ref = g[k]
            assert np.abs(sdD[k[6:]].grad.numpy() - ref).max() <= 2e-3 * (1e-4 + np.abs(ref).max()) + floor_d, k
        if k.startswith("D1/"):
            np.testing.assert_allclose(sdD[k[3:]].detach().numpy(), g[k], rtol=2e-4, atol=1e-5, err_msg=k)
        if k.startswith("G1/"):
            np.testing.assert_allclose(sdG[k[3:]].detach().numpy(), g[k], rtol=2e-4, atol=1e-5, err_msg=k)
    # generator phase (continues from the statistics left by the discriminator phase).  At B = 8 the discriminator's input
    # gradient is sensitive to 1e-5 perturbations of its input (ReLU masks under small-batch BatchNorm): the oracle's fake2
    # differs from the reference's by ~2e-5 and dL/dfake2 then moves by ~5e-3, although both agree to 2e-6 on identical
    # inputs.  The chain is therefore pinned link by link: (1) D's input gradient at the reference's own fake2,
    # (2) G's backward fed with the reference's dL/dfake2.
    sdD2 = {k: v.detach() for k, v in sdD.items()}
    x_ref = torch.from_numpy(g["fake2"]).clone().requires_grad_(True)
    gl = O.g_vanilla(O.dcgan_discriminator(dict(sdD2), x_ref)[0])
    np.testing.assert_allclose(gl.item(), g["g_loss"], rtol=1e-5)
    (dx,) = torch.autograd.grad(gl, x_ref)
    ref_dx = torch.from_numpy(g["dfake2"])
    assert float((dx - ref_dx).norm() / ref_dx.norm()) < 1e-4
    for k, v in sdG.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    fake2 = O.dcgan_generator(sdG, z)
    np.testing.assert_allclose(fake2.detach().numpy(), g["fake2"], rtol=1e-4, atol=2e-5)
    fake2.backward(ref_dx)
    for k in g.files:
        if k.startswith("Ggnorm/"):
            np.testing.assert_allclose(float(sdG[k[7:]].grad.norm()), float(g[k]), rtol=2e-3, atol=floor_g, err_msg=k)
        if k.startswith("Ggrad/"):
            ref = g[k]
            assert np.abs(sdG[k[6:]].grad.numpy() - ref).max() <= 2e-3 * (1e-4 + np.abs(ref).max()) + floor_g, k


def test_deep_256_bench_topology(golden_dir):
    """The oracle at BASELINE config 4's topology (256x256, depth 2, attention at 64x64: N = 4096, M = 1024) against the
    reference golden deep256_c16_attn_d2 (weights regenerated from the generator script's seed through the product
    modules, whose seeded initialisation equals the reference's bit for bit)."""
    import importlib
    from sgb200 import config as C
    from sgb200.utils import ops
    g = np.load(os.path.join(golden_dir, "deep256_c16_attn_d2.npz"))
    deep = importlib.import_module("sgb200.models.big_resnet_deep_legacy")
    torch.manual_seed(1234)
    M = C.make_modules(True, True, "cBN", "big_resnet_deep_legacy")
    MODEL = C._Section(info_type="N/A", g_info_injection="N/A")
    G = deep.Generator(z_dim=16, g_shared_dim=16, img_size=256, g_conv_dim=16, apply_attn=True, attn_g_loc=[4], g_cond_mtd="cBN",
                       num_classes=5, g_init="ortho", g_depth=2, mixed_precision=False, MODULES=M, MODEL=MODEL)
    D = deep.Discriminator(img_size=256, d_conv_dim=16, apply_d_sn=True, apply_attn=True, attn_d_loc=[2], d_cond_mtd="PD",
                           aux_cls_type="W/O", d_embed_dim="N/A", normalize_d_embed=False, num_classes=5, d_init="ortho",
                           d_depth=2, mixed_precision=False, MODULES=M, MODEL=MODEL)
    with torch.no_grad():
        for mod in list(G.modules()) + list(D.modules()):
            if isinstance(mod, ops.SelfAttention):
                mod.sigma.fill_(0.37)
    sdG = {k: v.detach().clone() for k, v in G.state_dict().items()}
    sdD = {k: v.detach().clone() for k, v in D.state_dict().items()}
    l1 = sum(float(p.detach().double().abs().sum()) for p in G.parameters())
    assert abs(l1 - float(g["param_l1_G"])) < 1e-6 * l1
    pD = [k for k, _ in D.named_parameters()]
    for k in pD:
        sdD[k].requires_grad_(True)
    gi = torch.Generator().manual_seed(77)
    z = torch.randn(16, 16, generator=gi)
    yf = torch.randint(0, 5, (16,), generator=gi)
    real = torch.rand(16, 3, 256, 256, generator=gi) * 2 - 1
    yr = torch.randint(0, 5, (16,), generator=gi)
    kw_g = dict(img_size=256, g_conv_dim=16, g_depth=2, attn_g_loc=(4,), apply_attn=True)
    kw_d = dict(img_size=256, d_conv_dim=16, d_depth=2, attn_d_loc=(2,), apply_attn=True)
    with torch.no_grad():
        fake = O.deep_generator(sdG, z, yf, **kw_g)
    np.testing.assert_allclose(fake[:, :, ::8, ::8].numpy(), g["fake_sub8"], rtol=1e-3, atol=2e-4)
    adv_r, h_r = O.deep_discriminator(sdD, real, yr, **kw_d)
    adv_f, _ = O.deep_discriminator(sdD, fake, yf, **kw_d)
    np.testing.assert_allclose(adv_r.detach().numpy(), g["adv_real"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(adv_f.detach().numpy(), g["adv_fake"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(h_r.detach().numpy(), g["h_real"], rtol=1e-3, atol=1e-3)
    (-adv_r.mean() + adv_f.mean()).backward()                 # losses.d_wasserstein (src/utils/losses.py:214-215)
    gmax = max(float(g["Dgrad/norm/" + k]) for k in pD)
    for k in pD:
        ref_n = float(g["Dgrad/norm/" + k])
        assert abs(float(sdD[k].grad.norm()) - ref_n) <= 5e-3 * (ref_n + 1e-3 * gmax), k
        if "Dgrad/full/" + k in g.files:
            ref = g["Dgrad/full/" + k]
            assert np.abs(sdD[k].grad.numpy() - ref).max() <= 5e-3 * (1e-3 * gmax + np.abs(ref).max()), k


def test_augmentations_closed_form_and_rng_order_match_reference(golden_dir):
    """DiffAugment / CR augmentation: the product's parameter draw (sgb200.utils.diffaug.draw_params / cr.draw_params)
    consumes the generator exactly like the reference, and the closed form the device kernels implement equals the
    reference's pass-by-pass result (values 1e-5: the contrast mean is taken before instead of after the saturation step;
    input gradients 1e-5)."""
    from sgb200.utils import cr, diffaug
    g = np.load(os.path.join(golden_dir, "augment.npz"))
    for tag in ("a", "b"):
        x = torch.from_numpy(g["x_" + tag])
        ct = torch.from_numpy(g["ct_" + tag])
        B, _, H, W = x.shape
        for pname, policy in (("full", "color,translation,cutout"), ("color", "color"), ("geo", "translation,cutout")):
            torch.manual_seed(4242)
            params = diffaug.draw_params(B, H, W, policy, "cpu")
            xr = x.clone().requires_grad_(True)
            names = policy.split(",")
            y = O.diffaug_closed_form(xr, params, "color" in names, "translation" in names, "cutout" in names)
            np.testing.assert_allclose(y.detach().numpy(), g["diffaug_%s_%s" % (pname, tag)], rtol=1e-5, atol=1e-5)
            y.backward(ct)
            np.testing.assert_allclose(xr.grad.numpy(), g["diffaug_%s_%s_dx" % (pname, tag)], rtol=1e-5, atol=1e-5)
        torch.manual_seed(777)
        f, tx, ty = cr.draw_params(B, H, W, "cpu")
        np.testing.assert_array_equal(O.cr_aug_closed_form(x, f, tx, ty).numpy(), g["cr_" + tag])


def test_epilogue_algorithms_equal_the_reference_operators():
    """The restatements of what the GEMM epilogues compute (two-pass softmax from per-part (max, sum exp) partials, softmax
    backward through delta = rowsum(dO * O), ReLU bit planes) against torch's own softmax / autograd and numpy (fp64)."""
    g = torch.Generator().manual_seed(7)
    S = torch.randn(3, 40, 256, generator=g, dtype=torch.float64) * 4
    for width in (64, 128, 256):
        P = O.softmax_from_partials(S, width)
        assert torch.allclose(P, torch.softmax(S, -1), rtol=1e-12, atol=1e-15)
    # backward: d/dS of <dO, softmax(S) @ V>
    V = torch.randn(3, 256, 32, generator=g, dtype=torch.float64)
    dO = torch.randn(3, 40, 32, generator=g, dtype=torch.float64)
    Sg = S.clone().requires_grad_(True)
    (torch.softmax(Sg, -1) @ V * dO).sum().backward()
    dS = O.softmax_backward_with_delta(torch.softmax(S, -1), dO, V)
    assert torch.allclose(dS, Sg.grad, rtol=1e-10, atol=1e-13)
    # bit planes: word c of a pixel = channels 64c .. 64c+63, little-endian
    y = torch.relu(torch.randn(2, 3, 3, 128, generator=g)).numpy()
    bits = O.relu_bit_planes(y)
    assert bits.shape == (2, 3, 3, 16) and bits.dtype == np.uint8
    words = bits.view(np.uint64).reshape(2, 3, 3, 2)
    for c in (0, 5, 63, 64, 100, 127):
        assert np.array_equal((words[..., c // 64] >> np.uint64(c % 64)) & np.uint64(1), (y[..., c] > 0).astype(np.uint64))

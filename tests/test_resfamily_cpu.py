"""BigGAN (big_resnet) and ResNetGAN (resnet: SNGAN / cBN / WGAN-style BN discriminator): oracle pinned to the reference
goldens, and sgb200 module state_dict keys + seeded initialisation bit-identical to the reference modules (CPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import studiogan_oracle as O
from sgb200 import config as C

CASES = {
    "biggan32_c32_attn": dict(family="big_resnet", conv_dim=32, attn=True, g_sn=True, d_sn=True, g_cond="cBN", d_cond="PD", adv="hinge", z_dim=20),
    "sngan32_c16": dict(family="resnet", attn=False, g_sn=False, d_sn=True, g_cond="W/O", d_cond="W/O", adv="hinge", z_dim=32),
    "resnet32_cbn_c16": dict(family="resnet", attn=False, g_sn=False, d_sn=True, g_cond="cBN", d_cond="PD", adv="hinge", z_dim=32),
    "wgan32_bn_c16": dict(family="resnet", attn=False, g_sn=False, d_sn=False, g_cond="W/O", d_cond="W/O", adv="wasserstein", z_dim=32),
}


def load_sd(npz, prefix, grad=False):
    """State dict ``prefix``; the G1/ and D1/ snapshots hold buffers only and take their parameters from G0/ and D0/."""
    sd = {}
    if prefix in ("G1/", "D1/"):
        sd = load_sd(npz, prefix[0] + "0/", grad)
    for k in npz.files:
        if k.startswith(prefix):
            t = torch.from_numpy(npz[k].copy())
            if grad and t.is_floating_point() and "running" not in k and not k.endswith(("weight_u", "weight_v")):
                t.requires_grad_(True)
            sd[k[len(prefix):]] = t
    return sd


def oracle_G(c, sd, z, y):
    if c["family"] == "big_resnet":
        return O.biggan_generator(sd, z, y, 32, c.get("conv_dim", 16), attn_g_loc=(2,), apply_attn=c["attn"])
    return O.resnet_generator(sd, z, y, 32, 16, num_classes=5, conditional=c["g_cond"] == "cBN")


def oracle_D(c, sd, x, y):
    return O.res_discriminator(sd, x, y, 32, c.get("conv_dim", 16), attn_d_loc=(1,), apply_attn=c["attn"], cond=c["d_cond"])


@pytest.mark.parametrize("tag", list(CASES))
def test_oracle_matches_reference(golden_dir, tag):
    c = CASES[tag]
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    z, yf = torch.from_numpy(g["z"]), torch.from_numpy(g["y_fake"])
    real, yr = torch.from_numpy(g["real"]), torch.from_numpy(g["y_real"])
    dl = {"hinge": O.d_hinge, "wasserstein": O.d_wasserstein}[c["adv"]]
    gl = {"hinge": O.g_hinge, "wasserstein": O.g_wasserstein}[c["adv"]]
    sdG, sdD = load_sd(g, "G0/"), load_sd(g, "D0/", grad=True)
    with torch.no_grad():
        fake = oracle_G(c, sdG, z, yf)
    np.testing.assert_allclose(fake.numpy(), g["fake"], rtol=1e-4, atol=3e-5)
    a, h = oracle_D(c, sdD, real, yr)
    b, _ = oracle_D(c, sdD, fake, yf)
    np.testing.assert_allclose(a.detach().numpy(), g["adv_real"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(b.detach().numpy(), g["adv_fake"], rtol=2e-4, atol=2e-4)
    loss = dl(a, b)
    np.testing.assert_allclose(loss.item(), g["d_loss"], rtol=1e-4)
    loss.backward()
    gmax = max(np.abs(g[k]).max() for k in g.files if k.startswith("Dgrad/"))
    for k in g.files:
        if k.startswith("Dgrad/"):      # floor: biases feeding a BatchNorm have analytically zero gradients (1e-6 round-off)
            ref, got = g[k], sdD[k[6:]].grad.numpy()
            assert np.abs(got - ref).max() <= 1e-2 * (1e-3 * gmax + np.abs(ref).max()), k   # single ReLU-boundary flips reach ~5e-3
    for k in g.files:
        if k.startswith("D1/") and ("weight_u" in k or "running_" in k):
            np.testing.assert_allclose(sdD[k[3:]].detach().numpy(), g[k], rtol=2e-4, atol=1e-5, err_msg=k)
    sdG2, sdD2 = load_sd(g, "G1/", grad=True), load_sd(g, "D1/")
    fake2 = oracle_G(c, sdG2, z, yf)
    np.testing.assert_allclose(fake2.detach().numpy(), g["fake2"], rtol=1e-4, atol=3e-5)
    a2, _ = oracle_D(c, sdD2, fake2, yf)
    l2 = gl(a2)
    np.testing.assert_allclose(l2.item(), g["g_loss"], rtol=1e-4)
    l2.backward()
    gmax = max(np.abs(g[k]).max() for k in g.files if k.startswith("Ggrad/"))
    for k in g.files:
        if k.startswith("Ggrad/"):
            ref, got = g[k], sdG2[k[6:]].grad.numpy()
            assert np.abs(got - ref).max() <= 1e-2 * (1e-3 * gmax + np.abs(ref).max()), k   # single ReLU-boundary flips reach ~5e-3


def build(c):
    import importlib
    mod = importlib.import_module("sgb200.models." + c["family"])
    M = C.make_modules(c["g_sn"], c["d_sn"], c["g_cond"], c["family"])
    MODEL = C._Section(info_type="N/A", g_info_injection="N/A")
    torch.manual_seed(4321)
    G = mod.Generator(z_dim=c["z_dim"], g_shared_dim=16, img_size=32, g_conv_dim=c.get("conv_dim", 16), apply_attn=c["attn"], attn_g_loc=[2],
                      g_cond_mtd=c["g_cond"], num_classes=5, g_init="ortho", g_depth="N/A", mixed_precision=False, MODULES=M, MODEL=MODEL)
    D = mod.Discriminator(img_size=32, d_conv_dim=c.get("conv_dim", 16), apply_d_sn=c["d_sn"], apply_attn=c["attn"], attn_d_loc=[1], d_cond_mtd=c["d_cond"],
                          aux_cls_type="W/O", d_embed_dim="N/A", normalize_d_embed=False, num_classes=5, d_init="ortho", d_depth="N/A",
                          mixed_precision=False, MODULES=M, MODEL=MODEL)
    return G, D


@pytest.mark.parametrize("tag", list(CASES))
def test_module_keys_and_seeded_init_match_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    G, D = build(CASES[tag])
    for net, prefix in ((G, "G0/"), (D, "D0/")):
        sd = net.state_dict()
        ref_keys = [k[len(prefix):] for k in g.files if k.startswith(prefix)]
        assert list(sd.keys()) == ref_keys
        for k in ref_keys:
            if k.endswith("sigma"):
                continue
            np.testing.assert_array_equal(sd[k].numpy(), g[prefix + k], err_msg=k)


GP_CASES = {
    "gp_resnet32_bn_c16": dict(family="resnet", conv_dim=16, d_sn=False, d_cond="W/O"),
    "gp_resnet32_sn_c16_pd": dict(family="resnet", conv_dim=16, d_sn=True, d_cond="PD"),
    "gp_deep32_sn_c8_pd": dict(family="deep", conv_dim=8, d_sn=True, d_cond="PD"),
}


@pytest.mark.parametrize("tag", list(GP_CASES))
def test_oracle_grad_penalty_matches_reference(golden_dir, tag):
    """O.grad_penalty over the oracle discriminators == the reference's cal_grad_penalty (src/utils/losses.py:301-316):
    interpolates (bit exact), penalty value, and dP/dtheta from the double backward."""
    c = GP_CASES[tag]
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    sdD = load_sd(g, "D0/", grad=True)
    real, fake, yr = torch.from_numpy(g["real"]), torch.from_numpy(g["fake"]), torch.from_numpy(g["y_real"])
    alpha = torch.from_numpy(g["alpha"])
    a4 = alpha.view(-1, 1, 1, 1)
    assert np.array_equal((a4 * real + (1 - a4) * fake).numpy(), g["x_hat"])

    def disc(x):
        if c["family"] == "deep":
            return O.deep_discriminator(sdD, x, yr, img_size=32, d_conv_dim=c["conv_dim"], d_depth=1)[0]
        return O.res_discriminator(sdD, x, yr, 32, c["conv_dim"], cond=c["d_cond"])[0]
    gp = O.grad_penalty(disc, real, fake, alpha)
    np.testing.assert_allclose(gp.item(), g["gp"], rtol=2e-4)
    gp.backward()
    gmax = max(np.abs(g[k]).max() for k in g.files if k.startswith("Dgrad/"))
    for k in g.files:
        if k.startswith("Dgrad/"):
            got = sdD[k[6:]].grad
            got = got.numpy() if got is not None else np.zeros_like(g[k])
            assert np.abs(got - g[k]).max() <= 1e-2 * (1e-3 * gmax + np.abs(g[k]).max()), k
    for k in g.files:
        if k.startswith("D1/") and ("weight_u" in k or "running_" in k):
            np.testing.assert_allclose(sdD[k[3:]].detach().numpy(), g[k], rtol=2e-4, atol=1e-5, err_msg=k)

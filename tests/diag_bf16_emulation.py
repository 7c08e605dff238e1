"""Diagnostic (not a test, CPU only): how much error does bf16 activation / gradient STORAGE alone introduce?

The fp32 oracle is run with every convolution input, weight and output (and their gradients) rounded to bf16, which is
what the CUDA path stores between kernels; the result is compared with the reference golden vectors.  Quoted in
DESIGN.md section 4 ("numerics notes") and in the tolerances of tests/test_gpu_parity.py:

    python tests/diag_bf16_emulation.py
    gp_resnet32_bn_c16     adv err 0.0038  g err 0.184
    gp_resnet32_sn_c16_pd  adv err 0.0016  g err 0.077
    gp_deep32_sn_c8_pd     adv err 0.0011  g err 0.050
(g = gradient of the summed discriminator output w.r.t. the input pixels, the first pass of the gradient penalty)."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "pytorch-studiogan_b200")):
    sys.path.insert(0, p)
from oracle import studiogan_oracle as O  # noqa: E402
from test_resfamily_cpu import load_sd  # noqa: E402


class RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


rb = RoundBF16.apply
_bn = O.batch_norm


def conv(sd, prefix, x, padding, training=True):
    w = O.weight(sd, prefix, training)
    return rb(F.conv2d(rb(x), rb(w), sd.get(prefix + "bias"), stride=1, padding=padding))


def bn(sd, prefix, x, training=True, track=True, affine=False):
    return rb(_bn(sd, prefix, x, training, track, affine))


if __name__ == "__main__":
    O.conv, O.batch_norm = conv, bn
    for tag, fam, cd, cond in [("gp_resnet32_bn_c16", "resnet", 16, "W/O"), ("gp_resnet32_sn_c16_pd", "resnet", 16, "PD"),
                               ("gp_deep32_sn_c8_pd", "deep", 8, "PD")]:
        g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
        sd = load_sd(g, "D0/", grad=True)
        yr = torch.from_numpy(g["y_real"])
        x = torch.from_numpy(g["x_hat"]).requires_grad_(True)
        if fam == "deep":
            adv = O.deep_discriminator(sd, x, yr, img_size=32, d_conv_dim=cd, d_depth=1)[0]
        else:
            adv = O.res_discriminator(sd, x, yr, 32, cd, cond=cond)[0]
        (gr,) = torch.autograd.grad(adv.sum(), x)
        ref_adv, ref_g = torch.from_numpy(g["adv_hat"]), torch.from_numpy(g["g"])
        print("%-22s adv err %.4f  g err %.3f" % (tag, float((adv.detach() - ref_adv).norm() / ref_adv.norm()),
                                                   float((gr - ref_g).norm() / ref_g.norm())))

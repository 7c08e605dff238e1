"""Launches the three conv-engine kernels at shapes of the B = 256 BigGAN-Deep-256 step, twice each (the first launch warms
the tensor-map / attribute caches), for `ncu --set full -k regex:<kernel> -s 1 -c 1` captures:
  rows   conv3x3_rows_kernel   3x3 64->64 @256x256, bias            (generator block 5 / discriminator block 0)
  one    conv_fprop_kernel     1x1 64->128 @256x256, mask bits + half-resolution residual (fused discriminator block entry dgrad)
  wgrad  conv_wgrad_kernel     3x3 128->128 @128x128
usage: python profiles/ncu_target.py rows|one|wgrad [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-studiogan_b200"))
from sgb200 import kernels as K  # noqa: E402

which = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)


def act(c, h, w):
    return torch.randn(B, h, w, c, device=dev, dtype=torch.bfloat16).permute(0, 3, 1, 2)


if which == "rows":
    x = act(64, 256, 256)
    wf, _ = K.weight_pack(torch.randn(64, 64, 3, 3, device=dev) / 24, None, 64, 64, 9, True, False)
    b = torch.randn(64, device=dev)
    for _ in range(2):
        y = K.conv_fprop(x, wf, 64, 3, 3, 1, 1, bias=b)
    print("algorithmic bytes", 2 * x.numel() + 2 * y.numel() + 2 * wf.numel())
elif which == "one":
    dz = act(64, 256, 256)
    a0 = act(128, 256, 256).relu()
    bits = torch.from_numpy(__import__("numpy").packbits((a0.permute(0, 2, 3, 1) > 0).cpu().numpy(), axis=-1, bitorder="little")).to(dev)
    r = act(128, 128, 128)
    _, wd = K.weight_pack(torch.randn(64, 128, 1, 1, device=dev) / 11, None, 64, 128, 1, True, True)
    for _ in range(2):
        y = K.conv_fprop(dz, wd, 128, 1, 1, 0, 0, mask_bits=bits, residual=r, res_up2=True, res_scale=0.25)
    print("algorithmic bytes", 2 * dz.numel() + 2 * y.numel() + bits.numel() + 2 * r.numel())
elif which == "wgrad":
    x, dy = act(128, 128, 128), act(128, 128, 128)
    for _ in range(2):
        K.conv_wgrad(x, dy, 3, 3, 1, 1)
    print("algorithmic bytes", 2 * x.numel() + 2 * dy.numel() + 4 * 128 * 128 * 9)
torch.cuda.synchronize()

"""Residual blocks shared by BigGAN (``big_resnet``) and ResNetGAN / SNGAN / SAGAN (``resnet``): the reference defines
them twice with identical arithmetic (src/models/big_resnet.py:15-42,161-242 and src/models/resnet.py:15-59,172-254).

Fusions used (all exact in real arithmetic; each intermediate is rounded to bf16 once):
* generator: cBN/BN + ReLU + nearest x2 in one pass; the 1x1 skip conv runs at LOW resolution and is added, up-sampled,
  in the epilogue of conv2d2 (a 1x1 conv commutes with nearest up-sampling);
* discriminator: ReLU of conv2d1's output in its epilogue; conv2d0's skip is added in conv2d2's epilogue BEFORE the
  2x2 average pooling (pooling is linear), so each block pools once.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import autograd_ops as A
from ..utils import ops


class GenBlock(nn.Module):
    """bn1 -> ReLU -> up x2 -> conv3x3 -> bn2 -> ReLU -> conv3x3, skip = conv1x1(up(x))."""

    def __init__(self, in_channels, out_channels, g_cond_mtd, affine_input_dim, MODULES, g_info_injection="N/A"):
        super().__init__()
        self.g_cond_mtd = g_cond_mtd
        self.g_info_injection = g_info_injection
        self.conditional = (g_cond_mtd == "cBN") or (g_info_injection == "cBN") or (MODULES.g_bn is ops.ConditionalBatchNorm2d)
        if self.conditional:
            self.bn1 = MODULES.g_bn(affine_input_dim, in_channels, MODULES)
            self.bn2 = MODULES.g_bn(affine_input_dim, out_channels, MODULES)
        else:
            self.bn1 = MODULES.g_bn(in_features=in_channels)
            self.bn2 = MODULES.g_bn(in_features=out_channels)
        self.activation = MODULES.g_act_fn
        self.conv2d0 = MODULES.g_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, stride=1, padding=0)
        self.conv2d1 = MODULES.g_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)
        self.conv2d2 = MODULES.g_conv2d(in_channels=out_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x, affine):
        main, side = A.ForkFn.call(x)
        if self.conditional:
            h = self.bn1(main, affine, relu=True, up2=True)
        else:
            h = self.bn1(main, relu=True, up2=True)
        h = self.conv2d1(h)
        h = self.bn2(h, affine, relu=True) if self.conditional else self.bn2(h, relu=True)
        skip = self.conv2d0(side)                               # low resolution; up-sampled inside conv2d2's epilogue
        return self.conv2d2(h, residual=skip, res_up2=True)


class _ImageSkipTangent:
    """Tangent of x0 = [bn0](avg_pool2d(img)) on the 3-channel image, written with tensor ops (a [B,3,H/2,W/2] array):
    pooling is linear; training-mode BN maps a tangent t to w * r * (t - mean(t) - xhat * mean(xhat * t))."""

    @staticmethod
    def tangent(args, out, tan):
        img, bn = args
        t = tan(img)
        if t is None:
            return None
        t0 = F.avg_pool2d(t, 2)
        if bn is None:
            return t0
        w = bn.weight.view(1, -1, 1, 1) if bn.weight is not None else 1.0
        if not bn.training:
            return w * torch.rsqrt(bn.running_var.view(1, -1, 1, 1) + bn.eps) * t0
        x0 = F.avg_pool2d(img, 2)
        mu = x0.mean((0, 2, 3), keepdim=True)
        r = torch.rsqrt(x0.var((0, 2, 3), unbiased=False, keepdim=True) + bn.eps)
        xh = (x0 - mu) * r
        return w * r * (t0 - t0.mean((0, 2, 3), keepdim=True) - xh * (xh * t0).mean((0, 2, 3), keepdim=True))


class DiscOptBlock(nn.Module):
    """First discriminator block, fed by the image: conv3x3 -> ReLU -> conv3x3 -> pool, skip = conv1x1(pool(image))."""

    def __init__(self, in_channels, out_channels, apply_d_sn, MODULES):
        super().__init__()
        self.apply_d_sn = apply_d_sn
        self.conv2d0 = MODULES.d_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, stride=1, padding=0)
        self.conv2d1 = MODULES.d_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)
        self.conv2d2 = MODULES.d_conv2d(in_channels=out_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)
        if not apply_d_sn:
            self.bn0 = MODULES.d_bn(in_features=in_channels)
            self.bn1 = MODULES.d_bn(in_features=out_channels)
        self.activation = MODULES.d_act_fn
        self.average_pooling = nn.AvgPool2d(2)

    def forward(self, img):
        """``img``: NCHW fp32 image."""
        h = self.conv2d1(A.ImageColFn.call(img), relu=self.apply_d_sn, premasked=self.apply_d_sn)
        if not self.apply_d_sn:
            h = self.bn1(h, relu=True)
            h = self.conv2d2(h)
        else:
            h = self.conv2d2(h, mask_input=True)
        h = A.PoolFn.call(h)
        x0 = F.avg_pool2d(img, 2)                               # 3-channel image: a [B,3,H/2,W/2] tensor op
        if not self.apply_d_sn:
            x0 = self.bn0(x0) if not isinstance(self.bn0, ops.BatchNorm2d) else F.batch_norm(
                x0, self.bn0.running_mean, self.bn0.running_var, self.bn0.weight, self.bn0.bias,
                self.bn0.training, self.bn0.momentum, self.bn0.eps)
        A.tape_record(_ImageSkipTangent, (img, None if self.apply_d_sn else self.bn0), x0)
        s = self.conv2d0(A.ImageInFn.call(x0))
        return A.AddFn.call(h, s)


class DiscBlock(nn.Module):
    """[bn1] -> ReLU -> conv3x3 -> [bn2] -> ReLU -> conv3x3 -> [pool]; skip = [pool](conv1x1([bn0](x))) or identity.
    With spectral norm (no BN) the reference's in-place ReLU also rectifies the aliased skip tensor, i.e. the skip path
    sees relu(x) (src/config.py:486 + src/models/big_resnet.py:225-228); with BN in between it sees x."""

    def __init__(self, in_channels, out_channels, apply_d_sn, MODULES, downsample=True):
        super().__init__()
        self.apply_d_sn = apply_d_sn
        self.downsample = downsample
        self.activation = MODULES.d_act_fn
        self.ch_mismatch = in_channels != out_channels
        if self.ch_mismatch or downsample:
            self.conv2d0 = MODULES.d_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, stride=1, padding=0)
            if not apply_d_sn:
                self.bn0 = MODULES.d_bn(in_features=in_channels)
        self.conv2d1 = MODULES.d_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)
        self.conv2d2 = MODULES.d_conv2d(in_channels=out_channels, out_channels=out_channels, kernel_size=3, stride=1, padding=1)
        if not apply_d_sn:
            self.bn1 = MODULES.d_bn(in_features=in_channels)
            self.bn2 = MODULES.d_bn(in_features=out_channels)
        self.average_pooling = nn.AvgPool2d(2)

    def forward(self, x):
        has_sc = self.downsample or self.ch_mismatch
        if self.apply_d_sn:
            a = A.ReluFn.call(x)                               # in-place ReLU of the reference: both branches see relu(x)
            main, side = A.ForkFn.call(a)
            h = self.conv2d1(main, relu=True, premasked=True)
            skip = self.conv2d0(side) if has_sc else side
            h = self.conv2d2(h, residual=skip, mask_input=True)
        else:
            main, side = A.ForkFn.call(x)
            h = self.conv2d1(self.bn1(main, relu=True))
            h = self.bn2(h, relu=True)
            skip = self.conv2d0(self.bn0(side)) if has_sc else side
            h = self.conv2d2(h, residual=skip)
        return A.PoolFn.call(h) if self.downsample else h

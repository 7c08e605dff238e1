"""ResNetGAN generator / discriminator — SNGAN, SAGAN, WGAN-GP backbones (reference ``src/models/resnet.py``):
whole-z linear0, plain BN or cBN conditioned on ONE-HOT labels (:111-112,140), discriminator with BN when spectral norm
is off.  Same constructor / forward signatures, sub-module names and registration order as the reference."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import autograd_ops as A
from ..snbatch import SNBatch
from ..utils import ops
from ._resblocks import DiscBlock, DiscOptBlock, GenBlock
from .big_resnet import D_DOWN, D_IN, D_OUT, G_IN, G_OUT


class Generator(nn.Module):
    def __init__(self, z_dim, g_shared_dim, img_size, g_conv_dim, apply_attn, attn_g_loc, g_cond_mtd, num_classes, g_init,
                 g_depth, mixed_precision, MODULES, MODEL):
        super().__init__()
        key = str(img_size)
        self.z_dim = z_dim
        self.num_classes = num_classes
        self.g_cond_mtd = g_cond_mtd
        self.mixed_precision = mixed_precision
        self.MODEL = MODEL
        self.in_dims = [g_conv_dim * m for m in G_IN[key]]
        self.out_dims = [g_conv_dim * m for m in G_OUT[key]]
        self.bottom = 4
        self.num_blocks = len(self.in_dims)
        self.affine_input_dim = 0
        if getattr(MODEL, "info_type", "N/A") != "N/A":
            raise NotImplementedError("InfoGAN heads are outside the sgb200 hot-path scope (SURVEY.md section 8)")
        self.g_info_injection = getattr(MODEL, "g_info_injection", "N/A")

        self.linear0 = MODULES.g_linear(in_features=self.z_dim, out_features=self.in_dims[0] * self.bottom * self.bottom, bias=True)
        if self.g_cond_mtd != "W/O" and self.g_cond_mtd == "cBN":
            self.affine_input_dim += self.num_classes

        blocks = []
        for index in range(self.num_blocks):
            blocks.append([GenBlock(in_channels=self.in_dims[index], out_channels=self.out_dims[index], g_cond_mtd=self.g_cond_mtd,
                                    affine_input_dim=self.affine_input_dim, MODULES=MODULES, g_info_injection=self.g_info_injection)])
            if index + 1 in attn_g_loc and apply_attn:
                blocks.append([ops.SelfAttention(self.out_dims[index], is_generator=True, MODULES=MODULES)])
        self.blocks = nn.ModuleList([nn.ModuleList(b) for b in blocks])

        self.bn4 = ops.batchnorm_2d(in_features=self.out_dims[-1])
        self.activation = MODULES.g_act_fn
        self.conv2d5 = MODULES.g_conv2d(in_channels=self.out_dims[-1], out_channels=3, kernel_size=3, stride=1, padding=1)
        self.tanh = nn.Tanh()
        ops.init_weights(self.modules, g_init)
        self.linear0._perm_S = self.bottom * self.bottom
        self._snb = SNBatch(self)

    def forward(self, z, label, shared_label=None, eval=False):
        self._snb.run()
        affines = None
        if self.g_cond_mtd != "W/O":
            onehot = F.one_hot(label, num_classes=self.num_classes).to(torch.float32)
            affines = A.ToBF16Fn.call(onehot)
        S = self.bottom * self.bottom
        act = self.linear0(z, perm_S=S)
        B = act.shape[0]
        act = act.reshape(B, self.bottom, self.bottom, self.in_dims[0]).permute(0, 3, 1, 2)
        for blocklist in self.blocks:
            for block in blocklist:
                act = block(act) if isinstance(block, ops.SelfAttention) else block(act, affines)
        act = self.bn4(act, relu=True)
        act = self.conv2d5(act)
        self._snb.clear()
        return A.ImageOutFn.call(act, 3)


class Discriminator(nn.Module):
    def __init__(self, img_size, d_conv_dim, apply_d_sn, apply_attn, attn_d_loc, d_cond_mtd, aux_cls_type, d_embed_dim,
                 normalize_d_embed, num_classes, d_init, d_depth, mixed_precision, MODULES, MODEL):
        super().__init__()
        key = str(img_size)
        self.d_cond_mtd = d_cond_mtd
        self.aux_cls_type = aux_cls_type
        self.normalize_d_embed = normalize_d_embed
        self.num_classes = num_classes
        self.mixed_precision = mixed_precision
        self.in_dims = [3] + [d_conv_dim * m for m in D_IN[key]]
        self.out_dims = [d_conv_dim * m for m in D_OUT[key]]
        self.MODEL = MODEL
        down = D_DOWN[key]
        if getattr(MODEL, "info_type", "N/A") != "N/A":
            raise NotImplementedError("InfoGAN heads are outside the sgb200 hot-path scope (SURVEY.md section 8)")

        blocks = []
        for index in range(len(self.in_dims)):
            if index == 0:
                blocks.append([DiscOptBlock(in_channels=self.in_dims[index], out_channels=self.out_dims[index],
                                            apply_d_sn=apply_d_sn, MODULES=MODULES)])
            else:
                blocks.append([DiscBlock(in_channels=self.in_dims[index], out_channels=self.out_dims[index], apply_d_sn=apply_d_sn,
                                         MODULES=MODULES, downsample=down[index])])
            if index + 1 in attn_d_loc and apply_attn:
                blocks.append([ops.SelfAttention(self.out_dims[index], is_generator=False, MODULES=MODULES)])
        self.blocks = nn.ModuleList([nn.ModuleList(b) for b in blocks])
        self.activation = MODULES.d_act_fn
        ops.build_discriminator_head(self, MODULES, self.out_dims[-1], d_cond_mtd, aux_cls_type, d_embed_dim, num_classes)
        if d_init:
            ops.init_weights(self.modules, d_init)
        self._snb = SNBatch(self)

    def forward(self, x, label, eval=False, adc_fake=False):
        self._snb.run()
        h = x
        for blocklist in self.blocks:
            for block in blocklist:
                h = block(h)
        h = A.SumHWFn.call(h, True)
        self._snb.clear()
        return ops.discriminator_head(self, h, label, adc_fake)

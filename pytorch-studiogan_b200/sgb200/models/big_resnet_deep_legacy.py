"""BigGAN-Deep (CompareGAN flavour) generator / discriminator on the sgb200 kernel set.

Drop-in for the reference ``src/models/big_resnet_deep_legacy.py``: identical constructor signatures, sub-module
names / registration order (so ``state_dict()`` keys, shapes and seeded initialisation agree), ``forward`` signatures
and the 12-key discriminator output dict (reference :400-413).  Internally activations are NHWC bf16 and every block is
a chain of fused kernels:

  GenBlock   (ref :49-73)   cBN+ReLU -> 1x1 -> cBN+ReLU(+nearest x2) -> 3x3 -> cBN+ReLU -> 3x3 -> cBN+ReLU -> 1x1 (+skip in the epilogue)
  DiscBlock  (ref :210-229) ReLU -> 1x1(+ReLU) -> 3x3(+ReLU) -> 3x3(+ReLU) -> [avgpool] -> 1x1 (+concat-skip in the epilogue);
             the skip is taken from relu(x) because the reference activation is in-place (see DBlockEntryFn)
"""
import torch
import torch.nn as nn

from .. import autograd_ops as A
from ..snbatch import SNBatch
from ..utils import ops

G_IN = {"32": [4, 4, 4], "64": [16, 8, 4, 2], "128": [16, 16, 8, 4, 2], "256": [16, 16, 8, 8, 4, 2],
        "512": [16, 16, 8, 8, 4, 2, 1]}
G_OUT = {"32": [4, 4, 4], "64": [8, 4, 2, 1], "128": [16, 8, 4, 2, 1], "256": [16, 8, 8, 4, 2, 1],
         "512": [16, 8, 8, 4, 2, 1, 1]}
D_IN = {"32": [4, 4, 4], "64": [1, 2, 4, 8], "128": [1, 2, 4, 8, 16], "256": [1, 2, 4, 8, 8, 16],
        "512": [1, 1, 2, 4, 8, 8, 16]}
D_OUT = {"32": [4, 4, 4], "64": [2, 4, 8, 16], "128": [2, 4, 8, 16, 16], "256": [2, 4, 8, 8, 16, 16],
         "512": [1, 2, 4, 8, 8, 16, 16]}
D_DOWN = {"32": [True, True, False, False], "64": [True, True, True, True, False],
          "128": [True, True, True, True, True, False], "256": [True, True, True, True, True, True, False],
          "512": [True, True, True, True, True, True, True, False]}
BOTTOM = 4


class GenBlock(nn.Module):
    def __init__(self, in_channels, out_channels, g_cond_mtd, affine_input_dim, upsample, MODULES, channel_ratio=4):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.g_cond_mtd = g_cond_mtd
        self.upsample = upsample
        self.hidden_channels = in_channels // channel_ratio
        hid = self.hidden_channels

        self.bn1 = MODULES.g_bn(affine_input_dim, in_channels, MODULES)
        self.bn2 = MODULES.g_bn(affine_input_dim, hid, MODULES)
        self.bn3 = MODULES.g_bn(affine_input_dim, hid, MODULES)
        self.bn4 = MODULES.g_bn(affine_input_dim, hid, MODULES)

        self.activation = MODULES.g_act_fn

        self.conv2d1 = MODULES.g_conv2d(in_channels=in_channels, out_channels=hid, kernel_size=1, stride=1, padding=0)
        self.conv2d2 = MODULES.g_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d3 = MODULES.g_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d4 = MODULES.g_conv2d(in_channels=hid, out_channels=out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, affine):
        main, skip = A.SplitResidualFn.call(x, self.out_channels)
        h = self.conv2d1(self.bn1(main, affine, relu=True))
        h = self.conv2d2(self.bn2(h, affine, relu=True, up2=self.upsample))
        h = self.conv2d3(self.bn3(h, affine, relu=True))
        return self.conv2d4(self.bn4(h, affine, relu=True), residual=skip, res_up2=self.upsample)


class Generator(nn.Module):
    def __init__(self, z_dim, g_shared_dim, img_size, g_conv_dim, apply_attn, attn_g_loc, g_cond_mtd, num_classes, g_init,
                 g_depth, mixed_precision, MODULES, MODEL):
        super().__init__()
        key = str(img_size)
        self.z_dim = z_dim
        self.g_shared_dim = g_shared_dim
        self.g_cond_mtd = g_cond_mtd
        self.num_classes = num_classes
        self.mixed_precision = mixed_precision
        self.MODEL = MODEL
        self.in_dims = [g_conv_dim * m for m in G_IN[key]]
        self.out_dims = [g_conv_dim * m for m in G_OUT[key]]
        self.bottom = BOTTOM
        self.num_blocks = len(self.in_dims)
        self.affine_input_dim = self.z_dim

        if getattr(MODEL, "info_type", "N/A") != "N/A":
            raise NotImplementedError("InfoGAN heads are outside the sgb200 hot-path scope (SURVEY.md section 8)")

        if self.g_cond_mtd != "W/O":
            self.affine_input_dim += self.g_shared_dim
            self.shared = ops.embedding(num_embeddings=self.num_classes, embedding_dim=self.g_shared_dim)

        self.linear0 = MODULES.g_linear(in_features=self.affine_input_dim,
                                        out_features=self.in_dims[0] * self.bottom * self.bottom, bias=True)

        blocks = []
        for index in range(self.num_blocks):
            for g_index in range(g_depth):
                last = g_index == g_depth - 1
                blocks.append([GenBlock(in_channels=self.in_dims[index],
                                        out_channels=self.out_dims[index] if g_index != 0 else self.in_dims[index],
                                        g_cond_mtd=g_cond_mtd, affine_input_dim=self.affine_input_dim,
                                        upsample=last, MODULES=MODULES)])
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
        self._snb.run()                                               # one power iteration + packs for every layer
        parts = []
        if self.g_cond_mtd != "W/O":
            if shared_label is None:
                shared_label = self.shared(label)
            parts.append(shared_label)
        if parts:
            z = torch.cat(parts + [z], 1)
        affine = A.ToBF16Fn.call(z)                                  # [B, K, 1, 1] bf16, shared by every cBN
        if self.g_cond_mtd == "cBN":
            self._snb.cbn_affine_all(affine)                         # gradient-free passes: all 96 gain / bias maps as one GEMM
        S = self.bottom * self.bottom
        act = self.linear0(affine, perm_S=S)                          # [B, S*C0, 1, 1], features already in (s, c) order
        B = act.shape[0]
        act = act.reshape(B, self.bottom, self.bottom, self.in_dims[0]).permute(0, 3, 1, 2)
        for blocklist in self.blocks:
            for block in blocklist:
                act = block(act) if isinstance(block, ops.SelfAttention) else block(act, affine)
        act = self.bn4(act, relu=True)
        act = self.conv2d5(act)
        self._snb.clear()
        return A.ImageOutFn.call(act, 3)


class DiscBlock(nn.Module):
    def __init__(self, in_channels, out_channels, MODULES, downsample=True, channel_ratio=4):
        super().__init__()
        self.downsample = downsample
        hid = out_channels // channel_ratio

        self.activation = MODULES.d_act_fn
        self.conv2d1 = MODULES.d_conv2d(in_channels=in_channels, out_channels=hid, kernel_size=1, stride=1, padding=0)
        self.conv2d2 = MODULES.d_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d3 = MODULES.d_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d4 = MODULES.d_conv2d(in_channels=hid, out_channels=out_channels, kernel_size=1, stride=1, padding=0)

        self.learnable_sc = in_channels != out_channels
        if self.learnable_sc:
            self.conv2d0 = MODULES.d_conv2d(in_channels=in_channels, out_channels=out_channels - in_channels, kernel_size=1,
                                            stride=1, padding=0)
        if self.downsample:
            self.average_pooling = nn.AvgPool2d(2)

    def forward(self, x, in_relu=False, out_relu=False):
        """in_relu: ``x`` already is relu(block input) -- the producing layer applied the reference's in-place ReLU
        (src/config.py:486) in its epilogue; out_relu: do the same for the next block / the head.  The gradient contract of
        such a tensor is "premasked": whoever consumes it returns a gradient that is already zero where it is zero."""
        if A.TAPE is not None:
            return self._forward_taped(x, in_relu, out_relu)
        a0 = x if in_relu else A.ReluPassFn.call(x)
        c1 = self.conv2d1
        h, px = A.DEntryConvFn.call(a0, ops._w(c1), c1.bias,
                                    {"downsample": self.downsample, "sn": getattr(c1, "_sn", None), "do_power_iteration": c1.training,
                                     "sn_cache": getattr(c1, "_sn_cache", None), "sn_pass": getattr(c1, "_sn_pass", None),
                                     "skip_channels": self.conv2d4.out_channels if self.learnable_sc else 0})
        h = self.conv2d2(h, relu=True, premasked=True, mask_input=True)
        h = self.conv2d3(h, relu=True, premasked=True, mask_input=True)
        if self.downsample:
            h = A.AvgPoolFn.call(h, True)
        skip = px
        if self.learnable_sc:
            c0 = self.conv2d0
            skip = A.ConcatSkipFn.call(px, ops._w(c0), c0.bias,
                                        {"sn": getattr(c0, "_sn", None), "do_power_iteration": c0.training,
                                         "sn_cache": getattr(c0, "_sn_cache", None), "sn_pass": getattr(c0, "_sn_pass", None)})
        return self.conv2d4(h, residual=skip, mask_input=not self.downsample, relu=out_relu, premasked=out_relu)

    def _forward_taped(self, x, in_relu, out_relu):
        """Op-by-op form recorded on the tangent tape (gradient penalty, utils/gp.py)."""
        if in_relu or out_relu:
            raise NotImplementedError("fused ReLU hand-over between blocks under the tangent tape")
        a0, px = A.DBlockEntryFn.call(x, self.downsample)
        h = self.conv2d1(a0, relu=True, premasked=True, mask_input=True)
        h = self.conv2d2(h, relu=True, premasked=True, mask_input=True)
        h = self.conv2d3(h, relu=True, premasked=True, mask_input=True)
        if self.downsample:
            h = A.AvgPoolFn.call(h, True)
        skip = px
        if self.learnable_sc:
            c0 = self.conv2d0
            skip = A.ConcatSkipFn.call(px, ops._w(c0), c0.bias,
                                        {"sn": getattr(c0, "_sn", None), "do_power_iteration": c0.training,
                                         "sn_cache": getattr(c0, "_sn_cache", None), "sn_pass": getattr(c0, "_sn_pass", None)})
        return self.conv2d4(h, residual=skip, mask_input=not self.downsample)


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
        self.in_dims = [d_conv_dim * m for m in D_IN[key]]
        self.out_dims = [d_conv_dim * m for m in D_OUT[key]]
        self.MODEL = MODEL
        down = D_DOWN[key]
        if getattr(MODEL, "info_type", "N/A") != "N/A":
            raise NotImplementedError("InfoGAN heads are outside the sgb200 hot-path scope (SURVEY.md section 8)")

        self.input_conv = MODULES.d_conv2d(in_channels=3, out_channels=self.in_dims[0], kernel_size=3, stride=1, padding=1)

        blocks = []
        for index in range(len(self.in_dims)):
            for d_index in range(d_depth):
                first = d_index == 0
                blocks.append([DiscBlock(in_channels=self.in_dims[index] if first else self.out_dims[index],
                                         out_channels=self.out_dims[index], MODULES=MODULES,
                                         downsample=bool(down[index] and first))])
            if (index + 1) in attn_d_loc and apply_attn:
                blocks.append([ops.SelfAttention(self.out_dims[index], is_generator=False, MODULES=MODULES)])
        self.blocks = nn.ModuleList([nn.ModuleList(b) for b in blocks])

        self.activation = MODULES.d_act_fn

        feat = self.out_dims[-1]
        if self.d_cond_mtd == "MH":
            self.linear1 = MODULES.d_linear(in_features=feat, out_features=1 + num_classes, bias=True)
        elif self.d_cond_mtd == "MD":
            self.linear1 = MODULES.d_linear(in_features=feat, out_features=num_classes, bias=True)
        else:
            self.linear1 = MODULES.d_linear(in_features=feat, out_features=1, bias=True)

        if self.aux_cls_type == "ADC":
            num_classes = num_classes * 2

        if self.d_cond_mtd == "AC":
            self.linear2 = MODULES.d_linear(in_features=feat, out_features=num_classes, bias=False)
        elif self.d_cond_mtd == "PD":
            self.embedding = MODULES.d_embedding(num_classes, feat)
        elif self.d_cond_mtd in ["2C", "D2DCE"]:
            self.linear2 = MODULES.d_linear(in_features=feat, out_features=d_embed_dim, bias=True)
            self.embedding = MODULES.d_embedding(num_classes, d_embed_dim)

        if self.aux_cls_type == "TAC":
            if self.d_cond_mtd == "AC":
                self.linear_mi = MODULES.d_linear(in_features=feat, out_features=num_classes, bias=False)
            elif self.d_cond_mtd in ["2C", "D2DCE"]:
                self.linear_mi = MODULES.d_linear(in_features=feat, out_features=d_embed_dim, bias=True)
                self.embedding_mi = MODULES.d_embedding(num_classes, d_embed_dim)
            else:
                raise NotImplementedError

        if d_init:
            ops.init_weights(self.modules, d_init)
        for name in ("linear1", "linear2", "linear_mi"):
            if hasattr(self, name):
                getattr(self, name)._head_layer = True      # fp32 head path, outside the batched SN / pack pass
        self._snb = SNBatch(self)

    def forward(self, x, label, eval=False, adc_fake=False):
        self._snb.run()
        seq = [b for bl in self.blocks for b in bl]
        fuse = A.TAPE is None              # ReLU hand-over between consecutive blocks through the conv epilogues
        first_relu = fuse and isinstance(seq[0], DiscBlock)
        h = self.input_conv(A.ImageColFn.call(x), relu=first_relu, premasked=first_relu)   # 3x3 patches of the image -> K = 32 GEMM
        relu_in = first_relu
        for i, block in enumerate(seq):
            if isinstance(block, DiscBlock):
                nxt = seq[i + 1] if i + 1 < len(seq) else None
                relu_out = fuse and (nxt is None or isinstance(nxt, DiscBlock))   # the head starts with the same ReLU (:344)
                h = block(h, in_relu=relu_in, out_relu=relu_out)
                relu_in = relu_out
            else:
                h = block(h)
                relu_in = False
        h = A.SumHWFn.call(h, True)                                  # relu + sum over (H, W), fp32 [B, C]
        self._snb.clear()
        return ops.discriminator_head(self, h, label, adc_fake)

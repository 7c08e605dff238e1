"""BigGAN-Deep, StudioGAN flavour (reference ``src/models/big_resnet_deep_studiogan.py``): like the legacy variant but
* generator block: the skip is a learnable 1x1 convolution ``conv2d0`` (in -> out channels) instead of dropping
  channels (ref :58-78); it runs at LOW resolution here and is added, nearest-up-sampled, in conv2d4's epilogue (a 1x1
  convolution commutes with nearest up-sampling);
* discriminator block (ref :193-253): the skip is ``conv2d0`` (in -> out) whenever the block down-samples or changes
  width (no channel concat), and the main branch pools BEFORE its last activation.  The in-place ``nn.ReLU`` of the
  reference (src/config.py:486) rectifies the aliased skip tensor here too, so the skip path sees relu(x); pooling and
  the 1x1 skip convolution commute, so the skip convolution always runs at the lower resolution.
Constructor signatures, sub-module names and registration order follow the reference (identical state_dict keys and
seeded initialisation)."""
import torch.nn as nn

from .. import autograd_ops as A
from . import big_resnet_deep_legacy as legacy

D_IN = dict(legacy.D_IN)
D_IN["32"] = [1, 4, 4]
D_IN["512"] = [1, 1, 2, 4, 8, 8, 16]


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
        self.conv2d0 = MODULES.g_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, stride=1, padding=0)
        self.conv2d1 = MODULES.g_conv2d(in_channels=in_channels, out_channels=hid, kernel_size=1, stride=1, padding=0)
        self.conv2d2 = MODULES.g_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d3 = MODULES.g_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d4 = MODULES.g_conv2d(in_channels=hid, out_channels=out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, affine):
        main, side = A.ForkFn.call(x)
        h = self.conv2d1(self.bn1(main, affine, relu=True))
        h = self.conv2d2(self.bn2(h, affine, relu=True, up2=self.upsample))
        h = self.conv2d3(self.bn3(h, affine, relu=True))
        skip = self.conv2d0(side)                               # low resolution; up-sampled inside conv2d4's epilogue
        return self.conv2d4(self.bn4(h, affine, relu=True), residual=skip, res_up2=self.upsample)


class Generator(legacy.Generator):
    """Same network plan as the legacy generator (ref :81-189) with the block above."""
    _block = GenBlock

    def __init__(self, *args, **kwargs):
        orig = legacy.GenBlock
        legacy.GenBlock = GenBlock              # the legacy constructor instantiates ``GenBlock`` by module-level name
        try:
            super().__init__(*args, **kwargs)
        finally:
            legacy.GenBlock = orig


class DiscBlock(nn.Module):
    def __init__(self, in_channels, out_channels, MODULES, optblock, downsample=True, channel_ratio=4):
        super().__init__()
        self.optblock = optblock
        self.downsample = downsample
        hid = out_channels // channel_ratio
        self.ch_mismatch = in_channels != out_channels
        if self.optblock:
            assert self.downsample and self.ch_mismatch, "downsample and ch_mismatch should be True."
        self.activation = MODULES.d_act_fn
        self.conv2d1 = MODULES.d_conv2d(in_channels=in_channels, out_channels=hid, kernel_size=1, stride=1, padding=0)
        self.conv2d2 = MODULES.d_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d3 = MODULES.d_conv2d(in_channels=hid, out_channels=hid, kernel_size=3, stride=1, padding=1)
        self.conv2d4 = MODULES.d_conv2d(in_channels=hid, out_channels=out_channels, kernel_size=1, stride=1, padding=0)
        if self.ch_mismatch or self.downsample:
            self.conv2d0 = MODULES.d_conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, stride=1, padding=0)
        if self.downsample:
            self.average_pooling = nn.AvgPool2d(2)

    def forward(self, x):
        a0, px = A.DBlockEntryFn.call(x, self.downsample)       # relu(x) and its 2x2 average (skip source)
        h = self.conv2d1(a0, relu=True, premasked=True, mask_input=True)
        h = self.conv2d2(h, relu=True, premasked=True, mask_input=True)
        if self.downsample:
            h = self.conv2d3(h, mask_input=True)                # pool first, then the activation (ref :239-240)
            h = A.ReluFn.call(A.PoolFn.call(h))
            skip = self.conv2d0(px)
            return self.conv2d4(h, residual=skip)
        h = self.conv2d3(h, relu=True, premasked=True, mask_input=True)
        skip = self.conv2d0(px) if self.ch_mismatch else px
        return self.conv2d4(h, residual=skip, mask_input=True)


class Discriminator(legacy.Discriminator):
    def __init__(self, img_size, d_conv_dim, *args, **kwargs):
        orig_block, orig_in = legacy.DiscBlock, legacy.D_IN
        state = {"first": True}

        def make_block(in_channels, out_channels, MODULES, downsample=True):
            optblock = state["first"]
            state["first"] = False
            return DiscBlock(in_channels=in_channels, out_channels=out_channels, MODULES=MODULES, optblock=optblock,
                             downsample=downsample)
        legacy.DiscBlock, legacy.D_IN = make_block, D_IN
        try:
            super().__init__(img_size, d_conv_dim, *args, **kwargs)
        finally:
            legacy.DiscBlock, legacy.D_IN = orig_block, orig_in

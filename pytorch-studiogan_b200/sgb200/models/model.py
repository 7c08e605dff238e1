"""Model factory and data-parallel preparation (reference ``src/models/model.py:19-200``).

``load_generator_discriminator`` keeps the reference signature and its 8-tuple result.  ``prepare_parallel_training``
replaces torch's SyncBatchNorm conversion + DDP wrapping: sync-BN is a process-group attribute on the sgb200
BatchNorm2d (its statistics vectors are all-reduced over NCCL between the stats and apply kernels), and gradients are
averaged by one NCCL all-reduce per network over the flat gradient arena (utils/arena.py) — NVSwitch bandwidth makes the
223 MB generator gradient a sub-millisecond exchange, so no bucketing / overlap machinery is needed on one box.
"""
import copy
import importlib

import torch
import torch.distributed as dist

from ..utils import ops
from ..utils.ema import Ema

BACKBONES = ("big_resnet_deep_legacy", "big_resnet_deep_studiogan", "big_resnet", "resnet", "deep_conv")


def load_generator_discriminator(DATA, OPTIMIZATION, MODEL, STYLEGAN, MODULES, RUN, device, logger):
    if MODEL.backbone not in BACKBONES:
        raise NotImplementedError("backbone '%s' is outside the sgb200 hot-path scope (SURVEY.md section 8)" % MODEL.backbone)
    module = importlib.import_module("sgb200.models." + MODEL.backbone)
    Gen = module.Generator(z_dim=MODEL.z_dim, g_shared_dim=MODEL.g_shared_dim, img_size=DATA.img_size,
                           g_conv_dim=MODEL.g_conv_dim, apply_attn=MODEL.apply_attn, attn_g_loc=MODEL.attn_g_loc,
                           g_cond_mtd=MODEL.g_cond_mtd, num_classes=DATA.num_classes, g_init=MODEL.g_init,
                           g_depth=MODEL.g_depth, mixed_precision=RUN.mixed_precision, MODULES=MODULES, MODEL=MODEL).to(device)
    Dis = module.Discriminator(img_size=DATA.img_size, d_conv_dim=MODEL.d_conv_dim, apply_d_sn=MODEL.apply_d_sn,
                               apply_attn=MODEL.apply_attn, attn_d_loc=MODEL.attn_d_loc, d_cond_mtd=MODEL.d_cond_mtd,
                               aux_cls_type=MODEL.aux_cls_type, d_embed_dim=MODEL.d_embed_dim,
                               normalize_d_embed=MODEL.normalize_d_embed, num_classes=DATA.num_classes, d_init=MODEL.d_init,
                               d_depth=MODEL.d_depth, mixed_precision=RUN.mixed_precision, MODULES=MODULES, MODEL=MODEL).to(device)
    if MODEL.apply_g_ema:
        Gen_ema = copy.deepcopy(Gen)
        for m in Gen_ema.modules():           # deep-copied spectral-norm states must point at the copy
            if hasattr(m, "_sn"):
                m._sn.module = m
                m._sn.ws = None
        if hasattr(Gen_ema, "_snb"):
            Gen_ema._snb.net, Gen_ema._snb.mods = Gen_ema, None
        ema = Ema(source=Gen, target=Gen_ema, decay=MODEL.g_ema_decay, start_iter=MODEL.g_ema_start)
    else:
        Gen_ema, ema = None, None
    return Gen, None, None, Dis, Gen_ema, None, None, ema


def prepare_parallel_training(Gen, Gen_mapping, Gen_synthesis, Dis, Gen_ema, Gen_ema_mapping, Gen_ema_synthesis, MODEL,
                              world_size, distributed_data_parallel, synchronized_bn, apply_g_ema, device):
    if not distributed_data_parallel:
        return Gen, Gen_mapping, Gen_synthesis, Dis, Gen_ema, Gen_ema_mapping, Gen_ema_synthesis
    group = dist.new_group([w for w in range(world_size)])
    for net in (Gen, Dis, Gen_ema if apply_g_ema else None):
        if net is None:
            continue
        net.sgb_group = group
        net.sgb_world_size = world_size
        if synchronized_bn:
            for m in net.modules():
                if isinstance(m, ops.BatchNorm2d):
                    m.sync_group = group
        # identical replicas at start (DDP broadcasts rank 0's parameters and buffers at construction)
        for t in list(net.parameters()) + list(net.buffers()):
            dist.broadcast(t.data, src=0, group=group)
    return Gen, Gen_mapping, Gen_synthesis, Dis, Gen_ema, Gen_ema_mapping, Gen_ema_synthesis


def allreduce_gradients(net, optimizer=None):
    """Average gradients across ranks: ONE all-reduce over the flat gradient arena when the optimiser keeps one (the
    1/world scale is folded into the fused Adam launch), else flatten / reduce / scatter back."""
    group = getattr(net, "sgb_group", None)
    if group is None:
        return
    world = net.sgb_world_size
    if optimizer is not None and hasattr(optimizer, "all_reduce"):
        optimizer.all_reduce(group, world)
        return
    grads = [p.grad for p in net.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    for g, s in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(s)

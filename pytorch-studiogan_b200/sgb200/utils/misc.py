"""Mode / gradient toggles the step runtime relies on (src/utils/misc.py:158-267,345-364), re-stated for the sgb200
modules (which subclass the same torch base classes, so the isinstance tests are the reference's own)."""
import random

import numpy as np
import torch
import torch.nn as nn


class dummy_context_mgr():
    def __enter__(self):
        return None

    def __exit__(self, exc_type, exc_value, traceback):
        return False


def fix_seed(seed):
    random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.cuda.manual_seed(seed)
    np.random.seed(seed)


def toggle_grad(model, grad, num_freeze_layers=-1, is_stylegan=False):
    model = peel_model(model)
    if num_freeze_layers == -1:
        for p in model.parameters():
            p.requires_grad = grad
        return
    try:
        num_blocks = len(model.in_dims)
        assert num_freeze_layers < num_blocks, "cannot freeze the {nfl}th block > total {nb} blocks.".format(
            nfl=num_freeze_layers, nb=num_blocks)
        for name, param in model.named_parameters():
            param.requires_grad = grad
            for layer in range(num_freeze_layers):
                if "blocks.{layer}.".format(layer=layer) in name:
                    param.requires_grad = False
    except AttributeError:
        for p in model.parameters():
            p.requires_grad = grad


def set_bn_trainable(m):
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        m.train()


def untrack_bn_statistics(m):
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        m.track_running_stats = False


def track_bn_statistics(m):
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        m.track_running_stats = True


def set_deterministic_op_trainable(m):
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear, nn.Embedding)):
        m.train()


def make_GAN_trainable(Gen, Gen_ema, Dis):
    Gen.train()
    Gen.apply(track_bn_statistics)
    if Gen_ema is not None:
        Gen_ema.train()
        Gen_ema.apply(track_bn_statistics)
    Dis.train()
    Dis.apply(track_bn_statistics)


def make_GAN_untrainable(Gen, Gen_ema, Dis):
    """eval() everything, then put conv / linear / embedding back in train mode so that the spectral-norm power
    iteration keeps running during evaluation, exactly as the reference does (src/utils/misc.py:356-364)."""
    Gen.eval()
    Gen.apply(set_deterministic_op_trainable)
    if Gen_ema is not None:
        Gen_ema.eval()
        Gen_ema.apply(set_deterministic_op_trainable)
    Dis.eval()
    Dis.apply(set_deterministic_op_trainable)


def peel_model(model):
    return model.module if hasattr(model, "module") else model


def peel_models(Gen, Gen_ema, Dis):
    return peel_model(Gen), (peel_model(Gen_ema) if Gen_ema is not None else None), peel_model(Dis)


def reset_bn_statistics(m):
    if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
        m.reset_running_stats()


def apply_standing_statistics(generator, standing_max_batch, standing_step, DATA, MODEL, LOSS, OPTIMIZATION, RUN, STYLEGAN=None,
                              device="cuda", global_rank=0, logger=None):
    """``-std_stat`` evaluation mode (src/utils/misc.py:301-333): reset the generator's BatchNorm running statistics and
    re-accumulate them over ``standing_step`` forward passes of random batch sizes (train mode, no gradient), then
    switch to eval.  Host logic only; the forward passes run on the sgb200 kernels like any other."""
    from . import sample
    generator.train()
    generator.apply(reset_bn_statistics)
    if global_rank == 0 and logger is not None:
        logger.info("Accumulate statistics of batchnorm layers to improve generation performance.")
    world = getattr(OPTIMIZATION, "world_size", 1)
    with torch.no_grad():
        for _ in range(standing_step):
            per_gpu = max(1, standing_max_batch // world)
            if RUN.distributed_data_parallel:
                rand_batch_size = random.randint(1, per_gpu)
                # sync-BN divides the all-reduced sums by (local count x world size): every rank must draw the SAME size.
                # (torch's SyncBatchNorm exchanges per-rank counts; here rank 0's draw is broadcast instead.)
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    t = torch.tensor([rand_batch_size], device=device if dist.get_backend() == "nccl" else "cpu")
                    dist.broadcast(t, src=0)
                    rand_batch_size = int(t.item())
            else:
                rand_batch_size = random.randint(1, per_gpu) * world
            sample.generate_images(z_prior=MODEL.z_prior, truncation_factor=-1, batch_size=rand_batch_size, z_dim=MODEL.z_dim,
                                   num_classes=DATA.num_classes, y_sampler="totally_random", radius="N/A", generator=generator,
                                   discriminator=None, is_train=True, LOSS=LOSS, RUN=RUN, MODEL=MODEL, device=device)
    generator.eval()


def identity(x):
    return x

"""Numeric check of the data-parallel path: W ranks x (B / W) images must reproduce ONE rank on the B-image batch.

What is exercised (reference semantics: ``src/models/model.py:157-200`` = torch SyncBatchNorm + DDP gradient averaging):
  * sync-BN forward  : all-reduce of [sum x, sum x^2] between the statistics and the apply kernels
  * sync-BN backward : all-reduce of [S1, S2] between the reduce and the apply kernels
  * gradients        : one all-reduce over the flat gradient arena, 1 / W folded into the Adam launch
  * running statistics (momentum update with the GLOBAL count)
on a small BigGAN-Deep (32x32, conv_dim 8): every rank builds the same two replicas, runs a discriminator phase and a
generator phase sharded (its slice of the global batch, groups attached) and unsharded (the whole batch, no groups),
and compares.  Both sides use the same kernels, so the differences are fp32 summation order plus the bf16 roundings it
flips; the stated bound is the single-GPU gradient tolerance of tests/test_gpu_parity.py (1e-1 relative L2, worst
parameter), typically met by a wide margin.
"""
import copy

import torch
import torch.distributed as dist


def _build(device, seed=1234):
    import importlib
    from .. import config as C
    deep = importlib.import_module("sgb200.models.big_resnet_deep_legacy")
    torch.manual_seed(seed)
    M = C.make_modules(True, True, "cBN", "big_resnet_deep_legacy")
    MODEL = C._Section(info_type="N/A", g_info_injection="N/A")
    G = deep.Generator(z_dim=16, g_shared_dim=16, img_size=32, g_conv_dim=8, apply_attn=False, attn_g_loc=[2], g_cond_mtd="cBN",
                       num_classes=5, g_init="ortho", g_depth=1, mixed_precision=False, MODULES=M, MODEL=MODEL)
    D = deep.Discriminator(img_size=32, d_conv_dim=8, apply_d_sn=True, apply_attn=False, attn_d_loc=[1], d_cond_mtd="PD",
                           aux_cls_type="W/O", d_embed_dim="N/A", normalize_d_embed=False, num_classes=5, d_init="ortho",
                           d_depth=1, mixed_precision=False, MODULES=M, MODEL=MODEL)
    return G.to(device).train(), D.to(device).train()


def _phases(G, D, optG, optD, z, yf, real, yr, world_for_reduce):
    """One discriminator update + one generator update (hinge), gradients left in the arenas; returns both loss values."""
    from . import losses, misc
    from ..models import model as model_lib
    misc.toggle_grad(G, False)
    misc.toggle_grad(D, True)
    optD.zero_grad()
    with torch.no_grad():
        fake = G(z, yf)
    d_loss = losses.d_wasserstein(D(real, yr)["adv_output"], D(fake, yf)["adv_output"])
    d_loss.backward()
    model_lib.allreduce_gradients(D, optD)
    gD = optD.grads.flat.clone() * optD.grad_scale
    optD.step()
    misc.toggle_grad(D, False)
    misc.toggle_grad(G, True)
    optG.zero_grad()
    g_loss = losses.g_wasserstein(D(G(z, yf), yf)["adv_output"])
    g_loss.backward()
    model_lib.allreduce_gradients(G, optG)
    gG = optG.grads.flat.clone() * optG.grad_scale
    optG.step()
    return gD, gG, d_loss.detach(), g_loss.detach()


def multirank_parity_check(device, per_rank=4):
    """Run on every rank of an initialised NCCL process group (world >= 2).  Returns a dict of worst relative-L2 errors
    (sharded vs unsharded) and ``ok``."""
    from ..models import model as model_lib
    from .optim import ArenaAdam
    world, rank = dist.get_world_size(), dist.get_rank()
    Gs, Ds = _build(device)
    Gf, Df = copy.deepcopy(Gs), copy.deepcopy(Ds)
    for net in (Gf, Df):                                   # deep-copied spectral-norm states must point at the copy
        for m in net.modules():
            if hasattr(m, "_sn"):
                m._sn.module, m._sn.ws = m, None
        net._snb.net, net._snb.mods = net, None
    model_lib.prepare_parallel_training(Gs, None, None, Ds, None, None, None, None, world, True, True, False, device)
    B = per_rank * world
    g = torch.Generator().manual_seed(99)
    z, yf = torch.randn(B, 16, generator=g).to(device), torch.randint(0, 5, (B,), generator=g).to(device)
    real, yr = (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1).to(device), torch.randint(0, 5, (B,), generator=g).to(device)
    sl = slice(rank * per_rank, (rank + 1) * per_rank)
    mk = lambda net, lr: ArenaAdam(net, lr, betas=(0.0, 0.999), eps=1e-6)
    gD_s, gG_s, dl_s, gl_s = _phases(Gs, Ds, mk(Gs, 5e-5), mk(Ds, 2e-4), z[sl], yf[sl], real[sl], yr[sl], world)
    gD_f, gG_f, dl_f, gl_f = _phases(Gf, Df, mk(Gf, 5e-5), mk(Df, 2e-4), z, yf, real, yr, 1)
    # the sharded loss is this rank's mean; its average over ranks is the global mean
    losses = torch.stack([dl_s, gl_s])
    dist.all_reduce(losses)
    losses /= world

    def rel(a, b):
        return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

    out = {"world": world, "global_batch": B,
           "d_grad": rel(gD_s, gD_f), "g_grad": rel(gG_s, gG_f),
           "d_loss": abs(float(losses[0]) - float(dl_f)) / (abs(float(dl_f)) + 1e-6),
           "g_loss": abs(float(losses[1]) - float(gl_f)) / (abs(float(gl_f)) + 1e-6)}
    bn = 0.0
    for (n, b), (_, c) in zip(Gs.named_buffers(), Gf.named_buffers()):
        if "running_" in n:
            bn = max(bn, rel(b, c))
    out["bn_running_stats"] = bn
    t = torch.tensor([out["d_grad"], out["g_grad"], out["d_loss"], out["g_loss"], bn], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.update(dict(zip(["d_grad", "g_grad", "d_loss", "g_loss", "bn_running_stats"], [float(v) for v in t])))
    out["ok"] = bool(out["d_grad"] < 1e-1 and out["g_grad"] < 1e-1 and out["d_loss"] < 5e-2 and out["g_loss"] < 5e-2
                     and out["bn_running_stats"] < 1e-2)
    return out

"""world_size-2 gloo tests (CPU) of the multi-rank host logic: replica broadcast, gradient averaging, the sync-BN
statistics algebra (sum / sum-of-squares all-reduce == statistics of the concatenated batch) and batch sharding."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "pytorch-studiogan_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sgb200.models import model as M
    torch.manual_seed(100 + rank)                      # different replicas before the broadcast
    net = nn.Sequential(nn.Linear(4, 3), nn.BatchNorm1d(3))
    M.prepare_parallel_training(net, None, None, net, None, None, None, None, world, True, False, False, "cpu")
    w0 = [p.detach().clone() for p in net.parameters()]
    # gradient averaging
    torch.manual_seed(7 + rank)
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    local = [p.grad.clone() for p in net.parameters()]
    M.allreduce_gradients(net)
    avg = [p.grad.clone() for p in net.parameters()]
    # sync-BN algebra: all-reduce of [sum, sumsq] and count gives the statistics of the global batch
    torch.manual_seed(11 + rank)
    x = torch.randn(5 + rank, 3) * (1 + rank) + rank
    stats = torch.stack([x.sum(0), (x * x).sum(0)])
    cnt = torch.tensor([float(x.shape[0])])
    dist.all_reduce(stats)
    dist.all_reduce(cnt)
    mean = stats[0] / cnt
    var = stats[1] / cnt - mean * mean
    q.put((rank, [w.numpy() for w in w0], [g.numpy() for g in local], [g.numpy() for g in avg], x.numpy(), mean.numpy(), var.numpy()))
    dist.destroy_process_group()


def test_two_rank_gloo_replicas_gradients_and_syncbn_stats():
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w_a, g_a, avg_a, x_a, mean_a, var_a), (_, w_b, g_b, avg_b, x_b, mean_b, var_b) = out
    for a, b in zip(w_a, w_b):
        assert np.array_equal(a, b)                                   # rank 0's replica everywhere
    for ga, gb, aa, ab in zip(g_a, g_b, avg_a, avg_b):
        np.testing.assert_allclose(aa, (ga + gb) / 2, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ab, aa, rtol=0, atol=0)
    allx = np.concatenate([x_a, x_b], 0)
    np.testing.assert_allclose(mean_a, allx.mean(0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var_a, allx.var(0), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(mean_b, mean_a)


def test_strong_scaling_batch_sharding():
    # bench.py / src/loader.py:162: per-rank batch = global // world, the whole-job value uses the global batch
    for world in (1, 2, 4, 8):
        assert 256 % world == 0 and (256 // world) * world == 256


def _d2dce_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "pytorch-studiogan_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sgb200.utils import losses
    g = torch.Generator().manual_seed(5)
    embed_all, proxy_all = torch.randn(8, 6, generator=g), torch.randn(8, 6, generator=g)
    label_all = torch.randint(0, 3, (8,), generator=g)
    sl = slice(4 * rank, 4 * rank + 4)
    embed = embed_all[sl].clone().requires_grad_(True)
    proxy = proxy_all[sl].clone().requires_grad_(True)
    out = {}
    for name, mod in (("d2dce", losses.Data2DataCrossEntropyLoss(3, 0.5, 0.98, DDP=True)),
                      ("c2", losses.ConditionalContrastiveLoss(3, 0.5, DDP=True))):
        embed.grad = proxy.grad = None
        loss = mod(embed=embed, proxy=proxy, label=label_all[sl])
        loss.backward()
        out[name] = (loss.item(), embed.grad.numpy().copy(), proxy.grad.numpy().copy())
    q.put((rank, out))
    dist.destroy_process_group()


def test_two_rank_gloo_conditioning_losses_gather_with_gradient():
    """DDP path of the contrastive conditioning losses (src/utils/losses.py:84-88,140-144): every rank all-gathers embed /
    proxy / label through GatherLayer, so each rank's loss equals the single-process loss on the concatenated batch and
    receives the gradient of its own shard."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "pytorch-studiogan_b200"))
    from sgb200.utils import losses
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_d2dce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(5)
    embed_all = torch.randn(8, 6, generator=g).requires_grad_(True)
    proxy_all = torch.randn(8, 6, generator=g).requires_grad_(True)
    label_all = torch.randint(0, 3, (8,), generator=g)
    for name, mod in (("d2dce", losses.Data2DataCrossEntropyLoss(3, 0.5, 0.98, DDP=False)),
                      ("c2", losses.ConditionalContrastiveLoss(3, 0.5, DDP=False))):
        embed_all.grad = proxy_all.grad = None
        ref = mod(embed=embed_all, proxy=proxy_all, label=label_all)
        ref.backward()
        for rank in (0, 1):
            val, de, dp = res[rank][name]
            np.testing.assert_allclose(val, ref.item(), rtol=1e-6)
            sl = slice(4 * rank, 4 * rank + 4)
            # GatherLayer hands each rank the gradient of ITS OWN copy of the gathered loss (src/utils/losses.py:19-37)
            np.testing.assert_allclose(de, embed_all.grad[sl].numpy(), rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(dp, proxy_all.grad[sl].numpy(), rtol=1e-5, atol=1e-7)


# ----------------------------------------------------------------------------------------------------------------------
# sgb200.utils.ops.BatchNorm2d with a sync group over gloo.  The module's forward / backward host logic (autograd_ops.
# BNActFn: where the all-reduces sit, which count divides the sums, what reaches the running statistics) runs unmodified;
# only the five CUDA kernel wrappers are replaced by fp32 torch restatements of their documented arithmetic
# (include/sgb200.h, "Batch-norm family"), since there is no GPU here.
# ----------------------------------------------------------------------------------------------------------------------
def _install_bn_kernel_shims(K):
    def bn_stats(x):
        xf = x.float()
        return torch.stack([xf.sum((0, 2, 3)), (xf * xf).sum((0, 2, 3))])

    def bn_finalize(stats, count, rm, rv, momentum, eps, use_batch_stats, track, mode, gain, bias, nb, C, device):
        if use_batch_stats:
            mean = stats[0] / count
            var = stats[1] / count - mean * mean
            if track:
                rm.mul_(1 - momentum).add_(momentum * mean)
                rv.mul_(1 - momentum).add_(momentum * var * count / max(count - 1.0, 1.0))
        else:
            mean, var = rm.clone(), rv.clone()
        rstd = torch.rsqrt(var + eps)
        g = (1.0 + gain) if mode == 0 else (gain.reshape(1, C) if mode == 1 else torch.ones(1, C))
        b = bias if mode == 0 else (bias.reshape(1, C) if mode == 1 else torch.zeros(1, C))
        scale = (rstd.reshape(1, C) * g).expand(nb, C).contiguous()
        shift = (b - mean.reshape(1, C) * scale).expand(nb, C).contiguous()
        return mean, rstd, scale, shift

    def scale_shift_act(x, scale, shift, per_image, relu, up2):
        y = x.float() * scale.reshape(-1, x.shape[1], 1, 1) + shift.reshape(-1, x.shape[1], 1, 1)
        return torch.relu(y) if relu else y

    def _dz(dy, x, scale, shift, relu):
        C = x.shape[1]
        y = x.float() * scale.reshape(-1, C, 1, 1) + shift.reshape(-1, C, 1, 1)
        return dy.float() * (y > 0) if relu else dy.float()

    def bn_bwd_reduce(dy, x, scale, shift, per_image, mean, rstd, relu, up2):
        C = x.shape[1]
        dz = _dz(dy, x, scale, shift, relu)
        xh = (x.float() - mean.reshape(1, C, 1, 1)) * rstd.reshape(1, C, 1, 1)
        s1, s2 = dz.sum((2, 3)), (dz * xh).sum((2, 3))
        g1 = scale / rstd.reshape(1, C)
        return torch.stack([s1, s2]), torch.stack([(g1 * s1).sum(0), (g1 * s2).sum(0)])

    def bn_bwd_apply(dy, x, scale, shift, per_image, mean, rstd, S12, count, relu, up2, use_batch_stats):
        C = x.shape[1]
        dz = _dz(dy, x, scale, shift, relu)
        dx = scale.reshape(-1, C, 1, 1) * dz
        if use_batch_stats:
            xh = (x.float() - mean.reshape(1, C, 1, 1)) * rstd.reshape(1, C, 1, 1)
            dx = dx - rstd.reshape(1, C, 1, 1) * (S12[0].reshape(1, C, 1, 1) + xh * S12[1].reshape(1, C, 1, 1)) / count
        return dx

    K.bn_stats, K.bn_finalize, K.scale_shift_act, K.bn_bwd_reduce, K.bn_bwd_apply = (bn_stats, bn_finalize, scale_shift_act,
                                                                                     bn_bwd_reduce, bn_bwd_apply)
    K.as_nhwc = lambda t: t


def _syncbn_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "pytorch-studiogan_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sgb200 import kernels as K
    from sgb200.utils import ops
    _install_bn_kernel_shims(K)
    g = torch.Generator().manual_seed(3)
    x_all = torch.randn(6, 5, 4, 4, generator=g) * 1.7 + 0.4
    w_all = torch.randn(6, 5, 4, 4, generator=g)
    bn = ops.batchnorm_2d(5)
    with torch.no_grad():
        bn.weight.copy_(torch.linspace(0.5, 1.5, 5))
        bn.bias.copy_(torch.linspace(-0.2, 0.2, 5))
    bn.sync_group = dist.group.WORLD
    bn.train()
    sl = slice(3 * rank, 3 * rank + 3)
    x = x_all[sl].permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)       # NHWC in memory
    y = bn(x, relu=True)
    (y * w_all[sl]).sum().backward()
    q.put((rank, y.detach().numpy(), x.grad.numpy(), bn.weight.grad.numpy(), bn.bias.grad.numpy(), bn.running_mean.numpy(),
           bn.running_var.numpy(), int(bn.num_batches_tracked)))
    dist.destroy_process_group()


def test_two_rank_gloo_sgb200_batchnorm_equals_one_rank_global_batch():
    """ops.BatchNorm2d(sync_group) on 2 ranks x 3 images == torch's BatchNorm on the 6-image batch: output, input gradient,
    running statistics (momentum 0.1, unbiased variance over the GLOBAL count, torch/nn/modules/_functions.py:7-212);
    the affine gradients are per-rank partial sums whose all-reduce (the gradient arena's) gives the global gradient."""
    import numpy as np
    import torch.nn.functional as F
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + os.getpid() % 2000
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(3)
    x_all = (torch.randn(6, 5, 4, 4, generator=g) * 1.7 + 0.4).double().requires_grad_(True)
    w_all = torch.randn(6, 5, 4, 4, generator=g).double()
    wt = torch.linspace(0.5, 1.5, 5).double().requires_grad_(True)
    bs = torch.linspace(-0.2, 0.2, 5).double().requires_grad_(True)
    rm, rv = torch.zeros(5).double(), torch.ones(5).double()
    y = torch.relu(F.batch_norm(x_all, rm, rv, wt, bs, True, 0.1, 1e-4))
    (y * w_all).sum().backward()
    y_got = np.concatenate([res[0][1], res[1][1]], 0)
    dx_got = np.concatenate([res[0][2], res[1][2]], 0)
    np.testing.assert_allclose(y_got, y.detach().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dx_got, x_all.grad.numpy(), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(res[0][3] + res[1][3], wt.grad.numpy(), rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(res[0][4] + res[1][4], bs.grad.numpy(), rtol=1e-3, atol=1e-4)
    for r in res:
        np.testing.assert_allclose(r[5], rm.numpy(), rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(r[6], rv.numpy(), rtol=1e-4, atol=1e-6)
        assert r[7] == 1

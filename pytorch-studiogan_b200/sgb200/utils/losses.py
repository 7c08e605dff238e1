"""Adversarial losses with the reference's names and signatures (src/utils/losses.py:197-239): scalar reductions over the
[B] logit vector coming out of the discriminator head, plus GatherLayer (:19-37) and the WGAN-GP penalty entry point."""
import torch
import torch.distributed as dist
import torch.nn.functional as F


class GatherLayer(torch.autograd.Function):
    """all_gather with a gradient (src/utils/losses.py:19-37): each rank receives the gradient of its own shard."""

    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        out = [torch.zeros_like(input) for _ in range(dist.get_world_size())]
        dist.all_gather(out, input)
        return tuple(out)

    @staticmethod
    def backward(ctx, *grads):
        (input,) = ctx.saved_tensors
        g = torch.zeros_like(input)
        g[:] = grads[dist.get_rank()]
        return g


def d_vanilla(d_logit_real, d_logit_fake, DDP=False):
    return torch.mean(F.softplus(-d_logit_real)) + torch.mean(F.softplus(d_logit_fake))


def g_vanilla(d_logit_fake, DDP=False):
    return torch.mean(F.softplus(-d_logit_fake))


def d_logistic(d_logit_real, d_logit_fake, DDP=False):
    return (F.softplus(-d_logit_real) + F.softplus(d_logit_fake)).mean()


def g_logistic(d_logit_fake, DDP=False):
    return F.softplus(-d_logit_fake).mean()


def d_ls(d_logit_real, d_logit_fake, DDP=False):
    return (0.5 * (d_logit_real - 1.0) ** 2 + 0.5 * d_logit_fake ** 2).mean()


def g_ls(d_logit_fake, DDP=False):
    return (0.5 * (d_logit_fake - 1.0) ** 2).mean()


def d_hinge(d_logit_real, d_logit_fake, DDP=False):
    return torch.mean(F.relu(1. - d_logit_real)) + torch.mean(F.relu(1. + d_logit_fake))


def g_hinge(d_logit_fake, DDP=False):
    return -torch.mean(d_logit_fake)


def d_wasserstein(d_logit_real, d_logit_fake, DDP=False):
    return torch.mean(d_logit_fake - d_logit_real)


def g_wasserstein(d_logit_fake, DDP=False):
    return -torch.mean(d_logit_fake)


def cal_deriv(inputs, outputs, device=None):
    from . import gp
    return gp.cal_deriv(inputs, outputs, device)


def cal_grad_penalty(real_images, real_labels, fake_images, discriminator, device=None, alpha=None):
    """src/utils/losses.py:301-316; see utils/gp.py for how the second-order term is evaluated."""
    from . import gp
    return gp.cal_grad_penalty(real_images, real_labels, fake_images, discriminator, device, alpha)

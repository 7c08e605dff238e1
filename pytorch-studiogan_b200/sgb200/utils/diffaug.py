"""DiffAugment (reference ``src/utils/diffaug.py``, policy "color,translation,cutout") as ONE device kernel per direction.

The reference runs seven tensor-op passes per call; per sample the whole chain is an affine colour map followed by an integer
shift and a box mask, so ``sgb_diffaug_fwd`` gathers the result in one pass from the per-sample parameters and the per-sample
input mean, and ``sgb_diffaug_bwd`` is its adjoint (the generator phase back-propagates through the augmentation into the
fake images).  ``draw_params`` consumes the device generator exactly in the reference's order (brightness, saturation,
contrast: ``torch.rand``; translation x / y, cutout x / y: ``torch.randint``), so a seeded run augments identically.
"""
import torch

from .. import _lib as L

_POLICIES = ("color", "translation", "cutout")


def draw_params(B, H, W, policy, device, dtype=torch.float32):
    """[B, 7] = (brightness offset, saturation factor, contrast factor, shift_h, shift_w, cut_h0, cut_w0) drawn in the order
    the reference's AUGMENT_FNS consume the RNG for ``policy`` (src/utils/diffaug.py:47-100)."""
    p = torch.zeros((B, 7), device=device, dtype=dtype)
    p[:, 1:3] = 1.0
    for name in policy.split(","):
        if name == "color":
            p[:, 0] = torch.rand(B, 1, 1, 1, dtype=dtype, device=device).view(B) - 0.5
            p[:, 1] = torch.rand(B, 1, 1, 1, dtype=dtype, device=device).view(B) * 2
            p[:, 2] = torch.rand(B, 1, 1, 1, dtype=dtype, device=device).view(B) + 0.5
        elif name == "translation":
            sh, sw = int(H * 0.125 + 0.5), int(W * 0.125 + 0.5)
            p[:, 3] = torch.randint(-sh, sh + 1, size=[B, 1, 1], device=device).view(B).to(dtype)
            p[:, 4] = torch.randint(-sw, sw + 1, size=[B, 1, 1], device=device).view(B).to(dtype)
        elif name == "cutout":
            ch, cw = int(H * 0.5 + 0.5), int(W * 0.5 + 0.5)
            p[:, 5] = torch.randint(0, H + (1 - ch % 2), size=[B, 1, 1], device=device).view(B).to(dtype)
            p[:, 6] = torch.randint(0, W + (1 - cw % 2), size=[B, 1, 1], device=device).view(B).to(dtype)
        else:
            raise NotImplementedError("DiffAugment policy '%s'" % name)
    return p


class _DiffAugFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, params, flags):
        x = x.contiguous()
        B, C, H, W = x.shape
        assert C == 3 and x.dtype == torch.float32
        mean_x = None
        if flags[0]:
            mean_x = torch.empty(B, device=x.device, dtype=torch.float32)
            L.call("sgb_sample_mean", L.ptr(x), B, C * H * W, L.ptr(mean_x), L.stream_ptr())
        y = torch.empty_like(x)
        L.call("sgb_diffaug_fwd", L.ptr(x), L.ptr(params), L.ptr(mean_x), L.ptr(y), B, H, W, flags[0], flags[1], flags[2], L.stream_ptr())
        ctx.flags = flags
        ctx.save_for_backward(params)
        return y

    @staticmethod
    def backward(ctx, dy):
        (params,) = ctx.saved_tensors
        dy = dy.contiguous()
        B, C, H, W = dy.shape
        dx = torch.empty_like(dy)
        ws = torch.empty(B, device=dy.device, dtype=torch.float32)
        f = ctx.flags
        L.call("sgb_diffaug_bwd", L.ptr(dy), L.ptr(params), L.ptr(dx), L.ptr(ws), B, H, W, f[0], f[1], f[2], L.stream_ptr())
        return dx, None, None


def apply_diffaug(x, policy="color,translation,cutout", channels_first=True, params=None):
    """Same call contract as the reference's ``apply_diffaug``; ``params`` (optional [B, 7]) replaces the random draw."""
    if not policy:
        return x
    if not channels_first:
        x = x.permute(0, 3, 1, 2)
    names = policy.split(",")
    if any(n not in _POLICIES for n in names):
        raise NotImplementedError("DiffAugment policy '%s'" % policy)
    B, _, H, W = x.shape
    if params is None:
        params = draw_params(B, H, W, policy, x.device)
    flags = (1 if "color" in names else 0, 1 if "translation" in names else 0, 1 if "cutout" in names else 0)
    y = _DiffAugFn.apply(x, params.to(torch.float32).contiguous(), flags)
    if not channels_first:
        y = y.permute(0, 2, 3, 1)
    return y.contiguous()

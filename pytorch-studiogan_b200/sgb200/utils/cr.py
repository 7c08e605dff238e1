"""Consistency-regularisation augmentation (reference ``src/utils/cr.py``): random horizontal flip (p = 0.5) followed by a
random translation of up to 1/8 of the image with reflect padding -- one gather kernel (``sgb_cr_aug``).  RNG order as the
reference: the flip probabilities come from the HOST generator (``torch.FloatTensor(n, 1).uniform_``), the two shift
vectors from the device generator (``torch.randint(..., device=x.device)``).  Used on detached images only (CR / bCR terms of
the discriminator phase), so there is no backward."""
import torch

from .. import _lib as L


def draw_params(B, H, W, device, flip=True, translation=True):
    f = tx = ty = None
    if flip:
        f = (torch.FloatTensor(B, 1).uniform_(0.0, 1.0).to(device) < 0.5).view(B).to(torch.uint8)
    if translation:
        mx, my = int(H * (1 / 8)), int(W * (1 / 8))
        tx = torch.randint(-mx, mx + 1, size=[B, 1, 1], device=device).view(B).to(torch.int32)
        ty = torch.randint(-my, my + 1, size=[B, 1, 1], device=device).view(B).to(torch.int32)
    return f, tx, ty


def apply_cr_aug(x, flip=True, translation=True, params=None):
    if not (flip or translation):
        return x
    x = x.detach().contiguous()
    B, C, H, W = x.shape
    f, tx, ty = params if params is not None else draw_params(B, H, W, x.device, flip, translation)
    y = torch.empty_like(x)
    L.call("sgb_cr_aug", L.ptr(x), L.ptr(f), L.ptr(tx), L.ptr(ty), L.ptr(y), B, C, H, W, L.stream_ptr())
    return y

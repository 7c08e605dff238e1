"""Generator weight EMA (src/utils/ema.py:10-40): p_ema <- lerp(p, p_ema, decay), decay = 0 before ``start_iter``;
float buffers are lerped the same way, ``num_batches_tracked`` is copied.

Source and target generators keep their parameters in identically laid-out flat arenas, so the parameter update is ONE
library launch (the reference, and a naive port, launch two kernels per tensor: 339 tensors for BigGAN-Deep); the
spectral-norm u / v buffers are lerped through the two flat arenas of the batched-SN pass, the remaining float buffers
(BatchNorm running statistics) per tensor."""
import torch

from .. import kernels as K
from .arena import param_arena


class Ema(object):
    def __init__(self, source, target, decay=0.9999, start_iter=0):
        self.source = source
        self.target = target
        self.decay = decay
        self.start_iter = start_iter
        with torch.no_grad():
            for p_ema, p in zip(self.target.parameters(), self.source.parameters()):
                p_ema.copy_(p)
            for b_ema, b in zip(self.target.buffers(), self.source.buffers()):
                b_ema.copy_(b)
        self.src_arena = self.tgt_arena = None

    def _arenas(self):
        p0 = next(self.source.parameters())
        if not p0.is_cuda:
            return False
        self.src_arena, self.tgt_arena = param_arena(self.source), param_arena(self.target)
        return True

    def update(self, iter=None):
        decay = 0.0 if (iter >= 0 and iter < self.start_iter) else self.decay
        with torch.no_grad():
            if self._arenas():
                K.ema_lerp(self.tgt_arena.flat, self.src_arena.flat, decay)
            else:
                for p_ema, p in zip(self.target.parameters(), self.source.parameters()):
                    K.ema_lerp(p_ema.data, p.data, decay)
            s_snb, t_snb = getattr(self.source, "_snb", None), getattr(self.target, "_snb", None)
            flat_uv = (s_snb is not None and t_snb is not None and s_snb.mods is not None and t_snb.mods is not None
                       and s_snb.u_flat.numel() == t_snb.u_flat.numel() and s_snb.v_flat.numel() == t_snb.v_flat.numel())
            if flat_uv:
                K.ema_lerp(t_snb.u_flat, s_snb.u_flat, decay)
                K.ema_lerp(t_snb.v_flat, s_snb.v_flat, decay)
            for (name, b_ema), (_, b) in zip(self.target.named_buffers(), self.source.named_buffers()):
                if "num_batches_tracked" in name or not b.is_floating_point():
                    b_ema.copy_(b)
                elif flat_uv and (name.endswith("weight_u") or name.endswith("weight_v")) and _in_flat(b, s_snb):
                    continue
                else:
                    K.ema_lerp(b_ema, b, decay)


def _in_flat(b, snb):
    """True if buffer ``b`` is a view into the batched-SN arenas (then the flat lerp above already covered it)."""
    for flat in (snb.u_flat, snb.v_flat):
        lo = flat.data_ptr()
        if lo <= b.data_ptr() < lo + flat.numel() * 4:
            return True
    return False

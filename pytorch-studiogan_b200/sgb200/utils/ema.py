"""Generator weight EMA (src/utils/ema.py:10-40): p_ema <- lerp(p, p_ema, decay), decay = 0 before ``start_iter``;
float buffers are lerped the same way, ``num_batches_tracked`` is copied.  One library launch per tensor."""
import torch

from .. import kernels as K


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

    def update(self, iter=None):
        decay = 0.0 if (iter >= 0 and iter < self.start_iter) else self.decay
        with torch.no_grad():
            for p_ema, p in zip(self.target.parameters(), self.source.parameters()):
                K.ema_lerp(p_ema.data, p.data, decay)
            for (name, b_ema), (_, b) in zip(self.target.named_buffers(), self.source.named_buffers()):
                if "num_batches_tracked" in name or not b.is_floating_point():
                    b_ema.copy_(b)
                else:
                    K.ema_lerp(b_ema, b, decay)

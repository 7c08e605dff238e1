"""Generator weight EMA (src/utils/ema.py:10-40): p_ema <- lerp(p, p_ema, decay), decay = 0 before ``start_iter``;
float buffers are lerped the same way, ``num_batches_tracked`` is copied.

Source and target generators keep their parameters in identically laid-out flat arenas, so the parameter update is ONE
library launch (the reference, and a naive port, launch two kernels per tensor: 339 tensors for BigGAN-Deep); the
spectral-norm u / v buffers are lerped through the two flat arenas of the batched-SN pass, the remaining float buffers
(BatchNorm running statistics) through one more flat arena and the integer counters through another: 4 lerps + 1 copy
per update instead of ~400 launches."""
import torch

from .. import kernels as K
from .arena import BufferArena, param_arena


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

    def _uv_flats(self):
        """Both generators' batched-SN arenas (built on demand: the EMA copy may never have run a forward pass yet)."""
        s_snb, t_snb = getattr(self.source, "_snb", None), getattr(self.target, "_snb", None)
        if s_snb is None or t_snb is None:
            return None
        dev = next(self.source.parameters()).device
        for snb, net in ((s_snb, self.source), (t_snb, self.target)):
            if snb.mods is None or snb.device != dev or any(w.data_ptr() != p for w, p in zip(snb.params, snb.ptrs)):
                snb._build(dev)
        if s_snb.u_flat.numel() != t_snb.u_flat.numel() or s_snb.v_flat.numel() != t_snb.v_flat.numel():
            return None
        return s_snb, t_snb

    def _buffer_arenas(self, uv):
        skip_s = (lambda b: _in_flat(b, uv[0])) if uv else (lambda b: False)
        skip_t = (lambda b: _in_flat(b, uv[1])) if uv else (lambda b: False)
        sb, tb = getattr(self, "_sbuf", None), getattr(self, "_tbuf", None)
        if sb is None or not (sb.intact() and tb.intact()):
            self._sbuf, self._tbuf = BufferArena(self.source, skip_s), BufferArena(self.target, skip_t)
        return self._sbuf, self._tbuf

    def update(self, iter=None):
        decay = 0.0 if (iter >= 0 and iter < self.start_iter) else self.decay
        with torch.no_grad():
            if not self._arenas():
                raise RuntimeError("sgb200 Ema.update needs the generators on a CUDA device (no host fallback)")
            K.ema_lerp(self.tgt_arena.flat, self.src_arena.flat, decay)
            uv = self._uv_flats()
            if uv is not None:
                K.ema_lerp(uv[1].u_flat, uv[0].u_flat, decay)
                K.ema_lerp(uv[1].v_flat, uv[0].v_flat, decay)
            sb, tb = self._buffer_arenas(uv)
            if sb.fflat is not None and tb.fflat is not None and sb.fflat.numel() == tb.fflat.numel():
                K.ema_lerp(tb.fflat, sb.fflat, decay)
            else:
                for b_ema, b in zip(tb.floats, sb.floats):
                    K.ema_lerp(b_ema, b, decay)
            if sb.iflat is not None and tb.iflat is not None:
                tb.iflat.copy_(sb.iflat)
            else:
                for b_ema, b in zip(tb.ints, sb.ints):
                    b_ema.copy_(b)
            for b_ema, b in zip(tb.other, sb.other):
                b_ema.copy_(b.float().lerp(b_ema.float(), decay).to(b_ema.dtype))


def _in_flat(b, snb):
    """True if buffer ``b`` is a view into the batched-SN arenas (then the flat lerp above already covered it)."""
    for flat in (snb.u_flat, snb.v_flat):
        lo = flat.data_ptr()
        if lo <= b.data_ptr() < lo + flat.numel() * 4:
            return True
    return False

"""Flat parameter storage: re-points every parameter of a module at a slice of one contiguous fp32 buffer (values,
shapes, names and state_dict are unchanged).  One arena per network lets whole-network updates (EMA lerp, fused Adam,
gradient all-reduce) run as a single launch instead of one per tensor."""
import torch


def _flatten(tensors, align=64):
    total = 0
    offs = []
    for t in tensors:
        offs.append(total)
        total += (t.numel() + align - 1) // align * align
    flat = torch.zeros(max(total, 1), device=tensors[0].device, dtype=tensors[0].dtype)
    for t, o in zip(tensors, offs):
        flat[o:o + t.numel()].copy_(t.detach().reshape(-1))
        t.data = flat[o:o + t.numel()].view(t.shape)
    return flat


class ParamArena:
    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.dtype == torch.float32]
        self.flat = _flatten(self.params) if self.params else None
        self.ptrs = [p.data_ptr() for p in self.params]

    def intact(self):
        return all(p.data_ptr() == q for p, q in zip(self.params, self.ptrs))


def param_arena(module):
    """The (cached) arena of ``module``; rebuilt if some parameter storage was replaced since (e.g. ``module.to``)."""
    a = getattr(module, "_sgb_param_arena", None)
    if a is None or not a.intact():
        a = ParamArena(module)
        module._sgb_param_arena = a
    return a


class GradArena:
    """Gradients of every arena parameter as views of one flat buffer laid out like the parameter arena: zero_grad is
    one memset, the data-parallel gradient exchange one all-reduce, the optimiser one launch."""

    def __init__(self, arena):
        self.arena = arena
        self.flat = torch.zeros_like(arena.flat)
        self.views = []
        o = 0
        for p in arena.params:
            self.views.append(self.flat[o:o + p.numel()].view(p.shape))
            o += (p.numel() + 63) // 64 * 64

    def attach(self):
        for p, g in zip(self.arena.params, self.views):
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g
            p._sgb_direct_grad = True      # autograd_ops.ConvFn may accumulate into p.grad itself

    def zero(self):
        self.flat.zero_()
        self.attach()


class BufferArena:
    """Float buffers (BatchNorm running statistics) and integer buffers (num_batches_tracked) of a module, minus those that
    already live in another flat arena (spectral-norm u / v), flattened per dtype so that the EMA buffer update is one
    lerp + one copy."""

    def __init__(self, module, skip=lambda b: False):
        bufs = [b for b in module.buffers() if not skip(b)]
        self.floats = [b for b in bufs if b.dtype == torch.float32]
        self.ints = [b for b in bufs if not b.is_floating_point()]
        self.other = [b for b in bufs if b.is_floating_point() and b.dtype != torch.float32]
        self.fflat = _flatten(self.floats, align=4) if self.floats else None
        self.iflat = None
        if self.ints and all(b.dtype == self.ints[0].dtype for b in self.ints):
            self.ints = [b if b.dim() > 0 else b for b in self.ints]
            self.iflat = _flatten(self.ints, align=1)
        self.ptrs = [b.data_ptr() for b in self.floats + self.ints]

    def intact(self):
        return all(b.data_ptr() == q for b, q in zip(self.floats + self.ints, self.ptrs))

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

    def zero(self):
        self.flat.zero_()
        self.attach()

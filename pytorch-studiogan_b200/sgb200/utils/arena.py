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

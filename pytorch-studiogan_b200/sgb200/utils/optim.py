"""Adam over a flat parameter arena: the whole network's update is ONE ``sgb_adam_ema_step`` launch.

Arithmetic is torch.optim.Adam's (the reference's optimiser, src/config.py:541-563: eps 1e-6, no amsgrad, weight decay
0 on every BASELINE config): m <- b1 m + (1-b1) g, v <- b2 v + (1-b2) g^2, p <- p - lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t)
+ eps).  torch skips a parameter whose ``.grad`` is None: here the parameters that are frozen at step time
(``requires_grad`` False -- ``misc.toggle_grad(..., freezeD)``, src/utils/misc.py:192-216) are left out of the update,
moments included, by launching over the contiguous arena ranges of the trainable ones (one range on the hot path).

``state_dict`` / ``load_state_dict`` speak torch.optim.Adam's format, so reference checkpoints of the optimiser state
load and save unchanged."""
import torch

from .. import kernels as K
from .arena import GradArena, param_arena


class ArenaAdam(object):
    def __init__(self, module, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0):
        if weight_decay != 0.0:
            raise NotImplementedError("weight decay is 0 on every hot-path config")
        self.module = module
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False)
        self.param_groups = [dict(self.defaults, params=[p for p in module.parameters()])]
        self._step_host = 0
        self.step_t = None                       # int32 device scalar: the step count lives on the device (graph replays)
        self.arena = self.grads = self.exp_avg = self.exp_avg_sq = None
        self.grad_scale = 1.0
        self._pending = None

    @property
    def step_count(self):
        return int(self.step_t) if self.step_t is not None else self._step_host

    @step_count.setter
    def step_count(self, v):
        self._step_host = int(v)
        if self.step_t is not None:
            self.step_t.fill_(int(v))

    def _ensure(self):
        arena = param_arena(self.module)
        if arena is not self.arena:
            if self.arena is not None and self.exp_avg is not None and arena.flat.numel() != self.exp_avg.numel():
                raise RuntimeError("parameter set changed under the optimiser")
            self.arena = arena
            self.grads = GradArena(arena)
            if self.step_t is None or self.step_t.device != arena.flat.device:
                self.step_t = torch.full((1,), self._step_host, device=arena.flat.device, dtype=torch.int32)
            if self.exp_avg is None or self.exp_avg.device != arena.flat.device:
                old = (self.exp_avg, self.exp_avg_sq)
                self.exp_avg = torch.zeros_like(arena.flat)
                self.exp_avg_sq = torch.zeros_like(arena.flat)
                if old[0] is not None:
                    self.exp_avg.copy_(old[0])
                    self.exp_avg_sq.copy_(old[1])
            if self._pending is not None:
                self._load_pending()
        if any(p.dtype != torch.float32 for p in self.module.parameters()):
            raise RuntimeError("ArenaAdam needs fp32 master parameters")

    def zero_grad(self, set_to_none=False):
        self._ensure()
        self.grads.zero()
        self.grad_scale = 1.0

    def all_reduce(self, group, world_size):
        import torch.distributed as dist
        self._ensure()
        dist.all_reduce(self.grads.flat, group=group)
        self.grad_scale = 1.0 / world_size

    @torch.no_grad()
    def step(self):
        self._ensure()
        self.grads.attach()
        self.step_t.add_(1)
        g = self.param_groups[0]
        for lo, hi in self._trainable_ranges():
            K.adam_ema_step(self.arena.flat[lo:hi], self.grads.flat[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], g["lr"],
                            g["betas"][0], g["betas"][1], g["eps"], 0, grad_scale=self.grad_scale, step_dev=self.step_t)
        self.grad_scale = 1.0

    def _trainable_ranges(self):
        """Contiguous [lo, hi) element ranges of the arena that hold parameters with requires_grad set."""
        key = tuple(p.requires_grad for p in self.arena.params)
        if getattr(self, "_ranges_key", None) != key:
            ranges, cur = [], None
            for o, p in self._offsets():
                end = o + (p.numel() + 63) // 64 * 64
                if p.requires_grad:
                    cur = [o, end] if cur is None else [cur[0], end]
                elif cur is not None:
                    ranges.append(tuple(cur))
                    cur = None
            if cur is not None:
                ranges.append(tuple(cur))
            self._ranges_key, self._ranges = key, ranges
        return self._ranges

    # ---------------------------------------------------------------------------------------- torch.optim.Adam format
    def _offsets(self):
        o = 0
        for p in self.arena.params:
            yield o, p
            o += (p.numel() + 63) // 64 * 64

    def state_dict(self):
        self._ensure()
        state = {}
        if self.step_count > 0:
            for i, (o, p) in enumerate(self._offsets()):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[o:o + p.numel()].view(p.shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.param_groups[0]["params"])))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        self._pending = sd
        for k, v in sd["param_groups"][0].items():
            if k != "params" and k in self.param_groups[0]:
                self.param_groups[0][k] = tuple(v) if k == "betas" else v
        self._ensure()
        if self._pending is not None:
            self._load_pending()

    def _load_pending(self):
        sd, self._pending = self._pending, None
        steps = set()
        for i, (o, p) in enumerate(self._offsets()):
            st = sd["state"].get(i)
            if st is None:
                continue
            self.exp_avg[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise NotImplementedError("per-parameter step counts differ; the arena optimiser keeps one")
        self.step_count = steps.pop() if steps else 0

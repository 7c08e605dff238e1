"""HBM probe: pure-write (fill), pure-read (sum), copy bandwidth of one B200 -- the ceiling of write-dominated layers."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 30                                  # 2 GiB of bf16 per buffer
a = torch.empty(n, device=dev, dtype=torch.bfloat16).normal_()
b = torch.empty(n, device=dev, dtype=torch.bfloat16)


def best(fn, it=10):
    ts = []
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)


out = {}
out["fill_write_GBs"] = n * 2 / best(lambda: b.zero_()) * 1e-9
out["memset_write_GBs"] = n * 2 / best(lambda: b.fill_(1.0)) * 1e-9
out["sum_read_GBs"] = n * 2 / best(lambda: a.view(torch.int16).sum()) * 1e-9
out["copy_rw_GBs"] = 2 * n * 2 / best(lambda: b.copy_(a)) * 1e-9
# 1 read : 2 written (the 1x1 64->128 layer's ratio): out[2n] = cat(a, a)
c = torch.empty(2 * n, device=dev, dtype=torch.bfloat16)
out["read1_write2_GBs"] = 3 * n * 2 / best(lambda: (c[:n].copy_(a), c[n:].copy_(a))) * 1e-9
print(json.dumps(out))

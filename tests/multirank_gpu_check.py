"""torchrun entry: `python -m torch.distributed.run --nproc-per-node N tests/multirank_gpu_check.py` -- N ranks x 4 images
against one rank x 4N images (sgb200.utils.ddp_check); rank 0 prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-studiogan_b200"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    from sgb200.utils.ddp_check import multirank_parity_check
    out = multirank_parity_check(device)
    if dist.get_rank() == 0:
        print("MULTIRANK_CHECK " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

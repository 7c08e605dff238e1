"""bench.py — BigGAN-Deep 256x256 G+D step throughput (BASELINE.json config 4) on N B200s, one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--config yaml]

step     = WORKER.train_discriminator (d_updates_per_step = 2 discriminator updates) + WORKER.train_generator
           (one generator update + EMA), i.e. one iteration of the reference loop (src/loader.py:392-398).
metric   = images/s = global batch x acml_steps / step time, whole job over all ranks (strong scaling: the global
           batch is fixed at 256, each rank takes 256/N).
value    = synthetic real images already resident in HBM.
e2e      = the same step driven from pinned HOST buffers: the basket's H2D copy and a D2H read of both losses are inside
           the timed region.
--impl reference : the CPU path (oracle port of the reference step) on the host cores, bounded sample, rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "pytorch-studiogan_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

# conv + linear + attention-bmm FLOPs per image (2 x MAC), BASELINE.md section 3 / SURVEY.md 8(d)
# WGAN-GP-128res (G 18.2 / D 9.4 GF per image forward): 5 D updates x (G fwd + 2 D fwd + 2 D bwd + penalty ~ 7 D fwd)
# + 1 G update (G fwd + D fwd + D dgrad + G bwd) ~ 775 GF per batch image per step
STEP_GFLOP_PER_IMAGE = {"BigGAN-Deep-256res": 1141.0, "WGAN-GP-128res": 775.0}
METRIC_NAME = {"BigGAN-Deep-256res": "BigGAN-Deep 256x256 G+D step images/sec",
               "WGAN-GP-128res": "WGAN-GP ResNetGAN 128x128 G+D step images/sec (5 D updates with gradient penalty + 1 G update)"}
G_FWD_GF, D_FWD_GF = 58.80, 60.50


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="override the global batch (diagnostics only)")
    ap.add_argument("--config", default=os.path.join(PKG, "configs", "BigGAN-Deep-256res.yaml"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--graphs", default="auto", choices=["auto", "on", "off"],
                    help="capture each training phase in a CUDA graph; auto = on when the per-GPU batch is <= 64 (launch-bound regime)")
    ap.add_argument("--no-fid", action="store_true", help="skip the FID-50k evaluation timing (N = 1 only)")
    ap.add_argument("--fid-num", type=int, default=50000)
    ap.add_argument("--cpu-batch", type=int, default=4, help="images per CPU-baseline step (4 -> ~15 s of CPU work on 32 threads)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="0 = min(32, host cores): more threads only add contention for these layer sizes")
    return ap.parse_args()


class Logger:
    def info(self, *a, **k):
        pass


class SyntheticBasketLoader:
    """Infinite loader with the reference basket contract (src/loader.py:178-193): one item = batch x acml x d_updates
    images in [-1, 1] fp32 NCHW + int64 labels.  ``device`` None -> pinned host memory (the e2e leg), else HBM-resident."""

    def __init__(self, per_rank_batch, n_items, img_size, num_classes, seed, device=None, pool=2):
        g = torch.Generator().manual_seed(seed)
        self.items = []
        for _ in range(pool):
            img = torch.rand(per_rank_batch * n_items, 3, img_size, img_size, generator=g) * 2 - 1
            lab = torch.randint(0, num_classes, (per_rank_batch * n_items,), generator=g)
            if device is None:
                img, lab = img.pin_memory(), lab.pin_memory()
            else:
                img, lab = img.to(device), lab.to(device)
            self.items.append((img, lab))
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        self.i += 1
        return self.items[self.i % len(self.items)]


class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) > 8:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_worker(args, rank, world, device):
    from sgb200 import config as C
    from sgb200.models import model as M
    from sgb200.utils import misc
    from sgb200.worker import WORKER
    cfgs = C.Configurations(args.config)
    if args.batch:
        cfgs.OPTIMIZATION.batch_size = args.batch
    global_batch = cfgs.OPTIMIZATION.batch_size
    assert global_batch % world == 0
    cfgs.OPTIMIZATION.batch_size = global_batch // world           # per-rank batch, as src/loader.py:162
    cfgs.RUN.cuda_graphs = args.graphs == "on" or (args.graphs == "auto" and cfgs.OPTIMIZATION.batch_size <= 64)
    cfgs.RUN.distributed_data_parallel = world > 1
    cfgs.RUN.synchronized_bn = world > 1
    misc.fix_seed(0 + rank)                                         # seed + rank (src/loader.py:99)
    Gen, _, _, Dis, Gen_ema, _, _, ema = M.load_generator_discriminator(cfgs.DATA, cfgs.OPTIMIZATION, cfgs.MODEL, cfgs.STYLEGAN,
                                                                        cfgs.MODULES, cfgs.RUN, device, Logger())
    if world > 1:
        Gen, _, _, Dis, Gen_ema, _, _ = M.prepare_parallel_training(Gen, None, None, Dis, Gen_ema, None, None, cfgs.MODEL, world,
                                                                    True, True, cfgs.MODEL.apply_g_ema, device)
    cfgs.define_optimizer(Gen, Dis)
    worker = WORKER(cfgs=cfgs, run_name="bench", Gen=Gen, Gen_mapping=None, Gen_synthesis=None, Dis=Dis, Gen_ema=Gen_ema,
                    Gen_ema_mapping=None, Gen_ema_synthesis=None, ema=ema, eval_model=None, train_dataloader=None,
                    eval_dataloader=None, global_rank=rank, local_rank=device, mu=None, sigma=None, real_feats=None, logger=Logger())
    return cfgs, worker, global_batch


def run_steps(worker, n, read_losses):
    out = None
    for s in range(n):
        _, d_loss = worker.train_discriminator(s)
        g_loss = worker.train_generator(s)
        if read_losses:
            out = (float(d_loss.detach()), float(g_loss.detach()))      # D2H read of the step's results
    return out


def timed(worker, steps, world, read_losses):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = run_steps(worker, steps, read_losses)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps, losses


def fid_eval_seconds(worker, cfgs, device, num_eval, batch):
    """BASELINE metric, second half: wall seconds of one FID-N evaluation (WORKER.evaluate: N generated images through
    G_ema -> quantise/resize/normalise -> InceptionV3 -> IS + FID), preceded by the reference-statistics pass over N
    synthetic uint8-valued reference images.  Inception weights are seeded (the pretrained file cannot be downloaded
    here), which changes no shape or FLOP."""
    from sgb200.metrics import features, fid
    from sgb200.metrics.preparation import LoadEvalModel
    ev = LoadEvalModel("InceptionV3_tf", "legacy", 1, False, device)
    gen = torch.Generator(device=device).manual_seed(1234)
    S = cfgs.DATA.img_size

    def ref_batches():
        for i in range(0, num_eval, batch):
            n = min(batch, num_eval - i)
            yield torch.randint(0, 256, (n, 3, S, S), generator=gen, device=device).float()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rf, _ = features.stack_real_features(ref_batches(), ev, False, device)
    mu, sigma = fid.calculate_moments(rf)
    torch.cuda.synchronize()
    t_ref = time.perf_counter() - t0
    worker.eval_model, worker.mu, worker.sigma, worker.num_eval = ev, mu, sigma, num_eval
    bs = cfgs.OPTIMIZATION.batch_size
    cfgs.OPTIMIZATION.batch_size = batch
    try:
        t0 = time.perf_counter()
        worker.evaluate(step=0, metrics=["is", "fid"], writing=False, training=True)
        torch.cuda.synchronize()
        t_eval = time.perf_counter() - t0
    finally:
        cfgs.OPTIMIZATION.batch_size = bs
    m = worker.last_metrics or {}
    return {"seconds": t_eval, "ref_stats_seconds": t_ref, "num_eval": num_eval, "batch": batch, "img_per_s": num_eval / t_eval,
            "FID": m.get("FID"), "IS": m.get("IS"), "inception_weights": "seeded (pretrained FID weights unavailable offline)"}


def cpu_step_images_per_sec(config_path, batch, threads, n_steps=1):
    """The reference step restated by the oracle (fp32 CPU torch): 2 discriminator updates + 1 generator update with Adam,
    BigGAN-Deep 256x256, on ``threads`` host threads.  A bounded sample: ``batch`` images per step."""
    import yaml
    from oracle import studiogan_oracle as O
    from sgb200 import config as C
    from sgb200.models import model as M
    torch.set_num_threads(threads)
    cfgs = C.Configurations(config_path)
    torch.manual_seed(0)
    Gen, _, _, Dis, _, _, _, _ = M.load_generator_discriminator(cfgs.DATA, cfgs.OPTIMIZATION, _no_ema(cfgs.MODEL), cfgs.STYLEGAN,
                                                                 cfgs.MODULES, cfgs.RUN, "cpu", Logger())
    m = cfgs.MODEL
    sdG = {k: v.detach().clone() for k, v in Gen.state_dict().items()}
    sdD = {k: v.detach().clone() for k, v in Dis.state_dict().items()}
    pG = [k for k, _ in Gen.named_parameters()]
    pD = [k for k, _ in Dis.named_parameters()]
    del Gen, Dis
    kw_g = dict(img_size=cfgs.DATA.img_size, g_conv_dim=m.g_conv_dim, g_depth=m.g_depth, attn_g_loc=tuple(m.attn_g_loc), apply_attn=m.apply_attn)
    kw_d = dict(img_size=cfgs.DATA.img_size, d_conv_dim=m.d_conv_dim, d_depth=m.d_depth, attn_d_loc=tuple(m.attn_d_loc), apply_attn=m.apply_attn)
    optG = torch.optim.Adam([sdG[k].requires_grad_(True) for k in pG], lr=cfgs.OPTIMIZATION.g_lr, betas=(cfgs.OPTIMIZATION.beta1, cfgs.OPTIMIZATION.beta2), eps=1e-6)
    optD = torch.optim.Adam([sdD[k].requires_grad_(True) for k in pD], lr=cfgs.OPTIMIZATION.d_lr, betas=(cfgs.OPTIMIZATION.beta1, cfgs.OPTIMIZATION.beta2), eps=1e-6)
    S, nc = cfgs.DATA.img_size, cfgs.DATA.num_classes
    t0 = time.perf_counter()
    for _ in range(n_steps):
        for _ in range(cfgs.OPTIMIZATION.d_updates_per_step):
            optD.zero_grad()
            real, yr = torch.rand(batch, 3, S, S) * 2 - 1, torch.randint(0, nc, (batch,))
            yf = torch.randint(0, nc, (batch,))
            z = torch.randn(batch, m.z_dim)
            with torch.no_grad():
                fake = O.deep_generator(sdG, z, yf, track=False, **kw_g)
            a, _ = O.deep_discriminator(sdD, real, yr, **kw_d)
            b, _ = O.deep_discriminator(sdD, fake, yf, **kw_d)
            O.d_hinge(a, b).backward()
            optD.step()
        optG.zero_grad()
        yf = torch.randint(0, nc, (batch,))
        z = torch.randn(batch, m.z_dim)
        fake = O.deep_generator(sdG, z, yf, **kw_g)
        a, _ = O.deep_discriminator({k: v.detach() for k, v in sdD.items()}, fake, yf, **kw_d)
        O.g_hinge(a).backward()
        optG.step()
    dt = time.perf_counter() - t0
    return batch * n_steps / dt, dt


def _no_ema(MODEL):
    import copy
    m = copy.copy(MODEL)
    m.apply_g_ema = False
    return m


def shutdown(worker, world):
    """Leave without hanging: captured CUDA graphs that contain NCCL kernels must die before the process group does, and a
    watchdog hard-exits if the teardown still blocks (the JSON line is already flushed by then)."""
    sys.stdout.flush()
    if world <= 1:
        return

    def _hard_exit():
        time.sleep(20.0)
        os._exit(0)
    threading.Thread(target=_hard_exit, daemon=True).start()
    for name in ("_d_graph", "_g_graph"):
        if getattr(worker, name, None) is not None:
            setattr(worker, name, None)
    import gc
    gc.collect()
    torch.cuda.synchronize()
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass
    os._exit(0)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = os.path.splitext(os.path.basename(args.config))[0]

    if args.impl == "reference":
        if rank != 0:
            return
        threads = args.cpu_threads or min(32, os.cpu_count() or 1)
        # up to K timed measurements of one full CPU step each (the model is rebuilt per measurement, only the step itself
        # is inside the timed region); the loop stops early so that the whole run stays within ~4 minutes.
        vals, t_used = [], 0.0
        for i in range(max(1, args.steps)):
            v, dt = cpu_step_images_per_sec(args.config, args.cpu_batch, threads)
            vals.append(v)
            t_used += dt
            if t_used + dt > 240.0:
                break
        v = float(np.mean(vals))
        line = {"impl": "reference", "metric": "BigGAN-Deep 256x256 G+D step images/sec", "value": v, "unit": "img/s",
                "n_gpus": args.gpus, "steps": len(vals), "warmup": 0, "ms_per_step": 1000.0 * args.cpu_batch / v,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "global_batch": args.cpu_batch, "img_size": 256, "d_updates_per_step": 2,
                           "note": "CPU restatement (oracle port) of the reference step; /root/reference is not present on the GPU box"},
                "cpu_baseline": {"value": v, "unit": "img/s", "cores": threads, "kind": "port",
                                 "sample": "%d-image batch, one full step (2 D updates + 1 G update, fwd+bwd+Adam) per measurement" % args.cpu_batch},
                "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from sgb200 import _lib
    _lib.check(_lib.load().sgb_device_check(), "sgb_device_check")
    cfgs, worker, global_batch = build_worker(args, rank, world, device)
    opt = cfgs.OPTIMIZATION
    per_rank = opt.batch_size
    n_items = opt.acml_steps * opt.d_updates_per_step
    S = cfgs.DATA.img_size

    # ---- device-resident leg (value)
    dev_loader = SyntheticBasketLoader(per_rank, n_items, S, cfgs.DATA.num_classes, 100 + rank, device=device)
    worker.train_dataloader, worker.train_iter = dev_loader, iter(dev_loader)
    run_steps(worker, args.warmup, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.LAUNCHES[0]
    ms_step, _ = timed(worker, args.steps, world, False)
    launches = (_lib.LAUNCHES[0] - launches0)
    sampler.stop_flag = True
    value = global_batch * opt.acml_steps / (ms_step * 1e-3)

    # ---- per-kernel accounting on one extra (untimed) step: CUDA events around every library call
    prof, prof_top = None, None
    try:
        _lib.PROFILE["events"] = []
        _lib.PROFILE["enabled"] = True
        graphs_on, cfgs.RUN.cuda_graphs = cfgs.RUN.cuda_graphs, False      # the accounting step runs eagerly (events per call)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        run_steps(worker, 1, False)
        torch.cuda.synchronize()
        prof_wall_ms = (time.perf_counter() - w0) * 1e3
        _lib.PROFILE["enabled"] = False
        cfgs.RUN.cuda_graphs = graphs_on
        agg, by_tag = {}, {}
        for tag, flops, e0, e1, nbytes in _lib.PROFILE["events"]:
            ms = e0.elapsed_time(e1)
            kind = tag.split(" ")[0]
            a = agg.setdefault(kind, [0.0, 0.0, 0, 0.0])
            a[0] += ms; a[1] += flops; a[2] += 1; a[3] += nbytes
            t = by_tag.setdefault(tag, [0.0, 0.0, 0, 0.0])
            t[0] += ms; t[1] += flops; t[2] += 1; t[3] += nbytes
        _lib.PROFILE["events"] = []
        prof = {k: {"ms": v[0], "tflops": (v[1] / (v[0] * 1e-3) * 1e-12) if v[0] > 0 and v[1] > 0 else None, "launches": v[2], "flop": v[1],
                    "gbytes": v[3] * 1e-9, "gb_per_s": (v[3] / (v[0] * 1e-3) * 1e-9) if v[0] > 0 and v[3] > 0 else None}
                for k, v in agg.items()}
        prof_top = [{"tag": k, "ms": round(v[0], 3), "n": v[2], "tflops": round(v[1] / (v[0] * 1e-3) * 1e-12, 1) if v[1] > 0 else None}
                    for k, v in sorted(by_tag.items(), key=lambda kv: -kv[1][0])[:40]]
        if rank == 0:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "kernel_breakdown_n%d_b%d.json" % (world, global_batch)), "w") as fh:
                json.dump({"wall_ms_of_profiled_step": prof_wall_ms, "sum_ms": sum(v[0] for v in agg.values()),
                           "by_kind": prof, "by_tag": [{"tag": k, "ms": v[0], "n": v[2], "flop": v[1], "bytes": v[3]} for k, v in
                                                       sorted(by_tag.items(), key=lambda kv: -kv[1][0])]}, fh, indent=1)
    except Exception as ex:  # accounting must never take the bench line down
        prof = {"error": repr(ex)}

    # ---- end-to-end leg: pinned host baskets, H2D inside the timed region, loss read back every step
    e2e = None
    if not args.no_e2e:
        host_loader = SyntheticBasketLoader(per_rank, n_items, S, cfgs.DATA.num_classes, 200 + rank, device=None)
        worker.train_dataloader, worker.train_iter = host_loader, iter(host_loader)
        run_steps(worker, 1, True)
        ms_e2e, _ = timed(worker, args.steps, world, True)
        h2d = world * per_rank * n_items * (3 * S * S * 4 + 8)
        e2e = {"value": global_batch * opt.acml_steps / (ms_e2e * 1e-3), "unit": "img/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e2e}

    fid50k = None
    if not args.no_fid and world == 1:
        try:
            fid50k = fid_eval_seconds(worker, cfgs, device, args.fid_num, min(256, per_rank))
        except Exception as ex:  # the evaluation timing must never take the bench line down
            fid50k = {"error": repr(ex)}

    if rank != 0:
        shutdown(worker, world)
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if "bf16_tflops_sustained" in peaks else "fallback 1.4 PFLOP/s sustained"
    step_flop = STEP_GFLOP_PER_IMAGE.get(workload, 0.0) * 1e9 * global_batch
    # dram__bytes_read.sum + dram__bytes_write.sum of one captured launch (ncu --set full, profiles/r01_ncu_full_summary.txt):
    # conv3x3_rows_kernel, 3x3 64->64 @256x256, B=32 -- 268.7 MB read + 221.4 MB written for 536.9 MB of algorithmic bytes
    # (the tail of the output was still in L2 when the kernel ended): no re-reads.
    roof = {"bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf, "peak_source": peak_src, "traffic": 490.1e6,
            "traffic_note": "bytes per launch of the captured conv3x3_rows_kernel launch (B=32 3x3 64->64 256^2; algorithmic 536.9e6)",
            "step_algorithmic_tflop": step_flop * 1e-12,
            "step_achieved": step_flop / (ms_step * 1e-3) * 1e-12 / world,
            "step_frac": step_flop / (ms_step * 1e-3) * 1e-12 / world / peak_tf}
    if isinstance(prof, dict) and "conv_fprop" in prof:
        k = prof["conv_fprop"]
        prof = {kk: vv for kk, vv in prof.items() if kk in ("conv_fprop", "conv_wgrad")}
        roof.update({"kernel": "conv_fprop_kernel (fprop + dgrad, tcgen05)", "achieved": k["tflops"], "frac": k["tflops"] / peak_tf,
                     "kernel_ms_per_step": k["ms"], "kernel_share_of_step": k["ms"] / ms_step, "kernels": prof})
    else:
        roof.update({"achieved": roof["step_achieved"], "frac": roof["step_frac"], "kernels": prof})

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        threads = args.cpu_threads or min(32, os.cpu_count() or 1)
        try:
            v, dt = cpu_step_images_per_sec(args.config, args.cpu_batch, threads)
            cpu = {"value": v, "unit": "img/s", "cores": threads, "kind": "port",
                   "sample": "%d-image batch, one full step (2 D updates + 1 G update, fwd+bwd+Adam), %.1f s" % (args.cpu_batch, dt)}
        except Exception as ex:
            cpu = {"value": None, "unit": "img/s", "cores": threads, "kind": "port", "sample": "failed: %r" % (ex,)}

    line = {"metric": METRIC_NAME.get(workload, workload + " G+D step images/sec"), "value": value, "unit": "img/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "global_batch": global_batch, "per_gpu_batch": per_rank, "img_size": S,
                       "d_updates_per_step": opt.d_updates_per_step, "acml_steps": opt.acml_steps, "parallelism": "dp%d" % world, "cuda_graphs": bool(cfgs.RUN.cuda_graphs),
                       "l2": "per-step working set (tens of GB of activations) >> 126 MB L2; no explicit flush needed"},
            "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "fid50k": fid50k,
            "clocks": sampler.summary()}
    print(json.dumps(line))
    shutdown(worker, world)


if __name__ == "__main__":
    main()

"""Checkpoint files in the reference's format (src/utils/ckpt.py:28-75, src/utils/misc.py ``save_model``,
src/worker.py:940-985): one ``model=<G|D|G_ema>-<current|best>-weights-step=<N>.pth`` per network holding
``state_dict`` (+ ``optimizer`` and, for D, the run bookkeeping).  Module ``state_dict`` keys / shapes and the arena
optimiser's ``state_dict`` are the reference's, so files written by either side load in the other."""
import glob
import os
from os.path import join

import torch


def make_ckpt_dir(ckpt_dir):
    os.makedirs(ckpt_dir, exist_ok=True)
    return ckpt_dir


def save_model(model, when, step, ckpt_dir, states):
    """Write ``states`` and drop the previous file of the same (model, when) -- src/utils/misc.py save_model."""
    make_ckpt_dir(ckpt_dir)
    pattern = join(ckpt_dir, "model={model}-{when}-weights-step=".format(model=model, when=when))
    for old in glob.glob(glob.escape(pattern) + "*.pth"):
        os.remove(old)
    path = pattern + "{step}.pth".format(step=step)
    torch.save(states, path)
    return path


def _cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu()
    if isinstance(obj, dict):
        return {k: _cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_cpu(v) for v in obj)
    return obj


def save(worker, step, is_best, epoch=0, topk="initialize", aa_p=None, lecam_emas=None):
    """WORKER.save (src/worker.py:940-985)."""
    when = "best" if is_best is True else "current"
    opt, run = worker.OPTIMIZATION, worker.RUN
    g_states = {"state_dict": _cpu(worker.Gen.state_dict()), "optimizer": _cpu(opt.g_optimizer.state_dict())}
    d_states = {"state_dict": _cpu(worker.Dis.state_dict()), "optimizer": _cpu(opt.d_optimizer.state_dict()),
                "seed": getattr(run, "seed", 0), "run_name": worker.run_name, "step": step, "epoch": epoch, "topk": topk,
                "aa_p": aa_p, "best_step": worker.best_step, "best_fid": worker.best_fid,
                "best_fid_ckpt": getattr(run, "ckpt_dir", None), "lecam_emas": lecam_emas}
    ckpt_dir = run.ckpt_dir
    paths = [save_model("G", when, step, ckpt_dir, g_states), save_model("D", when, step, ckpt_dir, d_states)]
    if worker.Gen_ema is not None:
        e_states = {"state_dict": _cpu(worker.Gen_ema.state_dict())}
        paths.append(save_model("G_ema", when, step, ckpt_dir, e_states))
    if when == "best":
        save_model("G", "current", step, ckpt_dir, g_states)
        save_model("D", "current", step, ckpt_dir, d_states)
        if worker.Gen_ema is not None:
            save_model("G_ema", "current", step, ckpt_dir, e_states)
    return paths


def load_ckpt(model, optimizer, ckpt_path, load_model=False, load_opt=False, load_misc=False):
    """src/utils/ckpt.py:28-75 (strict state-dict load; optimiser state in torch.optim.Adam's format)."""
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    if load_model:
        model.load_state_dict(ckpt["state_dict"], strict=True)
    if load_opt:
        optimizer.load_state_dict(ckpt["optimizer"])
    if load_misc:
        return (ckpt["seed"], ckpt["run_name"], ckpt["step"], ckpt.get("epoch", 0), ckpt.get("topk", "initialize"),
                ckpt.get("aa_p", ckpt.get("ada_p")), ckpt["best_step"], ckpt["best_fid"],
                ckpt.get("best_fid_checkpoint_path", ckpt.get("best_fid_ckpt")), ckpt.get("lecam_emas"))
    return None


def load_StudioGAN_ckpts(ckpt_dir, load_best, Gen, Dis, g_optimizer, d_optimizer, apply_g_ema, Gen_ema, ema, is_train=True):
    """src/utils/ckpt.py:78-141 without the logging / seed side effects: returns the D file's bookkeeping tuple."""
    when = "best" if load_best is True else "current"

    def find(model):
        return glob.glob(glob.escape(join(ckpt_dir, "model={m}-{w}-weights-step=".format(m=model, w=when))) + "*.pth")[0]
    load_ckpt(Gen, g_optimizer, find("G"), load_model=True, load_opt=is_train)
    misc_ = load_ckpt(Dis, d_optimizer, find("D"), load_model=True, load_opt=is_train, load_misc=True)
    if apply_g_ema:
        load_ckpt(Gen_ema, None, find("G_ema"), load_model=True)
        ema.source, ema.target = Gen, Gen_ema
    return misc_

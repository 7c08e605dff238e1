"""Standalone dset1-vs-dset2 evaluator (reference ``src/evaluate.py:112-288``): IS of each folder, FID and improved
precision / recall / density / coverage between two image folders (or between dset2 and pre-computed ``dset1_moments`` /
``dset1_feats`` files), single process / single GPU.

    python -m sgb200.evaluate --dset1 DIR --dset2 DIR [--dset1_moments npz] [--dset1_feats npz]
                              [--eval_metrics is fid prdc] [--post_resizer legacy|friendly] [--batch_size 256]

Same pipeline as the reference with ``quantize=False`` (``src/metrics/preparation.py:106-107``): images are decoded to uint8,
shipped as uint8, resized to 299x299 and normalised by the fused device kernel, passed through InceptionV3 on the conv
engine; FID statistics are accumulated on the fly in fp64 (``MomentsAccumulator``), PRDC runs on the fp64 tile kernels.
A folder is either ``DIR/<class>/<image>`` (ImageFolder, as the reference expects) or a flat directory of images.
"""
import argparse
import json
import os

import numpy as np
import torch

from .metrics import fid, ins, prdc
from .metrics.preparation import LoadEvalModel

IMG_EXT = (".png", ".jpg", ".jpeg", ".bmp", ".webp")


def list_images(root):
    """Image paths of ``root`` in the order torchvision's ImageFolder would enumerate them (sorted classes, sorted files); a
    directory without sub-directories is read as one class."""
    subdirs = sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d)))
    paths = []
    for d in (subdirs or [""]):
        base = os.path.join(root, d)
        paths += [os.path.join(base, f) for f in sorted(os.listdir(base)) if f.lower().endswith(IMG_EXT)]
    return paths


def folder_batches(paths, batch_size):
    """uint8-valued NCHW float batches (what ``get_outputs(x, quantize=False)`` consumes), decoded with PIL; all images of a
    folder must share one size, as in the reference (its DataLoader stacks them)."""
    from PIL import Image
    for s in range(0, len(paths), batch_size):
        arr = np.stack([np.asarray(Image.open(p).convert("RGB"), dtype=np.uint8) for p in paths[s:s + batch_size]], 0)
        yield torch.from_numpy(arr).permute(0, 3, 1, 2).contiguous()


def extract(paths, eval_model, batch_size, device, want_feats):
    """(MomentsAccumulator, probs [N, 1008], feats [N, 2048] | None) of a folder."""
    acc = fid.MomentsAccumulator(2048, device)
    probs, feats = [], []
    for batch in folder_batches(paths, batch_size):
        with torch.no_grad():
            f, logits = eval_model.get_outputs(batch.to(device, non_blocking=True).float(), quantize=False)
        acc.update(f)
        probs.append(torch.nn.functional.softmax(logits, dim=1))
        if want_feats:
            feats.append(f)
    return acc, torch.cat(probs, 0), (torch.cat(feats, 0) if want_feats else None)


def evaluate(dset1=None, dset2=None, dset1_moments=None, dset1_feats=None, eval_metrics=("is", "fid", "prdc"),
             post_resizer="legacy", batch_size=256, device="cuda:0", state_dict=None, nearest_k=5):
    assert dset2 is not None, "dset2 (the folder under evaluation) is required"
    if "fid" in eval_metrics:
        assert dset1 is not None or dset1_moments is not None, "Either dset1 or dset1_moments should be given to compute FID."
    if "prdc" in eval_metrics:
        assert dset1 is not None or dset1_feats is not None, "Either dset1 or dset1_feats should be given to compute PRDC."
    device = torch.device(device)
    ev = LoadEvalModel("InceptionV3_tf", post_resizer, 1, False, device, state_dict=state_dict)
    want_feats = "prdc" in eval_metrics
    load1 = ("fid" in eval_metrics and dset1_moments is None) or ("prdc" in eval_metrics and dset1_feats is None)
    out = {}
    if load1:
        p1 = list_images(dset1)
        acc1, probs1, feats1 = extract(p1, ev, batch_size, device, want_feats)
        out["dset1_size"] = len(p1)
    p2 = list_images(dset2)
    acc2, probs2, feats2 = extract(p2, ev, batch_size, device, want_feats)
    out["dset2_size"] = len(p2)
    if "is" in eval_metrics:
        if load1:
            out["IS_dset1"] = float(ins.eval_features(probs1, None, None, len(p1), 1, False)[0])
        out["IS"] = float(ins.eval_features(probs2, None, None, len(p2), 1, False)[0])
    if "fid" in eval_metrics:
        if dset1_moments is None:
            mu1, sigma1 = acc1.finalize()
        else:
            z = np.load(dset1_moments)
            mu1, sigma1 = torch.as_tensor(z["mu"], device=device), torch.as_tensor(z["sigma"], device=device)
        mu2, sigma2 = acc2.finalize()
        out["FID"] = fid.frechet_distance_device(mu1, sigma1, mu2, sigma2)
    if "prdc" in eval_metrics:
        real = feats1.double() if dset1_feats is None else torch.as_tensor(np.load(dset1_feats)["real_feats"], dtype=torch.float64,
                                                                            device=device)
        pr = prdc.compute_prdc(real, feats2.double(), nearest_k)
        out.update({"Improved_Precision": pr["precision"], "Improved_Recall": pr["recall"], "Density": pr["density"],
                    "Coverage": pr["coverage"]})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dset1", type=str, default=None, help="reference image folder")
    ap.add_argument("--dset1_feats", type=str, default=None, help="npz with 'real_feats' (pre-computed features of dset1)")
    ap.add_argument("--dset1_moments", type=str, default=None, help="npz with 'mu' / 'sigma' (pre-computed moments of dset1)")
    ap.add_argument("--dset2", type=str, default=None, help="image folder under evaluation")
    ap.add_argument("--batch_size", type=int, default=256)
    ap.add_argument("--eval_backbone", type=str, default="InceptionV3_tf")
    ap.add_argument("--post_resizer", type=str, default="legacy")
    ap.add_argument("--eval_metrics", nargs="+", default=["is", "fid", "prdc"])
    args = ap.parse_args()
    if args.eval_backbone != "InceptionV3_tf":
        raise NotImplementedError("only the InceptionV3_tf backbone is on the sgb200 path")
    res = evaluate(args.dset1, args.dset2, args.dset1_moments, args.dset1_feats, tuple(args.eval_metrics), args.post_resizer,
                   args.batch_size)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

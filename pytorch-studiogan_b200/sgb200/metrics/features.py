"""Feature stacking for evaluation (reference ``src/metrics/features.py:17-65``): generate ceil(N/B) batches with the
generator, run the evaluation model on each, concatenate, all-gather across ranks."""
import math
import os
import sys
import time

import torch

from ..utils import losses, sample


_T0 = [None]


def _tick(what, device):
    """SGB_EVAL_TIMING=1: wall seconds of the evaluation phases on stderr (synchronising; diagnostics only)."""
    if os.environ.get("SGB_EVAL_TIMING", "0") == "0":
        return
    torch.cuda.synchronize()
    now = time.perf_counter()
    if _T0[0] is not None and what:
        sys.stderr.write("[sgb200 eval] %-52s %.3f s\n" % (what, now - _T0[0]))
    _T0[0] = now


def generate_images_and_stack_features(generator, discriminator, eval_model, num_generate, y_sampler, batch_size, z_prior,
                                       truncation_factor, z_dim, num_classes, LOSS, RUN, MODEL, is_stylegan=False,
                                       generator_mapping=None, generator_synthesis=None, quantize=True, world_size=1, DDP=False,
                                       device="cuda", logger=None, disable_tqdm=True, moments=None):
    """``moments`` (metrics.fid.MomentsAccumulator, optional): every batch's features are also folded into the running
    [sum f, sum f f^T]; only the rows that survive the reference's ``[:num_generate]`` truncation of the rank-major
    gathered matrix are counted, so the statistics are those of exactly the same feature set."""
    eval_model.eval()
    _tick(None, device)
    feature_holder, prob_holder, fake_label_holder = [], [], []
    if device == 0 and logger is not None:
        logger.info("generate images and stack features ({} images).".format(num_generate))
    num_batches = int(math.ceil(float(num_generate) / float(batch_size)))
    if DDP:
        num_batches = num_batches // world_size + 1
    rank = 0
    if DDP:
        import torch.distributed as dist
        rank = dist.get_rank()
    first_row, done = rank * num_batches * batch_size, 0          # this rank's rows in the gathered (rank-major) matrix
    for _ in range(num_batches):
        fake_images, fake_labels, _, _, _, _, _ = sample.generate_images(
            z_prior=z_prior, truncation_factor=truncation_factor, batch_size=batch_size, z_dim=z_dim, num_classes=num_classes,
            y_sampler=y_sampler, radius="N/A", generator=generator, discriminator=discriminator, is_train=False, LOSS=LOSS, RUN=RUN,
            MODEL=MODEL, device=device)
        with torch.no_grad():
            features, logits = eval_model.get_outputs(fake_images, quantize=quantize)
            probs = torch.nn.functional.softmax(logits, dim=1)
        if moments is not None:
            keep = max(0, min(features.shape[0], num_generate - (first_row + done)))
            moments.update(features[:keep])
            done += features.shape[0]
        feature_holder.append(features)
        prob_holder.append(probs)
        fake_label_holder.append(fake_labels)
    feature_holder = torch.cat(feature_holder, 0)
    prob_holder = torch.cat(prob_holder, 0)
    fake_label_holder = torch.cat(fake_label_holder, 0)
    _tick("generate + extract (%d batches of %d)" % (num_batches, batch_size), device)
    if DDP:
        feature_holder = torch.cat(losses.GatherLayer.apply(feature_holder), dim=0)
        prob_holder = torch.cat(losses.GatherLayer.apply(prob_holder), dim=0)
        fake_label_holder = torch.cat(losses.GatherLayer.apply(fake_label_holder), dim=0)
        _tick("gather features / probabilities / labels", device)
    return feature_holder, prob_holder, list(fake_label_holder.detach().cpu().numpy())


def stack_real_features(images_iter, eval_model, quantize, device):
    """Features of a stream of real image batches (reference ``sample_images_from_loader_and_stack_features`` :68-104)."""
    feats, probs = [], []
    for images in images_iter:
        with torch.no_grad():
            f, logits = eval_model.get_outputs(images.to(device), quantize=quantize)
        feats.append(f)
        probs.append(torch.nn.functional.softmax(logits, dim=1))
    return torch.cat(feats, 0), torch.cat(probs, 0)

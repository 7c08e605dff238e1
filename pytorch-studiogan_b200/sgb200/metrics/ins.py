"""Inception score (reference ``src/metrics/ins.py:28-79``)."""
import torch


def calculate_kl_div(ps, splits):
    scores = []
    num_samples = ps.shape[0]
    with torch.no_grad():
        for j in range(splits):
            part = ps[(j * num_samples // splits):((j + 1) * num_samples // splits), :]
            kl = part * (torch.log(part) - torch.log(torch.unsqueeze(torch.mean(part, 0), 0)))
            kl = torch.exp(torch.mean(torch.sum(kl, 1)))
            scores.append(kl.unsqueeze(0))
        scores = torch.cat(scores, 0)
        m_scores = torch.mean(scores).detach().cpu().numpy()
        m_std = torch.std(scores).detach().cpu().numpy()      # unbiased: NaN for splits = 1, as in the reference
    return m_scores, m_std


TF_LABEL_TABLE = "./src/utils/tf_imagenet_folder_label_pairs.txt"     # where the reference reads it (src/utils/misc.py:587)


def load_imagenet_label_dict(path=None):
    """folder (WNID) -> row index of the TF-Inception label table, as misc.load_ImageNet_label_dict
    (src/utils/misc.py:582-595): one folder per line, first blank-separated token.  The table is a data file of the
    StudioGAN checkout (it is not sorted by WNID, so it cannot be re-derived); returns None when it is not there."""
    import os
    path = path or os.environ.get("SGB_TF_LABEL_TABLE", TF_LABEL_TABLE)
    if not os.path.exists(path):
        return None
    table = {}
    with open(path, "r") as fh:
        for i, line in enumerate(l for l in fh if l.strip()):
            table[line.split(" ")[0].strip()] = i
    return table


def eval_features(probs, labels, data_loader, num_features, split, is_acc, is_torch_backbone=False, label_table=None):
    """src/metrics/ins.py:45-79.  Top-1 / Top-5 of the generated classes under the TF InceptionV3: the loader's class index
    (sorted-folder order) is first mapped to the TF label index through the folder table (:47-49,69-76), then compared
    with logits[1:1001] shifted by the background class.  Without the table (or without a folder dataset behind the
    loader) the accuracies are reported as "N/A" rather than computed against the wrong index space."""
    probs = probs[:num_features]
    m_scores, m_std = calculate_kl_div(probs, splits=split)
    top1 = top5 = "N/A"
    if is_acc and labels is not None and not is_torch_backbone:
        table = label_table if label_table is not None else load_imagenet_label_dict()
        class_to_idx = getattr(getattr(getattr(data_loader, "dataset", None), "data", None), "class_to_idx", None)
        if table is None or class_to_idx is None:
            import warnings
            warnings.warn("sgb200: Top-1/Top-5 need the TF ImageNet folder table and an ImageFolder dataset; reporting N/A")
            return m_scores, m_std, top1, top5
        lut = torch.full((max(class_to_idx.values()) + 1,), -1, dtype=torch.long)
        for folder, idx in class_to_idx.items():
            lut[idx] = table[folder]
        lab = lut.to(probs.device)[torch.as_tensor(labels[:num_features], device=probs.device).long()]
        cls = probs[:, 1:1001]                                   # the TF Inception logits carry a background class first
        top5_idx = torch.topk(cls, 5, dim=1).indices
        top1 = float((top5_idx[:, 0] == lab).float().mean())
        top5 = float((top5_idx == lab[:, None]).any(1).float().mean())
    elif is_acc and labels is not None:
        lab = torch.as_tensor(labels[:num_features], device=probs.device)
        top5_idx = torch.topk(probs, 5, dim=1).indices
        top1 = float((top5_idx[:, 0] == lab).float().mean())
        top5 = float((top5_idx == lab[:, None]).any(1).float().mean())
    return m_scores, m_std, top1, top5

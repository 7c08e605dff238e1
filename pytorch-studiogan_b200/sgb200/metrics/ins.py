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


def eval_features(probs, labels, data_loader, num_features, split, is_acc, is_torch_backbone=False):
    probs = probs[:num_features]
    m_scores, m_std = calculate_kl_div(probs, splits=split)
    top1 = top5 = "N/A"
    if is_acc and labels is not None:
        lab = torch.as_tensor(labels[:num_features], device=probs.device)
        cls = probs[:, 1:1001] if not is_torch_backbone else probs       # the TF Inception logits carry a background class first
        top5_idx = torch.topk(cls, 5, dim=1).indices
        top1 = float((top5_idx[:, 0] == lab).float().mean())
        top5 = float((top5_idx == lab[:, None]).any(1).float().mean())
    return m_scores, m_std, top1, top5

"""Frechet Inception Distance (reference ``src/metrics/fid.py:34-136``).

``frechet_inception_distance`` keeps the reference's numpy / scipy.linalg.sqrtm arithmetic (host, fp64).
``frechet_distance_device`` evaluates the same quantity on the GPU in fp64 through the symmetric form
tr sqrt(S1 S2) = sum_i sqrt(lambda_i(S1^{1/2} S2 S1^{1/2})) (two symmetric eigendecompositions instead of a 2048^2
Schur-based sqrtm: ~0.2 s instead of ~12 s of host time), and ``calculate_moments`` accumulates mean / covariance on the
device instead of copying the [N, 2048] feature matrix to the host."""
import numpy as np
import torch


def frechet_inception_distance(mu1, sigma1, mu2, sigma2, eps=1e-6):
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean)


class MomentsAccumulator(object):
    """Running [sum f, sum f f^T] in fp64 on the device (sgb_feat_moments_accumulate): features are folded in batch by batch
    as the Inception pipeline produces them, the [N, 2048] matrix is never needed for FID, and under DDP the two
    accumulators (33.5 MB) are all-reduced instead of gathering the features.  ``finalize`` returns the mean
    and np.cov(rowvar=False) (src/metrics/fid.py:65-98)."""

    def __init__(self, dim, device):
        self.dim, self.n = dim, 0
        self.sum = torch.zeros(dim, dtype=torch.float64, device=device)
        self.outer = torch.zeros((dim, dim), dtype=torch.float64, device=device)

    def update(self, feats):
        from .. import _lib as L
        f = feats.detach().to(torch.float32).contiguous()
        assert f.dim() == 2 and f.shape[1] == self.dim
        if f.shape[0] == 0:
            return
        L.call("sgb_feat_moments_accumulate", L.ptr(f), f.shape[0], self.dim, L.ptr(self.sum), L.ptr(self.outer), L.stream_ptr())
        self.n += f.shape[0]

    def all_reduce(self, group=None):
        import torch.distributed as dist
        n = torch.tensor([float(self.n)], dtype=torch.float64, device=self.sum.device)
        for t in (self.sum, self.outer, n):
            dist.all_reduce(t, group=group)
        self.n = int(n.item())

    def finalize(self):
        from .. import _lib as L
        mu = torch.empty_like(self.sum)
        sigma = torch.empty_like(self.outer)
        L.call("sgb_feat_moments_finalize", L.ptr(self.sum), L.ptr(self.outer), float(self.n), self.dim, L.ptr(mu), L.ptr(sigma),
               L.stream_ptr())
        return mu, sigma


def calculate_moments(feats):
    """mean and unbiased covariance (np.cov(rowvar=False)) of an [N, D] feature tensor, fp64 on its device: the moment
    kernels on CUDA tensors (in row blocks of 4096), plain tensor algebra on the host (CPU tests / oracle comparisons)."""
    if feats.is_cuda:
        acc = MomentsAccumulator(feats.shape[1], feats.device)
        for s in range(0, feats.shape[0], 4096):
            acc.update(feats[s:s + 4096])
        return acc.finalize()
    f = feats.to(torch.float64)
    n = f.shape[0]
    mu = f.mean(0)
    fc = f - mu
    sigma = fc.t() @ fc / (n - 1)
    return mu, sigma


def _sym_sqrt(m):
    w, v = torch.linalg.eigh(m)
    return (v * torch.sqrt(torch.clamp(w, min=0))[None, :]) @ v.t()


def frechet_distance_device(mu1, sigma1, mu2, sigma2):
    mu1, mu2 = torch.as_tensor(mu1, dtype=torch.float64), torch.as_tensor(mu2, dtype=torch.float64, device=torch.as_tensor(mu1).device)
    s1 = torch.as_tensor(sigma1, dtype=torch.float64, device=mu1.device)
    s2 = torch.as_tensor(sigma2, dtype=torch.float64, device=mu1.device)
    diff = mu1 - mu2
    a = _sym_sqrt(s1)
    m = a @ s2 @ a
    m = (m + m.t()) * 0.5
    tr_covmean = torch.sqrt(torch.clamp(torch.linalg.eigvalsh(m), min=0)).sum()
    return float(diff.dot(diff) + torch.trace(s1) + torch.trace(s2) - 2 * tr_covmean)


def calculate_fid(fake_feats, pre_cal_mean, pre_cal_std, num_generate):
    m1, s1 = calculate_moments(fake_feats[:num_generate])
    fid = frechet_distance_device(m1, s1, torch.as_tensor(pre_cal_mean, device=m1.device), torch.as_tensor(pre_cal_std, device=m1.device))
    return fid, m1.cpu().numpy(), s1.cpu().numpy()

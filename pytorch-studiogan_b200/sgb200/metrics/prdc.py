"""Improved precision / recall and density / coverage (reference ``src/metrics/prdc.py:129-168``) on the device.

The reference materialises three N x N fp64 distance matrices on the host (20 GB each at N = 50 000).  Here distances are
produced in tiles ( |a|^2 + |b|^2 - 2 a.b in fp64 ), reduced immediately to what the four metrics need (k-th smallest per
row, per-column count, per-row any / minimum), and never stored: on CUDA tensors by the hand-written tile kernels of
csrc/metrics.cu, on host tensors (CPU tests) by the same algebra in tensor ops."""
import torch


def _rows(n, tile):
    for s in range(0, n, tile):
        yield s, min(s + tile, n)


def _sqdist(a, b, a2, b2):
    d = a2[:, None] + b2[None, :] - 2.0 * (a @ b.t())
    return torch.clamp(d, min=0)


def kth_nn_distances(x, k, tile=4096):
    """Distance to the k-th nearest neighbour, self included as the 0-th (reference get_kth_value(k = nearest_k + 1))."""
    x2 = (x * x).sum(1)
    out = torch.empty(x.shape[0], dtype=x.dtype, device=x.device)
    for s, e in _rows(x.shape[0], tile):
        d = _sqdist(x[s:e], x, x2[s:e], x2)
        out[s:e] = torch.sqrt(torch.kthvalue(d, k + 1, dim=1).values)
    return out


def _compute_prdc_cuda(real, fake, nearest_k):
    """The tile kernels of csrc/metrics.cu: two radii passes (real-real, fake-fake) and one real-fake pass, every distance
    tile reduced in registers / shared memory; nothing of size N x N is ever written."""
    from .. import _lib as L
    real, fake = real.contiguous(), fake.contiguous()
    nr, D = real.shape
    nf = fake.shape[0]
    dev = real.device
    rn, fn = torch.empty(nr, dtype=torch.float64, device=dev), torch.empty(nf, dtype=torch.float64, device=dev)
    r_real, r_fake = torch.empty_like(rn), torch.empty_like(fn)
    L.call("sgb_prdc_radii", L.ptr(real), nr, D, nearest_k, L.ptr(rn), L.ptr(r_real), L.stream_ptr())
    L.call("sgb_prdc_radii", L.ptr(fake), nf, D, nearest_k, L.ptr(fn), L.ptr(r_fake), L.stream_ptr())
    col_count = torch.empty(nf, dtype=torch.int32, device=dev)
    row_any = torch.empty(nr, dtype=torch.uint8, device=dev)
    row_cov = torch.empty(nr, dtype=torch.uint8, device=dev)
    L.call("sgb_prdc_cross", L.ptr(real), L.ptr(rn), L.ptr(fake), L.ptr(fn), L.ptr(r_real), L.ptr(r_fake), nr, nf, D,
           L.ptr(col_count), L.ptr(row_any), L.ptr(row_cov), L.stream_ptr())
    return dict(precision=float((col_count > 0).double().mean()), recall=float(row_any.double().mean()),
                density=float(col_count.double().mean() / float(nearest_k)), coverage=float(row_cov.double().mean()))


def compute_prdc(real_features, fake_features, nearest_k, tile=4096):
    real = torch.as_tensor(real_features, dtype=torch.float64)
    fake = torch.as_tensor(fake_features, dtype=torch.float64, device=real.device)
    if real.is_cuda and nearest_k + 1 <= 8:
        return _compute_prdc_cuda(real, fake, nearest_k)
    r_real = kth_nn_distances(real, nearest_k, tile)
    r_fake = kth_nn_distances(fake, nearest_k, tile)
    real2, fake2 = (real * real).sum(1), (fake * fake).sum(1)
    nf = fake.shape[0]
    prec_any = torch.zeros(nf, dtype=torch.bool, device=real.device)
    dens_cnt = torch.zeros(nf, dtype=torch.float64, device=real.device)
    rec_any = torch.zeros(real.shape[0], dtype=torch.bool, device=real.device)
    cov = torch.zeros(real.shape[0], dtype=torch.bool, device=real.device)
    for s, e in _rows(real.shape[0], tile):
        d = torch.sqrt(_sqdist(real[s:e], fake, real2[s:e], fake2))          # [tile, nf] real-to-fake distances
        inside_real = d < r_real[s:e, None]
        prec_any |= inside_real.any(0)
        dens_cnt += inside_real.sum(0).to(torch.float64)
        rec_any[s:e] = (d < r_fake[None, :]).any(1)
        cov[s:e] = d.min(1).values < r_real[s:e]
    return dict(precision=float(prec_any.double().mean()), recall=float(rec_any.double().mean()),
                density=float((1.0 / float(nearest_k)) * dens_cnt.mean()), coverage=float(cov.double().mean()))

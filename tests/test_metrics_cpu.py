"""Evaluation arithmetic of the product (sgb200.metrics.*) on CPU tensors against the reference golden values
(the same functions run on the GPU in fp64; no kernel is involved in these formulas)."""
import os

import numpy as np
import torch

from sgb200.metrics import fid, ins, prdc


def test_fid_is_prdc_match_reference_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    m1, s1 = fid.calculate_moments(torch.from_numpy(g["feat_fake"]))
    m2, s2 = fid.calculate_moments(torch.from_numpy(g["feat_real"]))
    np.testing.assert_allclose(s1.numpy(), np.cov(g["feat_fake"].astype(np.float64), rowvar=False), rtol=1e-10, atol=1e-12)
    # symmetric-eigendecomposition form vs the reference's scipy.linalg.sqrtm value: well inside the +-0.5 % budget
    np.testing.assert_allclose(fid.frechet_distance_device(m1, s1, m2, s2), float(g["fid"]), rtol=1e-6)
    np.testing.assert_allclose(fid.frechet_inception_distance(m1.numpy(), s1.numpy(), m2.numpy(), s2.numpy()), float(g["fid"]), rtol=1e-7)   # golden used an fp32 mean for the fake set
    m, s = ins.calculate_kl_div(torch.from_numpy(g["probs"]), 1)
    np.testing.assert_allclose(float(m), g["is_1"][0], rtol=1e-6)
    assert np.isnan(float(s))
    m4, s4 = ins.calculate_kl_div(torch.from_numpy(g["probs"]), 4)
    np.testing.assert_allclose([float(m4), float(s4)], g["is_4"], rtol=1e-5)
    pr = prdc.compute_prdc(g["feat_real"][:200], g["feat_fake"][:180].astype(np.float64), 5, tile=64)
    np.testing.assert_allclose([pr["precision"], pr["recall"], pr["density"], pr["coverage"]], g["prdc"], rtol=1e-12)


def test_fid_identical_sets_is_zero():
    rng = np.random.RandomState(1)
    f = torch.from_numpy(rng.randn(500, 64))
    m, s = fid.calculate_moments(f)
    assert abs(fid.frechet_distance_device(m, s, m, s)) < 1e-8

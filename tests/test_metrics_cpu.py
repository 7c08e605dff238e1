"""Evaluation arithmetic of the product (sgb200.metrics.*) on CPU tensors against the reference golden values
(the same functions run on the GPU in fp64; no kernel is involved in these formulas)."""
import os

import numpy as np
import pytest
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


def test_inception_accuracy_maps_loader_labels_through_the_folder_table(tmp_path):
    """Top-1 / Top-5 (src/metrics/ins.py:45-76): loader class index (sorted folders) -> TF label row through
    tf_imagenet_folder_label_pairs.txt, compared with logits[1:1001].  A table that is NOT in sorted-folder order must
    change the answer; without the table the result is "N/A", never a number in the wrong index space."""
    import types
    from sgb200.metrics import ins
    folders = ["n03", "n01", "n02", "n00"]                       # TF row order (unsorted)
    table = tmp_path / "tf_pairs.txt"
    table.write_text("".join("%s %d name%d\n" % (f, i + 1, i) for i, f in enumerate(folders)))
    class_to_idx = {f: i for i, f in enumerate(sorted(folders))}   # ImageFolder order: n00, n01, n02, n03
    loader = types.SimpleNamespace(dataset=types.SimpleNamespace(data=types.SimpleNamespace(class_to_idx=class_to_idx)))
    n = 8
    labels = torch.tensor([0, 1, 2, 3, 0, 1, 2, 3])                # loader labels
    tf_rows = torch.tensor([folders.index(sorted(folders)[int(l)]) for l in labels])
    probs = torch.full((n, 1008), 1e-4)
    probs[torch.arange(n), 1 + tf_rows] = 0.9                     # the classifier is always right, in TF index space
    _, _, top1, top5 = ins.eval_features(probs, labels, loader, n, 1, True, label_table=ins.load_imagenet_label_dict(str(table)))
    assert top1 == 1.0 and top5 == 1.0
    # the un-mapped comparison would be wrong for this table (n00 is loader class 0 but TF row 3)
    assert float(((probs[:, 1:1001].argmax(1)) == labels).float().mean()) < 1.0
    with pytest.warns(UserWarning):
        out = ins.eval_features(probs, labels, loader, n, 1, True, label_table=None)
    assert out[2] == "N/A" and out[3] == "N/A"

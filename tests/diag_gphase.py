"""Diagnostic (not a test): per-parameter gradient errors of the G phase against the reference golden vectors."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "pytorch-studiogan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
from sgb200.utils import losses
dev = torch.device("cuda:0")
for tag, cd, depth, attn in [("deep32_c8", 8, 1, False), ("deep32_c16_attn_d2", 16, 2, True)]:
    g = np.load(os.path.join(ROOT, "tests", "golden", tag + ".npz"))
    G, D = T._build_from_golden(g, cd, depth, attn, dev)
    z, yf = torch.from_numpy(g["z"]).to(dev), torch.from_numpy(g["y_fake"]).to(dev)
    real, yr = torch.from_numpy(g["real"]).to(dev), torch.from_numpy(g["y_real"]).to(dev)
    for p in G.parameters(): p.requires_grad_(False)
    fake = G(z, yf); rd = D(real, yr); fd = D(fake.detach(), yf)
    losses.d_hinge(rd["adv_output"], fd["adv_output"]).backward()
    D.zero_grad(set_to_none=True)
    for p in G.parameters(): p.requires_grad_(True)
    for p in D.parameters(): p.requires_grad_(False)
    fake2 = G(z, yf)
    fake2.retain_grad()
    gl = losses.g_hinge(D(fake2, yf)["adv_output"]); gl.backward()
    print("==", tag, "g_loss", float(gl), float(g["g_loss"]), "fake2 l2", T.l2_err(fake2, torch.from_numpy(g["fake2"])))
    rows = []
    for n, p in G.named_parameters():
        ref = torch.from_numpy(g["Ggrad/" + n]).double(); got = p.grad.detach().double().cpu()
        cos = float((got * ref).sum() / (got.norm() * ref.norm() + 1e-30))
        rows.append((float((got - ref).norm() / (ref.norm() + 1e-30)), cos, float(ref.norm()), float(got.norm()), n))
    for r in sorted(rows, reverse=True)[:25]:
        print("  err %.3f cos %.4f |ref| %.3e |got| %.3e  %s" % r)
    print("  median err %.4f" % float(np.median([r[0] for r in rows])))

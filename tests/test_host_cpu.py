"""CPU-side checks of the host layer: C-ABI surface, ctypes signatures, module/state_dict compatibility and seeded
initialisation parity with the reference (golden G0/D0 tensors come from the real StudioGAN modules)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from sgb200 import _lib as L
from sgb200 import config as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|int64_t)\s+(sgb_\w+)\s*\(([^;{]*)\)\s*;", src):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        decls[m.group(2)] = n
    return decls


def test_header_and_ctypes_table_agree():
    decls = _declared()
    assert len(decls) >= 28
    assert set(decls) == set(L.SIGNATURES), set(decls) ^ set(L.SIGNATURES)
    for name, n in decls.items():
        assert len(L.SIGNATURES[name][1]) == n, name


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.sgb_abi_version() == 1          # pure host call, no GPU needed


def test_struct_layout_matches_header():
    # field order / count of the two descriptor structs (a mismatch would silently corrupt kernel arguments)
    src = open(os.path.join(ROOT, "include", "sgb200.h")).read()
    for sname, cls in (("sgb_conv_desc", L.ConvDesc), ("sgb_wgrad_desc", L.WgradDesc)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (sname, sname), src, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            for part in stmt.split(","):
                names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
        assert names == [f[0] for f in cls._fields_], (sname, names)


def _build(tag_cfg):
    from sgb200.models import big_resnet_deep_legacy as deep
    M = C.make_modules(True, True, "cBN", "big_resnet_deep_legacy")
    MODEL = C._Section(info_type="N/A", g_info_injection="N/A")
    torch.manual_seed(1234)
    G = deep.Generator(z_dim=16, g_shared_dim=16, img_size=32, g_conv_dim=tag_cfg["conv_dim"], apply_attn=tag_cfg["attn"],
                       attn_g_loc=[2], g_cond_mtd="cBN", num_classes=5, g_init="ortho", g_depth=tag_cfg["depth"],
                       mixed_precision=False, MODULES=M, MODEL=MODEL)
    D = deep.Discriminator(img_size=32, d_conv_dim=tag_cfg["conv_dim"], apply_d_sn=True, apply_attn=tag_cfg["attn"],
                           attn_d_loc=[1], d_cond_mtd="PD", aux_cls_type="W/O", d_embed_dim="N/A", normalize_d_embed=False,
                           num_classes=5, d_init="ortho", d_depth=tag_cfg["depth"], mixed_precision=False, MODULES=M, MODEL=MODEL)
    return G, D


@pytest.mark.parametrize("tag,cfg", [("deep32_c8", dict(conv_dim=8, depth=1, attn=False)),
                                     ("deep32_c16_attn_d2", dict(conv_dim=16, depth=2, attn=True))])
def test_state_dict_keys_and_seeded_init_match_reference(golden_dir, tag, cfg):
    g = np.load(os.path.join(golden_dir, tag + ".npz"))
    G, D = _build(cfg)
    for net, prefix in ((G, "G0/"), (D, "D0/")):
        sd = net.state_dict()
        ref_keys = [k[len(prefix):] for k in g.files if k.startswith(prefix)]
        assert list(sd.keys()) == ref_keys                     # same keys in the same registration order
        for k in ref_keys:
            ref = g[prefix + k]
            assert tuple(sd[k].shape) == tuple(ref.shape), k
            if k.endswith("sigma"):
                continue                                        # the golden script overwrote the attention gate
            np.testing.assert_array_equal(sd[k].numpy(), ref, err_msg=k)   # identical RNG consumption -> bit-identical init


def test_reference_checkpoint_loads_strictly(golden_dir):
    g = np.load(os.path.join(golden_dir, "deep32_c16_attn_d2.npz"))
    G, D = _build(dict(conv_dim=16, depth=2, attn=True))
    for net, prefix in ((G, "G1/"), (D, "D1/")):
        sd = {k[3:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(prefix[0] + "0/")}
        sd.update({k[len(prefix):]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(prefix)})
        net.load_state_dict(sd, strict=True)


def test_yaml_configs_drop_in():
    ref_cfgs = os.environ.get("SGB_REFERENCE_CONFIGS", "/root/reference/src/configs")
    if not os.path.isdir(ref_cfgs):
        pytest.skip("reference config tree not present on this box")
    for rel in ["ImageNet/BigGAN-Deep-256.yaml", "CIFAR10/BigGAN.yaml", "CIFAR10/SNGAN.yaml", "CIFAR10/WGAN-GP.yaml",
                "CIFAR10/BigGAN-Deep.yaml"]:
        cfg = C.Configurations(os.path.join(ref_cfgs, rel))
        assert callable(cfg.MODULES.g_conv2d) and callable(cfg.LOSS.d_loss)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        L.load()

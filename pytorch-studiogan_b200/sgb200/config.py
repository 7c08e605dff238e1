"""Configuration object with the reference's section names and YAML schema (src/config.py): DATA / MODEL / LOSS /
OPTIMIZATION / RUN / MODULES, a strict YAML overlay (unknown key -> AttributeError, reference :400-409) and the three
factories the hot path is built from: ``define_modules`` (:435-495), ``define_losses`` (:411-433) and
``define_optimizer`` (:497-565).  Only the keys that reach the BigGAN / ResNetGAN hot path are carried; sections that
belong to StyleGAN, augmentation or analysis tooling are accepted by the overlay but otherwise ignored.
"""
import types

import torch
import torch.nn as nn
import yaml

from .utils import losses, ops


class _Section(types.SimpleNamespace):
    pass


def _defaults():
    c = {}
    c["DATA"] = _Section(name="CIFAR10", img_size=32, num_classes=10, img_channels=3)
    c["MODEL"] = _Section(
        backbone="resnet", g_cond_mtd="W/O", d_cond_mtd="W/O", aux_cls_type="W/O", normalize_d_embed=False, d_embed_dim="N/A",
        apply_g_sn=False, apply_d_sn=False, g_act_fn="ReLU", d_act_fn="ReLU", apply_attn=False, attn_g_loc=["N/A"],
        attn_d_loc=["N/A"], z_prior="gaussian", z_dim=128, w_dim="N/A", g_shared_dim="N/A", g_conv_dim=64, d_conv_dim=64,
        g_depth="N/A", d_depth="N/A", apply_g_ema=False, g_ema_decay="N/A", g_ema_start="N/A", g_init="ortho", d_init="ortho",
        info_type="N/A", g_info_injection="N/A", info_num_discrete_c="N/A", info_num_conti_c="N/A", info_dim_discrete_c="N/A")
    c["LOSS"] = _Section(
        adv_loss="vanilla", cond_lambda="N/A", tac_gen_lambda="N/A", tac_dis_lambda="N/A", mh_lambda="N/A", apply_fm=False,
        fm_lambda="N/A", apply_r1_reg=False, r1_place="N/A", r1_lambda="N/A", m_p="N/A", temperature="N/A", apply_wc=False,
        wc_bound="N/A", apply_gp=False, gp_lambda="N/A", apply_dra=False, dra_lambda="N/A", apply_maxgp=False,
        maxgp_lambda="N/A", apply_cr=False, cr_lambda="N/A", apply_bcr=False, real_lambda="N/A", fake_lambda="N/A",
        apply_zcr=False, radius="N/A", g_lambda="N/A", d_lambda="N/A", apply_lo=False, lo_alpha="N/A", lo_beta="N/A",
        lo_rate="N/A", lo_lambda="N/A", lo_steps4train="N/A", lo_steps4eval="N/A", apply_topk=False, topk_gamma="N/A",
        topk_nu="N/A", infoGAN_loss_discrete_lambda="N/A", infoGAN_loss_conti_lambda="N/A", apply_lecam=False,
        lecam_lambda="N/A", lecam_ema_start_iter="N/A", lecam_ema_decay="N/A", apply_inv_reg=False, inv_reg_lambda="N/A")
    c["OPTIMIZATION"] = _Section(
        type_="Adam", batch_size=64, acml_steps=1, g_lr=0.0002, d_lr=0.0002, g_weight_decay=0.0, d_weight_decay=0.0,
        momentum="N/A", nesterov="N/A", alpha="N/A", beta1=0.5, beta2=0.999, d_first=True, g_updates_per_step=1,
        d_updates_per_step=5, total_steps=100000)
    c["PRE"] = _Section(apply_rflip=True)
    c["AUG"] = _Section(apply_diffaug=False, apply_ada=False, apply_apa=False, cr_aug_type="W/O", bcr_aug_type="W/O",
                        diffaug_type="W/O")
    c["STYLEGAN"] = _Section()
    c["RUN"] = _Section(mixed_precision=False, distributed_data_parallel=False, synchronized_bn=False, batch_statistics=False,
                        standing_statistics=False, standing_step=-1, standing_max_batch=-1, freezeD=-1, cuda_graphs=False, langevin_sampling=False,
                        truncation_factor=-1.0, eval_backbone="InceptionV3_tf", post_resizer="legacy", seed=-1)
    c["MISC"] = _Section(no_proc_data=["CIFAR10", "CIFAR100", "Tiny_ImageNet"])
    c["MODULES"] = _Section()
    return c


class Configurations(object):
    def __init__(self, cfg_file=None):
        self.cfg_file = cfg_file
        self.super_cfgs = _defaults()
        for k, v in self.super_cfgs.items():
            setattr(self, k, v)
        if cfg_file is not None:
            self._overwrite_cfgs(cfg_file)
        self.define_modules()
        self.define_losses()
        self.define_augments()

    def update_cfgs(self, cfgs, super="RUN"):
        for attr, value in cfgs.items():
            setattr(self.super_cfgs[super], attr, value)

    def _overwrite_cfgs(self, cfg_file):
        with open(cfg_file, "r") as f:
            yaml_cfg = yaml.load(f, Loader=yaml.FullLoader)
        for section, entries in yaml_cfg.items():
            tolerant = section in ("STYLEGAN", "AUG", "PRE")
            for attr, value in entries.items():
                if hasattr(self.super_cfgs[section], attr) or tolerant:
                    setattr(self.super_cfgs[section], attr, value)
                else:
                    raise AttributeError("There does not exist '{cls}.{attr}' attribute in the config.py.".format(
                        cls=section, attr=attr))

    def define_losses(self):
        g_losses = {"vanilla": losses.g_vanilla, "logistic": losses.g_logistic, "least_square": losses.g_ls,
                    "hinge": losses.g_hinge, "wasserstein": losses.g_wasserstein}
        d_losses = {"vanilla": losses.d_vanilla, "logistic": losses.d_logistic, "least_square": losses.d_ls,
                    "hinge": losses.d_hinge, "wasserstein": losses.d_wasserstein}
        self.LOSS.g_loss = g_losses[self.LOSS.adv_loss]
        self.LOSS.d_loss = d_losses[self.LOSS.adv_loss]

    def define_augments(self, device=None):
        """``AUG.series_augment`` (applied to every image the discriminator sees, src/worker.py:276-285,549) and
        ``AUG.parallel_augment`` (the CR / bCR copy, :326-355) as in src/config.py:567-629, for the augmentation types built on
        this path: DiffAugment ("diffaug") and the CR augmentation ("cr" / "bcr").  ADA / APA / SimCLR types raise."""
        from .utils import cr, diffaug, misc
        self.AUG.series_augment = misc.identity
        self.AUG.parallel_augment = misc.identity
        table = {"diffaug": diffaug.apply_diffaug, "cr": cr.apply_cr_aug, "bcr": cr.apply_cr_aug}
        if self.AUG.apply_ada or self.AUG.apply_apa:
            raise NotImplementedError("AUG.apply_ada / apply_apa are outside the sgb200 hot-path scope")
        if self.AUG.apply_diffaug:
            if self.AUG.diffaug_type not in ("cr", "diffaug"):
                raise NotImplementedError("AUG.diffaug_type '%s' (built: 'diffaug', 'cr')" % self.AUG.diffaug_type)
            self.AUG.series_augment = table[self.AUG.diffaug_type]
        if self.LOSS.apply_cr:
            if self.AUG.cr_aug_type not in ("cr", "diffaug"):
                raise NotImplementedError("AUG.cr_aug_type '%s' (built: 'cr', 'diffaug')" % self.AUG.cr_aug_type)
            self.AUG.parallel_augment = table[self.AUG.cr_aug_type]
        if self.LOSS.apply_bcr:
            if self.AUG.bcr_aug_type not in ("bcr", "diffaug"):
                raise NotImplementedError("AUG.bcr_aug_type '%s' (built: 'bcr', 'diffaug')" % self.AUG.bcr_aug_type)
            self.AUG.parallel_augment = table[self.AUG.bcr_aug_type]

    def define_modules(self):
        return make_modules(self.MODEL.apply_g_sn, self.MODEL.apply_d_sn, self.MODEL.g_cond_mtd, self.MODEL.backbone,
                            self.MODEL.g_act_fn, self.MODEL.d_act_fn, self.MODEL.g_info_injection, out=self.MODULES)

    def define_optimizer(self, Gen, Dis):
        """Adam with eps=1e-6 (src/config.py:541-563): on the device the one-launch arena optimiser (same arithmetic and
        state_dict format as torch.optim.Adam), on the CPU (host-logic tests) torch.optim.Adam itself."""
        opt = self.OPTIMIZATION
        if opt.type_ != "Adam":
            raise NotImplementedError("only Adam is on the BASELINE configs' hot path")
        betas_g = [opt.beta1, opt.beta2]
        if next(Gen.parameters()).is_cuda and opt.g_weight_decay == 0.0 and opt.d_weight_decay == 0.0:
            from .utils.optim import ArenaAdam
            self.OPTIMIZATION.g_optimizer = ArenaAdam(Gen, lr=opt.g_lr, betas=betas_g, eps=1e-6)
            self.OPTIMIZATION.d_optimizer = ArenaAdam(Dis, lr=opt.d_lr, betas=betas_g, eps=1e-6)
            return
        self.OPTIMIZATION.g_optimizer = torch.optim.Adam(params=[p for p in Gen.parameters()], lr=opt.g_lr, betas=betas_g,
                                                         weight_decay=opt.g_weight_decay, eps=1e-6)
        self.OPTIMIZATION.d_optimizer = torch.optim.Adam(params=[p for p in Dis.parameters()], lr=opt.d_lr, betas=betas_g,
                                                         weight_decay=opt.d_weight_decay, eps=1e-6)


def make_modules(apply_g_sn, apply_d_sn, g_cond_mtd="cBN", backbone="big_resnet", g_act_fn="ReLU", d_act_fn="ReLU",
                 g_info_injection="N/A", out=None):
    """The operator plug-in table (reference ``Configurations.define_modules``, src/config.py:435-495)."""
    M = out if out is not None else _Section()
    M.g_conv2d = ops.snconv2d if apply_g_sn else ops.conv2d
    M.g_deconv2d = ops.sndeconv2d if apply_g_sn else ops.deconv2d
    M.g_linear = ops.snlinear if apply_g_sn else ops.linear
    M.g_embedding = ops.sn_embedding if apply_g_sn else ops.embedding
    M.d_conv2d = ops.snconv2d if apply_d_sn else ops.conv2d
    M.d_deconv2d = ops.sndeconv2d if apply_d_sn else ops.deconv2d
    M.d_linear = ops.snlinear if apply_d_sn else ops.linear
    M.d_embedding = ops.sn_embedding if apply_d_sn else ops.embedding
    if g_cond_mtd == "cBN" or g_info_injection == "cBN" or backbone == "big_resnet":
        M.g_bn = ops.ConditionalBatchNorm2d
    elif g_cond_mtd == "W/O":
        M.g_bn = ops.batchnorm_2d
    else:
        raise NotImplementedError(g_cond_mtd)
    if not apply_d_sn:
        M.d_bn = ops.batchnorm_2d
    if g_act_fn != "ReLU" or d_act_fn != "ReLU":
        raise NotImplementedError("the fused kernels implement ReLU (every BASELINE config); got %s / %s" % (g_act_fn, d_act_fn))
    M.g_act_fn = nn.ReLU(inplace=True)
    M.d_act_fn = nn.ReLU(inplace=True)
    return M

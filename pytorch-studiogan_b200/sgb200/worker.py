"""Step runtime with the reference's API (``src/worker.py``): ``WORKER.train_discriminator(step)`` ->
(real_cond_loss, dis_acml_loss), ``WORKER.train_generator(step)`` -> gen_acml_loss, in the reference's order of
operations (D phase :213-497, G phase :502-681): toggle_grad, BN tracking toggles, ``d_updates_per_step`` x ``acml_steps``
loops, labels-then-z sampling, two separate discriminator passes for real and fake, loss / acml_steps, optimiser step,
EMA update after the generator step.  Mixed precision / GradScaler do not exist here: the kernels compute in bf16 with
fp32 accumulation and fp32 master weights, which needs no loss scaling.
"""
import torch

from .models import model as model_lib
from .utils import losses, misc, sample


class WORKER(object):
    def __init__(self, cfgs, run_name, Gen, Gen_mapping, Gen_synthesis, Dis, Gen_ema, Gen_ema_mapping, Gen_ema_synthesis, ema,
                 eval_model, train_dataloader, eval_dataloader, global_rank, local_rank, mu, sigma, real_feats, logger,
                 aa_p=None, best_step=0, best_fid=None, best_ckpt_path=None, lecam_emas=None, num_eval=None, loss_list_dict=None,
                 metric_dict_during_train=None):
        self.cfgs = cfgs
        self.run_name = run_name
        self.Gen, self.Dis, self.Gen_ema, self.ema = Gen, Dis, Gen_ema, ema
        self.eval_model = eval_model
        self.train_dataloader = train_dataloader
        self.eval_dataloader = eval_dataloader
        self.global_rank, self.local_rank = global_rank, local_rank
        self.mu, self.sigma, self.real_feats = mu, sigma, real_feats
        self.logger = logger
        self.best_step, self.best_fid, self.best_ckpt_path = best_step, best_fid, best_ckpt_path
        self.DATA, self.MODEL, self.LOSS, self.OPTIMIZATION, self.RUN = cfgs.DATA, cfgs.MODEL, cfgs.LOSS, cfgs.OPTIMIZATION, cfgs.RUN
        self.DDP = self.RUN.distributed_data_parallel
        self.train_iter = iter(train_dataloader) if train_dataloader is not None else None
        self.is_stylegan = False
        if self.LOSS.adv_loss == "MH" or self.MODEL.aux_cls_type != "W/O" or self.MODEL.d_cond_mtd in ("AC", "2C", "D2DCE"):
            raise NotImplementedError("conditioning losses other than PD / W/O are queued behind the hot path (SURVEY 8f-3)")
        for flag in ("apply_cr", "apply_bcr", "apply_zcr", "apply_lo", "apply_topk", "apply_lecam", "apply_r1_reg",
                     "apply_dra", "apply_maxgp", "apply_fm", "apply_wc"):
            if getattr(self.LOSS, flag, False):
                raise NotImplementedError("LOSS.%s is outside the sgb200 hot-path scope" % flag)

    # ------------------------------------------------------------------------------------------------ data
    def sample_data_basket(self):
        """One loader item carries batch_size * acml_steps * d_updates_per_step samples (src/worker.py:194-208)."""
        try:
            real_image_basket, real_label_basket = next(self.train_iter)
        except StopIteration:
            self.train_iter = iter(self.train_dataloader)
            real_image_basket, real_label_basket = next(self.train_iter)
        bs = self.OPTIMIZATION.batch_size
        return torch.split(real_image_basket, bs), torch.split(real_label_basket, bs)

    def _generate(self, is_train=True):
        return sample.generate_images(z_prior=self.MODEL.z_prior, truncation_factor=-1.0, batch_size=self.OPTIMIZATION.batch_size,
                                      z_dim=self.MODEL.z_dim, num_classes=self.DATA.num_classes, y_sampler="totally_random",
                                      radius="N/A", generator=self.Gen, discriminator=self.Dis, is_train=is_train, LOSS=self.LOSS,
                                      RUN=self.RUN, MODEL=self.MODEL, device=self.local_rank)

    # ------------------------------------------------------------------------------------------------ D phase
    def train_discriminator(self, current_step):
        misc.make_GAN_trainable(self.Gen, self.Gen_ema, self.Dis)
        misc.toggle_grad(self.Gen, False)
        misc.toggle_grad(self.Dis, True, self.RUN.freezeD)
        # the generator uses batch statistics here but must not move its running statistics (src/worker.py:225)
        self.Gen.apply(misc.untrack_bn_statistics)
        real_image_basket, real_label_basket = self.sample_data_basket()
        batch_counter = 0
        dis_acml_loss = None
        for _ in range(self.OPTIMIZATION.d_updates_per_step):
            self.OPTIMIZATION.d_optimizer.zero_grad()
            for _ in range(self.OPTIMIZATION.acml_steps):
                real_images = real_image_basket[batch_counter].to(self.local_rank, non_blocking=True)
                real_labels = real_label_basket[batch_counter].to(self.local_rank, non_blocking=True)
                fake_images, fake_labels, _, _, _, _, _ = self._generate(True)
                real_dict = self.Dis(real_images, real_labels)
                fake_dict = self.Dis(fake_images, fake_labels, adc_fake=False)
                dis_acml_loss = self.LOSS.d_loss(real_dict["adv_output"], fake_dict["adv_output"], DDP=self.DDP)
                if self.LOSS.apply_gp:
                    from .utils import gp
                    dis_acml_loss = dis_acml_loss + self.LOSS.gp_lambda * gp.cal_grad_penalty(
                        real_images=real_images, real_labels=real_labels, fake_images=fake_images, discriminator=self.Dis,
                        device=self.local_rank)
                dis_acml_loss = dis_acml_loss / self.OPTIMIZATION.acml_steps
                dis_acml_loss.backward()
                batch_counter += 1
            model_lib.allreduce_gradients(self.Dis)
            self.OPTIMIZATION.d_optimizer.step()
        return "N/A", dis_acml_loss

    # ------------------------------------------------------------------------------------------------ G phase
    def train_generator(self, current_step):
        misc.make_GAN_trainable(self.Gen, self.Gen_ema, self.Dis)
        misc.toggle_grad(self.Dis, False)
        misc.toggle_grad(self.Gen, True)
        self.Gen.apply(misc.track_bn_statistics)
        gen_acml_loss = None
        for _ in range(self.OPTIMIZATION.g_updates_per_step):
            self.OPTIMIZATION.g_optimizer.zero_grad()
            for _ in range(self.OPTIMIZATION.acml_steps):
                fake_images, fake_labels, _, _, _, _, _ = self._generate(True)
                fake_dict = self.Dis(fake_images, fake_labels)
                gen_acml_loss = self.LOSS.g_loss(fake_dict["adv_output"], DDP=self.DDP)
                gen_acml_loss = gen_acml_loss / self.OPTIMIZATION.acml_steps
                gen_acml_loss.backward()
            model_lib.allreduce_gradients(self.Gen)
            self.OPTIMIZATION.g_optimizer.step()
            if self.MODEL.apply_g_ema:
                self.ema.update(current_step)
        return gen_acml_loss

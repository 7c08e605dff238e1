"""Latent / label sampling and G(z, y) with the reference's RNG order (src/utils/sample.py:28-178): labels are drawn
first (torch.randint on the device), then z (torch.randn on the device).  ``generate_images`` returns the reference's
7-tuple so worker code can unpack it unchanged."""
import torch


def sample_normal(batch_size, z_dim, truncation_factor, device):
    if truncation_factor == -1.0:
        return torch.randn(batch_size, z_dim, device=device)
    if truncation_factor > 0:
        from scipy.stats import truncnorm
        vals = truncnorm.rvs(-truncation_factor, truncation_factor, size=[batch_size, z_dim])
        return torch.FloatTensor(vals).to(device)
    raise ValueError("truncated_factor must be positive.")


def sample_y(y_sampler, batch_size, num_classes, device):
    """src/utils/sample.py:41-66.  "acending_some" / "acending_all" are the visualisation samplers (8 images per class:
    a numpy permutation of the classes, or every class)."""
    if y_sampler == "totally_random":
        return torch.randint(low=0, high=num_classes, size=(batch_size,), dtype=torch.long, device=device)
    if y_sampler in ("acending_some", "acending_all"):
        if y_sampler == "acending_some":
            assert batch_size % 8 == 0, "The size of batches should be a multiple of 8."
            import numpy as np
            classes = np.random.permutation(num_classes)[:batch_size // 8]
        else:
            classes = range(num_classes)
        return torch.tensor([int(c) for c in classes for _ in range(8)], dtype=torch.long).to(device)
    if isinstance(y_sampler, int):
        return torch.tensor([y_sampler] * batch_size, dtype=torch.long).to(device)
    return None


def sample_zy(z_prior, batch_size, z_dim, num_classes, truncation_factor, y_sampler, radius, device):
    fake_labels = sample_y(y_sampler, batch_size, num_classes, device)
    if fake_labels is not None:
        batch_size = fake_labels.shape[0]
    if z_prior == "gaussian":
        zs = sample_normal(batch_size, z_dim, truncation_factor, device)
    elif z_prior == "uniform":
        zs = torch.FloatTensor(batch_size, z_dim).uniform_(-1.0, 1.0).to(device)
    else:
        raise NotImplementedError(z_prior)
    zs_eps = None
    if isinstance(radius, float) and radius > 0.0:
        if z_prior == "gaussian":
            zs_eps = zs + radius * sample_normal(batch_size, z_dim, -1.0, device)
        else:
            zs_eps = zs + radius * torch.FloatTensor(batch_size, z_dim).uniform_(-1.0, 1.0).to(device)
    return zs, fake_labels, zs_eps


def generate_images(z_prior, truncation_factor, batch_size, z_dim, num_classes, y_sampler, radius, generator, discriminator,
                    is_train, LOSS, RUN, MODEL, device, is_stylegan=False, generator_mapping=None, generator_synthesis=None,
                    style_mixing_p=0.0, stylegan_update_emas=False, cal_trsp_cost=False):
    if is_stylegan:
        raise NotImplementedError("StyleGAN is outside the sgb200 hot-path scope")
    if is_train:
        truncation_factor = -1.0
    zs, fake_labels, zs_eps = sample_zy(z_prior, batch_size, z_dim, num_classes, truncation_factor, y_sampler, radius, device)
    fake_images = generator(zs, fake_labels, eval=not is_train)
    fake_images_eps = generator(zs_eps, fake_labels, eval=not is_train) if zs_eps is not None else None
    return fake_images, fake_labels, fake_images_eps, None, None, None, None

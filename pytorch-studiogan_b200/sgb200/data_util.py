"""Data path (reference ``src/data_util.py:59-142``, ``src/utils/hdf5.py``, basket loader ``src/loader.py:178-193``).

The reference decodes, flips and normalises every sample on CPU workers and ships fp32 NCHW tensors to the GPU.  Here the
dataset stays what it is on disk -- uint8 NHWC (the HDF5 ``imgs`` / ``labels`` arrays the reference's ``make_hdf5`` writes,
an ``.npz`` with the same two arrays, or a decoded image folder) -- baskets are gathered as uint8 into pinned host
memory, copied to the device on a side stream while the previous step computes, and ``sgb_u8_to_img`` applies
RandomHorizontalFlip + ToTensor + Normalize(0.5, 0.5) there (bit-identical fp32 arithmetic).  PCIe carries 1 byte per
value instead of 4 and no CPU core touches a pixel.

``Dataset_`` keeps the reference's constructor so ``loader.py``-style code can build it; ``DeviceBasketLoader`` yields what
``WORKER.sample_data_basket_raw`` expects: (images [B * acml * d_updates, 3, H, W] fp32 in [-1, 1], labels int64), already
on the device.
"""
import os

import numpy as np
import torch


class Dataset_(torch.utils.data.Dataset):
    """uint8 image dataset.  Sources, in the reference's order of preference: ``hdf5_path`` (arrays ``imgs`` [N,H,W,3] uint8
    and ``labels`` [N]; an ``.npz`` with the same keys is accepted where h5py is not installed), else ``data_dir``/train|valid as
    an image folder (decoded once with PIL: centre-crop to the short edge if ``crop_long_edge``, resize to ``resize_size``)."""

    def __init__(self, data_name, data_dir, train, crop_long_edge=False, resize_size=None, resizer="lanczos", random_flip=False,
                 normalize=True, hdf5_path=None, load_data_in_memory=False):
        super().__init__()
        self.data_name, self.data_dir, self.train = data_name, data_dir, train
        self.random_flip, self.normalize = random_flip, normalize
        self.hdf5_path = hdf5_path
        if hdf5_path is not None:
            self.imgs, self.labels = _load_arrays(hdf5_path)
        else:
            self.imgs, self.labels, self.class_to_idx = _decode_folder(os.path.join(data_dir, "train" if train else "valid"),
                                                                       crop_long_edge, resize_size, resizer)
        assert self.imgs.dtype == np.uint8 and self.imgs.ndim == 4 and self.imgs.shape[3] == 3
        self.labels = np.asarray(self.labels).astype(np.int64)

    @classmethod
    def from_arrays(cls, imgs, labels, data_name="arrays", random_flip=False, normalize=True):
        """Dataset over in-memory arrays (``imgs`` uint8 [N,H,W,3], ``labels`` [N]) -- synthetic data, pre-loaded HDF5."""
        self = cls.__new__(cls)
        torch.utils.data.Dataset.__init__(self)
        self.data_name, self.data_dir, self.train, self.hdf5_path = data_name, None, True, None
        self.random_flip, self.normalize = random_flip, normalize
        self.imgs, self.labels = np.ascontiguousarray(imgs), np.asarray(labels).astype(np.int64)
        assert self.imgs.dtype == np.uint8 and self.imgs.ndim == 4 and self.imgs.shape[3] == 3
        return self

    def __len__(self):
        return self.imgs.shape[0]

    def __getitem__(self, index):
        """Per-sample access with the reference's semantics (fp32 CHW in [-1,1] if ``normalize`` else uint8 CHW); the training
        path does not use it -- it gathers uint8 batches (``gather``) and normalises on the device."""
        img = torch.from_numpy(self.imgs[index]).permute(2, 0, 1)
        if self.random_flip and float(torch.rand(1)) < 0.5:
            img = img.flip(2)
        if self.normalize:
            img = (img.to(torch.float32).div(255) - 0.5) / 0.5
        return img, int(self.labels[index])

    def gather(self, indices, out_imgs, out_labels):
        """uint8 NHWC rows ``indices`` into (pinned) ``out_imgs`` [n,H,W,3] / ``out_labels`` [n]."""
        idx = np.asarray(indices)
        np.take(self.imgs, idx, axis=0, out=out_imgs.numpy())
        out_labels.copy_(torch.from_numpy(self.labels[idx]))


def _load_arrays(path):
    if path.endswith(".npz"):
        z = np.load(path)
        return z["imgs"], z["labels"]
    try:
        import h5py
    except ImportError as ex:
        raise RuntimeError("reading %s needs h5py (not installed here); an .npz with arrays 'imgs' / 'labels' is accepted" % path) from ex
    with h5py.File(path, "r") as f:
        return f["imgs"][:], f["labels"][:]


def _decode_folder(root, crop_long_edge, resize_size, resizer):
    from PIL import Image
    filt = {"nearest": Image.NEAREST, "box": Image.BOX, "bilinear": Image.BILINEAR, "hamming": Image.HAMMING,
            "bicubic": Image.BICUBIC, "lanczos": Image.LANCZOS}
    classes = sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d)))
    class_to_idx = {c: i for i, c in enumerate(classes)}
    imgs, labels = [], []
    for c in classes:
        for name in sorted(os.listdir(os.path.join(root, c))):
            if not name.lower().endswith((".png", ".jpg", ".jpeg", ".bmp", ".webp")):
                continue
            im = Image.open(os.path.join(root, c, name)).convert("RGB")
            if crop_long_edge:
                s = min(im.size)
                left, top = int(round((im.size[0] - s) / 2.0)), int(round((im.size[1] - s) / 2.0))
                im = im.crop((left, top, left + s, top + s))
            if resize_size is not None and resizer != "wo_resize":
                im = im.resize((resize_size, resize_size), resample=filt[resizer])
            imgs.append(np.asarray(im, dtype=np.uint8))
            labels.append(class_to_idx[c])
    return np.stack(imgs, 0), np.asarray(labels), class_to_idx


class DeviceBasketLoader(object):
    """Infinite basket iterator (one item = ``basket`` = batch * acml_steps * d_updates_per_step samples, src/loader.py:178-193)
    with the device-side transform.  Two slots of (pinned staging, device uint8) buffers; a staging thread gathers the next
    basket (torch.index_select: the GIL is released during the copy, so the training thread keeps issuing kernels) and
    enqueues its H2D copy on a side stream while the consumer works on the current one; ``__next__`` makes the compute stream
    wait for that copy, runs the flip / normalise kernel and returns device tensors."""

    def __init__(self, dataset, basket, device, shuffle=True, random_flip=None, seed=0, drop_last=True):
        import queue
        import threading
        self.ds, self.basket, self.device = dataset, int(basket), torch.device(device)
        self.flip = dataset.random_flip if random_flip is None else random_flip
        self.rng = np.random.RandomState(seed)
        self.gen = torch.Generator().manual_seed(seed)
        self.shuffle = shuffle
        n, H, W, _ = dataset.imgs.shape
        assert n >= self.basket or not drop_last, "dataset smaller than one basket"
        self.imgs_t = torch.from_numpy(dataset.imgs)          # shares memory with the dataset
        self.labels_t = torch.from_numpy(dataset.labels)
        self.h_img = [torch.empty((self.basket, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.h_lab = [torch.empty(self.basket, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.h_flip = [torch.empty(self.basket, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.d_img = [torch.empty((self.basket, H, W, 3), dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.d_lab = [torch.empty(self.basket, dtype=torch.int64, device=self.device) for _ in range(2)]
        self.d_flip = [torch.empty(self.basket, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.stream = torch.cuda.Stream(device=self.device)
        self.ready = [None, None]             # H2D of the slot finished (recorded on self.stream)
        self.consumed = [None, None]          # the consumer's kernels on the slot's device buffers (recorded on the compute stream)
        self.order, self.pos = self._new_order(), 0
        self.h2d_bytes = self.basket * (H * W * 3 + 8 + (1 if self.flip else 0))
        self.free, self.full = queue.Queue(), queue.Queue()
        self.free.put(0)
        self.free.put(1)
        self._stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _new_order(self):
        n = len(self.ds)
        return self.rng.permutation(n) if self.shuffle else np.arange(n)

    def _run(self):
        torch.cuda.set_device(self.device)
        while not self._stop:
            slot = self.free.get()
            if slot is None:
                return
            self._stage(slot)
            self.full.put(slot)

    def _stage(self, slot):
        if self.ready[slot] is not None:
            self.ready[slot].synchronize()        # the copy that last read this slot's pinned buffers is done
        if self.pos + self.basket > len(self.order):
            self.order, self.pos = self._new_order(), 0
        idx = torch.from_numpy(np.ascontiguousarray(self.order[self.pos:self.pos + self.basket]).astype(np.int64))
        self.pos += self.basket
        torch.index_select(self.imgs_t, 0, idx, out=self.h_img[slot])
        torch.index_select(self.labels_t, 0, idx, out=self.h_lab[slot])
        if self.flip:
            self.h_flip[slot].copy_((torch.rand(self.basket, generator=self.gen) < 0.5).to(torch.uint8))
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])     # the kernels that read this slot's device buffers are done
            self.d_img[slot].copy_(self.h_img[slot], non_blocking=True)
            self.d_lab[slot].copy_(self.h_lab[slot], non_blocking=True)
            if self.flip:
                self.d_flip[slot].copy_(self.h_flip[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[slot] = ev

    def close(self):
        self._stop = True
        self.free.put(None)

    def __iter__(self):
        return self

    def __next__(self):
        from . import kernels as K
        slot = self.full.get()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[slot])
        imgs = K.u8_to_img(self.d_img[slot], self.d_flip[slot] if self.flip else None)
        labels = self.d_lab[slot].clone()
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[slot] = ev
        self.free.put(slot)                       # the staging thread refills it while this basket is being trained on
        return imgs, labels

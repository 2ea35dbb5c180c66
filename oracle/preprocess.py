"""Oracle (test infrastructure): the reference's host-side image preprocessing, restated on arrays.

Follows /root/reference/data/dataset.py:135-155 (PIL RGB image -> channels re-merged as (b, g, r) -> optional
tf.hflip) and /root/reference/data/dataloader.py:15-19 (transforms.ToTensor + Normalize([0.5]*3, [0.5]*3)).
Pinned against the real PIL / torchvision pipeline by tools/make_golden.py -> tests/golden/preprocess_ref.npz.
"""
import numpy as np
import torch


def preprocess(img_u8_hwc, flip=False, swap_rb=True):
    """img_u8_hwc: uint8 (H,W,3) RGB as decoded -> fp32 (3,H,W) in [-1,1], channel 0 = blue (dataset.py:138-141)."""
    a = np.asarray(img_u8_hwc)
    if swap_rb:
        a = a[:, :, ::-1]                      # Image.merge('RGB', (b, g, r))
    if flip:
        a = a[:, ::-1, :]                      # tf.hflip
    t = torch.from_numpy(np.ascontiguousarray(a)).permute(2, 0, 1).float().div(255)      # ToTensor
    return t.sub(0.5).div(0.5)                 # Normalize(0.5, 0.5)


def preprocess_batch(imgs_u8_nhwc, flips=None, swap_rb=True):
    n = len(imgs_u8_nhwc)
    flips = [False] * n if flips is None else list(flips)
    return torch.stack([preprocess(imgs_u8_nhwc[i], bool(flips[i]), swap_rb) for i in range(n)])


def synth_images_u8(n, size=112, seed=0):
    """Deterministic synthetic decoded images: uint8 (n,size,size,3)."""
    g = np.random.RandomState(seed)
    return g.randint(0, 256, size=(n, size, size, 3)).astype(np.uint8)

"""Device side of the view transforms (reference ``code/dataset.py:24-109``).

The reference builds every labeled / weak / strong view on the host -- PIL ops, then ``ToTensor`` and ``Normalize`` -- and
moves fp32 NCHW tensors to the GPU.  Here the host stops at the uint8 HWC image (what PIL and the decoder produce):

* ``draw_view_params``  draws the random decisions of ``RandomHorizontalFlip`` + ``RandomCrop(size, padding,
  padding_mode='reflect')`` (``dataset.py:35-38``) from torch's global generator in torchvision's own order, so a seeded
  run makes the same decisions as the reference's ``Compose``;
* ``normalize_views``   applies flip, reflect-padded crop, ``ToTensor`` and ``Normalize`` in ONE launch
  (``b200ssl_normalize_views``), bit-exact with torchvision in fp32.

A batch crosses PCIe at one byte per sample instead of four, and the per-image ToTensor/Normalize work leaves the loader
workers.  ``RandAugmentMC`` / ``ColorJitter`` (``randaugment.py:207-222``) stay PIL ops on the host: put them before the
uint8 hand-over (they precede ``ToTensor`` in the reference too).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _native as N

__all__ = ["IMAGENET_MEAN", "IMAGENET_STD", "draw_view_params", "normalize_views"]

IMAGENET_MEAN = (0.485, 0.456, 0.406)     # dataset.py:21-22
IMAGENET_STD = (0.229, 0.224, 0.225)


def draw_view_params(n: int, height: int, width: int, size: int, padding: int = 0, p_flip: float = 0.5,
                     crop: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Per-image ``(flip [n] int32, crop_xy [n, 2] int32 = (left, top))`` drawn from torch's global generator exactly like
    ``RandomHorizontalFlip(p)`` followed by ``RandomCrop(size, padding)`` would for ``n`` images in sequence
    (``torch.rand(1) < p``; then ``torch.randint(0, H + 2*padding - size + 1)`` for the top and the same for the left)."""
    flips, xy = [], []
    for _ in range(n):
        flips.append(int(torch.rand(1) < p_flip))
        if crop:
            ph, pw = height + 2 * padding, width + 2 * padding
            if ph == size and pw == size:
                top, left = 0, 0                       # torchvision returns without drawing
            else:
                top = int(torch.randint(0, ph - size + 1, size=(1,)).item())
                left = int(torch.randint(0, pw - size + 1, size=(1,)).item())
            xy.append((left, top))
    return torch.tensor(flips, dtype=torch.int32), (torch.tensor(xy, dtype=torch.int32) if crop else None)


def normalize_views(images_u8: torch.Tensor, size: Optional[int] = None, padding: int = 0, flip: Optional[torch.Tensor] = None,
                    crop_xy: Optional[torch.Tensor] = None, mean: Sequence[float] = IMAGENET_MEAN,
                    std: Sequence[float] = IMAGENET_STD, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """``[N, H, W, 3]`` uint8 CUDA images -> ``[N, 3, size, size]`` normalised views.

    ``flip`` (``[N]``, non-zero = mirrored) and ``crop_xy`` (``[N, 2]`` = left, top of the crop window in the image padded by
    ``padding`` with reflection) are per-image; ``crop_xy=None`` takes the centred window (the identity when
    ``size == H == W`` and ``padding == 0``).  fp32 output is bit-exact with ``ToTensor`` + ``Normalize``."""
    import ctypes
    dev = N.require_cuda(images_u8, flip, crop_xy, what="normalize_views")
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
        raise ValueError(f"expected uint8 [N, H, W, 3] images, got {images_u8.dtype} {tuple(images_u8.shape)}")
    x = images_u8.contiguous()
    n, h, w, _ = x.shape
    size = int(size) if size is not None else h
    out = torch.empty(n, 3, size, size, dtype=dtype, device=dev)
    fl = flip.to(torch.int32).contiguous() if flip is not None else None
    xy = crop_xy.to(torch.int32).contiguous() if crop_xy is not None else None
    if fl is not None and fl.numel() != n or xy is not None and tuple(xy.shape) != (n, 2):
        raise ValueError("flip is [N], crop_xy is [N, 2]")
    m3, s3 = (ctypes.c_float * 3)(*mean), (ctypes.c_float * 3)(*std)
    N.check(N.lib().b200ssl_normalize_views(x.data_ptr(), out.data_ptr(), n, h, w, size, int(padding), N.ptr(fl), N.ptr(xy), m3, s3,
                                            N.dtype_enum(out), N.stream_ptr(dev)), "normalize_views")
    return out

"""ATen (PyTorch fp32) restatement of the reference's training loss -- TEST INFRASTRUCTURE: the fp32 reference the fused
CUDA loss kernels (csrc/gms_loss.cuh) are compared with, and the loss of the reference-ordered A/B trainer arm.

l1_loss / ssim restate utils/loss_utils.py:17-64 (11x11 Gaussian window, sigma 1.5, grouped conv2d, zero padding);
training_loss = (1 - lambda) * L1 + lambda * (1 - SSIM), train.py:105-107.  Pinned to the reference's own outputs by
tests/golden/loss.npz (tests/test_oracle_golden.py)."""
from __future__ import annotations

from math import exp

import torch
import torch.nn.functional as F

_window_cache = {}


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()


def _window(size: int, channel: int, device, dtype):
    key = (size, channel, str(device), dtype)
    w = _window_cache.get(key)
    if w is None:
        g = torch.tensor([exp(-(x - size // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(size)])
        g = (g / g.sum()).unsqueeze(1)
        w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
        w = w2.expand(channel, 1, size, size).contiguous().to(device=device, dtype=dtype)
        _window_cache[key] = w
    return w


def ssim(img1, img2, window_size: int = 11):
    channel = img1.size(-3)
    w = _window(window_size, channel, img1.device, img1.dtype)
    pad = window_size // 2
    mu1 = F.conv2d(img1, w, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, w, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=pad, groups=channel) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=pad, groups=channel) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


def training_loss(image, gt, lambda_dssim: float = 0.2):
    """PyTorch (ATen/cuDNN) restatement -- the fp32 reference the fused kernel is tested against."""
    return (1.0 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1.0 - ssim(image, gt))



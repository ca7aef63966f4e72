"""Training loss of the reference: (1 - lambda) * L1 + lambda * (1 - SSIM)   (train.py:105-107).

l1_loss / ssim restate utils/loss_utils.py:17-64 (11x11 Gaussian window, sigma 1.5, grouped conv2d, zero padding).
The loss produces dL/dimage for the rasterizer backward; it is host-side glue here (ATen / cuDNN), listed as
"next" in SURVEY.md section 8(f) rank 2."""
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


class _FusedLoss(torch.autograd.Function):
    """gms_l1_ssim_loss: loss and dL/dimage in two launches (csrc/gms_loss.cuh)."""

    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        import ctypes as C
        from . import _lib
        if not image.is_cuda:
            raise RuntimeError("fused_training_loss: CUDA tensors required (use training_loss for the ATen reference)")
        L = _lib.lib()
        img = image.detach().contiguous().float()
        g = gt.detach().contiguous().float()
        Cn, H, W = img.shape[-3], img.shape[-2], img.shape[-1]
        nb = C.c_size_t()
        _lib.check(L.gms_loss_scratch_bytes(Cn, H, W, C.byref(nb)), "gms_loss_scratch_bytes")
        scratch = torch.empty(nb.value, dtype=torch.uint8, device=img.device)
        out = torch.empty(3, dtype=torch.float32, device=img.device)
        need_grad = image.requires_grad
        dimg = torch.empty_like(img) if need_grad else None
        a = _lib.LossArgs(Cn, H, W, img.data_ptr(), g.data_ptr(), float(lambda_dssim), None, out.data_ptr(),
                          dimg.data_ptr() if need_grad else None, scratch.data_ptr(), nb.value)
        with torch.cuda.device(img.device):
            _lib.check(L.gms_l1_ssim_loss(C.byref(a), torch.cuda.current_stream(img.device).cuda_stream), "gms_l1_ssim_loss")
        ctx.dimg = dimg
        ctx.stats = out
        return out[0]

    @staticmethod
    def backward(ctx, grad_loss):
        d = ctx.dimg
        if d is None:
            return None, None, None
        return d * grad_loss, None, None


def fused_training_loss(image, gt, lambda_dssim: float = 0.2):
    """(1-lambda)*L1 + lambda*(1-SSIM) through the fused CUDA kernels; the gradient w.r.t. `image` is produced in the
    same call and handed to autograd."""
    return _FusedLoss.apply(image, gt, lambda_dssim)

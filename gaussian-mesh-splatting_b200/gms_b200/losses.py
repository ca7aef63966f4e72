"""Training loss of the reference, (1 - lambda) * L1 + lambda * (1 - SSIM) (train.py:105-107, utils/loss_utils.py:17-64), through
the fused CUDA kernels of csrc/gms_loss.cuh: loss value and dL/dimage in two launches.
The ATen restatement the kernels are tested against lives in tests/aten_reference.py."""
from __future__ import annotations

import torch


class _FusedLoss(torch.autograd.Function):
    """gms_l1_ssim_loss: loss and dL/dimage in two launches (csrc/gms_loss.cuh)."""

    @staticmethod
    def forward(ctx, image, gt, lambda_dssim):
        import ctypes as C
        from . import _lib
        if not image.is_cuda:
            raise RuntimeError("fused_training_loss: CUDA tensors required (no CPU path in the product)")
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

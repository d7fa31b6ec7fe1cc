"""On-device reconstruction metrics.

  mse / nmse / psnr / ssim     mridc/collections/common/metrics/reconstruction_metrics.py:11-41
  normalized_magnitude         reconstruction/models/base.py:415-420 (``|x| / max|x|``)
  evaluate                     the metric block of test_step, base.py:415-436

The reference moves the prediction to the host and evaluates with numpy / scikit-image; here the reductions run on the
GPU (fp64 accumulation) and only the final scalars cross to the host.  Inputs are CUDA tensors (no CPU fallback).
"""
from typing import Dict, Optional

import torch

from . import _lib

__all__ = ["mse", "nmse", "psnr", "ssim", "normalized_magnitude", "evaluate", "METRIC_FUNCS"]


def _ws(B, device):
    n = _lib.load().mrb_metrics_workspace_bytes(int(B))
    return torch.empty((n + 7) // 8, dtype=torch.float64, device=device)


def _pair(gt, pred):
    gt = _lib.require_cuda(gt, "gt").contiguous()
    pred = _lib.require_cuda(pred, "pred").contiguous()
    if gt.shape != pred.shape:
        raise ValueError("operands could not be broadcast together with shapes %s %s" % (tuple(gt.shape), tuple(pred.shape)))
    return gt, pred


def _run(gt, pred, maxval_mode, maxval=0.0):
    """-> device tensor of 5 doubles (mse, nmse, psnr, ssim, data range); gt / pred [B, H, W]."""
    B, H, W = gt.shape
    res = torch.empty(5, dtype=torch.float64, device=gt.device)
    _lib.check(_lib.load().mrb_recon_metrics(_lib.ptr(gt), _lib.ptr(pred), B, H, W, maxval_mode, float(maxval),
                                             _lib.ptr(res), _lib.ptr(_ws(B, gt.device)), _lib.stream_ptr()))
    return res


def _as3d(gt, pred):
    gt, pred = _pair(gt, pred)
    if gt.dim() == 2:
        gt, pred = gt[None], pred[None]
    elif gt.dim() != 3:
        gt, pred = gt.reshape(-1, *gt.shape[-2:]), pred.reshape(-1, *pred.shape[-2:])
    return gt, pred


def mse(gt: torch.Tensor, pred: torch.Tensor) -> float:
    """reconstruction_metrics.py:11-13."""
    return float(_run(*_as3d(gt, pred), 4)[0])


def nmse(gt: torch.Tensor, pred: torch.Tensor) -> float:
    """reconstruction_metrics.py:16-18."""
    return float(_run(*_as3d(gt, pred), 4)[1])


def psnr(gt: torch.Tensor, pred: torch.Tensor, maxval: Optional[float] = None) -> float:
    """reconstruction_metrics.py:21-25."""
    gt, pred = _as3d(gt, pred)
    return float((_run(gt, pred, 4) if maxval is None else _run(gt, pred, 6, maxval))[2])


def ssim(gt: torch.Tensor, pred: torch.Tensor, maxval: Optional[float] = None) -> float:
    """reconstruction_metrics.py:28-41."""
    if gt.ndim != 3:
        raise ValueError("Unexpected number of dimensions in ground truth.")
    if gt.ndim != pred.ndim:
        raise ValueError("Ground truth dimensions does not match pred.")
    gt, pred = _pair(gt, pred)
    return float((_run(gt, pred, 0) if maxval is None else _run(gt, pred, 2, maxval))[3])


METRIC_FUNCS = dict(MSE=mse, NMSE=nmse, PSNR=psnr, SSIM=ssim)


def normalized_magnitude(x: torch.Tensor) -> torch.Tensor:
    """base.py:415-420: ``torch.abs(x) / torch.abs(x).max()`` for complex64 (or float32) input of any shape."""
    _lib.require_cuda(x, "x", None)
    if x.is_complex():
        if x.dtype != torch.complex64:
            raise TypeError("mridc_b200: x must be complex64 or float32")
        flat, cplx = torch.view_as_real(x.contiguous()), 1
    elif x.dtype == torch.float32:
        flat, cplx = x.contiguous(), 0
    else:
        raise TypeError("mridc_b200: x must be complex64 or float32")
    out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    if out.numel():
        _lib.check(_lib.load().mrb_abs_max_normalize(_lib.ptr(flat), out.numel(), cplx, _lib.ptr(out),
                                                     _lib.ptr(_ws(0, x.device)), _lib.stream_ptr()))
    return out


def evaluate(pred: torch.Tensor, target: torch.Tensor) -> Dict[str, float]:
    """The metric block of test_step (base.py:415-436): both images are magnitude / max normalised, the data range of
    PSNR and SSIM is ``output.max() - output.min()``.  pred / target: [B, h, w] complex64 (or float32).  One device ->
    host copy of five doubles."""
    output = normalized_magnitude(pred)
    tgt = normalized_magnitude(target)
    if tgt.dim() != 3:
        raise ValueError("Unexpected number of dimensions in ground truth.")
    tgt, output = _pair(tgt, output)
    r = _run(tgt, output, 1).cpu()
    return {"mse": float(r[0]), "nmse": float(r[1]), "psnr": float(r[2]), "ssim": float(r[3])}

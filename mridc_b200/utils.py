"""Drop-in for the hot-path functions of ``mridc.collections.common.parts.utils`` on B200.

complex_mul / complex_conj / complex_abs(_sq) / rss / rss_complex / sense / coil_combination run as
hand-written sm_100a kernels behind the C-ABI; the crop helpers are pure index arithmetic (views), kept in
Python exactly as in the reference.  Same signatures and error texts (utils.py line numbers cited).
"""
import ctypes
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

__all__ = [
    "is_none", "to_tensor", "tensor_to_complex_np", "complex_mul", "complex_conj", "complex_abs", "complex_abs_sq",
    "rss", "rss_complex", "sense", "coil_combination", "check_stacked_complex", "apply_mask", "mask_center",
    "batched_mask_center", "center_crop", "complex_center_crop", "center_crop_to_smallest",
]


def is_none(x) -> bool:
    """utils.py:38-50."""
    return x is None or str(x).lower() == "none"


def to_tensor(data: np.ndarray) -> torch.Tensor:
    """utils.py:53-71: complex numpy -> real tensor with trailing (re, im) dim."""
    if np.iscomplexobj(data):
        data = np.stack((data.real, data.imag), axis=-1)
    return torch.from_numpy(data)


def tensor_to_complex_np(data: torch.Tensor) -> np.ndarray:
    """utils.py:74-87."""
    data = data.cpu().numpy()
    return data[..., 0] + 1j * data[..., 1]


def _cview(x: torch.Tensor) -> torch.Tensor:
    """complex64 view of a float32 [...,2] tensor without copying when strides allow."""
    if x.stride(-1) != 1 or any(s % 2 for s in x.stride()[:-1]) or x.storage_offset() % 2:
        x = x.contiguous()
    return torch.view_as_complex(x)


def complex_mul(x: torch.Tensor, y: torch.Tensor, _conj_y: bool = False) -> torch.Tensor:
    """utils.py:96-118 (broadcasting allowed)."""
    if not x.shape[-1] == y.shape[-1] == 2:
        raise ValueError("Tensors do not have separate complex dim.")
    _lib.require_cuda(x, "x")
    _lib.require_cuda(y, "y")
    xc, yc = _cview(x), _cview(y)
    shape = torch.broadcast_shapes(xc.shape, yc.shape)
    xe, ye = xc.expand(shape), yc.expand(shape)
    if len(shape) > 6:
        xe, ye = xe.contiguous().reshape(-1), ye.contiguous().reshape(-1)
    out = torch.empty(shape, dtype=torch.complex64, device=x.device)
    if out.numel() == 0:  # empty batch: nothing to launch (data_ptr() is null)
        return torch.view_as_real(out)
    nd = xe.dim()
    arr = ctypes.c_longlong * max(nd, 1)
    _lib.check(_lib.load().mrb_complex_mul(
        _lib.ptr(xe), _lib.ptr(ye), _lib.ptr(out), nd, arr(*xe.shape), arr(*xe.stride()), arr(*ye.stride()),
        int(_conj_y), _lib.stream_ptr()))
    return torch.view_as_real(out)


def complex_conj(x: torch.Tensor) -> torch.Tensor:
    """utils.py:121-139."""
    if x.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    out = torch.empty_like(x)
    if out.numel() == 0:
        return out
    _lib.check(_lib.load().mrb_complex_conj(_lib.ptr(x), _lib.ptr(out), x.numel() // 2, _lib.stream_ptr()))
    return out


def _abs(data, squared):
    if data.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    _lib.require_cuda(data, "data")
    data = data.contiguous()
    out = torch.empty(data.shape[:-1], dtype=torch.float32, device=data.device)
    if out.numel() == 0:
        return out
    _lib.check(_lib.load().mrb_complex_abs(_lib.ptr(data), _lib.ptr(out), out.numel(), int(squared),
                                           _lib.stream_ptr()))
    return out


def complex_abs(data: torch.Tensor) -> torch.Tensor:
    """utils.py:142-157."""
    return _abs(data, False)


def complex_abs_sq(data: torch.Tensor) -> torch.Tensor:
    """utils.py:160-175."""
    return _abs(data, True)


def check_stacked_complex(data: torch.Tensor) -> torch.Tensor:
    """utils.py:178-191."""
    return torch.view_as_complex(data) if data.shape[-1] == 2 else data


def _split(shape, dim):
    outer = 1
    for s in shape[:dim]:
        outer *= s
    inner = 1
    for s in shape[dim + 1:]:
        inner *= s
    return outer, shape[dim], inner


def rss(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """utils.py:194-209: sqrt(sum(data**2, dim)) on the real view (re/im planes reduced separately)."""
    _lib.require_cuda(data, "data")
    data = data.contiguous()
    d = dim % data.dim()
    outer, C, inner = _split(data.shape, d)
    out = torch.empty(data.shape[:d] + data.shape[d + 1:], dtype=torch.float32, device=data.device)
    if out.numel() == 0 or C == 0:  # empty batch / no coils: an empty sum
        return out.zero_()
    _lib.check(_lib.load().mrb_rss(_lib.ptr(data), _lib.ptr(out), outer, C, inner, _lib.stream_ptr()))
    return out


def rss_complex(data: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """utils.py:212-227: sqrt(sum_dim |data|^2)."""
    if data.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    _lib.require_cuda(data, "data")
    data = data.contiguous()
    d = dim % (data.dim() - 1)  # `dim` indexes the tensor after the complex dim has been reduced
    cshape = data.shape[:-1]
    outer, C, inner = _split(cshape, d)
    out = torch.empty(cshape[:d] + cshape[d + 1:], dtype=torch.float32, device=data.device)
    if out.numel() == 0 or C == 0:
        return out.zero_()
    _lib.check(_lib.load().mrb_rss_complex(_lib.ptr(data), _lib.ptr(out), outer, C, inner, _lib.stream_ptr()))
    return out


def sense(data: torch.Tensor, sensitivity_maps: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """utils.py:230-248: complex_mul(data, complex_conj(S)).sum(dim)."""
    if not data.shape[-1] == sensitivity_maps.shape[-1] == 2:
        raise ValueError("Tensors do not have separate complex dim.")
    _lib.require_cuda(data, "data")
    _lib.require_cuda(sensitivity_maps, "sensitivity_maps")
    d = dim % data.dim()
    if data.shape != sensitivity_maps.shape or d == data.dim() - 1:
        # broadcasting / reduction over the complex dim: compose the two kernels' results
        return complex_mul(data, sensitivity_maps, _conj_y=True).sum(dim)
    data, sensitivity_maps = data.contiguous(), sensitivity_maps.contiguous()
    cshape = data.shape[:-1]
    outer, C, inner = _split(cshape, d)
    out = torch.empty(cshape[:d] + cshape[d + 1:] + (2,), dtype=torch.float32, device=data.device)
    if out.numel() == 0 or C == 0:
        return out.zero_()
    _lib.check(_lib.load().mrb_sense_combine(_lib.ptr(data), _lib.ptr(sensitivity_maps), _lib.ptr(out), outer, C,
                                             inner, _lib.stream_ptr()))
    return out


def coil_combination(data: torch.Tensor, sensitivity_maps: torch.Tensor, method: str = "SENSE",
                     dim: int = 0) -> torch.Tensor:
    """utils.py:251-272 (method is case-sensitive, as in the reference)."""
    if method == "SENSE":
        return sense(data, sensitivity_maps, dim)
    if method == "RSS":
        return rss(data, dim)
    raise ValueError("Output type not supported.")


def apply_mask(data: torch.Tensor, mask_func, seed=None, padding: Optional[Sequence[int]] = None,
               shift: bool = False, half_scan_percentage: Optional[float] = 0.0,
               center_scale: Optional[float] = 0.02, existing_mask: Optional[torch.Tensor] = None):
    """utils.py:293-343 (input preparation; the mask itself is generated on the host by ``mask_func``)."""
    shape = np.array(data.shape)
    shape[:-3] = 1
    if existing_mask is None:
        mask, acc = mask_func(shape, seed, half_scan_percentage=half_scan_percentage, scale=center_scale)
    else:
        mask = existing_mask
        acc = mask.size / mask.sum()
    mask = mask.to(data.device)
    if padding is not None and padding[0] != 0:
        mask[:, :, : padding[0]] = 0
        mask[:, :, padding[1]:] = 0
    if shift:
        mask = torch.fft.fftshift(mask, dim=(1, 2))
    masked_data = data * mask + 0.0
    return masked_data, mask, acc


def mask_center(x: torch.Tensor, mask_from, mask_to, mask_type: str = "2D") -> torch.Tensor:
    """utils.py:346-372."""
    mask = torch.zeros_like(x)
    if isinstance(mask_from, list):
        mask_from = mask_from[0]
    if isinstance(mask_to, list):
        mask_to = mask_to[0]
    if mask_type == "1D":
        mask[:, :, :, mask_from:mask_to] = x[:, :, :, mask_from:mask_to]
    elif mask_type == "2D":
        mask[:, :, mask_from:mask_to] = x[:, :, mask_from:mask_to]
    return mask


def batched_mask_center(x: torch.Tensor, mask_from: torch.Tensor, mask_to: torch.Tensor,
                        mask_type: str = "2D") -> torch.Tensor:
    """utils.py:375-410."""
    if mask_from.shape != mask_to.shape:
        raise ValueError("mask_from and mask_to must match shapes.")
    if mask_from.ndim != 1:
        raise ValueError("mask_from and mask_to must have 1 dimension.")
    if mask_from.shape[0] not in (1, x.shape[0]) or x.shape[0] != mask_to.shape[0]:
        raise ValueError("mask_from and mask_to must have batch_size length.")
    if mask_from.shape[0] == 1:
        mask = mask_center(x, int(mask_from), int(mask_to), mask_type=mask_type)
    else:
        mask = torch.zeros_like(x)
        for i, (start, end) in enumerate(zip(mask_from, mask_to)):
            mask[i, :, :, start:end] = x[i, :, :, start:end]
    return mask


def center_crop(data: torch.Tensor, shape: Tuple[int, int]) -> torch.Tensor:
    """utils.py:413-435 (index arithmetic only; returns a view)."""
    if not (0 < shape[0] <= data.shape[-2] and 0 < shape[1] <= data.shape[-1]):
        raise ValueError("Invalid shapes.")
    w_from = (data.shape[-2] - shape[0]) // 2
    h_from = (data.shape[-1] - shape[1]) // 2
    return data[..., w_from: w_from + shape[0], h_from: h_from + shape[1]]


def complex_center_crop(data: torch.Tensor, shape: Tuple[int, int]) -> torch.Tensor:
    """utils.py:438-460."""
    if not (0 < shape[0] <= data.shape[-3] and 0 < shape[1] <= data.shape[-2]):
        raise ValueError("Invalid shapes.")
    w_from = (data.shape[-3] - shape[0]) // 2
    h_from = (data.shape[-2] - shape[1]) // 2
    return data[..., w_from: w_from + shape[0], h_from: h_from + shape[1], :]


def center_crop_to_smallest(x: Union[torch.Tensor, np.ndarray], y: Union[torch.Tensor, np.ndarray]):
    """utils.py:463-486."""
    smallest_width = min(x.shape[-1], y.shape[-1])
    smallest_height = min(x.shape[-2], y.shape[-2])
    x = center_crop(x, (smallest_height, smallest_width))
    y = center_crop(y, (smallest_height, smallest_width))
    return x, y

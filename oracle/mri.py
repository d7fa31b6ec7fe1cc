"""Oracle (TEST INFRASTRUCTURE): MRI primitives restated on PyTorch-CPU.

Follows mridc/collections/common/parts/fft.py and utils.py (line numbers cited per function).
Tensors are "complex-last-2" fp32 unless noted.  Not imported by the product package.
"""
from typing import Optional, Sequence

import torch


def _shift_amounts(shape, dims, inverse: bool):
    # fft.py:276-279 (fftshift: n // 2) and fft.py:317-320 (ifftshift: (n + 1) // 2)
    return [((shape[d] + 1) // 2) if inverse else (shape[d] // 2) for d in dims]


def roll(data: torch.Tensor, shift: Sequence[int], dim: Sequence[int]) -> torch.Tensor:
    """fft.py:205-240 -- circular shift per (shift, dim) pair; ValueError on length mismatch."""
    if len(shift) != len(dim):
        raise ValueError("len(shift) must match len(dim)")
    for s, d in zip(shift, dim):
        # fft.py:169-202: narrow + cat == torch.roll along one dim (pure copy, bit-identical)
        data = torch.roll(data, int(s) % data.size(d), d)
    return data


def fftshift(data: torch.Tensor, dim: Optional[Sequence[int]] = None) -> torch.Tensor:
    """fft.py:243-281."""
    dim = list(range(data.dim())) if dim is None else list(dim)
    return roll(data, _shift_amounts(data.shape, dim, False), dim)


def ifftshift(data: torch.Tensor, dim: Optional[Sequence[int]] = None) -> torch.Tensor:
    """fft.py:284-322."""
    dim = list(range(data.dim())) if dim is None else list(dim)
    return roll(data, _shift_amounts(data.shape, dim, True), dim)


def _transform(data, centered, normalization, spatial_dims, inverse):
    # fft.py:66-86 / 144-164
    if data.shape[-1] == 2:
        data = torch.view_as_complex(data)
    dims = [-2, -1] if spatial_dims is None else list(spatial_dims)
    if centered:
        data = ifftshift(data, dim=dims)
    norm = normalization if normalization.lower() != "none" else None
    fn = torch.fft.ifft2 if inverse else torch.fft.fft2
    nd = data.dim()
    pdims = [d % nd for d in dims]
    if sorted(pdims) == list(range(nd - len(pdims), nd)):
        data = fn(data, dim=dims, norm=norm)
    else:
        # Same transform, evaluated over trailing contiguous axes: torch 2.11's MKL path corrupts the heap for
        # strided multi-dim C2C transforms over non-trailing dims (reproducer: ifft2 of a [3, n, 5] complex64
        # tensor over dims (0, 1)); the reference itself pins torch 1.12.
        last = list(range(nd - len(pdims), nd))
        data = fn(data.movedim(pdims, last).contiguous(), dim=last, norm=norm).movedim(last, pdims)
    if centered:
        data = fftshift(data, dim=dims)
    return torch.view_as_real(data)


def fft2(data, centered=False, normalization="backward", spatial_dims=None):
    """fft.py:13-88."""
    return _transform(data, centered, normalization, spatial_dims, False)


def ifft2(data, centered=False, normalization="backward", spatial_dims=None):
    """fft.py:91-166."""
    return _transform(data, centered, normalization, spatial_dims, True)


def complex_mul(x, y):
    """utils.py:96-118."""
    if not x.shape[-1] == y.shape[-1] == 2:
        raise ValueError("Tensors do not have separate complex dim.")
    xr, xi = x[..., 0], x[..., 1]
    yr, yi = y[..., 0], y[..., 1]
    return torch.stack((xr * yr - xi * yi, xr * yi + xi * yr), dim=-1)


def complex_conj(x):
    """utils.py:121-139."""
    if x.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    return torch.stack((x[..., 0], -x[..., 1]), dim=-1)


def complex_abs_sq(data):
    """utils.py:160-175."""
    if data.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    return (data**2).sum(dim=-1)


def complex_abs(data):
    """utils.py:142-157."""
    if data.shape[-1] != 2:
        raise ValueError("Tensor does not have separate complex dim.")
    return (data**2).sum(dim=-1).sqrt()


def check_stacked_complex(data):
    """utils.py:178-191."""
    return torch.view_as_complex(data) if data.shape[-1] == 2 else data


def rss(data, dim=0):
    """utils.py:194-209 (on complex-last-2 input the re/im planes are reduced separately)."""
    return torch.sqrt((data**2).sum(dim))


def rss_complex(data, dim=0):
    """utils.py:212-227."""
    return torch.sqrt(complex_abs_sq(data).sum(dim))


def sense(data, sensitivity_maps, dim=0):
    """utils.py:230-248."""
    return complex_mul(data, complex_conj(sensitivity_maps)).sum(dim)


def coil_combination(data, sensitivity_maps, method="SENSE", dim=0):
    """utils.py:251-272."""
    if method == "SENSE":
        return sense(data, sensitivity_maps, dim)
    if method == "RSS":
        return rss(data, dim)
    raise ValueError("Output type not supported.")


def center_crop(data, shape):
    """utils.py:413-435."""
    if not (0 < shape[0] <= data.shape[-2] and 0 < shape[1] <= data.shape[-1]):
        raise ValueError("Invalid shapes.")
    w0 = (data.shape[-2] - shape[0]) // 2
    h0 = (data.shape[-1] - shape[1]) // 2
    return data[..., w0 : w0 + shape[0], h0 : h0 + shape[1]]


def complex_center_crop(data, shape):
    """utils.py:438-460."""
    if not (0 < shape[0] <= data.shape[-3] and 0 < shape[1] <= data.shape[-2]):
        raise ValueError("Invalid shapes.")
    w0 = (data.shape[-3] - shape[0]) // 2
    h0 = (data.shape[-2] - shape[1]) // 2
    return data[..., w0 : w0 + shape[0], h0 : h0 + shape[1], :]


def center_crop_to_smallest(x, y):
    """utils.py:463-486."""
    sw = min(x.shape[-1], y.shape[-1])
    sh = min(x.shape[-2], y.shape[-2])
    return center_crop(x, (sh, sw)), center_crop(y, (sh, sw))


def apply_mask_existing(data, mask):
    """utils.py:325-343 with ``existing_mask`` given: data * mask + 0.0."""
    return data * mask + 0.0

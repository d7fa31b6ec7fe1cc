"""Drop-in for ``mridc.collections.common.parts.fft`` (fft2, ifft2, roll, fftshift, ifftshift) on B200.

Same signatures, argument meaning and error behaviour as the reference (file:line cited per function,
relative to the upstream repo root).  Work is done by hand-written sm_100a kernels behind the C-ABI
(``mrb_fft1d_c2c`` / ``mrb_fft2_c2c`` / ``mrb_roll``); CPU tensors raise -- there is no fallback.
"""
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

__all__ = ["fft2", "ifft2", "fftshift", "ifftshift", "roll", "roll_one_dim"]

_NORMS = {"backward": 0, "none": 0, "ortho": 1, "forward": 2}


def _norm_code(normalization: str) -> int:
    # fft.py:80 / :158: "none" (any case) maps to torch's default (= backward)
    key = normalization.lower() if normalization.lower() == "none" else normalization
    if key not in _NORMS:
        raise RuntimeError("Invalid normalization mode: %s" % normalization)
    return _NORMS[key]


def _as_complex(data: torch.Tensor) -> torch.Tensor:
    # fft.py:66-67: only a trailing dim of size 2 is viewed as complex; otherwise the input is taken as is
    if data.shape[-1] == 2 and not data.is_complex():
        _lib.require_cuda(data, "data", torch.float32)
        return torch.view_as_complex(data.contiguous())
    _lib.require_cuda(data, "data", None)
    if data.dtype == torch.float32:
        return torch.complex(data, torch.zeros_like(data))
    if data.dtype != torch.complex64:
        raise TypeError("mridc_b200: fft supports float32 / complex64 data (got %s)" % data.dtype)
    return data.contiguous()


def _transform(data, centered, normalization, spatial_dims, inverse):
    lib = _lib.load()
    x = _as_complex(data)
    if spatial_dims is None:
        dims = [-2, -1]
    else:
        dims = [int(d) for d in spatial_dims]  # ListConfig -> list (fft.py:71-72)
    nd = x.dim()
    dims = [d % nd if -nd <= d < nd else d for d in dims]
    for d in dims:
        if not 0 <= d < nd:
            raise IndexError("Dimension out of range")
    if len(set(dims)) != len(dims):
        raise RuntimeError("FFT dims must be unique")
    code = _norm_code(normalization)
    out = torch.empty_like(x)
    shape = list(x.shape)
    npts = 1
    for d in dims:
        npts *= shape[d]
    if x.numel() == 0 or not dims:
        out.copy_(x)
        return torch.view_as_real(out)
    if code == 1:
        scale = 1.0 / float(np.sqrt(npts))
    elif code == 0:
        scale = 1.0 / npts if inverse else 1.0
    else:
        scale = 1.0 if inverse else 1.0 / npts
    st = _lib.stream_ptr()
    if dims == [nd - 2, nd - 1] or dims == [nd - 1, nd - 2]:
        batch = 1
        for s in shape[:-2]:
            batch *= s
        _lib.check(lib.mrb_fft2_c2c(_lib.ptr(x), _lib.ptr(out), batch, shape[-2], shape[-1], int(inverse),
                                    int(bool(centered)), code, st))
        return torch.view_as_real(out)
    src = x
    for i, d in enumerate(dims):
        outer = 1
        for s in shape[:d]:
            outer *= s
        inner = 1
        for s in shape[d + 1:]:
            inner *= s
        n = shape[d]
        rot = n // 2 if centered else 0
        sc = scale if i == len(dims) - 1 else 1.0
        _lib.check(lib.mrb_fft1d_c2c(_lib.ptr(src), _lib.ptr(out), outer, n, inner, int(inverse), rot, rot, sc, st))
        src = out
    return torch.view_as_real(out)


def fft2(data: torch.Tensor, centered: bool = False, normalization: str = "backward",
         spatial_dims: Sequence[int] = None) -> torch.Tensor:
    """2-D FFT; mirrors mridc/collections/common/parts/fft.py:13-88."""
    return _transform(data, centered, normalization, spatial_dims, False)


def ifft2(data: torch.Tensor, centered: bool = False, normalization: str = "backward",
          spatial_dims: Sequence[int] = None) -> torch.Tensor:
    """2-D inverse FFT; mirrors fft.py:91-166."""
    return _transform(data, centered, normalization, spatial_dims, True)


def roll_one_dim(data: torch.Tensor, shift: int, dim: int) -> torch.Tensor:
    """fft.py:169-202 (bit-exact circular shift along one dim, any dtype)."""
    _lib.require_cuda(data, "data", None)
    n = data.size(dim)
    shift = int(shift) % n if n > 0 else 0
    if shift == 0:
        return data
    lib = _lib.load()
    x = data.contiguous()
    # complex/bool/etc. are moved as raw bytes
    es = x.element_size()
    out = torch.empty_like(x)
    d = dim % x.dim()
    outer = 1
    for s in x.shape[:d]:
        outer *= s
    inner = 1
    for s in x.shape[d + 1:]:
        inner *= s
    _lib.check(lib.mrb_roll(_lib.ptr(x), _lib.ptr(out), outer, n, inner, es, shift, _lib.stream_ptr()))
    return out


def roll(data: torch.Tensor, shift: List[int], dim: Union[List[int], Sequence[int]]) -> torch.Tensor:
    """fft.py:205-240."""
    if len(shift) != len(dim):
        raise ValueError("len(shift) must match len(dim)")
    dim = list(dim)
    for s, d in zip(shift, dim):
        data = roll_one_dim(data, s, d)
    return data


def fftshift(data: torch.Tensor, dim: Optional[Union[List[int], Sequence[int]]] = None) -> torch.Tensor:
    """fft.py:243-281 (shift by n // 2)."""
    dim = list(range(data.dim())) if dim is None else list(dim)
    shift = [int(np.floor_divide(data.shape[d], 2)) for d in dim]
    return roll(data, shift, dim)


def ifftshift(data: torch.Tensor, dim: Optional[Union[List[int], Sequence[int]]] = None) -> torch.Tensor:
    """fft.py:284-322 (shift by (n + 1) // 2)."""
    dim = list(range(data.dim())) if dim is None else list(dim)
    shift = [int(np.floor_divide(data.shape[d] + 1, 2)) for d in dim]
    return roll(data, shift, dim)

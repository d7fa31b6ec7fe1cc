"""Drop-in for ``VarNetBlock`` (mridc/collections/reconstruction/models/varnet/vn_block.py:12-119)."""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, _ops

__all__ = ["VarNetBlock"]


class VarNetBlock(nn.Module):
    """Soft data consistency + learned regulariser; the two FFT halves run as fused sm_100a operators."""

    def __init__(self, model: nn.Module, fft_centered: bool = True, fft_normalization: str = "ortho",
                 spatial_dims: Optional[Tuple[int, int]] = None, coil_dim: int = 1, no_dc: bool = False):
        super().__init__()
        self.model = model
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.spatial_dims = spatial_dims if spatial_dims is not None else [-2, -1]
        self.coil_dim = coil_dim
        self.no_dc = no_dc
        self.dc_weight = nn.Parameter(torch.ones(1))

    def _check(self):
        _ops.check_spatial_dims(self.spatial_dims)
        if self.coil_dim != 1:
            raise NotImplementedError("mridc_b200: VarNetBlock expects coil_dim == 1")

    def sens_expand(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        """vn_block.py:51-69: fft2(complex_mul(x, S)); x is [B, 1, H, W, 2]."""
        self._check()
        return _ops.sens_expand_softdc(x, sens_maps, None, None, None, None, None, True, self.fft_centered,
                                       self.fft_normalization)

    def sens_reduce(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        """vn_block.py:71-87: sum_c ifft2(x) * conj(S), keepdim -> [B, 1, H, W, 2]."""
        self._check()
        return _ops.sens_reduce(x, sens_maps, self.fft_centered, self.fft_normalization).unsqueeze(self.coil_dim)

    @torch.no_grad()
    def forward(self, pred: torch.Tensor, ref_kspace: torch.Tensor, sens_maps: torch.Tensor,
                mask: torch.Tensor) -> torch.Tensor:
        """vn_block.py:89-119."""
        self._check()
        pred = _lib.require_cuda(pred, "pred").contiguous()
        B, C, H, W, _ = pred.shape
        ws = torch.empty((1, B, C, H, W, 2), dtype=torch.float32, device=pred.device)
        eta = _ops.sens_reduce(pred, sens_maps, self.fft_centered, self.fft_normalization, ws=ws).unsqueeze(1)
        eta = self.model(eta)
        return _ops.sens_expand_softdc(eta, sens_maps, pred, pred, ref_kspace, mask, self.dc_weight.detach(),
                                       self.no_dc, self.fft_centered, self.fft_normalization, ws=ws)

"""Drop-in for ``CascadeNetBlock`` (mridc/collections/reconstruction/models/cascadenet/ccnn_block.py:12-139), the block
of CascadeNet and CRNNet: soft data consistency around a ``[B, 2, H, W]`` image regulariser.  SURVEY.md section 8 (f) 4:
the same fused operators as ``VarNetBlock`` (``mrb_sens_reduce``, ``mrb_sens_expand_softdc``), different model layout."""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, _ops

__all__ = ["CascadeNetBlock"]


class CascadeNetBlock(nn.Module):
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
            raise NotImplementedError("mridc_b200: CascadeNetBlock expects coil_dim == 1")

    def sens_expand(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        """ccnn_block.py:58-79: fft2(complex_mul(x, S)), x [B, 1, H, W, 2]."""
        self._check()
        return _ops.sens_expand_softdc(x, sens_maps, None, None, None, None, None, True, self.fft_centered,
                                       self.fft_normalization)

    def sens_reduce(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        """ccnn_block.py:81-102: sum_c ifft2(x) conj(S), keepdim."""
        self._check()
        return _ops.sens_reduce(x, sens_maps, self.fft_centered, self.fft_normalization).unsqueeze(self.coil_dim)

    @torch.no_grad()
    def forward(self, pred: torch.Tensor, ref_kspace: torch.Tensor, sens_maps: torch.Tensor,
                mask: torch.Tensor) -> torch.Tensor:
        """ccnn_block.py:104-139: pred - where(mask, pred - ref, 0) * dc_weight - expand(model(reduce(pred)))."""
        self._check()
        pred = _lib.require_cuda(pred, "pred").contiguous()
        B, C, H, W, _ = pred.shape
        ws = torch.empty((1, B, C, H, W, 2), dtype=torch.float32, device=pred.device)
        eta = _ops.sens_reduce(pred, sens_maps, self.fft_centered, self.fft_normalization, ws=ws)
        eta = self.model(eta.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        if eta.shape[0] * eta.shape[1] * eta.shape[2] * eta.shape[3] != B * H * W * 2:
            raise ValueError("the regulariser must return [B, 2, H, W] (got %s)" % (tuple(eta.permute(0, 3, 1, 2).shape),))
        return _ops.sens_expand_softdc(eta, sens_maps, pred, pred, ref_kspace, mask, self.dc_weight.detach(), self.no_dc,
                                       self.fft_centered, self.fft_normalization, ws=ws)

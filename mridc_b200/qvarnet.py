"""Drop-in for ``qVarNetBlock`` (mridc/collections/quantitative/models/qvarnet/qvn_block.py:13-160; SURVEY.md section 8
(f) 2): one cascade of the quantitative variational network -- MEGRE signal model of the current maps, soft data
consistency in k-space, SENSE reduce per echo, map regulariser.

Composition of this package's CUDA operators (``mrb_megre_signal``, centred FFT, strided-broadcast complex multiply);
the reference's axis conventions are kept, including the leading singleton it puts in front of the maps
(``unsqueeze(0)``, :137-140), which makes the block a batch-size-1 operator upstream as well.
"""
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, fft, utils
from .qrim import SignalForwardModel

__all__ = ["qVarNetBlock"]


class qVarNetBlock(nn.Module):
    def __init__(self, model: nn.Module, fft_centered: bool = True, fft_normalization: str = "ortho",
                 spatial_dims: Optional[Tuple[int, int]] = None, coil_dim: int = 1, no_dc: bool = False,
                 linear_forward_model=None):
        super().__init__()
        self.linear_forward_model = (SignalForwardModel(sequence="MEGRE") if linear_forward_model is None
                                     else linear_forward_model)
        self.model = model
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.spatial_dims = spatial_dims if spatial_dims is not None else [-2, -1]
        self.coil_dim = coil_dim
        self.no_dc = no_dc
        self.dc_weight = nn.Parameter(torch.ones(1))

    def _kw(self):
        return dict(centered=self.fft_centered, normalization=self.fft_normalization, spatial_dims=self.spatial_dims)

    def sens_expand(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        """qvn_block.py:63-82: image(s) -> coil k-space, F(S x)."""
        return fft.fft2(utils.complex_mul(x, sens_maps), **self._kw())

    def sens_reduce(self, x: torch.Tensor, sens_maps: torch.Tensor) -> torch.Tensor:
        """qvn_block.py:84-101: coil k-space -> image, sum over the coil axis of conj(S) F^-1 x (no keepdim)."""
        return utils.complex_mul(fft.ifft2(x, **self._kw()), sens_maps, _conj_y=True).sum(dim=self.coil_dim)

    @torch.no_grad()
    def forward(self, prediction: torch.Tensor, masked_kspace: torch.Tensor, R2star_map_init: torch.Tensor,
                S0_map_init: torch.Tensor, B0_map_init: torch.Tensor, phi_map_init: torch.Tensor, TEs: List,
                sensitivity_maps: torch.Tensor, sampling_mask: torch.Tensor, gamma: torch.Tensor = None) -> torch.Tensor:
        """qvn_block.py:103-160; ``prediction`` is unused upstream too."""
        _lib.require_cuda(masked_kspace, "masked_kspace")
        init_eta = torch.stack([R2star_map_init, S0_map_init, B0_map_init, phi_map_init], dim=1)
        g = [float(v) for v in gamma]  # type: ignore
        maps = [m * s for m, s in zip((R2star_map_init, S0_map_init, B0_map_init, phi_map_init), g)]
        # the reference stacks the echoes at dim 1 of [1, B, H, W] maps: [1, E, B, H, W, 2]
        signal = self.linear_forward_model(*maps, TEs)  # [B, E, H, W, 2]
        init_pred = signal.transpose(0, 1).unsqueeze(0)
        S = sensitivity_maps.unsqueeze(self.coil_dim - 1)
        pred_kspace = self.sens_expand(init_pred, S)
        soft_dc = (pred_kspace - masked_kspace) * sampling_mask * self.dc_weight
        init_pred = self.sens_reduce(soft_dc, S)
        eta = torch.view_as_real(init_eta + torch.view_as_complex(self.model(init_pred)))
        eta[:, 0, ...] = eta[:, 0, ...].clamp_min(0)  # :156-158 (NaNs stay NaNs, like the masked assignment)
        return eta

"""Drop-in for the coil-sensitivity estimation network of the end-to-end models.

  BaseSensitivityModel   mridc/collections/reconstruction/models/base.py:715-932

``ifft2`` of the auto-calibration region -> per-coil ``NormUnet`` (coils folded into the batch) -> division by the
root-sum-of-squares over coils; every step runs on the sm_100a kernels (FFT engine, exact-fp32 U-Net kernels,
``mrb_divide_rss``).  The low-frequency bookkeeping on the mask (a handful of integers) stays in torch, as in the
reference.  Inference only.
"""
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib, fft, utils
from .unet import NormUnet

__all__ = ["BaseSensitivityModel"]


class BaseSensitivityModel(nn.Module):
    """base.py:715-932."""

    def __init__(self, chans: int = 8, num_pools: int = 4, in_chans: int = 2, out_chans: int = 2, drop_prob: float = 0.0,
                 padding_size: int = 15, mask_type: str = "2D", fft_centered: bool = True,
                 fft_normalization: str = "ortho", spatial_dims: Sequence[int] = None, coil_dim: int = 1,
                 normalize: bool = True, mask_center: bool = True):
        super().__init__()
        self.mask_type = mask_type
        self.norm_unet = NormUnet(chans, num_pools, in_chans=in_chans, out_chans=out_chans, drop_prob=drop_prob,
                                  padding_size=padding_size, normalize=normalize)
        self.mask_center = mask_center
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.spatial_dims = spatial_dims if spatial_dims is not None else [-2, -1]
        self.coil_dim = coil_dim
        self.normalize = normalize

    @staticmethod
    def chans_to_batch_dim(x: torch.Tensor) -> Tuple[torch.Tensor, int]:
        """base.py:789-806."""
        b, c, h, w, comp = x.shape
        return x.view(b * c, 1, h, w, comp), b

    @staticmethod
    def batch_chans_to_chan_dim(x: torch.Tensor, batch_size: int) -> torch.Tensor:
        """base.py:808-824."""
        bc, _, h, w, comp = x.shape
        c = bc // batch_size
        return x.view(batch_size, c, h, w, comp)

    @staticmethod
    def divide_root_sum_of_squares(x: torch.Tensor, coil_dim: int) -> torch.Tensor:
        """base.py:826-840: x / rss_complex(x, coil_dim) in one kernel."""
        _lib.require_cuda(x, "x")
        if x.shape[-1] != 2:
            raise ValueError("Tensor does not have separate complex dim.")
        x = x.contiguous()
        d = coil_dim % (x.dim() - 1)
        outer, C, inner = utils._split(x.shape[:-1], d)
        out = torch.empty_like(x)
        _lib.check(_lib.load().mrb_divide_rss(_lib.ptr(x), _lib.ptr(out), outer, C, inner, _lib.stream_ptr()))
        return out

    @staticmethod
    def get_pad_and_num_low_freqs(mask: torch.Tensor,
                                  num_low_frequencies: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """base.py:842-878 (integer bookkeeping on the mask, bit-exact)."""
        if num_low_frequencies is None or num_low_frequencies == 0:
            squeezed_mask = mask[:, 0, 0, :, 0].to(torch.int8)
            cent = squeezed_mask.shape[1] // 2
            left = torch.argmin(squeezed_mask[:, :cent].flip(1), dim=1)  # first zero left / right of the centre
            right = torch.argmin(squeezed_mask[:, cent:], dim=1)
            num_low_frequencies_tensor = torch.max(2 * torch.min(left, right), torch.ones_like(left))
        else:
            num_low_frequencies_tensor = num_low_frequencies * torch.ones(mask.shape[0], dtype=mask.dtype,
                                                                          device=mask.device)
        pad = torch.div(mask.shape[-2] - num_low_frequencies_tensor + 1, 2, rounding_mode="trunc")
        return pad, num_low_frequencies_tensor

    @torch.no_grad()
    def forward(self, masked_kspace: torch.Tensor, mask: torch.Tensor,
                num_low_frequencies: Optional[int] = None) -> torch.Tensor:
        """base.py:880-932: [B, C, H, W, 2] k-space, [B, 1, H|1, W, 1] mask -> [B, C, H, W, 2] sensitivity maps."""
        _lib.require_cuda(masked_kspace, "masked_kspace")
        if self.mask_center:
            pad, num_low_freqs = self.get_pad_and_num_low_freqs(mask, num_low_frequencies)
            masked_kspace = utils.batched_mask_center(masked_kspace, pad, pad + num_low_freqs, mask_type=self.mask_type)
        images, batches = self.chans_to_batch_dim(
            fft.ifft2(masked_kspace, centered=self.fft_centered, normalization=self.fft_normalization,
                      spatial_dims=self.spatial_dims))
        images = self.batch_chans_to_chan_dim(self.norm_unet(images), batches)
        if self.normalize:
            images = self.divide_root_sum_of_squares(images, self.coil_dim)
        return images

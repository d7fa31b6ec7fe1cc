"""Drop-in for ``JRSCIRIMBlock`` (mridc/collections/segmentation/models/jrscirim_base/jrscirim_block.py:19-377; SURVEY.md
section 8 (f) 4): joint reconstruction + segmentation -- a CIRIM (cascades of ``RIMBlock``, the DC operator's consumer) whose
last estimate feeds a segmentation head.

The reconstruction half is this package's ``RIMBlock`` (fused DC gradient, tensor-core regulariser, one hybrid k-space per
forward shared by the cascades).  Segmentation heads that exist in this package are built the way the reference builds
them -- ``"UNet"`` (``unet_block.Unet``) and ``"ConvLayer"`` (one ``ConvNonlinear``); AttentionUNet / LambdaUNet / VNet are
segmentation-only networks outside the reconstruction path and raise.  Same sub-module names as the reference
(``reconstruction_module.<i>.*``, ``segmentation_module.*``, ``dc_weight``).
"""
import math
from typing import Any, Dict, List, Optional, Tuple

import torch

from . import _lib, _ops, fft, utils
from .rim import ConvNonlinear, RIMBlock
from .unet import Unet

__all__ = ["JRSCIRIMBlock"]


class JRSCIRIMBlock(torch.nn.Module):
    def __init__(self, reconstruction_module_params: Dict, segmentation_module_params: Dict, input_channels: int,
                 magnitude_input: bool = True, fft_centered: bool = True, fft_normalization: str = "ortho",
                 spatial_dims: Optional[Tuple[int, int]] = None, coil_dim: int = 1, dimensionality: int = 2,
                 consecutive_slices: int = 1, coil_combination_method: str = "SENSE",
                 normalize_segmentation_output: bool = True):
        super().__init__()
        self.input_channels = input_channels
        self.magnitude_input = magnitude_input
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.spatial_dims = spatial_dims
        self.coil_dim = coil_dim
        self.dimensionality = dimensionality
        if self.dimensionality != 2:
            raise NotImplementedError(f"Currently only 2D is supported for segmentation, got {self.dimensionality}D.")
        self.consecutive_slices = consecutive_slices
        self.coil_combination_method = coil_combination_method

        rp = self.reconstruction_module_params = reconstruction_module_params  # jrscirim_block.py:96-131
        self.reconstruction_module_recurrent_filters = rp["recurrent_filters"]
        self.reconstruction_module_time_steps = 8 * math.ceil(rp["time_steps"] / 8)
        self.no_dc = rp["no_dc"]
        self.keep_eta = rp["keep_eta"]
        self.reconstruction_module_dimensionality = rp["dimensionality"]
        slices = self.consecutive_slices if self.reconstruction_module_dimensionality == 3 else 1
        self.reconstruction_module = torch.nn.ModuleList([
            RIMBlock(recurrent_layer=rp["recurrent_layer"], conv_filters=rp["conv_filters"], conv_kernels=rp["conv_kernels"],
                     conv_dilations=rp["conv_dilations"], conv_bias=rp["conv_bias"],
                     recurrent_filters=self.reconstruction_module_recurrent_filters,
                     recurrent_kernels=rp["recurrent_kernels"], recurrent_dilations=rp["recurrent_dilations"],
                     recurrent_bias=rp["recurrent_bias"], depth=rp["depth"], time_steps=self.reconstruction_module_time_steps,
                     conv_dim=rp["conv_dim"], no_dc=self.no_dc, fft_centered=self.fft_centered,
                     fft_normalization=self.fft_normalization, spatial_dims=self.spatial_dims, coil_dim=self.coil_dim - 1,
                     dimensionality=self.reconstruction_module_dimensionality, consecutive_slices=slices)
            for _ in range(rp["num_cascades"])])
        self.reconstruction_module_keep_eta = rp["keep_eta"]
        # :136-138: rnn_weights_init touches only Linear / Embedding / LayerNorm -> a no-op for these convolutions
        self.dc_weight = torch.nn.Parameter(torch.ones(1))
        self.reconstruction_module_accumulate_estimates = rp["accumulate_estimates"]

        sp = self.segmentation_module_params = segmentation_module_params  # :142-196
        kind = sp["segmentation_module"]
        self.segmentation_module_output_channels = sp["output_channels"]
        if kind.lower() == "unet":
            seg = Unet(in_chans=self.input_channels, out_chans=self.segmentation_module_output_channels, chans=sp["channels"],
                       num_pool_layers=sp["pooling_layers"], drop_prob=sp["dropout"])
        elif kind.lower() == "convlayer":
            seg = torch.nn.Sequential(ConvNonlinear(self.input_channels, self.segmentation_module_output_channels,
                                                    conv_dim=sp["conv_dim"], kernel_size=3, dilation=1, bias=False,
                                                    nonlinear=None))
        elif kind.lower() in ("attentionunet", "lambdaunet", "vnet"):
            raise NotImplementedError("mridc_b200: the %s segmentation head is outside the reconstruction path" % kind)
        else:
            raise ValueError(f"Segmentation module {kind} not implemented.")
        self.segmentation_module = seg
        self.normalize_segmentation_output = normalize_segmentation_output

    @torch.no_grad()
    def forward(self, y: torch.Tensor, sensitivity_maps: torch.Tensor, mask: torch.Tensor,
                init_reconstruction_pred: torch.Tensor, target_reconstruction: torch.Tensor, hx: torch.Tensor = None,
                sigma: float = 1.0) -> Tuple[List[Any], Any, Optional[Any]]:
        """jrscirim_block.py:200-333 -> (list[cascades] of list[time steps] of complex images, segmentation, hx)."""
        _lib.require_cuda(y, "y")
        if self.consecutive_slices > 1 and self.reconstruction_module_dimensionality == 2:
            # :233-271: slice by slice through the 2-D reconstruction module
            per_slice = []
            for s in range(self.consecutive_slices):
                y_s = y[:, s, ...].contiguous()
                S_s = sensitivity_maps[:, s, ...].contiguous()
                init_s = init_reconstruction_pred[:, s, ...]
                init_s = None if init_s is None or init_s.dim() < 4 else init_s
                cas, hx = self._cascades(y_s, S_s, mask[:, 0, ...], init_s, target_reconstruction[:, s, ...], hx, sigma)
                per_slice.append(torch.stack([torch.stack(c, dim=0) for c in cas], dim=0))
            preds = torch.stack(per_slice, dim=3)
            cascades_etas = [[preds[c, t, ...] for t in range(preds.shape[1])] for c in range(preds.shape[0])]
        else:
            init = (None if init_reconstruction_pred is None or init_reconstruction_pred.dim() < 4
                    else init_reconstruction_pred)
            cascades_etas, hx = self._cascades(y, sensitivity_maps, mask, init, target_reconstruction, hx, 1.0)
        pred_reconstruction = cascades_etas

        x = pred_reconstruction[-1][-1]  # :297-325: the last estimate is the segmentation input
        if x.shape[-1] != 2:
            x = torch.view_as_real(x)
        if self.consecutive_slices > 1 and x.dim() == 5:
            x = x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])
        if x.shape[-1] == 2:
            if self.input_channels == 1:
                x = torch.view_as_complex(x.contiguous()).unsqueeze(1)
                if self.magnitude_input:
                    x = torch.abs(x)
            elif self.input_channels == 2:
                if self.magnitude_input:
                    raise ValueError("Magnitude input is not supported for 2-channel input.")
                x = x.permute(0, 3, 1, 2)
            else:
                raise ValueError("The input channels must be either 1 or 2. Found: {}".format(self.input_channels))
        else:
            x = x.unsqueeze(1)
        x = torch.nn.functional.group_norm(x, num_groups=1)  # :327-328 (one group over the whole sample, no affine)
        pred_segmentation = torch.abs(self.segmentation_module(x.contiguous()))
        if self.normalize_segmentation_output:
            pred_segmentation = pred_segmentation / torch.max(pred_segmentation)
        if self.consecutive_slices > 1:
            pred_segmentation = pred_segmentation.view([y.shape[0], y.shape[1], *pred_segmentation.shape[1:]])
        return pred_reconstruction, pred_segmentation, hx

    def _cascades(self, y, sensitivity_maps, mask, init_pred, target, hx, sigma):
        """:273-295 (and the per-slice twin :246-268): the CIRIM loop; one hybrid k-space per call shared by the cascades."""
        prediction = y  # the reference clones; nothing below writes y
        y = y.contiguous()
        hybrid_cache: dict = {}
        cascades_etas = []
        for i, cascade in enumerate(self.reconstruction_module):
            prediction, hx = cascade(prediction, y, sensitivity_maps, mask, init_pred, hx, sigma,
                                     keep_eta=False if i == 0 else self.keep_eta, y_hybrid=hybrid_cache)
            cascades_etas.append([self.process_intermediate_pred(p, sensitivity_maps, target) for p in prediction])
        return cascades_etas, hx

    def process_intermediate_pred(self, pred, sensitivity_maps, target, do_coil_combination=False):
        """:335-377."""
        if not self.no_dc or do_coil_combination:
            if self.coil_combination_method == "SENSE" and self.coil_dim == 1 and pred.dim() == 5:
                _ops.check_spatial_dims(self.spatial_dims)
                pred = _ops.sens_reduce(pred, sensitivity_maps, self.fft_centered, self.fft_normalization)
            else:
                pred = fft.ifft2(pred, centered=self.fft_centered, normalization=self.fft_normalization,
                                 spatial_dims=self.spatial_dims)
                pred = utils.coil_combination(pred, sensitivity_maps, method=self.coil_combination_method, dim=self.coil_dim)
        pred = torch.view_as_complex(pred.contiguous())
        if target.shape[-1] == 2:
            target = torch.view_as_complex(target.contiguous())
        _, pred = utils.center_crop_to_smallest(target, pred)
        return pred

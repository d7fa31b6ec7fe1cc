"""Drop-ins for the quantitative (qMRI) RIM building blocks of the reference (BASELINE.json configs[4]).

  RescaleByMax, SignalForwardModel, expand_op, analytical_log_likelihood_gradient
                mridc/collections/quantitative/models/qrim/utils.py:12-295
  qRIMBlock     mridc/collections/quantitative/models/qrim/qrim_block.py:13-240

The data-consistency part of the analytic gradient is the same fused operator as the RIM log-likelihood gradient
(``mrb_dc_rim_grad``) with the echoes folded into the batch; the MEGRE signal model in front of it and the analytic
d/d(R2*, S0) behind it are pointwise kernels (``csrc/qmri.cu``).  The reference evaluates the gradient from the
``*_init`` maps, which never change inside the time loop (qrim_block.py:196-224), so the block computes it ONCE per
forward instead of ``time_steps`` times (bit-identical to re-evaluating it).
Inference only; CUDA fp32 tensors only (no CPU fallback).
"""
import ctypes
from typing import Any, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import _lib, _ops
from .rim import ConvGRUCell, ConvMGUCell, ConvNonlinear, ConvRNNStack, IndRNNCell

__all__ = ["RescaleByMax", "SignalForwardModel", "expand_op", "analytical_log_likelihood_gradient", "qRIMBlock"]


def _tes_list(TEs) -> List[float]:
    if isinstance(TEs, torch.Tensor):
        return [float(t) for t in TEs.detach().reshape(-1).cpu()]
    return [float(t) for t in TEs]


def _c_doubles(vals):
    return (ctypes.c_double * len(vals))(*vals)


def _c_floats(vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals]) if vals is not None else None


def _maps(R2star_map, S0_map, B0_map, phi_map):
    """-> four contiguous CUDA fp32 tensors of one common shape [..., H, W]."""
    out = []
    for name, t in (("R2star_map", R2star_map), ("S0_map", S0_map), ("B0_map", B0_map), ("phi_map", phi_map)):
        out.append(_lib.require_cuda(t, name).contiguous())
        if out[-1].shape != out[0].shape:
            raise ValueError("quantitative maps must share one shape (got %s and %s)" % (
                tuple(out[0].shape), tuple(out[-1].shape)))
    return out


class RescaleByMax:
    """qrim/utils.py:12-28."""

    def __init__(self, slack=1e-6):
        self.slack = slack

    def forward(self, data):
        gamma = torch.max(torch.max(torch.abs(data), 3, keepdim=True)[0], 2, keepdim=True)[0] + self.slack
        return data / gamma, gamma

    @staticmethod
    def reverse(data, gamma, _take_abs=False):
        """``stack([data[i] * gamma[i] for i in range(batch)])`` -- the factor is indexed by the BATCH position
        (qrim/utils.py:27-28), so ``len(gamma)`` bounds the batch size exactly as in the reference."""
        _lib.require_cuda(data, "data")
        B = data.shape[0]
        if B > len(gamma):
            raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (len(gamma), len(gamma)))
        data = data.contiguous()
        out = torch.empty_like(data)
        for b0 in range(0, B, 4):  # the entry point takes up to four per-sample factors by value
            nb = min(4, B - b0)
            scales = _c_floats([float(gamma[b0 + i]) for i in range(nb)])
            _lib.check(_lib.load().mrb_scale_batch(_lib.ptr(data[b0:]), _lib.ptr(out[b0:]), nb, data[0].numel(), scales,
                                                   int(_take_abs), _lib.stream_ptr()))
        return out

    @staticmethod
    def reverse_abs(data, gamma):
        """``reverse(torch.abs(data), gamma)`` in one pass (qcirim.py:287-289)."""
        return RescaleByMax.reverse(data, gamma, _take_abs=True)


class SignalForwardModel:
    """qrim/utils.py:31-155: MEGRE / MEGRE-no-phase signal model, maps [B, H, W] -> [B, n_echoes, H, W, 2]."""

    def __init__(self, sequence: Union[str, None] = None):
        super().__init__()
        self.sequence = sequence.lower() if isinstance(sequence, str) else None
        self.scaling = 1e-3

    def __call__(self, R2star_map, S0_map, B0_map, phi_map, TEs=None):
        if TEs is None:
            TEs = [3.0, 11.5, 20.0, 28.5]
        if self.sequence == "megre":
            return self.MEGRESignalModel(R2star_map, S0_map, B0_map, phi_map, TEs)
        if self.sequence == "megre_no_phase":
            return self.MEGRENoPhaseSignalModel(R2star_map, S0_map, TEs)
        raise ValueError(
            "Only MEGRE and MEGRE no phase are supported are signal forward model at the moment. "
            f"Found {self.sequence}"
        )

    def _run(self, maps, TEs, no_phase, gamma=None):
        tes = _tes_list(TEs)
        ref = maps[0]
        if ref.dim() < 2:
            raise ValueError("quantitative maps must be [batch_size, n_x, n_y]")
        lead = ref.shape[:-2]
        B = 1
        for s in lead:
            B *= int(s)
        HW = int(ref.shape[-2] * ref.shape[-1])
        out = torch.empty((*lead, len(tes), ref.shape[-2], ref.shape[-1], 2), dtype=torch.float32, device=ref.device)
        _lib.check(_lib.load().mrb_megre_signal(
            _lib.ptr(maps[0]), _lib.ptr(maps[1]), _lib.ptr(maps[2]) if not no_phase else None,
            _lib.ptr(maps[3]) if not no_phase else None, _c_floats(gamma), _c_doubles(tes), len(tes), float(self.scaling),
            B, HW, int(no_phase), _lib.ptr(out), _lib.stream_ptr()))
        return out

    def MEGRESignalModel(self, R2star_map, S0_map, B0_map, phi_map, TEs):
        return self._run(_maps(R2star_map, S0_map, B0_map, phi_map), TEs, False)

    def MEGRENoPhaseSignalModel(self, R2star_map, S0_map, TEs):
        r2, s0 = _maps(R2star_map, S0_map, R2star_map, S0_map)[:2]
        return self._run([r2, s0, None, None], TEs, True)


def expand_op(x, sensitivity_maps):
    """qrim/utils.py:158-163 (complex_mul with NaN -> 0)."""
    from . import utils

    x = utils.complex_mul(x, sensitivity_maps)
    return torch.nan_to_num_(x, nan=0.0, posinf=float("inf"), neginf=float("-inf"))


def _sampling_mask_3d(sampling_mask, B, H, W):
    """Reference masks broadcast against [B, E, C, H, W, 2]: [B|1, 1, (1,) H|1, W, 1] -> [mb, mh, W] view."""
    m = _lib.require_cuda(sampling_mask, "sampling_mask", None)
    if m.dim() == 6 and m.shape[1] == 1:
        m = m[:, 0]
    if m.dim() == 5 and m.shape[1] == 1 and m.shape[4] in (1, 2):
        m = m[:, 0, :, :, 0]
    elif m.dim() == 4 and m.shape[3] in (1, 2) and m.shape[0] in (1, B):
        m = m[..., 0]
    elif m.dim() != 3:
        raise ValueError("unsupported sampling_mask shape %s" % (tuple(sampling_mask.shape),))
    if m.shape[0] not in (1, B) or m.shape[1] not in (1, H) or m.shape[2] != W:
        raise ValueError("sampling_mask %s does not broadcast against k-space [%d, E, C, %d, %d, 2]" % (
            tuple(sampling_mask.shape), B, H, W))
    return m


def _qmri_gradient(model: SignalForwardModel, maps, gamma, tes, sens, y, sampling_mask, fft_centered,
                   fft_normalization, scaling, divisor, zero_nan, out, out_channels):
    """Batched analytic gradient: maps 4 x [B,H,W], sens [B,C,H,W,2], y [B,E,C,H,W,2] -> out[:, :4] ([B,>=4,H,W])."""
    if model.sequence not in ("megre", "megre_no_phase"):
        model(*maps, tes)  # raises the reference's ValueError
    B, E, C, H, W, _ = y.shape
    if len(tes) != E:
        raise ValueError("n_echoes mismatch: %d echo times, k-space %s" % (len(tes), tuple(y.shape)))
    eta = model._run(maps if model.sequence == "megre" else [maps[0], maps[1], None, None], tes,
                     model.sequence != "megre", gamma).reshape(B * E, H, W, 2)
    m = _sampling_mask_3d(sampling_mask, B, H, W)
    if m.shape[0] != 1:
        m = m[:, None].expand(B, E, m.shape[1], W).reshape(B * E, m.shape[1], W)
    # the echoes share the coil maps: the fused operator wants one map set per batch entry
    S = sens[:, None].expand(B, E, C, H, W, 2).reshape(B * E, C, H, W, 2)
    d = _ops.dc_rim_grad(eta, y.reshape(B * E, C, H, W, 2), S, m, 1.0, fft_centered, fft_normalization)
    _lib.check(_lib.load().mrb_megre_grad(
        _lib.ptr(d), _lib.ptr(maps[0]), _lib.ptr(maps[1]), _lib.ptr(maps[2]), _lib.ptr(maps[3]), _c_floats(gamma),
        _c_doubles(tes), E, float(scaling), B, H * W, float(divisor), int(bool(zero_nan)), _lib.ptr(out),
        int(out_channels), _lib.stream_ptr()))
    return out


def analytical_log_likelihood_gradient(linear_forward_model: SignalForwardModel, R2star_map: torch.Tensor,
                                       S0_map: torch.Tensor, B0_map: torch.Tensor, phi_map: torch.Tensor, TEs: List,
                                       sensitivity_maps: torch.Tensor, masked_kspace: torch.Tensor,
                                       sampling_mask: torch.Tensor, fft_centered: bool, fft_normalization: str,
                                       spatial_dims: Sequence[int], coil_dim: int,
                                       coil_combination_method: str = "SENSE", scaling: float = 1e-3) -> torch.Tensor:
    """qrim/utils.py:166-295 for ONE sample: maps [H, W], sensitivity_maps [C, H, W, 2], masked_kspace [E, C, H, W, 2],
    sampling_mask broadcastable to [1, E, C, H, W, 2] -> [4, H, W] = (R2*_re, S0_re, R2*_im, S0_im)."""
    if coil_dim != 2:
        raise NotImplementedError("mridc_b200: analytical_log_likelihood_gradient expects coil_dim == 2")
    if coil_combination_method != "SENSE":
        raise NotImplementedError("mridc_b200: analytical_log_likelihood_gradient supports SENSE coil combination only")
    _ops.check_spatial_dims(spatial_dims)
    maps = [t.unsqueeze(0) for t in _maps(R2star_map, S0_map, B0_map, phi_map)]
    if maps[0].dim() != 3:
        raise ValueError("maps must be [n_x, n_y] (one sample)")
    sens = _lib.require_cuda(sensitivity_maps, "sensitivity_maps").contiguous().unsqueeze(0)
    y = _lib.require_cuda(masked_kspace, "masked_kspace").contiguous().unsqueeze(0)
    if y.dim() != 6 or sens.dim() != 5:
        raise ValueError("expected masked_kspace [E, C, H, W, 2] and sensitivity_maps [C, H, W, 2]")
    H, W = y.shape[3], y.shape[4]
    mask = _lib.require_cuda(sampling_mask, "sampling_mask", None)
    while mask.dim() > 3 and mask.shape[0] == 1:
        mask = mask[0]
    if mask.dim() == 3 and mask.shape[-1] in (1, 2) and mask.shape[-2] == W:  # [H|1, W, 1]
        mask = mask[..., 0]
    if mask.dim() == 2:
        mask = mask.unsqueeze(0)
    out = torch.empty((1, 4, H, W), dtype=torch.float32, device=y.device)
    _qmri_gradient(linear_forward_model, maps, None, _tes_list(TEs), sens, y, mask, fft_centered, fft_normalization,
                   scaling, 1.0, False, out, 4)
    return out[0]


class qRIMBlock(nn.Module):
    """qrim_block.py:13-240."""

    def __init__(self, recurrent_layer=None, conv_filters=None, conv_kernels=None, conv_dilations=None,
                 conv_bias=None, recurrent_filters=None, recurrent_kernels=None, recurrent_dilations=None,
                 recurrent_bias=None, depth: int = 2, time_steps: int = 8, conv_dim: int = 2, no_dc: bool = False,
                 linear_forward_model=None, fft_centered: bool = True, fft_normalization: str = "ortho",
                 spatial_dims: Optional[Tuple[int, int]] = None, coil_dim: int = 1,
                 coil_combination_method: str = "SENSE", dimensionality: int = 2):
        super().__init__()
        self.linear_forward_model = (
            SignalForwardModel(sequence="MEGRE") if linear_forward_model is None else linear_forward_model
        )
        self.input_size = depth * 4
        self.time_steps = time_steps
        self.layers = nn.ModuleList()
        conv_layer = None
        for (
            (conv_features, conv_k_size, conv_dilation, l_conv_bias, nonlinear),
            (rnn_features, rnn_k_size, rnn_dilation, rnn_bias, rnn_type),
        ) in zip(
            zip(conv_filters, conv_kernels, conv_dilations, conv_bias, ["relu", "relu", None]),
            zip(recurrent_filters, recurrent_kernels, recurrent_dilations, recurrent_bias,
                [recurrent_layer, recurrent_layer, None]),
        ):
            conv_layer = None
            if conv_features != 0:
                conv_layer = ConvNonlinear(self.input_size, conv_features, conv_dim=conv_dim, kernel_size=conv_k_size,
                                           dilation=conv_dilation, bias=l_conv_bias, nonlinear=nonlinear)
                self.input_size = conv_features
            if rnn_features != 0 and rnn_type is not None:
                if rnn_type.upper() == "GRU":
                    rnn_cls = ConvGRUCell
                elif rnn_type.upper() == "MGU":
                    rnn_cls = ConvMGUCell
                elif rnn_type.upper() == "INDRNN":
                    rnn_cls = IndRNNCell
                else:
                    raise ValueError("Please specify a proper recurrent layer type.")
                rnn_layer = rnn_cls(self.input_size, rnn_features, conv_dim=conv_dim, kernel_size=rnn_k_size,
                                    dilation=rnn_dilation, bias=rnn_bias)
                self.input_size = rnn_features
                self.layers.append(ConvRNNStack(conv_layer, rnn_layer))
        self.final_layer = nn.Sequential(conv_layer)
        self.recurrent_filters = recurrent_filters
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.spatial_dims = spatial_dims if spatial_dims is not None else [-2, -1]
        self.coil_dim = coil_dim
        self.coil_combination_method = coil_combination_method

    @torch.no_grad()
    def forward(self, pred: torch.Tensor, masked_kspace: torch.Tensor, R2star_map_init: torch.Tensor,
                S0_map_init: torch.Tensor, B0_map_init: torch.Tensor, phi_map_init: torch.Tensor, TEs: List,
                sensitivity_maps: torch.Tensor, sampling_mask: torch.Tensor, eta: torch.Tensor = None,
                hx: torch.Tensor = None, gamma: torch.Tensor = None,
                keep_eta: bool = False) -> Tuple[Any, Union[list, torch.Tensor, None]]:
        """qrim_block.py:134-240 -> (list[time_steps] of eta [B, 4, H, W], None)."""
        if self.coil_dim != 2:
            raise NotImplementedError("mridc_b200: qRIMBlock expects coil_dim == 2 ([B, E, C, H, W, 2] k-space)")
        if self.coil_combination_method != "SENSE":
            raise NotImplementedError("mridc_b200: qRIMBlock supports SENSE coil combination only")
        _ops.check_spatial_dims(self.spatial_dims)
        y = _lib.require_cuda(masked_kspace, "masked_kspace").contiguous()
        sens = _lib.require_cuda(sensitivity_maps, "sensitivity_maps").contiguous()
        if y.dim() != 6 or y.shape[-1] != 2 or sens.dim() != 5:
            raise ValueError("expected masked_kspace [B, E, C, H, W, 2] and sensitivity_maps [B, C, H, W, 2]")
        B, E, C, H, W, _ = y.shape
        maps = _maps(R2star_map_init, S0_map_init, B0_map_init, phi_map_init)
        if tuple(maps[0].shape) != (B, H, W):
            raise ValueError("maps must be [batch_size, n_x, n_y] = %s (got %s)" % ((B, H, W), tuple(maps[0].shape)))
        if eta is None:  # :184-185
            eta = torch.stack(maps, dim=1)
        eta = _lib.require_cuda(eta, "eta")
        if hx is None:  # :187-192
            hx = [y.new_zeros((B, f, H, W)) for f in self.recurrent_filters if f != 0]
        g4 = [float(gamma[k]) for k in range(4)]  # :196-199 (TypeError on gamma=None, like the reference)
        # conv input [grad_eta | eta] (:226): the gradient half is step-invariant, the eta half is updated in place
        x = torch.empty((B, 8, H, W), dtype=torch.float32, device=y.device)
        x[:, 4:].copy_(eta)
        _qmri_gradient(self.linear_forward_model, maps, g4, _tes_list(TEs), sens, y, sampling_mask, self.fft_centered,
                       self.fft_normalization, 1e-3, 100.0, True, x, 8)  # :204-223
        lib = _lib.load()
        etas = []
        final = self.final_layer[0]
        for _ in range(self.time_steps):
            g = x
            for h, convrnn in enumerate(self.layers):  # :228-230
                hx[h] = convrnn(g, hx[h])
                g = hx[h]
            delta = final(g)  # :232
            eta_t = torch.empty((B, 4, H, W), dtype=torch.float32, device=y.device)
            _lib.check(lib.mrb_qrim_eta_update(_lib.ptr(x), 8, 4, _lib.ptr(delta), _lib.ptr(eta_t), B, H * W,
                                               _lib.stream_ptr()))  # :233-236
            etas.append(eta_t)
        return etas, None

"""Drop-ins for the Recurrent Variational Network blocks:
``Conv2dGRU`` (mridc/collections/reconstruction/models/recurrentvarnet/conv2gru.py:11-163), ``RecurrentInit`` and
``RecurrentVarNetBlock`` (recurrentvarnet/recurrentvarnet.py:16-240).

Same sub-module / parameter names as the reference (its ``state_dict`` loads with ``strict=True``).  The k-space halves run
on the fused operators of the DC path (``mrb_sens_reduce``, ``mrb_sens_expand_softdc``: the refinement enters with the
opposite sign of VarNet's, so the regulariser's image is negated first); every convolution is ``mrb_conv2d`` with its
activation fused; the GRU gate arithmetic is pointwise.
"""
from typing import List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import _lib, _ops

__all__ = ["Conv2dGRU", "RecurrentInit", "RecurrentVarNetBlock"]


def _conv(seq: nn.Sequential, x: torch.Tensor, act=_ops.ACT_NONE) -> torch.Tensor:
    """Run an ``nn.Sequential([ReplicationPad2d,] Conv2d)`` parameter holder on the CUDA conv kernel."""
    conv = seq[-1]
    k, dil = int(conv.kernel_size[0]), int(conv.dilation[0])
    replicate = isinstance(seq[0], nn.ReplicationPad2d)
    want = dil * (k - 1) // 2
    have = int(seq[0].padding[0]) if replicate else int(conv.padding[0])
    if have != want or k % 2 == 0:
        raise NotImplementedError("mridc_b200: only 'same' odd-kernel convolutions (padding %d for k=%d, dilation %d; got %d)"
                                  % (want, k, dil, have))
    bias = conv.bias.detach() if conv.bias is not None else None
    return _ops.conv2d(x.contiguous(), conv.weight.detach(), bias, k, dil,
                       _ops.PAD_REPLICATE if replicate else _ops.PAD_ZERO, act=act)


class Conv2dGRU(nn.Module):
    """conv2gru.py:11-163."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: Optional[int] = None, num_layers: int = 2,
                 gru_kernel_size=1, orthogonal_initialization: bool = True, instance_norm: bool = False,
                 dense_connect: int = 0, replication_padding: bool = True):
        super().__init__()
        if out_channels is None:
            out_channels = in_channels
        if instance_norm:
            raise NotImplementedError("mridc_b200: Conv2dGRU(instance_norm=True) is not built (RecurrentVarNetBlock uses False)")
        self.num_layers = num_layers
        self.hidden_channels = hidden_channels
        self.dense_connect = dense_connect
        self.reset_gates = nn.ModuleList([])
        self.update_gates = nn.ModuleList([])
        self.out_gates = nn.ModuleList([])
        self.conv_blocks = nn.ModuleList([])
        # conv2gru.py:67-91 (layer 0: 5x5, layer 1: 3x3 dilation 2, then 3x3)
        for idx in range(num_layers + 1):
            in_ch = in_channels if idx == 0 else (1 + min(idx, dense_connect)) * hidden_channels
            out_ch = hidden_channels if idx < num_layers else out_channels
            k, dil = (5 if idx == 0 else 3), (2 if idx == 1 else 1)
            pad = 2 if idx in (0, 1) else 1
            block = [nn.ReplicationPad2d(pad)] if replication_padding else []
            block.append(nn.Conv2d(in_ch, out_ch, k, dilation=dil, padding=0 if replication_padding else (2 if idx == 0 else 1)))
            self.conv_blocks.append(nn.Sequential(*block))
        # :93-107
        for _ in range(num_layers):
            for gru_part in (self.reset_gates, self.update_gates, self.out_gates):
                gru_part.append(nn.Sequential(nn.Conv2d(2 * hidden_channels, hidden_channels, gru_kernel_size,
                                                        padding=gru_kernel_size // 2)))
        if orthogonal_initialization:  # :109-117
            for reset_gate, update_gate, out_gate in zip(self.reset_gates, self.update_gates, self.out_gates):
                nn.init.orthogonal_(reset_gate[-1].weight)
                nn.init.orthogonal_(update_gate[-1].weight)
                nn.init.orthogonal_(out_gate[-1].weight)
                nn.init.constant_(reset_gate[-1].bias, -1.0)
                nn.init.constant_(update_gate[-1].bias, 0.0)
                nn.init.constant_(out_gate[-1].bias, 0.0)

    @torch.no_grad()
    def forward(self, cell_input: torch.Tensor, previous_state: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """:119-163; states are [B, hidden, H, W, num_layers]."""
        _lib.require_cuda(cell_input, "cell_input")
        new_states: List[torch.Tensor] = []
        conv_skip: List[torch.Tensor] = []
        if previous_state is None:
            previous_state = torch.zeros(cell_input.size(0), self.hidden_channels, cell_input.size(2), cell_input.size(3),
                                         self.num_layers, dtype=cell_input.dtype, device=cell_input.device)
        for idx in range(self.num_layers):
            if conv_skip:
                cell_input = torch.cat([*conv_skip[-self.dense_connect:], cell_input], dim=1)
            cell_input = _conv(self.conv_blocks[idx], cell_input, _ops.ACT_RELU)
            if self.dense_connect > 0:
                conv_skip.append(cell_input)
            h = previous_state[:, :, :, :, idx]
            stacked = torch.cat([cell_input, h], dim=1)
            update = torch.sigmoid(_conv(self.update_gates[idx], stacked))
            reset = torch.sigmoid(_conv(self.reset_gates[idx], stacked))
            delta = torch.tanh(_conv(self.out_gates[idx], torch.cat([cell_input, h * reset], dim=1)))
            cell_input = h * (1 - update) + delta * update
            new_states.append(cell_input)
            cell_input = torch.relu(cell_input)
        if conv_skip:
            cell_input = torch.cat([*conv_skip[-self.dense_connect:], cell_input], dim=1)
        out = _conv(self.conv_blocks[self.num_layers], cell_input)
        return out, torch.stack(new_states, dim=-1)


class RecurrentInit(nn.Module):
    """recurrentvarnet.py:16-109: learned initial hidden state, [B, in, H, W] -> [B, out, H, W, depth]."""

    def __init__(self, in_channels: int, out_channels: int, channels: Tuple[int, ...], dilations: Tuple[int, ...],
                 depth: int = 2, multiscale_depth: int = 1):
        super().__init__()
        self.conv_blocks = nn.ModuleList()
        self.out_blocks = nn.ModuleList()
        self.depth = depth
        self.multiscale_depth = multiscale_depth
        tch = in_channels
        for (curr_channels, curr_dilations) in zip(channels, dilations):
            self.conv_blocks.append(nn.Sequential(nn.ReplicationPad2d(curr_dilations),
                                                  nn.Conv2d(tch, curr_channels, 3, padding=0, dilation=curr_dilations)))
            tch = curr_channels
        tch = int(np.sum(channels[-multiscale_depth:]))
        for _ in range(depth):
            self.out_blocks.append(nn.Sequential(nn.Conv2d(tch, out_channels, 1, padding=0)))

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _lib.require_cuda(x, "x")
        features = []
        for block in self.conv_blocks:
            x = _conv(block, x, _ops.ACT_RELU)
            if self.multiscale_depth > 1:
                features.append(x)
        if self.multiscale_depth > 1:
            x = torch.cat(features[-self.multiscale_depth:], dim=1)
        return torch.stack([_conv(block, x, _ops.ACT_RELU) for block in self.out_blocks], dim=-1)


class RecurrentVarNetBlock(nn.Module):
    """recurrentvarnet.py:112-240: k <- k - alpha * where(mask == 0, 0, k - y) + F(S * H_theta(sum_c conj(S) F^-1 k))."""

    def __init__(self, in_channels: int = 2, hidden_channels: int = 64, num_layers: int = 4, fft_centered: bool = True,
                 fft_normalization: str = "ortho", spatial_dims: Optional[Tuple[int, int]] = None, coil_dim: int = 1):
        super().__init__()
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.spatial_dims = spatial_dims if spatial_dims is not None else [-2, -1]
        self.coil_dim = coil_dim
        self.learning_rate = nn.Parameter(torch.tensor([1.0]))
        self.regularizer = Conv2dGRU(in_channels=in_channels, hidden_channels=hidden_channels, num_layers=num_layers,
                                     replication_padding=True)

    @torch.no_grad()
    def forward(self, current_kspace: torch.Tensor, masked_kspace: torch.Tensor, sampling_mask: torch.Tensor,
                sensitivity_map: torch.Tensor, hidden_state: Union[None, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        _ops.check_spatial_dims(self.spatial_dims)
        if self.coil_dim != 1:
            raise NotImplementedError("mridc_b200: RecurrentVarNetBlock expects coil_dim == 1")
        current_kspace = _lib.require_cuda(current_kspace, "current_kspace").contiguous()
        if current_kspace.dim() != 5 or current_kspace.shape[-1] != 2:
            raise NotImplementedError("mridc_b200: RecurrentVarNetBlock expects [B, C, H, W, 2] k-space")
        B, C, H, W, _ = current_kspace.shape
        ws = torch.empty((1, B, C, H, W, 2), dtype=torch.float32, device=current_kspace.device)
        # :205-221 (one complex image: the last dimension holds a single (re, im) pair)
        image = _ops.sens_reduce(current_kspace, sensitivity_map, self.fft_centered, self.fft_normalization, ws=ws)
        recurrent_term, hidden_state = self.regularizer(image.permute(0, 3, 1, 2), hidden_state)
        # :224-240: the fused operator computes base - where(mask, pred - y, 0) * w - F(S * img)
        refinement = -recurrent_term.permute(0, 2, 3, 1)
        new_kspace = _ops.sens_expand_softdc(refinement, sensitivity_map, current_kspace, current_kspace, masked_kspace,
                                             sampling_mask, self.learning_rate.detach(), False, self.fft_centered,
                                             self.fft_normalization, ws=ws)
        return new_kspace, hidden_state

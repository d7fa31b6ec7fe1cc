"""Drop-ins for the U-Net regulariser of the reference (unet_base/unet_block.py:11-308) on B200 kernels.

Sub-module names match the reference (``unet.down_sample_layers.N.layers.{0,4}.weight`` ...) so reference
checkpoints load key-for-key; ``torch.nn`` layers only hold parameters / initialise them.  The 3x3 convolutions run as
implicit GEMMs on tcgen05 with error-compensated fp16-split operands (conv_tc2.cu, uconv3_kernel: x = hi + lo to 2^-22,
three products per MAC, fp32 accumulation -- the bf16 split of the RIM kernels (2^-17) does not survive E2EVN's metric
gate), the first conv (caller-scaled input) and everything else on exact-fp32 CUDA-core kernels, InstanceNorm statistics in
fp64.  Skip connections are written straight into the concat buffer
(no ``torch.cat``).  Inference only (Dropout2d is the identity).
"""
import math
import os
import weakref
from typing import List, Tuple

import torch
import torch.nn as nn

from . import _lib, _ops

__all__ = ["NormUnet", "Unet", "ConvBlock", "TransposeConvBlock"]


def _tc_convs():
    """tcgen05 split-bf16 3x3 convolutions (conv_tc2.cu, uconv3_kernel); MRIDC_B200_UNET_FP32=1 keeps the exact-fp32
    CUDA-core kernel."""
    return os.environ.get("MRIDC_B200_UNET_FP32", "0") != "1"


_PACKS = {}  # id(parameter) -> (weak reference, key, packed weights); rebuilt when the parameter changes


def _packed(weight):
    """fp16 hi / lo weight image of uconv3_kernel, packed once per parameter version.  The entry holds a weak reference to
    the parameter: ids and device addresses are recycled once a model is freed, so (id, data_ptr, version) alone could hand
    a new model the packed weights of a dead one."""
    key = (weight.data_ptr(), weight._version, str(weight.device))
    hit = _PACKS.get(id(weight))
    if hit is not None and hit[0]() is weight and hit[1] == key:
        return hit[2]
    lib = _lib.load()
    Cout, Cin = weight.shape[0], weight.shape[1]
    w = weight.detach().contiguous()
    pk = torch.empty(lib.mrb_tc2_unet_packed_bytes(Cin, Cout), dtype=torch.uint8, device=weight.device)
    _lib.check(lib.mrb_tc2_unet_pack(_lib.ptr(w), _lib.ptr(pk), Cin, Cout, _lib.stream_ptr()))
    if len(_PACKS) > 1024:  # drop the entries of freed parameters
        for k in [k for k, v in _PACKS.items() if v[0]() is None]:
            del _PACKS[k]
    _PACKS[id(weight)] = (weakref.ref(weight), key, pk)
    return pk


def _conv3x3(x, weight, x_bs, N, Cin, H, W, normalised=True):
    """Conv2d(k = 3, padding = 1, bias = False) of the buffer x ([N, Cin, H, W] with batch stride x_bs) -> contiguous
    [N, Cout, H, W].  ``normalised``: the input is the output of an InstanceNorm (O(1) values), which the fp16-split
    tensor-core kernel requires; the first conv of a U-Net sees caller-scaled data and stays on the exact-fp32 kernel."""
    Cout = weight.shape[0]
    if normalised and _tc_convs() and Cin <= 64:
        out = torch.empty((N, Cout, H, W), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().mrb_tc2_unet_conv3x3(_lib.ptr(x), x_bs, _lib.ptr(_packed(weight)), _lib.ptr(out), Cout * H * W,
                                                    N, Cin, Cout, H, W, _lib.stream_ptr()))
        return out
    return _ops.conv2d(x, weight, None, 3, 1, _ops.PAD_ZERO, x_bstride=x_bs, N=N, Cin=Cin, H=H, W=W)


def _instnorm_lrelu(x, x_bs, out, out_bs, N, C, HW, slope=0.2, eps=1e-5):
    stats = torch.empty((2 * N * C,), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().mrb_instnorm_lrelu(_lib.ptr(x), x_bs, _lib.ptr(out), out_bs, N, C, HW, eps, slope,
                                              _lib.ptr(stats), _lib.stream_ptr()))
    return out


class ConvBlock(nn.Module):
    """unet_block.py:230-271: (3x3 conv, bias=False -> InstanceNorm2d -> LeakyReLU(0.2) -> Dropout2d) x 2."""

    def __init__(self, in_chans: int, out_chans: int, drop_prob: float):
        super().__init__()
        self.in_chans, self.out_chans, self.drop_prob = in_chans, out_chans, drop_prob
        self.layers = nn.Sequential(
            nn.Conv2d(in_chans, out_chans, kernel_size=3, padding=1, bias=False),
            nn.InstanceNorm2d(out_chans),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Dropout2d(drop_prob),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
            nn.InstanceNorm2d(out_chans),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            nn.Dropout2d(drop_prob),
        )

    def run(self, x, x_bs, N, H, W, out=None, out_bs=None, normalised=True):
        """x: buffer holding [N, in_chans, H, W] with batch stride x_bs.  Result goes to ``out`` (batch stride
        out_bs) if given, else to a fresh contiguous tensor.  ``normalised``: see _conv3x3."""
        if self.training and self.drop_prob > 0:
            raise NotImplementedError("mridc_b200 is inference only (Dropout2d with p > 0 in training mode)")
        C = self.out_chans
        HW = H * W
        t = _conv3x3(x, self.layers[0].weight, x_bs, N, self.in_chans, H, W, normalised)
        _instnorm_lrelu(t, C * HW, t, C * HW, N, C, HW)
        u = _conv3x3(t, self.layers[4].weight, C * HW, N, C, H, W)
        if out is None:
            out, out_bs = u, C * HW
        _instnorm_lrelu(u, C * HW, out, out_bs, N, C, HW)
        return out

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        image = _lib.require_cuda(image, "image").contiguous()
        N, C, H, W = image.shape
        return self.run(image, C * H * W, N, H, W, normalised=False)


class TransposeConvBlock(nn.Module):
    """unet_block.py:274-308: ConvTranspose2d(k=2, s=2, bias=False) -> InstanceNorm2d -> LeakyReLU(0.2)."""

    def __init__(self, in_chans: int, out_chans: int):
        super().__init__()
        self.in_chans, self.out_chans = in_chans, out_chans
        self.layers = nn.Sequential(
            nn.ConvTranspose2d(in_chans, out_chans, kernel_size=2, stride=2, bias=False),
            nn.InstanceNorm2d(out_chans),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
        )

    def run(self, x, N, H, W, out=None, out_bs=None):
        C = self.out_chans
        t = torch.empty((N, C, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().mrb_conv_transpose2x2(_lib.ptr(x), self.in_chans * H * W, _lib.ptr(self.layers[0].weight),
                                                     _lib.ptr(t), C * 4 * H * W, N, self.in_chans, C, H, W,
                                                     _lib.stream_ptr()))
        if out is None:
            out, out_bs = t, C * 4 * H * W
        _instnorm_lrelu(t, C * 4 * H * W, out, out_bs, N, C, 4 * H * W)
        return out

    def forward(self, image: torch.Tensor) -> torch.Tensor:
        image = _lib.require_cuda(image, "image").contiguous()
        N, _, H, W = image.shape
        return self.run(image, N, H, W)


class Unet(nn.Module):
    """unet_block.py:139-227."""

    def __init__(self, in_chans: int, out_chans: int, chans: int = 32, num_pool_layers: int = 4,
                 drop_prob: float = 0.0):
        super().__init__()
        self.in_chans, self.out_chans, self.chans = in_chans, out_chans, chans
        self.num_pool_layers, self.drop_prob = num_pool_layers, drop_prob
        self.down_sample_layers = nn.ModuleList([ConvBlock(in_chans, chans, drop_prob)])
        ch = chans
        for _ in range(num_pool_layers - 1):
            self.down_sample_layers.append(ConvBlock(ch, ch * 2, drop_prob))
            ch *= 2
        self.conv = ConvBlock(ch, ch * 2, drop_prob)
        self.up_conv = nn.ModuleList()
        self.up_transpose_conv = nn.ModuleList()
        for _ in range(num_pool_layers - 1):
            self.up_transpose_conv.append(TransposeConvBlock(ch * 2, ch))
            self.up_conv.append(ConvBlock(ch * 2, ch, drop_prob))
            ch //= 2
        self.up_transpose_conv.append(TransposeConvBlock(ch * 2, ch))
        self.up_conv.append(nn.Sequential(ConvBlock(ch * 2, ch, drop_prob),
                                          nn.Conv2d(ch, self.out_chans, kernel_size=1, stride=1)))

    @torch.no_grad()
    def forward(self, image: torch.Tensor, normalised_input: bool = False) -> torch.Tensor:
        """unet_block.py:187-227.  ``normalised_input`` (not in the reference signature): the caller guarantees O(1) input
        values (NormUnet with normalize=True), so the first convolution may use the fp16-split tensor-core kernel too."""
        lib = _lib.load()
        image = _lib.require_cuda(image, "image").contiguous()
        N, C, H, W = image.shape
        dev = image.device
        cats = []  # (concat buffer [N, 2*ch, h, w], ch, h, w): skip lives in channels [ch, 2ch)
        cur, cur_bs, cur_c, h, w = image, C * H * W, C, H, W
        for li, layer in enumerate(self.down_sample_layers):
            ch = layer.out_chans
            cat = torch.empty((N, 2 * ch, h, w), dtype=torch.float32, device=dev)
            skip = cat[:, ch:]  # view: batch stride 2*ch*h*w
            # the very first conv reads the caller's image (any scale); every later input is an average of
            # instance-normalised activations
            layer.run(cur, cur_bs, N, h, w, out=skip, out_bs=2 * ch * h * w, normalised=li > 0 or normalised_input)
            cats.append((cat, ch, h, w))
            pooled = torch.empty((N, ch, h // 2, w // 2), dtype=torch.float32, device=dev)
            _lib.check(lib.mrb_avgpool2(_lib.ptr(skip), 2 * ch * h * w, _lib.ptr(pooled), ch * (h // 2) * (w // 2), N,
                                        ch, h, w, _lib.stream_ptr()))
            cur, cur_bs, cur_c, h, w = pooled, ch * (h // 2) * (w // 2), ch, h // 2, w // 2
        out = self.conv.run(cur, cur_bs, N, h, w)
        for i, (tconv, conv) in enumerate(zip(self.up_transpose_conv, self.up_conv)):
            cat, ch, sh, sw = cats.pop()
            if (2 * h, 2 * w) == (sh, sw):
                tconv.run(out, N, h, w, out=cat, out_bs=2 * ch * sh * sw)
            else:
                # odd skip size: reflect-pad right/bottom by one (unet_block.py:216-222)
                up = tconv.run(out, N, h, w)
                _lib.check(lib.mrb_pad2d(_lib.ptr(up), ch * 4 * h * w, _lib.ptr(cat), 2 * ch * sh * sw, N, ch, 2 * h,
                                         2 * w, sh, sw, 0, 0, 2, _lib.stream_ptr()))
            h, w = sh, sw
            block = conv if isinstance(conv, ConvBlock) else conv[0]
            out = block.run(cat, 2 * ch * h * w, N, h, w)
            if not isinstance(conv, ConvBlock):
                fin = conv[1]  # unet_block.py:185: 1x1 conv with bias
                if (h * w) % 4 == 0:
                    res = torch.empty((N, fin.out_channels, h, w), dtype=torch.float32, device=dev)
                    _lib.check(lib.mrb_conv1x1(_lib.ptr(out), ch * h * w, _lib.ptr(fin.weight), _lib.ptr(fin.bias), _lib.ptr(res),
                                               fin.out_channels * h * w, N, ch, fin.out_channels, h * w, _lib.stream_ptr()))
                    out = res
                else:
                    out = _ops.conv2d(out, fin.weight, fin.bias, 1, 1, _ops.PAD_ZERO)
        return out


class NormUnet(nn.Module):
    """unet_block.py:11-136: complex -> channels, group-norm (2 groups, unbiased std), pad, U-Net, unpad, unnorm."""

    def __init__(self, chans: int, num_pools: int, in_chans: int = 2, out_chans: int = 2, drop_prob: float = 0.0,
                 padding_size: int = 15, normalize: bool = True, norm_groups: int = 2):
        super().__init__()
        self.unet = Unet(in_chans=in_chans, out_chans=out_chans, chans=chans, num_pool_layers=num_pools,
                         drop_prob=drop_prob)
        self.padding_size = padding_size
        self.normalize = normalize
        self.norm_groups = norm_groups

    def pad_sizes(self, h, w):
        w_mult = ((w - 1) | self.padding_size) + 1
        h_mult = ((h - 1) | self.padding_size) + 1
        w_pad = [math.floor((w_mult - w) / 2), math.ceil((w_mult - w) / 2)]
        h_pad = [math.floor((h_mult - h) / 2), math.ceil((h_mult - h) / 2)]
        return h_pad, w_pad, h_mult, w_mult

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        x = _lib.require_cuda(x, "x").contiguous()
        if x.shape[-1] != 2 or x.dim() != 5:
            raise NotImplementedError("mridc_b200: NormUnet expects complex input [B, C, H, W, 2]")
        if self.norm_groups != 2:
            raise NotImplementedError("mridc_b200: NormUnet supports norm_groups == 2")
        B, C, H, W, _ = x.shape
        dev = x.device
        st = _lib.stream_ptr()
        planar = torch.empty((B, 2 * C, H, W), dtype=torch.float32, device=dev)
        mean_std = torch.empty((B, 2, 2), dtype=torch.float32, device=dev)
        stats = torch.empty((4 * B,), dtype=torch.float64, device=dev)
        _lib.check(lib.mrb_normunet_in(_lib.ptr(x), _lib.ptr(planar), _lib.ptr(mean_std), B, C, H * W,
                                       int(bool(self.normalize)), _lib.ptr(stats), st))
        h_pad, w_pad, h_mult, w_mult = self.pad_sizes(H, W)
        if (h_mult, w_mult) != (H, W):
            padded = torch.empty((B, 2 * C, h_mult, w_mult), dtype=torch.float32, device=dev)
            _lib.check(lib.mrb_pad2d(_lib.ptr(planar), 2 * C * H * W, _lib.ptr(padded), 2 * C * h_mult * w_mult, B,
                                     2 * C, H, W, h_mult, w_mult, h_pad[0], w_pad[0], 0, st))
        else:
            padded = planar
        y = self.unet(padded, normalised_input=bool(self.normalize))  # group-normalised: zero mean, unit std
        Co = y.shape[1]
        if (h_mult, w_mult) != (H, W):
            un = torch.empty((B, Co, H, W), dtype=torch.float32, device=dev)
            _lib.check(lib.mrb_pad2d(_lib.ptr(y), Co * h_mult * w_mult, _lib.ptr(un), Co * H * W, B, Co, h_mult, w_mult,
                                     H, W, -h_pad[0], -w_pad[0], 0, st))
            y = un
        out = torch.empty((B, Co // 2, H, W, 2), dtype=torch.float32, device=dev)
        _lib.check(lib.mrb_normunet_out(_lib.ptr(y), _lib.ptr(mean_std), _lib.ptr(out), B, Co // 2, H * W,
                                        int(bool(self.normalize)), st))
        return out

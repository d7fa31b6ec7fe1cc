"""Oracle (TEST INFRASTRUCTURE): RIM / VarNet / U-Net blocks restated functionally on PyTorch-CPU.

Weights come as a flat ``state_dict``-style mapping with the reference's own key names (SURVEY.md
section 8a "Weights"); ``prefix`` selects a sub-module.  Hyper-parameters come as a plain dict with the
reference cfg keys.  Not imported by the product package.
"""
import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import mri


# --------------------------------------------------------------------------------------------------
# RIM
# --------------------------------------------------------------------------------------------------
def log_likelihood_gradient(eta, masked_kspace, sense, mask, sigma, fft_centered, fft_normalization,
                            spatial_dims, coil_dim):
    """reconstruction/models/rim/rim_utils.py:11-67."""
    if coil_dim == 0:
        coil_dim += 1  # :41-42
    er = eta[..., 0:1].unsqueeze(coil_dim)
    ei = eta[..., 1:2].unsqueeze(coil_dim)
    sr, si = sense[..., 0:1], sense[..., 1:2]
    pred = torch.cat((er * sr - ei * si, er * si + ei * sr), -1)  # :47-49
    pred = mri.fft2(pred, fft_centered, fft_normalization, spatial_dims)  # :51
    pred = mri.ifft2(mask * (pred - masked_kspace), fft_centered, fft_normalization, spatial_dims)  # :53-58
    pr, pi = pred[..., 0:1], pred[..., 1:2]
    re_out = torch.sum(pr * sr + pi * si, coil_dim) / (sigma**2.0)  # :61
    im_out = torch.sum(pi * sr - pr * si, coil_dim) / (sigma**2.0)  # :62
    return torch.cat((er.squeeze(coil_dim), ei.squeeze(coil_dim), re_out, im_out), -1).permute(0, 3, 1, 2)  # :67


def conv_nonlinear(x, weight, bias, kernel_size, dilation, nonlinear):
    """rim/conv_layers.py:36-123: ReplicationPad2d(dil*(k-1)//2) -> Conv2d(padding=0) -> nonlinearity."""
    pad = (dilation * (kernel_size - 1)) // 2
    if pad > 0:
        x = F.pad(x, (pad, pad, pad, pad), mode="replicate")
    x = F.conv2d(x, weight, bias, padding=0, dilation=dilation)
    if nonlinear is None:
        return x
    if nonlinear.upper() == "RELU":
        return F.relu(x)
    if nonlinear.upper() == "LEAKYRELU":
        return F.leaky_relu(x, 0.01)
    raise ValueError("Please specify a proper nonlinearity")


def _zero_pad_conv(x, w, b, k, dil):
    return F.conv2d(x, w, b, padding=(dil * (k - 1)) // 2, dilation=dil)


def conv_gru_cell(x, h, w_ih, b_ih, w_hh, kernel_size, dilation):
    """rim/rnn_cells.py:93-127 (conv_dim == 2 path)."""
    ih = _zero_pad_conv(x, w_ih, b_ih, kernel_size, dilation).chunk(3, 1)
    hh = _zero_pad_conv(h, w_hh, None, kernel_size, dilation).chunk(3, 1)
    r = torch.sigmoid(ih[0] + hh[0])
    z = torch.sigmoid(ih[1] + hh[1])
    n = torch.tanh(ih[2] + r * hh[2])
    return n * (1 - z) + z * h


def conv_mgu_cell(x, h, w_ih, b_ih, w_hh, kernel_size, dilation):
    """rim/rnn_cells.py:230-261."""
    ih = _zero_pad_conv(x, w_ih, b_ih, kernel_size, dilation).chunk(2, 1)
    hh = _zero_pad_conv(h, w_hh, None, kernel_size, dilation).chunk(2, 1)
    f = torch.sigmoid(ih[0] + hh[0])
    c = torch.tanh(ih[1] + f * hh[1])
    return c + f * (h - c)


def indrnn_cell(x, h, w_ih, b_ih, hh, kernel_size, dilation):
    """rim/rnn_cells.py:367-391."""
    return F.relu(_zero_pad_conv(x, w_ih, b_ih, kernel_size, dilation) + hh * h)


def _sub(sd: Dict[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def rim_block(sd, hp, pred, masked_kspace, sense, mask, eta=None, hx=None, sigma=1.0, keep_eta=False):
    """rim/rim_block.py:139-269, dimensionality == 2, consecutive_slices == 1.

    ``hp`` keys: recurrent_layer, conv_filters, conv_kernels, conv_dilations, recurrent_filters,
    recurrent_kernels, recurrent_dilations, time_steps, no_dc, fft_centered, fft_normalization,
    spatial_dims, coil_dim.  Returns (list of etas | list of k-spaces, hx).
    """
    cen, nrm = hp["fft_centered"], hp["fft_normalization"]
    sdims, cdim = hp.get("spatial_dims") or [-2, -1], hp["coil_dim"]
    if isinstance(pred, list):
        pred = pred[-1].detach()  # :185-186
    if hx is None:  # :188-193
        hx = [masked_kspace.new_zeros((masked_kspace.size(0), f, *masked_kspace.size()[2:-1]))
              for f in hp["recurrent_filters"] if f != 0]
    if eta is None or eta.ndim < 3:  # :195-211
        eta = pred if keep_eta else torch.sum(
            mri.complex_mul(mri.ifft2(pred, cen, nrm, sdims), mri.complex_conj(sense)), cdim)
    rl = hp["recurrent_layer"].upper()
    nlayers = sum(1 for f in hp["recurrent_filters"] if f != 0)
    etas = []
    for _ in range(hp["time_steps"]):  # :217
        g = log_likelihood_gradient(eta, masked_kspace, sense, mask, sigma, cen, nrm, sdims, cdim).contiguous()
        for l in range(nlayers):  # :233-237
            p = "layers.%d." % l
            g = conv_nonlinear(g, sd[p + "convs.conv_layer.weight"], sd.get(p + "convs.conv_layer.bias"),
                               hp["conv_kernels"][l], hp["conv_dilations"][l], "relu")
            k, d = hp["recurrent_kernels"][l], hp["recurrent_dilations"][l]
            if rl == "GRU":
                hx[l] = conv_gru_cell(g, hx[l], sd[p + "rnn.ih.weight"], sd.get(p + "rnn.ih.bias"),
                                      sd[p + "rnn.hh.weight"], k, d)
            elif rl == "MGU":
                hx[l] = conv_mgu_cell(g, hx[l], sd[p + "rnn.ih.weight"], sd.get(p + "rnn.ih.bias"),
                                      sd[p + "rnn.hh.weight"], k, d)
            elif rl == "INDRNN":
                hx[l] = indrnn_cell(g, hx[l], sd[p + "rnn.ih.weight"], sd.get(p + "rnn.ih.bias"),
                                    sd[p + "rnn.hh"], k, d)
            else:
                raise ValueError("Please specify a proper recurrent layer type.")
            g = hx[l]
        L = nlayers  # final layer = the conv of the last zip entry (:122)
        g = conv_nonlinear(g, sd["final_layer.0.conv_layer.weight"], sd.get("final_layer.0.conv_layer.bias"),
                           hp["conv_kernels"][L], hp["conv_dilations"][L], None)  # :239
        eta = eta + g.permute(0, 2, 3, 1)  # :241-248
        etas.append(eta)
    if hp["no_dc"]:
        return etas, hx  # :253-254
    zero = torch.zeros(1, 1, 1, 1, 1).to(masked_kspace)
    soft_dc = torch.where(mask, pred - masked_kspace, zero) * sd["dc_weight"]  # :256 (bool mask required)
    ks = [masked_kspace - soft_dc - mri.fft2(mri.complex_mul(e.unsqueeze(cdim), sense), cen, nrm, sdims)
          for e in etas]  # :257-267
    return ks, hx


def rim_block_3d(sd, hp, pred, masked_kspace, sense, mask, eta=None, hx=None, sigma=1.0, keep_eta=False):
    """rim/rim_block.py:168-254 with dimensionality == 3 and conv_dim == 3 (IndRNN cell, no_dc).

    Inputs [batch, slices, coils, H, W, 2] are folded to [batch*slices, ...] (:168-180); in the time loop the gradient
    [batch*slices, 4, H, W] becomes the UNBATCHED Conv3d input [4, D = batch*slices, H, W] (:230-231), ConvNonlinear pads
    with ReplicationPad3d (conv_layers.py:72-76), the IndRNN cell runs its Conv3d on [1, C, D, H, W] with a
    (1, C, 1, 1, 1) recurrent weight (rnn_cells.py:297-312, :386-391) and the block permutes the hidden states back to
    [D, C, H, W] (:243-246).  Note that D mixes batch and slices: the 3-D convolutions see one stack of batch*slices.
    The GRU / MGU cells cannot run in this mode in the reference: their gates are nn.Conv2d whatever conv_dim says
    (rnn_cells.py:23-38, :160-175) and the 5-D tensors of :114-116 raise -- restated as the same RuntimeError.
    """
    cen, nrm = hp["fft_centered"], hp["fft_normalization"]
    sdims, cdim = hp.get("spatial_dims") or [-2, -1], hp["coil_dim"]
    batch, slices = masked_kspace.shape[0], masked_kspace.shape[1]
    fold = lambda t: t.reshape([t.shape[0] * t.shape[1], *t.shape[2:]])
    pred = pred[-1].detach() if isinstance(pred, (tuple, list)) else fold(pred)
    masked_kspace, mask, sense = fold(masked_kspace), fold(mask), fold(sense)
    if hx is None:
        hx = [masked_kspace.new_zeros((masked_kspace.size(0), f, *masked_kspace.size()[2:-1]))
              for f in hp["recurrent_filters"] if f != 0]
    if eta is None or eta.ndim < 3:
        eta = pred if keep_eta else torch.sum(
            mri.complex_mul(mri.ifft2(pred, cen, nrm, sdims), mri.complex_conj(sense)), cdim)
    if eta.dim() == 5:
        eta = fold(eta)
    rl = hp["recurrent_layer"].upper()
    nlayers = sum(1 for f in hp["recurrent_filters"] if f != 0)

    def conv3(x, w, b, k, dil, nonlinear):  # x [C, D, H, W]
        pad = (dil * (k - 1)) // 2
        if pad > 0:
            x = F.pad(x.unsqueeze(0), (pad,) * 6, mode="replicate").squeeze(0)
        x = F.conv3d(x.unsqueeze(0), w, b, padding=0, dilation=dil).squeeze(0)
        return F.relu(x) if nonlinear == "relu" else x

    etas = []
    for _ in range(hp["time_steps"]):
        g = log_likelihood_gradient(eta, masked_kspace, sense, mask, sigma, cen, nrm, sdims, cdim).contiguous()
        g = g.view([batch * slices, 4, g.shape[2], g.shape[3]]).permute(1, 0, 2, 3)
        for l in range(nlayers):
            p = "layers.%d." % l
            g = conv3(g, sd[p + "convs.conv_layer.weight"], sd.get(p + "convs.conv_layer.bias"),
                      hp["conv_kernels"][l], hp["conv_dilations"][l], "relu")
            k, d = hp["recurrent_kernels"][l], hp["recurrent_dilations"][l]
            x5, h5 = g.unsqueeze(0), hx[l].permute(1, 0, 2, 3).unsqueeze(0)
            if rl != "INDRNN":
                raise RuntimeError("Expected 3D (unbatched) or 4D (batched) input to conv2d, but got input of size: %s"
                                   % list(x5.shape))
            h5 = F.relu(F.conv3d(x5, sd[p + "rnn.ih.weight"], sd.get(p + "rnn.ih.bias"), padding=(d * (k - 1)) // 2,
                                 dilation=d) + sd[p + "rnn.hh"] * h5)
            hx[l] = h5.squeeze(0)
            g = hx[l]
        L = nlayers
        g = conv3(g, sd["final_layer.0.conv_layer.weight"], sd.get("final_layer.0.conv_layer.bias"),
                  hp["conv_kernels"][L], hp["conv_dilations"][L], None)
        g = g.permute(1, 2, 3, 0)
        for l in range(len(hx)):
            hx[l] = hx[l].permute(1, 0, 2, 3)
        eta = eta + g
        etas.append(eta)
    if not hp["no_dc"]:
        raise NotImplementedError("rim_block_3d restates the no_dc path")
    return etas, hx


# --------------------------------------------------------------------------------------------------
# U-Net regulariser
# --------------------------------------------------------------------------------------------------
def _conv_block(x, sd, p):
    """unet_base/unet_block.py:230-271 (Dropout2d(p=0) is the identity in eval)."""
    for i in (0, 4):
        x = F.conv2d(x, sd[p + "layers.%d.weight" % i], None, padding=1)
        x = F.instance_norm(x, eps=1e-5)
        x = F.leaky_relu(x, 0.2)
    return x


def _tconv_block(x, sd, p):
    """unet_block.py:274-308."""
    x = F.conv_transpose2d(x, sd[p + "layers.0.weight"], None, stride=2)
    return F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)


def unet(x, sd, num_pool_layers):
    """unet_block.py:139-227."""
    stack = []
    out = x
    for i in range(num_pool_layers):
        out = _conv_block(out, sd, "down_sample_layers.%d." % i)
        stack.append(out)
        out = F.avg_pool2d(out, kernel_size=2, stride=2, padding=0)
    out = _conv_block(out, sd, "conv.")
    for i in range(num_pool_layers):
        skip = stack.pop()
        out = _tconv_block(out, sd, "up_transpose_conv.%d." % i)
        padding = [0, 0, 0, 0]
        if out.shape[-1] != skip.shape[-1]:
            padding[1] = 1
        if out.shape[-2] != skip.shape[-2]:
            padding[3] = 1
        if sum(padding) != 0:
            out = F.pad(out, padding, "reflect")
        out = torch.cat([out, skip], dim=1)
        if i < num_pool_layers - 1:
            out = _conv_block(out, sd, "up_conv.%d." % i)
        else:
            out = _conv_block(out, sd, "up_conv.%d.0." % i)
            out = F.conv2d(out, sd["up_conv.%d.1.weight" % i], sd["up_conv.%d.1.bias" % i])
    return out


def norm_unet(x, sd, num_pools, padding_size=15, normalize=True, norm_groups=2):
    """unet_block.py:11-136.  ``sd`` keys are prefixed ``unet.``."""
    iscomplex = x.shape[-1] == 2
    if iscomplex:
        b, c, h, w, _ = x.shape
        x = x.permute(0, 4, 1, 2, 3).reshape(b, 2 * c, h, w)  # :55-60
    mean = std = 1.0
    if normalize:  # :71-85 (unbiased std)
        b, c, h, w = x.shape
        xg = x.reshape(b, norm_groups, -1)
        mean = xg.mean(-1, keepdim=True)
        std = xg.std(-1, keepdim=True)
        x = ((xg - mean) / std).reshape(b, c, h, w)
    _, _, h, w = x.shape  # :93-106
    w_mult = ((w - 1) | padding_size) + 1
    h_mult = ((h - 1) | padding_size) + 1
    w_pad = [math.floor((w_mult - w) / 2), math.ceil((w_mult - w) / 2)]
    h_pad = [math.floor((h_mult - h) / 2), math.ceil((h_mult - h) / 2)]
    x = F.pad(x, w_pad + h_pad)
    x = unet(x, _sub(sd, "unet."), num_pools)
    x = x[..., h_pad[0] : h_mult - h_pad[1], w_pad[0] : w_mult - w_pad[1]]  # :108-111
    if normalize:
        b, c, h, w = x.shape
        x = (x.reshape(b, norm_groups, -1) * std + mean).reshape(b, c, h, w)
    if iscomplex:
        b, c2, h, w = x.shape
        x = x.view(b, 2, c2 // 2, h, w).permute(0, 2, 3, 4, 1).contiguous()  # :62-69
    return x


# --------------------------------------------------------------------------------------------------
# VarNet block
# --------------------------------------------------------------------------------------------------
def sens_expand(x, sens, cen, nrm, sdims):
    """varnet/vn_block.py:51-69."""
    return mri.fft2(mri.complex_mul(x, sens), cen, nrm, sdims)


def sens_reduce(x, sens, cen, nrm, sdims, coil_dim):
    """varnet/vn_block.py:71-87."""
    x = mri.ifft2(x, cen, nrm, sdims)
    return mri.complex_mul(x, mri.complex_conj(sens)).sum(dim=coil_dim, keepdim=True)


def varnet_block(sd, hp, pred, ref_kspace, sens, mask):
    """varnet/vn_block.py:89-119.  ``sd``: dc_weight + model.unet.* ; hp: pooling_layers, padding_size,
    normalize, no_dc, fft_*, spatial_dims, coil_dim."""
    cen, nrm = hp["fft_centered"], hp["fft_normalization"]
    sdims, cdim = hp.get("spatial_dims") or [-2, -1], hp["coil_dim"]
    zero = torch.zeros(1, 1, 1, 1, 1).to(pred)
    soft_dc = torch.where(mask.bool(), pred - ref_kspace, zero) * sd["dc_weight"]
    eta = sens_reduce(pred, sens, cen, nrm, sdims, cdim)
    eta = norm_unet(eta, _sub(sd, "model."), hp["pooling_layers"], hp["padding_size"], hp["normalize"])
    eta = sens_expand(eta, sens, cen, nrm, sdims)
    if not hp["no_dc"]:
        eta = pred - soft_dc - eta
    return eta


# --------------------------------------------------------------------------------------------------
# Sensitivity-estimation network
# --------------------------------------------------------------------------------------------------
def sens_pad_and_num_low_freqs(mask, num_low_frequencies=None):
    """reconstruction/models/base.py:842-878."""
    if num_low_frequencies is None or num_low_frequencies == 0:
        sq = mask[:, 0, 0, :, 0].to(torch.int8)
        cent = sq.shape[1] // 2
        left = torch.argmin(sq[:, :cent].flip(1), dim=1)
        right = torch.argmin(sq[:, cent:], dim=1)
        nlf = torch.max(2 * torch.min(left, right), torch.ones_like(left))
    else:
        nlf = num_low_frequencies * torch.ones(mask.shape[0], dtype=mask.dtype, device=mask.device)
    pad = torch.div(mask.shape[-2] - nlf + 1, 2, rounding_mode="trunc")
    return pad, nlf


def sensitivity_model(sd, hp, masked_kspace, mask, num_low_frequencies=None):
    """BaseSensitivityModel.forward, reconstruction/models/base.py:880-932.  ``sd`` keys are prefixed ``norm_unet.``;
    hp keys: sens_pools, padding_size (15), sens_mask_type, sens_normalize, sens_mask_center, fft_centered,
    fft_normalization, spatial_dims, coil_dim."""
    x = masked_kspace
    if hp.get("sens_mask_center", True):
        pad, nlf = sens_pad_and_num_low_freqs(mask, num_low_frequencies)
        keep = torch.zeros_like(x)  # utils.batched_mask_center, common/parts/utils.py:379-417
        if pad.shape[0] == 1:
            a, b = int(pad), int(pad + nlf)
            if hp["sens_mask_type"] == "1D":
                keep[:, :, :, a:b] = x[:, :, :, a:b]
            elif hp["sens_mask_type"] == "2D":
                keep[:, :, a:b] = x[:, :, a:b]
        else:
            for i, (a, b) in enumerate(zip(pad, pad + nlf)):
                keep[i, :, :, a:b] = x[i, :, :, a:b]
        x = keep
    img = mri.ifft2(x, hp["fft_centered"], hp["fft_normalization"], hp.get("spatial_dims") or [-2, -1])
    b, c, h, w, comp = img.shape
    out = norm_unet(img.reshape(b * c, 1, h, w, comp), _sub(sd, "norm_unet."), hp["sens_pools"],
                    hp.get("padding_size", 15), hp.get("sens_normalize", True)).view(b, c, h, w, comp)
    if hp.get("sens_normalize", True):
        cd = hp["coil_dim"]
        out = out / mri.rss_complex(out, dim=cd).unsqueeze(-1).unsqueeze(cd)  # :826-840
    return out

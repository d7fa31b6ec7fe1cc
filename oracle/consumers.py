"""Oracle (TEST INFRASTRUCTURE): the other consumers of the DC operator restated
functionally on PyTorch-CPU.

  data_gd / data_vs / dc_layer / prox_cg   mridc/collections/reconstruction/models/sigmanet/dc_layers.py:22-478
  cascadenet_block                         mridc/collections/reconstruction/models/cascadenet/ccnn_block.py:104-139
  conv2dgru / recurrent_init /
  recurrentvarnet_block                    .../recurrentvarnet/conv2gru.py:119-163, recurrentvarnet.py:89-109,:176-240
  qvarnet_block                            mridc/collections/quantitative/models/qvarnet/qvn_block.py:103-160
  jrscirim_block                           mridc/collections/segmentation/models/jrscirim_base/jrscirim_block.py:200-377

Pinned by oracle/make_golden.py::gen_consumers against the unmodified reference modules (tests/golden/consumers.npz).
Not imported by the product package.
"""
from typing import Callable, Dict, List

import torch
import torch.nn.functional as F

from . import mri
from .nets import conv_nonlinear, norm_unet, rim_block, unet
from .qnets import megre_signal


def _A(x, smaps, mask, cen, nrm, sd, apply_mask=True):
    """sum over dim -4 (keepdim) of [mask *] fft2(S x): dc_layers.py:69-81 / :222-231 / :374-383."""
    k = mri.fft2(mri.complex_mul(x.expand_as(smaps), smaps), cen, nrm, sd)
    if apply_mask:
        k = k * mask
    return torch.sum(k, -4, keepdim=True)


def _AT(k, smaps, mask, cen, nrm, sd, apply_mask=True):
    """sum over dim -5 of conj(S) ifft2([mask *] k): dc_layers.py:83-95 / :233-243 / :385-396."""
    if apply_mask:
        k = k * mask
    return torch.sum(mri.complex_mul(mri.ifft2(k, cen, nrm, sd), mri.complex_conj(smaps)), dim=-5)


def data_gd(x, y, smaps, mask, weight, cen, nrm, sd):
    """DataGDLayer.forward, dc_layers.py:54-96."""
    r = _A(x.unsqueeze(-5), smaps, mask, cen, nrm, sd) - y
    return x - weight * _AT(r, smaps, mask, cen, nrm, sd)


def data_vs(x, y, smaps, mask, alpha, beta, cen, nrm, sd):
    """DataVSLayer.forward, dc_layers.py:359-397."""
    A_x = _A(x.unsqueeze(-5), smaps, mask, cen, nrm, sd, apply_mask=False)
    k_dc = (1 - mask) * A_x + mask * (alpha * A_x + (1 - alpha) * y)
    x_dc = _AT(k_dc, smaps, mask, cen, nrm, sd, apply_mask=False)
    return beta * x + (1 - beta) * x_dc


def dc_layer(x, y, mask, lam, cen, nrm, sd):
    """DCLayer.forward, dc_layers.py:448-467."""
    A_x = mri.fft2(x, cen, nrm, sd)
    return mri.ifft2((1 - mask) * A_x + mask * (lam * A_x + (1 - lam) * y), cen, nrm, sd)


def _cdot(a, b):
    """ConjugateGradient.complexDot, dc_layers.py:161-166."""
    n = a.shape[0]
    m = mri.complex_mul(a, mri.complex_conj(b))
    return torch.stack([m[..., 0].reshape(n, -1).sum(-1), m[..., 1].reshape(n, -1).sum(-1)], -1)


def prox_cg(z, lam, y, smaps, mask, tol, max_iter, cen, nrm, sd):
    """ConjugateGradient.forward + solve, dc_layers.py:168-255 ((re, im) pair arithmetic as upstream)."""

    def M(p):
        return lam * _AT(_A(p, smaps, mask, cen, nrm, sd), smaps, mask, cen, nrm, sd) + p

    x0 = lam * _AT(y, smaps, mask, cen, nrm, sd) + z
    n = x0.shape[0]
    x = torch.zeros(x0.shape)
    r, p = x0.clone(), x0.clone()
    x0x0 = x0.pow(2).view(n, -1).sum(-1)
    rr = torch.stack([r.pow(2).view(n, -1).sum(-1), torch.zeros(n)], dim=-1)
    it = 0
    while torch.min(rr[..., 0] / x0x0) > tol and it < max_iter:
        it += 1
        q = M(p)
        d2 = _cdot(p, q)
        re1, im1 = rr.unbind(-1)
        re2, im2 = d2.unbind(-1)
        alpha = torch.stack([re1 * re2 + im1 * im2, im1 * re2 - re1 * im2], -1) / mri.complex_abs(d2) ** 2
        x = x + mri.complex_mul(alpha.reshape(n, 1, 1, 1, -1), p)
        r = r - mri.complex_mul(alpha.reshape(n, 1, 1, 1, -1), q)
        rr_new = torch.stack([r.pow(2).view(n, -1).sum(-1), torch.zeros(n)], dim=-1)
        beta = torch.stack([rr_new[..., 0] / rr[..., 0], torch.zeros(n)], dim=-1)
        p = r + mri.complex_mul(beta.reshape(n, 1, 1, 1, -1), p)
        rr = rr_new
    return x


def cascadenet_block(model: Callable, dc_weight, pred, ref_kspace, sens, mask, cen, nrm, sd, coil_dim=1, no_dc=False):
    """CascadeNetBlock.forward, ccnn_block.py:104-139; ``model`` maps [B, 2, H, W] -> [B, 2, H, W]."""
    soft_dc = torch.where(mask.bool(), pred - ref_kspace, torch.zeros(1, 1, 1, 1, 1).to(pred)) * dc_weight
    eta = mri.complex_mul(mri.ifft2(pred, cen, nrm, sd), mri.complex_conj(sens)).sum(dim=coil_dim, keepdim=True)
    eta = model(eta.squeeze(coil_dim).permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    if eta.dim() < sens.dim():
        eta = eta.unsqueeze(1)
    eta = mri.fft2(mri.complex_mul(eta, sens), cen, nrm, sd)
    return eta if no_dc else pred - soft_dc - eta


def _seq_conv(x, sd, prefix, replicate, k, dil):
    """nn.Sequential([ReplicationPad2d(p),] Conv2d): the conv sits at index 1 behind a pad module, else 0."""
    p = dil * (k - 1) // 2
    if replicate:
        return F.conv2d(F.pad(x, (p, p, p, p), mode="replicate"), sd[prefix + "1.weight"], sd[prefix + "1.bias"], dilation=dil)
    return F.conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], padding=p, dilation=dil)


def conv2dgru(sd: Dict[str, torch.Tensor], num_layers, hidden, x, state, dense_connect=0, gru_kernel=1):
    """Conv2dGRU.forward (replication padding, no instance norm), conv2gru.py:119-163."""
    if state is None:
        state = torch.zeros(x.size(0), hidden, x.size(2), x.size(3), num_layers, dtype=x.dtype)
    new_states: List[torch.Tensor] = []
    skip: List[torch.Tensor] = []
    geo = lambda i: (5 if i == 0 else 3, 2 if i == 1 else 1)
    for i in range(num_layers):
        if skip:
            x = torch.cat([*skip[-dense_connect:], x], dim=1)
        x = F.relu(_seq_conv(x, sd, "conv_blocks.%d." % i, True, *geo(i)))
        if dense_connect > 0:
            skip.append(x)
        h = state[..., i]
        st = torch.cat([x, h], dim=1)
        u = torch.sigmoid(_seq_conv(st, sd, "update_gates.%d." % i, False, gru_kernel, 1))
        r = torch.sigmoid(_seq_conv(st, sd, "reset_gates.%d." % i, False, gru_kernel, 1))
        d = torch.tanh(_seq_conv(torch.cat([x, h * r], dim=1), sd, "out_gates.%d." % i, False, gru_kernel, 1))
        x = h * (1 - u) + d * u
        new_states.append(x)
        x = F.relu(x)
    if skip:
        x = torch.cat([*skip[-dense_connect:], x], dim=1)
    out = _seq_conv(x, sd, "conv_blocks.%d." % num_layers, True, *geo(num_layers))
    return out, torch.stack(new_states, dim=-1)


def recurrent_init(sd, dilations, depth, multiscale_depth, x):
    """RecurrentInit.forward, recurrentvarnet.py:89-109."""
    feats = []
    for i, d in enumerate(dilations):
        x = F.relu(_seq_conv(x, sd, "conv_blocks.%d." % i, True, 3, d))
        if multiscale_depth > 1:
            feats.append(x)
    if multiscale_depth > 1:
        x = torch.cat(feats[-multiscale_depth:], dim=1)
    return torch.stack([F.relu(F.conv2d(x, sd["out_blocks.%d.0.weight" % j], sd["out_blocks.%d.0.bias" % j]))
                        for j in range(depth)], dim=-1)


def recurrentvarnet_block(sd, num_layers, hidden, current, masked, mask, sens, state, cen, nrm, sdims, coil_dim=1):
    """RecurrentVarNetBlock.forward, recurrentvarnet.py:176-240; ``sd``: learning_rate + regularizer.*"""
    err = torch.where(mask == 0, torch.tensor([0.0], dtype=masked.dtype), current - masked)
    term = torch.cat([mri.complex_mul(mri.ifft2(k, cen, nrm, sdims), mri.complex_conj(sens)).sum(coil_dim)
                      for k in torch.split(current, 2, -1)], dim=-1).permute(0, 3, 1, 2)
    reg = {k[len("regularizer."):]: v for k, v in sd.items() if k.startswith("regularizer.")}
    term, state = conv2dgru(reg, num_layers, hidden, term, state)
    term = term.permute(0, 2, 3, 1)
    term = torch.cat([mri.fft2(mri.complex_mul(im.unsqueeze(coil_dim), sens), cen, nrm, sdims)
                      for im in torch.split(term, 2, -1)], dim=-1)
    return current - sd["learning_rate"] * err + term, state


def qvarnet_block(model_sd, unet_hp, dc_weight, masked_kspace, R2, S0, B0, phi, TEs, sens, sampling_mask, gamma, cen, nrm,
                  sdims, coil_dim):
    """qVarNetBlock.forward, qvn_block.py:133-160; the regulariser is a NormUnet (``model_sd`` keys ``unet.*``)."""
    init_eta = torch.stack([R2, S0, B0, phi], dim=1)
    maps = [(m * gamma[i]).unsqueeze(0) for i, m in enumerate((R2, S0, B0, phi))]
    init_pred = megre_signal(*maps, TEs)
    S = sens.unsqueeze(coil_dim - 1)
    soft_dc = (mri.fft2(mri.complex_mul(init_pred, S), cen, nrm, sdims) - masked_kspace) * sampling_mask * dc_weight
    init_pred = mri.complex_mul(mri.ifft2(soft_dc, cen, nrm, sdims), mri.complex_conj(S)).sum(dim=coil_dim)
    out = norm_unet(init_pred, model_sd, unet_hp["num_pools"], unet_hp.get("padding_size", 15), unet_hp.get("normalize", True))
    eta = torch.view_as_real(init_eta + torch.view_as_complex(out))
    e0 = eta[:, 0, ...]
    e0[e0 < 0] = 0
    eta[:, 0, ...] = e0
    return eta


def jrscirim_block(sd, rim_hp, num_cascades, keep_eta, seg_kind, seg_hp, input_channels, magnitude_input, consecutive_slices,
                   y, sens, mask, init_pred, target, normalize_output=True):
    """JRSCIRIMBlock.forward, jrscirim_block.py:200-333, for no_dc reconstruction modules of dimensionality 2.
    ``sd``: reconstruction_module.<i>.* + segmentation_module.*; returns (list[cascades][time steps] of complex images,
    segmentation)."""

    def cascades(y_, S_, m_, init_, tgt_, hx):
        pred, out = y_.clone(), []
        for i in range(num_cascades):
            bsd = {k[len("reconstruction_module.%d." % i):]: v for k, v in sd.items()
                   if k.startswith("reconstruction_module.%d." % i)}
            pred, hx = rim_block(bsd, rim_hp, pred, y_, S_, m_, init_, hx, 1.0, keep_eta=False if i == 0 else keep_eta)
            steps = []
            for p in pred:  # process_intermediate_pred (:335-377) with no_dc: view + crop to the target
                p = torch.view_as_complex(p)
                t = torch.view_as_complex(tgt_) if tgt_.shape[-1] == 2 else tgt_
                steps.append(mri.center_crop_to_smallest(t, p)[1])
            out.append(steps)
        return out, hx

    hx = None
    if consecutive_slices > 1:
        per_slice = []
        for s_ in range(consecutive_slices):
            init_s = init_pred[:, s_, ...]
            cas, hx = cascades(y[:, s_, ...], sens[:, s_, ...], mask[:, 0, ...], None if init_s.dim() < 4 else init_s,
                               target[:, s_, ...], hx)
            per_slice.append(torch.stack([torch.stack(c, 0) for c in cas], 0))
        preds = torch.stack(per_slice, dim=3)
        etas = [[preds[c, t] for t in range(preds.shape[1])] for c in range(preds.shape[0])]
    else:
        etas, hx = cascades(y, sens, mask, None if init_pred is None or init_pred.dim() < 4 else init_pred, target, hx)
    x = etas[-1][-1]
    if x.shape[-1] != 2:
        x = torch.view_as_real(x)
    if consecutive_slices > 1 and x.dim() == 5:
        x = x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])
    if input_channels == 1:
        x = torch.view_as_complex(x).unsqueeze(1)
        if magnitude_input:
            x = torch.abs(x)
    else:
        x = x.permute(0, 3, 1, 2)
    x = F.group_norm(x, num_groups=1)
    seg_sd = {k[len("segmentation_module."):]: v for k, v in sd.items() if k.startswith("segmentation_module.")}
    if seg_kind == "unet":
        seg = unet(x, seg_sd, seg_hp["pooling_layers"])
    else:
        seg = conv_nonlinear(x, seg_sd["0.conv_layer.weight"], None, 3, 1, None)
    seg = torch.abs(seg)
    if normalize_output:
        seg = seg / torch.max(seg)
    if consecutive_slices > 1:
        seg = seg.view([y.shape[0], y.shape[1], *seg.shape[1:]])
    return etas, seg

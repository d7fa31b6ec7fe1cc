"""Oracle (TEST INFRASTRUCTURE): the quantitative qRIM / qCIRIM path restated functionally on PyTorch-CPU.

  megre_signal / analytical_log_likelihood_gradient   mridc/collections/quantitative/models/qrim/utils.py:68-295
  qrim_block                                          mridc/collections/quantitative/models/qrim/qrim_block.py:134-240
  qcirim_forward                                      mridc/collections/quantitative/models/qcirim.py:247-341
                                                      (use_reconstruction_module = False)

Pinned by oracle/make_golden.py against the unmodified reference modules (tests/golden/qmri.npz).  Not imported by the
product package.
"""
from typing import Dict, List

import torch

from . import mri
from .nets import conv_gru_cell, conv_mgu_cell, conv_nonlinear, indrnn_cell


def megre_signal(R2star_map, S0_map, B0_map, phi_map, TEs, scaling=1e-3, no_phase=False):
    """qrim/utils.py:68-155: maps [B, H, W] -> [B, E, H, W, 2]."""
    echoes = []
    for te in TEs:
        f = torch.exp(-te * scaling * R2star_map)
        if no_phase:  # :141-152
            echoes.append(torch.stack((S0_map * f, S0_map * f), -1))
            continue
        c = torch.cos(B0_map * scaling * -te)
        s = torch.sin(B0_map * scaling * -te)
        echoes.append(torch.stack((S0_map * f * c - phi_map * f * s, S0_map * f * s + phi_map * f * c), -1))  # :107-119
    pred = torch.stack(echoes, 1)
    pred[pred != pred] = 0.0  # :121
    return pred


def analytical_log_likelihood_gradient(R2star_map, S0_map, B0_map, phi_map, TEs, sensitivity_maps, masked_kspace,
                                       sampling_mask, fft_centered, fft_normalization, spatial_dims, coil_dim,
                                       coil_combination_method="SENSE", scaling=1e-3, no_phase=False):
    """qrim/utils.py:166-295, one sample: maps [H, W], sens [C, H, W, 2], k-space [E, C, H, W, 2] -> [4, H, W]."""
    R2, S0, B0, ph = (m.unsqueeze(0) for m in (R2star_map, S0_map, B0_map, phi_map))  # :221-224
    pred = megre_signal(R2, S0, B0, ph, TEs, 1e-3, no_phase)  # :226 (the model's own scaling)
    sens = sensitivity_maps.unsqueeze(0).unsqueeze(coil_dim - 1)
    x = mri.complex_mul(pred.unsqueeze(coil_dim), sens)  # expand_op :158-163
    x[x != x] = 0
    pred_kspace = mri.fft2(x, fft_centered, fft_normalization, spatial_dims)  # :231-236
    diff = (pred_kspace - masked_kspace) * sampling_mask  # :238
    d = mri.coil_combination(mri.ifft2(diff, fft_centered, fft_normalization, spatial_dims), sens,
                             method=coil_combination_method, dim=coil_dim)  # :239-244
    s0g, r2g = [], []
    for te in TEs:
        f = torch.exp(-te * scaling * R2)
        c = torch.cos(B0 * scaling * -te)
        s = torch.sin(B0 * scaling * -te)
        s0g.append(torch.stack((f * c, -f * s), -1))  # :255-257
        r2g.append(torch.stack((-te * scaling * f * (S0 * c - ph * s), -te * scaling * f * (-S0 * s - ph * c)), -1))  # :259-271
    s0d, r2d = torch.stack(s0g, 1), torch.stack(r2g, 1)
    dr, di = d[..., 0], d[..., 1]
    s0_grad = torch.stack([dr * s0d[..., 0] - di * s0d[..., 1], dr * s0d[..., 1] + di * s0d[..., 0]], -1).squeeze()
    r2_grad = torch.stack([dr * r2d[..., 0] - di * r2d[..., 1], dr * r2d[..., 1] + di * r2d[..., 0]], -1).squeeze()
    s0_grad, r2_grad = torch.mean(s0_grad, 0), torch.mean(r2_grad, 0)  # :286-289
    return torch.stack([r2_grad[..., 0], s0_grad[..., 0], r2_grad[..., 1], s0_grad[..., 1]], 0)  # :291-294


def qrim_block(sd: Dict[str, torch.Tensor], hp: dict, masked_kspace, R2star_map_init, S0_map_init, B0_map_init,
               phi_map_init, TEs, sensitivity_maps, sampling_mask, eta=None, hx=None, gamma=None):
    """qrim_block.py:134-240.  hp keys: recurrent_layer, conv_kernels, conv_dilations, recurrent_filters,
    recurrent_kernels, recurrent_dilations, time_steps, fft_centered, fft_normalization, spatial_dims, coil_dim,
    coil_combination_method, sequence."""
    B = masked_kspace.shape[0]
    if eta is None:
        eta = torch.stack([R2star_map_init, S0_map_init, B0_map_init, phi_map_init], dim=1)  # :184-185
    if hx is None:
        hx = [eta.new_zeros((eta.size(0), f, *eta.size()[2:])) for f in hp["recurrent_filters"] if f != 0]  # :187-192
    R2 = R2star_map_init * gamma[0]  # :196-199
    S0 = S0_map_init * gamma[1]
    B0 = B0_map_init * gamma[2]
    ph = phi_map_init * gamma[3]
    rl = hp["recurrent_layer"].upper()
    nlayers = sum(1 for f in hp["recurrent_filters"] if f != 0)
    no_phase = hp.get("sequence", "MEGRE").lower() == "megre_no_phase"
    etas = []
    for _ in range(hp["time_steps"]):
        grad_eta = torch.zeros_like(eta)
        for idx in range(B):  # :204-223
            g = analytical_log_likelihood_gradient(
                R2[idx], S0[idx], B0[idx], ph[idx], TEs, sensitivity_maps[idx], masked_kspace[idx], sampling_mask[idx],
                hp["fft_centered"], hp["fft_normalization"], hp.get("spatial_dims") or [-2, -1], hp["coil_dim"],
                hp.get("coil_combination_method", "SENSE"), no_phase=no_phase).contiguous()
            grad_eta[idx] = g / 100
            grad_eta[grad_eta != grad_eta] = 0.0
        g = torch.cat([grad_eta, eta], dim=hp["coil_dim"] - 1)  # :226
        for l in range(nlayers):  # :228-230
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
                hx[l] = indrnn_cell(g, hx[l], sd[p + "rnn.ih.weight"], sd.get(p + "rnn.ih.bias"), sd[p + "rnn.hh"], k, d)
            else:
                raise ValueError("Please specify a proper recurrent layer type.")
            g = hx[l]
        g = conv_nonlinear(g, sd["final_layer.0.conv_layer.weight"], sd.get("final_layer.0.conv_layer.bias"),
                           hp["conv_kernels"][nlayers], hp["conv_dilations"][nlayers], None)  # :232
        eta = eta + g  # :233
        eta[:, 0][eta[:, 0] < 0] = 0  # :234-236
        etas.append(eta)
    return etas, None


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def qcirim_forward(sd, cfg, R2star_map_init, S0_map_init, B0_map_init, phi_map_init, TEs: List, y, sensitivity_maps,
                   mask_brain, sampling_mask):
    """qcirim.py:247-341 with use_reconstruction_module == False -> [pred, R2*, S0, B0, phi] (cascades x steps x [B,H,W])."""
    gamma = torch.tensor(list(cfg["quantitative_module_gamma_regularization_factors"]))  # :141
    hp = dict(
        recurrent_layer=cfg["quantitative_module_recurrent_layer"],
        conv_kernels=list(cfg["quantitative_module_conv_kernels"]),
        conv_dilations=list(cfg["quantitative_module_conv_dilations"]),
        recurrent_filters=list(cfg["quantitative_module_recurrent_filters"]),
        recurrent_kernels=list(cfg["quantitative_module_recurrent_kernels"]),
        recurrent_dilations=list(cfg["quantitative_module_recurrent_dilations"]),
        time_steps=cfg["quantitative_module_time_steps"], fft_centered=cfg["fft_centered"],
        fft_normalization=cfg["fft_normalization"], spatial_dims=list(cfg["spatial_dims"]), coil_dim=cfg["coil_dim"],
        coil_combination_method=cfg["coil_combination_method"],
        sequence=cfg["quantitative_module_signal_forward_model_sequence"])
    maps = [R2star_map_init / gamma[0], S0_map_init / gamma[1], B0_map_init / gamma[2], phi_map_init / gamma[3]]  # :247-250
    out = [[], [], [], []]
    for i in range(cfg["quantitative_module_num_cascades"]):
        prediction, _ = qrim_block(_sub(sd, "qcirim.%d." % i), hp, y, maps[0], maps[1], maps[2], maps[3], TEs,
                                   sensitivity_maps, sampling_mask, None, None, gamma)
        maps = [prediction[-1][:, k] for k in range(4)]  # :279-284
        steps = []
        for pred in prediction:  # :299-307 -> process_intermediate_pred(abs(pred)) -> RescaleByMax.reverse (batch-indexed)
            a = torch.abs(pred)
            x = torch.stack([a[b] * gamma[b] for b in range(a.shape[0])], 0)
            steps.append([x[:, k, ...] for k in range(4)])
        for k in range(4):
            out[k].append([s[k] for s in steps])
    return [torch.empty([]), out[0], out[1], out[2], out[3]]

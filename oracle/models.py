"""Oracle (TEST INFRASTRUCTURE): model forwards restated on PyTorch-CPU.

The reference model classes need pytorch_lightning/omegaconf (absent), so the ~15 lines of glue of each
``forward`` are restated here around the block functions of ``oracle.nets``.
"""
import math

import torch

from . import mri, nets


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def cirim_hparams(cfg):
    """reconstruction/models/cirim.py:46-88 (time_steps rounded up to a multiple of 8, :51)."""
    hp = dict(cfg)
    hp["time_steps"] = 8 * math.ceil(cfg["time_steps"] / 8)
    return hp


def cirim_process_intermediate_pred(pred, sens, target, hp, do_coil_combination=False):
    """cirim.py:167-197."""
    if not hp["no_dc"] or do_coil_combination:
        pred = mri.ifft2(pred, hp["fft_centered"], hp["fft_normalization"], hp.get("spatial_dims"))
        pred = mri.coil_combination(pred, sens, method=hp["coil_combination_method"], dim=hp["coil_dim"])
    pred = torch.view_as_complex(pred)
    _, pred = mri.center_crop_to_smallest(target, pred)
    return pred


def cirim_forward(sd, cfg, y, sens, mask, init_pred, target):
    """cirim.py:115-165.  The reference *yields* the result once; this returns it."""
    hp = cirim_hparams(cfg)
    prediction = y.clone()
    init_pred = None if init_pred is None or init_pred.dim() < 4 else init_pred
    cascades_etas = []
    for i in range(hp["num_cascades"]):
        prediction, _ = nets.rim_block(_sub(sd, "cirim.%d." % i), hp, prediction, y, sens, mask, init_pred, None,
                                       1.0, keep_eta=False if i == 0 else hp["keep_eta"])
        cascades_etas.append([cirim_process_intermediate_pred(p, sens, target, hp) for p in prediction])
    return cascades_etas


def varnet_forward(sd, cfg, y, sens, mask, init_pred, target):
    """reconstruction/models/vn.py:94-142."""
    hp = dict(cfg)
    hp["pooling_layers"] = cfg["pooling_layers"]
    est = y.clone()
    for i in range(cfg["num_cascades"]):
        est = nets.varnet_block(_sub(sd, "cascades.%d." % i), hp, est, y, sens, mask)
    est = mri.ifft2(est, cfg["fft_centered"], cfg["fft_normalization"], cfg.get("spatial_dims"))
    est = mri.coil_combination(est, sens, method=cfg["coil_combination_method"], dim=cfg["coil_dim"])
    est = torch.view_as_complex(est)
    _, est = mri.center_crop_to_smallest(target, est)
    return est


def unet_forward(sd, cfg, y, sens, mask, init_pred, target):
    """reconstruction/models/unet.py:77-121."""
    eta = torch.view_as_complex(mri.coil_combination(
        mri.ifft2(y, cfg["fft_centered"], cfg["fft_normalization"], cfg.get("spatial_dims")),
        sens, method=cfg["coil_combination_method"], dim=cfg["coil_dim"]))
    _, eta = mri.center_crop_to_smallest(target, eta)
    out = nets.norm_unet(torch.view_as_real(eta.unsqueeze(cfg["coil_dim"])), _sub(sd, "unet."),
                         cfg["pooling_layers"], cfg["padding_size"], cfg["normalize"])
    return torch.view_as_complex(out).squeeze(cfg["coil_dim"])


def zf_forward(cfg, y, sens, mask, target=None):
    """reconstruction/models/zf.py:62-100."""
    pred = mri.coil_combination(
        mri.ifft2(y, cfg["fft_centered"], cfg["fft_normalization"], cfg.get("spatial_dims")),
        sens, method=cfg["coil_combination_method"].upper(), dim=cfg["coil_dim"])
    pred = mri.check_stacked_complex(pred)
    _, pred = mri.center_crop_to_smallest(target, pred)
    return pred

"""Thin tensor-level wrappers over the C-ABI (argument checking, mask canonicalisation, workspaces).

Everything here takes / returns CUDA fp32 tensors; nothing computes on the CPU.
"""
from typing import Optional, Tuple

import torch

from . import _lib

NORM_CODES = {"backward": 0, "none": 0, "ortho": 1, "forward": 2}
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
PAD_ZERO, PAD_REPLICATE = 0, 1


def norm_code(normalization: str) -> int:
    key = "none" if normalization.lower() == "none" else normalization
    if key not in NORM_CODES:
        raise RuntimeError("Invalid normalization mode: %s" % normalization)
    return NORM_CODES[key]


def check_spatial_dims(spatial_dims, ndim_complex=4):
    """The fused operators transform the last two (H, W) axes of [B, C, H, W] complex data."""
    if spatial_dims is None:
        return
    dims = sorted(int(d) % ndim_complex for d in spatial_dims)
    if dims != [ndim_complex - 2, ndim_complex - 1]:
        raise NotImplementedError(
            "mridc_b200 fused operators support spatial_dims over the last two (H, W) axes only; got %s" % (
                list(spatial_dims),))


def canonical_mask(mask: torch.Tensor, B: int, H: int, W: int) -> Tuple[torch.Tensor, int, int, int]:
    """-> (contiguous [mb, mh, W] tensor of uint8 or float32, dtype code, mb, mh).

    Accepts what the reference broadcasts against [B, C, H, W, 2] k-space: [B|1, 1, H|1, W, 1] (uint8 from the
    data pipeline, float32 from apply_mask, bool), SURVEY section 7 'API quirks'."""
    _lib.require_cuda(mask, "mask", None)
    m = mask
    if m.dim() == 5:
        if m.shape[1] != 1 or m.shape[4] != 1:
            m = None
        else:
            m = m[:, 0, :, :, 0]
    elif m.dim() == 4 and m.shape[1] == 1:  # [B|1, 1, H|1, W]
        m = m[:, 0]
    elif m.dim() == 3:
        pass
    else:
        m = None
    if m is None or m.shape[0] not in (1, B) or m.shape[1] not in (1, H) or m.shape[2] not in (1, W):
        # generic broadcast (rare): materialise [B, H, W]
        full = mask.expand(B, 1, H, W, 1) if mask.dim() == 5 else torch.broadcast_to(mask, (B, 1, H, W, 1))
        m = full[:, 0, :, :, 0]
    if m.shape[2] == 1 and W != 1:
        m = m.expand(m.shape[0], m.shape[1], W)
    if m.dtype in (torch.uint8, torch.bool):
        m = m.contiguous().view(torch.uint8) if m.dtype == torch.bool else m.contiguous()
        code = 0
    else:
        m = m.to(torch.float32).contiguous()
        code = 1
    return m, code, int(m.shape[0]), int(m.shape[1])


def _ws(B, C, H, W, device, halves=2):
    return torch.empty((halves, B, C, H, W, 2), dtype=torch.float32, device=device)


def _check5(t, name):
    _lib.require_cuda(t, name)
    if t.dim() != 5 or t.shape[-1] != 2:
        raise ValueError("%s must be [B, C, H, W, 2] (got %s)" % (name, tuple(t.shape)))
    return t.contiguous()


def dc_hybrid_prepare(y, mask, centered, ws=None):
    """Hybrid-space k-space for 1-D (column) masks: (1/H) * centred inverse DFT of y along H, sampled columns packed at the
    front of each row.  Prepared once per slice batch; feeds ``dc_rim_grad(..., y_hybrid=...)`` (see mridc_b200.h).
    Returns None when the mask depends on k_h (the general three-pass operator applies)."""
    y = _check5(y, "masked_kspace")
    B, C, H, W, _ = y.shape
    m, code, mb, mh = canonical_mask(mask, B, H, W)
    if mh != 1:
        return None
    if ws is None:
        ws = _ws(B, C, H, W, y.device, halves=1)
    yh = torch.empty_like(y)
    _lib.check(_lib.load().mrb_dc_hybrid_prepare(_lib.ptr(y), _lib.ptr(m), code, mb, _lib.ptr(yh), B, C, H, W,
                                                 int(bool(centered)), _lib.ptr(ws), ws.numel() * 4, _lib.stream_ptr()))
    return yh


def dc_rim_grad(eta, y, S, mask, sigma, centered, normalization, out=None, ws=None, nhwc=False, y_hybrid=None):
    """rim_utils.py:11-67 -> [B, 4, H, W] (or channels-last [B, H, W, 4] when nhwc).
    y_hybrid: result of dc_hybrid_prepare(y, mask, centered) -> single-kernel row form (1-D masks only).
    nhwc == 2 (hybrid form, W == 320, C <= 16): ``out`` is a zero-initialised G8 byte buffer (mrb_g8_bytes) and receives the
    split-bf16 conv input with its replicate border."""
    y = _check5(y, "masked_kspace")
    S = _check5(S, "sense")
    B, C, H, W, _ = y.shape
    _lib.require_cuda(eta, "eta")
    if tuple(eta.shape) != (B, H, W, 2) or S.shape != y.shape:
        raise ValueError("shape mismatch: eta %s, y %s, S %s" % (tuple(eta.shape), tuple(y.shape), tuple(S.shape)))
    eta = eta.contiguous()
    m, code, mb, mh = canonical_mask(mask, B, H, W)
    if out is None:
        out = torch.empty((B, H, W, 4) if nhwc else (B, 4, H, W), dtype=torch.float32, device=y.device)
    lib = _lib.load()
    if y_hybrid is not None:
        if mh != 1 or y_hybrid.shape != y.shape:
            raise ValueError("y_hybrid needs a 1-D column mask and the shape of masked_kspace")
        _lib.check(lib.mrb_dc_rim_grad_hybrid(_lib.ptr(eta), _lib.ptr(y_hybrid), _lib.ptr(S), _lib.ptr(m), code, mb,
                                              1.0 / (float(sigma) ** 2.0), _lib.ptr(out), 2 if nhwc == 2 else int(bool(nhwc)),
                                              B, C, H, W,
                                              int(bool(centered)), norm_code(normalization), _lib.stream_ptr()))
        return out
    if ws is None:
        ws = _ws(B, C, H, W, y.device)
    _lib.check(lib.mrb_dc_rim_grad(_lib.ptr(eta), _lib.ptr(y), _lib.ptr(S), _lib.ptr(m), code, mb, mh,
                                   1.0 / (float(sigma) ** 2.0), _lib.ptr(out), int(bool(nhwc)), B, C, H, W,
                                   int(bool(centered)),
                                   norm_code(normalization), _lib.ptr(ws), ws.numel() * 4, _lib.stream_ptr()))
    return out


def sens_reduce(x, S, centered, normalization, out=None, ws=None):
    """sum_c ifft2(x) * conj(S): [B,C,H,W,2] -> [B,H,W,2]  (vn_block.py:71-87 without keepdim)."""
    x = _check5(x, "x")
    S = _check5(S, "sensitivity_maps")
    if S.shape != x.shape:
        raise ValueError("shape mismatch: x %s, S %s" % (tuple(x.shape), tuple(S.shape)))
    B, C, H, W, _ = x.shape
    if out is None:
        out = torch.empty((B, H, W, 2), dtype=torch.float32, device=x.device)
    if ws is None:
        ws = _ws(B, C, H, W, x.device, halves=1)
    _lib.check(_lib.load().mrb_sens_reduce(_lib.ptr(x), _lib.ptr(S), _lib.ptr(out), B, C, H, W, int(bool(centered)),
                                           norm_code(normalization), _lib.ptr(ws), ws.numel() * 4,
                                           _lib.stream_ptr()))
    return out


def sens_expand_softdc(img, S, base, pred, y, mask, dc_weight, no_dc, centered, normalization, out=None, ws=None):
    """E = fft2(S*img); out = no_dc ? E : base - where(mask, pred - y, 0)*dc_weight - E."""
    S = _check5(S, "sensitivity_maps")
    B, C, H, W, _ = S.shape
    _lib.require_cuda(img, "img")
    img = img.contiguous()
    if img.numel() != B * H * W * 2:
        raise ValueError("img must hold [B, H, W, 2] (got %s)" % (tuple(img.shape),))
    m, code, mb, mh = None, 0, 1, 1
    if not no_dc:
        base, pred, y = _check5(base, "base"), _check5(pred, "pred"), _check5(y, "ref_kspace")
        _lib.require_cuda(dc_weight, "dc_weight")  # device-resident learnable scalar: no host sync
        m, code, mb, mh = canonical_mask(mask, B, H, W)
    if out is None:
        out = torch.empty_like(S)
    if ws is None:
        ws = _ws(B, C, H, W, S.device, halves=1)
    _lib.check(_lib.load().mrb_sens_expand_softdc(
        _lib.ptr(img), _lib.ptr(S), _lib.ptr(base) if not no_dc else None, _lib.ptr(pred) if not no_dc else None,
        _lib.ptr(y) if not no_dc else None, _lib.ptr(m), code, mb, mh,
        _lib.ptr(dc_weight) if not no_dc else None, int(bool(no_dc)),
        _lib.ptr(out), B, C, H, W, int(bool(centered)), norm_code(normalization), _lib.ptr(ws), ws.numel() * 4,
        _lib.stream_ptr()))
    return out


def conv2d(x, weight, bias, k, dil, pad_mode, act=ACT_NONE, slope=0.0, add=None, add_scale=None, residual=None,
           out=None, x_bstride=None, out_bstride=None, N=None, Cin=None, H=None, W=None):
    """'same' conv (odd k).  x [N,Cin,H,W]; weight [Cout,Cin,k,k]; optional fused epilogues (see the header)."""
    _lib.require_cuda(x, "x")
    _lib.require_cuda(weight, "weight")
    if N is None:
        if x.dim() != 4:
            raise ValueError("conv input must be [N, C, H, W]")
        x = x.contiguous()
        N, Cin, H, W = x.shape
    Cout = weight.shape[0]
    if weight.shape[1] != Cin or weight.shape[2] != k or weight.shape[3] != k:
        raise RuntimeError("input has inconsistent input_size: got %d, expected %d" % (Cin, weight.shape[1]))
    weight = weight.contiguous()
    nhwc = residual is not None
    if out is None:
        out = torch.empty((N, H, W, Cout) if nhwc else (N, Cout, H, W), dtype=torch.float32, device=x.device)
    xbs = Cin * H * W if x_bstride is None else x_bstride
    obs = Cout * H * W if out_bstride is None else out_bstride
    _lib.check(_lib.load().mrb_conv2d(
        _lib.ptr(x), xbs, _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(out), obs, N, Cin, Cout, H, W, int(k), int(dil),
        pad_mode, act, float(slope), _lib.ptr(add), _lib.ptr(add_scale), _lib.ptr(residual), int(nhwc),
        _lib.stream_ptr()))
    return out

"""Generate tests/golden/*.npz from the UNMODIFIED reference (container only; needs /root/reference).

    python -m oracle.make_golden

For every case the reference leaf modules (imported through oracle/ref_import.py) are run on seeded
inputs; the oracle restatement is run on the same inputs and must agree before the vectors are written.
The vectors travel with the repo, so the GPU box (no /root/reference) can still pin the oracle and the
CUDA path against real reference outputs.  TEST INFRASTRUCTURE.
"""
import os
import sys

import numpy as np
import torch

from . import metrics as ometrics  # noqa: F401
from . import models as omodels
from . import mri as omri
from . import nets as onets
from . import qnets as oqnets
from .ref_import import Ref

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _close(a, b, what, rtol=1e-6, atol=1e-7):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.is_complex():
        a, b = torch.view_as_real(a), torch.view_as_real(b)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    if not torch.allclose(a, b, rtol=rtol, atol=atol):
        raise AssertionError("oracle != reference for %s (max abs err %g)" % (what, err))
    return err


def _np(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v = v.detach()
            v = torch.view_as_real(v) if v.is_complex() else v
            v = v.numpy()
        out[k] = np.asarray(v)
    return out


def small_inputs(B, C, H, W, seed, mask_kind="1d", mask_dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(B, C, H, W, 2, generator=g)
    S = torch.randn(B, C, H, W, 2, generator=g) * 0.5
    eta = torch.randn(B, H, W, 2, generator=g)
    if mask_kind == "1d":
        m = (torch.rand(1, 1, 1, W, 1, generator=g) < 0.4).float()
        m[..., W // 2, :] = 1
    elif mask_kind == "2d":
        m = (torch.rand(1, 1, H, W, 1, generator=g) < 0.4).float()
    else:  # per-batch 2-D
        m = (torch.rand(B, 1, H, W, 1, generator=g) < 0.4).float()
    y = y * m
    return y, S, eta, m.to(mask_dtype)


def gen_masks(R):
    from mridc_b200 import synth

    out = {}
    cases = [("random", R.subsample.RandomMaskFunc, synth.RandomMask1D, [0.08], [4], (1, 320, 320, 2), 123),
             ("random8", R.subsample.RandomMaskFunc, synth.RandomMask1D, [0.04], [8], (1, 640, 320, 2), 7),
             ("equi", R.subsample.Equispaced1DMaskFunc, synth.Equispaced1DMask, [0.08], [4], (1, 320, 320, 2), 123),
             ("equi8", R.subsample.Equispaced1DMaskFunc, synth.Equispaced1DMask, [0.04], [8], (1, 640, 320, 2), 123),
             ("equi_multi", R.subsample.Equispaced1DMaskFunc, synth.Equispaced1DMask, [0.08, 0.04], [4, 8],
              (1, 218, 170, 2), (1, 2, 3))]
    for name, rcls, mcls, cf, acc, shape, seed in cases:
        rm, ra = rcls(cf, acc)(shape, seed)
        mm, ma = mcls(cf, acc)(shape, seed)
        assert torch.equal(rm, mm) and ra == ma, name
        out[name] = rm.numpy()
        out[name + "_acc"] = np.asarray(ra)
    for name, acc in (("gauss4", 4), ("gauss8", 8)):
        np.random.seed(123)
        rm, ra = R.subsample_nn.Gaussian1DMaskFunc([0.7], [acc])((1, 320, 320, 2), 0, scale=0.02)
        np.random.seed(123)
        mm, ma = synth.Gaussian1DMask([0.7], [acc])((1, 320, 320, 2), 0, scale=0.02)
        assert torch.equal(rm, mm) and ra == ma, name
        out[name] = rm.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "masks.npz"), **out)
    print("masks.npz: %d arrays, bit-exact vs reference" % len(out))


def gen_prims(R):
    out = {}
    g = torch.Generator().manual_seed(11)
    shapes = [(3, 3, 2), (4, 6, 2), (10, 8, 4, 2), (2, 3, 7, 5, 2), (2, 12, 10, 2), (1, 2, 15, 9, 2)]
    i = 0
    for shp in shapes:
        x = torch.randn(*shp, generator=g)
        for cen in (False, True):
            for nrm in ("backward", "ortho", "forward", "none"):
                for inv in (False, True):
                    rf = R.fft.ifft2 if inv else R.fft.fft2
                    of = omri.ifft2 if inv else omri.fft2
                    r = rf(x, centered=cen, normalization=nrm)
                    _close(of(x, cen, nrm), r, "fft %s" % (shp,))
                    out["fft%d_x" % i] = x.numpy()
                    out["fft%d_out" % i] = r.numpy()
                    out["fft%d_cfg" % i] = np.asarray([int(cen), ["backward", "ortho", "forward", "none"].index(nrm),
                                                       int(inv)])
                    i += 1
    # non-default spatial dims
    x = torch.randn(3, 6, 5, 4, 2, generator=g)
    r = R.fft.fft2(x, centered=True, normalization="ortho", spatial_dims=[-3, -2])
    _close(omri.fft2(x, True, "ortho", [-3, -2]), r, "fft spatial dims")
    out["fftsd_x"], out["fftsd_out"] = x.numpy(), r.numpy()
    out["nfft"] = np.asarray(i)
    # shifts / roll on integer data (bit exact)
    a = torch.arange(7 * 6 * 5).reshape(7, 6, 5)
    out["roll_x"] = a.numpy()
    out["roll_a"] = R.fft.roll(a, [2, -3], [0, 2]).numpy()
    out["roll_b"] = R.fft.roll(a, [9], [1]).numpy()
    out["fftshift"] = R.fft.fftshift(a).numpy()
    out["ifftshift"] = R.fft.ifftshift(a).numpy()
    out["fftshift_d"] = R.fft.fftshift(a, dim=[0, 1]).numpy()
    assert torch.equal(omri.fftshift(a), R.fft.fftshift(a)) and torch.equal(omri.ifftshift(a), R.fft.ifftshift(a))
    assert torch.equal(omri.roll(a, [2, -3], [0, 2]), R.fft.roll(a, [2, -3], [0, 2]))
    # complex utilities
    x = torch.randn(2, 5, 6, 7, 2, generator=g)
    y = torch.randn(2, 5, 6, 7, 2, generator=g)
    yb = torch.randn(1, 1, 6, 7, 2, generator=g)
    U = R.utils
    out["cx"], out["cy"], out["cyb"] = x.numpy(), y.numpy(), yb.numpy()
    for name, rv, ov in [
        ("cmul", U.complex_mul(x, y), omri.complex_mul(x, y)),
        ("cmulb", U.complex_mul(x, yb), omri.complex_mul(x, yb)),
        ("cconj", U.complex_conj(x), omri.complex_conj(x)),
        ("cabs", U.complex_abs(x), omri.complex_abs(x)),
        ("cabssq", U.complex_abs_sq(x), omri.complex_abs_sq(x)),
        ("rss1", U.rss(x, 1), omri.rss(x, 1)),
        ("rss0", U.rss(x, 0), omri.rss(x, 0)),
        ("rssc1", U.rss_complex(x, 1), omri.rss_complex(x, 1)),
        ("sense1", U.sense(x, y, 1), omri.sense(x, y, 1)),
        ("cc_sense", U.coil_combination(x, y, "SENSE", 1), omri.coil_combination(x, y, "SENSE", 1)),
        ("cc_rss", U.coil_combination(x, y, "RSS", 1), omri.coil_combination(x, y, "RSS", 1)),
    ]:
        _close(ov, rv, name, rtol=0, atol=0)
        out[name] = rv.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "prims.npz"), **out)
    print("prims.npz: %d arrays" % len(out))


def gen_dc(R):
    out = {}
    i = 0
    for (B, C, H, W), mk, md, cen, nrm, sigma in [
        ((2, 3, 12, 10), "1d", torch.float32, True, "ortho", 1.0),
        ((2, 3, 12, 10), "1d", torch.uint8, False, "backward", 1.0),
        ((1, 4, 9, 15), "2d", torch.float32, True, "backward", 0.5),
        ((2, 2, 16, 6), "b2d", torch.uint8, True, "forward", 1.0),
        ((1, 5, 20, 24), "1d", torch.bool, False, "ortho", 2.0),
    ]:
        y, S, eta, m = small_inputs(B, C, H, W, 100 + i, mk, md)
        r = R.rim_utils.log_likelihood_gradient(eta, y, S, m, sigma, cen, nrm, [-2, -1], 1)
        _close(onets.log_likelihood_gradient(eta, y, S, m, sigma, cen, nrm, [-2, -1], 1), r, "dc grad %d" % i)
        vb = R.vn_block.VarNetBlock(torch.nn.Identity(), cen, nrm, [-2, -1], 1)
        red = vb.sens_reduce(y, S)
        exp = vb.sens_expand(red, S)
        _close(onets.sens_reduce(y, S, cen, nrm, [-2, -1], 1), red, "sens_reduce %d" % i)
        _close(onets.sens_expand(red, S, cen, nrm, [-2, -1]), exp, "sens_expand %d" % i)
        out.update({"dc%d_%s" % (i, k): v for k, v in _np(dict(
            y=y, S=S, eta=eta, mask=m.to(torch.uint8) if m.dtype == torch.bool else m, grad=r, red=red, exp=exp)).items()})
        out["dc%d_cfg" % i] = np.asarray([int(cen), ["backward", "ortho", "forward"].index(nrm), sigma,
                                          {"torch.float32": 1, "torch.uint8": 0, "torch.bool": 2}[str(md)]])
        i += 1
    out["ndc"] = np.asarray(i)
    np.savez_compressed(os.path.join(GOLDEN, "dc.npz"), **out)
    print("dc.npz: %d arrays" % len(out))


RIM_HP = dict(conv_filters=[16, 16, 2], conv_kernels=[5, 3, 3], conv_dilations=[1, 2, 1],
              conv_bias=[True, True, False], recurrent_filters=[16, 16, 0], recurrent_kernels=[1, 1, 0],
              recurrent_dilations=[1, 1, 0], recurrent_bias=[True, True, False], depth=2, time_steps=8, conv_dim=2,
              spatial_dims=[-2, -1], coil_dim=1, dimensionality=2)


def _ref_rim(R, layer, hp, no_dc, cen, nrm, seed):
    torch.manual_seed(seed)
    blk = R.rim_block.RIMBlock(recurrent_layer=layer, conv_filters=hp["conv_filters"], conv_kernels=hp["conv_kernels"],
                               conv_dilations=hp["conv_dilations"], conv_bias=hp["conv_bias"],
                               recurrent_filters=hp["recurrent_filters"], recurrent_kernels=hp["recurrent_kernels"],
                               recurrent_dilations=hp["recurrent_dilations"], recurrent_bias=hp["recurrent_bias"],
                               depth=2, time_steps=hp["time_steps"], conv_dim=2, no_dc=no_dc, fft_centered=cen,
                               fft_normalization=nrm, spatial_dims=[-2, -1], coil_dim=1, dimensionality=2)
    return blk.eval()


def gen_rim(R):
    out = {}
    i = 0
    for layer, rk, no_dc, cen, nrm, md in [("GRU", 1, True, False, "backward", torch.uint8),
                                           ("GRU", 3, True, True, "ortho", torch.float32),
                                           ("IndRNN", 1, True, True, "ortho", torch.float32),
                                           ("MGU", 1, True, False, "ortho", torch.float32),
                                           ("GRU", 1, False, True, "ortho", torch.bool)]:
        hp = dict(RIM_HP, recurrent_layer=layer, no_dc=no_dc, fft_centered=cen, fft_normalization=nrm,
                  recurrent_kernels=[rk, rk, 0])
        blk = _ref_rim(R, layer, hp, no_dc, cen, nrm, seed=1 + i)
        sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
        y, S, _, m = small_inputs(2, 3, 18, 14, 200 + i, "1d", md)
        with torch.no_grad():
            etas, hx = blk(y.clone(), y, S, m, None, None, 1.0, False)
            o_etas, o_hx = onets.rim_block(sd, hp, y.clone(), y, S, m, None, None, 1.0, False)
        for a, b in zip(o_etas, etas):
            _close(a, b, "rim %s step" % layer, rtol=1e-5, atol=1e-6)
        out.update({"rim%d_%s" % (i, k): v for k, v in _np(dict(
            y=y, S=S, mask=m.to(torch.uint8) if m.dtype == torch.bool else m, last=etas[-1], first=etas[0], h0=hx[0],
            h1=hx[1])).items()})
        out.update({"rim%d_w_%s" % (i, k): v.numpy() for k, v in sd.items()})
        out["rim%d_cfg" % i] = np.asarray([["GRU", "IndRNN", "MGU"].index(layer), rk, int(no_dc), int(cen),
                                           ["backward", "ortho", "forward"].index(nrm),
                                           {"torch.float32": 1, "torch.uint8": 0, "torch.bool": 2}[str(md)]])
        i += 1
    out["nrim"] = np.asarray(i)
    np.savez_compressed(os.path.join(GOLDEN, "rim.npz"), **out)
    print("rim.npz: %d arrays" % len(out))


def gen_rim3d(R):
    """RIMBlock with dimensionality == 3 / conv_dim == 3 and the IndRNN cell -- the only cell whose 3-D path runs in the
    reference (test_cirim.py:155-290 uses it): reference block -> oracle restatement -> rim3d.npz."""
    out = {}
    i = 0
    for layer, (batch, slices), cen, nrm in [("IndRNN", (1, 3), True, "ortho"), ("IndRNN", (2, 2), False, "backward")]:
        hp = dict(RIM_HP, recurrent_layer=layer, no_dc=True, fft_centered=cen, fft_normalization=nrm, time_steps=4,
                  conv_dim=3, dimensionality=3)
        torch.manual_seed(40 + i)
        blk = R.rim_block.RIMBlock(recurrent_layer=layer, conv_filters=hp["conv_filters"], conv_kernels=hp["conv_kernels"],
                                   conv_dilations=hp["conv_dilations"], conv_bias=hp["conv_bias"],
                                   recurrent_filters=hp["recurrent_filters"], recurrent_kernels=hp["recurrent_kernels"],
                                   recurrent_dilations=hp["recurrent_dilations"], recurrent_bias=hp["recurrent_bias"],
                                   depth=2, time_steps=hp["time_steps"], conv_dim=3, no_dc=True, fft_centered=cen,
                                   fft_normalization=nrm, spatial_dims=[-2, -1], coil_dim=1, dimensionality=3).eval()
        sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
        y, S, _, m = small_inputs(batch * slices, 3, 15, 12, 300 + i, "1d", torch.float32)
        y, S = y.reshape(batch, slices, *y.shape[1:]), S.reshape(batch, slices, *S.shape[1:])
        m = m.reshape(1, 1, *m.shape[1:]).expand(batch, slices, *m.shape[1:]).contiguous()
        with torch.no_grad():
            etas, hx = blk(y.clone(), y, S, m, None, None, 1.0, False)
            o_etas, o_hx = onets.rim_block_3d(sd, hp, y.clone(), y, S, m, None, None, 1.0, False)
        for a, b in zip(o_etas, etas):
            _close(a, b, "rim3d %s step" % layer, rtol=1e-5, atol=1e-6)
        for a, b in zip(o_hx, hx):
            _close(a, b, "rim3d %s hidden" % layer, rtol=1e-5, atol=1e-6)
        out.update({"rim%d_%s" % (i, k): v for k, v in _np(dict(y=y, S=S, mask=m, last=etas[-1], first=etas[0],
                                                                 h0=hx[0], h1=hx[1])).items()})
        out.update({"rim%d_w_%s" % (i, k): v.numpy() for k, v in sd.items()})
        out["rim%d_cfg" % i] = np.asarray([["GRU", "IndRNN", "MGU"].index(layer), int(cen),
                                           ["backward", "ortho", "forward"].index(nrm), hp["time_steps"]])
        i += 1
    out["nrim"] = np.asarray(i)
    np.savez_compressed(os.path.join(GOLDEN, "rim3d.npz"), **out)
    print("rim3d.npz: %d arrays" % len(out))


def gen_unet(R):
    out = {}
    i = 0
    for (B, H, W), chans, pools, padsz in [((2, 20, 24), 6, 2, 11), ((1, 15, 18), 4, 2, 15), ((1, 18, 22), 4, 3, 1)]:
        torch.manual_seed(5 + i)
        nu = R.unet_block.NormUnet(chans=chans, num_pools=pools, padding_size=padsz, normalize=True).eval()
        sd = {k: v.detach().clone() for k, v in nu.state_dict().items()}
        g = torch.Generator().manual_seed(300 + i)
        x = torch.randn(B, 1, H, W, 2, generator=g)
        with torch.no_grad():
            r = nu(x)
        _close(onets.norm_unet(x, sd, pools, padsz, True), r, "normunet %d" % i, rtol=1e-5, atol=1e-6)
        out["unet%d_x" % i], out["unet%d_out" % i] = x.numpy(), r.numpy()
        out["unet%d_cfg" % i] = np.asarray([chans, pools, padsz])
        out.update({"unet%d_w_%s" % (i, k): v.numpy() for k, v in sd.items()})
        i += 1
    out["nunet"] = np.asarray(i)
    # VarNetBlock with a NormUnet regulariser
    for j, (cen, nrm, no_dc, md) in enumerate([(True, "ortho", False, torch.float32),
                                               (False, "backward", False, torch.uint8),
                                               (True, "ortho", True, torch.float32)]):
        torch.manual_seed(40 + j)
        vb = R.vn_block.VarNetBlock(R.unet_block.NormUnet(chans=4, num_pools=2, padding_size=11, normalize=True),
                                    cen, nrm, [-2, -1], 1, no_dc).eval()
        with torch.no_grad():
            vb.dc_weight.fill_(0.7)
        sd = {k: v.detach().clone() for k, v in vb.state_dict().items()}
        y, S, _, m = small_inputs(2, 3, 16, 12, 400 + j, "1d", md)
        g = torch.Generator().manual_seed(450 + j)
        pred = torch.randn(2, 3, 16, 12, 2, generator=g)
        with torch.no_grad():
            r = vb(pred, y, S, m)
        hp = dict(pooling_layers=2, padding_size=11, normalize=True, no_dc=no_dc, fft_centered=cen,
                  fft_normalization=nrm, spatial_dims=[-2, -1], coil_dim=1)
        _close(onets.varnet_block(sd, hp, pred, y, S, m), r, "varnet block %d" % j, rtol=1e-5, atol=1e-6)
        out.update({"vn%d_%s" % (j, k): v for k, v in _np(dict(pred=pred, y=y, S=S, mask=m, out=r)).items()})
        out["vn%d_cfg" % j] = np.asarray([int(cen), ["backward", "ortho", "forward"].index(nrm), int(no_dc)])
        out.update({"vn%d_w_%s" % (j, k): v.numpy() for k, v in sd.items()})
    out["nvn"] = np.asarray(3)
    np.savez_compressed(os.path.join(GOLDEN, "unet_vn.npz"), **out)
    print("unet_vn.npz: %d arrays" % len(out))


def gen_models(R):
    """Model-level vectors: reference *blocks* composed with the restated forward glue (the model classes
    themselves need pytorch_lightning and cannot be imported)."""
    from mridc_b200 import synth

    out = {}
    B, C, H, W = 1, 4, 32, 24
    batch = synth.make_batch(B, C, H, W, synth.Equispaced1DMask([0.08], [4]), seed=123, centered=True,
                             normalization="ortho", mask_dtype="uint8")
    y, S, m, tgt = batch["y"], batch["sensitivity_maps"], batch["mask"], batch["target"]
    out.update({"in_" + k: v for k, v in _np(dict(y=y, S=S, mask=m, target=tgt)).items()})
    # CIRIM 2 cascades x 8 steps, GRU, 16 filters
    cfg = dict(RIM_HP, recurrent_layer="GRU", no_dc=True, fft_centered=True, fft_normalization="ortho",
               num_cascades=2, keep_eta=True, coil_combination_method="SENSE")
    sd = {}
    blocks = []
    for c in range(2):
        blk = _ref_rim(R, "GRU", cfg, True, True, "ortho", seed=70 + c)
        blocks.append(blk)
        sd.update({"cirim.%d.%s" % (c, k): v.detach().clone() for k, v in blk.state_dict().items()})
    with torch.no_grad():
        pred = y.clone()
        casc = []
        for c, blk in enumerate(blocks):
            pred, _ = blk(pred, y, S, m, None, None, 1.0, keep_eta=False if c == 0 else True)
            casc.append([torch.view_as_complex(p) for p in pred])
        o = omodels.cirim_forward(sd, cfg, y, S, m, None, tgt)
    for c in range(2):
        for t in range(8):
            _close(o[c][t], casc[c][t], "cirim c%d t%d" % (c, t), rtol=1e-5, atol=1e-6)
    out["cirim_out"] = torch.view_as_real(torch.stack([torch.stack(c) for c in casc])).numpy()
    out.update({"cirim_w_" + k: v.numpy() for k, v in sd.items()})
    # E2EVN 3 cascades, chans 4, 2 pools
    vcfg = dict(num_cascades=3, channels=4, pooling_layers=2, padding_size=11, normalize=True, no_dc=False,
                fft_centered=True, fft_normalization="ortho", spatial_dims=[-2, -1], coil_dim=1,
                coil_combination_method="SENSE")
    vsd = {}
    vbs = []
    for c in range(3):
        torch.manual_seed(80 + c)
        vb = R.vn_block.VarNetBlock(R.unet_block.NormUnet(chans=4, num_pools=2, padding_size=11, normalize=True),
                                    True, "ortho", [-2, -1], 1, False).eval()
        vbs.append(vb)
        vsd.update({"cascades.%d.%s" % (c, k): v.detach().clone() for k, v in vb.state_dict().items()})
    with torch.no_grad():
        est = y.clone()
        for vb in vbs:
            est = vb(est, y, S, m)
        est = R.utils.coil_combination(R.fft.ifft2(est, True, "ortho", [-2, -1]), S, "SENSE", 1)
        est = torch.view_as_complex(est)
        o = omodels.varnet_forward(vsd, vcfg, y, S, m, None, tgt)
    _close(o, est, "varnet forward", rtol=1e-4, atol=1e-5)
    out["vn_out"] = torch.view_as_real(est).numpy()
    out.update({"vn_w_" + k: v.numpy() for k, v in vsd.items()})
    # ZF (SENSE and RSS) and UNet model
    for meth in ("SENSE", "RSS"):
        with torch.no_grad():
            r = R.utils.check_stacked_complex(R.utils.coil_combination(R.fft.ifft2(y, True, "ortho", [-2, -1]), S,
                                                                       meth, 1))
        zc = dict(coil_combination_method=meth.lower(), fft_centered=True, fft_normalization="ortho",
                  spatial_dims=[-2, -1], coil_dim=1)
        _close(omodels.zf_forward(zc, y, S, m, tgt), r, "zf " + meth)
        out["zf_" + meth] = torch.view_as_real(r).numpy()
    torch.manual_seed(90)
    nu = R.unet_block.NormUnet(chans=4, num_pools=2, padding_size=11, normalize=True).eval()
    usd = {"unet." + k: v.detach().clone() for k, v in nu.state_dict().items()}
    ucfg = dict(channels=4, pooling_layers=2, padding_size=11, normalize=True, fft_centered=True,
                fft_normalization="ortho", spatial_dims=[-2, -1], coil_dim=1, coil_combination_method="SENSE")
    with torch.no_grad():
        eta = torch.view_as_complex(R.utils.coil_combination(R.fft.ifft2(y, True, "ortho", [-2, -1]), S, "SENSE", 1))
        r = torch.view_as_complex(nu(torch.view_as_real(eta.unsqueeze(1)))).squeeze(1)
    _close(omodels.unet_forward(usd, ucfg, y, S, m, None, tgt), r, "unet forward", rtol=1e-5, atol=1e-6)
    out["unet_out"] = torch.view_as_real(r).numpy()
    out.update({"unet_w_" + k: v.numpy() for k, v in usd.items()})
    np.savez_compressed(os.path.join(GOLDEN, "models.npz"), **out)
    print("models.npz: %d arrays" % len(out))


QRIM_HP = dict(conv_filters=[16, 16, 4], conv_kernels=[5, 3, 3], conv_dilations=[1, 2, 1], conv_bias=[True, True, False],
               recurrent_filters=[16, 16, 0], recurrent_kernels=[1, 1, 0], recurrent_dilations=[1, 1, 0],
               recurrent_bias=[True, True, False], time_steps=8, spatial_dims=[-2, -1], coil_dim=2,
               coil_combination_method="SENSE")
QGAMMA = [150.0, 150.0, 1000.0, 150.0]


def qmri_inputs(B, E, C, H, W, seed, mask_kind):
    """Seeded qMRI-shaped inputs: maps in physical ranges (R2* ~ 0..100 1/s, B0 ~ +-60 Hz-ish), k-space of the MEGRE
    signal of slightly different maps (so the residual is neither zero nor huge), mask 1-D / 2-D / per-batch 2-D."""
    g = torch.Generator().manual_seed(seed)
    r2 = torch.rand(B, H, W, generator=g) * 90 + 5
    s0 = torch.randn(B, H, W, generator=g)
    b0 = torch.randn(B, H, W, generator=g) * 40
    ph = torch.randn(B, H, W, generator=g)
    S = torch.randn(B, C, H, W, 2, generator=g) * 0.5
    tes = [3.0, 11.5, 20.0, 28.5][:E] if E <= 4 else [3.0 + 4.25 * e for e in range(E)]
    sig = oqnets.megre_signal(r2 * 1.1, s0 + 0.2, b0 * 0.9, ph - 0.1, tes)
    k = omri.fft2(omri.complex_mul(sig.unsqueeze(2), S.unsqueeze(1)), False, "backward", [-2, -1])
    k = k + 0.05 * torch.randn(k.shape, generator=g)
    if mask_kind == "1d":
        m = (torch.rand(B, 1, 1, W, 1, generator=g) < 0.4).float()
        m[..., W // 2, :] = 1
    elif mask_kind == "2d":
        # one pattern for the whole batch; the reference indexes sampling_mask[idx], so it still carries the batch dim
        m = (torch.rand(1, 1, H, W, 1, generator=g) < 0.35).float().expand(B, 1, H, W, 1).contiguous()
    else:
        m = (torch.rand(B, 1, H, W, 1, generator=g) < 0.35).float()
    y = k * m.unsqueeze(1)
    return r2, s0, b0, ph, tes, y, S, m


def gen_qmri(R):
    """qRIM / qCIRIM: signal model, analytic gradient, qRIMBlock, model-level forward."""
    out = {}
    # --- signal model + analytic gradient, one sample (qrim/utils.py) ---
    i = 0
    for (E, C, H, W, mk, cen, nrm, seq) in [(4, 3, 12, 10, "1d", False, "backward", "MEGRE"),
                                            (4, 5, 9, 14, "2d", True, "ortho", "MEGRE"),
                                            (3, 2, 16, 6, "2d", False, "forward", "MEGRE"),
                                            (2, 3, 8, 8, "1d", True, "backward", "MEGRE_no_phase")]:
        r2, s0, b0, ph, tes, y, S, m = qmri_inputs(1, E, C, H, W, 700 + i, mk)
        fm = R.qrim_utils.SignalForwardModel(sequence=seq)
        sig = fm(r2, s0, b0, ph, tes)
        o_sig = oqnets.megre_signal(r2, s0, b0, ph, tes, no_phase=seq.lower() == "megre_no_phase")
        _close(o_sig, sig, "qmri signal", rtol=0, atol=0)
        g = R.qrim_utils.analytical_log_likelihood_gradient(fm, r2[0], s0[0], b0[0], ph[0], tes, S[0], y[0], m[0], cen,
                                                            nrm, [-2, -1], 2)
        o_g = oqnets.analytical_log_likelihood_gradient(r2[0], s0[0], b0[0], ph[0], tes, S[0], y[0], m[0], cen, nrm,
                                                        [-2, -1], 2, no_phase=seq.lower() == "megre_no_phase")
        _close(o_g, g, "qmri gradient", rtol=1e-6, atol=1e-6 * g.abs().max().item())
        out.update({"grad%d_%s" % (i, k): v for k, v in _np(dict(
            r2=r2, s0=s0, b0=b0, ph=ph, tes=np.asarray(tes), y=y, S=S, mask=m, signal=sig, grad=g)).items()})
        out["grad%d_cfg" % i] = np.asarray([int(cen), ["backward", "ortho", "forward"].index(nrm),
                                            int(seq.lower() == "megre_no_phase")])
        i += 1
    out["ngrad"] = np.asarray(i)
    # --- qRIMBlock (qrim/qrim_block.py) ---
    i = 0
    for (layer, B, mk, cen, nrm) in [("IndRNN", 2, "1d", False, "backward"), ("GRU", 1, "2d", True, "ortho"),
                                     ("MGU", 2, "2db", False, "ortho")]:
        hp = dict(QRIM_HP, recurrent_layer=layer, fft_centered=cen, fft_normalization=nrm, sequence="MEGRE")
        torch.manual_seed(40 + i)
        blk = R.qrim_block.qRIMBlock(
            recurrent_layer=layer, conv_filters=hp["conv_filters"], conv_kernels=hp["conv_kernels"],
            conv_dilations=hp["conv_dilations"], conv_bias=hp["conv_bias"], recurrent_filters=hp["recurrent_filters"],
            recurrent_kernels=hp["recurrent_kernels"], recurrent_dilations=hp["recurrent_dilations"],
            recurrent_bias=hp["recurrent_bias"], depth=2, time_steps=hp["time_steps"], conv_dim=2, no_dc=True,
            linear_forward_model=R.qrim_utils.SignalForwardModel(sequence="MEGRE"), fft_centered=cen,
            fft_normalization=nrm, spatial_dims=[-2, -1], coil_dim=2, coil_combination_method="SENSE",
            dimensionality=2).eval()
        sd = {k: v.detach().clone() for k, v in blk.state_dict().items()}
        r2, s0, b0, ph, tes, y, S, m = qmri_inputs(B, 4, 3, 14, 12, 800 + i, mk)
        gamma = torch.tensor(QGAMMA)
        maps = [r2 / gamma[0], s0 / gamma[1], b0 / gamma[2], ph / gamma[3]]
        with torch.no_grad():
            etas, _ = blk(y.clone(), y, maps[0], maps[1], maps[2], maps[3], tes, S, m, None, None, gamma, False)
            o_etas, _ = oqnets.qrim_block(sd, hp, y, maps[0], maps[1], maps[2], maps[3], tes, S, m, None, None, gamma)
        for a, b in zip(o_etas, etas):
            _close(a, b, "qrim %s step" % layer, rtol=1e-5, atol=1e-6)
        out.update({"blk%d_%s" % (i, k): v for k, v in _np(dict(
            r2=maps[0], s0=maps[1], b0=maps[2], ph=maps[3], tes=np.asarray(tes), y=y, S=S, mask=m, first=etas[0],
            last=etas[-1])).items()})
        out.update({"blk%d_w_%s" % (i, k): v.numpy() for k, v in sd.items()})
        out["blk%d_cfg" % i] = np.asarray([["GRU", "IndRNN", "MGU"].index(layer), int(cen),
                                           ["backward", "ortho", "forward"].index(nrm)])
        i += 1
    out["nblk"] = np.asarray(i)
    # --- model level: qCIRIM.forward glue (qcirim.py:247-341) restated around the REFERENCE blocks ---
    cfg = oqnets_cfg(num_cascades=2)
    torch.manual_seed(77)
    blocks = [R.qrim_block.qRIMBlock(
        recurrent_layer=cfg["quantitative_module_recurrent_layer"], conv_filters=cfg["quantitative_module_conv_filters"],
        conv_kernels=cfg["quantitative_module_conv_kernels"], conv_dilations=cfg["quantitative_module_conv_dilations"],
        conv_bias=cfg["quantitative_module_conv_bias"], recurrent_filters=cfg["quantitative_module_recurrent_filters"],
        recurrent_kernels=cfg["quantitative_module_recurrent_kernels"],
        recurrent_dilations=cfg["quantitative_module_recurrent_dilations"],
        recurrent_bias=cfg["quantitative_module_recurrent_bias"], depth=2,
        time_steps=cfg["quantitative_module_time_steps"], conv_dim=2, no_dc=True,
        linear_forward_model=R.qrim_utils.SignalForwardModel(sequence="MEGRE"), fft_centered=cfg["fft_centered"],
        fft_normalization=cfg["fft_normalization"], spatial_dims=[-2, -1], coil_dim=2, coil_combination_method="SENSE",
        dimensionality=2).eval() for _ in range(2)]
    sd = {}
    for ci, b in enumerate(blocks):
        sd.update({"qcirim.%d.%s" % (ci, k): v.detach().clone() for k, v in b.state_dict().items()})
    r2, s0, b0, ph, tes, y, S, m = qmri_inputs(2, 4, 4, 16, 12, 900, "2d")
    gamma = torch.tensor(QGAMMA)
    with torch.no_grad():
        # reference blocks driven by the reference's forward glue, transcribed line by line (qcirim.py:247-312)
        mp = [r2 / gamma[0], s0 / gamma[1], b0 / gamma[2], ph / gamma[3]]
        prediction = y.clone()
        ref_maps = [[], [], [], []]
        for ci, cascade in enumerate(blocks):
            prediction, _ = cascade(prediction, y, mp[0], mp[1], mp[2], mp[3], tes, S, m, None, None, gamma,
                                    keep_eta=ci != 0)
            mp = [prediction[-1][:, k] for k in range(4)]
            steps = [R.qrim_utils.RescaleByMax.reverse(torch.abs(p), gamma) for p in prediction]
            for k in range(4):
                ref_maps[k].append([s[:, k, ...] for s in steps])
        o = oqnets.qcirim_forward(sd, cfg, r2, s0, b0, ph, tes, y, S, torch.ones_like(m), m)
    for k in range(4):
        for ci in range(2):
            for a, b in zip(o[1 + k][ci], ref_maps[k][ci]):
                _close(a, b, "qcirim map %d" % k, rtol=1e-5, atol=1e-5)
    out.update({"model_%s" % k: v for k, v in _np(dict(
        r2=r2, s0=s0, b0=b0, ph=ph, tes=np.asarray(tes), y=y, S=S, mask=m,
        r2_last=ref_maps[0][-1][-1], s0_last=ref_maps[1][-1][-1], b0_last=ref_maps[2][-1][-1],
        ph_last=ref_maps[3][-1][-1], r2_first=ref_maps[0][0][0])).items()})
    out.update({"model_w_%s" % k: v.numpy() for k, v in sd.items()})
    np.savez_compressed(os.path.join(GOLDEN, "qmri.npz"), **out)
    print("qmri.npz: %d arrays" % len(out))


def oqnets_cfg(num_cascades=1, filters=16, layer="IndRNN", centered=False, normalization="backward"):
    """Flat qCIRIM cfg with the keys of projects/quantitative/model_zoo/conf/base_qcirim_run.yaml:7-121."""
    return dict(
        use_reconstruction_module=False, quantitative_module_recurrent_layer=layer,
        quantitative_module_conv_filters=[filters, filters, 4], quantitative_module_conv_kernels=[5, 3, 3],
        quantitative_module_conv_dilations=[1, 2, 1], quantitative_module_conv_bias=[True, True, False],
        quantitative_module_recurrent_filters=[filters, filters, 0], quantitative_module_recurrent_kernels=[1, 1, 0],
        quantitative_module_recurrent_dilations=[1, 1, 0], quantitative_module_recurrent_bias=[True, True, False],
        quantitative_module_depth=2, quantitative_module_time_steps=8, quantitative_module_conv_dim=2,
        quantitative_module_num_cascades=num_cascades, quantitative_module_no_dc=True,
        quantitative_module_keep_eta=True, quantitative_module_accumulate_estimates=True,
        quantitative_module_signal_forward_model_sequence="MEGRE", quantitative_module_dimensionality=2,
        quantitative_module_gamma_regularization_factors=list(QGAMMA), shift_B0_input=False, dimensionality=2,
        coil_combination_method="SENSE", use_sens_net=False, fft_centered=centered, fft_normalization=normalization,
        spatial_dims=[-2, -1], coil_dim=2)


def gen_poisson(R):
    """BASELINE.json configs[4]: 12x Poisson-disc 2-D mask.  The reference draws it from numba's unseeded generator
    (subsample.py:584-600), so it is generated ONCE here with the reference function and cached bit-packed; both
    implementations are always fed this tensor."""
    H, W = 232, 288
    np.random.seed(123)
    mf = R.subsample.Poisson2DMaskFunc([0.7], [12])
    m, acc = mf((1, H, W, 2), seed=123)
    m = m.numpy().reshape(H, W)
    assert set(np.unique(m)) <= {0.0, 1.0}
    np.savez_compressed(os.path.join(GOLDEN, "poisson_mask.npz"), shape=np.asarray([H, W]), acc=np.asarray(acc),
                        bits=np.packbits(m.astype(np.uint8)))
    print("poisson_mask.npz: %dx%d, acceleration %s, sampled fraction %.4f" % (H, W, acc, m.mean()))


def _ref_sens_model(R):
    """BaseSensitivityModel lives in reconstruction/models/base.py, whose module imports need Lightning: run the class
    definition alone (oracle/ref_import.py::ref_class_from_source) against the imported reference leaf modules."""
    from abc import ABC
    from typing import Optional, Sequence, Tuple

    from .ref_import import ref_class_from_source

    ns = dict(nn=torch.nn, torch=torch, ABC=ABC, Optional=Optional, Sequence=Sequence, Tuple=Tuple, utils=R.utils,
              fft=R.fft, unet_block=R.unet_block)
    return ref_class_from_source("mridc/collections/reconstruction/models/base.py", "BaseSensitivityModel", ns)


def gen_sens(R):
    """Sensitivity-estimation network, reference class executed from its own source."""
    Sens = _ref_sens_model(R)
    out = {}
    i = 0
    for (B, C, H, W, chans, pools, mtype, cen, nrm, normalize, mcenter, nlf) in [
            (2, 3, 20, 24, 4, 2, "2D", True, "ortho", True, True, None),
            (1, 4, 18, 30, 4, 3, "1D", False, "backward", True, True, None),
            (1, 2, 16, 16, 4, 2, "1D", True, "ortho", True, True, 6),
            (1, 3, 17, 21, 4, 2, "2D", False, "backward", False, False, None)]:
        torch.manual_seed(60 + i)
        net = Sens(chans, pools, fft_centered=cen, fft_normalization=nrm, spatial_dims=[-2, -1], coil_dim=1,
                   mask_type=mtype, normalize=normalize, mask_center=mcenter).eval()
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        g = torch.Generator().manual_seed(600 + i)
        y = torch.randn(B, C, H, W, 2, generator=g)
        # per-sample column masks with a fully sampled centre band of different widths
        m = (torch.rand(B, 1, 1, W, 1, generator=g) < 0.3).float()
        for b in range(B):
            half = 2 + b
            m[b, 0, 0, W // 2 - half: W // 2 + half, 0] = 1
            m[b, 0, 0, W // 2 - half - 1, 0] = 0
            m[b, 0, 0, W // 2 + half, 0] = 0
        y = y * m
        with torch.no_grad():
            ref = net(y, m, nlf)
            hp = dict(sens_pools=pools, padding_size=15, sens_mask_type=mtype, sens_normalize=normalize,
                      sens_mask_center=mcenter, fft_centered=cen, fft_normalization=nrm, spatial_dims=[-2, -1], coil_dim=1)
            o = onets.sensitivity_model(sd, hp, y, m, nlf)
        _close(o, ref, "sens model %d" % i, rtol=1e-5, atol=1e-6)
        pad, n = Sens.get_pad_and_num_low_freqs(m, nlf)
        out.update({"sens%d_%s" % (i, k): v for k, v in _np(dict(y=y, mask=m, out=ref, pad=pad, nlf=n)).items()})
        out.update({"sens%d_w_%s" % (i, k): v.numpy() for k, v in sd.items()})
        out["sens%d_cfg" % i] = np.asarray([chans, pools, ["1D", "2D"].index(mtype), int(cen),
                                            ["backward", "ortho", "forward"].index(nrm), int(normalize), int(mcenter),
                                            -1 if nlf is None else nlf])
        i += 1
    out["nsens"] = np.asarray(i)
    np.savez_compressed(os.path.join(GOLDEN, "sens.npz"), **out)
    print("sens.npz: %d arrays" % len(out))


def gen_apply_mask(R):
    """common/parts/utils.py:293-343: the reference's apply_mask on seeded k-space with its own mask functions
    (seeded 1-D equispaced / random, padding zeroing, fftshift of a 2-D mask); mridc_b200.utils.apply_mask must
    reproduce data, mask and acceleration bit for bit (tests/test_host_logic.py, tests/test_gpu_parity.py)."""
    from mridc_b200 import synth, utils as mutils

    out = {}
    g = torch.Generator().manual_seed(11)
    cases = [
        ("equi", R.subsample.Equispaced1DMaskFunc, synth.Equispaced1DMask, [0.08], [4], (3, 24, 40, 2), 123, None, False),
        ("rand_pad", R.subsample.RandomMaskFunc, synth.RandomMask1D, [0.08], [4], (2, 3, 16, 48, 2), (1, 2, 3), (6, 40), False),
        ("equi2d_shift", R.subsample.Equispaced2DMaskFunc, None, [0.08], [4], (2, 30, 36, 2), 5, None, True),
        ("rand_pad0", R.subsample.RandomMaskFunc, synth.RandomMask1D, [0.04], [8], (1, 12, 64, 2), 9, (0, 50), False),
    ]
    for i, (name, rcls, mcls, cf, acc, shape, seed, padding, shift) in enumerate(cases):
        data = torch.randn(*shape, generator=g)
        data[..., 0, 0, :] = -0.0  # the "+ 0.0" of utils.py:341 must clear the sign of zeros
        rd, rm, ra = R.utils.apply_mask(data.clone(), rcls(cf, acc), seed=seed, padding=padding, shift=shift)
        if mcls is None:
            # 2-D mask functions are not restated in the package (a23 covers the 1-D ones + Gaussian + the cached
            # Poisson mask): replay the reference's raw mask through a fixed mask function
            shp = np.array(data.shape); shp[:-3] = 1
            raw, racc = rcls(cf, acc)(shp, seed, half_scan_percentage=0.0, scale=0.02)
            out["am%d_raw" % i] = raw.numpy()

            def fixed(shape, seed, half_scan_percentage=0.0, scale=0.02, _m=raw, _a=racc):
                return _m.clone(), _a

            md, mm, ma = mutils.apply_mask(data.clone(), fixed, seed=seed, padding=padding, shift=shift)
        else:
            md, mm, ma = mutils.apply_mask(data.clone(), mcls(cf, acc), seed=seed, padding=padding, shift=shift)
        assert torch.equal(rd, md) and torch.equal(rm, mm) and ra == ma, name
        assert torch.equal(torch.signbit(rd), torch.signbit(md)), name
        out.update({"am%d_in" % i: data.numpy(), "am%d_out" % i: rd.numpy(), "am%d_mask" % i: rm.numpy(),
                    "am%d_acc" % i: np.asarray(ra), "am%d_seed" % i: np.asarray(seed),
                    "am%d_pad" % i: np.asarray(padding if padding is not None else (-1, -1)),
                    "am%d_cfg" % i: np.asarray([["equi", "rand_pad", "equi2d_shift", "rand_pad0"].index(name), int(shift)]),
                    "am%d_cf" % i: np.asarray(cf), "am%d_accs" % i: np.asarray(acc)})
    out["nam"] = np.asarray(len(cases))
    np.savez_compressed(os.path.join(GOLDEN, "apply_mask.npz"), **out)
    print("apply_mask: %d cases" % len(cases))


def gen_consumers(R):
    """The other consumers of the DC operator: sigmanet DC layers in the sigmanet layout
    ([B, C, sets, H, W, 2] maps) and in DUNet's ([1, C, H, W, 2] maps, dunet.py:177-186), CascadeNetBlock,
    RecurrentInit / RecurrentVarNetBlock, qVarNetBlock.  Reference modules run unmodified; the oracle restatement
    (oracle/consumers.py) is checked against them."""
    from . import consumers as oc

    out = {}
    g = torch.Generator().manual_seed(77)
    rn = lambda *s: torch.randn(*s, generator=g)
    sd_ = [-2, -1]
    # ---- sigmanet layers, sigmanet layout: x [B, sets, H, W, 2], y [B, C, 1, H, W, 2], smaps [B, C, sets, H, W, 2]
    i = 0
    for B, C, S, H, W, cen, nrm in [(2, 3, 1, 12, 10, True, "ortho"), (1, 4, 2, 9, 14, False, "backward")]:
        x, smaps = rn(B, S, H, W, 2), rn(B, C, S, H, W, 2) * 0.5
        mask = (torch.rand(1, 1, 1, 1, W, 1, generator=g) < 0.45).float()
        y = rn(B, C, 1, H, W, 2) * mask
        gd = R.dc_layers.DataGDLayer(0.7, fft_centered=cen, fft_normalization=nrm, spatial_dims=sd_)
        vs = R.dc_layers.DataVSLayer(0.3, 0.6, fft_centered=cen, fft_normalization=nrm, spatial_dims=sd_)
        with torch.no_grad():
            r_gd, r_vs = gd(x, y, smaps, mask), vs(x, y, smaps, mask)
        _close(oc.data_gd(x, y, smaps, mask, 0.7, cen, nrm, sd_), r_gd, "DataGDLayer %d" % i)
        _close(oc.data_vs(x, y, smaps, mask, 0.3, 0.6, cen, nrm, sd_), r_vs, "DataVSLayer %d" % i)
        case = dict(x=x, y=y, smaps=smaps, mask=mask, gd=r_gd, vs=r_vs)
        if B == 1:  # the prox layer broadcasts x over the coil axis with expand_as: batch size 1 upstream
            cg = R.dc_layers.DataProxCGLayer(0.5, tol=1e-6, iter=6, fft_centered=cen, fft_normalization=nrm, spatial_dims=sd_)
            with torch.no_grad():
                r_cg = cg(x, y, smaps, mask)
            _close(oc.prox_cg(x, 0.5, y, smaps, mask, 1e-6, 6, cen, nrm, sd_), r_cg, "DataProxCGLayer %d" % i, rtol=1e-5, atol=1e-6)
            case["cg"] = r_cg
        out.update({"sig%d_%s" % (i, k): v for k, v in _np(case).items()})
        out["sig%d_cfg" % i] = np.asarray([int(cen), ["backward", "ortho", "forward"].index(nrm)])
        i += 1
    out["nsig"] = np.asarray(i)
    # ---- DUNet layout, two iterations of the gradient layer (the image becomes per-coil after the first, dunet.py:186)
    C, H, W = 3, 10, 12
    x, smaps = rn(1, H, W, 2), rn(1, C, H, W, 2) * 0.5
    mask = (torch.rand(1, 1, 1, W, 1, generator=g) < 0.5).float()
    y = rn(1, C, H, W, 2) * mask
    gd = R.dc_layers.DataGDLayer(0.4, fft_centered=True, fft_normalization="ortho", spatial_dims=sd_)
    vs = R.dc_layers.DataVSLayer(0.2, 0.5, fft_centered=True, fft_normalization="ortho", spatial_dims=sd_)
    with torch.no_grad():
        x1 = gd(x, y, smaps, mask)
        x2 = gd(x1, y, smaps, mask)
        v1 = vs(x, y, smaps, mask)
    _close(oc.data_gd(oc.data_gd(x, y, smaps, mask, 0.4, True, "ortho", sd_), y, smaps, mask, 0.4, True, "ortho", sd_), x2, "GD dunet")
    _close(oc.data_vs(x, y, smaps, mask, 0.2, 0.5, True, "ortho", sd_), v1, "VS dunet")
    out.update({"dun_" + k: v for k, v in _np(dict(x=x, y=y, smaps=smaps, mask=mask, gd1=x1, gd2=x2, vs=v1)).items()})
    # ---- DCLayer (single coil)
    x, mask = rn(2, 1, 8, 10, 2), (torch.rand(2, 1, 8, 10, 1, generator=g) < 0.4).float()
    y = rn(2, 1, 8, 10, 2) * mask
    dl = R.dc_layers.DCLayer(0.25, fft_centered=False, fft_normalization="ortho", spatial_dims=sd_)
    with torch.no_grad():
        r = dl(x, y, mask)
    _close(oc.dc_layer(x, y, mask, 0.25, False, "ortho", sd_), r, "DCLayer")
    out.update({"dcl_" + k: v for k, v in _np(dict(x=x, y=y, mask=mask, out=r)).items()})
    # ---- CascadeNetBlock around a two-conv regulariser
    torch.manual_seed(5)
    reg = torch.nn.Sequential(torch.nn.Conv2d(2, 8, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(8, 2, 3, padding=1))
    i = 0
    for no_dc, md in [(False, torch.float32), (True, torch.uint8)]:
        y, S, _, m = small_inputs(2, 3, 12, 10, 300 + i, "1d", md)
        pred = y + 0.1 * rn(*y.shape)
        blk = R.ccnn_block.CascadeNetBlock(reg, True, "ortho", sd_, 1, no_dc)
        blk.dc_weight.data.fill_(0.8)
        with torch.no_grad():
            r = blk(pred, y, S, m)
            _close(oc.cascadenet_block(reg, blk.dc_weight, pred, y, S, m, True, "ortho", sd_, 1, no_dc), r, "CascadeNetBlock %d" % i)
        out.update({"ccnn%d_%s" % (i, k): v for k, v in _np(dict(pred=pred, y=y, S=S, mask=m, out=r)).items()})
        i += 1
    out.update({"ccnn_w_" + k.replace(".", "_"): v.numpy() for k, v in reg.state_dict().items()})
    # ---- RecurrentInit + two RecurrentVarNetBlock steps (hidden state carried)
    torch.manual_seed(6)
    init = R.recurrentvarnet.RecurrentInit(2, 8, (8, 8), (1, 2), depth=2, multiscale_depth=2).eval()
    blk = R.recurrentvarnet.RecurrentVarNetBlock(2, 8, 2, True, "ortho", sd_, 1).eval()
    blk.learning_rate.data.fill_(0.9)
    y, S, _, m = small_inputs(2, 3, 12, 10, 310, "1d", torch.float32)
    cur = y + 0.1 * rn(*y.shape)
    with torch.no_grad():
        img0 = R.utils.complex_mul(R.fft.ifft2(y, True, "ortho", sd_), R.utils.complex_conj(S)).sum(1).permute(0, 3, 1, 2)
        h0 = init(img0)
        k1, h1 = blk(cur, y, m, S, h0)
        k2, h2 = blk(k1, y, m, S, h1)
        k1n, h1n = blk(cur, y, m, S, None)
    isd, bsd = init.state_dict(), blk.state_dict()
    _close(oc.recurrent_init(isd, (1, 2), 2, 2, img0), h0, "RecurrentInit")
    o1, oh1 = oc.recurrentvarnet_block(bsd, 2, 8, cur, y, m, S, h0, True, "ortho", sd_)
    o2, oh2 = oc.recurrentvarnet_block(bsd, 2, 8, o1, y, m, S, oh1, True, "ortho", sd_)
    _close(o2, k2, "RecurrentVarNetBlock k-space", rtol=1e-5, atol=1e-6)
    _close(oh2, h2, "RecurrentVarNetBlock state", rtol=1e-5, atol=1e-6)
    out.update({"rvn_" + k: v for k, v in _np(dict(y=y, S=S, mask=m, cur=cur, img0=img0, h0=h0, k1=k1, h1=h1, k2=k2, h2=h2,
                                                  k1n=k1n, h1n=h1n)).items()})
    out.update({"rvn_init_" + k.replace(".", "_"): v.numpy() for k, v in isd.items()})
    out.update({"rvn_blk_" + k.replace(".", "_"): v.numpy() for k, v in bsd.items()})
    # ---- qVarNetBlock (batch 1, 4 echoes, coil_dim 2) around a NormUnet(8 -> 8 channels)
    torch.manual_seed(7)
    E, C, H, W = 4, 3, 16, 12
    unet = R.unet_block.NormUnet(chans=4, num_pools=2, in_chans=2 * E, out_chans=2 * E, padding_size=3, normalize=True)
    qb = R.qvn_block.qVarNetBlock(unet, True, "ortho", sd_, 2, False).eval()
    qb.dc_weight.data.fill_(0.7)
    maps = [rn(1, H, W).abs() * s for s in (30.0, 1.0, 10.0, 0.5)]
    maps[0][0, :2] = -1.0  # negative R2* estimates are clipped on the way out (:156-158)
    S = rn(1, C, H, W, 2) * 0.5
    sm = (torch.rand(1, 1, 1, H, W, 1, generator=g) < 0.4).float()
    yq = rn(1, E, C, H, W, 2) * sm
    TEs = [3.0, 11.5, 20.0, 28.5]
    gamma = torch.tensor([150.0, 150.0, 1000.0, 150.0])
    with torch.no_grad():
        r = qb(yq.clone(), yq, *[m_ / gamma[j] for j, m_ in enumerate(maps)], TEs, S, sm, gamma)
        o = oc.qvarnet_block({k[len("model."):]: v for k, v in qb.state_dict().items() if k.startswith("model.")},
                             dict(num_pools=2, padding_size=3, normalize=True), qb.dc_weight, yq,
                             *[m_ / gamma[j] for j, m_ in enumerate(maps)], TEs, S, sm, gamma, True, "ortho", sd_, 2)
    _close(o, r, "qVarNetBlock", rtol=1e-5, atol=1e-6)
    out.update({"qvn_" + k: v for k, v in _np(dict(y=yq, S=S, mask=sm, gamma=gamma, R2=maps[0] / gamma[0], S0=maps[1] / gamma[1],
                                                  B0=maps[2] / gamma[2], phi=maps[3] / gamma[3], out=r)).items()})
    out.update({"qvn_w_" + k.replace(".", "_"): v.numpy() for k, v in qb.state_dict().items()})
    np.savez_compressed(os.path.join(GOLDEN, "consumers.npz"), **out)
    print("consumers.npz: %d arrays" % len(out))


TRANSFORM_CASES = [
    # name, ctor keyword arguments (mask functions by name), extra inputs
    ("sense_norm", dict(coil_combination_method="SENSE", mask_func=["equi"], normalize_inputs=True, fft_centered=True,
                        fft_normalization="ortho", coil_dim=1), {}),
    ("rss_crop", dict(coil_combination_method="RSS", mask_func=["rand", "equi"], crop_size=(10, 8), normalize_inputs=False,
                      fft_centered=False, fft_normalization="backward", coil_dim=1), {}),
    # cropping after masking needs a 2-D mask upstream (center_crop of the squeezed mask, :524): a precomputed one
    ("kcrop_after", dict(coil_combination_method="SENSE", mask_func=None, crop_size=(12, 10), kspace_crop=True,
                         crop_before_masking=False, normalize_inputs=True, fft_centered=True, fft_normalization="backward",
                         coil_dim=1), {"eta": True, "mask": True}),
    # a tuple of mask functions takes the single-mask branch (:485-496)
    ("single_func", dict(coil_combination_method="SENSE", mask_func=("equi",), crop_size=(12, 10), kspace_crop=False,
                         normalize_inputs=True, fft_centered=True, fft_normalization="ortho", coil_dim=1), {"eta": True}),
    ("zero_fill_full", dict(coil_combination_method="SENSE", mask_func=None, kspace_zero_filling_size=(20, 18),
                            normalize_inputs=True, fft_centered=True, fft_normalization="ortho", coil_dim=1), {}),
    ("given_mask", dict(coil_combination_method="SENSE", mask_func=None, shift_mask=True, normalize_inputs=True,
                        fft_centered=False, fft_normalization="forward", coil_dim=1), {"mask": True}),
    ("prewhiten", dict(apply_prewhitening=True, prewhitening_scale_factor=1.5, prewhitening_patch_start=2,
                       prewhitening_patch_length=6, coil_combination_method="SENSE", mask_func=["equi"], normalize_inputs=True,
                       fft_centered=True, fft_normalization="ortho", coil_dim=1), {}),
    ("none_norm", dict(coil_combination_method="RSS", mask_func=["equi"], normalize_inputs=True, fft_centered=False,
                       fft_normalization="none", coil_dim=1), {}),
]


def transform_inputs(seed, C=4, H=16, W=14):
    rng = np.random.RandomState(seed)
    k = (rng.randn(C, H, W) + 1j * rng.randn(C, H, W)).astype(np.complex64)
    S = (rng.randn(C, H, W) + 1j * rng.randn(C, H, W)).astype(np.complex64) * 0.5
    eta = (rng.randn(H, W) + 1j * rng.randn(H, W)).astype(np.complex64)
    m = (rng.rand(H, W) < 0.4).astype(np.float32)
    return k, S, eta, m


def gen_transforms(R):
    """MRIDataTransforms (reconstruction/parts/transforms.py:155-619) run unmodified on small slices: every output of the
    9-tuple is stored.  Mask functions are the reference's; the product is fed the bit-identical restatements."""
    sub = R.subsample
    mk = {"equi": lambda: sub.Equispaced1DMaskFunc([0.08], [4]), "rand": lambda: sub.RandomMaskFunc([0.1], [3])}
    out = {}
    for idx, (name, kw, extra) in enumerate(TRANSFORM_CASES):
        kw = dict(kw)
        if kw.get("mask_func"):
            kw["mask_func"] = type(kw["mask_func"])(mk[n]() for n in kw["mask_func"])
        tr = R.transforms.MRIDataTransforms(**kw)
        k, S, eta, m = transform_inputs(500 + idx)
        res = tr(k, S, [m] if extra.get("mask") else None, eta if extra.get("eta") else None, None, {}, "file%d.h5" % idx, 3)
        kspace, masked, sens, mask, eta_o, target, fname, sl, acc = res
        d = dict(kspace=kspace, sens=sens, target=target)
        if isinstance(masked, list):
            for j, (y_, m_) in enumerate(zip(masked, mask)):
                d["masked%d" % j], d["mask%d" % j] = y_, m_
            d["acc"] = np.asarray([float(a) for a in acc])
        else:
            d["masked0"], d["mask0"] = masked, mask
            d["acc"] = np.asarray([float(torch.as_tensor(acc).reshape(-1)[0])])
        if extra.get("eta"):
            d["eta"] = eta_o
        out.update({"tr%d_%s" % (idx, k_): v for k_, v in _np(d).items()})
        assert fname == "file%d.h5" % idx and sl == 3
    np.savez_compressed(os.path.join(GOLDEN, "transforms.npz"), **out)
    print("transforms.npz: %d arrays" % len(out))


JRS_RIM = dict(recurrent_layer="GRU", conv_filters=[8, 8, 2], conv_kernels=[5, 3, 3], conv_dilations=[1, 2, 1],
               conv_bias=[True, True, False], recurrent_filters=[8, 8, 0], recurrent_kernels=[1, 1, 0],
               recurrent_dilations=[1, 1, 0], recurrent_bias=[True, True, False], depth=2, time_steps=8, conv_dim=2,
               num_cascades=2, no_dc=True, keep_eta=True, dimensionality=2, pretrained=False, accumulate_estimates=True)
JRS_CASES = [
    # name, segmentation params, input channels, magnitude input, consecutive slices
    ("unet_mag", dict(segmentation_module="UNet", output_channels=3, channels=4, pooling_layers=2, dropout=0.0), 1, True, 1),
    ("conv_cplx", dict(segmentation_module="ConvLayer", output_channels=2, conv_dim=2), 2, False, 1),
    ("unet_slices", dict(segmentation_module="UNet", output_channels=2, channels=4, pooling_layers=1, dropout=0.0), 1, True, 2),
]


def jrs_inputs(idx, slices):
    g = torch.Generator().manual_seed(900 + idx)
    B, C, H, W = 2, 3, 16, 12
    lead = (B, slices) if slices > 1 else (B,)
    y = torch.randn(*lead, C, H, W, 2, generator=g)
    S = torch.randn(*lead, C, H, W, 2, generator=g) * 0.5
    m = (torch.rand(1, 1, 1, W, 1, generator=g) < 0.45).float()
    m[..., W // 2, :] = 1
    m = m.reshape(1, 1, 1, 1, W, 1) if slices > 1 else m
    y = y * m
    init = torch.randn(*lead, H, W, 2, generator=g)
    target = torch.randn(*lead, H, W, 2, generator=g)
    return y, S, m, init, target


def gen_jrscirim(R):
    """JRSCIRIMBlock (segmentation/models/jrscirim_base/jrscirim_block.py) run unmodified: CIRIM cascades + UNet / ConvLayer
    segmentation heads, single slices and the per-slice loop of consecutive_slices = 2."""
    from . import consumers as oc

    out = {}
    for idx, (name, sp, in_ch, mag, slices) in enumerate(JRS_CASES):
        torch.manual_seed(40 + idx)
        blk = R.jrscirim_block.JRSCIRIMBlock(dict(JRS_RIM), dict(sp), in_ch, mag, True, "ortho", [-2, -1], 2, 2, slices,
                                             "SENSE", True).eval()
        y, S, m, init, target = jrs_inputs(idx, slices)
        use_init = idx == 1  # one case starts every cascade from a given image (rim_block.py:195)
        no_init = torch.zeros(2, slices) if slices > 1 else torch.zeros(1)  # < 4-D: ignored (:243-247, :277-281)
        with torch.no_grad():
            rec, seg, _ = blk(y, S, m, init if use_init else no_init, target)
            hp = dict(JRS_RIM, fft_centered=True, fft_normalization="ortho", spatial_dims=[-2, -1], coil_dim=1)
            orec, oseg = oc.jrscirim_block(blk.state_dict(), hp, 2, True, "unet" if "UNet" in sp["segmentation_module"] else "conv",
                                           sp, in_ch, mag, slices, y, S, m, init if use_init else no_init, target)
        _close(torch.stack([torch.stack(c) for c in orec]), torch.stack([torch.stack(c) for c in rec]), "JRSCIRIM rec " + name,
               rtol=1e-5, atol=1e-6)
        _close(oseg, seg, "JRSCIRIM seg " + name, rtol=1e-4, atol=1e-5)
        d = dict(y=y, S=S, mask=m, init=init, target=target, rec=torch.stack([torch.stack(c) for c in rec]), seg=seg)
        out.update({"jrs%d_%s" % (idx, k): v for k, v in _np(d).items()})
        out.update({"jrs%d_w_%s" % (idx, k.replace(".", "_")): v.numpy() for k, v in blk.state_dict().items()})
    np.savez_compressed(os.path.join(GOLDEN, "jrscirim.npz"), **out)
    print("jrscirim.npz: %d arrays" % len(out))


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLDEN, exist_ok=True)
    R = Ref()
    only = set(sys.argv[1:])  # e.g. `python -m oracle.make_golden qmri` regenerates one fixture file
    for name, fn in (("masks", gen_masks), ("prims", gen_prims), ("dc", gen_dc), ("rim", gen_rim), ("rim3d", gen_rim3d),
                     ("unet", gen_unet),
                     ("models", gen_models), ("qmri", gen_qmri), ("poisson", gen_poisson), ("sens", gen_sens),
                     ("apply_mask", gen_apply_mask), ("consumers", gen_consumers), ("transforms", gen_transforms),
                     ("jrscirim", gen_jrscirim)):
        if not only or name in only:
            fn(R)
    tot = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    print("golden fixtures total %.1f KB" % (tot / 1024))


if __name__ == "__main__":
    sys.exit(main())

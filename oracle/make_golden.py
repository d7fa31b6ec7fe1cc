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


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(GOLDEN, exist_ok=True)
    R = Ref()
    gen_masks(R)
    gen_prims(R)
    gen_dc(R)
    gen_rim(R)
    gen_unet(R)
    gen_models(R)
    tot = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    print("golden fixtures total %.1f KB" % (tot / 1024))


if __name__ == "__main__":
    sys.exit(main())

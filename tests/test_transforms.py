"""MRIDataTransforms on the device against the outputs of the unmodified reference class
(tests/golden/transforms.npz, oracle/make_golden.py::gen_transforms): every element of the 9-tuple, for SENSE / RSS
targets, image- and k-space crops before / after masking, zero filling, noise pre-whitening, fully sampled data, precomputed and generated
masks, all normalisation modes.

CPU part: the transform's host logic with the CUDA wrappers swapped for the oracle's CPU functions (test scaffolding; the
package has no CPU path), and the output assembly.  GPU part (``-m gpu``): the same cases through the C-ABI.
Tolerance: rel-L2 <= 2e-6 per tensor (two FFT round trips); masks and acceleration factors bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import mri as omri
from oracle.make_golden import TRANSFORM_CASES, transform_inputs


def _mask_funcs():
    from mridc_b200 import synth

    return {"equi": lambda: synth.Equispaced1DMask([0.08], [4]), "rand": lambda: synth.RandomMask1D([0.1], [3])}


def _run_case(mb, idx, name, kw, extra):
    kw = dict(kw)
    mk = _mask_funcs()
    if kw.get("mask_func"):
        kw["mask_func"] = type(kw["mask_func"])(mk[n]() for n in kw["mask_func"])
    tr = mb.MRIDataTransforms(**kw)
    k, S, eta, m = transform_inputs(500 + idx)
    return tr(k, S, [m] if extra.get("mask") else None, eta if extra.get("eta") else None, None, {}, "file%d.h5" % idx, 3)


def _check_case(g, idx, res, extra, tol):
    kspace, masked, sens, mask, eta, target, fname, sl, acc = res
    assert fname == "file%d.h5" % idx and sl == 3
    p = "tr%d_" % idx
    for key, val in (("kspace", kspace), ("sens", sens), ("target", target)):
        assert tuple(val.shape) == g[p + key].shape, (idx, key)
        assert rel_l2(val, g[p + key]) < tol, (idx, key, rel_l2(val, g[p + key]))
    masked = masked if isinstance(masked, list) else [masked]
    mask = mask if isinstance(mask, list) else [mask]
    acc = acc if isinstance(acc, list) else [acc]
    for j, (y, mm, a) in enumerate(zip(masked, mask, acc)):
        assert np.array_equal(mm.cpu().numpy(), g[p + "mask%d" % j]) and mm.dtype == torch.from_numpy(g[p + "mask%d" % j]).dtype
        assert rel_l2(y, g[p + "masked%d" % j]) < tol, (idx, j)
        assert float(torch.as_tensor(a).reshape(-1)[0]) == float(g[p + "acc"][j])
    if extra.get("eta"):
        assert rel_l2(eta, g[p + "eta"]) < tol


class _OracleUtils:
    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):  # host-side helpers (to_tensor, crops, apply_mask, is_none) are the package's own
        return getattr(self._real, name)

    rss = staticmethod(omri.rss)
    sense = staticmethod(omri.sense)
    complex_abs = staticmethod(omri.complex_abs)


class _OracleOps:
    PAD_ZERO = 0

    @staticmethod
    def conv2d(x, weight, bias, k, dil, pad_mode, **kw):
        return torch.nn.functional.conv2d(x, weight, bias)


def test_host_logic_against_reference_vectors(golden, monkeypatch):
    import mridc_b200 as mb
    import mridc_b200.transforms as tm

    monkeypatch.setattr(tm, "_ops", _OracleOps)
    monkeypatch.setattr(tm, "fft", omri)
    monkeypatch.setattr(tm, "utils", _OracleUtils(tm.utils))
    monkeypatch.setattr(tm, "_device", lambda: torch.device("cpu"))
    g = golden("transforms")
    for idx, (name, kw, extra) in enumerate(TRANSFORM_CASES):
        _check_case(g, idx, _run_case(mb, idx, name, kw, extra), extra, 1e-6)


def test_unbuilt_options_raise():
    import mridc_b200 as mb

    with pytest.raises(NotImplementedError):
        mb.MRIDataTransforms(apply_gcc=True)
    with pytest.raises(NotImplementedError):
        mb.MRIDataTransforms(dimensionality=3)


def test_assemble_reconstructions_orders_slices(tmp_path):
    """models/base.py:576-582: slices sorted per file and stacked; save_reconstructions writes `reconstruction`."""
    import mridc_b200 as mb

    outs = [("b.h5", 1, np.full((2, 3), 11.0)), ("a.h5", 2, torch.full((2, 3), 2.0)), ("b.h5", 0, np.full((2, 3), 10.0)),
            ("a.h5", 0, np.full((2, 3), 0.0)), ("a.h5", 1, np.full((2, 3), 1.0))]
    rec = mb.assemble_reconstructions(outs)
    assert sorted(rec) == ["a.h5", "b.h5"] and rec["a.h5"].shape == (3, 2, 3) and rec["b.h5"].shape == (2, 2, 3)
    assert [float(s[0, 0]) for s in rec["a.h5"]] == [0.0, 1.0, 2.0] and [float(s[0, 0]) for s in rec["b.h5"]] == [10.0, 11.0]
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is None or not hasattr(h5py, "File"):  # absent, or the oracle's stub module
        with pytest.raises(ImportError, match="h5py"):
            mb.save_reconstructions(rec, tmp_path)
        return
    mb.save_reconstructions(rec, tmp_path)
    with h5py.File(tmp_path / "a.h5") as hf:
        assert np.array_equal(hf["reconstruction"][()], rec["a.h5"])


@pytest.mark.gpu
def test_device_transforms_against_reference_vectors(golden):
    import mridc_b200 as mb

    g = golden("transforms")
    for idx, (name, kw, extra) in enumerate(TRANSFORM_CASES):
        res = _run_case(mb, idx, name, kw, extra)
        assert res[0].is_cuda and res[5].is_cuda
        _check_case(g, idx, res, extra, 2e-6)


@pytest.mark.gpu
def test_device_transforms_feed_cirim_full_size():
    """15 x 320 x 320: the transform's outputs go straight into CIRIM.forward (no host round trip) and the masked k-space
    is consistent with the mask (off the sampled columns only the rounding of the max-normalisation round trip is left)."""
    import mridc_b200 as mb
    from mridc_b200 import synth

    C, H, W = 15, 320, 320
    k = synth._fft2c(synth.coil_maps(C, H, W) * synth.phantom(H, W)[None], True, "ortho").astype(np.complex64)
    tr = mb.MRIDataTransforms(coil_combination_method="SENSE", mask_func=[synth.Equispaced1DMask([0.08], [4])],
                              normalize_inputs=True, fft_centered=True, fft_normalization="ortho", coil_dim=1)
    kspace, masked, sens, mask, eta, target, *_ = tr(k, synth.coil_maps(C, H, W).astype(np.complex64), None, None, None, {},
                                                     "file.h5", 0)
    y, m = masked[0], mask[0]
    assert y.is_cuda and m.dtype == torch.uint8 and float(target.max()) == 1.0
    assert float((y * (1 - m.float())).abs().max()) < 1e-5 * float(y.abs().max())
    model = mb.CIRIM(synth.cirim_cfg(num_cascades=1, centered=True, normalization="ortho")).cuda()
    out = next(model.forward(y.unsqueeze(0), sens.unsqueeze(0), m.unsqueeze(0), None, target.unsqueeze(0)))
    assert out[-1][-1].shape == (1, H, W) and torch.isfinite(torch.view_as_real(out[-1][-1])).all()

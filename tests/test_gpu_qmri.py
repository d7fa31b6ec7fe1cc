"""GPU parity of the quantitative path (BASELINE.json configs[4]): MEGRE signal model, analytic
log-likelihood gradient, qRIMBlock and qCIRIM.forward through the C-ABI against (1) the committed reference vectors
(tests/golden/qmri.npz, produced by the unmodified reference modules), (2) the CPU oracle on seeded inputs,
(3) properties at the configs[4] size (32 coils x 4 echoes, 232 x 288, cached 12x Poisson-disc mask).

Tolerances: pointwise signal model rel-L2 <= 1e-6 (libdevice exp / sin / cos vs the host libm), gradient <= 5e-6
(fused DC operator + pointwise epilogue), blocks <= 1e-5 on eta after 8 steps, model maps <= 1e-4.
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

NRM3 = ["backward", "ortho", "forward"]
LAYERS = ["GRU", "IndRNN", "MGU"]
QGAMMA = [150.0, 150.0, 1000.0, 150.0]
QRIM_KW = dict(conv_filters=[16, 16, 4], conv_kernels=[5, 3, 3], conv_dilations=[1, 2, 1], conv_bias=[True, True, False],
               recurrent_filters=[16, 16, 0], recurrent_kernels=[1, 1, 0], recurrent_dilations=[1, 1, 0],
               recurrent_bias=[True, True, False], depth=2, time_steps=8, conv_dim=2, no_dc=True, spatial_dims=[-2, -1],
               coil_dim=2, coil_combination_method="SENSE", dimensionality=2)


def cu(a):
    t = torch.from_numpy(a) if isinstance(a, np.ndarray) else a
    return t.cuda()


def test_signal_and_gradient_golden(golden):
    import mridc_b200 as mb

    g = golden("qmri")
    for i in range(int(g["ngrad"])):
        cen, nrm, nophase = (int(v) for v in g["grad%d_cfg" % i])
        t = {k: cu(g["grad%d_%s" % (i, k)]) for k in ("r2", "s0", "b0", "ph", "y", "S", "mask")}
        tes = [float(v) for v in g["grad%d_tes" % i]]
        fm = mb.SignalForwardModel(sequence="MEGRE_no_phase" if nophase else "MEGRE")
        sig = fm(t["r2"], t["s0"], t["b0"], t["ph"], tes)
        assert sig.shape == g["grad%d_signal" % i].shape
        assert rel_l2(sig, g["grad%d_signal" % i]) < 1e-6, i
        gr = mb.analytical_log_likelihood_gradient(fm, t["r2"][0], t["s0"][0], t["b0"][0], t["ph"][0], tes, t["S"][0],
                                                   t["y"][0], t["mask"][0], bool(cen), NRM3[nrm], [-2, -1], 2)
        assert gr.shape == g["grad%d_grad" % i].shape
        assert rel_l2(gr, g["grad%d_grad" % i]) < 5e-6, (i, rel_l2(gr, g["grad%d_grad" % i]))


def test_signal_model_api_and_errors():
    import mridc_b200 as mb

    m = torch.rand(2, 6, 5).cuda()
    out = mb.SignalForwardModel("MEGRE")(m, m, m, m)  # default echo times (qrim/utils.py:63-64)
    assert out.shape == (2, 4, 6, 5, 2)
    with pytest.raises(ValueError, match="Only MEGRE and MEGRE no phase"):
        mb.SignalForwardModel("SE")(m, m, m, m)
    with pytest.raises(ValueError, match="Only MEGRE and MEGRE no phase"):
        mb.SignalForwardModel(None)(m, m, m, m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mb.SignalForwardModel("MEGRE")(m.cpu(), m, m, m)
    # NaN inputs are zeroed like pred[pred != pred] = 0 (qrim/utils.py:121)
    bad = m.clone()
    bad[0, 0, 0] = float("nan")
    out = mb.SignalForwardModel("MEGRE")(bad, m, m, m)
    assert torch.isfinite(out).all() and (out[0, :, 0, 0] == 0).all()
    # RescaleByMax.reverse indexes its factors by batch position (qrim/utils.py:27-28)
    x = torch.randn(2, 4, 6, 5).cuda()
    r = mb.RescaleByMax.reverse(x, torch.tensor(QGAMMA))
    assert torch.equal(r.cpu(), torch.stack([x.cpu()[i] * torch.tensor(QGAMMA)[i] for i in range(2)], 0))
    with pytest.raises(IndexError):
        mb.RescaleByMax.reverse(torch.randn(5, 4, 3, 3).cuda(), torch.tensor(QGAMMA))
    g6 = torch.tensor([2.0, 3.0, 5.0, 7.0, 11.0, 13.0])  # more factors than the four the entry point takes per call
    x = torch.randn(6, 4, 3, 3).cuda()
    assert torch.equal(mb.RescaleByMax.reverse(x, g6).cpu(), torch.stack([x.cpu()[i] * g6[i] for i in range(6)], 0))


@pytest.mark.parametrize("B,E,C,H,W,mk", [(1, 2, 1, 4, 3, "2d"), (2, 4, 3, 17, 13, "2db"), (3, 2, 5, 16, 24, "1d"),
                                          (1, 6, 2, 31, 8, "2d")])
def test_gradient_vs_oracle(B, E, C, H, W, mk):
    """Batched product path (echoes folded into the DC operator's batch) vs the per-sample oracle, incl. the /100 and
    NaN handling of the block (qrim_block.py:204-223) and a per-batch mask."""
    from mridc_b200 import qrim
    from oracle import qnets as oq
    from oracle.make_golden import qmri_inputs

    r2, s0, b0, ph, tes, y, S, m = qmri_inputs(B, E, C, H, W, 50 + B + E, mk)
    for cen, nrm in ((False, "backward"), (True, "ortho")):
        ref = torch.stack([oq.analytical_log_likelihood_gradient(r2[b], s0[b], b0[b], ph[b], tes, S[b], y[b], m[b], cen, nrm,
                                                                 [-2, -1], 2) / 100 for b in range(B)])
        out = torch.full((B, 8, H, W), 7.0).cuda()
        maps = [t.cuda() for t in (r2, s0, b0, ph)]
        qrim._qmri_gradient(qrim.SignalForwardModel("MEGRE"), maps, None, tes, S.cuda(), y.cuda(), m.cuda(), cen, nrm, 1e-3,
                            100.0, True, out, 8)
        assert rel_l2(out[:, :4], ref) < 5e-6, (cen, nrm, rel_l2(out[:, :4], ref))
        assert (out[:, 4:] == 7.0).all()  # only the gradient half of the conv input is written


def _block(kw, sd):
    import mridc_b200 as mb

    blk = mb.qRIMBlock(**kw).cuda().eval()
    blk.load_state_dict(sd, strict=True)  # the reference's own key names
    return blk


def test_qrim_block_golden(golden):
    g = golden("qmri")
    gamma = torch.tensor(QGAMMA)
    for i in range(int(g["nblk"])):
        layer, cen, nrm = (int(v) for v in g["blk%d_cfg" % i])
        kw = dict(QRIM_KW, recurrent_layer=LAYERS[layer], fft_centered=bool(cen), fft_normalization=NRM3[nrm])
        blk = _block(kw, golden.weights(g, "blk%d_w_" % i))
        t = {k: cu(g["blk%d_%s" % (i, k)]) for k in ("r2", "s0", "b0", "ph", "y", "S", "mask")}
        tes = [float(v) for v in g["blk%d_tes" % i]]
        etas, hx = blk(t["y"], t["y"], t["r2"], t["s0"], t["b0"], t["ph"], tes, t["S"], t["mask"], None, None, gamma, False)
        assert hx is None and len(etas) == 8 and etas[0].shape == g["blk%d_first" % i].shape
        assert rel_l2(etas[0], g["blk%d_first" % i]) < 1e-5, (i, rel_l2(etas[0], g["blk%d_first" % i]))
        assert rel_l2(etas[-1], g["blk%d_last" % i]) < 1e-5, (i, rel_l2(etas[-1], g["blk%d_last" % i]))
        assert (etas[-1][:, 0] >= 0).all()
        assert len({e.data_ptr() for e in etas}) == 8  # every step's estimate is its own tensor


def test_qcirim_golden_and_structure(golden):
    import mridc_b200 as mb
    from mridc_b200 import synth

    g = golden("qmri")
    cfg = synth.qcirim_cfg(num_cascades=2, filters=16)
    model = mb.qCIRIM(cfg).cuda().eval()
    missing = model.load_state_dict(golden.weights(g, "model_w_"), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    t = {k: cu(g["model_" + k]) for k in ("r2", "s0", "b0", "ph", "y", "S", "mask")}
    y0 = t["y"].clone()
    gen = model.forward(t["r2"], t["s0"], t["b0"], t["ph"], [float(v) for v in g["model_tes"]], t["y"], t["S"],
                        torch.ones_like(t["mask"]), t["mask"])
    out = next(gen)  # a generator, like the reference (qcirim.py:312)
    assert len(out) == 5 and out[0].shape == torch.Size([])
    assert all(len(out[k]) == 2 and len(out[k][0]) == 8 for k in range(1, 5))
    assert out[1][0][0].shape == (2, 16, 12)
    for k, name in enumerate(("r2", "s0", "b0", "ph")):
        e = rel_l2(out[1 + k][-1][-1], g["model_%s_last" % name])
        assert e < 1e-4, (name, e)
    assert rel_l2(out[1][0][0], g["model_r2_first"]) < 1e-5
    assert torch.equal(t["y"], y0)  # inputs are never written


def test_qcirim_config5_vs_oracle_and_properties():
    """configs[4] geometry: 32 coils, 4 echoes, 232 x 288, cached 12x Poisson-disc mask, IndRNN-128 quantitative module."""
    import mridc_b200 as mb
    from mridc_b200 import synth
    from oracle import qnets as oq

    d = synth.make_qmri_batch(1)
    cfg = synth.qcirim_cfg()
    torch.manual_seed(1)
    model = mb.qCIRIM(cfg).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    args = [d[k] for k in ("R2star_map_init", "S0_map_init", "B0_map_init", "phi_map_init")] + [d["TEs"]] + [
        d[k] for k in ("y", "sensitivity_maps", "mask_brain", "sampling_mask")]
    out = next(model.forward(*[a.cuda() if isinstance(a, torch.Tensor) else a for a in args]))
    ref = oq.qcirim_forward(sd, cfg, *args)
    for k, name in enumerate(("R2star", "S0", "B0", "phi")):
        e = rel_l2(out[1 + k][-1][-1], ref[1 + k][-1][-1])
        assert e < 1e-4, (name, e)
    # property: with k-space that is exactly the forward model of the (init x gamma / gamma) maps the residual, hence
    # the data-consistency gradient, vanishes to fp32 round-off of the signal model
    from mridc_b200 import qrim

    maps = [d[k].cuda() for k in ("R2star_map", "S0_map", "B0_map", "phi_map")]
    S, m = d["sensitivity_maps"].cuda(), d["sampling_mask"].cuda()
    sig = mb.SignalForwardModel("MEGRE")(*maps, d["TEs"])
    full = mb.fft2(mb.complex_mul(sig.unsqueeze(2), S.unsqueeze(1)), centered=False, normalization="backward",
                   spatial_dims=[-2, -1])
    y_consistent = full * m.unsqueeze(1)
    g0 = torch.empty(1, 4, *synth.QMRI_SHAPE).cuda()
    qrim._qmri_gradient(qrim.SignalForwardModel("MEGRE"), maps, None, d["TEs"], S, y_consistent, m, False, "backward", 1e-3,
                        1.0, False, g0, 4)
    g1 = torch.empty_like(g0)
    qrim._qmri_gradient(qrim.SignalForwardModel("MEGRE"), maps, None, d["TEs"], S, torch.zeros_like(y_consistent), m, False,
                        "backward", 1e-3, 1.0, False, g1, 4)
    assert g0.abs().max().item() < 1e-4 * g1.abs().max().item()

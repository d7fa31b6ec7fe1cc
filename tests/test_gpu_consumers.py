"""GPU parity of the other consumers of the DC operator through the C-ABI: sigmanet DC layers
(gradient, prox-CG, variable splitting, single-coil), CascadeNetBlock, RecurrentInit / RecurrentVarNetBlock (Conv2dGRU),
qVarNetBlock -- against the reference's own outputs (tests/golden/consumers.npz) and, at the fastMRI size, against
properties of the operators.  Tolerances: layers rel-L2 <= 2e-6, CG (6 iterations) <= 2e-5, blocks <= 1e-5."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
NRM3 = ["backward", "ortho", "forward"]
SD = [-2, -1]


def cu(a):
    return (torch.from_numpy(np.asarray(a)) if not isinstance(a, torch.Tensor) else a).cuda()


def load_sd(module, g, prefix):
    sd = {k: torch.from_numpy(g[prefix + k.replace(".", "_")]) for k in module.state_dict().keys()}
    module.load_state_dict(sd, strict=True)
    return module.cuda()


def test_sigmanet_layers_golden(golden):
    import mridc_b200 as mb

    g = golden("consumers")
    for i in range(int(g["nsig"])):
        cen, nrm = bool(g["sig%d_cfg" % i][0]), NRM3[int(g["sig%d_cfg" % i][1])]
        x, y, S, m = (cu(g["sig%d_%s" % (i, k)]) for k in ("x", "y", "smaps", "mask"))
        kw = dict(fft_centered=cen, fft_normalization=nrm, spatial_dims=SD)
        assert rel_l2(mb.DataGDLayer(0.7, **kw).cuda()(x, y, S, m), g["sig%d_gd" % i]) < 2e-6
        assert rel_l2(mb.DataVSLayer(0.3, 0.6, **kw).cuda()(x, y, S, m), g["sig%d_vs" % i]) < 2e-6
        if "sig%d_cg" % i in g:
            e = rel_l2(mb.DataProxCGLayer(0.5, tol=1e-6, iter=6, **kw).cuda()(x, y, S, m), g["sig%d_cg" % i])
            assert e < 2e-5, e
    x, y, S, m = (cu(g["dun_" + k]) for k in ("x", "y", "smaps", "mask"))
    gd = mb.DataGDLayer(0.4, fft_centered=True, fft_normalization="ortho", spatial_dims=SD).cuda()
    x1 = gd(x, y, S, m)
    assert rel_l2(x1, g["dun_gd1"]) < 2e-6 and rel_l2(gd(x1, y, S, m), g["dun_gd2"]) < 2e-6
    assert rel_l2(mb.DataVSLayer(0.2, 0.5).cuda()(x, y, S, m), g["dun_vs"]) < 2e-6
    dl = mb.DCLayer(0.25, fft_centered=False, fft_normalization="ortho", spatial_dims=SD).cuda()
    assert rel_l2(dl(cu(g["dcl_x"]), cu(g["dcl_y"]), cu(g["dcl_mask"])), g["dcl_out"]) < 2e-6
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dl(torch.from_numpy(g["dcl_x"]), torch.from_numpy(g["dcl_y"]), torch.from_numpy(g["dcl_mask"]))


def test_prox_cg_solves_the_normal_equations_full_size():
    """15 coils x 320 x 320: the returned x satisfies (lambda A^H A + I) x = lambda A^H y + z to the CG tolerance."""
    import mridc_b200 as mb

    g = torch.Generator().manual_seed(3)
    C, H, W = 15, 320, 320
    S = torch.randn(1, C, 1, H, W, 2, generator=g) * 0.2
    m = (torch.rand(1, 1, 1, 1, W, 1, generator=g) < 0.3).float()
    y = torch.randn(1, C, 1, H, W, 2, generator=g) * m
    z = torch.randn(1, 1, H, W, 2, generator=g)
    S, m, y, z = S.cuda(), m.cuda(), y.cuda(), z.cuda()
    lam = 0.6
    x = mb.DataProxCGLayer(lam, tol=1e-10, iter=30).cuda()(z, y, S, m)
    kw = dict(centered=True, normalization="ortho", spatial_dims=SD)
    AT = lambda k: torch.sum(mb.complex_mul(mb.ifft2(k * m, **kw), mb.complex_conj(S)), dim=-5)
    A = lambda v: torch.sum(mb.fft2(mb.complex_mul(v.expand_as(S), S), **kw) * m, dim=-4, keepdim=True)
    lhs = lam * AT(A(x)) + x
    rhs = lam * AT(y) + z
    assert rel_l2(lhs, rhs) < 1e-4


class _Reg(torch.nn.Module):
    """conv3x3(2 -> 8) + ReLU + conv3x3(8 -> 2), zero padding, on the package's conv kernel."""

    def __init__(self, g):
        super().__init__()
        self.w = {k: cu(g["ccnn_w_" + k]) for k in ("0_weight", "0_bias", "2_weight", "2_bias")}

    def forward(self, x):
        from mridc_b200 import _ops

        x = _ops.conv2d(x.contiguous(), self.w["0_weight"], self.w["0_bias"], 3, 1, _ops.PAD_ZERO, act=_ops.ACT_RELU)
        return _ops.conv2d(x, self.w["2_weight"], self.w["2_bias"], 3, 1, _ops.PAD_ZERO)


def test_cascadenet_block_golden(golden):
    import mridc_b200 as mb

    g = golden("consumers")
    for i, no_dc in enumerate((False, True)):
        pred, y, S = (cu(g["ccnn%d_%s" % (i, k)]) for k in ("pred", "y", "S"))
        m = cu(g["ccnn%d_mask" % i])
        blk = mb.CascadeNetBlock(_Reg(g), True, "ortho", SD, 1, no_dc).cuda()
        blk.dc_weight.data.fill_(0.8)
        out = blk(pred, y, S, m)
        assert out.shape == g["ccnn%d_out" % i].shape
        assert rel_l2(out, g["ccnn%d_out" % i]) < 5e-6, rel_l2(out, g["ccnn%d_out" % i])
        assert rel_l2(blk.sens_expand(blk.sens_reduce(pred, S), S),
                      mb.fft2(mb.complex_mul(blk.sens_reduce(pred, S), S), True, "ortho")) < 2e-6


def test_recurrentvarnet_golden(golden):
    import mridc_b200 as mb

    g = golden("consumers")
    init = load_sd(mb.RecurrentInit(2, 8, (8, 8), (1, 2), depth=2, multiscale_depth=2), g, "rvn_init_")
    blk = load_sd(mb.RecurrentVarNetBlock(2, 8, 2, True, "ortho", SD, 1), g, "rvn_blk_")
    y, S, m, cur = (cu(g["rvn_" + k]) for k in ("y", "S", "mask", "cur"))
    h0 = init(cu(g["rvn_img0"]))
    assert rel_l2(h0, g["rvn_h0"]) < 2e-6
    k1, h1 = blk(cur, y, m, S, h0)
    k2, h2 = blk(k1, y, m, S, h1)
    assert k2.shape == g["rvn_k2"].shape and h2.shape == g["rvn_h2"].shape
    assert rel_l2(k1, g["rvn_k1"]) < 5e-6 and rel_l2(h1, g["rvn_h1"]) < 5e-6
    assert rel_l2(k2, g["rvn_k2"]) < 1e-5 and rel_l2(h2, g["rvn_h2"]) < 1e-5
    k1n, h1n = blk(cur, y, m, S, None)
    assert rel_l2(k1n, g["rvn_k1n"]) < 5e-6 and rel_l2(h1n, g["rvn_h1n"]) < 5e-6


def test_recurrentvarnet_block_full_size_fixed_point():
    """15 x 320 x 320: with a zero regulariser output layer and learning rate 1 the block is hard data consistency:
    sampled columns become the measurements (k - (k - y), one rounding), the others are untouched (bit-exact)."""
    import mridc_b200 as mb

    g = torch.Generator().manual_seed(4)
    B, C, H, W = 2, 15, 320, 320
    blk = mb.RecurrentVarNetBlock(2, 16, 2, False, "backward", SD, 1).cuda()
    blk.regularizer.conv_blocks[2][1].weight.data.zero_()
    blk.regularizer.conv_blocks[2][1].bias.data.zero_()
    m = (torch.rand(1, 1, 1, W, 1, generator=g) < 0.25).float().cuda()
    y = torch.randn(B, C, H, W, 2, generator=g).cuda() * m
    S = torch.randn(B, C, H, W, 2, generator=g).cuda() * 0.2
    cur = torch.randn(B, C, H, W, 2, generator=g).cuda()
    new, h = blk(cur, y, m, S, None)
    assert h.shape == (B, 16, H, W, 2)
    mm = m.bool().expand_as(cur)
    assert torch.allclose(new[mm], y[mm], rtol=0, atol=2e-6) and torch.equal(new[~mm], cur[~mm])


def test_qvarnet_block_golden(golden):
    import mridc_b200 as mb

    g = golden("consumers")
    unet = mb.NormUnet(chans=4, num_pools=2, in_chans=8, out_chans=8, padding_size=3, normalize=True)
    qb = load_sd(mb.qVarNetBlock(unet, True, "ortho", SD, 2, False), g, "qvn_w_")
    y, S, sm = cu(g["qvn_y"]), cu(g["qvn_S"]), cu(g["qvn_mask"])
    maps = [cu(g["qvn_" + k]) for k in ("R2", "S0", "B0", "phi")]
    out = qb(y.clone(), y, *maps, [3.0, 11.5, 20.0, 28.5], S, sm, torch.from_numpy(g["qvn_gamma"]))
    assert out.shape == g["qvn_out"].shape
    assert rel_l2(out, g["qvn_out"]) < 1e-5, rel_l2(out, g["qvn_out"])
    assert float(out[:, 0].min()) >= 0.0  # negative R2* estimates are clipped (qvn_block.py:156-158)


def test_jrscirim_block_golden(golden):
    """JRSCIRIMBlock: CIRIM cascades on the fused DC operator + UNet / ConvLayer segmentation heads, single slices, a given
    initial image, and the per-slice loop of consecutive_slices = 2 -- against the reference's own outputs.  The 8-channel
    ConvGRU of the fixture runs on the exact-fp32 kernels: blocks <= 1e-5, segmentation (after group-norm + U-Net) <= 1e-4."""
    import mridc_b200 as mb
    from oracle.make_golden import JRS_CASES, JRS_RIM

    g = golden("jrscirim")
    for idx, (name, sp, in_ch, mag, slices) in enumerate(JRS_CASES):
        p = "jrs%d_" % idx
        blk = load_sd(mb.JRSCIRIMBlock(dict(JRS_RIM), dict(sp), in_ch, mag, True, "ortho", SD, 2, 2, slices, "SENSE", True), g,
                      p + "w_")
        y, S, m, init, target = (cu(g[p + k]) for k in ("y", "S", "mask", "init", "target"))
        no_init = torch.zeros(2, slices).cuda() if slices > 1 else torch.zeros(1).cuda()
        rec, seg, hx = blk(y, S, m, init if idx == 1 else no_init, target)
        rec = torch.view_as_real(torch.stack([torch.stack(c) for c in rec]))
        assert rec.shape == g[p + "rec"].shape and seg.shape == g[p + "seg"].shape, name
        e_r, e_s = rel_l2(rec, g[p + "rec"]), rel_l2(seg, g[p + "seg"])
        print("[jrscirim] %s rec %.2e seg %.2e" % (name, e_r, e_s))
        assert e_r < 1e-5 and e_s < 1e-4, (name, e_r, e_s)
        assert len(hx) == 2 and hx[0].shape[1] == 8
    with pytest.raises(NotImplementedError):
        mb.JRSCIRIMBlock(dict(JRS_RIM), dict(segmentation_module="AttentionUNet", output_channels=2), 1)
    with pytest.raises(ValueError, match="not implemented"):
        mb.JRSCIRIMBlock(dict(JRS_RIM), dict(segmentation_module="nope", output_channels=2), 1)

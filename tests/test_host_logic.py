"""CPU: host-side logic of the drop-in layer (no compute calls): parameter/key compatibility with the reference,
initialiser parity, error behaviour, masks, sharding arithmetic."""
import numpy as np
import pytest
import torch

from oracle import ref_import

RIM_HP = dict(conv_filters=[16, 16, 2], conv_kernels=[5, 3, 3], conv_dilations=[1, 2, 1],
              conv_bias=[True, True, False], recurrent_filters=[16, 16, 0], recurrent_kernels=[1, 1, 0],
              recurrent_dilations=[1, 1, 0], recurrent_bias=[True, True, False])


def test_masks_bit_exact_vs_reference_golden(golden):
    from mridc_b200 import synth

    g = golden("masks")
    cases = [("random", synth.RandomMask1D, [0.08], [4], (1, 320, 320, 2), 123),
             ("random8", synth.RandomMask1D, [0.04], [8], (1, 640, 320, 2), 7),
             ("equi", synth.Equispaced1DMask, [0.08], [4], (1, 320, 320, 2), 123),
             ("equi8", synth.Equispaced1DMask, [0.04], [8], (1, 640, 320, 2), 123),
             ("equi_multi", synth.Equispaced1DMask, [0.08, 0.04], [4, 8], (1, 218, 170, 2), (1, 2, 3))]
    for name, cls, cf, acc, shape, seed in cases:
        m, a = cls(cf, acc)(shape, seed)
        assert m.dtype == torch.float32 and tuple(m.shape) == g[name].shape
        assert np.array_equal(m.numpy(), g[name]), name
        assert a == int(g[name + "_acc"])
    for name, acc in (("gauss4", 4), ("gauss8", 8)):
        np.random.seed(123)
        m, _ = synth.Gaussian1DMask([0.7], [acc])((1, 320, 320, 2), 0, scale=0.02)
        assert np.array_equal(m.numpy(), g[name]), name
        assert abs(m.mean().item() - 1.0 / acc) < 0.05
    # seeded masks do not disturb the generator's state and are reproducible
    f = synth.Equispaced1DMask([0.08], [4])
    assert torch.equal(f((1, 64, 64, 2), 5)[0], f((1, 64, 64, 2), 5)[0])
    with pytest.raises(ValueError):
        synth.RandomMask1D([0.08], [4, 8])


def test_state_dict_keys_match_reference_layout(golden):
    """Reference state_dicts (stored in the golden files) load strictly into the drop-in modules."""
    import mridc_b200 as mb

    g = golden("rim")
    blk = mb.RIMBlock(recurrent_layer="GRU", depth=2, time_steps=8, conv_dim=2, no_dc=True, **RIM_HP)
    blk.load_state_dict(golden.weights(g, "rim0_w_"), strict=True)
    assert sorted(blk.state_dict()) == sorted(golden.weights(g, "rim0_w_"))
    blk = mb.RIMBlock(recurrent_layer="IndRNN", depth=2, time_steps=8, conv_dim=2, no_dc=True, **RIM_HP)
    blk.load_state_dict(golden.weights(g, "rim2_w_"), strict=True)
    blk = mb.RIMBlock(recurrent_layer="GRU", depth=2, time_steps=8, conv_dim=2, no_dc=False, **RIM_HP)
    assert "dc_weight" in blk.state_dict()
    g = golden("unet_vn")
    vb = mb.VarNetBlock(mb.NormUnet(chans=4, num_pools=2, padding_size=11))
    vb.load_state_dict(golden.weights(g, "vn0_w_"), strict=True)
    full = mb.CIRIM(dict(RIM_HP, recurrent_layer="GRU", depth=2, time_steps=5, conv_dim=2, no_dc=True,
                         num_cascades=3, dimensionality=2, keep_eta=True, fft_centered=False,
                         fft_normalization="backward", spatial_dims=[-2, -1], coil_dim=1,
                         coil_combination_method="SENSE"))
    assert full.time_steps == 8  # rounded up to a multiple of 8 (cirim.py:51)
    keys = set(full.state_dict())
    assert "dc_weight" in keys and "cirim.2.layers.1.rnn.hh.weight" in keys
    assert "cirim.0.final_layer.0.conv_layer.weight" in keys and "cirim.0.final_layer.0.conv_layer.bias" not in keys
    # 94,080 parameters per GRU cascade at 64 filters
    big = mb.RIMBlock(recurrent_layer="GRU", conv_filters=[64, 64, 2], conv_kernels=[5, 3, 3],
                      conv_dilations=[1, 2, 1], conv_bias=[True, True, False], recurrent_filters=[64, 64, 0],
                      recurrent_kernels=[1, 1, 0], recurrent_dilations=[1, 1, 0], recurrent_bias=[True, True, False],
                      depth=2, time_steps=8, conv_dim=2, no_dc=True)
    assert sum(p.numel() for p in big.parameters()) == 94080
    vn = mb.VarNetBlock(mb.NormUnet(chans=14, num_pools=2, padding_size=11))
    assert sum(p.numel() for p in vn.parameters()) == 89267


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_initialisers_reproduce_reference_weights():
    """Same seed -> same random-init weights as the reference modules (same init calls in the same order)."""
    import mridc_b200 as mb

    R = ref_import.Ref()
    for layer in ("GRU", "MGU", "IndRNN"):
        kw = dict(recurrent_layer=layer, depth=2, time_steps=8, conv_dim=2, no_dc=True, **RIM_HP)
        torch.manual_seed(1)
        a = R.rim_block.RIMBlock(**kw).state_dict()
        torch.manual_seed(1)
        b = mb.RIMBlock(**kw).state_dict()
        assert sorted(a) == sorted(b)
        for k in a:
            assert torch.equal(a[k], b[k]), (layer, k)
    torch.manual_seed(2)
    a = R.unet_block.NormUnet(chans=6, num_pools=3, padding_size=11).state_dict()
    torch.manual_seed(2)
    b = mb.NormUnet(chans=6, num_pools=3, padding_size=11).state_dict()
    assert sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_error_behaviour_matches_reference_texts():
    import mridc_b200 as mb

    x = torch.zeros(2, 3, 4, 2)
    with pytest.raises(ValueError, match="Tensors do not have separate complex dim."):
        mb.complex_mul(x[..., :1], x)
    with pytest.raises(ValueError, match="Tensor does not have separate complex dim."):
        mb.complex_conj(x[..., :1])
    with pytest.raises(ValueError, match="Output type not supported."):
        mb.coil_combination(x, x, method="espirit")
    with pytest.raises(ValueError, match="len\\(shift\\) must match len\\(dim\\)"):
        mb.roll(x, [1], [0, 1])
    with pytest.raises(ValueError, match="Invalid shapes."):
        mb.center_crop(x, (9, 1))
    with pytest.raises(ValueError, match="Please specify a proper recurrent layer type."):
        mb.RIMBlock(recurrent_layer="LSTM", depth=2, conv_dim=2, **RIM_HP)
    with pytest.raises(ValueError, match="Please specify a proper nonlinearity"):
        mb.ConvNonlinear(4, 8, 2, 3, 1, True, "gelu")
    # no CPU fallback anywhere on the compute path
    for fn in (lambda: mb.fft2(x), lambda: mb.ifft2(x), lambda: mb.complex_abs(x), lambda: mb.rss(x, 1),
               lambda: mb.sense(x, x, 1), lambda: mb.fftshift(x)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fn()
    # crops are pure index arithmetic and work on any device, bit exact
    a = torch.arange(6 * 8).reshape(6, 8)
    assert torch.equal(mb.center_crop(a, (2, 4)), a[2:4, 2:6])
    assert torch.equal(mb.check_stacked_complex(x), torch.view_as_complex(x))


def test_partition_and_shard_arithmetic():
    from mridc_b200 import sharding

    assert sharding.partition(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sharding.partition(8, 8) == [(i, i + 1) for i in range(8)]
    assert sharding.partition(3, 8)[3:] == [(3, 3)] * 5
    assert sharding.partition(0, 2) == [(0, 0), (0, 0)]
    for n in range(0, 40):
        for w in (1, 2, 3, 8):
            p = sharding.partition(n, w)
            assert p[0][0] == 0 and p[-1][1] == n and all(a[1] == b[0] for a, b in zip(p, p[1:]))
    y = torch.arange(10).reshape(10, 1)
    m = torch.ones(1, 5)
    (ys, ms, none), (a, b) = sharding.shard_slices([y, m, None], rank=1, world_size=4)
    assert (a, b) == (3, 6) and torch.equal(ys, y[3:6]) and ms is m and none is None
    with pytest.raises(ValueError):
        sharding.partition(4, 0)


def test_synth_batch_shapes_and_determinism():
    from mridc_b200 import synth

    b1 = synth.make_batch(2, 4, 32, 24, centered=True, normalization="ortho")
    b2 = synth.make_batch(2, 4, 32, 24, centered=True, normalization="ortho")
    assert b1["y"].shape == (2, 4, 32, 24, 2) and b1["mask"].shape == (1, 1, 1, 24, 1)
    assert b1["mask"].dtype == torch.uint8 and b1["target"].dtype == torch.complex64
    assert all(torch.equal(b1[k], b2[k]) for k in ("y", "sensitivity_maps", "mask", "target"))
    # masked columns of y are exactly zero; sampled ones are not
    m = b1["mask"].bool().reshape(-1)
    assert torch.count_nonzero(b1["y"][..., ~m, :]) == 0 and torch.count_nonzero(b1["y"][..., m, :]) > 0
    assert not torch.equal(b1["y"][0], b1["y"][1])


def test_host_prefetcher_rejects_cpu_device():
    import pytest
    import mridc_b200 as mb

    with pytest.raises(RuntimeError, match="CUDA device"):
        mb.HostPrefetcher([], "cpu")


def test_qcirim_ctor_errors():
    import mridc_b200 as mb
    from mridc_b200 import synth

    with pytest.raises(ValueError, match="Only 2D is currently supported"):
        mb.qCIRIM(dict(synth.qcirim_cfg(filters=8), quantitative_module_dimensionality=3))
    with pytest.raises(ValueError, match="does not support explicit DC"):
        mb.qCIRIM(dict(synth.qcirim_cfg(filters=8), quantitative_module_no_dc=False))
    with pytest.raises(NotImplementedError):
        mb.qCIRIM(dict(synth.qcirim_cfg(filters=8), use_reconstruction_module=True))
    # no CPU fallback on the quantitative path either
    m = torch.zeros(1, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mb.SignalForwardModel("MEGRE")(m, m, m, m)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mb.RescaleByMax.reverse(torch.zeros(1, 4, 4, 4), torch.ones(4))


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_qrim_initialisers_reproduce_reference_weights():
    import mridc_b200 as mb

    R = ref_import.Ref()
    kw = dict(conv_filters=[8, 8, 4], conv_kernels=[5, 3, 3], conv_dilations=[1, 2, 1], conv_bias=[True, True, False],
              recurrent_filters=[8, 8, 0], recurrent_kernels=[1, 1, 0], recurrent_dilations=[1, 1, 0],
              recurrent_bias=[True, True, False], depth=2, time_steps=8, conv_dim=2, no_dc=True, coil_dim=2)
    for layer in ("GRU", "MGU", "IndRNN"):
        torch.manual_seed(3)
        a = R.qrim_block.qRIMBlock(recurrent_layer=layer, **kw).state_dict()
        torch.manual_seed(3)
        b = mb.qRIMBlock(recurrent_layer=layer, **kw).state_dict()
        assert sorted(a) == sorted(b)
        for k in a:
            assert torch.equal(a[k], b[k]), (layer, k)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")
def test_sens_net_initialiser_and_bookkeeping_match_reference():
    import mridc_b200 as mb
    from oracle.make_golden import _ref_sens_model

    Sens = _ref_sens_model(ref_import.Ref())
    torch.manual_seed(9)
    a = Sens(4, 3, mask_type="1D").state_dict()
    torch.manual_seed(9)
    b = mb.BaseSensitivityModel(4, 3, mask_type="1D").state_dict()
    assert sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a)
    g = torch.Generator().manual_seed(3)
    for _ in range(5):
        m = (torch.rand(3, 1, 1, 24, 1, generator=g) < 0.5).float()
        m[:, 0, 0, 10:14, 0] = 1
        for nlf in (None, 0, 4):
            pa, na = Sens.get_pad_and_num_low_freqs(m, nlf)
            pb, nb = mb.BaseSensitivityModel.get_pad_and_num_low_freqs(m, nlf)
            assert torch.equal(pa, pb) and torch.equal(na, nb)
    # a model with use_sens_net builds the network first (base.py:81-94): same keys / RNG order as the reference
    from mridc_b200 import synth

    cfg = dict(synth.varnet_cfg(num_cascades=1), use_sens_net=True, sens_chans=4, sens_pools=2, sens_mask_type="2D",
               sens_normalize=True, sens_mask_center=True)
    torch.manual_seed(11)
    ours = mb.VarNet(cfg).state_dict()
    torch.manual_seed(11)
    ref_first = Sens(4, 2, fft_centered=False, fft_normalization="backward", spatial_dims=[-2, -1], coil_dim=1,
                     mask_type="2D", normalize=True, mask_center=True).state_dict()
    for k, v in ref_first.items():
        assert torch.equal(ours["sens_net." + k], v), k


def _replay_apply_mask(golden, device):
    """tests/golden/apply_mask.npz (oracle/make_golden.py::gen_apply_mask): the reference's apply_mask
    (common/parts/utils.py:293-343) on seeded k-space -- data, mask and acceleration must come back bit for bit,
    including the cleared sign of zeros (":341 + 0.0"), the padding zeroing (:332-335) and the mask fftshift (:337-338)."""
    from mridc_b200 import synth, utils as mutils

    g = golden("apply_mask")
    for i in range(int(g["nam"])):
        kind, shift = (int(v) for v in g["am%d_cfg" % i])
        cf, acc = [float(v) for v in g["am%d_cf" % i]], [int(v) for v in g["am%d_accs" % i]]
        seed = g["am%d_seed" % i]
        seed = int(seed) if seed.ndim == 0 else tuple(int(v) for v in seed)
        pad = tuple(int(v) for v in g["am%d_pad" % i])
        pad = None if pad == (-1, -1) else pad
        if kind == 2:  # 2-D equispaced mask drawn by the reference, replayed through a fixed mask function
            raw, racc = torch.from_numpy(g["am%d_raw" % i]), float(g["am%d_acc" % i])

            def fn(shape, seed, half_scan_percentage=0.0, scale=0.02, _m=raw, _a=racc):
                return _m.clone(), _a
        else:
            fn = (synth.Equispaced1DMask if kind == 0 else synth.RandomMask1D)(cf, acc)
        data = torch.from_numpy(g["am%d_in" % i]).to(device)
        out, mask, a = mutils.apply_mask(data, fn, seed=seed, padding=pad, shift=bool(shift))
        assert out.device == data.device and mask.device == data.device
        ref = torch.from_numpy(g["am%d_out" % i])
        assert torch.equal(out.cpu(), ref), i
        assert torch.equal(torch.signbit(out.cpu()), torch.signbit(ref)), i
        assert torch.equal(mask.cpu(), torch.from_numpy(g["am%d_mask" % i])), i
        assert float(a) == float(g["am%d_acc" % i]), i


def test_apply_mask_golden(golden):
    _replay_apply_mask(golden, "cpu")

"""The reference's own numeric known-answer tests (tests/collections/reconstruction/test_fft.py:17-199) restated
(np.prod instead of the removed np.product): oracle on CPU, CUDA path on the GPU box.  Same shapes, same
`arange` inputs, numpy.fft as the truth, np.allclose defaults (rtol 1e-5, atol 1e-8) for the oracle; the CUDA
kernels use their own butterflies, so exact-zero bins come out at fp32 round-off of the row norm: rtol 1e-5 and
atol 1e-5 * max|truth| are used there (stated tolerance)."""
import numpy as np
import pytest
import torch

from oracle import mri as omri

SHAPES = [[3, 3], [4, 6], [10, 8, 4]]


def create_input(shape):
    x = np.arange(np.prod(shape)).reshape(shape)
    return torch.from_numpy(x).float()


def _np_c(x):
    x = x.numpy()
    return x[..., 0] + 1j * x[..., 1]


def _truth(x, centered, norm, inverse):
    a = _np_c(x)
    if centered:
        a = np.fft.ifftshift(a, (-2, -1))
    a = (np.fft.ifft2 if inverse else np.fft.fft2)(a, norm=norm)
    if centered:
        a = np.fft.fftshift(a, (-2, -1))
    return a


CASES = [(s, c, n, i) for s in SHAPES for c in (True, False) for n in ("ortho", "backward", "forward")
         for i in (False, True)]


@pytest.mark.parametrize("shape,centered,norm,inverse", CASES)
def test_oracle_fft_kat(shape, centered, norm, inverse):
    x = create_input(shape + [2])
    fn = omri.ifft2 if inverse else omri.fft2
    out = _np_c(fn(x, centered=centered, normalization=norm, spatial_dims=[-2, -1]))
    assert np.allclose(out, _truth(x, centered, norm, inverse))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,centered,norm,inverse", CASES)
def test_cuda_fft_kat(shape, centered, norm, inverse):
    import mridc_b200 as mb

    x = create_input(shape + [2])
    fn = mb.ifft2 if inverse else mb.fft2
    out = _np_c(fn(x.cuda(), centered=centered, normalization=norm, spatial_dims=[-2, -1]).cpu())
    truth = _truth(x, centered, norm, inverse)
    assert np.allclose(out, truth, rtol=1e-5, atol=1e-5 * np.abs(truth).max())


def test_oracle_complex_abs_and_shifts():
    # test_fft.py:149-199
    x = create_input([3, 3, 2])
    assert np.allclose(omri.complex_abs(x).numpy(), np.abs(_np_c(x)))
    for shape in ([5, 6, 2], [3, 7]):
        a = create_input(shape)
        assert np.array_equal(omri.fftshift(a).numpy(), np.fft.fftshift(a.numpy()))
        assert np.array_equal(omri.ifftshift(a).numpy(), np.fft.ifftshift(a.numpy()))
    a = create_input([4, 5, 6])
    for shift, dim in ((0, 0), (1, 0), (-1, 0), (100, 0), (2, 1), (-7, 2)):
        assert np.array_equal(omri.roll(a, [shift], [dim]).numpy(), np.roll(a.numpy(), shift, dim))


@pytest.mark.gpu
def test_cuda_complex_abs_and_shifts():
    import mridc_b200 as mb

    x = create_input([3, 3, 2])
    assert np.allclose(mb.complex_abs(x.cuda()).cpu().numpy(), np.abs(_np_c(x)))
    for shape in ([5, 6, 2], [3, 7]):
        a = create_input(shape)
        assert np.array_equal(mb.fftshift(a.cuda()).cpu().numpy(), np.fft.fftshift(a.numpy()))
        assert np.array_equal(mb.ifftshift(a.cuda()).cpu().numpy(), np.fft.ifftshift(a.numpy()))
    a = torch.arange(4 * 5 * 6).reshape(4, 5, 6)  # int64, as the reference's roll tests
    for shift, dim in ((0, 0), (1, 0), (-1, 0), (100, 0), (2, 1), (-7, 2)):
        assert np.array_equal(mb.roll(a.cuda(), [shift], [dim]).cpu().numpy(), np.roll(a.numpy(), shift, dim))

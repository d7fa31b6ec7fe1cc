"""Deterministic synthetic fastMRI-shaped inputs and sampling masks (host side, numpy only).

Input preparation, not the hot path: the reference builds its inputs on the host too (h5 -> numpy ->
``MRIDataTransforms``).  The generator follows SURVEY.md section 8(d): analytic phantom, Gaussian-profile coil
maps normalised by their RSS, k-space = fft2(S * x), reference-style masking and max-normalisation
(mridc/collections/reconstruction/parts/transforms.py:526-614).

Mask generators restate the host algorithms of mridc/collections/reconstruction/data/subsample.py
(RandomMaskFunc :113-153, Equispaced1DMaskFunc :175-222) and the *working* Gaussian-1D of
mridc/collections/common/data/subsample.py:377-470; tests/golden/masks.npz pins them bit-exactly against
the reference functions.
"""
from typing import Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------------------
# masks
# --------------------------------------------------------------------------------------------------
class _SeededMask:
    def __init__(self, center_fractions: Sequence[float], accelerations: Sequence[int]):
        if len(center_fractions) != len(accelerations):
            raise ValueError("Number of center fractions should match number of accelerations")
        self.center_fractions = list(center_fractions)
        self.accelerations = list(accelerations)
        self.rng = np.random.RandomState()

    def _pick(self):
        i = self.rng.randint(0, len(self.accelerations))
        return self.center_fractions[i], self.accelerations[i]

    def _seeded(self, seed, fn):
        if seed is None:
            return fn()
        state = self.rng.get_state()
        self.rng.seed(seed)
        try:
            return fn()
        finally:
            self.rng.set_state(state)

    @staticmethod
    def _as_tensor(mask_1d, shape):
        import torch

        dims = [1] * len(shape)
        dims[-2] = mask_1d.shape[0]
        return torch.from_numpy(np.ascontiguousarray(mask_1d.reshape(dims).astype(np.float32)))


class RandomMask1D(_SeededMask):
    """Random column mask with a fully sampled centre (subsample.py:94-153)."""

    def __call__(self, shape, seed=None, half_scan_percentage=0.0, scale=0.02):
        if len(shape) < 3:
            raise ValueError("Shape should have 3 or more dimensions")

        def gen():
            ncol = int(shape[-2])
            cf, acc = self._pick()
            nlow = int(round(ncol * cf))
            prob = (ncol / acc - nlow) / (ncol - nlow)
            m = self.rng.uniform(size=ncol) < prob
            start = (ncol - nlow + 1) // 2
            m[start:start + nlow] = True
            return self._as_tensor(m, shape), acc

        return self._seeded(seed, gen)


class Equispaced1DMask(_SeededMask):
    """Equispaced columns with random offset plus fully sampled centre (subsample.py:156-222)."""

    def __call__(self, shape, seed=None, half_scan_percentage=0.0, scale=0.02):
        if len(shape) < 3:
            raise ValueError("Shape should have 3 or more dimensions")

        def gen():
            cf, acc = self._pick()
            ncol = int(shape[-2])
            nlow = int(round(ncol * cf))
            m = np.zeros(ncol, dtype=np.float32)
            start = (ncol - nlow + 1) // 2
            m[start:start + nlow] = 1.0
            adj = (acc * (nlow - ncol)) / (nlow * acc - ncol)
            offset = self.rng.randint(0, round(adj))
            cols = np.around(np.arange(offset, ncol - 1, adj)).astype(np.uint)
            m[cols] = 1.0
            return self._as_tensor(m, shape), acc

        return self._seeded(seed, gen)


class Gaussian1DMask(_SeededMask):
    """Gaussian-density 1-D column mask (the fixed variant, common/data/subsample.py:377-470): columns drawn
    without replacement from a Gaussian profile with the *global* numpy RNG (``seed`` is ignored, as upstream),
    plus a fully sampled centre band of ``int(ncol * scale)`` columns.  center_fractions = FWHM."""

    def __call__(self, shape, seed=None, half_scan_percentage=0.0, scale=0.02):
        nrow, ncol = int(shape[-3]), int(shape[-2])
        i = self.rng.randint(0, len(self.accelerations))
        fwhm, acc = self.center_fractions[i], self.accelerations[i]
        # upstream works on the transposed (ncol, nrow) grid and transposes back at the end
        scaled = int(ncol * scale)
        top = (ncol - scaled) // 2
        grid = np.concatenate((np.zeros((top, nrow)), np.ones((scaled, nrow)),
                               np.zeros((ncol - scaled - top, nrow))))
        n_sample = int(ncol / acc)
        sigma = fwhm / np.sqrt(8 * np.log(2))
        xs = np.linspace(-1.0, 1.0, ncol)
        kern = np.exp(-(xs**2 / (2 * sigma**2)))
        kern = kern / kern.sum()
        idxs = np.random.choice(range(ncol), size=n_sample, replace=False, p=kern)
        grid[idxs, :] = 1.0
        grid = np.fft.ifftshift(np.fft.ifftshift(np.fft.ifftshift(grid, axes=0), axes=0), axes=(0, 1))
        if half_scan_percentage != 0:
            grid[: int(np.round(grid.shape[0] * half_scan_percentage)), :] = 0.0
        grid = grid.T
        return self._as_tensor(grid[0], shape), acc


# --------------------------------------------------------------------------------------------------
# phantom
# --------------------------------------------------------------------------------------------------
def _fft2c(x, centered, norm, inverse=False):
    fn = np.fft.ifft2 if inverse else np.fft.fft2
    nm = None if norm in ("backward", "none") else norm
    if centered:
        x = np.fft.ifftshift(x, axes=(-2, -1))
    x = fn(x, axes=(-2, -1), norm=nm)
    if centered:
        x = np.fft.fftshift(x, axes=(-2, -1))
    return x


def phantom(H: int, W: int, slice_index: int = 0) -> np.ndarray:
    """Complex image [H, W]; slices differ by a seeded phase ramp / shift."""
    u = np.linspace(-1, 1, H)[:, None]
    v = np.linspace(-1, 1, W)[None, :]
    rng = np.random.RandomState(1234 + slice_index)
    du, dv, ph = (rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0, 2 * np.pi)) if slice_index else (0, 0, 0)
    uu, vv = u - du, v - dv
    x = np.exp(-2 * ((uu / 0.6) ** 2 + (vv / 0.8) ** 2)) * (1 + 0.3 * np.sin(8 * uu) * np.cos(6 * vv))
    return x * np.exp(1j * (0.5 * uu + ph))


def coil_maps(C: int, H: int, W: int) -> np.ndarray:
    u = np.linspace(-1, 1, H)[None, :, None]
    v = np.linspace(-1, 1, W)[None, None, :]
    a = (2 * np.pi * np.arange(C) / C)[:, None, None]
    S = np.exp(-((u - 0.9 * np.cos(a)) ** 2 + (v - 0.9 * np.sin(a)) ** 2) / 0.8) * np.exp(
        1j * (u * np.cos(a) + v * np.sin(a)))
    return S / np.sqrt((np.abs(S) ** 2).sum(0, keepdims=True))


def make_batch(B: int, C: int, H: int, W: int, mask_func=None, seed: Optional[int] = 123, centered: bool = False,
               normalization: str = "backward", first_slice: int = 0, mask_dtype="uint8"):
    """-> dict of CPU torch tensors: y [B,C,H,W,2], sensitivity_maps [B,C,H,W,2], mask [1,1,1,W,1],
    target [B,H,W] complex64, init_pred [B,H,W,2] (zero-filled SENSE image), kspace (fully sampled)."""
    import torch

    if mask_func is None:
        mask_func = Equispaced1DMask([0.08], [4])
    S = coil_maps(C, H, W)
    mask_t, acc = mask_func((1, H, W, 2), seed)  # [1, W, 1] over the last three dims (H, W, 2) -> [1,W,1]
    m = mask_t.numpy().reshape(1, 1, W)
    ys, ks, tg = [], [], []
    for b in range(B):
        x = phantom(H, W, first_slice + b)
        k = _fft2c(S * x[None], centered, normalization)
        y = k * m + 0.0
        img = _fft2c(y, centered, normalization, inverse=True)
        # transforms.py:526-614 max-normalisation y <- fft2(ifft2(y) / max|ifft2(y)|); the transform is linear, so
        # scaling y directly is the same operator and keeps the unsampled columns exactly zero
        y = y / np.max(np.abs(img))
        ys.append(y)
        ks.append(k)
        tg.append(x / np.max(np.abs(x)))
    Sn = S / np.max(np.abs(S))

    def c2r(a):
        a = np.asarray(a)
        return torch.from_numpy(np.stack((a.real, a.imag), -1).astype(np.float32))

    y = c2r(np.stack(ys))
    sens = c2r(np.broadcast_to(Sn[None], (B, C, H, W)).copy())
    mask = torch.from_numpy(m.reshape(1, 1, 1, W, 1).astype(np.float32))
    if mask_dtype == "uint8":
        mask = mask.byte()  # reconstruction/parts/transforms.py:427 mask.byte()
    elif mask_dtype == "bool":
        mask = mask.bool()
    zf = np.stack([(np.conj(Sn) * _fft2c(yy, centered, normalization, inverse=True)).sum(0) for yy in ys])
    return {
        "y": y, "sensitivity_maps": sens, "mask": mask, "target": torch.from_numpy(np.stack(tg).astype(np.complex64)),
        "init_pred": c2r(zf), "kspace": c2r(np.stack(ks)), "acc": acc,
    }


# cfg dictionaries of the BASELINE.json configurations (hyper-parameters from projects/reconstruction/model_zoo/conf)
def cirim_cfg(recurrent_layer="GRU", num_cascades=5, time_steps=8, centered=False, normalization="backward"):
    return dict(recurrent_layer=recurrent_layer, conv_filters=[64, 64, 2], conv_kernels=[5, 3, 3],
                conv_dilations=[1, 2, 1], conv_bias=[True, True, False], recurrent_filters=[64, 64, 0],
                recurrent_kernels=[1, 1, 0], recurrent_dilations=[1, 1, 0], recurrent_bias=[True, True, False],
                depth=2, time_steps=time_steps, conv_dim=2, num_cascades=num_cascades, dimensionality=2, no_dc=True,
                keep_eta=True, accumulate_estimates=True, train_loss_fn="l1", val_loss_fn="l1",
                coil_combination_method="SENSE", use_sens_net=False, fft_centered=centered,
                fft_normalization=normalization, spatial_dims=[-2, -1], coil_dim=1)


def varnet_cfg(num_cascades=12, channels=14, pooling_layers=2, padding_size=11, centered=False,
               normalization="backward", no_dc=False):
    return dict(num_cascades=num_cascades, channels=channels, pooling_layers=pooling_layers,
                padding_size=padding_size, normalize=True, no_dc=no_dc, train_loss_fn="l1", val_loss_fn="l1",
                coil_combination_method="SENSE", use_sens_net=False, fft_centered=centered,
                fft_normalization=normalization, spatial_dims=[-2, -1], coil_dim=1)


def unet_cfg(channels=64, pooling_layers=2, padding_size=11, centered=False, normalization="backward"):
    return dict(channels=channels, pooling_layers=pooling_layers, padding_size=padding_size, normalize=True,
                train_loss_fn="l1", val_loss_fn="l1", coil_combination_method="SENSE", use_sens_net=False,
                fft_centered=centered, fft_normalization=normalization, spatial_dims=[-2, -1], coil_dim=1)


def zf_cfg(method="SENSE", centered=False, normalization="backward"):
    return dict(coil_combination_method=method, use_sens_net=False, fft_centered=centered,
                fft_normalization=normalization, spatial_dims=[-2, -1], coil_dim=1)

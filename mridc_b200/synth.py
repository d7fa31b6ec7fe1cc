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


# --------------------------------------------------------------------------------------------------
# quantitative (BASELINE.json configs[4]: qCIRIM R2* mapping, 32-coil multi-echo 7T-shaped slices)
# --------------------------------------------------------------------------------------------------
QMRI_SHAPE = (232, 288)  # AHEAD-like in-plane matrix (non-power-of-2: 232 = 8*29, 288 = 32*9)
QMRI_TES = [3.0, 11.5, 20.0, 28.5]  # ms, base_qcirim_run.yaml / qrim/utils.py:64
QMRI_GAMMA = [150.0, 150.0, 1000.0, 150.0]  # base_qcirim_run.yaml:91-95


def qcirim_cfg(num_cascades=1, filters=128, recurrent_layer="IndRNN", time_steps=8, centered=False,
               normalization="backward", sequence="MEGRE"):
    """Flat cfg with the quantitative-module keys of projects/quantitative/model_zoo/conf/base_qcirim_run.yaml:7-121."""
    return dict(
        use_reconstruction_module=False, quantitative_module_recurrent_layer=recurrent_layer,
        quantitative_module_conv_filters=[filters, filters, 4], quantitative_module_conv_kernels=[5, 3, 3],
        quantitative_module_conv_dilations=[1, 2, 1], quantitative_module_conv_bias=[True, True, False],
        quantitative_module_recurrent_filters=[filters, filters, 0], quantitative_module_recurrent_kernels=[1, 1, 0],
        quantitative_module_recurrent_dilations=[1, 1, 0], quantitative_module_recurrent_bias=[True, True, False],
        quantitative_module_depth=2, quantitative_module_time_steps=time_steps, quantitative_module_conv_dim=2,
        quantitative_module_num_cascades=num_cascades, quantitative_module_no_dc=True,
        quantitative_module_keep_eta=True, quantitative_module_accumulate_estimates=True,
        quantitative_module_signal_forward_model_sequence=sequence, quantitative_module_dimensionality=2,
        quantitative_module_gamma_regularization_factors=list(QMRI_GAMMA), shift_B0_input=False, dimensionality=2,
        coil_combination_method="SENSE", use_sens_net=False, fft_centered=centered, fft_normalization=normalization,
        spatial_dims=[-2, -1], coil_dim=2)


def cached_poisson_mask():
    """The 12x Poisson-disc mask of configs[4], generated once by the reference's Poisson2DMaskFunc (its numba RNG is
    unseeded, subsample.py:584-600) and committed bit-packed (oracle/make_golden.py::gen_poisson) -> [H, W] float32."""
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "poisson_mask.npz")
    d = np.load(path)
    H, W = (int(v) for v in d["shape"])
    return np.unpackbits(d["bits"])[: H * W].reshape(H, W).astype(np.float32)


def make_qmri_batch(B: int, C: int = 32, H: Optional[int] = None, W: Optional[int] = None, mask: np.ndarray = None,
                    tes: Sequence[float] = tuple(QMRI_TES), centered: bool = False, normalization: str = "backward",
                    first_slice: int = 0):
    """Synthetic multi-echo GRE slices -> dict of CPU torch tensors: y [B,E,C,H,W,2] (undersampled MEGRE k-space),
    sensitivity_maps [B,C,H,W,2], sampling_mask [B,1,H,W,1] float32, mask_brain [B,1,H,W,1], the ground-truth maps
    R2star / S0 / B0 / phi [B,H,W] and their *_init versions (a smoothed / biased guess, standing in for the
    reference's least-squares initial fit), TEs (list of ms)."""
    import torch

    if H is None or W is None:
        H, W = QMRI_SHAPE
    if mask is None:
        mask = cached_poisson_mask()
        if mask.shape != (H, W):
            raise ValueError("the cached Poisson mask is %s; pass mask= for %s" % (mask.shape, (H, W)))
        if not centered:
            # the pattern is drawn with its calibration disc at the array centre; a non-centred FFT has DC at [0, 0]
            # (utils.apply_mask's `shift` option, common/parts/utils.py:335-336)
            mask = np.fft.ifftshift(mask)
    S = coil_maps(C, H, W)
    S = S / np.max(np.abs(S))
    u = np.linspace(-1, 1, H)[:, None]
    v = np.linspace(-1, 1, W)[None, :]
    maps, inits, ys = [], [], []
    for b in range(B):
        x = phantom(H, W, first_slice + b)
        mag = np.abs(x) / np.max(np.abs(x))
        brain = (mag > 0.05).astype(np.float64)
        r2 = brain * (25.0 + 30.0 * mag + 10.0 * np.sin(5 * u) * np.cos(4 * v))       # 1/s
        b0 = brain * (40.0 * u * v + 15.0 * np.sin(3 * v))                              # rad/s-scaled field map
        s0 = mag * np.cos(np.angle(x))                                                  # Re S0
        ph = mag * np.sin(np.angle(x))                                                  # Im S0
        sig = np.stack([(s0 + 1j * ph) * np.exp(-te * 1e-3 * r2) * np.exp(-1j * b0 * 1e-3 * te) for te in tes])
        k = _fft2c(sig[:, None] * S[None], centered, normalization)                     # [E, C, H, W]
        ys.append(k * mask[None, None])
        maps.append((r2, s0, b0, ph))
        rng = np.random.RandomState(4321 + first_slice + b)
        inits.append(tuple(m * (1.0 + 0.1 * rng.standard_normal()) + 0.02 * np.abs(m).max() * rng.standard_normal(m.shape)
                           for m in (r2, s0, b0, ph)))

    def c2r(a):
        a = np.asarray(a)
        return torch.from_numpy(np.stack((a.real, a.imag), -1).astype(np.float32))

    def f32(a):
        return torch.from_numpy(np.asarray(a).astype(np.float32))

    m5 = np.broadcast_to(mask.reshape(1, 1, H, W, 1), (B, 1, H, W, 1)).astype(np.float32).copy()
    out = {
        "y": c2r(np.stack(ys)), "sensitivity_maps": c2r(np.broadcast_to(S[None], (B, C, H, W)).copy()),
        "sampling_mask": torch.from_numpy(m5), "mask_brain": torch.ones(B, 1, H, W, 1), "TEs": [float(t) for t in tes],
    }
    for i, name in enumerate(("R2star_map", "S0_map", "B0_map", "phi_map")):
        out[name] = f32(np.stack([m[i] for m in maps]))
        out[name + "_init"] = f32(np.stack([m[i] for m in inits]))
    return out

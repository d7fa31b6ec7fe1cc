"""On-device drop-in for ``MRIDataTransforms`` (mridc/collections/reconstruction/parts/transforms.py:16-619; SURVEY.md
section 8 (f) 3) and the reconstruction output layout of the test loop (``models/base.py:576-587``,
``common/parts/utils.py:275-290``).

One slice of raw k-space goes to the GPU once; zero filling, target formation (RSS / SENSE of the fully sampled data),
|x| / max normalisation, image- or k-space cropping, masking and the max-normalisation round trips of
``normalize_inputs`` then run on the device with this package's FFT / coil-combination kernels, so the slice that enters
``CIRIM.forward`` never returns to the host.  Same constructor arguments, call signature and 9-tuple as the reference.
Masks are still drawn on the host by the reference's mask functions (bit-exact inputs).
Noise pre-whitening (``NoisePreWhitening``, transforms.py:622-663) runs on the device too: the C x C whitening matrix from
a noise patch, applied as a 1 x 1 channel-mixing convolution.  Geometric coil compression (transforms.py:666-905; an SVD per
read-out position) is not built.
"""
import math
import os
from collections import defaultdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _ops, fft, utils

__all__ = ["MRIDataTransforms", "NoisePreWhitening", "assemble_reconstructions", "save_reconstructions"]


def _unset(v) -> bool:
    return v is None or v in ("", "None")


def _device() -> torch.device:
    return torch.device("cuda", torch.cuda.current_device())


class NoisePreWhitening:
    """transforms.py:622-663: coil decorrelation.  psi = inv(chol(N N^T / (n - 1))) * sqrt(2 scale) from the noise patch
    ``data[:, x0:x1, y0:y1]`` of the (re, im)-stacked view -- a REAL C x C matrix, as upstream computes it -- then
    ``psi @ data``.  The C x C factorisation is a few hundred flops (torch.linalg on the device, no host round trip); the
    product over the slice is this package's convolution kernel with psi as a 1 x 1 kernel."""

    def __init__(self, patch_size: List[int], scale_factor: float = 1.0):
        self.patch_size = patch_size
        self.scale_factor = scale_factor

    @torch.no_grad()
    def __call__(self, data):
        if not self.patch_size:
            raise ValueError("Patch size must be defined for noise prewhitening.")
        if data.shape[-1] != 2:
            data = torch.view_as_real(data)
        noise = data[:, self.patch_size[0]: self.patch_size[1], self.patch_size[-2]: self.patch_size[-1]]
        noise_int = torch.reshape(noise, (noise.shape[0], int(torch.numel(noise) / noise.shape[0])))
        deformation_matrix = (1 / (float(noise_int.shape[1]) - 1)) * torch.mm(noise_int, noise_int.t())
        psi = torch.linalg.inv(torch.linalg.cholesky(deformation_matrix)) * math.sqrt(2) * math.sqrt(self.scale_factor)
        C = data.shape[0]
        flat = data.contiguous().reshape(1, C, data.shape[1], -1)  # pointwise: any 2-D arrangement of the samples works
        out = _ops.conv2d(flat, psi.reshape(C, C, 1, 1).contiguous(), None, 1, 1, _ops.PAD_ZERO)
        return out.reshape(data.shape)


class MRIDataTransforms:
    def __init__(self, apply_prewhitening: bool = False, prewhitening_scale_factor: float = 1.0,
                 prewhitening_patch_start: int = 10, prewhitening_patch_length: int = 30, apply_gcc: bool = False,
                 gcc_virtual_coils: int = 10, gcc_calib_lines: int = 24, gcc_align_data: bool = True,
                 coil_combination_method: str = "SENSE", dimensionality: int = 2, mask_func: Optional[List] = None,
                 shift_mask: bool = False, mask_center_scale: Optional[float] = 0.02, half_scan_percentage: float = 0.0,
                 remask: bool = False, crop_size: Optional[Tuple[int, int]] = None, kspace_crop: bool = False,
                 crop_before_masking: bool = True, kspace_zero_filling_size: Optional[Tuple] = None,
                 normalize_inputs: bool = False, fft_centered: bool = True, fft_normalization: str = "ortho",
                 max_norm: bool = True, spatial_dims: Sequence[int] = None, coil_dim: int = 0, use_seed: bool = True):
        if apply_gcc:
            raise NotImplementedError("mridc_b200: geometric coil compression is not built")
        if dimensionality != 2:
            raise NotImplementedError("mridc_b200: MRIDataTransforms handles dimensionality == 2 (one slice per call)")
        self.coil_combination_method = coil_combination_method
        self.dimensionality = dimensionality
        self.mask_func = mask_func
        self.shift_mask = shift_mask
        self.mask_center_scale = mask_center_scale
        self.half_scan_percentage = half_scan_percentage
        self.remask = remask
        self.crop_size = crop_size
        self.kspace_crop = kspace_crop
        self.crop_before_masking = crop_before_masking
        self.kspace_zero_filling_size = kspace_zero_filling_size
        self.normalize_inputs = normalize_inputs
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.max_norm = max_norm
        self.spatial_dims = spatial_dims if spatial_dims is not None else [-2, -1]
        self.coil_dim = coil_dim - 1  # transforms.py:130 (2-D data: the batch axis is not there yet)
        self.apply_prewhitening = apply_prewhitening
        self.prewhitening = NoisePreWhitening(  # transforms.py:132-143
            patch_size=[prewhitening_patch_start, prewhitening_patch_length + prewhitening_patch_start,
                        prewhitening_patch_start, prewhitening_patch_length + prewhitening_patch_start],
            scale_factor=prewhitening_scale_factor) if apply_prewhitening else None
        self.gcc = None
        self.use_seed = use_seed

    # ---- helpers ----------------------------------------------------------------------------------------------------
    def _f(self, x):
        return fft.fft2(x, centered=self.fft_centered, normalization=self.fft_normalization, spatial_dims=self.spatial_dims)

    def _i(self, x):
        return fft.ifft2(x, centered=self.fft_centered, normalization=self.fft_normalization, spatial_dims=self.spatial_dims)

    def _crop_image(self, x):  # image-space tensor; k-space crop when kspace_crop (:323-360)
        return self._i(utils.complex_center_crop(self._f(x), self.crop_size)) if self.kspace_crop \
            else utils.complex_center_crop(x, self.crop_size)

    def _crop_kspace(self, k):  # k-space tensor (:362-383, :505-546)
        return utils.complex_center_crop(k, self.crop_size) if self.kspace_crop \
            else self._f(utils.complex_center_crop(self._i(k), self.crop_size))

    def _max_normalise(self, k):
        """k-space -> image / max|image| -> k-space (:548-616); the un-normalised "none" mode uses plain DFT scaling."""
        if self.fft_normalization in ("backward", "ortho", "forward"):
            im = self._i(k)
            if self.max_norm:
                im = im / torch.max(torch.abs(im))
            return self._f(im)
        if self.fft_normalization in ("none", None) and self.max_norm:
            im = fft.ifft2(k, centered=False, normalization="backward", spatial_dims=self.spatial_dims)
            im = im / torch.max(utils.complex_abs(im))  # upstream divides a complex tensor here: the magnitude
            return fft.fft2(im, centered=False, normalization="backward", spatial_dims=self.spatial_dims)
        return k

    # ---- the transform ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def __call__(self, kspace: np.ndarray, sensitivity_map: np.ndarray, mask: np.ndarray, eta: np.ndarray,
                 target: np.ndarray, attrs: Dict, fname: str, slice_idx: int):
        dev = _device()
        kspace = utils.to_tensor(kspace).to(dev)
        have_sens = sensitivity_map is not None and sensitivity_map.size != 0
        if have_sens:
            sensitivity_map = utils.to_tensor(sensitivity_map).to(dev)
        if self.apply_prewhitening:  # :223-224
            kspace = self.prewhitening(kspace)

        if not _unset(self.kspace_zero_filling_size):  # :229-262
            top = int(np.floor_divide(abs(int(self.kspace_zero_filling_size[0]) - kspace.shape[1]), 2))
            left = int(np.floor_divide(abs(int(self.kspace_zero_filling_size[1]) - kspace.shape[2]), 2))
            pad = (0, 0, left, left, top, top)
            kspace = torch.nn.functional.pad(kspace, pad=pad, mode="constant", value=0)
            sensitivity_map = self._i(torch.nn.functional.pad(self._f(sensitivity_map), pad=pad, mode="constant", value=0))

        eta = utils.to_tensor(eta).to(dev) if eta is not None and eta.size != 0 else torch.tensor([])

        method = self.coil_combination_method.upper()  # :270-298
        if method == "RSS":
            target = utils.rss(self._i(kspace), dim=self.coil_dim)
        elif method == "SENSE":
            if have_sens:
                target = utils.sense(self._i(kspace), sensitivity_map, dim=self.coil_dim)
            else:
                target = utils.to_tensor(target).to(dev)
        elif target is not None and target.size != 0:
            target = utils.to_tensor(target).to(dev)
        elif "target" in attrs or "target_rss" in attrs:
            target = torch.tensor(attrs["target"]).to(dev)
        else:
            raise ValueError("No target found")
        target = torch.view_as_complex(target.contiguous())
        target = torch.abs(target / torch.max(torch.abs(target)))

        seed = tuple(map(ord, fname)) if self.use_seed else None
        acq_start = attrs["padding_left"] if "padding_left" in attrs else 0
        acq_end = attrs["padding_right"] if "padding_left" in attrs else 0

        cropping = not _unset(self.crop_size)
        if cropping:  # :309-360; the clamped size sticks to the instance like upstream
            h = min(int(self.crop_size[0]), target.shape[0])
            w = min(int(self.crop_size[1]), target.shape[1])
            self.crop_size = (int(h), int(w))
            target = utils.center_crop(target, self.crop_size)
            if have_sens:
                sensitivity_map = self._crop_image(sensitivity_map)
            if eta is not None and eta.ndim > 2:
                eta = self._crop_image(eta)
        if cropping and self.crop_before_masking:
            kspace = self._crop_kspace(kspace)

        if not utils.is_none(mask):  # :385-407 precomputed masks
            for _mask in mask:
                if list(_mask.shape) == [kspace.shape[-3], kspace.shape[-2]]:
                    mask = torch.from_numpy(_mask).unsqueeze(0).unsqueeze(-1)
                    break
            if (not utils.is_none(acq_start) and not utils.is_none(acq_end)) and acq_start != 0:
                mask[:, :, :acq_start] = 0
                mask[:, :, acq_end:] = 0
            if isinstance(mask, np.ndarray):
                mask = torch.from_numpy(mask).unsqueeze(0).unsqueeze(-1)
            mask = mask.to(dev)
            if self.shift_mask:
                mask = torch.fft.fftshift(mask, dim=(self.spatial_dims[0] - 1, self.spatial_dims[1] - 1))
            if cropping and self.crop_before_masking:
                mask = utils.complex_center_crop(mask, self.crop_size)
            masked_kspace = kspace * mask + 0.0
            acc = 1
        elif utils.is_none(self.mask_func):  # :408-435 fully sampled
            masked_kspace = kspace.clone()
            acc = torch.tensor([1])
            mask = torch.ones(masked_kspace.shape[-3], masked_kspace.shape[-2], dtype=torch.float32, device=dev)
            if cropping:
                mask = utils.center_crop(mask, self.crop_size)
            mask = mask.unsqueeze(0).unsqueeze(-1)
            if self.shift_mask:
                mask = torch.fft.fftshift(mask, dim=(1, 2))
            masked_kspace = masked_kspace * mask
            mask = mask.byte()
        elif isinstance(self.mask_func, list):  # :436-484
            masked_kspace, mask, acc = [], [], []
            for m in self.mask_func:
                _y, _m, _a = utils.apply_mask(kspace, m, seed, (acq_start, acq_end), shift=self.shift_mask,
                                              half_scan_percentage=self.half_scan_percentage,
                                              center_scale=self.mask_center_scale)
                masked_kspace.append(_y)
                mask.append(_m.byte())
                acc.append(_a)
        else:  # :485-496
            masked_kspace, mask, acc = utils.apply_mask(kspace, self.mask_func[0], seed, (acq_start, acq_end),
                                                        shift=self.shift_mask,
                                                        half_scan_percentage=self.half_scan_percentage,
                                                        center_scale=self.mask_center_scale)
            mask = mask.byte()

        if cropping and not self.crop_before_masking:  # :498-546
            kspace = self._crop_kspace(kspace)
            masked_kspace = self._crop_kspace(masked_kspace)
            mask = utils.center_crop(mask.squeeze(-1), self.crop_size).unsqueeze(-1)

        if self.normalize_inputs:  # :548-616
            kspace = self._max_normalise(kspace)
            if isinstance(masked_kspace, list):
                masked_kspace = [self._max_normalise(y) for y in masked_kspace]
            else:
                masked_kspace = self._max_normalise(masked_kspace)
            if self.max_norm:
                if have_sens:
                    sensitivity_map = sensitivity_map / torch.max(torch.abs(sensitivity_map))
                if eta.ndim > 2:
                    eta = eta / torch.max(torch.abs(eta))
                target = target / torch.max(torch.abs(target))

        return kspace, masked_kspace, sensitivity_map, mask, eta, target, fname, slice_idx, acc


def assemble_reconstructions(outputs) -> Dict[str, np.ndarray]:
    """``models/base.py:576-582``: (fname, slice_num, output) triples -> {fname: [slices, ...] stacked in slice order}."""
    per_file = defaultdict(list)
    for fname, slice_num, output in outputs:
        if isinstance(output, torch.Tensor):
            output = output.detach().cpu().numpy()
        per_file[fname].append((int(slice_num), output))
    return {fname: np.stack([out for _, out in sorted(items, key=lambda t: t[0])]) for fname, items in per_file.items()}


def save_reconstructions(reconstructions: Dict[str, np.ndarray], out_dir):
    """``common/parts/utils.py:275-290`` / ``models/base.py:583-587``: one h5 file per input file with the dataset
    ``reconstruction``.  h5py is the reference's own dependency for this step; it is imported here, not at package import."""
    try:
        import h5py
    except ImportError as e:  # pragma: no cover - depends on the installation
        raise ImportError("mridc_b200.save_reconstructions writes the reference's h5 layout and needs h5py") from e
    if not hasattr(h5py, "File"):  # a stub module registered under the name (the oracle's import recipe does that)
        raise ImportError("mridc_b200.save_reconstructions writes the reference's h5 layout and needs h5py")
    out_dir = os.fspath(out_dir)
    os.makedirs(out_dir, exist_ok=True)
    for fname, recons in reconstructions.items():
        with h5py.File(os.path.join(out_dir, fname), "w") as hf:
            hf.create_dataset("reconstruction", data=recons)

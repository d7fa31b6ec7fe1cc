"""Oracle (TEST INFRASTRUCTURE): MSE / NMSE / PSNR / SSIM as used by the reference's test_step.

Call sites: mridc/collections/common/metrics/reconstruction_metrics.py:11-41, used from
reconstruction/models/base.py:415-436 on ``|pred|/max`` and ``|target|/max``.  The arithmetic lives in
scikit-image (``scikit-image>=0.18.3``, unpinned, not vendored, not installed here); its published
algorithm (Wang et al. 2004 as implemented by ``skimage.metrics.structural_similarity`` with
``win_size=7``, uniform filter, ``use_sample_covariance=True``, K1=0.01, K2=0.03, 3-px border crop) is
restated on scipy.  No reference test touches the metrics -> PARITY UNPINNED for this file.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def mse(gt, pred):
    """reconstruction_metrics.py:11-13."""
    return np.mean((gt - pred) ** 2)


def nmse(gt, pred):
    """reconstruction_metrics.py:16-18."""
    return np.linalg.norm(gt - pred) ** 2 / np.linalg.norm(gt) ** 2


def psnr(gt, pred, maxval=None):
    """reconstruction_metrics.py:21-25 -> skimage.peak_signal_noise_ratio = 10 log10(R^2 / mse)."""
    if maxval is None:
        maxval = np.max(gt)
    err = np.mean((np.asarray(gt, dtype=np.float64) - np.asarray(pred, dtype=np.float64)) ** 2)
    return 10 * np.log10((maxval**2) / err)


def _ssim2d(im1, im2, data_range, win_size=7, K1=0.01, K2=0.03):
    im1 = im1.astype(np.float64)
    im2 = im2.astype(np.float64)
    NP = win_size**2
    cov_norm = NP / (NP - 1)
    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)
    C1 = (K1 * data_range) ** 2
    C2 = (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux**2 + uy**2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return S[pad:-pad, pad:-pad].mean(dtype=np.float64)


def ssim(gt, pred, maxval=None):
    """reconstruction_metrics.py:28-41: mean over the leading (slice) dim of 2-D SSIMs."""
    if gt.ndim != 3:
        raise ValueError("Unexpected number of dimensions in ground truth.")
    if gt.ndim != pred.ndim:
        raise ValueError("Ground truth dimensions does not match pred.")
    maxval = np.max(gt) if maxval is None else maxval
    return sum(_ssim2d(gt[i], pred[i], maxval) for i in range(gt.shape[0])) / gt.shape[0]

"""Drop-ins for the data-consistency layers of ``mridc/collections/reconstruction/models/sigmanet/dc_layers.py``
(other consumers of the DC operator): ``DataIDLayer`` :15, ``DataGDLayer`` :22-96,
``DataProxCGLayer`` / ``ConjugateGradient`` :98-325, ``DataVSLayer`` :327-410, ``DCLayer`` :416-478.

The reference writes these layers against ``[..., sets, coils, H, W, 2]`` tensors of the sigmanet code base and calls them
from ``DUNet`` with ``[B, C, H, W, 2]`` maps and a ``[B, H, W, 2]`` image (``dunet.py:177-186``), where ``unsqueeze(-5)`` /
``sum(-4, keepdim)`` / ``sum(-5)`` land on the batch and coil axes.  The layers below keep exactly those axis semantics:
they are compositions of this package's CUDA operators (centred FFT, strided-broadcast complex multiply, fused
expand / reduce where the operands have the fastMRI layout) with the reference's reductions, inference only.
"""
from typing import Optional, Tuple

import torch

from . import fft, utils

__all__ = ["DataIDLayer", "DataGDLayer", "DataProxCGLayer", "ConjugateGradient", "DataVSLayer", "DCLayer"]


class DataIDLayer(torch.nn.Module):
    """dc_layers.py:15-19: placeholder, no parameters and no forward."""

    def __init__(self, *args, **kwargs):
        super().__init__()


def _scalar_param(value):
    p = torch.nn.Parameter(torch.Tensor(1))
    p.data = torch.tensor(value, dtype=p.dtype)
    return p


class _FFTCfg(torch.nn.Module):
    def _cfg(self, fft_centered, fft_normalization, spatial_dims):
        self.fft_centered = fft_centered
        self.fft_normalization = fft_normalization
        self.spatial_dims = spatial_dims if spatial_dims is not None else [-2, -1]

    def _fft(self, x):
        return fft.fft2(x, centered=self.fft_centered, normalization=self.fft_normalization, spatial_dims=self.spatial_dims)

    def _ifft(self, x):
        return fft.ifft2(x, centered=self.fft_centered, normalization=self.fft_normalization, spatial_dims=self.spatial_dims)


class DataGDLayer(_FFTCfg):
    """dc_layers.py:22-96: x - lambda * A^H(mask * (sum_{-4} mask * A x - y))."""

    def __init__(self, lambda_init, learnable=True, fft_centered: bool = True, fft_normalization: str = "ortho",
                 spatial_dims: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.lambda_init = lambda_init
        self.data_weight = _scalar_param(lambda_init)
        self.data_weight.requires_grad = learnable
        self._cfg(fft_centered, fft_normalization, spatial_dims)

    @torch.no_grad()
    def forward(self, x, y, smaps, mask):
        # :69-82
        A_x_y = torch.sum(self._fft(utils.complex_mul(x.unsqueeze(-5).expand_as(smaps), smaps)) * mask, -4, keepdim=True) - y
        # :83-95
        gradD_x = torch.sum(utils.complex_mul(self._ifft(A_x_y * mask), smaps, _conj_y=True), dim=(-5))
        return x - self.data_weight * gradD_x


class ConjugateGradient:
    """dc_layers.py:157-255, forward half (inference): conjugate gradients on (lambda A^H A + I) x = lambda A^H y + z in
    complex arithmetic (the reference carries (re, im) pairs; its step alpha = rr conj(<p,q>) / |<p,q>|^2 is rr / <p,q>).
    The reference wraps the solver in an autograd Function; ``apply`` is kept as the entry point."""

    @staticmethod
    def complexDot(data1, data2):
        """:161-166: per-sample sum of data1 * conj(data2) as (re, im)."""
        n = data1.shape[0]
        d = (torch.view_as_complex(data1.contiguous()) * torch.view_as_complex(data2.contiguous()).conj()).reshape(n, -1).sum(-1)
        return torch.view_as_real(d)

    @staticmethod
    def solve(x0, M, tol, max_iter):
        """:168-196.  One host read per iteration for the stopping test, like the reference's ``while``."""
        n = x0.shape[0]
        b = torch.view_as_complex(x0.contiguous())
        bshape = (n,) + (1,) * (b.dim() - 1)
        x = torch.zeros_like(b)
        r, p = b.clone(), b.clone()
        bb = (x0 * x0).reshape(n, -1).sum(-1)
        rr = bb.clone()
        for _ in range(max_iter):
            if not bool(torch.min(rr / bb) > tol):
                break
            q = torch.view_as_complex(M(torch.view_as_real(p)).contiguous())
            pq = (p * q.conj()).reshape(n, -1).sum(-1)
            alpha = (rr / pq).reshape(bshape)
            x = x + alpha * p
            r = r - alpha * q
            rr_new = torch.view_as_real(r).pow(2).reshape(n, -1).sum(-1)
            p = r + (rr_new / rr).reshape(bshape) * p
            rr = rr_new
        return torch.view_as_real(x)

    @staticmethod
    @torch.no_grad()
    def apply(z, lambdaa, y, smaps, mask, tol, max_iter, fft_centered, fft_normalization, spatial_dims):
        """:198-255."""
        kw = dict(centered=fft_centered, normalization=fft_normalization, spatial_dims=spatial_dims)

        def normal(p):  # lambda A^H A p + p, A = sum_{-4} mask F S, A^H = sum_{-5} conj(S) F^-1 mask   (:222-246)
            k = torch.sum(fft.fft2(utils.complex_mul(p.expand_as(smaps), smaps), **kw) * mask, dim=-4, keepdim=True)
            return lambdaa * _adjoint(k, smaps, mask, kw) + p

        return ConjugateGradient.solve(lambdaa * _adjoint(y, smaps, mask, kw) + z, normal, tol, max_iter)


def _adjoint(k, smaps, mask, kw):
    return torch.sum(utils.complex_mul(fft.ifft2(k * mask, **kw), smaps, _conj_y=True), dim=-5)


class DataProxCGLayer(_FFTCfg):
    """dc_layers.py:98-154: prox of the data term by conjugate gradient (Aggarwal et al.)."""

    def __init__(self, lambda_init, tol=1e-6, iter=10, learnable=True, fft_centered: bool = True,
                 fft_normalization: str = "ortho", spatial_dims: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.lambdaa = torch.nn.Parameter(torch.Tensor(1))
        self.lambdaa.data = torch.tensor(lambda_init)
        self.lambdaa_init = lambda_init
        self.lambdaa.requires_grad = learnable
        self.tol = tol
        self.iter = iter
        self.op = ConjugateGradient
        self._cfg(fft_centered, fft_normalization, spatial_dims)

    def forward(self, x, f, smaps, mask):
        return self.op.apply(x, self.lambdaa.detach(), f, smaps, mask, self.tol, self.iter, self.fft_centered,
                             self.fft_normalization, self.spatial_dims)

    def set_learnable(self, flag):
        self.lambdaa.requires_grad = flag


class DataVSLayer(_FFTCfg):
    """dc_layers.py:327-410: variable-splitting data consistency + weighted averaging."""

    def __init__(self, alpha_init, beta_init, learnable=True, fft_centered: bool = True, fft_normalization: str = "ortho",
                 spatial_dims: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.alpha = _scalar_param(alpha_init)
        self.beta = _scalar_param(beta_init)
        self.learnable = learnable
        self.set_learnable(learnable)
        self._cfg(fft_centered, fft_normalization, spatial_dims)

    @torch.no_grad()
    def forward(self, x, y, smaps, mask):
        # :374-383
        A_x = torch.sum(self._fft(utils.complex_mul(x.unsqueeze(-5).expand_as(smaps), smaps)), -4, keepdim=True)
        # :384
        k_dc = (1 - mask) * A_x + mask * (self.alpha * A_x + (1 - self.alpha) * y)
        # :385-396
        x_dc = torch.sum(utils.complex_mul(self._ifft(k_dc), smaps, _conj_y=True), dim=(-5))
        return self.beta * x + (1 - self.beta) * x_dc

    def set_learnable(self, flag):
        self.learnable = flag
        self.alpha.requires_grad = self.learnable
        self.beta.requires_grad = self.learnable


class DCLayer(_FFTCfg):
    """dc_layers.py:416-478: single-coil data consistency of DC-CNN."""

    def __init__(self, lambda_init=0.0, learnable=True, fft_centered: bool = True, fft_normalization: str = "ortho",
                 spatial_dims: Optional[Tuple[int, int]] = None):
        super().__init__()
        self.lambda_ = _scalar_param(lambda_init)
        self.learnable = learnable
        self.set_learnable(learnable)
        self._cfg(fft_centered, fft_normalization, spatial_dims)

    @torch.no_grad()
    def forward(self, x, y, mask):
        A_x = self._fft(x)
        k_dc = (1 - mask) * A_x + mask * (self.lambda_ * A_x + (1 - self.lambda_) * y)
        return self._ifft(k_dc)

    def set_learnable(self, flag):
        self.learnable = flag
        self.lambda_.requires_grad = self.learnable

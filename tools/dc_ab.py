"""A/B timing of the W = 320 hybrid DC gradient kernels (run once per variant: MRIDC_B200_DC_V3=1 selects the half-warp / lane-FFT
experiment).  Inputs stream from HBM (B = 16: 393 MB of S and yh) like in the bench."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _ops
C, H, W = 15, 320, 320
dev = torch.device("cuda")
def t(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tag = "v3 (lane FFT)" if os.environ.get("MRIDC_B200_DC_V3") else "v2 (smem transposes)"
for B in (4, 16):
    y = torch.randn(B, C, H, W, 2, device=dev); S = torch.randn(B, C, H, W, 2, device=dev)
    eta = torch.randn(B, H, W, 2, device=dev)
    mask = torch.zeros(1, 1, 1, W, 1, device=dev); mask[..., ::4, :] = 1; mask[..., 147:173, :] = 1
    y = y * mask
    out = torch.empty((B, H, W, 4), device=dev)
    bytes_alg = B * (2 * C * H * W * 8 + 3 * H * W * 8) + W
    yh = _ops.dc_hybrid_prepare(y, mask, False)
    us = t(lambda: _ops.dc_rim_grad(eta, y, S, mask, 1.0, False, "backward", out=out, nhwc=True, y_hybrid=yh))
    print("%s B=%d: %7.1f us  -> %6.0f GB/s algorithmic (%.3f of 6457)" % (tag, B, us, bytes_alg / us / 1e3, bytes_alg / us / 1e3 / 6457))

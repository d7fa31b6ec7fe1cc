"""Time the fused DC operators (CUDA events, L2-warm steady state as inside the CIRIM loop)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _ops
C, H, W = 15, 320, 320
dev = torch.device("cuda")
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for B in (1, 4, 8):
    y = torch.randn(B, C, H, W, 2, device=dev); S = torch.randn(B, C, H, W, 2, device=dev)
    eta = torch.randn(B, H, W, 2, device=dev)
    mask = (torch.rand(1, 1, 1, W, 1, device=dev) < 0.25).to(torch.uint8)
    ws = torch.empty((2, B, C, H, W, 2), device=dev); out = torch.empty((B, H, W, 4), device=dev)
    bytes_alg = B * (2 * C * H * W * 8 + 3 * H * W * 8) + W
    for cen, nrm in ((False, "backward"), (True, "ortho")):
        us = t(lambda: _ops.dc_rim_grad(eta, y, S, mask, 1.0, cen, nrm, out=out, ws=ws, nhwc=True))
        print("B=%d centered=%s rim_grad (3-pass) %7.1f us  -> %6.0f GB/s algorithmic" % (B, cen, us, bytes_alg / us / 1e3))
        yh = _ops.dc_hybrid_prepare(y, mask, cen)
        us = t(lambda: _ops.dc_rim_grad(eta, y, S, mask, 1.0, cen, nrm, out=out, nhwc=True, y_hybrid=yh))
        print("B=%d centered=%s rim_grad (hybrid) %7.1f us  -> %6.0f GB/s algorithmic" % (B, cen, us, bytes_alg / us / 1e3))
        us = t(lambda: _ops.dc_hybrid_prepare(y, mask, cen, ws=ws[0]))
        print("B=%d centered=%s hybrid prepare   %7.1f us (once per slice batch)" % (B, cen, us))
    m2 = (torch.rand(1, 1, H, W, 1, device=dev) < 0.25).to(torch.uint8)  # 2-D mask: the general three-pass operator
    us = t(lambda: _ops.dc_rim_grad(eta, y, S, m2, 1.0, True, "ortho", out=out, ws=ws, nhwc=True))
    os.environ["MRIDC_B200_DC_STOCKHAM"] = "1"
    us_s = t(lambda: _ops.dc_rim_grad(eta, y, S, m2, 1.0, True, "ortho", out=out, ws=ws, nhwc=True))
    del os.environ["MRIDC_B200_DC_STOCKHAM"]
    print("B=%d 2-D mask rim_grad (3-pass, 320-point kernels) %7.1f us -> %6.0f GB/s algorithmic   (Stockham kernels: %7.1f us)"
          % (B, us, bytes_alg / us / 1e3, us_s))
    us = t(lambda: _ops.sens_reduce(y, S, False, "backward", ws=ws))
    print("B=%d sens_reduce %7.1f us" % (B, us))
    us = t(lambda: _ops.sens_expand_softdc(eta, S, None, None, None, None, None, True, False, "backward", ws=ws))
    print("B=%d sens_expand %7.1f us" % (B, us))

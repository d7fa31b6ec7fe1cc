"""A/B of the W = 320 hybrid DC gradient kernel's output modes in one process, interleaved: fp32 channels-last (OUT_MODE 2)
vs the G8 split-bf16 conv input with its replicate border (OUT_MODE 3, what the time loop launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _lib, _ops
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
y = torch.randn(B, C, H, W, 2, device=dev); S = torch.randn(B, C, H, W, 2, device=dev)
eta = torch.randn(B, H, W, 2, device=dev)
mask = torch.zeros(1, 1, 1, W, 1, device=dev); mask[..., ::4, :] = 1; mask[..., 147:173, :] = 1
y = y * mask
out2 = torch.empty((B, H, W, 4), device=dev)
out3 = torch.zeros(_lib.load().mrb_g8_bytes(B, H, W), dtype=torch.uint8, device=dev)
yh = _ops.dc_hybrid_prepare(y, mask, False)
alg = B * (2 * C * H * W * 8 + 3 * H * W * 8) + W
for rnd in range(3):
    u2 = t(lambda: _ops.dc_rim_grad(eta, y, S, mask, 1.0, False, "backward", out=out2, nhwc=True, y_hybrid=yh))
    u3 = t(lambda: _ops.dc_rim_grad(eta, y, S, mask, 1.0, False, "backward", out=out3, nhwc=2, y_hybrid=yh))
    print("B=%d round %d: fp32 out %.1f us (%.0f GB/s)   G8 out %.1f us (%.0f GB/s)" % (B, rnd, u2, alg / u2 / 1e3, u3, alg / u3 / 1e3))

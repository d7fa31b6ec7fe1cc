"""InstanceNorm + LeakyReLU: single-pass cluster kernel vs the two-kernel form vs torch, over shapes / strides / in-place."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _lib
lib = _lib.load(); st = _lib.stream_ptr()
def run(x, xbs, out, obs, N, C, HW):
    stats = torch.empty(2 * N * C, dtype=torch.float64, device="cuda")
    _lib.check(lib.mrb_instnorm_lrelu(_lib.ptr(x), xbs, _lib.ptr(out), obs, N, C, HW, 1e-5, 0.2, _lib.ptr(stats), st))
torch.manual_seed(0)
for N, C, H, W, extra in ((2, 3, 37, 45, 0), (2, 14, 320, 320, 0), (1, 28, 160, 160, 28), (3, 5, 7, 9, 2), (1, 4, 64, 48, 0), (2, 56, 80, 80, 0)):
    HW = H * W
    full = torch.randn(N, C + extra, H, W, device="cuda") * 3 + 0.7
    x = full[:, extra:]
    ref = torch.nn.functional.leaky_relu(torch.nn.functional.instance_norm(x.double(), eps=1e-5), 0.2).float()
    out = torch.empty(N, C, H, W, device="cuda")
    run(x, (C + extra) * HW, out, C * HW, N, C, HW)
    e1 = ((out - ref).norm() / ref.norm()).item()
    xin = x.clone().contiguous(); run(xin, C * HW, xin, C * HW, N, C, HW)
    e2 = ((xin - ref).norm() / ref.norm()).item()
    print("N=%d C=%d %dx%d extra=%d: rel-L2 out-of-place %.2e in-place %.2e  nan=%s" % (N, C, H, W, extra, e1, e2, bool(torch.isnan(out).any())))

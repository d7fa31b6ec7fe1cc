import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _lib
lib = _lib.load(); st = _lib.stream_ptr()
torch.manual_seed(0)
def pack_conv(w, k):
    p = torch.empty(lib.mrb_tc_packed_floats(0, w.shape[0], 64, k), device="cuda")
    _lib.check(lib.mrb_tc_pack_conv(_lib.ptr(w), _lib.ptr(p), w.shape[0], 64, k, st)); return p
B,H,W = 1,16,8
x = torch.arange(B*H*W*64, dtype=torch.float32, device="cuda").reshape(B,H,W,64) * 0.001
# 1) identity 1x1
w = torch.eye(64, device="cuda").reshape(64,64,1,1).contiguous()
out = torch.full((B,H,W,64), -7.0, device="cuda")
_lib.check(lib.mrb_tc_conv_nhwc(_lib.ptr(x), _lib.ptr(pack_conv(w,1)), None, _lib.ptr(out), B,H,W,64,1,1,0,st))
torch.cuda.synchronize()
d = (out - x).abs()
print("identity 1x1: max err", d.max().item(), "n bad", (d > 1e-4).sum().item(), "of", d.numel())
xf = x.reshape(-1,64); of = out.reshape(-1,64)
for p in (0,1,7,8,9,127):
    print("p", p, "x", [round(v,3) for v in xf[p,:10].tolist()], "\n     out", [round(v,3) for v in of[p,:10].tolist()], " cols32..35", [round(v,3) for v in of[p,32:36].tolist()])
# which (pixel, ch) pairs are wrong?
bad = (d.reshape(-1,64) > 1e-4)
print("bad per channel:", bad.sum(0).tolist())
print("bad per pixel (first 32):", bad.sum(1)[:32].tolist())
# 2) random 1x1
w = torch.randn(64,64,1,1, device="cuda")*0.1
xr = torch.randn(B,H,W,64, device="cuda")
_lib.check(lib.mrb_tc_conv_nhwc(_lib.ptr(xr), _lib.ptr(pack_conv(w,1)), None, _lib.ptr(out), B,H,W,64,1,1,0,st))
ref = (xr.reshape(-1,64).double() @ w.reshape(64,64).double().t()).reshape(B,H,W,64)
print("random 1x1 rel err", ((out.double()-ref).norm()/ref.norm()).item())

"""One-shot process: first TC GRU launch of the process at the tiny size; report bad elements."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _lib
from oracle import nets as onets
lib = _lib.load(); st = _lib.stream_ptr()
B, H, W = 1, 16, 8
g = torch.Generator().manual_seed(16)
x = torch.randn(B, 64, H, W, generator=g); h = torch.randn(B, 64, H, W, generator=g)
wih = torch.randn(192, 64, 1, 1, generator=g) * 0.1; whh = torch.randn(192, 64, 1, 1, generator=g) * 0.1
bih = torch.randn(192, generator=g)
ref = onets.conv_gru_cell(x, h, wih, bih, whh, 1, 1)
xd = x.permute(0, 2, 3, 1).contiguous().cuda(); hd = h.permute(0, 2, 3, 1).contiguous().cuda()
wi, wh, bd = wih.cuda().contiguous(), whh.cuda().contiguous(), bih.cuda()
pg = torch.empty(lib.mrb_tc_packed_floats(1, 64, 64, 1), device="cuda")
_lib.check(lib.mrb_tc_pack_gru(_lib.ptr(wi), _lib.ptr(wh), _lib.ptr(pg), 64, 64, st))
out = torch.empty(B, H, W, 64, device="cuda")
_lib.check(lib.mrb_tc_gru_nhwc(_lib.ptr(xd), _lib.ptr(hd), _lib.ptr(pg), _lib.ptr(bd), _lib.ptr(out), B, H, W, 64, st))
o = out.permute(0, 3, 1, 2).cpu()
e = ((o.double() - ref.double()).norm() / ref.double().norm()).item()
d = (o - ref).abs()
idx = (d > 1e-5).nonzero()
print("err %.2e n>1e-5: %d" % (e, idx.shape[0]), "chans", sorted(set(i[1].item() for i in idx))[:70], "pix", sorted(set(i[2].item() * W + i[3].item() for i in idx))[:130])

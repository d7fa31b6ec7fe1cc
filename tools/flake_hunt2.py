import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_tc as T
from mridc_b200 import _lib
from oracle import nets as onets
lib = _lib.load(); st = _lib.stream_ptr()
B, H, W = 1, 16, 8
nbad = 0
for it in range(150):
    g = torch.Generator().manual_seed(H)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().cuda()
    x4 = torch.randn(B, 4, H, W, generator=g); w1 = torch.randn(64, 4, 5, 5, generator=g) * 0.2; b1 = torch.randn(64, generator=g)
    out = torch.empty(B, H, W, 64, device="cuda")
    x4d, p1, b1d = nhwc(x4), T._pack(2, w1.cuda()), b1.cuda()
    _lib.check(lib.mrb_tc_conv5x5x4_nhwc(_lib.ptr(x4d), _lib.ptr(p1), _lib.ptr(b1d), _lib.ptr(out), B, H, W, 64, 1, st))
    x = torch.randn(B, 64, H, W, generator=g); xd = nhwc(x)
    for dil, relu in ((2, 1), (1, 0)):
        w2 = torch.randn(64, 64, 3, 3, generator=g) * 0.05; b2 = torch.randn(64, generator=g)
        p2, b2d = T._pack(0, w2.cuda(), k=3), b2.cuda()
        _lib.check(lib.mrb_tc_conv_nhwc(_lib.ptr(xd), _lib.ptr(p2), _lib.ptr(b2d), _lib.ptr(out), B, H, W, 64, 3, dil, relu, st))
    h = torch.randn(B, 64, H, W, generator=g)
    wih = torch.randn(192, 64, 1, 1, generator=g) * 0.1; whh = torch.randn(192, 64, 1, 1, generator=g) * 0.1
    bih = torch.randn(192, generator=g)
    ref = onets.conv_gru_cell(x, h, wih, bih, whh, 1, 1)
    hd, pg, bihd = nhwc(h), T._pack(1, wih.cuda(), whh.cuda()), bih.cuda()
    _lib.check(lib.mrb_tc_gru_nhwc(_lib.ptr(xd), _lib.ptr(hd), _lib.ptr(pg), _lib.ptr(bihd), _lib.ptr(out), B, H, W, 64, st))
    o = out.permute(0, 3, 1, 2).cpu()
    e = ((o.double() - ref.double()).norm() / ref.double().norm()).item()
    if e > 2e-6:
        nbad += 1
        d = (o - ref).abs()
        idx = (d > 1e-5).nonzero()
        print("iter", it, "err", e, "n>1e-5:", idx.shape[0], "chans", sorted(set(i[1].item() for i in idx))[:70], "pix", sorted(set(i[2].item() * W + i[3].item() for i in idx))[:40])
print("bad", nbad, "of 150")

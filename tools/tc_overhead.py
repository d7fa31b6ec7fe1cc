import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mridc_b200 as mb
from mridc_b200 import _lib, synth
from mridc_b200.rim_tc import RimTcEngine
lib = _lib.load(); st = _lib.stream_ptr()
dev = torch.device("cuda")
model = mb.CIRIM(synth.cirim_cfg("GRU")).cuda().eval()
blk = model.cirim[0]; eng = RimTcEngine(blk); packs = eng.packs()
c0, c1, r0 = blk.layers[0].convs, blk.layers[1].convs, blk.layers[0].rnn
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, H, W) in ((1, 16, 8), (1, 128, 148), (1, 320, 320)):
    g4 = torch.randn(B, H, W, 4, device=dev); x = torch.randn(B, H, W, 64, device=dev); h = torch.randn(B, H, W, 64, device=dev)
    out = torch.empty(B, H, W, 64, device=dev)
    a = t(lambda: lib.mrb_tc_conv5x5x4_nhwc(_lib.ptr(g4), _lib.ptr(packs[0][0]), _lib.ptr(c0.conv_layer.bias), _lib.ptr(out), B, H, W, 64, 1, st))
    b = t(lambda: lib.mrb_tc_gru_nhwc(_lib.ptr(x), _lib.ptr(h), _lib.ptr(packs[0][1]), _lib.ptr(r0.ih.bias), _lib.ptr(out), B, H, W, 64, st))
    c = t(lambda: lib.mrb_tc_conv_nhwc(_lib.ptr(x), _lib.ptr(packs[1][0]), _lib.ptr(c1.conv_layer.bias), _lib.ptr(out), B, H, W, 64, 3, 2, 1, st))
    print("B=%d %dx%d (%d tiles): conv5x5x4 %.1f us  gru %.1f us  conv3x3d2 %.1f us" % (B, H, W, (B*H*W+127)//128, a, b, c))

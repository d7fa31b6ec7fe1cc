"""Time the final RIM conv (64 -> 2, fused eta update) right after a GRU kernel wrote its input (L2 state as in the loop)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mridc_b200 as mb
from mridc_b200 import _lib, synth
from mridc_b200.rim_tc import RimTcEngine
lib = _lib.load(); st = _lib.stream_ptr()
B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 320, 320
dev = torch.device("cuda")
model = mb.CIRIM(synth.cirim_cfg("GRU")).cuda().eval()
blk = model.cirim[0]; eng = RimTcEngine(blk); packs = eng.packs()
x = torch.randn(B, H, W, 64, device=dev); h = torch.randn(B, H, W, 64, device=dev); out = torch.empty(B, H, W, 64, device=dev)
r0 = blk.layers[0].rnn
w3 = blk.final_layer[0].conv_layer.weight
eta = torch.randn(B, H, W, 2, device=dev); o2 = torch.empty_like(eta)
gru = lambda: lib.mrb_tc_gru_nhwc(_lib.ptr(x), _lib.ptr(h), _lib.ptr(packs[0][1]), _lib.ptr(r0.ih.bias), _lib.ptr(out), B, H, W, 64, st)
c2 = lambda: lib.mrb_conv_c2_nhwc_residual(_lib.ptr(out), _lib.ptr(w3), None, _lib.ptr(eta), _lib.ptr(o2), B, H, W, 64, 3, 1, st)
for _ in range(3): gru(); c2()
torch.cuda.synchronize()
n = 20; tot = 0.0
for _ in range(n):
    gru()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c2(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print("conv_c2 after GRU: %.1f us" % (tot / n * 1e3))

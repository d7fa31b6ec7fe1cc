"""Profiling driver: runs the hot operators a few times (for ncu launch lists / full captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mridc_b200 as mb
from mridc_b200 import _ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
what = sys.argv[3] if len(sys.argv) > 3 else "all"
C, H, W = 15, 320, 320
dev = torch.device("cuda")
torch.manual_seed(0)
y = torch.randn(B, C, H, W, 2, device=dev)
S = torch.randn(B, C, H, W, 2, device=dev)
eta = torch.randn(B, H, W, 2, device=dev)
mask = (torch.rand(1, 1, 1, W, 1, device=dev) < 0.25).to(torch.uint8)
ws = torch.empty((2, B, C, H, W, 2), device=dev)
cfg = synth.cirim_cfg("GRU")
model = mb.CIRIM(cfg).cuda().eval()
blk = model.cirim[0]
from mridc_b200.rim_tc import RimTcEngine
eng = RimTcEngine(blk)
conv_step, _ = eng.bench_step(B, H, W, dev)   # one time step of the regulariser in the engine's own layout
yhyb = _ops.dc_hybrid_prepare(y, mask, False, ws=ws[0])
from mridc_b200 import _lib
g8o = torch.zeros(_lib.load().mrb_g8_bytes(B, H, W), dtype=torch.uint8, device=dev)
for _ in range(reps):
    if what in ("all", "dc"):  # the production form at W = 320: G8 output (split-bf16 conv input with its border)
        _ops.dc_rim_grad(eta, y, S, mask, 1.0, False, "backward", out=g8o, nhwc=2, y_hybrid=yhyb)
    if what in ("all", "conv"):
        conv_step()
    if what in ("all", "vn"):
        pass
torch.cuda.synchronize()
if what in ("all", "vn"):
    vn = mb.VarNet(synth.varnet_cfg(num_cascades=1)).cuda().eval()
    tgt = torch.zeros(B, H, W, dtype=torch.complex64, device=dev)
    for _ in range(reps):
        vn(y, S, mask, None, tgt)
    torch.cuda.synchronize()
print("done")

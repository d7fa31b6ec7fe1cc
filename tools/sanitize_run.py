"""Workload for compute-sanitizer: one CIRIM forward (BH tensor-core engine) with 2 cascades + the DC operators + one
E2EVN cascade at a reduced batch (the tools are 10-100x slower than native)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mridc_b200 as mb
from mridc_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H = int(sys.argv[2]) if len(sys.argv) > 2 else 320
what = sys.argv[3] if len(sys.argv) > 3 else "all"
batch = synth.make_batch(B, 15, H, 320)
dev = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
if what in ("all", "cirim"):
    model = mb.CIRIM(synth.cirim_cfg("GRU", num_cascades=2)).cuda().eval()
    out = next(model(dev["y"], dev["sensitivity_maps"], dev["mask"], None, dev["target"]))
    torch.cuda.synchronize()
    print("cirim ok", float(out[-1][-1].abs().mean()))
if what in ("all", "indrnn"):
    model = mb.CIRIM(synth.cirim_cfg("IndRNN", num_cascades=1)).cuda().eval()
    out = next(model(dev["y"], dev["sensitivity_maps"], dev["mask"], None, dev["target"]))
    torch.cuda.synchronize()
    print("cirim-indrnn ok", float(out[-1][-1].abs().mean()))
if what in ("all", "vn"):
    vn = mb.VarNet(synth.varnet_cfg(num_cascades=1)).cuda().eval()
    o = vn(dev["y"], dev["sensitivity_maps"], dev["mask"], None, dev["target"])
    torch.cuda.synchronize()
    print("varnet ok", float(o.abs().mean()))

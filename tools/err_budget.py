"""Where does the end-to-end error of CIRIM 5x8 come from?  CUDA (tensor-core on/off) vs CPU fp32 oracle vs CPU fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mridc_b200 as mb
from mridc_b200 import synth
from oracle import models as omodels

def rel(a, b):
    a, b = torch.view_as_real(a.cpu()).double(), torch.view_as_real(b.cpu()).double()
    return ((a - b).norm() / b.norm()).item()

for cen, nrm in ((False, "backward"), (True, "ortho")):
    cfg = synth.cirim_cfg("GRU", centered=cen, normalization=nrm)
    batch = synth.make_batch(1, 15, 320, 320, centered=cen, normalization=nrm)
    torch.manual_seed(1)
    model = mb.CIRIM(cfg).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        r32 = omodels.cirim_forward(sd, cfg, batch["y"], batch["sensitivity_maps"], batch["mask"], None, batch["target"])[-1][-1]
        sd64 = {k: v.double() for k, v in sd.items()}
        r64 = omodels.cirim_forward(sd64, cfg, batch["y"].double(), batch["sensitivity_maps"].double(), batch["mask"],
                                    None, batch["target"])[-1][-1]
    model = model.cuda()
    args = (batch["y"].cuda(), batch["sensitivity_maps"].cuda(), batch["mask"].cuda(), None, batch["target"].cuda())
    os.environ["MRIDC_B200_DISABLE_TC"] = "0"
    g_tc = next(model(*args))[-1][-1]
    os.environ["MRIDC_B200_DISABLE_TC"] = "1"
    for blk in model.cirim: blk._tc_engine = None
    g_32 = next(model(*args))[-1][-1]
    os.environ["MRIDC_B200_DISABLE_TC"] = "0"
    for blk in model.cirim: blk._tc_engine = None
    print("centered=%s %s: cpu32-vs-fp64 %.2e | cuda_fp32-vs-fp64 %.2e  cuda_tc-vs-fp64 %.2e | cuda_fp32-vs-cpu32 %.2e  cuda_tc-vs-cpu32 %.2e  cuda_tc-vs-cuda_fp32 %.2e" % (
        cen, nrm, rel(r32, r64), rel(g_32, r64), rel(g_tc, r64), rel(g_32, r32), rel(g_tc, r32), rel(g_tc, g_32)))

"""Secondary measurement (not the driver's bench line): E2EVN (VarNet) 12 cascades, BASELINE.json configs[1].

15-coil 320x320 knee-shaped slices, 4x Gaussian 1-D mask, U-Net 14 channels / 2 pools, random-init weights.
Prints one JSON line: slices/s device-resident and end to end through VarNet.forward (pinned host inputs, H2D + D2H in
the timed region), the soft-DC block alone (sens_reduce + sens_expand_softdc: algorithmic HBM bytes / time, SURVEY 8d
contract figure 6*CHW8 + 2*HW8 per cascade), ZF (configs[0]) and the CPU oracle port for one slice.

    python tools/bench_varnet.py [--batch 8] [--steps 10] [--warmup 3] [--no-cpu-baseline]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import mridc_b200 as mb  # noqa: E402
from mridc_b200 import _ops, synth  # noqa: E402

C, H, W = 15, 320, 320


def timed(fn, n, w):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    B = a.batch
    np.random.seed(123)  # the Gaussian mask draws from numpy's global generator (as upstream)
    d = synth.make_batch(B, C, H, W, mask_func=synth.Gaussian1DMask([0.7], [4]), seed=123, mask_dtype="uint8")
    cfg = synth.varnet_cfg()
    torch.manual_seed(1)
    model = mb.VarNet(cfg).eval()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda()
    host = {k: d[k].pin_memory() for k in ("y", "sensitivity_maps", "mask", "target")}
    dev = {k: v.cuda() for k, v in host.items()}

    def step_dev():
        return model(dev["y"], dev["sensitivity_maps"], dev["mask"], None, dev["target"])

    def step_e2e():
        t = {k: v.cuda(non_blocking=True) for k, v in host.items()}
        return model(t["y"], t["sensitivity_maps"], t["mask"], None, t["target"]).cpu()

    ms = timed(step_dev, a.steps, a.warmup)
    ms_e2e = timed(step_e2e, a.steps, a.warmup)
    # the data-consistency half of one cascade: reduce (IFFT + conj-coil sum) and expand + soft DC
    ws = torch.empty((2, B, C, H, W, 2), device="cuda")
    img = torch.randn(B, H, W, 2, device="cuda")
    dcw = torch.ones(1, device="cuda")
    out = torch.empty_like(dev["y"])

    def dc_block():
        _ops.sens_reduce(dev["y"], dev["sensitivity_maps"], False, "backward", ws=ws)
        _ops.sens_expand_softdc(img, dev["sensitivity_maps"], dev["y"], dev["y"], dev["y"], dev["mask"], dcw, False, False,
                                "backward", out=out, ws=ws)

    ms_dc = timed(dc_block, 20, 3)
    chw8, hw8 = C * H * W * 8, H * W * 8
    dc_bytes = B * (6 * chw8 + 2 * hw8)
    zf = mb.ZF(synth.zf_cfg())
    ms_zf = timed(lambda: zf(dev["y"], dev["sensitivity_maps"], dev["mask"], dev["target"]), 20, 3)
    res = {
        "metric": "e2evn_320x320x15coil_slices_per_sec", "unit": "slices/s", "value": B / ms * 1e3, "ms_per_step": ms,
        "steps": a.steps, "warmup": a.warmup, "n_gpus": 1, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "E2EVN (VarNet) 12 cascades, U-Net 14 channels / 2 pools, 15-coil 320x320 knee-shaped slices, "
                               "4x Gaussian 1-D mask (BASELINE.json configs[1])", "slices_per_step": B},
        "e2e": {"value": B / ms_e2e * 1e3, "unit": "slices/s",
                "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values())),
                "d2h_bytes_per_step": B * H * W * 8},
        "soft_dc_block": {"ms": ms_dc, "algorithmic_bytes": dc_bytes, "achieved_gbs": dc_bytes / ms_dc / 1e6,
                          "share_of_step": 12 * ms_dc / ms,
                          "note": "sens_reduce + sens_expand_softdc of one cascade; contract figure 6*CHW8 + 2*HW8 per slice"},
        "zf_configs0": {"ms": ms_zf, "slices_per_sec": B / ms_zf * 1e3,
                        "achieved_gbs": B * (2 * chw8 + hw8) / ms_zf / 1e6},
    }
    if not a.no_cpu_baseline:
        from oracle import models as omodels

        torch.set_num_threads(os.cpu_count() or 1)
        np.random.seed(123)
        one = synth.make_batch(1, C, H, W, mask_func=synth.Gaussian1DMask([0.7], [4]), seed=123, mask_dtype="uint8")
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = omodels.varnet_forward(sd, cfg, one["y"], one["sensitivity_maps"], one["mask"], None, one["target"])
        dt = time.perf_counter() - t0
        o1 = model(one["y"].cuda(), one["sensitivity_maps"].cuda(), one["mask"].cuda(), None, one["target"].cuda())
        a_, b_ = torch.view_as_real(o1.cpu()).double(), torch.view_as_real(ref).double()
        t0 = time.perf_counter()
        with torch.no_grad():
            omodels.zf_forward(synth.zf_cfg(), one["y"], one["sensitivity_maps"], one["mask"], one["target"])
        dtz = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 1.0 / dt, "unit": "slices/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": "1 slice of the same workload, torch-CPU oracle port, no warm-up, %.1f s" % dt,
                               "parity_rel_l2_vs_cuda": float((a_ - b_).norm() / b_.norm()),
                               "zf_slices_per_sec": 1.0 / dtz}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

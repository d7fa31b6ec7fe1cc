"""Attribute uconv3_kernel time to its roles by switching them off in the tools build (results are garbage, timing only):
1 no global loads, 2 no MMAs, 4 no output stores, 8 no split / staging stores."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _lib
import _toolslib
lib = _toolslib.load(); st = _lib.stream_ptr()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for cin, cout, hw in ((14, 14, 320), (28, 28, 160), (56, 28, 160)):
    x = torch.randn(B, cin, hw, hw, device="cuda"); w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    o = torch.empty(B, cout, hw, hw, device="cuda")
    pk = torch.empty(_lib.load().mrb_tc2_unet_packed_bytes(cin, cout), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mrb_tc2_unet_pack(_lib.ptr(w), _lib.ptr(pk), cin, cout, st))
    run = lambda: lib.mrb_tc2_unet_conv3x3(_lib.ptr(x), cin * hw * hw, _lib.ptr(pk), _lib.ptr(o), cout * hw * hw, B, cin, cout, hw, hw, st)
    res = []
    for flags in (0, 1, 2, 4, 8, 9, 3, 6, 15):
        lib.mrb_tc2_set_debug(flags)
        res.append("%d:%.0f" % (flags, t(run)))
    lib.mrb_tc2_set_debug(0)
    print("%2d -> %2d @ %3d  us by switched-off roles  %s" % (cin, cout, hw, "  ".join(res)))
# per-role cycle counters (mean over CTAs): loaders (two groups), MMA lane, epilogue
prof = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
lib.mrb_tc2_set_prof(_lib.ptr(prof))
for cin, cout, hw in ((14, 14, 320), (28, 28, 160), (56, 28, 160)):
    x = torch.randn(B, cin, hw, hw, device="cuda"); w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    o = torch.empty(B, cout, hw, hw, device="cuda")
    pk = torch.empty(_lib.load().mrb_tc2_unet_packed_bytes(cin, cout), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mrb_tc2_unet_pack(_lib.ptr(w), _lib.ptr(pk), cin, cout, st))
    prof.zero_()
    lib.mrb_tc2_unet_conv3x3(_lib.ptr(x), cin * hw * hw, _lib.ptr(pk), _lib.ptr(o), cout * hw * hw, B, cin, cout, hw, hw, st)
    torch.cuda.synchronize()
    p = prof.view(148, 16).double().mean(0).tolist()
    print("%2d -> %2d @ %3d cycles: prologue %6.0f | loader0 total %7.0f wait_empty %7.0f load+stage %7.0f | loader1 %7.0f %7.0f %7.0f | "
          "mma total %7.0f wait_acc %7.0f wait_full %7.0f | epi total %7.0f wait %7.0f" % (
              cin, cout, hw, p[15], p[0], p[1], p[2], p[4], p[5], p[6], p[8], p[9], p[10], p[12], p[13]))
lib.mrb_tc2_set_prof(None)

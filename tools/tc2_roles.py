"""Role attribution for the second-generation GRU kernel (tools build: cycle counters + role switches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from mridc_b200 import _lib
import _toolslib
lib = _toolslib.load(); st = _lib.stream_ptr()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = W = 320
dev = "cuda"
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
nb = lib.mrb_bh_bytes(B, H, W)
x = torch.randn(B, H, W, 64, device=dev); h = torch.randn(B, H, W, 64, device=dev)
xb = torch.empty(nb, dtype=torch.uint8, device=dev); hb = torch.empty_like(xb); ob = torch.empty_like(xb)
lib.mrb_bh_from_nhwc(_lib.ptr(x), _lib.ptr(xb), B, H, W, st); lib.mrb_bh_from_nhwc(_lib.ptr(h), _lib.ptr(hb), B, H, W, st)
wih = torch.randn(192, 64, device=dev) * 0.1; whh = torch.randn(192, 64, device=dev) * 0.1; b = torch.randn(192, device=dev)
pk = torch.empty(lib.mrb_tc2_gru_packed_bytes(), dtype=torch.uint8, device=dev)
assert 0 == lib.mrb_tc2_pack_gru(_lib.ptr(wih), _lib.ptr(whh), _lib.ptr(pk), 64, 64, st)
run = lambda: lib.mrb_tc2_gru(_lib.ptr(xb), _lib.ptr(hb), _lib.ptr(pk), _lib.ptr(b), _lib.ptr(ob), B, H, W, st)
names = {0: "full", 1: "no MMA", 2: "no TMA loads", 4: "no gate math", 8: "no TMA stores", 4 | 8: "no math, no stores",
         1 | 2: "no MMA, no TMA", 1 | 2 | 4 | 8: "skeleton"}
for f, nm in names.items():
    lib.mrb_tc2_set_debug(f)
    print("%-28s %7.1f us" % (nm, t(run)), flush=True)
lib.mrb_tc2_set_debug(0)
prof = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.mrb_tc2_set_prof(_lib.ptr(prof))
run(); torch.cuda.synchronize()
p = prof.view(148, 16).double().mean(0).tolist()
print("producer: total %.0f wait_empty %.0f | mma: total %.0f wait_acc %.0f wait_full %.0f | epi(warp 0): total %.0f wait_acc %.0f tmem_ld %.0f math %.0f split+store %.0f (cycles, mean over CTAs)" % tuple(p[:10]))
lib.mrb_tc2_set_prof(None)

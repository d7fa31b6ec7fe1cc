"""Per-layer timing of the U-Net 3x3 convolutions of E2EVN (BASELINE.json configs[1]: 14 channels, 2 pools) at B slices:
tensor-core fp16-split kernel vs the exact-fp32 CUDA-core kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _lib, _ops
lib = _lib.load(); st = _lib.stream_ptr()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
tot_tc = tot_ff = 0.0
for cin, cout, hw in ((14, 14, 320), (14, 28, 160), (28, 28, 160), (28, 56, 80), (56, 56, 80), (56, 28, 160), (28, 28, 160),
                      (28, 14, 320), (14, 14, 320)):
    x = torch.randn(B, cin, hw, hw, device="cuda"); w = torch.randn(cout, cin, 3, 3, device="cuda") * 0.05
    o = torch.empty(B, cout, hw, hw, device="cuda")
    pk = torch.empty(lib.mrb_tc2_unet_packed_bytes(cin, cout), dtype=torch.uint8, device="cuda")
    _lib.check(lib.mrb_tc2_unet_pack(_lib.ptr(w), _lib.ptr(pk), cin, cout, st))
    tc = t(lambda: _lib.check(lib.mrb_tc2_unet_conv3x3(_lib.ptr(x), cin * hw * hw, _lib.ptr(pk), _lib.ptr(o), cout * hw * hw, B, cin,
                                                       cout, hw, hw, st)))
    ff = t(lambda: _ops.conv2d(x, w, None, 3, 1, _ops.PAD_ZERO, out=o))
    fl = 2.0 * B * hw * hw * cin * cout * 9
    tot_tc += tc; tot_ff += ff
    print("%2d -> %2d @ %3d: tc %7.1f us (%6.1f TFLOP/s, %5.0f GB/s of in+out)   fp32 %7.1f us (%5.1f TFLOP/s)" % (
        cin, cout, hw, tc, fl / tc / 1e6, B * hw * hw * (cin + cout) * 4 / tc / 1e3, ff, fl / ff / 1e6))
print("sum of the 9 convs at B=%d: tc %.0f us, fp32 %.0f us" % (B, tot_tc, tot_ff))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mridc_b200 import _lib
import _toolslib
lib = _toolslib.load(); st = _lib.stream_ptr()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
iters = 960
print("tcgen05.mma kind::tf32 M=128 K=8: cycles per MMA (issue / to-completion), %d MMAs" % iters)
for a_tmem in (2, 6):
    for N in (64, 96, 128, 192):
        row = []
        for nacc in (1,):
            if nacc * N > 448:
                continue
            assert 0 == (lib.mrb_tc_microbench(N, nacc, iters, a_tmem, _lib.ptr(out), st))
            torch.cuda.synchronize()
            o = out.tolist()
            row.append("nacc=%d: %5.1f / %5.1f" % (nacc, o[0] / iters, o[1] / iters))
        print("A in %s N=%3d  " % ({6: "TMEM(tight, ONE accumulator)", 5: "TMEM(tight, TWO issuing threads; per-thread cycles per MMA)", 4: "TMEM(tight+commit+fence/12)", 3: "TMEM(tight+commit/12)", 2: "TMEM(tight)", 1: "TMEM", 0: "smem"}[a_tmem], N) + "   ".join(row))

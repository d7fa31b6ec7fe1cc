"""Uninitialised-memory detector: poison the caching allocator's free blocks with NaNs, then run the module-level conv /
cell parity cases (and a few fused operators) and report every case whose result changes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import rel_l2
from mridc_b200.rim import ConvGRUCell, ConvMGUCell, ConvNonlinear, IndRNNCell
from oracle import nets as onets


def poison():
    junk = [torch.full((32, 1024, 1024), float("nan"), device="cuda") for _ in range(6)]
    junk += [torch.full((n,), float("nan"), device="cuda") for n in (1 << 10, 1 << 14, 1 << 18, 1 << 20, 1 << 22) for _ in range(8)]
    torch.cuda.synchronize()
    del junk


bad = 0
for rep in range(3):
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(4)
    for cin, cout, k, dil, nl, H, W in ((4, 64, 5, 1, "relu", 37, 45), (64, 64, 3, 2, "relu", 20, 70),
                                        (64, 2, 3, 1, None, 33, 33), (3, 7, 7, 1, "leakyrelu", 16, 19),
                                        (5, 20, 3, 3, "relu", 24, 31), (9, 40, 1, 1, None, 10, 12)):
        mod = ConvNonlinear(cin, cout, 2, k, dil, True, nl)
        with torch.no_grad():
            mod.conv_layer.bias.normal_()
        x = torch.randn(2, cin, H, W, generator=g)
        ref = onets.conv_nonlinear(x, mod.conv_layer.weight.detach(), mod.conv_layer.bias.detach(), k, dil, nl)
        poison()
        e = rel_l2(mod.cuda()(x.cuda()), ref)
        if not e < 2e-6:
            bad += 1
            print("FAIL conv", (cin, cout, k, dil), e, flush=True)
    for cls, fn, key in ((ConvGRUCell, onets.conv_gru_cell, None), (ConvMGUCell, onets.conv_mgu_cell, None),
                         (IndRNNCell, onets.indrnn_cell, "hh")):
        for cx, ch, k, dil in ((64, 64, 1, 1), (16, 24, 3, 1), (8, 70, 1, 1), (6, 10, 3, 2)):
            mod = cls(cx, ch, 2, k, dil, True)
            with torch.no_grad():
                mod.ih.bias.normal_()
            x = torch.randn(2, cx, 19, 23, generator=g)
            h = torch.randn(2, ch, 19, 23, generator=g)
            hh = mod.hh.detach() if key else mod.hh.weight.detach()
            ref = fn(x, h, mod.ih.weight.detach(), mod.ih.bias.detach(), hh, k, dil)
            poison()
            e = rel_l2(mod.cuda()(x.cuda(), h.cuda()), ref)
            if not e < 2e-6:
                bad += 1
                print("FAIL cell", cls.__name__, (cx, ch, k, dil), e, flush=True)
print("poisoned-allocator run: %d failing case(s)" % bad)

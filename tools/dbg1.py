import faulthandler, sys, torch
faulthandler.enable()
sys.path.insert(0, '.')
import mridc_b200 as mb
from oracle import mri as omri
g = torch.Generator().manual_seed(5)
x = torch.randn(4, 6, 10, 2, generator=g)
xc = torch.view_as_complex(x)
print('A', flush=True)
a = mb.fft2(xc.cuda(), True, "ortho"); torch.cuda.synchronize(); print('B', flush=True)
xt = x.permute(1, 0, 2, 3)
xg = xt.cuda(); torch.cuda.synchronize(); print('C', xg.shape, xg.stride(), flush=True)
xcg = xg.contiguous(); torch.cuda.synchronize(); print('D', flush=True)
v = torch.view_as_complex(xcg); print('E', flush=True)
b = mb.ifft2(xg, False, "forward"); torch.cuda.synchronize(); print('F', flush=True)

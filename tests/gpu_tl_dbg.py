import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aadff_b200
from oracle import focal_stack_oracle as orc
from deeplens.psfnet import ThinLens
gen = torch.Generator().manual_seed(5)
for (N, C, H, W, ks) in [(1, 3, 40, 56, 11), (2, 3, 72, 200, 11), (1, 1, 96, 160, 31)]:
    tlk = ThinLens(foc_len=50.0, fnum=1.8, kernel_size=ks, sensor_size=[36.0, 24.0], sensor_res=(H, W)).to("cuda")
    im = torch.rand(N, C, H, W, generator=gen); dp = 300 + 6000 * torch.rand(N, 1, H, W, generator=gen); fc = 500 + 3000 * torch.rand(N, generator=gen)
    ref = orc.thinlens_render(im, dp, fc, ks, 50.0, 1.8, tlk.ps)
    for flags in (16, 0):
        aadff_b200.native.lib.aadff_debug_set_flags(flags)
        got = tlk.render(im.cuda(), dp.cuda(), fc.cuda())
        torch.cuda.synchronize()
        print(N, C, H, W, ks, "flags", flags, "err", float((got.cpu() - ref).abs().max()), flush=True)

"""Workload for the ncu capture of the final secondary kernels (profiles/r02b_ncu_rows_final.md): thin lens, strip
gather k = 11 / 7, fp32 mode, tensor-core pred.   ncu --set full -k regex:... --launch-skip 5 -c 5 python tests/gpu_ncu_rows2.py"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import aadff_b200
from aadff_b200 import synthetic
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
N, H, W, ks = 16, 512, 512, 11
img = torch.rand(N, 3, H, W, device="cuda")
dep = -(300 + 5000 * torch.rand(N, 1, H, W, device="cuda"))
foc = -(500 + 3000 * torch.rand(N, device="cuda"))
tl = aadff_b200.ThinLens(foc_len=50.0, fnum=1.8, kernel_size=ks, sensor_size=[36.0, 24.0], sensor_res=(H, W)).to("cuda")
psf = torch.rand(N, H, W, ks, ks, device="cuda")
psf7 = torch.rand(N, H, W, 7, 7, device="cuda")
lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
i1, dm = synthetic.synthetic_rgbd(1, 256, 256, seed=7)
f1 = -synthetic.synthetic_focus(dm, 5).cuda() * 1e3
probes = torch.rand(1 << 18, 4, device="cuda")
for _ in range(2):
    tl.render(img, dep, foc)
    aadff_b200.local_psf_render(img, psf, ks)
    aadff_b200.local_psf_render(img, psf7, 7)
    lens.render_stack(i1.cuda(), -dm.cuda() * 1e3, f1, mode="fp32")
    lens.pred(probes, mode="parity")
torch.cuda.synchronize()

"""Workload for an ncu capture of the secondary kernels (thin lens, stand-alone gather, PSF convolution, dataset
preparation): one launch each at a representative size.   ncu --set full -k regex:... python tests/gpu_ncu_rows.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from deeplens.psfnet import ThinLens  # noqa: E402
from deeplens.render_psf import local_psf_render, render_psf  # noqa: E402

N, H, W, ks = 16, 512, 512, 11
img = torch.rand(N, 3, H, W, device="cuda")
dep = -(300 + 5000 * torch.rand(N, 1, H, W, device="cuda"))
foc = -(500 + 3000 * torch.rand(N, device="cuda"))
tl = ThinLens(foc_len=50.0, fnum=1.8, kernel_size=ks, sensor_size=[36.0, 24.0], sensor_res=(H, W)).to("cuda")
psf = torch.rand(4, H, W, ks, ks, device="cuda")
bgr = torch.randint(0, 255, (8, 1024, 1280, 3), dtype=torch.uint8, device="cuda")
d16 = torch.randint(0, 8000, (8, 1024, 1280), dtype=torch.int16, device="cuda").view(torch.uint16)
for _ in range(3):
    tl.render(img, dep, foc)
    local_psf_render(img[:4], psf, ks)
    render_psf(img[:4], torch.rand(3, ks, ks, device="cuda"))
    aadff_b200.preprocess_rgbd(bgr, d16, (480, 640))
torch.cuda.synchronize()

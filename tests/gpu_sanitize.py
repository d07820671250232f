"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck), not a pytest file:

    compute-sanitizer --tool memcheck  python tests/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tests/gpu_sanitize.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from oracle import focal_stack_oracle as orc  # noqa: E402

lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
img, dm = orc.synthetic_rgbd(1, 24, 40, seed=1)
foc = -orc.synthetic_focus(dm, 2).cuda() * 1e3
img, dep = img.cuda(), -dm.cuda() * 1e3
for mode in ("parity", "econ", "mixed", "fast", "fp32"):
    out = lens.render_stack(img, dep, foc, mode=mode)
    torch.cuda.synchronize()
    print(mode, float(out.mean()))
print("pred", float(lens.pred(torch.rand(200, 4).cuda(), mode="parity").sum()))
psf = torch.rand(1, 24, 40, 11, 11, device="cuda")
print("gather", float(aadff_b200.local_psf_render(img, psf, 11).mean()))
img5 = torch.rand(2, 5, 19, 45, device="cuda")                   # ragged: partial groups / tiles, 3 + 1 + 1 channel passes
for ks in (3, 7, 31):
    psf = torch.rand(2, 19, 45, ks, ks, device="cuda")
    print("gather ragged ks", ks, float(aadff_b200.local_psf_render(img5, psf, ks).mean()))
tl = aadff_b200.ThinLens(50.0, 1.8, 11, [36.0, 24.0], (24, 40)).to("cuda")
print("thinlens", float(tl.render(img, dep, foc[:, 0]).mean()))

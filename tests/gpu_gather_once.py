"""One launch of local_psf_render for profilers (ncu): python tests/gpu_gather_once.py [N H W ks]."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import aadff_b200  # noqa: E402

N, H, W, ks = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (16, 512, 512, 11)
img = torch.rand(N, 3, H, W, device="cuda")
psf = torch.rand(N, H, W, ks, ks, device="cuda")
for _ in range(3):
    out = aadff_b200.local_psf_render(img, psf, ks)
torch.cuda.synchronize()
print(float(out.mean()))

"""Timing of the CUDA-core fp32 mode (mlp_fp32_kernel) on c2 and of pred() through it."""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from aadff_b200 import synthetic  # noqa: E402

lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
img, dm = synthetic.synthetic_rgbd(1, 512, 512, seed=7)
foc = -synthetic.synthetic_focus(dm, 5).cuda() * 1e3
img, dep = img.cuda(), -dm.cuda() * 1e3


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms = timeit(lambda: lens.render_stack(img, dep, foc, mode="fp32"), 3)
ref = lens.render_stack(img, dep, foc, mode="parity")
out = lens.render_stack(img, dep, foc, mode="fp32")
flops = 5 * 512 * 512 * 1143808
print(f"fp32 mode c2: {ms:.2f} ms  {5 * 512 * 512 / ms / 1e3:.1f} Mpix*slices/s  {flops / ms / 1e9:.1f} TFLOP/s fp32 "
      f"({flops / ms / 1e9 / 74.5:.2f} of 148 SM x 128 FFMA x 2 x 1.965 GHz = 74.5)  max|fp32 - parity| {float((out - ref).abs().max()):.2e}")
inp = torch.rand(1 << 18, 4, device="cuda")
ms = timeit(lambda: lens.pred(inp), 3)
print(f"pred (fp32 kernel) M=256Ki: {ms:.2f} ms  {inp.shape[0] / ms / 1e3:.1f} Mprobes/s")

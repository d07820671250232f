"""Time the fused kernel of whatever library AADFF_LIB_PATH points to (build-variant experiments):
burst (c2) and sustained (c5b16 back to back), parity / econ / fast.   AADFF_LIB_PATH=... python tests/gpu_variant_time.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from aadff_b200 import synthetic  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
tag = os.path.basename(os.environ.get("AADFF_LIB_PATH") or "default")
for (N, S, H, W, iters, name) in [(1, 5, 512, 512, 30, "c2"), (16, 5, 512, 512, 12, "c5b16")]:
    img, dm = synthetic.synthetic_rgbd(N, H, W, seed=7)
    foc = -synthetic.synthetic_focus(dm, S).cuda() * 1e3
    img, dep = img.cuda(), -dm.cuda() * 1e3
    for mode in ("parity", "econ", "fast"):
        ms = min(timeit(lambda: lens.render_stack(img, dep, foc, mode=mode), iters) for _ in range(2))
        print(f"[variant {tag}] {name:6s} {mode:7s} {ms:8.3f} ms  {N * S * H * W / ms / 1e3:7.1f} Mpix*s/s  checksum {float(lens.render_stack(img, dep, foc, mode=mode).double().sum()):.6f}", flush=True)

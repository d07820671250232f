"""A/B timing of ThinLens.render: flags 0 = two pixels per thread (default), 2048 = one pixel per thread."""
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from deeplens.psfnet import ThinLens  # noqa: E402

lib = aadff_b200.native.lib


def timeit(fn, iters=10):
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


for (N, H, W, ks) in [(4, 512, 512, 11), (16, 512, 512, 11), (16, 512, 512, 7), (16, 512, 512, 3), (16, 512, 512, 15),
                      (4, 512, 512, 21), (1, 1080, 1920, 31), (4, 1080, 1920, 31), (16, 480, 640, 11)]:
    tl = ThinLens(foc_len=50.0, fnum=1.8, kernel_size=ks, sensor_size=[36.0, 24.0], sensor_res=(H, W)).to("cuda")
    img = torch.rand(N, 3, H, W, device="cuda")
    dep = -(300 + 5000 * torch.rand(N, 1, H, W, device="cuda"))
    foc = -(500 + 3000 * torch.rand(N, device="cuda"))
    res = {}
    for flags in (2048, 0):
        lib.aadff_debug_set_flags(flags)
        ms = timeit(lambda: tl.render(img, dep, foc))
        res[flags] = (ms, tl.render(img, dep, foc).clone())
        lib.aadff_debug_set_flags(0)
    px = N * H * W
    d = float((res[0][1] - res[2048][1]).abs().max())
    print(f"thinlens N{N} {H}x{W} k{ks}: one-pixel {res[2048][0]:.3f} ms {px / res[2048][0] / 1e6:.2f} Gpix/s | two-pixel "
          f"{res[0][0]:.3f} ms {px / res[0][0] / 1e6:.2f} Gpix/s ({res[2048][0] / res[0][0]:.2f}x, {px * ks * ks / res[0][0] / 1e9:.2f} T taps/s) "
          f"max|d| {d:.1e}", flush=True)

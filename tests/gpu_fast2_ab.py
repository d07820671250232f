"""A/B of AADFF_MODE_FAST: two tiles in flight per CTA (debug flag 32) against the default one-tile kernel:
results (both evaluate the same single-term arithmetic) and device time per launch.    python tests/gpu_fast2_ab.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from aadff_b200 import synthetic  # noqa: E402

nat = aadff_b200.native


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


def main():
    lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
    lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
    for (N, S, H, W, iters, tag) in [(1, 1, 8, 16, 3, "one tile"), (1, 3, 37, 50, 3, "ragged"), (2, 2, 24, 40, 3, "small"),
                                     (1, 5, 512, 512, 20, "c2"), (16, 5, 256, 256, 20, "c3"), (16, 5, 512, 512, 10, "c5b16")]:
        img, dm = synthetic.synthetic_rgbd(N, H, W, seed=7)
        foc = -synthetic.synthetic_focus(dm, S).cuda() * 1e3
        img, dep = img.cuda(), -dm.cuda() * 1e3
        nat.lib.aadff_debug_set_flags(0)
        ref = lens.render_stack(img, dep, foc, mode="fast")
        ms1 = timeit(lambda: lens.render_stack(img, dep, foc, mode="fast"), iters)
        nat.lib.aadff_debug_set_flags(32)
        out = lens.render_stack(img, dep, foc, mode="fast")
        ms2 = timeit(lambda: lens.render_stack(img, dep, foc, mode="fast"), iters)
        nat.lib.aadff_debug_set_flags(0)
        par = lens.render_stack(img, dep, foc, mode="parity")
        px = N * S * H * W / 1e3
        print(f"[fast2 A/B] {tag:9s} max|two-tile - one-tile| = {float((out - ref).abs().max()):.2e}  max|fast - parity| = "
              f"{float((out - par).abs().max()):.2e}  one-tile {ms1:7.3f} ms ({px / ms1:7.1f} Mpix*s/s)  two-tile {ms2:7.3f} ms "
              f"({px / ms2:7.1f})  {ms1 / ms2:.3f}x", flush=True)


if __name__ == "__main__":
    main()

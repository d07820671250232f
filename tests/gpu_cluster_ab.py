"""A/B of the 2-CTA cluster variant (multicast weight stream, debug flag 8) against the default single-CTA kernel:
bit-equality of the results and device time per launch, burst (c2) and sustained (back-to-back c5b16 stacks).
    python tests/gpu_cluster_ab.py [mode ...]"""
import os
import sys
import time

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
    modes = sys.argv[1:] or ["parity", "econ", "fast", "mixed"]
    lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
    lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
    for (N, S, H, W, iters, tag) in [(1, 5, 512, 512, 20, "c2"), (1, 3, 37, 50, 5, "ragged"), (1, 1, 8, 16, 5, "one tile"),
                                     (16, 5, 256, 256, 20, "c3"), (16, 5, 512, 512, 12, "c5b16 (sustained)")]:
        img, dm = synthetic.synthetic_rgbd(N, H, W, seed=7)
        foc = -synthetic.synthetic_focus(dm, S).cuda() * 1e3
        img, dep = img.cuda(), -dm.cuda() * 1e3
        for mode in modes:
            nat.lib.aadff_debug_set_flags(0)  # single-CTA
            ref = lens.render_stack(img, dep, foc, mode=mode)
            ms1 = timeit(lambda: lens.render_stack(img, dep, foc, mode=mode), iters)
            nat.lib.aadff_debug_set_flags(8)
            out = lens.render_stack(img, dep, foc, mode=mode)
            ms2 = timeit(lambda: lens.render_stack(img, dep, foc, mode=mode), iters)
            nat.lib.aadff_debug_set_flags(0)
            px = N * S * H * W / 1e3
            print(f"[cluster A/B] {tag:18s} {mode:7s} equal={bool(torch.equal(out, ref))} single-CTA {ms1:8.3f} ms ({px / ms1:7.1f} Mpix*s/s) "
                  f"cluster-2 {ms2:8.3f} ms ({px / ms2:7.1f})  {ms1 / ms2:.3f}x", flush=True)
    lens31 = aadff_b200.PSFNet(kernel_size=31, device="cuda")
    img, dm = synthetic.synthetic_rgbd(1, 1080, 1920, seed=7)
    foc = -synthetic.synthetic_focus(dm, 4).cuda() * 1e3
    img, dep = img.cuda(), -dm.cuda() * 1e3
    nat.lib.aadff_debug_set_flags(0)  # single-CTA
    ref = lens31.render_stack(img, dep, foc)
    ms1 = timeit(lambda: lens31.render_stack(img, dep, foc), 6)
    nat.lib.aadff_debug_set_flags(8)
    out = lens31.render_stack(img, dep, foc)
    ms2 = timeit(lambda: lens31.render_stack(img, dep, foc), 6)
    nat.lib.aadff_debug_set_flags(0)
    print(f"[cluster A/B] c4-like 4x1080p k31 parity equal={bool(torch.equal(out, ref))} single-CTA {ms1:.3f} ms cluster-2 {ms2:.3f} ms {ms1 / ms2:.3f}x", flush=True)


if __name__ == "__main__":
    main()

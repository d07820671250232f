"""Generic A/B of a debug flag on the fused kernel: python tests/gpu_flag_ab.py <flag> [modes...]"""
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
    flag = int(sys.argv[1])
    modes = sys.argv[2:] or ["parity", "econ", "fast"]
    lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
    lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
    for (N, S, H, W, iters, tag) in [(1, 5, 512, 512, 30, "c2"), (16, 5, 256, 256, 20, "c3"), (16, 5, 512, 512, 8, "c5b16")]:
        img, dm = synthetic.synthetic_rgbd(N, H, W, seed=7)
        foc = -synthetic.synthetic_focus(dm, S).cuda() * 1e3
        img, dep = img.cuda(), -dm.cuda() * 1e3
        for mode in modes:
            res = []
            for rep in range(2):
                for f in (0, flag):
                    nat.lib.aadff_debug_set_flags(f)
                    out = lens.render_stack(img, dep, foc, mode=mode)
                    res.append((f, timeit(lambda: lens.render_stack(img, dep, foc, mode=mode), iters), out))
            nat.lib.aadff_debug_set_flags(0)
            base = min(t for f, t, _ in res if f == 0)
            alt = min(t for f, t, _ in res if f == flag)
            eq = torch.equal(res[0][2], res[1][2])
            px = N * S * H * W / 1e3
            print(f"[flag {flag} A/B] {tag:6s} {mode:7s} equal={eq} base {base:8.3f} ms ({px / base:7.1f}) flag {alt:8.3f} ms ({px / alt:7.1f}) {base / alt:.3f}x", flush=True)


if __name__ == "__main__":
    main()

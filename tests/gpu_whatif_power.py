"""What-if for the power-capped regime (not a pytest file): 64 x 5 x 512 x 512 stacks back to back, with and without the
weight stream (debug flag 1: results invalid).  Measures what the L2 -> shared-memory weight traffic costs under the power cap.

    python tests/gpu_whatif_power.py
"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aadff_b200
from aadff_b200 import synthetic
lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
lens.load_net(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'rf50mm_PSFNet480x640_ks11.pkl'))
N, S, H, W = 64, 5, 512, 512
img, dm = synthetic.synthetic_rgbd(N, H, W, seed=1)
foc = -synthetic.synthetic_focus(dm, S).cuda() * 1e3
img, dep = img.cuda(), -dm.cuda() * 1e3
for flags in (0, 1, 0, 1):
    aadff_b200.native.lib.aadff_debug_set_flags(flags)
    for _ in range(3): lens.render_stack(img, dep, foc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(12): lens.render_stack(img, dep, foc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 12
    print(f"flags={flags} ({'no weight copies' if flags else 'normal'}): {ms:.1f} ms per 64x5x512x512 stack, {N*S*H*W/ms/1e3:.1f} Mpix*slices/s", flush=True)
aadff_b200.native.lib.aadff_debug_set_flags(0)

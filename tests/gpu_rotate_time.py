"""Timing of preprocess_rgbd with and without AutoAgument's spline rotation against scipy.ndimage.rotate on the host
(profiles/r02b_rotate_time.txt)."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import aadff_b200
B, H, W = 8, 1024, 1280
bgr = torch.randint(0, 255, (B, H, W, 3), dtype=torch.uint8, device="cuda")
d16 = torch.randint(0, 8000, (B, H, W), dtype=torch.int16, device="cuda").view(torch.uint16)
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
jit = torch.rand(B, 2); fl = torch.randint(0, 4, (B,), dtype=torch.uint8)
ms00 = timeit(lambda: aadff_b200.preprocess_rgbd(bgr, d16, (480, 640)))
print(f"plain (no jitter, no flips): {ms00:.3f} ms")
ms0 = timeit(lambda: aadff_b200.preprocess_rgbd(bgr, d16, (480, 640), jitter=jit, flips=fl))
ms1 = timeit(lambda: aadff_b200.preprocess_rgbd(bgr, d16, (480, 640), jitter=jit, flips=fl, rotate_deg=[17.0 + 20 * i for i in range(B)]))
ms2 = timeit(lambda: aadff_b200.preprocess_rgbd(bgr, d16, (480, 640), jitter=jit, flips=fl, rotate_deg=[17.0, None] * (B // 2)))
print(f"preprocess_rgbd 8 x 1024x1280 -> 480x640: no rotation {ms0:.3f} ms; all 8 rotated {ms1:.3f} ms ({ms1 / B * 1e3:.0f} us per sample); 4 of 8 rotated {ms2:.3f} ms")
from scipy import ndimage
a = (bgr[0].cpu().numpy()[..., ::-1] / 255.)
t0 = time.perf_counter(); r = ndimage.rotate(a, 17, reshape=False); rd = ndimage.rotate(d16[0].cpu().view(torch.int16).numpy() / 4000, 17, reshape=False); t1 = time.perf_counter()
print(f"scipy.ndimage.rotate of one sample (image + depth) on this box's CPU: {(t1 - t0) * 1e3:.0f} ms")

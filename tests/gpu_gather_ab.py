"""A/B timing of the stand-alone gather kernels: python tests/gpu_gather_ab.py [N H W]
flags 0 = strip-walking kernel (ks <= 15), 8192 / 1024 / 4096 = its shared-memory plans 0 / 1 / 2 forced (0 = automatic choice), 512 = register-streaming kernel.
Prints algorithmic GB/s (4 ks^2 + 24 B per pixel at C = 3) and the fraction of the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402

N, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 512, 512)
peak = 6531.6
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
lib = aadff_b200.native.lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for ks in (3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 31):
    n = N if ks <= 13 else max(1, N // 4)
    if len(sys.argv) > 4 and str(ks) not in sys.argv[4].split(','):
        continue
    img = torch.rand(n, 3, H, W, device="cuda")
    psf = torch.rand(n, H, W, ks, ks, device="cuda")
    ref = None
    for flags in ((512, 0, 8192, 1024, 4096) if ks <= 15 else (512,)):
        lib.aadff_debug_set_flags(flags)
        for _ in range(3):
            out = aadff_b200.local_psf_render(img, psf, ks)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = aadff_b200.local_psf_render(img, psf, ks)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        lib.aadff_debug_set_flags(0)
        ms = sorted(ts)[len(ts) // 2]
        gb = n * H * W * (4 * ks * ks + 24) / 1e9
        if ref is None:
            ref = out.clone()
        err = float((out - ref).abs().max())
        print(f"ks={ks:2d} flags={flags:4d} {n}x3x{H}x{W}: {ms:8.3f} ms  {gb / ms * 1e3:7.1f} GB/s  "
              f"{gb / ms * 1e3 / peak:5.3f} of {peak:.0f}  {n * H * W / ms / 1e6:6.2f} Gpix/s  max|d| vs 512: {err:.2e}",
              flush=True)
    del img, psf

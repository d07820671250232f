"""Where could a two-term evaluation start?  Sweep of AADFF_MODE_ECON's first two-term layer group (debug flags bits
16..19) on the rf50mm checkpoint: worst-case certificate (half the L1 distance of the PSFs to the fp32 PSFs = the largest
error any [0,1] image can show at a pixel), max-abs on a 5 x 512^2 noise image and on the c2 workload, executed MMA terms.
    python tests/gpu_econ_sweep.py [n_probes]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from aadff_b200 import synthetic  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_econ_certificate import probes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
lib = aadff_b200.native.lib
lens = aadff_b200.PSFNet(kernel_size=11, device="cuda")
lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
x = probes(n, 5)
ref = lens.pred(x).double()
img, dm = synthetic.synthetic_rgbd(1, 512, 512, seed=7)
foc = -synthetic.synthetic_focus(dm, 5).cuda() * 1e3
img, dep = img.cuda(), -dm.cuda() * 1e3
noise = torch.rand(1, 3, 512, 512, generator=torch.Generator().manual_seed(1)).cuda()
ref_c2 = lens.render_stack(img, dep, foc, mode="fp32")
ref_noise = lens.render_stack(noise, dep, foc, mode="fp32")
# cost in layer units: L1 = 0.25 (K = 64), L2..L9 = 1 each, head = 121 / 256 -> group g of 10
units = [0.25] + [1.0] * 8 + [128 / 256]
for first in (None, 4, 5, 6, 7, 8, 9, 10):
    mode = "parity" if first is None else "econ"
    lib.aadff_debug_set_flags(0 if first in (None, 4) else first << 16)
    l1 = (lens.pred(x, mode=mode).double() - ref).abs().sum((-1, -2)) / 2
    q = torch.quantile(l1[:1 << 20].float(), torch.tensor([0.999, 0.9999], device="cuda"))
    e_c2 = float((lens.render_stack(img, dep, foc, mode=mode) - ref_c2).abs().max())
    e_noise = float((lens.render_stack(noise, dep, foc, mode=mode) - ref_noise).abs().max())
    for _ in range(2):
        lens.render_stack(img, dep, foc, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lens.render_stack(img, dep, foc, mode=mode)
    e1.record()
    torch.cuda.synchronize()
    rate = 10 * 5 * 512 * 512 / e0.elapsed_time(e1) / 1e3
    lib.aadff_debug_set_flags(0)
    f = 10 if first is None else first
    terms = (sum(units[:f]) * 3 + sum(units[f:]) * 2) / sum(units)
    name = "parity (3 terms everywhere)" if first is None else f"2 terms from group {first} ({'L%d' % (first + 1) if first < 9 else 'head only' if first == 9 else 'nothing'})"
    print(f"[econ sweep] {name:38s} avg terms {terms:.3f} -> ceiling {1 / terms:.3f}  L1/2 max {float(l1.max()):.3e} p99.99 {float(q[1]):.3e} "
          f"p99.9 {float(q[0]):.3e}  max-abs c2 {e_c2:.2e}  noise image {e_noise:.2e}  c2 {rate:.0f} Mpix*slices/s"
          f"{'' if first in (None, 4, 10) else ' (generic kernel)'}", flush=True)

"""Diagnostic (GPU): worst-case certificate of a reduced-term mode.  For a probe (x, y, z, foc_z) the largest image
error ANY [0,1] image can show at that pixel is half the L1 distance between the mode's PSF and the fp32 PSF
(both sum to 1).  Prints max / quantiles of L1/2 over N probes incl. faces and corners of the input box, for the
shipped checkpoint and seeded random networks.      python tests/gpu_econ_certificate.py [n_probes]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aadff_b200  # noqa: E402
from aadff_b200 import synthetic  # noqa: E402


def probes(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 4, generator=g)
    x[:, :2] = x[:, :2] * 2 - 1
    m = n // 20
    x[:m, 0] = torch.sign(x[:m, 0]); x[m:2 * m, 1] = torch.sign(x[m:2 * m, 1])
    x[2 * m:3 * m, 2] = torch.round(x[2 * m:3 * m, 2]); x[3 * m:4 * m, 3] = torch.round(x[3 * m:4 * m, 3])
    x[4 * m:5 * m, :2] = torch.sign(x[4 * m:5 * m, :2]); x[5 * m:6 * m, 2:] = torch.round(x[5 * m:6 * m, 2:])
    return x.cuda()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    nets = [("rf50mm ckpt k=11", 11, None)] + [(f"seeded k={ks} seed={s}", ks, s) for ks, s in ((11, 0), (11, 1), (31, 0), (7, 2))]
    for name, ks, seed in nets:
        lens = aadff_b200.PSFNet(kernel_size=ks, device="cuda")
        if seed is None:
            lens.load_net(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"))
        else:
            Ws, bs = synthetic.seeded_psfnet_weights(ks, seed=seed)
            g = torch.Generator().manual_seed(100 + seed)
            sd = {}
            for i, (W, b) in enumerate(zip(Ws, bs)):
                sd[f"net.{2 * i}.weight"], sd[f"net.{2 * i}.bias"] = W, (torch.rand(b.shape, generator=g) - 0.5) * 0.2
            lens.psfnet.load_state_dict(sd)
        x = probes(n if ks < 20 else n // 8, 5)
        ref = lens.pred(x).double()
        for mode in ("parity", "econ8", "econ", "mixed", "fast"):
            l1 = (lens.pred(x, mode=mode).double() - ref).abs().sum((-1, -2)) / 2
            q = torch.quantile(l1[:1 << 20].float(), torch.tensor([0.999, 0.9999], device="cuda"))
            print(f"[certificate] {name:24s} {mode:7s} probes={x.shape[0]} L1/2 max {float(l1.max()):.3e} "
                  f"p99.99 {float(q[1]):.3e} p99.9 {float(q[0]):.3e} mean {float(l1.mean()):.3e}", flush=True)


if __name__ == "__main__":
    main()

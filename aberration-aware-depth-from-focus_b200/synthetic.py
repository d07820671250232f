"""Synthetic workloads for bench.py / smoke runs: seeded RGB-D batches, focus distances and PSFNet weights of the
shapes BASELINE.json names.  Data generation only -- nothing here touches the hot path.

The CPU checker used by the tests carries its own copy of these generators; tests/test_host_cpu.py checks that both
produce identical tensors, so the CPU baseline and the GPU arm of bench.py see the same inputs while the product
imports nothing from the checker.
"""
from __future__ import annotations

import math

import torch


def synthetic_rgbd(N: int, H: int, W: int, seed: int):
    """img ~ U[0,1]; depth = 3 random planes + 8 random rectangles rescaled to [0.5, 5] m, 0.5 % invalid (0) pixels
    (Middlebury has 0.6 %).  Returns (img [N,3,H,W], depth_m [N,1,H,W])."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(N, 3, H, W, generator=g)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    d = torch.zeros(N, H, W)
    for n in range(N):
        for _ in range(3):
            a, b, c = torch.rand(3, generator=g).tolist()
            d[n] += a * xx + b * yy + c
        for _ in range(8):
            y0, x0, hh, ww, v = torch.rand(5, generator=g).tolist()
            y1, x1 = int(y0 * H), int(x0 * W)
            d[n, y1:y1 + max(1, int(hh * H / 2)), x1:x1 + max(1, int(ww * W / 2))] += 2 * v
        lo, hi = d[n].min(), d[n].max()
        d[n] = 0.5 + 4.5 * (d[n] - lo) / (hi - lo + 1e-12)
    invalid = torch.rand(N, H, W, generator=g) < 0.005
    d[invalid] = 0.0
    return img, d.unsqueeze(1)


def synthetic_focus(depth_m: torch.Tensor, S: int) -> torch.Tensor:
    """Focus distances [N,S] in metres: dff.utils.select_focus_dist (dff/utils.py:4-51) for S > 3 -- the reference
    asserts num > 3 -- else a linspace over the valid depth range."""
    B = depth_m.shape[0]
    d = depth_m.reshape(B, -1)
    dmax = d.amax(dim=1)
    dmin = torch.stack([d[i][d[i] > 0].min() for i in range(B)])
    if S == 1:
        return ((dmin + dmax) / 2).view(B, 1)
    steps = torch.stack([dmin + i * (dmax - dmin) / (S - 1) for i in range(S)], dim=1)
    return torch.sort(steps, dim=-1)[0] if S > 3 else steps


def seeded_psfnet_weights(ks: int, seed: int = 0):
    """Random PSFNet weights for kernel sizes without a shipped checkpoint: the reference's initialiser
    (psfnet_arch.py:251-264: kaiming_uniform_ weights, zero bias) over MLP(4, ks^2, 256, 8) (psfnet.py:58), seeded."""
    g = torch.Generator().manual_seed(seed)
    dims = [4, 64, 256] + [256] * 8 + [ks * ks]
    Ws, bs = [], []
    for fin, fout in zip(dims[:-1], dims[1:]):
        bound = math.sqrt(2.0) * math.sqrt(3.0 / fin)
        Ws.append((torch.rand(fout, fin, generator=g) * 2 - 1) * bound)
        bs.append(torch.zeros(fout))
    return Ws, bs


def split_state_dict(state_dict):
    """state_dict of deeplens.psfnet_arch.MLP (keys net.{0,2,...}.weight / .bias) -> (weights, biases) in layer order."""
    idx = sorted({int(k.split(".")[1]) for k in state_dict if k.startswith("net.")})
    return ([state_dict[f"net.{i}.weight"].float() for i in idx], [state_dict[f"net.{i}.bias"].float() for i in idx])

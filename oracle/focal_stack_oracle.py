"""CPU oracle for the aberrated focal-stack synthesis path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
file; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker or as the
timed CPU baseline -- never as the thing shipped.

It restates, in plain fp32 torch-CPU / numpy arithmetic, the algorithm of the
reference (singer-yang/Aberration-Aware-Depth-from-Focus):

* ``depth2z``            <- deeplens/psfnet.py:447-450
* ``coord_grid``         <- deeplens/psfnet.py:427-431 (torch.linspace semantics)
* ``mlp_forward``        <- deeplens/psfnet_arch.py:24-47 (MLP.net + L1 normalise)
* ``local_psf_render``   <- deeplens/render_psf.py:76-107 (per-pixel PSF gather)
* ``render``             <- deeplens/psfnet.py:393-441 (4-D and 3-D branches)
* ``render_stack``       <- 2_aber_aware_dff_aif.py:108-114 (slice loop + stack)
* ``thinlens_psf`` / ``thinlens_render`` <- deeplens/psfnet.py:503-570
* ``select_focus_dist``  <- dff/utils.py:4-51 ('linear' mode)
* ``render_reference_ops`` follows the reference's *operator sequence*
  (pad -> unfold -> mul -> sum) so that timing it on CPU is representative of
  what the reference itself costs; the other functions use the direct
  definition of the gather, which needs k^2 times less memory.

Third-party arithmetic: everything the reference computes is done by torch
(ATen / MKL); the reference pins no version, this container has
torch 2.11.0+cu128.  Parity pin: the reference ships no tests or golden
vectors, so this oracle is pinned against outputs of the reference itself,
run in the build container by ``tests/golden/make_golden.py`` and committed as
``tests/golden/*.npz`` (see ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch

DMIN = 200.0      # [mm]  deeplens/psfnet.py:11
DMAX = 20000.0    # [mm]  deeplens/psfnet.py:12


# --------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------
def depth2z(depth: torch.Tensor, d_min: float = -DMIN, d_max: float = -DMAX) -> torch.Tensor:
    """z = clamp((d - d_min) / (d_max - d_min), 0, 1)   (deeplens/psfnet.py:447-450)."""
    z = (depth - d_min) / (d_max - d_min)
    return torch.clamp(z, min=0, max=1)


def linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """fp32 restatement of torch.linspace (ATen RangeFactories, CPU): with
    step = (end-start)/(steps-1) in fp32, element i is fma(step, i, start) in the
    first half and fma(-step, steps-1-i, end) in the second half.  The fused
    multiply-add (one rounding) is emulated in float64, where the product is exact."""
    start = np.float32(start)
    end = np.float32(end)
    if steps == 1:
        return np.array([start], dtype=np.float32)
    step = np.float64(np.float32((end - start) / np.float32(steps - 1)))
    i = np.arange(steps, dtype=np.float64)
    half = steps // 2
    lo = (np.float64(start) + step * i).astype(np.float32)
    hi = (np.float64(end) - step * (steps - 1 - i)).astype(np.float32)
    return np.where(np.arange(steps) < half, lo, hi).astype(np.float32)


def coord_grid(H: int, W: int) -> tuple[torch.Tensor, torch.Tensor]:
    """x[h,w] = linspace(-1,1,W)[w], y[h,w] = linspace(1,-1,H)[h]  (psfnet.py:427-431)."""
    xs = torch.from_numpy(linspace_f32(-1.0, 1.0, W))
    ys = torch.from_numpy(linspace_f32(1.0, -1.0, H))
    x = xs.view(1, W).expand(H, W)
    y = ys.view(H, 1).expand(H, W)
    return x, y


def split_state_dict(state_dict) -> tuple[list[torch.Tensor], list[torch.Tensor]]:
    """``net.{0,2,...}.weight/bias`` -> ordered lists (weight [out,in], bias [out])."""
    idx = sorted({int(k.split(".")[1]) for k in state_dict})
    Ws = [state_dict[f"net.{i}.weight"].detach().to("cpu", torch.float32) for i in idx]
    bs = [state_dict[f"net.{i}.bias"].detach().to("cpu", torch.float32) for i in idx]
    return Ws, bs


def mlp_forward(Ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], o: torch.Tensor,
                dtype=torch.float32) -> torch.Tensor:
    """[...,4] -> [...,k^2]: Linear+ReLU chain, Sigmoid head, L1 normalise
    (deeplens/psfnet_arch.py:31-47; F.normalize p=1 clamps the norm at 1e-12)."""
    h = o.to(dtype)
    n = len(Ws)
    for l, (W, b) in enumerate(zip(Ws, bs)):
        h = torch.nn.functional.linear(h, W.to(dtype), b.to(dtype))
        h = torch.relu(h) if l < n - 1 else torch.sigmoid(h)
    denom = h.abs().sum(dim=-1, keepdim=True).clamp_min(1e-12)
    return h / denom


def local_psf_render(img: torch.Tensor, psf: torch.Tensor, ks: int) -> torch.Tensor:
    """out[n,c,h,w] = sum_{i,j} img[n,c,clamp(h+i-r),clamp(w+j-r)] * psf[n,h,w,i,j]
    (definition of deeplens/render_psf.py:76-107: replicate pad, no kernel flip,
    PSF indexed by the output pixel, shared by all channels)."""
    if img.dim() < 4:
        img = img.unsqueeze(0)
    N, C, H, W = img.shape
    r = int((ks - 1) / 2)
    psf = psf.reshape(N, H, W, ks, ks)
    pad = torch.nn.functional.pad(img, (r, r, r, r), mode="replicate")
    out = torch.zeros_like(img)
    for i in range(ks):
        for j in range(ks):
            out += pad[:, :, i:i + H, j:j + W] * psf[:, None, :, :, i, j]
    return out


def model_input(depth: torch.Tensor, foc_dist: torch.Tensor,
                d_min: float = -DMIN, d_max: float = -DMAX) -> torch.Tensor:
    """o[n,h,w,:] = (x, y, z, foc_z)  (psfnet.py:426-437).  depth [N,H,W], foc [N]."""
    N, H, W = depth.shape
    x, y = coord_grid(H, W)
    z = depth2z(depth.float(), d_min, d_max)
    fz = depth2z(foc_dist.float().view(N, 1, 1).expand(N, H, W), d_min, d_max)
    return torch.stack((x.expand(N, H, W), y.expand(N, H, W), z, fz), -1).float()


def render(Ws, bs, img, depth, foc_dist, ks: int, dtype=torch.float32) -> torch.Tensor:
    """PSFNet.render (psfnet.py:393-441).  4-D: img [N,C,H,W], depth [N,1,H,W], foc [N];
    3-D: img [C,H,W], depth [H,W], scalar foc -> [1,C,H,W]."""
    if img.dim() == 3:
        img = img.unsqueeze(0)
        depth = depth.reshape(1, *depth.shape[-2:])
        foc_dist = torch.as_tensor([float(foc_dist)], dtype=torch.float32)
    else:
        depth = depth.reshape(depth.shape[0], *depth.shape[-2:])
    o = model_input(depth, foc_dist)
    psf = mlp_forward(Ws, bs, o, dtype)
    return local_psf_render(img.to(dtype), psf, ks)


def render_stack(Ws, bs, img, depth, foc_dists, ks: int, dtype=torch.float32) -> torch.Tensor:
    """S calls of ``render`` stacked on dim 2 -> [N,C,S,H,W] (2_aber_aware_dff_aif.py:108-114)."""
    S = foc_dists.shape[1]
    return torch.stack([render(Ws, bs, img, depth, foc_dists[:, s], ks, dtype) for s in range(S)], dim=2)


# --------------------------------------------------------------------------
# the reference's operator sequence (for the CPU timing baseline)
# --------------------------------------------------------------------------
def render_reference_ops(Ws, bs, img, depth, foc_dist, ks: int) -> torch.Tensor:
    """Same result as ``render`` but through the operators the reference uses
    (11x linear, replicate pad, unfold, broadcast multiply, sum over k^2), so its
    CPU cost is that of deeplens/psfnet.py:424-441 + render_psf.py:96-107."""
    N, C, H, W = img.shape
    o = model_input(depth.reshape(N, H, W), foc_dist)
    psf = mlp_forward(Ws, bs, o)                                   # [N,H,W,k^2]
    r = int((ks - 1) / 2)
    pad = torch.nn.functional.pad(img, (r, r, r, r), mode="replicate")
    cols = torch.nn.functional.unfold(pad, (ks, ks)).view(N, C, ks * ks, H * W)
    taps = torch.stack(C * [psf.reshape(-1, ks, ks)], 1)           # the reference's C-fold copy
    taps = taps.view(N, H * W, C, ks * ks).permute(0, 2, 3, 1)
    return (cols * taps).sum(2).view(N, C, H, W)


# --------------------------------------------------------------------------
# next rows of the scope table
# --------------------------------------------------------------------------
def thinlens_coc(depth, foc_dist, foc_len, fnum, ps):
    """ThinLens.coc (psfnet.py:503-511): circle of confusion in pixels, >= 0.1."""
    if (depth < 0).any():
        depth = -depth
        foc_dist = -foc_dist
    depth = torch.clamp(depth, DMIN, DMAX)
    coc = foc_len / fnum * torch.abs(depth - foc_dist) / depth * foc_len / (foc_dist - foc_len)
    return torch.clamp(coc / ps, min=0.1)


def thinlens_psf(depth, foc_dist, ks, foc_len, fnum, ps):
    """ThinLens.render's PSF (psfnet.py:549-566): clipped Gaussian, sigma = coc/2."""
    N, _, H, W = depth.shape
    foc = foc_dist.view(N, 1, 1, 1).expand(N, 1, H, W)
    lin = torch.linspace(-ks / 2 + 1 / 2, ks / 2 - 1 / 2, ks)
    x, y = torch.meshgrid(lin, torch.linspace(ks / 2 - 1 / 2, -ks / 2 + 1 / 2, ks), indexing="xy")
    rad = (thinlens_coc(depth, foc, foc_len, fnum, ps).squeeze(1) / 2)[..., None, None]
    r2 = x ** 2 + y ** 2
    psf = torch.exp(-r2 / 2 / rad ** 2) / (2 * np.pi * rad ** 2)
    psf = psf * (r2 < rad ** 2)
    return psf / psf.sum((-1, -2), keepdim=True)


def thinlens_render(img, depth, foc_dist, ks, foc_len, fnum, ps):
    return local_psf_render(img, thinlens_psf(depth, foc_dist, ks, foc_len, fnum, ps), ks)


def select_focus_dist(depth: torch.Tensor, num: int) -> torch.Tensor:
    """dff/utils.py:4-51, mode='linear': linear between min(valid depth) and max depth."""
    assert num > 3, "Focal stack size is too small"
    B = depth.shape[0]
    d = depth.reshape(B, -1)
    dmax = d.amax(dim=1)
    dmin = torch.stack([d[i][d[i] > 0].min() for i in range(B)])
    out = torch.stack([dmin + i * (dmax - dmin) / (num - 1) for i in range(num)], dim=1)
    return torch.sort(out, dim=-1)[0]


# --------------------------------------------------------------------------
# synthetic workload generator shared by tests and bench
# --------------------------------------------------------------------------
def synthetic_rgbd(N: int, H: int, W: int, seed: int):
    """img U[0,1]; depth = 3 planes + 8 rectangles rescaled to [0.5,5] m with 0.5 %
    invalid (0) pixels.  Returns (img [N,3,H,W], depth_m [N,1,H,W])."""
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
    """focus distances [N,S] in metres: select_focus_dist for S>3, else linspace min..max."""
    if S > 3:
        return select_focus_dist(depth_m, S)
    B = depth_m.shape[0]
    d = depth_m.reshape(B, -1)
    dmax = d.amax(dim=1)
    dmin = torch.stack([d[i][d[i] > 0].min() for i in range(B)])
    if S == 1:
        return ((dmin + dmax) / 2).view(B, 1)
    return torch.stack([dmin + i * (dmax - dmin) / (S - 1) for i in range(S)], dim=1)


def seeded_psfnet_weights(ks: int, seed: int = 0):
    """Random PSFNet weights for kernel sizes with no shipped checkpoint: the same
    initialiser the reference applies (psfnet_arch.py:251-264 -> kaiming_uniform_ on
    weights, zero bias) over MLP(4, ks^2, 256, 8) (psfnet.py:58), seeded."""
    g = torch.Generator().manual_seed(seed)
    dims = [4, 64, 256] + [256] * 8 + [ks * ks]
    Ws, bs = [], []
    for fin, fout in zip(dims[:-1], dims[1:]):
        bound = math.sqrt(2.0) * math.sqrt(3.0 / fin)          # kaiming_uniform_, a=0, fan_in
        Ws.append((torch.rand(fout, fin, generator=g) * 2 - 1) * bound)
        bs.append(torch.zeros(fout))
    return Ws, bs

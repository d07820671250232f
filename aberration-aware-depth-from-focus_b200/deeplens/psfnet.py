"""PSFNet host wrapper (mirror of the reference's deeplens/psfnet.py:14-76, 375-454, 489-570).

Same constructor, attributes and method signatures as the reference so that 0_warm_up.py,
dff/factory.get_lens and the 2_aber_aware_dff_* training loops run unchanged; ``render`` and
``pred`` dispatch through the C ABI (include/aadff.h) into the sm_100a kernels.

Differences, all deliberate:
  * the ray-traced lens model behind ``Lensgroup`` is out of scope: ``filename`` is recorded, not
    parsed, and ``analysis()`` only logs;
  * ``load_net`` passes ``map_location`` (the reference's bare torch.load fails on a box whose
    device differs from the one the checkpoint was saved on);
  * an input of unsupported rank raises instead of silently returning None;
  * ``render_stack`` renders all focal slices of a batch in one launch, writing [B,C,S,H,W]
    (AiFNet layout) or [B,S,C,H,W] (DFVNet layout) directly.
"""
import logging
import os

import numpy as np
import torch
import torch.nn as nn

import aadff_native as _nat
from .psfnet_arch import MLP, initialize_weights
from .render_psf import local_psf_render

DMIN = 200     # [mm]
DMAX = 20000   # [mm]


def _device_index(device) -> int:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"aadff-b200 runs on CUDA (sm_100a) only, got device '{device}'; there is no CPU path")
    return dev.index if dev.index is not None else torch.cuda.current_device()


class PSFNet(nn.Module):
    def __init__(self, filename=None, model_name='mlp', kernel_size=11, sensor_res=(512, 512), device='cuda',
                 mode=None):
        super().__init__()
        self.filename = filename
        self.sensor_res = sensor_res
        self.device = device
        self.in_features = 4
        self.kernel_size = kernel_size
        self.model_name = model_name
        self.mode = mode or os.environ.get("AADFF_MODE", "parity")
        self.d_max = -DMAX
        self.d_min = -DMIN
        self.foc_d_arr = np.array([-500, -600, -700, -800, -900, -1000, -1250, -1500, -1750, -2000,
                                   -2500, -3000, -4000, -5000, -6000, -8000, -10000, -12000, -15000, -20000])
        self.foc_z_arr = (self.foc_d_arr - self.d_min) / (self.d_max - self.d_min)
        self._native = None
        self._native_sig = None
        self.init_net()

    # ------------------------------------------------------------------ network
    def init_net(self):
        ks = self.kernel_size
        if self.model_name == 'mlp':
            self.psfnet = MLP(in_features=4, out_features=ks ** 2, hidden_features=256, hidden_layers=8)
        elif self.model_name in ('mlpconv', 'siren'):
            raise NotImplementedError(f"'{self.model_name}' is not on the focal-stack synthesis path")
        else:
            raise Exception('Unsupported PSF network architecture.')
        self.psfnet.apply(initialize_weights)
        self.psfnet.to(self.device)
        self.psfnet._evaluator = self._mlp_eval
        self._native = None

    def load_net(self, net_path):
        """Load pretrained network (state_dict with keys net.{0,2,...}.{weight,bias})."""
        self.psfnet.load_state_dict(torch.load(net_path, map_location=self.device))
        self._native = None

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.psfnet.parameters())

    def native(self) -> "_nat.NativePSFNet":
        """The pre-packed device copy of the weights; rebuilt when the parameters change."""
        sig = self._signature()
        if self._native is None or sig != self._native_sig:
            layers = self.psfnet.linear_layers()
            idx = _device_index(layers[0].weight.device)
            ws = [l.weight.detach().float().cpu().numpy() for l in layers]
            bs = [l.bias.detach().float().cpu().numpy() for l in layers]
            if self._native is not None:
                self._native.close()
            self._native = _nat.NativePSFNet(ws, bs, self.kernel_size, idx)
            self._native_sig = sig
        return self._native

    def analysis(self, *args, **kwargs):
        logging.info("PSFNet.analysis(): lens plots/ray tracing are outside the synthesis path; skipped.")

    # ------------------------------------------------------------------ inference
    def _mlp_eval(self, inp, mode="fp32"):
        nat = self.native()
        flat = inp.detach().reshape(-1, 4).to(f"cuda:{nat.device_index}", torch.float32).contiguous()
        out = torch.empty(flat.shape[0], self.kernel_size ** 2, device=flat.device, dtype=torch.float32)
        if flat.shape[0] == 0:
            return out.reshape(*inp.shape[:-1], self.kernel_size ** 2)
        with torch.cuda.device(flat.device):
            _nat.check(_nat.lib.aadff_psfnet_pred_tc_f32(nat.handle, flat.data_ptr(), out.data_ptr(), flat.shape[0],
                                                         _nat.MODES[mode], torch.cuda.current_stream().cuda_stream))
        return out.reshape(*inp.shape[:-1], self.kernel_size ** 2)

    def pred(self, inp, mode=None):
        """inp [...,4] = (x, y, z, foc_z) -> psf [..., ks, ks].  mode=None: fp32 CUDA-core kernel (operation for
        operation with the reference); 'parity' / 'econ' / 'mixed' / 'fast': the tensor-core kernel (~30x faster)."""
        psf = self.psfnet(inp) if mode is None else self._mlp_eval(inp, mode)
        return psf.reshape(*psf.shape[:-1], self.kernel_size, self.kernel_size)

    def _launch(self, img, depth, foc, out, strides, mode):
        nat = self.native()
        N, C, H, W = img.shape
        S = foc.shape[1]
        if out.numel() == 0:            # empty batch: nothing to launch (data_ptr() of an empty tensor is NULL)
            return
        arr = (_nat.ctypes.c_int64 * 5)(*strides)
        with torch.cuda.device(img.device):
            _nat.check(_nat.lib.aadff_render_stack_f32(
                nat.handle, img.data_ptr(), depth.data_ptr(), foc.data_ptr(), out.data_ptr(), arr,
                N, C, S, H, W, float(self.d_min), float(self.d_max), _nat.MODES[mode or self.mode],
                torch.cuda.current_stream().cuda_stream))

    def _prep(self, img, depth):
        nat = self.native()
        dev = torch.device(f"cuda:{nat.device_index}")
        img = img.detach().to(dev, torch.float32).contiguous()
        N, C, H, W = img.shape
        depth = depth.detach().to(dev, torch.float32).reshape(N, H, W).contiguous()
        return img, depth, dev

    @torch.no_grad()
    def render(self, img, depth, foc_dist, mode=None):
        """Render image with aif image and depth map.

        img [N,C,H,W], depth [N,1,H,W] (mm, negative), foc_dist [N] (mm, negative) -> [N,C,H,W];
        or img [C,H,W], depth [H,W], scalar foc_dist -> [1,C,H,W].
        """
        if img.dim() == 3:
            img = img.unsqueeze(0)
            depth = depth.reshape(1, *depth.shape[-2:])
            foc_dist = torch.as_tensor([float(foc_dist)], dtype=torch.float32)
        elif img.dim() != 4:
            raise ValueError(f"render expects a [N,C,H,W] or [C,H,W] image, got shape {tuple(img.shape)}")
        img, depth, dev = self._prep(img, depth)
        N, C, H, W = img.shape
        foc = torch.as_tensor(foc_dist).detach().to(dev, torch.float32).reshape(N, 1).contiguous()
        out = torch.empty_like(img)
        self._launch(img, depth, foc, out, (C * H * W, H * W, 0, W, 1), mode)
        return out

    @torch.no_grad()
    def render_stack(self, img, depth, foc_dists, layout="BCSHW", mode=None):
        """All S focal slices in one launch: foc_dists [N,S] (mm, negative) -> [N,C,S,H,W]
        (== torch.stack([render(img, depth, foc_dists[:, s]) for s], dim=2)), or [N,S,C,H,W]."""
        if img.dim() != 4:
            raise ValueError("render_stack expects a [N,C,H,W] image")
        img, depth, dev = self._prep(img, depth)
        N, C, H, W = img.shape
        foc = foc_dists.detach().to(dev, torch.float32).reshape(N, -1).contiguous()
        S = foc.shape[1]
        if layout == "BCSHW":
            out = torch.empty(N, C, S, H, W, device=dev, dtype=torch.float32)
            strides = (C * S * H * W, S * H * W, H * W, W, 1)
        elif layout == "BSCHW":
            out = torch.empty(N, S, C, H, W, device=dev, dtype=torch.float32)
            strides = (S * C * H * W, H * W, C * H * W, W, 1)
        else:
            raise ValueError("layout must be 'BCSHW' or 'BSCHW'")
        self._launch(img, depth, foc, out, strides, mode)
        return out

    @torch.no_grad()
    def simulate_focal_stack(self, aif, depth_m, n_stack, layout="BCSHW", mode=None):
        """The focal-stack simulation block of the training scripts (2_aber_aware_dff_aif.py:101-114) in one
        call with no host synchronisation: focus distances from `select_focus_dist(depth_m, n_stack, 'linear')`
        (metres, device-side reductions), then ONE fused launch for all slices.
        aif [B,C,H,W] in [0,1], depth_m [B,1,H,W] metres (0 = invalid) -> (stack [B,C,S,H,W], focus_dists [B,S] m)."""
        from dff.utils import select_focus_dist
        focus_dists = select_focus_dist(depth_m, n_stack, mode='linear')
        stack = self.render_stack(aif, -depth_m * 1e3, -focus_dists * 1e3, layout=layout, mode=mode)
        return stack, focus_dists

    # ------------------------------------------------------------------ utils
    def depth2z(self, depth):
        z = (depth - self.d_min) / (self.d_max - self.d_min)
        return torch.clamp(z, min=0, max=1)

    def z2depth(self, z):
        return z * (self.d_max - self.d_min) + self.d_min


class ThinLens(nn.Module):
    """Thin-lens baseline (mirror of deeplens/psfnet.py:489-570): clipped-Gaussian PSF from the
    circle of confusion, rendered with the same per-pixel gather kernel."""

    def __init__(self, foc_len, fnum, kernel_size, sensor_size, sensor_res, device='cpu'):
        super().__init__()
        self.d_max = DMAX
        self.d_min = DMIN
        self.kernel_size = kernel_size
        self.foc_len = foc_len
        self.fnum = fnum
        self.sensor_size = sensor_size
        self.sensor_res = sensor_res
        self.ps = self.sensor_size[0] / self.sensor_res[0]
        self.device = device

    def to(self, device):
        self.device = device
        return self

    def coc(self, depth, foc_dist):
        if (depth < 0).any():
            depth = -depth
            foc_dist = -foc_dist
        depth = torch.clamp(depth, self.d_min, self.d_max)
        coc = self.foc_len / self.fnum * torch.abs(depth - foc_dist) / depth * self.foc_len / (foc_dist - self.foc_len)
        return torch.clamp(coc / self.ps, min=0.1)

    @torch.no_grad()
    def render(self, img, depth, foc_dist):
        """img [N,C,H,W], depth [N,1,H,W], foc_dist [N] -> [N,C,H,W] (fused CUDA kernel: coc -> clipped
        Gaussian PSF -> gather; the PSF tensor of the reference is never materialised)."""
        if img.dim() != 4:
            raise ValueError("ThinLens.render expects a [N,C,H,W] image")
        if not img.is_cuda:
            raise RuntimeError("ThinLens.render: CUDA tensors required (no CPU fallback in this build)")
        N, C, H, W = img.shape
        img = img.detach().contiguous().float()
        dep = depth.detach().to(img.device, torch.float32).reshape(N, H, W).contiguous()
        foc = torch.as_tensor(foc_dist).detach().to(img.device, torch.float32).reshape(N).contiguous()
        out = torch.empty_like(img)
        if out.numel() == 0:
            return out
        flip = int(bool((dep < 0).any()))          # the reference's data-dependent sign convention (psfnet.py:504)
        with torch.cuda.device(img.device):
            _nat.check(_nat.lib.aadff_thinlens_render_f32(
                img.data_ptr(), dep.data_ptr(), foc.data_ptr(), out.data_ptr(), N, C, H, W, int(self.kernel_size),
                float(self.foc_len), float(self.fnum), float(self.ps), float(self.d_min), float(self.d_max), flip,
                torch.cuda.current_stream().cuda_stream))
        return out

    @torch.no_grad()
    def psf(self, depth, foc_dist):
        """The per-pixel thin-lens PSFs [N,H,W,ks,ks] as the reference builds them (psfnet.py:549-566); kept
        for inspection / for callers that want to feed local_psf_render themselves."""
        ks = self.kernel_size
        device = depth.device
        N, _, H, W = depth.shape
        foc = foc_dist.to(device).view(N, 1, 1, 1).expand(N, 1, H, W)
        lin = torch.linspace(-ks / 2 + 1 / 2, ks / 2 - 1 / 2, ks)
        x, y = torch.meshgrid(lin, torch.linspace(ks / 2 - 1 / 2, -ks / 2 + 1 / 2, ks), indexing='xy')
        x, y = x.to(device), y.to(device)
        radius = (self.coc(depth, foc).squeeze(1) / 2)[..., None, None]
        r2 = x ** 2 + y ** 2
        psf = torch.exp(-r2 / 2 / radius ** 2) / (2 * np.pi * radius ** 2)
        psf = psf * (r2 < radius ** 2)
        return psf / psf.sum((-1, -2), keepdim=True)

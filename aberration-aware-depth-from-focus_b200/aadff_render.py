"""Per-pixel PSF rendering (mirror of the reference's deeplens/render_psf.py:76-107)."""
import torch

import aadff_native as _nat


__all__ = ["local_psf_render", "render_psf", "render_psf_map", "local_psf_render_high_res"]


def local_psf_render(input, psf, kernel_size=11):
    """Blur ``input`` with a different PSF at every pixel.

    Same contract as the reference: input [N,C,H,W] (or [C,H,W] -> treated as N=1),
    psf [N,H,W,ks,ks] (any shape that reshapes to it), returns [N,C,H,W]; replicate border,
    no kernel flip, one PSF shared by all channels.  CUDA only.
    """
    if input.dim() < 4:
        input = input.unsqueeze(0)
    if not input.is_cuda:
        raise RuntimeError("local_psf_render: CUDA tensors required (no CPU fallback in this build)")
    n, c, h, w = input.shape
    ks = int(kernel_size)
    img = input.detach().contiguous().float()
    taps = psf.detach().to(img.device).contiguous().float()
    if taps.numel() != n * h * w * ks * ks:
        raise ValueError(f"psf has {taps.numel()} elements, expected {n}*{h}*{w}*{ks}*{ks}")
    out = torch.empty_like(img)
    with torch.cuda.device(img.device):
        _nat.check(_nat.lib.aadff_local_psf_render_f32(
            img.data_ptr(), taps.data_ptr(), out.data_ptr(), n, c, h, w, ks,
            torch.cuda.current_stream().cuda_stream))
    return out


def _psf_conv(img, psf_map, ks, grid):
    if not img.is_cuda:
        raise RuntimeError("render_psf / render_psf_map: CUDA tensors required (no CPU fallback in this build)")
    B, C, H, W = img.shape
    x = img.detach().contiguous().float()
    taps = psf_map.detach().to(x.device).contiguous().float()
    out = torch.zeros_like(x)            # the reference starts from zeros (rows past the last patch stay 0)
    hb = (_nat.ctypes.c_int * (grid + 1))(*[int(i / grid * H) for i in range(grid + 1)])
    wb = (_nat.ctypes.c_int * (grid + 1))(*[int(j / grid * W) for j in range(grid + 1)])
    with torch.cuda.device(x.device):
        _nat.check(_nat.lib.aadff_render_psf_map_f32(x.data_ptr(), taps.data_ptr(), out.data_ptr(), B, C, H, W, ks, grid,
                                                     hb, wb, torch.cuda.current_stream().cuda_stream))
    return out


def render_psf(img, psf):
    """Render an image with ONE PSF per channel (mirror of deeplens/render_psf.py:12-28): img [B,C,H,W],
    psf [C,ks,ks] -> [B,C,H,W]; reflect padding, true convolution (the PSF is flipped), grouped by channel."""
    _, ks, ks2 = psf.shape
    if ks != ks2 or ks % 2 == 0:
        raise ValueError("render_psf: square PSFs of odd size only (an even size changes the output shape in the reference)")
    if psf.shape[0] != img.shape[1]:
        raise ValueError("render_psf: PSF must have one kernel per image channel")
    return _psf_conv(img, psf, ks, 1)


def render_psf_map(img, psf_map, grid):
    """Render an image with a grid x grid map of PSFs, one per image patch (mirror of deeplens/render_psf.py:31-73):
    img [B,C,H,W] (or an HxWx3 uint8 array), psf_map [C, grid*ks, grid*ks] -> [B,C,H,W]."""
    if not torch.is_tensor(img):
        import numpy as np
        img = torch.tensor((img / 255.).astype(np.float32)).permute(2, 0, 1).unsqueeze(0).to(psf_map.device)
    assert len(img.shape) == 4, 'Input image should be [B, C, H, W]'
    Cpsf, Hpsf, Wpsf = psf_map.shape
    assert Hpsf % grid == 0 and Wpsf % grid == 0, 'PSF map size should be divisible by grid'
    ks = int(Hpsf / grid)
    assert ks % 2 == 1, 'PSF kernel size should be odd'
    assert img.shape[1] == Cpsf, 'PSF map should have the same channel as image'
    return _psf_conv(img, psf_map, ks, grid)


def local_psf_render_high_res(input, psf, patch_size=[320, 480], kernel_size=11):
    """Patch-based local_psf_render (mirror of deeplens/render_psf.py:110-127), kept with the reference's semantics:
    every patch is replicate-padded on its own, so a (ks-1)/2-pixel band at each seam differs from the full-frame
    gather.  The fused kernels never need it (no [N,H,W,k,k] unfold exists to bound); it is here for callers."""
    import math
    B, C, H, W = input.shape
    out = torch.zeros_like(input)
    for pi in range(int(math.ceil(H / patch_size[0]))):
        for pj in range(int(math.ceil(W / patch_size[1]))):
            li, ui = pi * patch_size[0], min((pi + 1) * patch_size[0], H)
            lj, uj = pj * patch_size[1], min((pj + 1) * patch_size[1], W)
            out[:, :, li:ui, lj:uj] = local_psf_render(input[:, :, li:ui, lj:uj], psf[:, li:ui, lj:uj, :, :],
                                                       kernel_size=kernel_size)
    return out

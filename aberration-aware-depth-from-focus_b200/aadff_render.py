"""Per-pixel PSF rendering (mirror of the reference's deeplens/render_psf.py:76-107)."""
import torch

import aadff_native as _nat


__all__ = ["local_psf_render"]


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

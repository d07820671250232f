"""Device-side sample preparation (the CPU half of the reference's Dataset classes, dff/dataset.py, after decoding).

The reference decodes, augments and resizes every sample on the CPU with ``num_workers=0``
(2_aber_aware_dff_aif.py:74); once the focal-stack simulation runs at hundreds of Mpix*slices/s that is the
bottleneck of a training step.  ``preprocess_rgbd`` does ToTensor + colour jitter + flips + the antialiased resize
for a whole batch in one launch from the raw decoded arrays (uint8 BGR image as ``cv.imread`` returns it, uint16 depth).
File decoding itself and AutoAgument's spline rotation stay with the reference.
"""
import torch

import aadff_native as _nat

__all__ = ["preprocess_rgbd"]


@torch.no_grad()
def preprocess_rgbd(bgr_u8, depth_u16, size, depth_div=4000.0, depth_mode="antialias", jitter=None, flips=None):
    """bgr_u8 [B,H,W,3] uint8 (cv.imread order) and/or depth_u16 [B,H,W] uint16/int16-as-uint16, CUDA tensors ->
    (aif [B,3,h,w] float32 RGB in [0,1], depth [B,1,h,w] float32 = raw / depth_div), resized to ``size`` = (h, w).

    depth_div: 4000 for Matterport3D (dataset.py:45), 1000 for Middlebury (:199).  depth_mode 'antialias' =
    torchvision Resize(antialias=True) as Matterport3D's transform applies it to the depth; 'cv2' = cv.resize
    INTER_LINEAR as Middlebury does.  jitter [B,2] = (contrast, brightness) of AutoAgument's colour jitter
    (:259-262), contrast < 0 = none; flips [B] uint8: bit 0 = np.flip(axis=1), bit 1 = np.flip(axis=0)."""
    ref = bgr_u8 if bgr_u8 is not None else depth_u16
    if not ref.is_cuda:
        raise RuntimeError("preprocess_rgbd: CUDA tensors required (no CPU fallback in this build)")
    dev = ref.device
    B, H, W = ref.shape[0], ref.shape[1], ref.shape[2]
    h, w = int(size[0]), int(size[1])
    aif = depth = None
    if bgr_u8 is not None:
        assert bgr_u8.dtype == torch.uint8 and bgr_u8.shape == (B, H, W, 3)
        bgr_u8 = bgr_u8.contiguous()
        aif = torch.empty(B, 3, h, w, device=dev, dtype=torch.float32)
    if depth_u16 is not None:
        assert depth_u16.dtype in (torch.uint16, torch.int16) and depth_u16.shape == (B, H, W)
        depth_u16 = depth_u16.contiguous()
        depth = torch.empty(B, 1, h, w, device=dev, dtype=torch.float32)
    jit = None if jitter is None else jitter.to(dev, torch.float32).reshape(B, 2).contiguous()
    flp = None if flips is None else flips.to(dev, torch.uint8).reshape(B).contiguous()
    ptr = lambda t: None if t is None else t.data_ptr()
    if B:
        with torch.cuda.device(dev):
            _nat.check(_nat.lib.aadff_preprocess_rgbd_u8(ptr(bgr_u8), ptr(depth_u16), ptr(aif), ptr(depth), B, H, W, h, w,
                                                         float(depth_div), {"antialias": 0, "cv2": 1}[depth_mode],
                                                         ptr(jit), ptr(flp), torch.cuda.current_stream().cuda_stream))
    return aif, depth

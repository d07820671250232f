"""Golden vectors for the f4 row: render_psf / render_psf_map of the reference
(deeplens/render_psf.py:12-73), loaded by file path and run on CPU.   python tests/golden/make_golden_f4.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_render_psf", "/root/reference/deeplens/render_psf.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

g = torch.Generator().manual_seed(404)
out = {}
for i, (B, C, H, W, ks) in enumerate([(2, 3, 40, 56, 11), (1, 1, 17, 23, 5), (1, 3, 33, 70, 31), (1, 2, 9, 9, 1)]):
    img = torch.rand(B, C, H, W, generator=g)
    psf = torch.rand(C, ks, ks, generator=g)
    psf = psf / psf.sum((1, 2), keepdim=True)
    out[f"conv{i}_img"], out[f"conv{i}_psf"], out[f"conv{i}_out"] = img, psf, ref.render_psf(img, psf)
for i, (B, C, H, W, ks, grid) in enumerate([(1, 3, 50, 64, 7, 3), (2, 3, 96, 128, 11, 4), (1, 3, 37, 41, 3, 5)]):
    img = torch.rand(B, C, H, W, generator=g)
    pm = torch.rand(C, grid * ks, grid * ks, generator=g)
    out[f"map{i}_img"], out[f"map{i}_psf"], out[f"map{i}_grid"], out[f"map{i}_out"] = img, pm, grid, ref.render_psf_map(img, pm, grid)
np.savez_compressed(os.path.join(HERE, "kat_i_psf_conv.npz"),
                    **{k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
print("written", {k: tuple(np.asarray(v).shape) for k, v in out.items() if k.endswith("_out")})

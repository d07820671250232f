"""Golden vectors at BASELINE.json's full sizes (c2, c3, c4), produced by running the REFERENCE itself.

    python tests/golden/make_golden_baseline_sizes.py      # build container only (/root/reference)

Inputs are bench.py's seeded synthetic workloads (oracle.synthetic_rgbd / synthetic_focus, seed 1234), so they are
regenerated on the GPU box instead of being stored; outputs are stored sub-sampled, with float64 sums per
(image, slice) over ALL pixels.

  c2  1 x 5 x 512 x 512, k=11, rf50mm checkpoint: all 5 slices through PSFNet.render (deeplens/psfnet.py:424-441)
  c3  16 x 5 x 256 x 256, k=11: all 16 images, all 5 slices, batched exactly as the DFV script calls render
  c4  1 x 10 x 1080 x 1920, k=31, seeded weights: the reference's unfold of a full frame needs 3 x 24 GB, so rows are
      rendered in bands: the reference's own PSFNet.pred on the band's (x, y, z, foc_z) -- coordinates built as
      psfnet.py:427-437 builds them for the FULL frame -- and its own local_psf_render on the band cut out of the
      row-padded full frame (r extra rows on either side, cropped afterwards), which is exactly what a full-frame
      call computes for those rows.  Bands: top border, middle, bottom border; slices 0, 5, 9.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, ROOT, REF      # noqa: E402

C4_BANDS = [(0, 8), (536, 544), (1072, 1080)]
C4_SLICES = [0, 5, 9]
C4_COL_STRIDE = 3


def reference_band(lens, local_psf_render, img, depth, foc_mm, h0, h1):
    """Rows [h0, h1) of lens.render(img, depth, foc) for a single image, by the reference's own pieces."""
    N, C, H, W = img.shape
    ks = lens.kernel_size
    r = (ks - 1) // 2
    # psfnet.py:426-437 for the full frame, then the band's rows
    z = lens.depth2z(depth).squeeze(1)
    x, y = torch.meshgrid(torch.linspace(-1, 1, W), torch.linspace(1, -1, H), indexing='xy')
    x, y = x.unsqueeze(0).repeat(N, 1, 1), y.unsqueeze(0).repeat(N, 1, 1)
    foc_z = lens.depth2z(foc_mm.unsqueeze(-1).unsqueeze(-1).repeat(1, H, W))
    o = torch.stack((x, y, z, foc_z), -1).float()[:, h0:h1]
    psf = lens.pred(o)                                                  # [N, hb, W, ks, ks]
    hb = h1 - h0
    rows_padded = torch.nn.functional.pad(img, (0, 0, r, r), mode='replicate')      # rows only
    band = rows_padded[:, :, h0:h1 + 2 * r, :]                                       # band with r real neighbours each side
    psf_ext = torch.zeros(N, hb + 2 * r, W, ks, ks)
    psf_ext[:, r:r + hb] = psf
    return local_psf_render(band, psf_ext, ks)[:, :, r:r + hb, :]


def main():
    PSFNet, ThinLens, ref_gather, MLP, ref_select_focus = import_reference()
    from oracle.focal_stack_oracle import synthetic_rgbd, synthetic_focus, seeded_psfnet_weights
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **{
        k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})
    ck = os.path.join(REF, "ckpt/rf50mm/PSFNet480x640_ks11.pkl")
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        lens = PSFNet(filename="./lenses/rf50mm/lens.json", sensor_res=(512, 512), kernel_size=11, device="cpu")
        lens.psfnet.load_state_dict(torch.load(ck, map_location="cpu"))

        # ---- c2
        img, dm = synthetic_rgbd(1, 512, 512, seed=1234)
        foc_m = synthetic_focus(dm, 5)
        stack = torch.stack([lens.render(img, -dm * 1e3, -foc_m[:, s] * 1e3) for s in range(5)], dim=2)   # [1,3,5,512,512]
        save("kat_c2_1x5x512x512.npz", seed=1234, foc_m=foc_m, out_sub=stack[..., ::4, ::4],
             out_rows=stack[..., [0, 1, 255, 256, 510, 511], :], sums=stack.double().sum((1, 3, 4)))
        print("c2 done", float(stack.mean()))

        # ---- c3
        img, dm = synthetic_rgbd(16, 256, 256, seed=1234)
        foc_m = synthetic_focus(dm, 5)
        stack = torch.stack([lens.render(img, -dm * 1e3, -foc_m[:, s] * 1e3) for s in range(5)], dim=2)   # [16,3,5,256,256]
        save("kat_c3_16x5x256x256.npz", seed=1234, foc_m=foc_m, out_sub=stack[..., ::8, ::8],
             out_full_img3=stack[3, :, :, ::2, ::2], sums=stack.double().sum((1, 3, 4)))
        print("c3 done", float(stack.mean()))

        # ---- c4 (k = 31, the bench's seeded weights: kaiming-uniform, zero bias)
        Ws, bs = seeded_psfnet_weights(31, seed=0)
        lens31 = PSFNet(filename="./lenses/rf50mm/lens.json", sensor_res=(480, 640), kernel_size=31, device="cpu")   # render() ignores sensor_res; the ray-traced constructor insists on the lens aspect ratio
        sd = {}
        for l, (Wl, bl) in enumerate(zip(Ws, bs)):
            sd[f"net.{2 * l}.weight"], sd[f"net.{2 * l}.bias"] = Wl, bl
        lens31.psfnet.load_state_dict(sd)
        img, dm = synthetic_rgbd(1, 1080, 1920, seed=1234)
        foc_m = synthetic_focus(dm, 10)
        # the banded evaluation against a full-frame reference call where that still fits: a 96 x 128 frame
        small_img, small_dm = synthetic_rgbd(1, 96, 128, seed=77)
        full = lens31.render(small_img, -small_dm * 1e3, torch.tensor([-1500.]))
        for (h0, h1) in [(0, 8), (40, 56), (88, 96)]:
            band = reference_band(lens31, ref_gather, small_img, -small_dm * 1e3, torch.tensor([-1500.]), h0, h1)
            assert torch.equal(band, full[:, :, h0:h1]), (h0, h1, float((band - full[:, :, h0:h1]).abs().max()))
        print("banded == full-frame reference on 96x128 (bit-exact)")
        bands = {}
        for s in C4_SLICES:
            for (h0, h1) in C4_BANDS:
                out = reference_band(lens31, ref_gather, img, -dm * 1e3, -foc_m[:, s] * 1e3, h0, h1)
                bands[f"s{s}_h{h0}"] = out[0, :, :, ::C4_COL_STRIDE]
                print("c4 band", s, h0, float(out.mean()))
        save("kat_c4_1x10x1080x1920_ks31.npz", seed=1234, weight_seed=0, foc_m=foc_m, col_stride=C4_COL_STRIDE,
             bands=np.asarray(C4_BANDS), slices=np.asarray(C4_SLICES), **bands)
    print("done")


if __name__ == "__main__":
    main()

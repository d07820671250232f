"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Run once in the build container (the only place /root/reference exists):

    python tests/golden/make_golden.py

It imports the reference's own classes (deeplens.psfnet.PSFNet / ThinLens,
deeplens.render_psf.local_psf_render, dff.utils.select_focus_dist) from
/root/reference with the three optional plotting/metric modules stubbed out
(they are never touched on the hot path), loads the shipped rf50mm checkpoint
with map_location='cpu', and records inputs + outputs as small .npz files.
The checkpoint (a data asset, 2.3 MB) is copied next to them because the GPU
box has no /root/reference.  Nothing here is imported by the product.
"""
import os
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def import_reference():
    for name in ["matplotlib", "matplotlib.pyplot", "lpips", "skimage", "skimage.metrics"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["skimage.metrics"].peak_signal_noise_ratio = lambda *a, **k: 0
    sys.modules["skimage.metrics"].structural_similarity = lambda *a, **k: 0
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    os.chdir(REF)
    from deeplens.psfnet import PSFNet, ThinLens
    from deeplens.render_psf import local_psf_render
    from deeplens.psfnet_arch import MLP
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_dff_utils", os.path.join(REF, "dff/utils.py"))
    dff_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dff_utils)
    return PSFNet, ThinLens, local_psf_render, MLP, dff_utils.select_focus_dist


def analytic_rgbd(N, H, W):
    """KAT-B inputs: RNG-free, recomputable anywhere."""
    n, c, h, w = np.meshgrid(np.arange(N), np.arange(3), np.arange(H), np.arange(W), indexing="ij")
    img = ((7 * h + 13 * w + 29 * c + 101 * n) % 256) / 255.0
    n, h, w = np.meshgrid(np.arange(N), np.arange(H), np.arange(W), indexing="ij")
    depth_m = 0.5 + 4.5 * (((h * W + w) * 31 + 17 * n) % 1000) / 999.0
    return (torch.tensor(img, dtype=torch.float32),
            torch.tensor(depth_m, dtype=torch.float32).unsqueeze(1))


def main():
    PSFNet, ThinLens, ref_gather, MLP, ref_select_focus = import_reference()
    from oracle.focal_stack_oracle import synthetic_rgbd, synthetic_focus, seeded_psfnet_weights

    ck_src = os.path.join(REF, "ckpt/rf50mm/PSFNet480x640_ks11.pkl")
    ck_dst = os.path.join(HERE, "rf50mm_PSFNet480x640_ks11.pkl")
    if not os.path.exists(ck_dst):
        shutil.copyfile(ck_src, ck_dst)

    lens = PSFNet(filename="./lenses/rf50mm/lens.json", sensor_res=(480, 640), kernel_size=11, device="cpu")
    lens.psfnet.load_state_dict(torch.load(ck_src, map_location="cpu"))
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **{
        k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in kw.items()})

    with torch.no_grad():
        # KAT-A: pred() on fixed probes + seeded random probes
        g = torch.Generator().manual_seed(7)
        probes = torch.tensor([[0, 0, 0, 0], [0, 0, .5, .5], [1, 1, 1, 0], [-1, .5, .25, .75],
                               [0, 0, .1, .1]], dtype=torch.float32)
        rnd = torch.rand(59, 4, generator=g)
        rnd[:, :2] = rnd[:, :2] * 2 - 1
        inp = torch.cat([probes, rnd])
        save("kat_a_pred.npz", inp=inp, psf=lens.pred(inp))

        # KAT-B: analytic images through render()
        for (N, H, W, foc) in [(1, 48, 64, [-2400.]), (2, 64, 64, [-600., -5000.])]:
            img, dm = analytic_rgbd(N, H, W)
            out = lens.render(img, -dm * 1e3, torch.tensor(foc))
            save(f"kat_b_{N}x{H}x{W}.npz", foc=np.float32(foc), out=out)
        img, dm = analytic_rgbd(1, 480, 640)
        out = lens.render(img, -dm * 1e3, torch.tensor([-2400.]))
        save("kat_b_1x480x640.npz", foc=np.float32([-2400.]), out_sub=out[:, :, ::8, ::8],
             out_rows=out[:, :, [0, 1, 239, 240, 478, 479], :],
             sum=np.float64(out.double().sum()), mean=np.float64(out.double().mean()))

        # KAT-C: depth clamps (invalid 0 and > 20 m)
        img, dm = analytic_rgbd(1, 32, 32)
        dm[:, :, :8] = 0.0
        dm[:, :, 8:16] = 30.0
        out = lens.render(img, -dm * 1e3, torch.tensor([-1000.]))
        save("kat_c_clamp.npz", foc=np.float32([-1000.]), out=out)

        # KAT-D: gather alone on arbitrary (un-normalised) PSFs, C=1/3, several k
        for i, (N, C, H, W, ks) in enumerate([(1, 3, 20, 28, 11), (2, 1, 17, 9, 5), (1, 3, 9, 33, 3),
                                              (1, 4, 12, 12, 7), (1, 3, 12, 10, 31)]):
            g = torch.Generator().manual_seed(100 + i)
            im = torch.rand(N, C, H, W, generator=g)
            psf = torch.rand(N, H, W, ks, ks, generator=g)
            save(f"kat_d_gather_{i}.npz", img=im, psf=psf, ks=ks, out=ref_gather(im, psf, ks))
        im3 = torch.rand(1, 12, 14, generator=g)      # 3-D input -> [1,1,H,W]
        psf = torch.rand(1, 12, 14, 5, 5, generator=g)
        save("kat_d_gather_3d.npz", img=im3, psf=psf, ks=5, out=ref_gather(im3, psf, 5))

        # seeded synthetic RGB-D stack (the bench generator), ragged size, S=5
        img, dm = synthetic_rgbd(2, 40, 56, seed=1234)
        foc_m = synthetic_focus(dm, 5)
        stack = torch.stack([lens.render(img, -dm * 1e3, -foc_m[:, s] * 1e3) for s in range(5)], dim=2)
        save("kat_e_stack_2x40x56.npz", img=img, depth_m=dm, foc_m=foc_m, out=stack)
        assert torch.equal(foc_m, ref_select_focus(dm, 5))
        save("kat_f_select_focus.npz", depth_m=dm, num=5, out=ref_select_focus(dm, 5),
             out8=ref_select_focus(dm, 8))

        # 3-D branch of render()
        out3 = lens.render(img[0], -dm[0, 0] * 1e3, -1500.0)
        save("kat_e_render3d.npz", img=img[0], depth_m=dm[0, 0], foc=np.float32(-1500.0), out=out3)

        # k = 31: no checkpoint exists -> seeded weights (+ random biases) in the reference MLP
        Ws, bs = seeded_psfnet_weights(31, seed=0)
        g = torch.Generator().manual_seed(31)
        bs = [(torch.rand(b.shape, generator=g) - 0.5) * 0.2 for b in bs]
        lens31 = PSFNet(filename="./lenses/rf50mm/lens.json", sensor_res=(480, 640), kernel_size=31, device="cpu")
        sd = {}
        for l, (Wl, bl) in enumerate(zip(Ws, bs)):
            sd[f"net.{2 * l}.weight"], sd[f"net.{2 * l}.bias"] = Wl, bl
        lens31.psfnet.load_state_dict(sd)
        img, dm = synthetic_rgbd(1, 40, 48, seed=4321)
        out = lens31.render(img, -dm * 1e3, torch.tensor([-1800.]))
        save("kat_g_ks31_1x40x48.npz", img=img, depth_m=dm, foc=np.float32([-1800.]),
             bias_seed=31, weight_seed=0, out=out,
             psf_probe=lens31.pred(torch.tensor([[0.1, -0.2, 0.3, 0.4]])))

        # thin-lens baseline (next-row f1)
        tl = ThinLens(foc_len=50.0, fnum=1.8, kernel_size=11, sensor_size=[36.0, 24.0], sensor_res=(40, 56))
        img, dm = synthetic_rgbd(2, 40, 56, seed=99)
        dm = dm.clamp_min(0.3)          # ThinLens has no invalid-depth handling of its own
        foc = torch.tensor([-1200., -3000.])
        save("kat_h_thinlens.npz", img=img, depth_m=dm, foc=foc, foc_len=50.0, fnum=1.8,
             sensor_size=[36.0, 24.0], out=tl.render(img, -dm * 1e3, foc))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()

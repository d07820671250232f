"""CPU: the oracle restatement against golden vectors produced by the reference
itself (tests/golden/make_golden.py).  Tolerances: the oracle uses the direct
gather definition and its own summation order, so it agrees with the reference
to fp32 rounding (<= 2e-6 on [0,1] images), not bit-for-bit."""
import numpy as np
import pytest
import torch

from oracle import focal_stack_oracle as orc
from conftest import analytic_rgbd, load_golden

TOL = 2e-6


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_linspace_restatement_is_bit_exact():
    for n in [1, 2, 3, 7, 48, 64, 255, 256, 480, 512, 640, 1080, 1920]:
        for a, b in [(-1.0, 1.0), (1.0, -1.0)]:
            assert np.array_equal(orc.linspace_f32(a, b, n), torch.linspace(a, b, n).numpy()), (n, a, b)


def test_kat_a_pred(rf50mm_weights):
    g = load_golden("kat_a_pred.npz")
    psf = orc.mlp_forward(*rf50mm_weights, T(g["inp"])).reshape(-1, 11, 11)
    assert (psf - T(g["psf"])).abs().max() < 1e-6
    # the literal values recorded in SURVEY.md section 8c
    assert abs(float(psf[0, 5, 5]) - 0.015142419) < 1e-7
    assert abs(float(psf[1, 5, 5]) - 0.810939252) < 1e-6
    assert abs(float(psf[2, 0, 0]) - 2.872148e-03) < 1e-8
    assert int(psf[2].argmax()) == 84 and int(psf[0].argmax()) == 51
    assert torch.allclose(psf.sum((-1, -2)), torch.ones(len(psf)), atol=1e-6)


@pytest.mark.parametrize("N,H,W", [(1, 48, 64), (2, 64, 64)])
def test_kat_b_render(rf50mm_weights, N, H, W):
    g = load_golden(f"kat_b_{N}x{H}x{W}.npz")
    img, dm = analytic_rgbd(N, H, W)
    out = orc.render(*rf50mm_weights, img, -dm * 1e3, T(g["foc"]), 11)
    assert (out - T(g["out"])).abs().max() < TOL


def test_kat_b_survey_literals(rf50mm_weights):
    g = load_golden("kat_b_1x48x64.npz")
    out = T(g["out"])
    assert abs(float(out.mean()) - 0.497834290) < 1e-6
    assert np.allclose(out[0, :, 24, 21].numpy(), [0.7247602, 0.8376201, 0.6984411], atol=1e-6)


def test_kat_b_full_frame(rf50mm_weights):
    g = load_golden("kat_b_1x480x640.npz")
    img, dm = analytic_rgbd(1, 480, 640)
    out = orc.render(*rf50mm_weights, img, -dm * 1e3, T(g["foc"]), 11)
    assert (out[:, :, ::8, ::8] - T(g["out_sub"])).abs().max() < TOL
    assert (out[:, :, [0, 1, 239, 240, 478, 479], :] - T(g["out_rows"])).abs().max() < TOL
    assert abs(float(out.double().sum()) - float(g["sum"])) < 0.25   # 2.7e-7 mean drift over 921 600 values
    assert abs(float(g["mean"]) - 0.499993414) < 1e-6


def test_kat_c_clamps(rf50mm_weights):
    g = load_golden("kat_c_clamp.npz")
    img, dm = analytic_rgbd(1, 32, 32)
    dm[:, :, :8] = 0.0
    dm[:, :, 8:16] = 30.0
    out = orc.render(*rf50mm_weights, img, -dm * 1e3, T(g["foc"]), 11)
    assert (out - T(g["out"])).abs().max() < TOL


@pytest.mark.parametrize("i", [0, 1, 2, 3, 4, "3d"])
def test_kat_d_gather(i):
    g = load_golden(f"kat_d_gather_{i}.npz")
    out = orc.local_psf_render(T(g["img"]), T(g["psf"]), int(g["ks"]))
    assert out.shape == g["out"].shape
    assert (out - T(g["out"])).abs().max() < 5e-5 * int(g["ks"])   # un-normalised PSFs: sums reach k^2/4


def test_kat_e_stack_and_3d(rf50mm_weights):
    g = load_golden("kat_e_stack_2x40x56.npz")
    img, dm, foc_m = T(g["img"]), T(g["depth_m"]), T(g["foc_m"])
    img2, dm2 = orc.synthetic_rgbd(2, 40, 56, seed=1234)
    assert torch.equal(img, img2) and torch.equal(dm, dm2)
    assert torch.equal(foc_m, orc.synthetic_focus(dm, 5))
    out = orc.render_stack(*rf50mm_weights, img, -dm * 1e3, -foc_m * 1e3, 11)
    assert out.shape == (2, 3, 5, 40, 56)
    assert (out - T(g["out"])).abs().max() < TOL
    one = orc.render_reference_ops(*rf50mm_weights, img, -dm * 1e3, -foc_m[:, 2] * 1e3, 11)
    assert (one - T(g["out"])[:, :, 2]).abs().max() < TOL
    g3 = load_golden("kat_e_render3d.npz")
    out3 = orc.render(*rf50mm_weights, T(g3["img"]), -T(g3["depth_m"]) * 1e3, float(g3["foc"]), 11)
    assert out3.shape == (1, 3, 40, 56)
    assert (out3 - T(g3["out"])).abs().max() < TOL


def test_kat_f_select_focus():
    g = load_golden("kat_f_select_focus.npz")
    assert torch.equal(orc.select_focus_dist(T(g["depth_m"]), 5), T(g["out"]))
    assert torch.equal(orc.select_focus_dist(T(g["depth_m"]), 8), T(g["out8"]))


def test_kat_g_ks31():
    g = load_golden("kat_g_ks31_1x40x48.npz")
    Ws, bs = orc.seeded_psfnet_weights(31, seed=int(g["weight_seed"]))
    gen = torch.Generator().manual_seed(int(g["bias_seed"]))
    bs = [(torch.rand(b.shape, generator=gen) - 0.5) * 0.2 for b in bs]
    out = orc.render(Ws, bs, T(g["img"]), -T(g["depth_m"]) * 1e3, T(g["foc"]), 31)
    assert (out - T(g["out"])).abs().max() < TOL


def test_kat_h_thinlens():
    g = load_golden("kat_h_thinlens.npz")
    out = orc.thinlens_render(T(g["img"]), -T(g["depth_m"]) * 1e3, T(g["foc"]), 11,
                              float(g["foc_len"]), float(g["fnum"]), float(g["sensor_size"][0]) / 40)
    assert (out - T(g["out"])).abs().max() < TOL


def test_oracle_matches_reference_at_baseline_sizes(rf50mm_weights):
    """The oracle against the reference's own output at BASELINE sizes (one slice of c2, one image of c3):
    tests/golden/make_golden_baseline_sizes.py ran the reference on bench.py's seeded workloads."""
    from conftest import load_golden
    g = load_golden("kat_c2_1x5x512x512.npz")
    img, dm = orc.synthetic_rgbd(1, 512, 512, seed=int(g["seed"]))
    foc_m = orc.synthetic_focus(dm, 5)
    assert torch.equal(foc_m, torch.from_numpy(g["foc_m"]))
    out = orc.render(*rf50mm_weights, img, -dm * 1e3, -foc_m[:, 2] * 1e3, 11)
    assert float((out[..., ::4, ::4] - torch.from_numpy(g["out_sub"])[:, :, 2]).abs().max()) < 2e-6
    assert abs(float(out.double().sum()) - float(g["sums"][0, 2])) < 0.1     # 786 432 values: mean bias < 1.3e-7
    g = load_golden("kat_c3_16x5x256x256.npz")
    img, dm = orc.synthetic_rgbd(16, 256, 256, seed=int(g["seed"]))
    foc_m = orc.synthetic_focus(dm, 5)
    out = orc.render(*rf50mm_weights, img[3:4], -dm[3:4] * 1e3, -foc_m[3:4, 4] * 1e3, 11)
    assert float((out[0, :, ::2, ::2] - torch.from_numpy(g["out_full_img3"])[:, 4]).abs().max()) < 2e-6


def test_spline_rotate_oracle_vs_scipy_and_reference_golden():
    """AutoAgument's rotation (dff/dataset.py:275-284): the oracle's restatement of scipy's order-3 spline rotation
    against scipy itself (prefilter: recursion and two-sided-sum form; rotation incl. multiples of 90 degrees, where
    the in/out-of-image decision sits exactly on the border) and against the full-resolution output stored in
    kat_l_rotate.npz (scipy.ndimage.rotate on the decoded float64 arrays, as the reference calls it)."""
    from oracle import spline_rotate_oracle as so
    ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(4)
    for (H, W) in [(9, 13), (40, 56), (5, 70), (2, 3)]:
        img, dep = rng.random((H, W, 3)), rng.random((H, W)) * 5
        f = ndimage.spline_filter(dep, 3, output=np.float64, mode="mirror")
        assert np.abs(f - so.prefilter_axis(so.prefilter_axis(dep, 0), 1)).max() < 1e-12
        if min(H, W) > 1:
            assert np.abs(f - so.prefilter_fir(so.prefilter_fir(dep, 0), 1)).max() < 1e-7      # z^17 truncation
        for deg in (0, 17, 45, 90, 133, 179):
            ri, rd = ndimage.rotate(img, deg, reshape=False), ndimage.rotate(dep, deg, reshape=False)
            rd[rd < 0] = 0
            oi, od = so.auto_augment_rotate(img, dep, deg)
            assert np.abs(ri - oi).max() < 1e-12 and np.abs(rd - od).max() < 1e-12, (H, W, deg)
    g = load_golden("kat_l_rotate.npz")
    a64 = g["bgr"][..., ::-1] / 255.
    oi, od = so.auto_augment_rotate(a64, g["depth"] / 4000, float(g["full_degree"]))
    assert np.abs(oi - g["full_aif"]).max() < 1e-6 and np.abs(od - g["full_depth"]).max() < 1e-6

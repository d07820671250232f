"""Golden vectors for the device-side sample preparation (f4 row): the reference's own Dataset classes
(dff/dataset.py: Matterport3D incl. AutoAgument's jitter/flips, Middlebury) run on small synthetic files.
    python tests/golden/make_golden_f4_data.py        # build container only
Inputs stored = the decoded arrays exactly as cv.imread returns them; outputs = what Dataset.__getitem__ returns."""
import importlib.util
import os
import sys
import tempfile
import types

import cv2 as cv
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ["skimage", "skimage.morphology"]:
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["skimage.morphology"].disk = sys.modules["skimage.morphology"].closing = lambda *a, **k: None
import scipy.ndimage
sys.modules.setdefault("scipy.ndimage.interpolation", scipy.ndimage)
spec = importlib.util.spec_from_file_location("ref_dataset", "/root/reference/dff/dataset.py")
ds = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ds)

rng = np.random.default_rng(12)
out = {}
with tempfile.TemporaryDirectory() as tmp:
    H, W = 96, 130
    # smooth + noise image so that the antialias filter matters; depth with invalid zeros
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.stack([(127 + 100 * np.sin(xx / 7.0 + c) * np.cos(yy / 5.0) + rng.integers(-20, 20, (H, W))).clip(0, 255) for c in range(3)], -1).astype(np.uint8)
    depth = (2000 + 1500 * np.sin(xx / 23.0) + 900 * (yy > 40) + rng.integers(0, 50, (H, W))).astype(np.uint16)
    depth[rng.random((H, W)) < 0.01] = 0
    os.makedirs(f"{tmp}/rgb/scene0/undistorted_color_images"); os.makedirs(f"{tmp}/dep/scene0/render_depth")
    cv.imwrite(f"{tmp}/rgb/scene0/undistorted_color_images/a.jpg", img, [cv.IMWRITE_JPEG_QUALITY, 95])
    cv.imwrite(f"{tmp}/dep/scene0/render_depth/a.png", depth)
    out["mp_bgr"] = cv.imread(f"{tmp}/rgb/scene0/undistorted_color_images/a.jpg")          # decoded, BGR uint8
    out["mp_depth"] = cv.imread(f"{tmp}/dep/scene0/render_depth/a.png", -1)
    for tag, size in (("a", (48, 65)), ("b", (40, 56)), ("c", (96, 130))):
        d = ds.Matterport3D(f"{tmp}/rgb", f"{tmp}/dep", resize=size, train=False)
        aif, dep = d[0]
        out[f"mp_{tag}_size"], out[f"mp_{tag}_aif"], out[f"mp_{tag}_depth"] = np.asarray(size), aif.numpy(), dep.numpy()
    # train=True: AutoAgument with a seed that draws jitter + both flips and no rotation (the spline rotation is not rebuilt)
    for seed in range(1000):
        np.random.seed(seed)
        r = [np.random.rand() for _ in range(3)]
        if r[0] > 0.5:
            f1, f0, rot = np.random.rand() > 0.5, np.random.rand() > 0.5, np.random.rand() > 0.5
            if f1 and f0 and not rot:
                break
    np.random.seed(seed)
    d = ds.Matterport3D(f"{tmp}/rgb", f"{tmp}/dep", resize=(48, 65), train=True)
    aif, dep = d[0]
    out["mp_aug_seed"], out["mp_aug_jitter"], out["mp_aug_flips"] = seed, np.float32([r[1], r[2]]), np.uint8(3)
    out["mp_aug_aif"], out["mp_aug_depth"] = aif.numpy(), dep.numpy()
    # Middlebury: im0.png + depth.png (mm), depth through cv.resize
    os.makedirs(f"{tmp}/mb/sceneA")
    cv.imwrite(f"{tmp}/mb/sceneA/im0.png", img)
    cv.imwrite(f"{tmp}/mb/sceneA/depth.png", depth)
    out["mb_bgr"], out["mb_depth"] = cv.imread(f"{tmp}/mb/sceneA/im0.png"), cv.imread(f"{tmp}/mb/sceneA/depth.png", -1)
    d = ds.Middlebury(f"{tmp}/mb", resize=(48, 64), train=False)
    aif, dep = d[0]
    out["mb_size"], out["mb_aif"], out["mb_depth_out"] = np.asarray((48, 64)), aif.numpy(), dep.numpy()
np.savez_compressed(os.path.join(HERE, "kat_j_dataset_prep.npz"), **out)
print({k: np.asarray(v).shape for k, v in out.items()})

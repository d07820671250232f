"""CPU oracle for AutoAgument's rotation (dff/dataset.py:275-284).

TEST INFRASTRUCTURE ONLY (see focal_stack_oracle.py): only ``tests/`` may import this file.

The reference calls ``scipy.ndimage.rotate(x, degree, reshape=False)`` with scipy's defaults (order 3, mode
'constant', cval 0, prefilter True).  The arithmetic lives in a third-party dependency that the reference does not pin
(this container: scipy 1.18.1); its published algorithm (scipy/ndimage/src/ni_splines.c, ni_interpolation.c) is
restated here in float64 numpy:

* ``prefilter_axis``   <- spline_filter1d, order 3, mode 'mirror' (what mode 'constant' filters with): gain 6, causal and
  anti-causal recursion with pole z = sqrt(3) - 2 and the mirror initial values
* ``prefilter_fir``    the same filter as the two-sided sum sqrt(3) z^|k| over the mirror-extended signal -- the form the
  CUDA kernel evaluates (csrc/spline_rotate_kernel.cuh)
* ``rotation_xform``   <- scipy.ndimage.rotate: matrix [[cos, sin], [-sin, cos]] (cosdg / sindg), offset = centre - M centre
* ``affine_plane``     <- NI_GeometricTransform: coordinate = shift + m0 row + m1 col; outside [0, n-1] -> cval; 4 x 4
  coefficients with mirrored indices and the order-3 weights of get_spline_interpolation_weights
* ``auto_augment_rotate`` <- the rotation block of AutoAgument: every channel plane of the image, the depth plane, and
  ``depth[depth < 0] = 0``

Parity pin: ``tests/test_oracle_golden.py::test_spline_rotate_oracle_vs_scipy`` compares every function with scipy itself
(<= 1e-13) and ``tests/golden/kat_l_rotate.npz`` holds outputs of the reference's own ``AutoAgument`` / ``Matterport3D``
(tests/golden/make_golden_f4_rotate.py).
"""
import math

import numpy as np

Z = math.sqrt(3.0) - 2.0


def prefilter_axis(x, axis):
    x = np.moveaxis(np.array(x, dtype=np.float64), axis, -1).copy()
    n = x.shape[-1]
    if n == 1:
        return np.moveaxis(x, -1, axis)
    c = x * 6.0
    zn1 = Z ** (n - 1)
    c0 = c[..., 0] + zn1 * c[..., n - 1]
    zi = Z
    for i in range(1, n - 1):
        c0 = c0 + zi * (c[..., i] + zn1 * c[..., n - 1 - i])
        zi *= Z
    c[..., 0] = c0 / (1 - zn1 * zn1)
    for i in range(1, n):
        c[..., i] += Z * c[..., i - 1]
    c[..., n - 1] = (Z * c[..., n - 2] + c[..., n - 1]) * Z / (Z * Z - 1)
    for i in range(n - 2, -1, -1):
        c[..., i] = Z * (c[..., i + 1] - c[..., i])
    return np.moveaxis(c, -1, axis)


def mirror_index(i, n):
    """NI_EXTEND_MIRROR edge mapping of ni_interpolation.c (vectorised)."""
    i = np.asarray(i)
    if n <= 1:
        return np.zeros_like(i)
    s2 = 2 * n - 2
    j = np.abs(i) % s2
    return np.where(j >= n, s2 - j, j)


def prefilter_fir(x, axis, half_width=16):
    x = np.moveaxis(np.array(x, dtype=np.float64), axis, -1)
    n = x.shape[-1]
    idx = np.arange(n)
    out = np.zeros_like(x)
    for k in range(-half_width, half_width + 1):
        out += math.sqrt(3.0) * (Z ** abs(k)) * x[..., mirror_index(idx + k, n)]
    return np.moveaxis(out, -1, axis)


def cosdg_sindg(degree):
    try:
        from scipy import special
        return float(special.cosdg(degree)), float(special.sindg(degree))
    except Exception:                                   # exact at multiples of 90 like cephes' cosdg / sindg
        d = degree % 360
        table = {0: (1.0, 0.0), 90: (0.0, 1.0), 180: (-1.0, 0.0), 270: (0.0, -1.0)}
        return table.get(d, (math.cos(math.radians(degree)), math.sin(math.radians(degree))))


def rotation_xform(degree, H, W):
    """(m00, m01, m10, m11, off0, off1) of scipy.ndimage.rotate(..., reshape=False) for an H x W plane."""
    c, s = cosdg_sindg(degree)
    M = np.array([[c, s], [-s, c]])
    centre = (np.array([H, W], dtype=np.float64) - 1) / 2
    off = centre - M @ centre
    return np.array([M[0, 0], M[0, 1], M[1, 0], M[1, 1], off[0], off[1]])


def _weights3(t):
    u = 1 - t
    w1 = (t * t * (t - 2) * 3 + 4) / 6
    w2 = (u * u * (u - 2) * 3 + 4) / 6
    w0 = u * u * u / 6
    return [w0, w1, w2, 1 - w0 - w1 - w2]


def affine_plane(img, xform, coef=None):
    """scipy.ndimage.affine_transform(img, M, off, order=3, mode='constant', cval=0) for a 2-D array."""
    img = np.asarray(img, dtype=np.float64)
    H, W = img.shape
    if coef is None:
        coef = prefilter_axis(prefilter_axis(img, 0), 1)
    m00, m01, m10, m11, o0, o1 = [float(v) for v in xform]
    r, q = np.mgrid[0:H, 0:W].astype(np.float64)
    y = (o0 + m00 * r) + m01 * q
    x = (o1 + m10 * r) + m11 * q
    inside = (y >= 0) & (y <= H - 1) & (x >= 0) & (x <= W - 1)
    y = np.where(inside, y, 0.0)
    x = np.where(inside, x, 0.0)
    fy, fx = np.floor(y), np.floor(x)
    wy, wx = _weights3(y - fy), _weights3(x - fx)
    sy, sx = fy.astype(np.int64) - 1, fx.astype(np.int64) - 1
    out = np.zeros((H, W))
    for a in range(4):
        ya = mirror_index(sy + a, H)
        for b in range(4):
            out += wy[a] * wx[b] * coef[ya, mirror_index(sx + b, W)]
    return np.where(inside, out, 0.0)


def auto_augment_rotate(img, depth, degree):
    """img [H,W,3] and/or depth [H,W] float arrays -> rotated copies (dff/dataset.py:275-284)."""
    out_img = out_depth = None
    if img is not None:
        H, W = img.shape[:2]
        xf = rotation_xform(degree, H, W)
        out_img = np.stack([affine_plane(img[..., c], xf) for c in range(img.shape[-1])], -1)
    if depth is not None:
        H, W = depth.shape
        out_depth = affine_plane(depth, rotation_xform(degree, H, W))
        out_depth[out_depth < 0] = 0
    return out_img, out_depth

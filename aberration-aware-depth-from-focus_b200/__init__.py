"""aadff-b200: B200-native aberrated focal-stack synthesis.

This directory is laid out as a *site directory*: put it on sys.path (``import aadff_b200`` from
the repository root does that) and the reference's own import lines keep working,

    from deeplens.psfnet import PSFNet, ThinLens
    from deeplens.render_psf import local_psf_render
    from dff.factory import get_lens
    from dff.utils import select_focus_dist

now backed by libaadff.so (csrc/, C ABI in include/aadff.h).
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)

import aadff_native as native                                    # noqa: E402
from deeplens.psfnet import PSFNet, ThinLens, DMIN, DMAX         # noqa: E402
from deeplens.render_psf import local_psf_render                 # noqa: E402
from dff.factory import get_lens                                 # noqa: E402
from dff.utils import select_focus_dist                          # noqa: E402
import sharding                                                  # noqa: E402

__all__ = ["native", "sharding", "PSFNet", "ThinLens", "local_psf_render", "get_lens", "select_focus_dist", "DMIN", "DMAX"]

"""aadff-b200: B200-native aberrated focal-stack synthesis.

Two ways to put it underneath the reference's scripts:

1. **Shadow packages.**  This directory is laid out as a *site directory*: with it in front of the
   reference checkout on sys.path (``import aadff_b200`` from the repository root does that), the
   reference's own import lines keep working,

       from deeplens.utils import set_seed, set_logger      # reference file (extended __path__)
       from deeplens.psfnet import *                         # B200 PSFNet / ThinLens
       from dff import *                                     # get_lens, select_focus_dist: B200;
                                                             # AiFDepthNet, datasets, metrics: reference files

   now backed by libaadff.so (csrc/, C ABI in include/aadff.h).

2. **install().**  If the reference's ``deeplens`` / ``dff`` are already imported (or come first on
   sys.path), ``aadff_b200.install()`` grafts the CUDA-backed methods onto the reference's own classes
   and module attributes in place (aadff_install.py).
"""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
if _here not in _sys.path:
    _sys.path.insert(0, _here)

# The implementation modules have names of their own (aadff_*), so these imports can never pick up a
# reference ``deeplens`` / ``dff`` that happens to be imported already.
import aadff_native as native                                    # noqa: E402
from aadff_lens import PSFNet, ThinLens, DMIN, DMAX              # noqa: E402
from aadff_render import local_psf_render                        # noqa: E402
from aadff_factory import get_lens, get_dataset                  # noqa: E402
from aadff_focus import select_focus_dist                        # noqa: E402
from aadff_install import install, uninstall                     # noqa: E402
from aadff_data import preprocess_rgbd                           # noqa: E402
import sharding                                                  # noqa: E402

__all__ = ["native", "sharding", "PSFNet", "ThinLens", "local_psf_render", "get_lens", "get_dataset",
           "select_focus_dist", "install", "uninstall", "preprocess_rgbd", "DMIN", "DMAX"]

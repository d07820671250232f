"""Shadow of the reference's ``deeplens`` package for the focal-stack synthesis path.

``psfnet``, ``psfnet_arch`` and ``render_psf`` are the B200 implementations (re-exported from aadff_lens /
aadff_arch / aadff_render).  Every other sub-module of the reference (``utils``, ``optics``, ``basics`` ...)
is *not* rebuilt: if a reference checkout is on sys.path, ``__path__`` is extended with its ``deeplens``
directory, so ``from deeplens.utils import set_seed`` and friends keep resolving to the reference's files
(what 2_aber_aware_dff_aif.py:23 needs), while the three hot-path modules resolve here first.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
_AADFF_SHADOW = True

from .psfnet import *          # noqa: E402,F401,F403
from .psfnet_arch import *     # noqa: E402,F401,F403
from .render_psf import *      # noqa: E402,F401,F403

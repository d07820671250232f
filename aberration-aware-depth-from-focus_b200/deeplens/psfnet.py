"""``deeplens.psfnet`` of the shadow package: the B200 PSFNet / ThinLens (aadff_lens.py)."""
import numpy as np  # noqa: F401  (the reference module exports these through its star import, scripts rely on it)
import torch        # noqa: F401

from aadff_lens import DMAX, DMIN, PSFNet, ThinLens            # noqa: F401
from aadff_arch import MLP, initialize_weights                 # noqa: F401
from aadff_render import local_psf_render                      # noqa: F401

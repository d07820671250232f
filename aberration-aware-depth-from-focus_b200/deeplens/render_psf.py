"""``deeplens.render_psf`` of the shadow package (aadff_render.py)."""
from aadff_render import *                                     # noqa: F401,F403

"""``deeplens.psfnet_arch`` of the shadow package (aadff_arch.py)."""
from aadff_arch import MLP, initialize_weights                 # noqa: F401

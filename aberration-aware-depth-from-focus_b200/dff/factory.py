"""``dff.factory`` of the shadow package (aadff_factory.py)."""
from aadff_factory import get_dataset, get_lens                # noqa: F401

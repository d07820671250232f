"""``dff.utils`` of the shadow package (aadff_focus.py)."""
from aadff_focus import select_focus_dist                      # noqa: F401

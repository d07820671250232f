"""Shadow of the reference's ``dff`` package: ``factory`` and ``utils`` are the B200 versions; ``AiFNet``,
``dataset`` and ``metrics`` (downstream consumers / host-side file IO, out of scope) resolve to the reference's own
files through the extended ``__path__`` when a reference checkout is on sys.path -- so ``from dff import *``
(2_aber_aware_dff_aif.py:25) yields get_lens, get_dataset, select_focus_dist, AiFDepthNet, the datasets, the metrics.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
_AADFF_SHADOW = True

from .factory import *         # noqa: E402,F401,F403
from .utils import *           # noqa: E402,F401,F403

for _name in ("AiFNet", "dataset", "metrics"):
    try:
        _mod = __import__(f"{__name__}.{_name}", fromlist=["*"])
    except ModuleNotFoundError as _e:             # no reference checkout on sys.path (or one of ITS dependencies missing)
        if _e.name not in (f"{__name__}.{_name}",):
            import logging as _logging
            _logging.getLogger(__name__).warning("dff.%s of the reference could not be imported: %s", _name, _e)
        continue
    for _k in getattr(_mod, "__all__", [k for k in vars(_mod) if not k.startswith("_")]):
        globals().setdefault(_k, getattr(_mod, _k))

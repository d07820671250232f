"""install(): put the B200 path underneath an already-imported reference, in place.

The shadow packages (deeplens/, dff/ next to this file) only win when this directory precedes the reference
checkout on sys.path.  When the reference's own ``deeplens`` / ``dff`` are what ``import`` finds (they were
imported first, or the reference root comes first on sys.path), ``install()`` patches them instead:

  * ``deeplens.psfnet.PSFNet``: ``render``, ``pred`` and the new ``render_stack`` / ``simulate_focal_stack`` are
    replaced by the CUDA-backed functions of aadff_lens.PSFNet (they read the lens by duck typing:
    ``kernel_size``, ``psfnet.net``, ``d_min``, ``d_max``); construction, ``load_net``, the ray tracer and
    ``analysis()`` stay the reference's (deeplens/psfnet.py:14-76).
  * ``deeplens.psfnet.ThinLens.render`` -> the fused thin-lens kernel.
  * ``local_psf_render`` in ``deeplens.render_psf`` and in every module that star-imported it
    (``deeplens.psfnet``, ``deeplens``) -> the gather kernel.
  * ``select_focus_dist`` in ``dff.utils`` / ``dff`` -> the device-side version.

``uninstall()`` restores the originals.
"""
import importlib
import sys

_saved = []          # (object, attribute name, had_it, old value)


def _set(obj, name, value):
    _saved.append((obj, name, hasattr(obj, name), getattr(obj, name, None)))
    setattr(obj, name, value)


def install(verbose: bool = False) -> bool:
    """Returns True if a reference installation was patched, False if the shadow packages are the ones in use
    (nothing to do) or no ``deeplens`` can be imported at all."""
    if _saved:
        return True
    import aadff_lens
    import aadff_render
    import aadff_focus
    try:
        ref_psfnet = importlib.import_module("deeplens.psfnet")
    except ModuleNotFoundError:
        return False
    if ref_psfnet.PSFNet is aadff_lens.PSFNet:
        return False                                   # the shadow package is active already
    for name in aadff_lens.PSFNET_GRAFT:
        _set(ref_psfnet.PSFNet, name, getattr(aadff_lens.PSFNet, name))
    for name in aadff_lens.THINLENS_GRAFT:
        _set(ref_psfnet.ThinLens, name, getattr(aadff_lens.ThinLens, name))
    for modname in ("deeplens.render_psf", "deeplens.psfnet", "deeplens", "dff.factory"):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "local_psf_render"):
            _set(mod, "local_psf_render", aadff_render.local_psf_render)
    try:
        importlib.import_module("dff.utils")
    except Exception:                                  # dff may need packages this box lacks; its utils is optional
        pass
    for modname in ("dff.utils", "dff"):
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "select_focus_dist"):
            _set(mod, "select_focus_dist", aadff_focus.select_focus_dist)
    if verbose:
        print(f"aadff_b200.install(): patched {len(_saved)} attributes of the imported reference")
    return True


def uninstall() -> None:
    while _saved:
        obj, name, had, old = _saved.pop()
        if had:
            setattr(obj, name, old)
        else:
            delattr(obj, name)

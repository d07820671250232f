"""Import the unmodified reference from baseline/_ref (bench.py's reference arm and the eager-GPU baseline only).

The reference imports matplotlib, lpips and skimage at module level (deeplens/optics.py:12, deeplens/utils.py:5-8);
none of them is touched by PSFNet.render, and none is installed in this image, so empty stand-ins are registered
first.  PSFNet.load_net is not used: its bare torch.load (deeplens/psfnet.py:76) fails for a CUDA-saved checkpoint on
a CPU box, so the state_dict is loaded with map_location and handed to load_state_dict -- same weights, same module.
"""
import os
import sys
import types

REF_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "deeplens", "psfnet.py"))


def import_reference():
    """-> the reference's deeplens.psfnet module (process-wide: call only in a process that does not use the shadow
    packages of aadff_b200)."""
    for name in ["matplotlib", "matplotlib.pyplot", "lpips", "skimage", "skimage.metrics"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sk = sys.modules["skimage.metrics"]
    if not hasattr(sk, "peak_signal_noise_ratio"):
        sk.peak_signal_noise_ratio = lambda *a, **k: 0
        sk.structural_similarity = lambda *a, **k: 0
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import deeplens.psfnet as ref_psfnet
    assert os.path.realpath(ref_psfnet.__file__).startswith(os.path.realpath(REF_ROOT)), ref_psfnet.__file__
    return ref_psfnet


def make_lens(ks, sensor_res, device, state_dict=None):
    """The reference's PSFNet for the rf50mm lens (ray-traced constructor, ~8 s on CPU), weights from `state_dict`
    (keys net.{0,2,..}.weight/bias) or its own seeded initialisation."""
    import torch
    ref_psfnet = import_reference()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)                       # the reference opens ./lenses/... relative to its root
    try:
        torch.manual_seed(0)
        lens = ref_psfnet.PSFNet(filename="./lenses/rf50mm/lens.json", sensor_res=sensor_res, kernel_size=ks,
                                 device=device)
    finally:
        os.chdir(cwd)
    if state_dict is not None:
        lens.psfnet.load_state_dict(state_dict)
    lens.psfnet.to(device)
    return lens


def load_aifnet():
    """dff/AiFNet.py of the reference, loaded by file path (BASELINE config 5's downstream consumer, used as is)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_aifnet", os.path.join(REF_ROOT, "dff", "AiFNet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

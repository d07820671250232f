"""PSFNet fitting with the division of labour of SURVEY 8f / f3: training targets from the REFERENCE's ray tracer
(PSFNet.get_training_data, deeplens/psfnet.py:135-170, unmodified copy in baseline/_ref), optimisation on the device by
the grafted train_psfnet (aadff_b200.install()).  Prints the loss every 10 iterations.
    python tests/gpu_fit_with_reference_raytracer.py [iters]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_import  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
ref_import.import_reference()
sd = torch.load(os.path.join(ROOT, "tests/golden/rf50mm_PSFNet480x640_ks11.pkl"), map_location="cpu")
lens = ref_import.make_lens(11, (480, 640), "cuda", sd)
import aadff_b200  # noqa: E402
assert aadff_b200.install()
os.chdir(ref_import.REF_ROOT)
torch.manual_seed(0)
inp, psf = lens.get_training_data(bs=128, spp=2048)                  # the reference's ray tracer
print("training batch from the reference ray tracer:", tuple(inp.shape), tuple(psf.shape), psf.device, float(psf.sum(-1).mean()))
before = lens.pred(inp.cuda()).reshape(128, -1)
print("MSE of the shipped checkpoint on this batch:", float(((before - psf.cuda()) ** 2).mean()))
t0 = time.time()
hist = lens.train_psfnet(iters=iters, bs=128, lr=1e-5, spp=2048, evaluate_every=10, save=False)
print(f"{iters + 1} iterations in {time.time() - t0:.1f} s (ray tracing included); losses:", [f"{l:.3e}" for _, l in hist])
print("FIT-OK")

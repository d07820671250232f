"""Golden for the body of 0_warm_up.py (BASELINE config c1), produced by the reference on CPU:
the script's own lines 9-22 with the image it names replaced by datasets/usaf1951.png (the Middlebury RGB files are not
shipped with the reference checkout; its depth maps are) and the real Adirondack-perfect/depth.png.
    python tests/golden/make_golden_warmup.py"""
import os
import sys

import cv2 as cv
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, REF      # noqa: E402

PSFNet = import_reference()[0]                       # also chdir()s into the reference
psfnet = PSFNet(filename='./lenses/rf50mm/lens.json', sensor_res=(480, 640), kernel_size=11, device='cpu')
psfnet.psfnet.load_state_dict(torch.load('./ckpt/rf50mm/PSFNet480x640_ks11.pkl', map_location='cpu'))
# 0_warm_up.py:14-17
img_u8 = cv.resize(cv.cvtColor(cv.imread('./datasets/usaf1951.png'), cv.COLOR_BGR2RGB), (640, 480))
img = torch.tensor(img_u8).permute(2, 0, 1).unsqueeze(0).float() / 255
depth_np = cv.resize(cv.imread('./datasets/Middlebury2014/Adirondack-perfect/depth.png', -1) / 1000., (640, 480))
depth = torch.tensor(depth_np).unsqueeze(0).unsqueeze(0).float()
# 0_warm_up.py:20-22
depth = - depth * 1e3
focus_dist = torch.tensor([-2400.])
with torch.no_grad():
    out = psfnet.render(img, depth, focus_dist)
np.savez_compressed(os.path.join(HERE, "kat_k_warmup_c1.npz"), img_u8=img_u8, depth_m=depth_np.astype(np.float32),
                    out_sub=out[..., ::3, ::3].numpy(), out_rows=out[..., [0, 240, 479], :].numpy(),
                    sum=np.float64(out.double().sum()))
print("written", out.shape, float(out.mean()), "invalid depth px:", int((depth_np == 0).sum()))

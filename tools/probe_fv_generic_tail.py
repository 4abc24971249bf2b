"""Accuracy tail of the generic-D FV path (3xTF32 `StatsGenPolicy`): bundled GMMs without the K=256 / D=64 special case, long
images, every image against the fp32 CUDA-core path."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
for name, d_in, T, n in (("OXFORD102_K256_ROOTSIFT", 128, 2000, 1024), ("OXFORD102_K256_VGG16_PCA", 514, 196, 4096), ("OXFORD102_K256_ROOTSIFT_PCA", 128, 2000, 1024)):
    enc = FisherVectorEncoder(feature_extractor=Descriptors(d_in), weights=getattr(GMMWeights, name))
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((n * T, d_in), device="cuda", generator=g).abs_()
    if d_in == 128:
        x = (x / (x.sum(1, keepdim=True) + 1e-7)).sqrt_()
    offs = torch.arange(n + 1, dtype=torch.int64) * T
    a = enc.encode_descriptors(x, offs)
    N.set_path(N.PATH_SIMT)
    u = torch.cat([enc.encode_descriptors(x[i * T:(i + 256) * T], offs[i:i + 257] - offs[i]) for i in range(0, n, 256)])
    N.set_path(N.PATH_AUTO)
    per = (a - u).norm(dim=1) / u.norm(dim=1)
    print(json.dumps({"gmm": name, "T": T, "images": n, "err_max": float(per.max()), "err_median": float(per.median()),
                      "n_gt_1e-4": int((per > 1e-4).sum()), "n_gt_5e-5": int((per > 5e-5).sum())}), flush=True)
    del x, a, u

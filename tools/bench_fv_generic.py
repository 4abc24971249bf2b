"""FV encode throughput on the shapes the tcgen05 FV kernels do not cover yet (CUDA-core path):
VGG16-PCA (514 -> 257, T = 196) and RootSIFT without PCA (D = 128, T = 2000)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
for name, member, d_in, T, n in (("vgg16_pca", GMMWeights.OXFORD102_K256_VGG16_PCA, 514, 196, 2048),
                                 ("rootsift_nopca", GMMWeights.OXFORD102_K256_ROOTSIFT, 128, 2000, 256)):
    enc = FisherVectorEncoder(feature_extractor=Descriptors(d_in), weights=member, output_dtype=np.float32)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((n * T, d_in), device="cuda", generator=g)
    if d_in == 128:
        x = x.abs_(); x = (x / (x.sum(1, keepdim=True) + 1e-7)).sqrt_()
    offs = torch.arange(n + 1, dtype=torch.int64) * T
    for _ in range(2):
        out = enc.encode_descriptors(x, offs)
    torch.cuda.synchronize()
    N.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = enc.encode_descriptors(x, offs, n_streams=1); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"case": name, "images": n, "T": T, "d_in": d_in, "ms": ms, "images_per_s": n / ms * 1e3,
                      "stages_ms": {k: round(v[0], 3) for k, v in N.profile_read().items()}}))
    N.profile_enable(False)

"""Role / phase cycle counters of the fused FV kernel (needs a -DPVS_TIMING build selected with PVS_LIB, and
PVS_TIMING_PRINT=1): one full-size pass, single stream, one library call."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
n_img, T = int(os.environ.get("N_IMG", 2368)), 2000
enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA, output_dtype=np.float32)
x = torch.empty((n_img * T, 128), dtype=torch.float32, device="cuda")
x.normal_(0, 40).abs_().clamp_(0, 255).floor_()
offs = torch.arange(n_img + 1, dtype=torch.int64) * T
for _ in range(2):
    enc.encode_descriptors(x, offs, images_per_call=n_img, n_streams=1)
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
enc.encode_descriptors(x, offs, images_per_call=n_img, n_streams=1)
e1.record(); torch.cuda.synchronize()
print(f"{n_img} images in one call: {e0.elapsed_time(e1):.3f} ms", file=sys.stderr)

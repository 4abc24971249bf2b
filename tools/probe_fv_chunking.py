"""FV C2 batch (8 189 x 2 000 SIFT-like descriptors, resident): images per library call x number of side streams."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
n, T = 8189, 2000
enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA)
gen = torch.Generator(device="cuda").manual_seed(99)
x = torch.empty((n * T, 128), dtype=torch.float32, device="cuda")
for r in range(0, n * T, 1 << 20):
    blk = x[r:r + (1 << 20)]
    blk.normal_(0, 40, generator=gen)
    blk.abs_().clamp_(0, 255).floor_()
offs = torch.arange(n + 1, dtype=torch.int64) * T
out = torch.empty((n, 33024), dtype=torch.float32, device="cuda")
for ipc, ns in ((592, 2), (1184, 2), (2368, 2), (296, 2), (592, 3), (1184, 3), (592, 1), (444, 2), (740, 2)):
    for _ in range(2):
        enc.encode_descriptors(x, offs, out=out, images_per_call=ipc, n_streams=ns)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        enc.encode_descriptors(x, offs, out=out, images_per_call=ipc, n_streams=ns)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(json.dumps({"images_per_call": ipc, "streams": ns, "ms": round(ms, 3), "images_per_s": round(n / ms * 1e3)}), flush=True)

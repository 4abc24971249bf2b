"""Per-stage times of the FV C2 batch (8 189 x 2 000, resident), stage timers on (one stream, events around every stage)."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
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
for _ in range(2):
    enc.encode_descriptors(x, offs, out=out)
torch.cuda.synchronize()
N.profile_enable(True)
reps = 4
for _ in range(reps):
    enc.encode_descriptors(x, offs, out=out, n_streams=1)
torch.cuda.synchronize()
print(json.dumps({"lib": os.environ.get("PVS_LIB", "default"), "stages_ms": {k: round(v[0] / reps, 3) for k, v in N.profile_read().items()}}))

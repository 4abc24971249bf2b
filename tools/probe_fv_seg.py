"""Accuracy / speed of the fused FV kernel as a function of the statistics segment length (PVS_FV_SEG tiles of 128
descriptors folded with fp32 adds): the full C2 batch against the fp32 CUDA-core path (itself 0.2e-5 .. 0.9e-5 of fp64)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
n, T = int(os.environ.get("N_IMG", 8189)), 2000
enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA)
gen = torch.Generator(device="cuda").manual_seed(99)
x = torch.empty((n * T, 128), dtype=torch.float32, device="cuda")
for r in range(0, n * T, 1 << 20):
    blk = x[r:r + (1 << 20)]
    blk.normal_(0, 40, generator=gen)
    blk.abs_().clamp_(0, 255).floor_()
offs = torch.arange(n + 1, dtype=torch.int64) * T
N.set_path(N.PATH_SIMT)
u = torch.empty((n, 33024), dtype=torch.float32, device="cuda")
for i0 in range(0, n, 1024):
    i1 = min(n, i0 + 1024)
    u[i0:i1] = enc.encode_descriptors(x[i0 * T:i1 * T], offs[i0:i1 + 1] - offs[i0])
N.set_path(N.PATH_AUTO)
torch.cuda.synchronize()
def run(tag):
    a = enc.encode_descriptors(x, offs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        a = enc.encode_descriptors(x, offs)
    e1.record(); torch.cuda.synchronize()
    per = (a - u).norm(dim=1) / u.norm(dim=1)
    print(json.dumps({"variant": tag, "ms": round(e0.elapsed_time(e1) / 3, 3), "images_per_s": round(n / (e0.elapsed_time(e1) / 3) * 1e3),
                      "err_max": float(per.max()), "err_median": float(per.median()), "n_gt_1e-4": int((per > 1e-4).sum()),
                      "n_gt_5e-5": int((per > 5e-5).sum()), "worst": per.topk(3).indices.tolist()}), flush=True)
plan = (("0", ("1", "2", "4", "16")), ("2", ("2", "4")), ("1", ("",)))
if os.environ.get("VARIANTS"):                                  # e.g. VARIANTS=2:1,2:2,0:2
    plan = tuple((v.split(":")[0], (v.split(":")[1],)) for v in os.environ["VARIANTS"].split(","))
for mode, segs in plan:
    for seg in segs:
        os.environ["PVS_FV_FUSED"] = mode
        if seg: os.environ["PVS_FV_SEG"] = seg
        else: os.environ.pop("PVS_FV_SEG", None)
        run(f"fused={mode} seg={seg or '-'}")

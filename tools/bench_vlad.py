"""VLAD encode throughput with per-stage device times (BASELINE.json configs[0] / configs[2] shapes).

    python tools/bench_vlad.py --shape c3 [--images 8192] [--reps 3]

c1: RootSIFT-like 128-D, T = 2000 per image;  c3: VGG16-conv-like 514-D, T = 196 per image.
K = 256 random-init centres (the reference checkout ships no K-Means file).  Descriptors are
resident in HBM; timed with CUDA events after warm-up.  Reports images/s, the algorithmic
HBM bytes per image (descriptors in + encoding out) and the fraction of the measured copy
bandwidth that corresponds to.
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import VLADEncoder
from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
from pyvisim_b200.features import Descriptors

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="c3", choices=["c1", "c3"])
ap.add_argument("--images", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--images-per-call", type=int, default=4096)
ap.add_argument("--path", default="auto", choices=["auto", "simt", "tensor"])
a = ap.parse_args()
T, D = (2000, 128) if a.shape == "c1" else (196, 514)
n = a.images or (2048 if a.shape == "c1" else 16384)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((n * T, D), device=dev, generator=g)
if a.shape == "c1":
    x = x.abs_()
    x = (x / (x.sum(1, keepdim=True) + 1e-7)).sqrt_()
centers = x[torch.randperm(n * T, device=dev, generator=g)[:256]].cpu().numpy() \
    + 0.01 * np.random.default_rng(0).standard_normal((256, D)).astype(np.float32)
enc = VLADEncoder(feature_extractor=Descriptors(D), kmeans_model=kmeans_from_centers(centers))
offs = torch.arange(n + 1, dtype=torch.int64) * T
N.set_path({"auto": N.PATH_AUTO, "simt": N.PATH_SIMT, "tensor": N.PATH_TENSOR}[a.path])
for _ in range(2):
    out = enc.encode_descriptors(x, offs, images_per_call=a.images_per_call)
torch.cuda.synchronize()
N.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    out = enc.encode_descriptors(x, offs, images_per_call=a.images_per_call)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
prof = N.profile_read()
N.profile_enable(False)
alg = T * D * 4 + 256 * D * 4
try:
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    hbm = 6650.0
gbs = alg * n / ms / 1e6
print(json.dumps({"shape": a.shape, "images": n, "T": T, "D": D, "ms": ms, "images_per_s": n / ms * 1e3,
                  "alg_bytes_per_image": alg, "path_gbs": gbs, "frac_hbm": gbs / hbm,
                  "stages_ms": {k: round(v[0] / a.reps, 4) for k, v in prof.items()},
                  "checksum": float(out.double().abs().sum().item())}))

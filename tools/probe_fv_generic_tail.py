"""Accuracy tail of the generic-D FV path (3xTF32 `StatsGenPolicy`): bundled GMMs without the K=256 / D=64 special case, long
images, every image against the fp32 CUDA-core path."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pvs_oracle as O                                          # the checker, for the worst images only
MF = os.path.join(ROOT, "python-visual-similarity_b200", "pyvisim_b200", "res", "model_files")
FILES = {"OXFORD102_K256_ROOTSIFT": ("gmm_k256_root_sift_no_pca", None), "OXFORD102_K256_VGG16_PCA": ("gmm_k256_deep_features_vgg16_pca", "pca_k256_deep_features_vgg16_f2"),
         "OXFORD102_K256_ROOTSIFT_PCA": ("gmm_k256_root_sift_pca", "pca_k256_root_sift_f2")}
def rel(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
for name, d_in, T, n in (("OXFORD102_K256_ROOTSIFT", 128, 2000, 1024), ("OXFORD102_K256_VGG16_PCA", 514, 196, 4096), ("OXFORD102_K256_ROOTSIFT_PCA", 128, 2000, 1024)):
    enc = FisherVectorEncoder(feature_extractor=Descriptors(d_in), weights=getattr(GMMWeights, name))
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((n * T, d_in), device="cuda", generator=g).abs_()
    if d_in == 128:
        x = (x / (x.sum(1, keepdim=True) + 1e-7)).sqrt_()
    offs = torch.arange(n + 1, dtype=torch.int64) * T
    a = enc.encode_descriptors(x, offs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a = enc.encode_descriptors(x, offs)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    N.set_path(N.PATH_SIMT)
    u = torch.cat([enc.encode_descriptors(x[i * T:(i + 256) * T], offs[i:i + 257] - offs[i]) for i in range(0, n, 256)])
    N.set_path(N.PATH_AUTO)
    per = (a - u).norm(dim=1) / u.norm(dim=1)
    print(json.dumps({"gmm": name, "T": T, "images": n, "seg": os.environ.get("PVS_FV_SEG", "default"), "ms": round(ms, 3), "err_max": float(per.max()), "err_median": float(per.median()),
                      "n_gt_1e-4": int((per > 1e-4).sum()), "n_gt_5e-5": int((per > 5e-5).sum())}), flush=True)
    worst = per.topk(3).indices.tolist()
    w = dict(np.load(os.path.join(MF, FILES[name][0] + ".npz")))
    pc = dict(np.load(os.path.join(MF, FILES[name][1] + ".npz"))) if FILES[name][1] else None
    ref = O.fv_encode([x[i * T:(i + 1) * T].cpu().numpy() for i in worst], w["weights"], w["means"], w["covariances"], w["precisions_cholesky"],
                      pca=(pc["components"], pc["mean"]) if pc else None)
    print(json.dumps({"gmm": name, "worst": worst, "tensor_vs_fp64": [rel(a[i].cpu().numpy(), ref[j]) for j, i in enumerate(worst)],
                      "cuda_core_vs_fp64": [rel(u[i].cpu().numpy(), ref[j]) for j, i in enumerate(worst)]}), flush=True)
    del x, a, u

"""Short program for `ncu --set full`: one launch of every hot kernel at a representative shape
(FV C2 chunk of 592 images; VLAD C1 / C3 chunks; bf16 similarity + top-100 and the fp32-accurate split variant at d = 32768)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import retrieval
from pyvisim_b200.encoders import FisherVectorEncoder, VLADEncoder, GMMWeights
from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
from pyvisim_b200.features import Descriptors
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
which = set((os.environ.get("PROF", "fv,vlad,sim,sim3")).split(","))
if "fv" in which:
    n, T = 592, 2000
    enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA, output_dtype=np.float32)
    x = torch.empty((n * T, 128), dtype=torch.float32, device=dev).normal_(0, 40, generator=g).abs_().clamp_(0, 255).floor_()
    offs = torch.arange(n + 1, dtype=torch.int64) * T
    for _ in range(2):
        enc.encode_descriptors(x, offs, images_per_call=n, n_streams=1)
    torch.cuda.synchronize()
    del x
if "vlad" in which:
    for n, T, D in ((1024, 2000, 128), (4096, 196, 514)):
        x = torch.randn((n * T, D), device=dev, generator=g).abs_()
        if D == 128:
            x = (x / (x.sum(1, keepdim=True) + 1e-7)).sqrt_()
        cen = x[torch.randperm(n * T, device=dev, generator=g)[:256]].cpu().numpy()
        enc = VLADEncoder(feature_extractor=Descriptors(D), kmeans_model=kmeans_from_centers(cen))
        offs = torch.arange(n + 1, dtype=torch.int64) * T
        for _ in range(2):
            enc.encode_descriptors(x, offs, images_per_call=n)
        torch.cuda.synchronize()
        del x
if "sim" in which or "sim3" in which:
    n, d = 8192, 32768
    v = torch.randn((n, d), device=dev, generator=g)
    if "sim" in which:
        vb = retrieval.l2_normalize(v, "bf16")
        for _ in range(2):
            retrieval.cosine_topk(vb, vb, 100)
        torch.cuda.synchronize()
        del vb
    if "sim3" in which:
        vs = retrieval.l2_normalize(v[:4096], "split")
        for _ in range(2):
            retrieval.cosine_topk(vs, vs, 100)
        torch.cuda.synchronize()
print("prof_cmd done")

"""Quick-start case (BASELINE.json configs[0]): label agreement, parity and latency breakdown."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pvs_oracle as O
from pyvisim_b200.encoders import VLADEncoder
from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
from pyvisim_b200.features import Descriptors
from pyvisim_b200._utils import cosine_similarity
rng = np.random.default_rng(0)
def rootsift_like(t):
    a = np.abs(rng.standard_normal((t, 128))).astype(np.float32); a /= a.sum(axis=1, keepdims=True) + 1e-7; return np.sqrt(a)
d1, d2 = rootsift_like(2000), rootsift_like(1900)
cen = np.vstack([d1, d2])[rng.choice(3900, 256, replace=False)]
enc = VLADEncoder(feature_extractor=Descriptors(128), kmeans_model=kmeans_from_centers(cen))
out, lab = enc.encode_descriptors([d1, d2], return_labels=True)
x = np.vstack([d1, d2]); gold = O.kmeans_predict(x, cen)
bad = np.flatnonzero(lab != gold)
s = O.kmeans_scores(x[bad], cen) if bad.size else None
print("label mismatches:", bad.size)
for i, r in enumerate(bad[:10]):
    gap = abs(s[i, lab[r]] - s[i, gold[r]]); print("  row", r, "ours", lab[r], "oracle", gold[r], "fp64 gap", gap, "rel", gap / np.abs(s[i]).max())
ref = O.vlad_encode([d1, d2], cen)
print("vlad rel-L2 vs oracle:", np.linalg.norm(out - ref) / np.linalg.norm(ref))
print("sim ours", enc.similarity_score([d1], [d2]), "oracle", O.similarity_score(ref[:1], ref[1:]), "oracle on our encodings", O.similarity_score(out[:1], out[1:]))
# fp32 argmin of the oracle's own scores in fp64 vs fp32
s64 = O.kmeans_scores(x, cen); print("oracle fp32 labels vs fp64 argmin mismatches:", int((gold != s64.argmin(1)).sum()), " ours vs fp64:", int((lab != s64.argmin(1)).sum()))
def t(fn, n=50):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
v1, v2 = out[:1], out[1:]
print("ms: encode(1 image)", t(lambda: enc.encode([d1])), " encode_descriptors(1)", t(lambda: enc.encode_descriptors([d1])),
      " cosine(1x1)", t(lambda: cosine_similarity(v1, v2)), " similarity_score", t(lambda: enc.similarity_score([d1], [d2])))

"""rel-L2 error of the FV golden cases on the current path (vs the fp64 reference output)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, split, rel_l2
from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
CASES = [("fv_sift_pca", "OXFORD102_K256_SIFT_PCA"), ("fv_sift_pca_gmmsampled", "OXFORD102_K256_SIFT_PCA"), ("fv_rootsift_pca", "OXFORD102_K256_ROOTSIFT_PCA"),
         ("fv_rootsift_nopca", "OXFORD102_K256_ROOTSIFT"), ("fv_sift_nopca", "OXFORD102_K256_SIFT"), ("fv_vgg_pca", "OXFORD102_K256_VGG16_PCA"),
         ("fv_vgg_pca_gmmsampled", "OXFORD102_K256_VGG16_PCA")]
for case, member in CASES:
    g = load_golden(case)
    d_in = g["desc"].shape[1]
    enc = FisherVectorEncoder(feature_extractor=Descriptors(d_in), weights=getattr(GMMWeights, member))
    res = {}
    for name, path in (("auto", N.PATH_AUTO), ("simt", N.PATH_SIMT)):
        N.set_path(path)
        res[name] = rel_l2(enc.encode_descriptors(g["desc"], g["offsets"]), g["out"])
    N.set_path(N.PATH_AUTO)
    print(f"{case:28s} auto {res['auto']:.3e}   simt {res['simt']:.3e}")

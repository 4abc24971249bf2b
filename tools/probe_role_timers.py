"""Role wait-time counters of the FV kernels on one 592-image chunk: run with PVS_LIB pointing at a library built with\nPVS_NVCC_EXTRA=-DPVS_TIMING (see pyvisim_b200/_build.py)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-visual-similarity_b200"))
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
n, T = 592, 2000
enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA)
g = torch.Generator(device="cuda").manual_seed(99)
x = torch.empty((n * T, 128), device="cuda").normal_(0, 40, generator=g).abs_().clamp_(0, 255).floor_()
offs = torch.arange(n + 1, dtype=torch.int64) * T
enc.encode_descriptors(x, offs); torch.cuda.synchronize()
os.environ["PVS_TIMING_PRINT"] = "1"
enc.encode_descriptors(x, offs); torch.cuda.synchronize()

import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA, output_dtype=np.float32)
for n_img in (1, 64, 512, 2048):
    T = 2000
    x = torch.randn((n_img * T, 128), device="cuda").abs_().mul_(40).floor_()
    offs = torch.arange(n_img + 1, dtype=torch.int64) * T
    for path in (N.PATH_TENSOR, N.PATH_SIMT):
        N.set_path(path)
        for _ in range(2):
            enc.encode_descriptors(x, offs, images_per_call=4096)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = 5
        for _ in range(reps):
            enc.encode_descriptors(x, offs, images_per_call=4096)
        e1.record()
        t_host = (time.perf_counter() - t0) / reps
        torch.cuda.synchronize()
        t_wall = (time.perf_counter() - t0) / reps
        print(f"n_img={n_img:5d} path={path} host-enqueue {t_host*1e3:8.3f} ms  wall {t_wall*1e3:8.3f} ms  gpu {e0.elapsed_time(e1)/reps:8.3f} ms  -> {n_img/t_wall:10.0f} img/s", flush=True)
N.set_path(N.PATH_AUTO)
# stage breakdown for 2048 images in one call
N.profile_enable(True)
enc.encode_descriptors(x, offs, images_per_call=4096)
torch.cuda.synchronize()
print(N.profile_read())

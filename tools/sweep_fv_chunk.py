"""FV C2 throughput as a function of images per library call and number of streams.
Small chunks keep a chunk's posteriors (1 KB per descriptor) inside the 126 MB L2 between the
posterior and the statistics kernel (the workspace is reused chunk after chunk)."""
import os, sys, json, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
from pyvisim_b200.features import Descriptors
n_img, T = int(os.environ.get("N_IMG", 4096)), 2000
dev = torch.device("cuda", 0)
enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA, output_dtype=np.float32)
x = torch.empty((n_img * T, 128), dtype=torch.float32, device=dev)
x.normal_(0, 40).abs_().clamp_(0, 255).floor_()
offs = torch.arange(n_img + 1, dtype=torch.int64) * T
res = torch.empty((n_img, 33024), dtype=torch.float32, device=dev)
ref = None
for ns in (1, 2):
    for ipc in (24, 32, 48, 64, 96, 148, 296, 592, 1184):
        f = lambda: enc.encode_descriptors(x, offs, images_per_call=ipc, out=res, n_streams=ns)
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        chk = float(res.double().sum())
        print(json.dumps({"streams": ns, "images_per_call": ipc, "ms": round(ms, 3), "images_per_s": round(n_img / ms * 1e3), "checksum": chk}), flush=True)

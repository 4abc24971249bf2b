"""CUDA-core GEMM (pvs_pca_project on the round-to-nearest path): 401 408 x 514 descriptors onto N columns, N around the
tile widths (the VGG16 PCA has N = 257 = two 128-wide tiles + one column)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N
M, K = 2048 * 196, 514
x = torch.randn((M, K), device="cuda")
rng = np.random.default_rng(0)
N.set_path(N.PATH_SIMT)
for n in (128, 256, 257, 288, 320):
    pca = N.Model.pca(rng.standard_normal((n, K)).astype(np.float32), rng.standard_normal(K).astype(np.float32))
    y = torch.empty((M, n), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        N.check(N.lib().pvs_pca_project(pca.handle, x.data_ptr(), M, y.data_ptr(), st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        N.check(N.lib().pvs_pca_project(pca.handle, x.data_ptr(), M, y.data_ptr(), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"n": n, "ms": round(ms, 3), "tflops": round(2.0 * M * K * n / ms / 1e9, 2)}), flush=True)

"""All-pairs cosine similarity + top-k throughput (BASELINE.json configs[3] shape, scaled).

    python tools/bench_sim.py --n 32768 --d 32768 --k 100 [--nq 8192]

VLAD-shaped rows: 256 blocks of 128, ~15 % of the blocks empty, every other block unit
norm.  Reports TFLOP/s (2*nq*n*d, no symmetry credit) and queries/s for the fused bf16
tensor-core kernel, timed with CUDA events after warm-up.
"""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import _native as N, retrieval

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32768)
ap.add_argument("--nq", type=int, default=0)
ap.add_argument("--d", type=int, default=32768)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", type=int, default=64, help="rows verified against a torch fp32 matmul of the same bf16 operands")
a = ap.parse_args()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.empty((a.n, a.d), dtype=torch.bfloat16, device=dev)
blk = 128
for r in range(0, a.n, 4096):
    v = torch.randn((min(4096, a.n - r), a.d // blk, blk), device=dev, generator=g)
    v = v / v.norm(dim=2, keepdim=True)
    v = v * (torch.rand((v.shape[0], v.shape[1], 1), device=dev, generator=g) > 0.15)
    v = v.reshape(v.shape[0], -1)
    x[r:r + v.shape[0]] = (v / v.norm(dim=1, keepdim=True).clamp_min(1e-30)).bfloat16()
nq = a.nq or a.n
q = x[:nq]
for _ in range(1):
    s, i = retrieval.cosine_topk(q, x, a.k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N.profile_enable(True)
e0.record()
for _ in range(a.reps):
    s, i = retrieval.cosine_topk(q, x, a.k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
prof = N.profile_read()
N.profile_enable(False)
flops = 2.0 * nq * a.n * a.d
ok = err = recall = None
if a.check:
    # fp64 scores of the same bf16 operands for the first rows: every returned score must be
    # the true score of the returned index (tensor-core fp32 accumulation error only), and
    # every returned index must belong to the true top-k up to that error
    rows = slice(0, min(a.check, nq))
    full = q[rows].double() @ x.double().T
    ts, ti = torch.topk(full, a.k, dim=1)
    got = torch.gather(full, 1, i[rows])
    err = float((got - s[rows].double()).abs().max())
    slack = float((ts[:, -1:] - got).clamp_min(0).max())          # how far below the true k-th best
    recall = float(sum(len(set(r1.tolist()) & set(r2.tolist())) for r1, r2 in zip(ti, i[rows]))) / ti.numel()
    ok = err <= 1e-3 and slack <= 2 * err + 1e-7
print(json.dumps({"n_db": a.n, "n_q": nq, "d": a.d, "k": a.k, "ms": ms, "tflops": flops / ms / 1e9,
                  "queries_per_s": nq / ms * 1e3, "stages_ms": {k: v[0] / a.reps for k, v in prof.items()},
                  "topk_ok": ok, "max_score_err_vs_fp64": err, "recall_vs_fp64_topk": recall}))

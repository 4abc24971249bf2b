"""Row-block sharded all-pairs cosine similarity + top-k across N GPUs (BASELINE.json configs[3]
shape, scaled by --n).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/bench_retrieval_mgpu.py --rows 65536 --dim 32768 --topk 100

Every rank generates the same seeded VLAD-shaped bf16 database (replicated, SURVEY.md 8e
variant i), scores its own block of query rows against it with the fused tensor-core kernel
and the ranks all-gather the final (rows / N, k) score and index lists over NCCL -- the only
collective on the path.  Timing: CUDA events around [top-k kernel + all-gather], barrier on
both sides, max over ranks.  Rank 0 also checks a few gathered rows against a single-GPU
pass over the same rows.
"""
import argparse, json, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
from pyvisim_b200 import retrieval

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=65536)
ap.add_argument("--dim", type=int, default=32768)
ap.add_argument("--topk", type=int, default=100)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--native-comm", action="store_true", help="all-gather through pvs_allgather_topk (library-owned NCCL communicator)")
ap.add_argument("--ring", action="store_true", help="database sharded, shards passed around a ring (retrieval.all_pairs_topk_ring); "
                                                    "checked against the replicated-database path")
ap.add_argument("--shard-only", action="store_true", help="with --ring: every rank generates ONLY its own shard (the database never "
                                                          "exists in one place, e.g. 524,288 x 164,608-D = 172 GB in bf16); checked by properties")
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
lo, hi = retrieval.shard_bounds(a.rows, world, rank)
gen_rows = (hi - lo) if a.shard_only else a.rows
g = torch.Generator(device=dev).manual_seed(rank + 1 if a.shard_only else 0)   # replicated: same data on every rank
x = torch.empty((gen_rows, a.dim), dtype=torch.bfloat16, device=dev)
for r in range(0, gen_rows, 4096):
    m = min(4096, gen_rows - r)
    v = torch.randn((m, a.dim // 128, 128), device=dev, generator=g)
    v = v / v.norm(dim=2, keepdim=True)
    v = v * (torch.rand((m, a.dim // 128, 1), device=dev, generator=g) > 0.15)
    v = v.reshape(m, -1)
    x[r:r + m] = (v / v.norm(dim=1, keepdim=True).clamp_min(1e-30)).bfloat16()
comm = retrieval.NativeComm() if (a.native_comm and world > 1) else None


def step():
    if a.ring:                                                   # this rank only ever touches two shards at a time
        mine = x if a.shard_only else x[lo:hi]
        return retrieval.all_pairs_topk_ring(mine, a.topk, rank=rank, world=world, normalized=True, gather=not a.shard_only)
    s, i = retrieval.cosine_topk(x[lo:hi], x, a.topk)            # this rank's query rows vs the whole database
    return retrieval.gather_topk(s, i, a.rows, comm=comm) if world > 1 else (s, i)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


s, i = step()
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    s, i = step()
e1.record()
barrier()
ms = e0.elapsed_time(e1) / a.reps
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
ok = None
if a.shard_only:
    # properties instead of a reference pass: every vector's best match is itself (global index = shard offset + row,
    # score 1 up to bf16 rounding), lists sorted, all indices inside the database
    own = torch.arange(lo, hi, device=dev)
    good = bool(torch.equal(i[:, 0], own)) and bool(((s[:, 0] - 1).abs() < 2e-2).all()) and bool((s[:, :-1] >= s[:, 1:]).all()) \
        and bool(((i >= 0) & (i < a.rows)).all())
    t = torch.tensor([1 if good else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "n": a.rows, "d": a.dim, "k": a.topk, "ms": ms, "tflops_total": 2.0 * a.rows * a.rows * a.dim / ms / 1e9,
                          "queries_per_s": a.rows / ms * 1e3, "database_bytes_bf16": 2 * a.rows * a.dim, "shard_bytes_bf16": 2 * (hi - lo) * a.dim,
                          "self_match_top1_sorted_in_range_on_every_rank": bool(t.item() == 1),
                          "collective": "ring pass of the database shards (NCCL send/recv); lists stay sharded", "scaling": "database and queries sharded"}))
elif rank == 0:
    rows = torch.tensor([0, a.rows // 2, a.rows - 1], device=dev)
    s1, i1 = retrieval.cosine_topk(x[rows], x, a.topk)
    ok = bool(torch.equal(i1, i[rows]) and torch.equal(s1, s[rows]))
    if a.ring:                                                   # the whole result against the replicated-database path
        s2, i2 = retrieval.cosine_topk(x[lo:hi], x, a.topk)
        ok = ok and bool(torch.equal(i2, i[lo:hi]) and torch.equal(s2, s[lo:hi]))
    print(json.dumps({"n_gpus": world, "n": a.rows, "d": a.dim, "k": a.topk, "ms": ms, "tflops_total": 2.0 * a.rows * a.rows * a.dim / ms / 1e9,
                      "queries_per_s": a.rows / ms * 1e3, "gathered_shape": list(s.shape),
                      "gathered_rows_match_single_gpu_pass": ok, "collective": ("ring pass of the database shards (NCCL send/recv) + " if a.ring else "") + ("pvs_allgather_topk" if comm else "torch.distributed all_gather_into_tensor"), "scaling": "strong (fixed database, query rows split)"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

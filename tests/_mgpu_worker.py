"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL).  SURVEY.md section 4(c): the sharded
retrieval paths must return, bit for bit, what a single-GPU pass returns, and the exact path must agree with the
fp64 oracle on a subset.  Also the device label metrics over the gathered lists and the FV encode of an image shard."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pvs_oracle as O
from pyvisim_b200 import retrieval

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, k = 3001, 1024, 20                                # uneven shards on purpose
g = torch.Generator(device=dev).manual_seed(11)          # same data on every rank
x = torch.randn((n, d), device=dev, generator=g)
x[7] = x[3]                                              # exact ties across shard boundaries
x[n - 1] = x[3]
lo, hi = retrieval.shard_bounds(n, world, rank)
comm = retrieval.NativeComm()


def same(a, b, what):
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), f"rank {rank}: {what} differs from the single-GPU pass"


for dtype in ("bf16", "fp32"):
    single = retrieval.all_pairs_topk(x, k, dtype=dtype)                                         # the whole job on this GPU
    # (i) database replicated, query rows sharded, lists all-gathered by torch.distributed and by the library's communicator
    s, i = retrieval.all_pairs_topk(x, k, dtype=dtype, rank=rank, world=world, gather=False)
    same((s, i), (single[0][lo:hi], single[1][lo:hi]), f"{dtype} replicated shard")
    same(retrieval.gather_topk(s, i, n), single, f"{dtype} gather (torch.distributed)")
    same(retrieval.gather_topk(s, i, n, comm=comm), single, f"{dtype} gather (pvs_allgather_topk)")
    # (ii) database sharded and all-gathered once
    same(retrieval.all_pairs_topk(x[lo:hi].contiguous(), k, dtype=dtype, rank=rank, world=world, database_is_sharded=True), single,
         f"{dtype} sharded database")
    # (iii) database sharded for good, shards travel around the ring.  bf16: bit-identical.  fp32: every shard's partial list
    # has its own near-tie runs re-scored exactly while the rest keeps its tensor score (error below the band), so scores
    # may differ from the single pass in the last bits and the merged order only at near-ties: checked against fp64 below.
    ring = retrieval.all_pairs_topk_ring(x[lo:hi].contiguous(), k, dtype=dtype, rank=rank, world=world, gather=True)
    if dtype == "bf16":
        same(ring, single, f"{dtype} ring")
    if dtype == "fp32" and rank == 0:                    # the exact paths against the fp64 ranking on a subset of the rows
        sub = np.arange(0, n, 37)
        xs = x.cpu().numpy().astype(np.float64)
        xn = xs / np.linalg.norm(xs, axis=1, keepdims=True)
        s64 = xn[sub] @ xn.T
        ref = np.lexsort((np.arange(n)[None, :].repeat(len(sub), 0), -s64), axis=1)[:, :k]
        for what, (sc, ix) in (("single", single), ("ring", ring)):
            got = ix[sub].cpu().numpy()
            for r in range(len(sub)):
                if not np.array_equal(got[r], ref[r]):   # only near-ties below 1e-6 may differ
                    a, b = np.sort(s64[r][got[r]])[::-1], np.sort(s64[r][ref[r]])[::-1]
                    assert np.abs(a - b).max() < 1e-6, f"{what} row {sub[r]}: disagrees with the fp64 ranking"
            assert np.abs(sc[sub].cpu().numpy() - np.take_along_axis(s64, got, 1)).max() <= 1e-5, what
# device label metrics over the gathered lists == the oracle's on the host
labels = torch.randint(0, 17, (n,), generator=torch.Generator().manual_seed(5))
s, i = retrieval.all_pairs_topk(x, k + 1, dtype="fp32", rank=rank, world=world, exclude_self=False)
hits, ap = retrieval.label_metrics(i[:, 1:].contiguous(), labels.to(dev), labels.to(dev))
acc_ref = O.top_k_accuracy_from_lists(i[:, 1:].cpu().numpy(), labels.numpy(), labels.numpy())
map_ref = O.top_k_map_from_lists(i[:, 1:].cpu().numpy(), labels.numpy(), labels.numpy())
assert abs(float((hits > 0).float().mean()) - acc_ref) <= 1e-6, (float((hits > 0).float().mean()), acc_ref)
assert abs(float(ap.double().mean()) - map_ref) <= 1e-6, (float(ap.double().mean()), map_ref)
dist.barrier()
comm.close()
if rank == 0:
    print("MGPU_OK")
dist.destroy_process_group()

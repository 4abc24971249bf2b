"""All-pairs cosine similarity with fused top-k, single GPU and row-block sharded.

Replaces the per-query loop of ``pyvisim/eval.py:31-43,69-80,126-132`` (one
``cosine_similarity`` call that re-normalises the whole database plus a full
``np.argsort`` *per query*) with: normalise the database once, then for every block of
query rows one contraction against the database with the k best columns kept per row.

Multi-GPU: rank r owns query rows ``[r*N/W, (r+1)*N/W)``; the
database is either replicated or resident as W shards that are all-gathered once.  Each
rank's top-k is final for its rows; the only exchange on the result path is an
all-gather of the ``(rows/W, k)`` score and index lists (``torch.distributed``, NCCL on
GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _native as N

__all__ = ["l2_normalize", "cosine_topk", "all_pairs_topk", "shard_bounds", "merge_topk",
           "gather_topk", "label_metrics", "NativeComm"]


class NativeComm:
    """NCCL communicator owned by ``libpvs_b200`` (``pvs_comm_*`` / ``pvs_allgather_topk``).

    Created collectively by all ranks of an initialised ``torch.distributed`` group: rank 0
    makes the 128-byte NCCL unique id, the group broadcasts it, every rank calls
    ``ncclCommInitRank`` on its current CUDA device.  ``gather_topk(..., comm=NativeComm())``
    then runs the all-gather of the top-k lists through the C ABI instead of
    ``torch.distributed``."""

    def __init__(self, group=None):
        import ctypes as C
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("torch.distributed must be initialised to exchange the NCCL id")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        lib = N.lib()
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            N.check(lib.pvs_comm_unique_id(C.addressof(buf)))
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        on_cuda = dist.get_backend(group) == "nccl"
        t = uid.cuda() if on_cuda else uid
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(t.cpu().tolist())
        self._handle = C.c_void_p()
        N.check(lib.pvs_comm_create(raw, self.world, self.rank, C.byref(self._handle)))

    def allgather_topk(self, scores, idx):
        """[rows, k] fp32 / int64 CUDA tensors (same rows on every rank) -> [world*rows, k]."""
        import torch
        rows, k = scores.shape
        s_all = torch.empty((self.world * rows, k), dtype=torch.float32, device=scores.device)
        i_all = torch.empty((self.world * rows, k), dtype=torch.int64, device=scores.device)
        with torch.cuda.device(scores.device):
            N.check(N.lib().pvs_allgather_topk(self._handle, scores.contiguous().data_ptr(), idx.contiguous().data_ptr(),
                                               rows, k, s_all.data_ptr(), i_all.data_ptr(),
                                               torch.cuda.current_stream(scores.device).cuda_stream))
        return s_all, i_all

    def close(self):
        if getattr(self, "_handle", None):
            N.lib().pvs_comm_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced row range of ``rank`` (first ``n % world`` ranks get one more)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def split_supported(d: int) -> bool:
    """Shapes the fp32-accurate tensor-core similarity takes (fp16 hi + lo operand planes)."""
    return d % 8 == 0 and d >= 64


def l2_normalize(x, dtype: str = "fp32"):
    """Row-normalise a CUDA fp32 tensor (zero rows stay zero).

    ``"fp32"`` -> fp32 ``[n, d]``; ``"bf16"`` -> bf16 ``[n, d]`` (one-pass tensor-core similarity, scores within
    1e-2); ``"split"`` -> fp16 ``[2, n, d]``: the planes ``hi``, ``lo`` with ``v * 2**15 = hi + lo`` -- the operand
    format of the fp32-accurate tensor-core similarity (``PVS_F16X2``)."""
    import torch
    x = x.contiguous().float()
    if dtype == "split":
        out = torch.empty((2,) + tuple(x.shape), dtype=torch.float16, device=x.device)
        code = N.F16X2
    else:
        out = torch.empty_like(x, dtype=torch.bfloat16 if dtype == "bf16" else torch.float32)
        code = N.BF16 if dtype == "bf16" else N.F32
    with torch.cuda.device(x.device):
        N.check(N.lib().pvs_l2_normalize_rows(x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(), code,
                                              torch.cuda.current_stream(x.device).cuda_stream))
    return out


def _is_split(t) -> bool:
    import torch
    return t.dtype == torch.float16 and t.ndim == 3 and t.shape[0] == 2


def take_rows(xn, lo: int, hi: int):
    """Rows ``lo:hi`` of a normalised matrix in any of the three formats (contiguous)."""
    return xn[:, lo:hi].contiguous() if _is_split(xn) else xn[lo:hi]


def cosine_topk(queries_n, database_n, k: int, index_offset: int = 0, return_stats: bool = False):
    """Top-k database rows per query row.  Inputs are ALREADY row-normalised CUDA tensors in the same format
    (fp32, bf16 or split planes, see :func:`l2_normalize`).  Returns (scores fp32 [nq,k], indices int64 [nq,k]),
    scores descending, lowest index first on exact ties.  With fp32 / split operands on a tcgen05 device the
    indices are those of the exact (fp64) ranking of the operands; ``return_stats`` adds
    ``{"rescored": ..., "unresolved": ...}`` of the exact tie resolution."""
    import ctypes as C
    import torch
    split = _is_split(queries_n)
    if split != _is_split(database_n) or queries_n.dtype != database_n.dtype or \
            (not split and queries_n.dtype not in (torch.float32, torch.bfloat16)):
        raise ValueError("queries and database must both be float32, both bfloat16 or both split planes")
    if queries_n.shape[-1] != database_n.shape[-1]:
        raise ValueError("feature dimensions differ")
    q, db = queries_n.contiguous(), database_n.contiguous()
    nq, d = q.shape[-2:]
    ndb = db.shape[-2]
    dt = N.F16X2 if split else (N.BF16 if q.dtype == torch.bfloat16 else N.F32)
    scores = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    idx = torch.empty((nq, k), dtype=torch.int64, device=q.device)
    lib = N.lib()
    need = lib.pvs_cosine_topk_workspace_bytes(nq, ndb, d, k, dt)
    ws = torch.empty((max(int(need), 1),), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        st = torch.cuda.current_stream(q.device).cuda_stream
        N.check(lib.pvs_cosine_topk(q.data_ptr(), db.data_ptr(), dt, nq, ndb, d, k, int(index_offset),
                                    scores.data_ptr(), idx.data_ptr(), ws.data_ptr(), ws.numel(), st))
        if return_stats:
            a, b = C.c_int64(0), C.c_int64(0)
            if dt != N.BF16 and nq > 0:
                N.check(lib.pvs_cosine_topk_exact_stats(ws.data_ptr(), nq, ndb, k, C.byref(a), C.byref(b), st))
            return scores, idx, {"rescored": a.value, "unresolved": b.value}
    return scores, idx


def _format(dtype: str, d: int) -> str:
    """Operand format for a requested precision: "fp32" means fp32-ACCURATE scores and exact indices -- split
    planes on the tensor cores when the shape allows, fp32 rows (CUDA cores) otherwise."""
    if dtype in ("fp32", "split"):
        return "split" if split_supported(d) else "fp32"
    if dtype == "fp32_rows":
        return "fp32"
    if dtype != "bf16":
        raise ValueError(f"dtype must be 'bf16', 'fp32' or 'fp32_rows', got {dtype!r}")
    return "bf16"


def merge_topk(scores, idx, k: int):
    """Merge ``parts`` top-k lists per row: inputs [parts, nq, k] -> [nq, k] (database-sharded
    variant; same ordering rule)."""
    import torch
    parts, nq, kk = scores.shape
    assert kk == k and idx.shape == scores.shape
    so = torch.empty((nq, k), dtype=torch.float32, device=scores.device)
    io = torch.empty((nq, k), dtype=torch.int64, device=scores.device)
    with torch.cuda.device(scores.device):
        N.check(N.lib().pvs_topk_merge(scores.contiguous().data_ptr(), idx.contiguous().data_ptr(), parts, nq, k,
                                       so.data_ptr(), io.data_ptr(),
                                       torch.cuda.current_stream(scores.device).cuda_stream))
    return so, io


def gather_topk(scores, idx, n_total: int, group=None, comm: "Optional[NativeComm]" = None):
    """All-gather the per-rank ``(rows_r, k)`` lists into ``(n_total, k)`` on every rank.
    Works on CUDA (NCCL) and CPU (gloo) tensors; row counts may differ by one between ranks.
    With ``comm`` (a :class:`NativeComm`) the exchange goes through ``pvs_allgather_topk``."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return scores, idx
    world = dist.get_world_size(group)
    k = scores.shape[1]
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)

    def padded(t):
        if t.shape[0] == pad:
            return t.contiguous()
        p = torch.zeros((pad, k), dtype=t.dtype, device=t.device)
        p[: t.shape[0]] = t
        return p

    if comm is not None:
        s_all, i_all = comm.allgather_topk(padded(scores), padded(idx))
    else:
        s_all = torch.empty((world * pad, k), dtype=scores.dtype, device=scores.device)
        i_all = torch.empty((world * pad, k), dtype=idx.dtype, device=idx.device)
        dist.all_gather_into_tensor(s_all, padded(scores), group=group)
        dist.all_gather_into_tensor(i_all, padded(idx), group=group)
    if all(hi - lo == pad for lo, hi in sizes):
        return s_all, i_all
    keep = torch.cat([torch.arange(r * pad, r * pad + (hi - lo)) for r, (lo, hi) in enumerate(sizes)]).to(scores.device)
    return s_all[keep], i_all[keep]


def all_pairs_topk(vectors, k: int, *, dtype: str = "bf16", rank: int = 0, world: int = 1,
                   database_is_sharded: bool = False, gather: bool = True, group=None,
                   exclude_self: bool = False):
    """All-pairs cosine similarity + top-k over a set of image vectors.

    ``vectors``: CUDA fp32 tensor.  ``dtype``: "bf16" (fastest, scores within 1e-2, indices near-optimal) or
    "fp32" (fp32-accurate tensor-core scores, indices of the exact ranking; about a third of the bf16 rate).  With ``database_is_sharded=False`` every rank holds the
    full ``(N, D)`` set (replicated database) and scores only its own query rows.  With
    ``database_is_sharded=True`` every rank holds only its ``shard_bounds`` rows; the
    normalised shards are all-gathered once (bf16 halves that traffic) and the rank's own
    rows are the queries.  Returns ``(scores, indices)``: for the rank's rows when
    ``gather=False``, else for all N rows on every rank.
    """
    import torch
    import torch.distributed as dist
    fmt = _format(dtype, vectors.shape[1])
    xn = l2_normalize(vectors, fmt)
    split = fmt == "split"
    rows = xn.shape[-2]
    if world > 1 and database_is_sharded:
        counts = gather_shard_counts(rows, world, group)
        n_total = sum(counts)
        pad = max(counts)
        mine = xn
        if rows != pad:
            z = xn.new_zeros(xn.shape[:-2] + (pad - rows, xn.shape[-1]))
            mine = torch.cat([xn, z], dim=-2)
        full = torch.empty((world,) + tuple(mine.shape), dtype=xn.dtype, device=xn.device)
        dist.all_gather_into_tensor(full, mine.contiguous(), group=group)
        # [world, (2,) pad, d] -> [(2,) world * pad, d]
        full = full.permute(1, 0, 2, 3).reshape(2, world * pad, -1) if split else full.reshape(world * pad, -1)
        if any(c != pad for c in counts):
            keep = torch.cat([torch.arange(r * pad, r * pad + c) for r, c in enumerate(counts)]).to(xn.device)
            full = full[:, keep] if split else full[keep]
        db, q = full.contiguous(), xn
        lo = sum(counts[:rank])
    else:
        n_total = rows
        lo, hi = shard_bounds(n_total, world, rank)
        db, q = xn, take_rows(xn, lo, hi)
    kk = k + 1 if exclude_self else k
    scores, idx = cosine_topk(q, db, kk)
    if exclude_self:
        own = torch.arange(lo, lo + q.shape[-2], device=idx.device).unsqueeze(1)
        keep = idx != own
        # drop the self column where present, else the last column
        order = torch.argsort((~keep).to(torch.int8), dim=1, stable=True)[:, :k]
        scores, idx = torch.gather(scores, 1, order), torch.gather(idx, 1, order)
    if gather and world > 1:
        return gather_topk(scores, idx, n_total, group)
    return scores, idx


def all_pairs_topk_ring(vectors, k: int, *, dtype: str = "bf16", rank: int = 0, world: int = 1, group=None,
                        gather: bool = False, normalized: bool = False):
    """All-pairs cosine top-k with the database SHARDED and never assembled on one GPU
    (variant ii; BASELINE.json configs[4]: 1 M x 164 608-D in bf16 is 329 GB,
    41 GB per rank on 8 GPUs).  Every rank holds its ``shard_bounds`` rows (queries = its own rows).
    The normalised shards travel around a ring: in step s a rank scores its queries against the
    shard of rank ``(rank + s) % world`` with the fused top-k kernel while that shard is already on its
    way to the left neighbour (NCCL send / recv overlapped with the GEMM), and folds the partial list
    into its running top-k with ``pvs_topk_merge``.  ``world`` steps, ``world - 1`` shard transfers per
    rank, two shard buffers.  Same ordering rule as a single pass (score descending, lowest global index
    first), so the result equals ``all_pairs_topk`` on the replicated database.
    Returns the rank's ``(scores, indices)`` or, with ``gather=True``, all rows on every rank."""
    import torch
    import torch.distributed as dist
    xn = vectors if normalized else l2_normalize(vectors, _format(dtype, vectors.shape[1]))
    if world == 1:
        return cosine_topk(xn, xn, k)
    rows = xn.shape[-2]
    counts = gather_shard_counts(rows, world, group)
    starts = np.concatenate([[0], np.cumsum(counts)])
    if min(counts) < k:
        raise ValueError(f"every shard needs at least k = {k} rows (smallest shard: {min(counts)})")
    pad = max(counts)
    cur = xn.new_zeros(xn.shape[:-2] + (pad, xn.shape[-1]))
    cur[..., :rows, :] = xn
    nxt = torch.empty_like(cur)
    best_s = best_i = None
    left, right = (rank - 1) % world, (rank + 1) % world
    for step in range(world):
        src = (rank + step) % world                      # whose shard `cur` holds
        reqs = []
        if step + 1 < world:                             # pass it on while it is being scored
            reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, cur, left, group),
                                           dist.P2POp(dist.irecv, nxt, right, group)])
        s, i = cosine_topk(xn, take_rows(cur, 0, counts[src]) if counts[src] != pad else cur, k, index_offset=int(starts[src]))
        if best_s is None:
            best_s, best_i = s, i
        else:
            best_s, best_i = merge_topk(torch.stack([best_s, s]), torch.stack([best_i, i]), k)
        for r in reqs:
            r.wait()
        cur, nxt = nxt, cur
    if gather:
        return gather_topk(best_s, best_i, int(starts[-1]), group)
    return best_s, best_i


def gather_shard_counts(local_rows: int, world: int, group=None) -> list[int]:
    """Row count of every rank's shard.  A collective: EVERY rank of the group must call it, every
    time (nothing is cached -- a cache keyed on the local count alone lets ranks disagree about
    whether to enter the all-gather when only a peer's shard changed)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(local_rows)], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return [int(o.item()) for o in out]


def label_metrics(topk_idx, db_labels, query_labels):
    """Top-k accuracy hits and per-query average precision (``eval.py:82-98,126-145``,
    quirk Q6) on the device.  Returns (hits int32 [nq], ap fp32 [nq])."""
    import torch
    nq, k = topk_idx.shape
    dev = topk_idx.device
    hits = torch.empty((nq,), dtype=torch.int32, device=dev)
    ap = torch.empty((nq,), dtype=torch.float32, device=dev)
    dbl = db_labels.to(device=dev, dtype=torch.int32).contiguous()
    ql = query_labels.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        N.check(N.lib().pvs_topk_label_metrics(topk_idx.contiguous().data_ptr(), dbl.data_ptr(), ql.data_ptr(), nq, k,
                                               hits.data_ptr(), ap.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream))
    return hits, ap

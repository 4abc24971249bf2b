#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its configuration, B200 path vs the reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4]

Default workload (the one BASELINE.json's metric is quoted on), at N=1 and per rank at N>1 (weak scaling):
configs[1] -- Fisher-vector encode, GMM K=256 (diag) on SIFT-PCA-64 (bundled gmm_k256_sift_pca +
pca_k256_sift_f2), 8,189 synthetic images x 2,000 SIFT-like 128-D fp32 descriptors.  One step = one pass of the
encode path over that batch.  --workload c3: configs[2], VLAD on 100,000 images x 196 x 514-D (sharded over the
ranks, strong scaling); --workload c4: configs[3], all-pairs cosine + top-100 over 262,144 VLAD-32768 vectors
(database replicated in bf16, query rows sharded, top-k lists all-gathered; strong scaling).  At every N the c2
line also carries `extra.retrieval_allgather`: a 65,536-row instance of configs[3] through the same sharded path,
so the one collective of the hot path is exercised by the scaling run.

value  : images/s, whole job, descriptors already resident in HBM, timed with CUDA events
         on the launching stream, max over ranks.  Inputs (8.4 GB) >> L2 (126 MB), so no
         explicit flush is needed between steps.
e2e    : the same through the public API with HOST buffers (pinned): every step copies the
         descriptors H2D and the encodings D2H inside the timed region.
roofline / cpu_baseline: see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))

METRIC = "fv_encode_images_per_s_k256"
UNIT = "images/s"
WORKLOADS = {
    # name: (n_images, T, d_in, description)
    "c2": (8189, 2000, 128, "FV encode, GMM K=256 diag, SIFT-128 -> PCA-64, 8189 images x 2000 descriptors"),
    "c3": (100000, 196, 514, "VLAD encode, K=256, VGG16 conv 514-D, 196 descriptors/image, 100k images"),
    "c4": (262144, 0, 32768, "all-pairs cosine + top-100 over 262144 VLAD vectors (256 x 128 = 32768-D)"),
    "c5": (1048576, 0, 164608, "Pipeline VLAD(RootSIFT-128) + FV(VGG16-PCA 514 -> 257) encode (164608-D) + retrieval top-100 / mAP over 1M synthetic images"),
}
# algorithmic work per image of the FV C2 path (PCA + logits + statistics; descriptors in,
# fp32 encoding out)
C2_FLOP_PER_IMAGE = 2 * 2000 * 128 * 64 + 2 * (2 * 2000 * 256 * 128)       # 294.9 MFLOP
C2_BYTES_PER_IMAGE = 2000 * 128 * 4 + 33024 * 4                              # 1 156 096 B


def host_threads():
    """(threads the BLAS pools will use now, vendor string) via threadpoolctl."""
    try:
        from threadpoolctl import threadpool_info
        info = [i for i in threadpool_info() if i.get("user_api") == "blas"]
        if info:
            return max(i.get("num_threads", 1) for i in info), ", ".join(sorted({f"{i.get('internal_api')} {i.get('version')}" for i in info}))
    except Exception:
        pass
    return os.cpu_count(), "unknown"


class stdout_to_stderr:
    """NCCL prints its version banner on stdout when a communicator is created; the contract is ONE JSON line there."""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


class blas_threads:
    """All host cores for the CPU legs (torchrun exports OMP_NUM_THREADS=1, which would cripple the reference arm), or 1."""
    def __init__(self, n):
        self.n, self.ctx = n, None
    def __enter__(self):
        try:
            from threadpoolctl import threadpool_limits
            self.ctx = threadpool_limits(limits=self.n)
            self.ctx.__enter__()
        except Exception:
            self.ctx = None
    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def sift_like_host(rng, rows, d=128):
    return np.floor(np.clip(np.abs(rng.normal(0, 40, (rows, d))), 0, 255)).astype(np.float32)


# --------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# --------------------------------------------------------------------------------------
def oracle_fv_rate(descs, seconds_hint=None):
    """images/s of the NumPy/fp64 restatement of FisherVectorEncoder.encode on this host."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pvs_oracle as O
    wdir = os.path.join(ROOT, "python-visual-similarity_b200", "pyvisim_b200", "res", "model_files")
    g = np.load(os.path.join(wdir, "gmm_k256_sift_pca.npz"))
    p = np.load(os.path.join(wdir, "pca_k256_sift_f2.npz"))
    t0 = time.perf_counter()
    out = O.fv_encode(descs, g["weights"], g["means"], g["covariances"], g["precisions_cholesky"],
                      pca=(p["components"], p["mean"]))
    dt = time.perf_counter() - t0
    return len(descs) / dt, dt, out


def vlad_like_rows(rng, n, d=32768):
    """VLAD-shaped unit rows: 256 unit blocks of 128, ~15 % of the blocks empty."""
    b = rng.standard_normal((n, d // 128, 128)).astype(np.float32)
    b /= np.linalg.norm(b, axis=2, keepdims=True)
    b *= (rng.random((n, d // 128, 1)) > 0.15)
    b = b.reshape(n, d)
    return b / np.maximum(np.linalg.norm(b, axis=1, keepdims=True), 1e-30)


def cpu_legs(sample_fv, sample_vlad=16, sim_rows=(256, 8192)):
    """The reference's CPU path (oracle port) for the three BASELINE.json metrics on bounded samples: all host threads
    and one thread (BASELINE.md section 3 asks for both and for the BLAS vendor)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pvs_oracle as O
    rng = np.random.default_rng(0)
    T = 2000
    descs = [sift_like_host(rng, T, 128) for _ in range(sample_fv)]
    a = np.abs(rng.standard_normal((sample_vlad * T, 128))).astype(np.float32)
    a = np.sqrt(a / (a.sum(1, keepdims=True) + 1e-7))
    cen = a[rng.choice(a.shape[0], 256, replace=False)]
    vl = [a[i * T:(i + 1) * T] for i in range(sample_vlad)]
    q, db = vlad_like_rows(rng, sim_rows[0]), vlad_like_rows(rng, sim_rows[1])
    out = {}
    for tag, n in (("all_threads", os.cpu_count()), ("one_thread", 1)):
        with blas_threads(n):
            used, vendor = host_threads()
            fv_descs = descs if n > 1 else descs[:max(2, sample_fv // 8)]
            oracle_fv_rate(fv_descs[:2]); O.vlad_encode(vl[:2], cen); O.cosine_topk(q[:8], db, 100)     # warm-up (file loads, pools)
            fv_rate = max(oracle_fv_rate(fv_descs)[0] for _ in range(2))
            t_v = t_s = 1e30
            for _ in range(2):
                t0 = time.perf_counter(); O.vlad_encode(vl, cen); t_v = min(t_v, time.perf_counter() - t0)
                t0 = time.perf_counter(); O.cosine_topk(q, db, 100); t_s = min(t_s, time.perf_counter() - t0)
        out[tag] = {"threads": used, "fv_c2_images_per_s": fv_rate, "vlad_c1_images_per_s": sample_vlad / t_v,
                    "cosine_top100_queries_per_s_at_8192_rows": sim_rows[0] / t_s,
                    "cosine_top100_tflops": 2.0 * sim_rows[0] * sim_rows[1] * 32768 / t_s / 1e12}
        out["blas"] = vendor
    return out


def run_reference(args):
    """The reference's own CPU implementation of the path (the oracle port: the reference is Python on scikit-learn /
    NumPy and cannot travel to the GPU box) on all host threads, same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pvs_oracle as O
    n_img, T, d_in, desc = WORKLOADS[args.workload]
    rng = np.random.default_rng(0)
    with blas_threads(os.cpu_count()):                    # torchrun exports OMP_NUM_THREADS=1
        used, vendor = host_threads()
        if args.workload == "c2":
            sample, metric, unit = 32, METRIC, UNIT
            descs = [sift_like_host(rng, T, d_in) for _ in range(sample)]
            fn = lambda: oracle_fv_rate(descs)
            what = f"{sample} of {n_img} images x {T} descriptors per step"
        elif args.workload == "c3":
            sample, metric, unit = 512, "vlad_encode_images_per_s_k256", UNIT
            x = rng.standard_normal((sample * T, d_in)).astype(np.float32)
            cen = x[rng.choice(x.shape[0], 256, replace=False)]
            descs = [x[i * T:(i + 1) * T] for i in range(sample)]
            fn = lambda: O.vlad_encode(descs, cen)
            what = f"{sample} of {n_img} images x {T} descriptors per step"
        elif args.workload == "c5":
            sample, metric, unit, db_rows = 32, "pipeline_retrieval_queries_per_s", "queries/s", 2048
            q = rng.standard_normal((sample, d_in)).astype(np.float32)
            db = rng.standard_normal((db_rows, d_in)).astype(np.float32)
            fn = lambda: O.cosine_topk(q, db, 100)
            what = f"{sample} queries against {db_rows} of {n_img} database rows of {d_in}-D per step (scaled by {db_rows} / {n_img}); retrieval only"
        else:
            sample, metric, unit, db_rows = 128, "cosine_top100_queries_per_s", "queries/s", 16384
            q, db = vlad_like_rows(rng, sample), vlad_like_rows(rng, 16384)
            fn = lambda: O.cosine_topk(q, db, 100)
            what = f"{sample} queries against 16384 of {n_img} database rows per step (scaled by 16384 / {n_img})"
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            fn()
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    if args.workload in ("c4", "c5"):
        value *= db_rows / n_img                           # queries/s against the full database
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if args.workload == "c2" else "strong", "vs_baseline": None, "dtype": "f64" if args.workload == "c2" else "f32",
        "data": "synthetic", "config": {"workload": desc, "sample": what},
        "cpu_baseline": {"value": value, "unit": unit, "cores": used, "kind": "port", "blas": vendor, "sample": what},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# clocks sampler (pynvml), runs during the timed region
# --------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.ok, self._stop = [], set(), False, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.max = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.ok:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max, "reasons": sorted(self.reasons)}



# --------------------------------------------------------------------------------------
# other BASELINE.json metrics, measured in the same run (N=1 only): VLAD images/s and
# all-pairs similarity TFLOP/s + top-k queries/s at single-GPU sizes
# --------------------------------------------------------------------------------------
def run_extras(dev, pk):
    import torch
    from pyvisim_b200 import _native as N, retrieval
    from pyvisim_b200.encoders import VLADEncoder
    from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
    from pyvisim_b200.features import Descriptors

    def timed(fn, reps):
        # the GPU may have idled (CPU baseline just ran): warm up for ~0.3 s so the clocks are
        # back up before anything is timed
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.3:
            fn()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out

    out = {}
    g = torch.Generator(device=dev).manual_seed(7)
    for name, (n, T, D) in {"vlad_c1_rootsift128": (4096, 2000, 128), "vlad_c3_vgg514": (16384, 196, 514)}.items():
        x = torch.randn((n * T, D), device=dev, generator=g)
        if D == 128:                                      # RootSIFT-like: non-negative, unit L2
            x = x.abs_()
            x = (x / (x.sum(1, keepdim=True) + 1e-7)).sqrt_()
        centers = x[torch.randperm(n * T, device=dev, generator=g)[:256]].cpu().numpy()
        enc = VLADEncoder(feature_extractor=Descriptors(D), kmeans_model=kmeans_from_centers(centers))
        offs = torch.arange(n + 1, dtype=torch.int64) * T
        res = torch.empty((n, 256 * D), dtype=torch.float32, device=dev)     # caller-owned output: no allocation in the timed calls
        ms, _ = timed(lambda: enc.encode_descriptors(x, offs, images_per_call=4096, out=res), 3)
        if os.environ.get("PVS_BENCH_DEBUG"):
            N.profile_enable(True)
            ms2, _ = timed(lambda: enc.encode_descriptors(x, offs, images_per_call=4096, out=res), 3)
            print(name, "ms", ms, "again", ms2, {k: v[0] / v[1] for k, v in N.profile_read().items()}, file=sys.stderr)
            N.profile_enable(False)
        alg = T * D * 4 + 256 * D * 4
        out[name] = {"images_per_s": n / ms * 1e3, "images": n, "descriptors_per_image": T, "d": D, "k": 256,
                     "frac_of_hbm_peak": alg * n / ms / 1e6 / pk["hbm_gbs"], "weights": "random-init K-Means (no file bundled)"}
        del x, enc, res
    # BASELINE.json configs[0]: README quick start, similarity_score of two images (~2k RootSIFT-128
    # descriptors each) through the drop-in API with host arrays; wall-clock latency per call
    rng = np.random.default_rng(0)

    def rootsift_like(t):
        a = np.abs(rng.standard_normal((t, 128))).astype(np.float32)
        a /= a.sum(axis=1, keepdims=True) + 1e-7
        return np.sqrt(a)

    d1, d2 = rootsift_like(2000), rootsift_like(1900)
    cen = np.vstack([d1, d2])[rng.choice(3900, 256, replace=False)]
    enc = VLADEncoder(feature_extractor=Descriptors(128), kmeans_model=kmeans_from_centers(cen))
    for _ in range(5):
        score = enc.similarity_score([d1], [d2])
    lat = []
    for _ in range(30):
        t0 = time.perf_counter()
        score = enc.similarity_score([d1], [d2])
        lat.append(time.perf_counter() - t0)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pvs_oracle as O
    t0 = time.perf_counter()
    ref = O.similarity_score(O.vlad_encode([d1], cen), O.vlad_encode([d2], cen))
    cpu_s = time.perf_counter() - t0
    _, lab = enc.encode_descriptors([d1, d2], return_labels=True)
    lab = np.asarray(lab)
    xs = np.vstack([d1, d2])
    gold = O.kmeans_predict(xs, cen)                      # fp32, this host's BLAS: sub-1e-6 ties fall by its rounding noise
    s64 = (cen.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * (xs.astype(np.float64) @ cen.astype(np.float64).T)
    exact = s64.argmin(1).astype(np.int32)                # what the device resolves near-ties to (fp64 re-evaluation)
    bad = np.flatnonzero(lab != gold)
    gaps = (np.abs(s64[bad, lab[bad]] - s64[bad, gold[bad]]) / np.abs(s64[bad]).max(axis=1)).tolist()

    def vlad_from_labels(x, lb):                          # vlad.py:98-111 with given labels (oracle arithmetic)
        v = np.zeros((256, 128), dtype=np.float32)
        np.add.at(v, lb, x - cen[lb])
        return (v / (np.linalg.norm(v, axis=1, ord=2, keepdims=True) + 1e-9)).flatten()
    ref_exact = float(np.asarray(O.similarity_score(vlad_from_labels(d1, exact[:2000])[None], vlad_from_labels(d2, exact[2000:])[None])).ravel()[0])
    sc = float(np.asarray(score).ravel()[0])
    out["quickstart_vlad_similarity_score"] = {"latency_ms_median": 1e3 * float(np.median(lat)), "latency_ms_min": 1e3 * float(np.min(lat)),
                                               "cpu_oracle_ms": 1e3 * cpu_s, "score": sc,
                                               "rel_diff_vs_oracle_with_exact_labels": abs(sc - ref_exact) / abs(ref_exact),
                                               "label_mismatches_vs_exact_argmin": int((lab != exact).sum()),
                                               "rel_diff_vs_fp32_oracle": float(abs(sc - np.asarray(ref).ravel()[0]) / abs(np.asarray(ref).ravel()[0])),
                                               "label_mismatches_vs_fp32_oracle": int(bad.size), "rel_score_gap_of_those": gaps,
                                               "note": "near-ties below 1e-6 are resolved by exact (fp64) scores on the device; the fp32 reference "
                                                       "resolves them by the rounding noise of its BLAS kernel",
                                               "descriptors": [2000, 1900]}
    del enc
    return out

# --------------------------------------------------------------------------------------
# BASELINE.json configs[3]: all-pairs cosine + top-k, database replicated in bf16, query rows sharded over the ranks,
# top-k lists all-gathered (pvs_allgather_topk through the library's NCCL communicator).  Used by --workload c4 and, with
# a smaller database, as the retrieval leg of every c2 line.
# --------------------------------------------------------------------------------------
def vlad_like_device(n, d, dev, gen, dtype):
    import torch
    v = torch.empty((n, d), dtype=dtype, device=dev)
    for r in range(0, n, 4096):                           # VLAD-shaped rows: 256 unit blocks of 128, ~15 % empty
        m = min(4096, n - r)
        b = torch.randn((m, d // 128, 128), device=dev, generator=gen)
        b = b / b.norm(dim=2, keepdim=True)
        b = b * (torch.rand((m, d // 128, 1), device=dev, generator=gen) > 0.15)
        b = b.reshape(m, -1)
        v[r:r + m] = (b / b.norm(dim=1, keepdim=True).clamp_min(1e-30)).to(dtype)
    return v


def run_retrieval(dev, rank, world, n, d, k, steps, warmup, pk, comm=None):
    import torch
    import torch.distributed as dist
    from pyvisim_b200 import _native as N, retrieval
    gen = torch.Generator(device=dev).manual_seed(4321)   # same seed on every rank: the database is replicated
    db = vlad_like_device(n, d, dev, gen, torch.bfloat16)
    lo, hi = retrieval.shard_bounds(n, world, rank)
    q = db[lo:hi]

    def step():
        s_, i_ = retrieval.cosine_topk(q, db, k)
        if world > 1:
            s_, i_ = retrieval.gather_topk(s_, i_, n, comm=comm)
        return s_, i_

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(1, warmup)):
        s_, i_ = step()
    barrier()
    N.lib().pvs_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s_, i_ = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    launches = int(N.lib().pvs_launch_count())
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # every row's best match is itself (index = global row), lists sorted, gathered lists complete on every rank
    ok = bool((i_[:, 0] == torch.arange(i_.shape[0], device=dev)).all()) and bool((s_[:, :-1] >= s_[:, 1:]).all()) \
        and i_.shape[0] == n
    tf = 2.0 * n * n * d / ms / 1e9
    out = {"ms": ms, "tflops": tf, "queries_per_s": n / ms * 1e3, "n": n, "d": d, "k": k, "dtype": "bf16", "n_gpus": world,
           "frac_of_bf16_sustained_peak": tf / world / pk["bf16_tflops_sustained"], "frac_of_bf16_burst_peak": tf / world / pk["bf16_tflops"],
           "collective": "pvs_allgather_topk (ncclAllGather x 2, library communicator)" if world > 1 else "none (one rank)",
           "allgather_bytes_per_rank": int((hi - lo) * k * 12) if world > 1 else 0,
           "self_match_and_order_ok": ok, "gpu_launches_per_step": launches // max(steps, 1)}
    del db
    return out, (s_, i_)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pyvisim_b200 import _native as N
    from pyvisim_b200.encoders import FisherVectorEncoder, GMMWeights
    from pyvisim_b200.features import Descriptors

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()

    n_img, T, d_in, desc = WORKLOADS["c2"]
    if args.images:
        n_img = args.images
    rows = n_img * T
    enc = FisherVectorEncoder(feature_extractor=Descriptors(d_in), weights=GMMWeights.OXFORD102_K256_SIFT_PCA,
                              output_dtype=np.float32)
    K, D = enc.clustering_model.means_.shape
    out_dim = 2 * K * D + K

    # synthetic SIFT-like descriptors, generated on the device (seeded per rank)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.empty((rows, d_in), dtype=torch.float32, device=dev)
    step_rows = 1 << 20
    for r in range(0, rows, step_rows):
        blk = x[r:r + step_rows]
        blk.normal_(0, 40, generator=gen)
        blk.abs_().clamp_(0, 255).floor_()
    offsets = torch.arange(n_img + 1, dtype=torch.int64) * T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = torch.empty((n_img, out_dim), dtype=torch.float32, device=dev)      # caller-owned output

    def step(n_streams=2):
        return enc.encode_descriptors(x, offsets, images_per_call=args.images_per_call, out=res, n_streams=n_streams)

    for _ in range(args.warmup):
        out = step()
    barrier()
    N.lib().pvs_launch_count_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            out = step()
        ev1.record()
        barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = int(N.lib().pvs_launch_count())
    # per-kernel durations: the timed steps alternate image chunks between two streams, so kernels
    # of neighbouring chunks overlap there; the stage times (and the roofline entry computed from
    # them) come from extra steps issued on ONE stream, CUDA events around every launch
    stages = {}
    prof_steps = 0
    if not args.no_profile:
        prof_steps = max(1, min(args.steps, 2))
        step(n_streams=1)
        barrier()
        N.profile_enable(True)
        for _ in range(prof_steps):
            step(n_streams=1)
        barrier()
        stages = N.profile_read()
        N.profile_enable(False)
    if world > 1:
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_img / (ms_per_step / 1e3)
    checksum = float(out.double().sum().item())

    # ---- e2e: host (pinned) in, host (pinned) out, through encode_descriptors ------------
    e2e_images = n_img if args.e2e_images <= 0 else min(n_img, args.e2e_images)
    e_rows = e2e_images * T
    x_host = torch.empty((e_rows, d_in), dtype=torch.float32, pin_memory=True)
    x_host.copy_(x[:e_rows])
    out_host = torch.empty((e2e_images, out_dim), dtype=torch.float32, pin_memory=True)
    x_np, out_np = x_host.numpy(), out_host.numpy()
    offs_np = (np.arange(e2e_images + 1, dtype=np.int64) * T)
    e2e_steps = max(1, min(args.steps, 3))
    enc.encode_descriptors(x_np, offs_np, out=out_np)                      # warm-up (arena allocation)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        enc.encode_descriptors(x_np, offs_np, out=out_np)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * e2e_images / e2e_s
    e2e_ok = bool(np.allclose(out_np[:4], out[:4].cpu().numpy(), atol=1e-6))
    # declared option, not the headline: the synthetic SIFT-like descriptors are integers in 0..255 (like OpenCV's), so they
    # may also be handed over as uint8 rows -- a quarter of the PCIe bytes, widened on the device, bit-identical encodings
    e2e_u8 = None
    if rank == 0 and world == 1 and not args.no_extra:
        xu8 = torch.empty((e_rows, d_in), dtype=torch.uint8, pin_memory=True)
        xu8.copy_(x[:e_rows].to(torch.uint8))
        out_u8 = torch.empty((e2e_images, out_dim), dtype=torch.float32, pin_memory=True)
        enc.encode_descriptors(xu8.numpy(), offs_np, out=out_u8.numpy())
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            enc.encode_descriptors(xu8.numpy(), offs_np, out=out_u8.numpy())
        torch.cuda.synchronize()
        u8_s = (time.perf_counter() - t0) / e2e_steps
        e2e_u8 = {"value": e2e_images / u8_s, "unit": UNIT, "h2d_bytes_per_step": int(e_rows * d_in + (e2e_images + 1) * 8),
                  "d2h_bytes_per_step": int(e2e_images * out_dim * 4), "bit_identical_to_float32_input": bool(np.array_equal(out_u8.numpy(), out_np)),
                  "note": "uint8 transport of integer-valued descriptors (pvs_fv_encode_host_u8); not the reference's dtype"}
        del xu8, out_u8

    # ---- roofline of the dominant kernel (live CUDA-event stage times) -------------------
    # per image the path does 294.9 MFLOP (PCA + logits + statistics) and moves 1 156 096 B
    # (descriptors in, fp32 encoding out).  `achieved` = that work for the images of one launch / the dominant kernel's
    # average launch time; both roofs are given, `bound` names the binding one.  The contractions run as three kind::f16
    # MMAs per product (fp16 hi + lo operands, fp32-class accuracy), so the tensor ceiling of this path is peak / 3.
    pk = peaks()
    dominant = max(stages.items(), key=lambda kv: kv[1][0]) if stages else (None, (0.0, 0))
    bytes_per_image = {"fv_finalize": K * 2 * D * 4 + 16 * K * 4 + out_dim * 4,  # S + zeroth-order partials in, encoding out
                       "tc_fv_project": T * (d_in + D) * 4,                    # X in, Y out
                       "tc_fv_posterior": T * (D + K) * 4,                     # Y in, Q out (fp16 hi + lo planes = 4 B)
                       "tc_fv_stats": T * (D + K) * 4 + K * 2 * D * 4,         # Q + Y in, S out
                       "tc_fv_poststats_fused": T * D * 4 + K * 2 * D * 4}     # Y in, S out (Q never leaves the SM)
    traffic_db = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic_db = json.load(f)
    except Exception:
        pass

    def two_roofs(ms_for_images, images):
        tf = C2_FLOP_PER_IMAGE * images / (ms_for_images / 1e3) / 1e12
        gb = C2_BYTES_PER_IMAGE * images / (ms_for_images / 1e3) / 1e9
        t_frac, h_frac = tf / pk["bf16_tflops_sustained"], gb / pk["hbm_gbs"]
        return {"tensor": {"achieved_tflops": tf, "peak_tflops": pk["bf16_tflops_sustained"], "frac": t_frac,
                           "frac_of_3pass_ceiling": 3 * t_frac, "flop_per_image": C2_FLOP_PER_IMAGE},
                "hbm": {"achieved_gbs": gb, "peak_gbs": pk["hbm_gbs"], "frac": h_frac, "bytes_per_image": C2_BYTES_PER_IMAGE},
                "binding": "tensor" if 3 * t_frac >= h_frac else "hbm"}

    roofline = None
    if dominant[0]:
        name, (ms, n) = dominant
        per_launch_ms = ms / n
        imgs_per_launch = n_img * prof_steps / n
        r = two_roofs(per_launch_ms, imgs_per_launch)
        tr = traffic_db.get(name)
        if r["binding"] == "tensor":
            roofline = {"bound": "tensor", "achieved": r["tensor"]["achieved_tflops"], "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                        "frac": r["tensor"]["frac"], "peak_source": f"bf16 dense sustained, {pk['source']} (kernel runs inside a long step)"}
        else:
            roofline = {"bound": "hbm", "achieved": r["hbm"]["achieved_gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": r["hbm"]["frac"], "peak_source": f"hbm copy, {pk['source']}"}
        own_flop = {"tc_fv_project": 2 * T * d_in * D, "tc_fv_posterior": 2 * T * K * 2 * D, "tc_fv_stats": 2 * T * K * 2 * D,
                    "tc_fv_poststats_fused": 4 * T * K * 2 * D}.get(name)
        if own_flop:                                       # the kernel's OWN share of the work, for orientation
            own_tf = own_flop * imgs_per_launch / (per_launch_ms / 1e3) / 1e12
            roofline["kernel_own"] = {"flop_per_image": own_flop, "achieved_tflops": own_tf,
                                      "frac_of_3pass_ceiling": 3 * own_tf / pk["bf16_tflops_sustained"],
                                      "hbm_frac_on_interface_bytes": bytes_per_image.get(name, 0) * imgs_per_launch / (per_launch_ms / 1e3) / 1e9 / pk["hbm_gbs"]}
        roofline.update({
            "kernel": name, "tensor": r["tensor"], "hbm": r["hbm"],
            "note": "SURVEY.md 8(d) work of the WHOLE path per image over the dominant kernel's time (so it can exceed the ceiling "
                    "when the step is spread over several kernels: path_roofline is the whole step, kernel_own the kernel's own "
                    "share); the contractions are three kind::f16 MMAs per product, so the tensor ceiling is peak / 3",
            "traffic": tr["dram_bytes_per_image"] * imgs_per_launch if tr else None,
            "traffic_source": tr["source"] if tr else None,
            "kernel_interface_bytes_per_launch": bytes_per_image.get(name, 0) * imgs_per_launch,
            "kernel_ms_per_launch": per_launch_ms, "images_per_launch": imgs_per_launch,
            "kernel_share_of_step": ms / sum(v[0] for v in stages.values()),
            "timing": f"{prof_steps} extra single-stream step(s), CUDA events around each launch"})
    path = two_roofs(ms_per_step, n_img)                   # the same two fractions for the whole step
    alg_bytes = C2_BYTES_PER_IMAGE
    path_gbs = path["hbm"]["achieved_gbs"]

    # ---- CPU baseline: oracle port on a bounded sample (rank 0, N=1 only) ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample
        descs = [x[i * T:(i + 1) * T].cpu().numpy() for i in range(sample)]
        with blas_threads(os.cpu_count()):
            used, vendor = host_threads()
            rate, dt, ref_out = oracle_fv_rate(descs)
        err = float(np.linalg.norm(out[:sample].cpu().numpy().astype(np.float64) - ref_out) / np.linalg.norm(ref_out))
        cpu = {"value": rate, "unit": UNIT, "cores": used, "kind": "port", "blas": vendor,
               "sample": f"first {sample} of {n_img} images, {dt:.1f} s, NumPy/{vendor}, {used} threads",
               "parity_rel_l2_vs_gpu": err, "legs": cpu_legs(32)}

    extra = None
    del x
    torch.cuda.empty_cache()
    if not args.no_extra:
        comm = None
        if world > 1:
            from pyvisim_b200 import retrieval
            with stdout_to_stderr():
                comm = retrieval.NativeComm()
                torch.cuda.synchronize()
        ret, _ = run_retrieval(dev, rank, world, 65536, 32768, 100, 2, 1, pk, comm)
        if rank == 0:
            extra = run_extras(dev, pk) if world == 1 else {}
            extra["retrieval_allgather"] = ret
            if e2e_u8:
                extra["e2e_uint8_descriptors"] = e2e_u8

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "images_per_gpu": n_img, "descriptors_per_image": T, "d_in": d_in,
                       "k": K, "d": D, "weights": "bundled gmm_k256_sift_pca + pca_k256_sift_f2",
                       "l2": "inputs (8.4 GB/GPU) larger than L2, no flush", "parallelism": f"dp{world} (images sharded, no collective)", "streams": "image chunks alternate between 2 CUDA streams per GPU",
                       "images_per_call": args.images_per_call or "4 per SM (592)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e_rows * d_in * 4 + (e2e_images + 1) * 8),
                    "d2h_bytes_per_step": int(e2e_images * out_dim * 4), "images": e2e_images,
                    "matches_device_path": e2e_ok},
            "gpu_launches": launches,
            "roofline": roofline,
            "path_roofline": path,
            "path_hbm": {"algorithmic_bytes_per_image": alg_bytes, "achieved_gbs": path_gbs,
                         "frac_of_hbm_peak": path_gbs / pk["hbm_gbs"]},
            "stages_ms": {k: round(v[0] / max(prof_steps, 1), 4) for k, v in stages.items()},
            # algorithmic bytes of each stage / its single-stream time, as a fraction of the measured HBM peak
            "stages_hbm_frac": {k: round(bytes_per_image[k] * n_img / (v[0] / max(prof_steps, 1) / 1e3) / 1e9 / pk["hbm_gbs"], 4)
                                for k, v in stages.items() if k in bytes_per_image and v[0] > 0.02 * sum(w[0] for w in stages.values())},
            "cpu_baseline": cpu,
            "extra": extra,
            "clocks": clocks.summary(),
            "checksum": checksum,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    return rank, local_rank, world, dev


def run_c3(args):
    """BASELINE.json configs[2]: VLAD on VGG16 last-conv features (514-D, K=256, 196 descriptors / image), 100,000 images
    sharded over the ranks (strong scaling, no collective)."""
    import torch
    import torch.distributed as dist
    from pyvisim_b200 import _native as N, retrieval
    from pyvisim_b200.encoders import VLADEncoder
    from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
    from pyvisim_b200.features import Descriptors
    rank, local_rank, world, dev = _dist_setup()
    n_total, T, D, desc = WORKLOADS["c3"]
    if args.images:
        n_total = args.images
    lo, hi = retrieval.shard_bounds(n_total, world, rank)
    n = hi - lo
    gen = torch.Generator(device=dev).manual_seed(77 + rank)
    x = torch.empty((n * T, D), dtype=torch.float32, device=dev)
    for r in range(0, n * T, 1 << 20):
        x[r:r + (1 << 20)].normal_(0, 1, generator=gen).abs_()       # post-ReLU-like, non-negative
    cg = torch.Generator(device=dev).manual_seed(5)                  # the same random-init centres on every rank
    centers = torch.randn((256, D), device=dev, generator=cg).abs_().cpu().numpy()
    enc = VLADEncoder(feature_extractor=Descriptors(D), kmeans_model=kmeans_from_centers(centers))
    offs = torch.arange(n + 1, dtype=torch.int64) * T
    res = torch.empty((n, 256 * D), dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = lambda: enc.encode_descriptors(x, offs, images_per_call=4096, out=res)
    for _ in range(args.warmup):
        step()
    barrier()
    N.lib().pvs_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = int(N.lib().pvs_launch_count())
    N.profile_enable(True)
    step()
    barrier()
    stages = N.profile_read()
    N.profile_enable(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    pk = peaks()
    alg = T * D * 4 + 256 * D * 4                          # descriptors in, encoding out
    gbs = alg * n_total / (ms / 1e3) / 1e9
    dom = max(stages.items(), key=lambda kv: kv[1][0])
    dom_ms = dom[1][0] / dom[1][1]
    dom_imgs = n / dom[1][1]
    dom_gbs = alg * dom_imgs / (dom_ms / 1e3) / 1e9
    tdb = {}
    try:
        tdb = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    tr = tdb.get(dom[0] + "_c3")
    # e2e on a bounded subset through the public API with pinned host buffers
    m = min(n, 16384)
    xh = torch.empty((m * T, D), dtype=torch.float32, pin_memory=True)
    xh.copy_(x[:m * T])
    oh = torch.empty((m, 256 * D), dtype=torch.float32, pin_memory=True)
    offs_np = np.arange(m + 1, dtype=np.int64) * T
    enc.encode_descriptors(xh.numpy(), offs_np, out=oh.numpy())
    barrier()
    t0 = time.perf_counter()
    enc.encode_descriptors(xh.numpy(), offs_np, out=oh.numpy())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    checksum = float(sum(res[r:r + 4096].double().sum().item() for r in range(0, n, 4096)))
    if rank == 0:
        print(json.dumps({
            "metric": "vlad_encode_images_per_s_k256", "value": n_total / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": desc, "images_total": n_total, "images_per_gpu": n, "descriptors_per_image": T, "d": D, "k": 256,
                                            "weights": "random-init K-Means (no 514-D K-Means file is bundled)",
                                            "l2": f"inputs ({n * T * D * 4 / 1e9:.1f} GB/GPU) larger than L2, no flush",
                                            "parallelism": f"dp{world} (images sharded, no collective)"},
            "e2e": {"value": world * m / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(m * T * D * 4 + (m + 1) * 8), "d2h_bytes_per_step": int(m * 256 * D * 4),
                    "images": m, "note": "bounded subset per rank through encode_descriptors with pinned host buffers"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": dom_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": dom_gbs / pk["hbm_gbs"],
                         "traffic": tr["dram_bytes_per_image"] * dom_imgs if tr else None, "traffic_source": tr["source"] if tr else None,
                         "bytes_per_image": alg, "kernel_ms_per_launch": dom_ms, "images_per_launch": dom_imgs,
                         "note": "SURVEY.md 8(d) bytes of the whole path per image over the dominant kernel's time"},
            "path_hbm": {"achieved_gbs": gbs / world, "frac_of_hbm_peak": gbs / world / pk["hbm_gbs"], "note": "per GPU"},
            "stages_ms": {k: round(v[0], 4) for k, v in stages.items()},
            "clocks": clocks.summary(), "checksum": checksum}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c4(args):
    """BASELINE.json configs[3]: all-pairs cosine + top-100 over 262,144 VLAD-32768 vectors; database replicated in bf16
    (17.2 GB), query rows sharded over the ranks, top-k lists all-gathered (strong scaling)."""
    import torch
    import torch.distributed as dist
    from pyvisim_b200 import retrieval
    rank, local_rank, world, dev = _dist_setup()
    n, _, d, desc = WORKLOADS["c4"]
    if args.images:
        n = args.images
    pk = peaks()
    comm = None
    if world > 1:
        with stdout_to_stderr():
            comm = retrieval.NativeComm()
            torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        r, (s_, i_) = run_retrieval(dev, rank, world, n, d, 100, args.steps, max(1, min(args.warmup, 2)), pk, comm)
    # e2e on a bounded block: fp32 query vectors from pinned host memory -> normalise -> top-100 against the resident database
    # -> lists back to the host
    gen = torch.Generator(device=dev).manual_seed(4321)
    db = vlad_like_device(min(n, 65536), d, dev, gen, torch.bfloat16)
    m = 4096
    qh = torch.empty((m, d), dtype=torch.float32, pin_memory=True)
    qh.copy_(db[:m].float())
    sh, ih = torch.empty((m, 100), dtype=torch.float32, pin_memory=True), torch.empty((m, 100), dtype=torch.int64, pin_memory=True)

    def e2e():
        q = retrieval.l2_normalize(qh.to(dev, non_blocking=True), "bf16")
        a, b = retrieval.cosine_topk(q, db, 100)
        sh.copy_(a, non_blocking=True); ih.copy_(b, non_blocking=True)
        torch.cuda.synchronize()
    e2e()
    t0 = time.perf_counter()
    e2e()
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        tdb = {}
        try:
            tdb = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        tr = tdb.get("tc_sim_topk")
        print(json.dumps({
            "metric": "cosine_top100_queries_per_s", "value": r["queries_per_s"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": {"workload": desc, "n": n, "d": d, "k": 100, "parallelism": f"query rows sharded over {world} rank(s), "
                                            "database replicated in bf16, top-k lists all-gathered", "l2": "database (17.2 GB) larger than L2, no flush"},
            "e2e": {"value": m / e2e_s * (db.shape[0] / n), "unit": "queries/s", "h2d_bytes_per_step": int(m * d * 4), "d2h_bytes_per_step": int(m * 100 * 12),
                    "note": f"{m} fp32 query rows from pinned host memory against {db.shape[0]} resident rows, scaled to {n} rows"},
            "gpu_launches": r["gpu_launches_per_step"] * args.steps,
            "roofline": {"bound": "tensor", "kernel": "tc_sim_topk (SimPolicy: bf16 GEMM + fused top-k)", "achieved": r["tflops"] / world,
                         "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": r["tflops"] / world / pk["bf16_tflops_sustained"],
                         "frac_of_burst_peak": r["tflops"] / world / pk["bf16_tflops"], "traffic": tr["dram_bytes_per_launch"] if tr else None,
                         "traffic_source": tr["source"] if tr else None, "flop": 2.0 * n * n * d, "note": "per GPU; 2 N^2 D FLOP / step time"},
            "retrieval": r, "clocks": clocks.summary()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(args):
    """BASELINE.json configs[4]: Pipeline VLAD(RootSIFT) + FV(VGG16-PCA) concatenated encode, then all-pairs retrieval with
    top-100 lists and mAP / top-k accuracy over the whole set, images sharded over the ranks.  The encodings (164,608-D,
    bf16 after row normalisation: 329 GB for 1 M images) never exist in one place: every rank keeps its shard and the
    shards travel around a ring (`retrieval.all_pairs_topk_ring`); label metrics are evaluated on the device
    (`pvs_topk_label_metrics`) and reduced with one all-reduce.  Synthetic images: every class has a prototype descriptor
    for each extractor, an image is its class prototype plus noise (so that retrieval has something to find)."""
    import torch
    import torch.distributed as dist
    from pyvisim_b200 import _native as N, retrieval
    from pyvisim_b200.encoders import FisherVectorEncoder, VLADEncoder, GMMWeights, Pipeline
    from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
    from pyvisim_b200.features import Descriptors
    rank, local_rank, world, dev = _dist_setup()
    n_total, _, dim, desc = WORKLOADS["c5"]
    if args.images:
        n_total = args.images
    lo, hi = retrieval.shard_bounds(n_total, world, rank)
    n = hi - lo
    T1, T2, k = 2000, 196, 100
    per_class = 64
    classes = max(2, n_total // per_class)
    pg = torch.Generator(device=dev).manual_seed(2024)               # prototypes and centres: the same on every rank
    proto1 = torch.randn((classes, 128), device=dev, generator=pg).abs_()
    proto2 = torch.randn((classes, 514), device=dev, generator=pg)
    cen = torch.randn((256, 128), device=dev, generator=pg).abs_()
    cen = (cen / cen.sum(1, keepdim=True)).sqrt_().cpu().numpy()
    pipe = Pipeline([VLADEncoder(feature_extractor=Descriptors(128), kmeans_model=kmeans_from_centers(cen)),
                     FisherVectorEncoder(feature_extractor=Descriptors(514), weights=GMMWeights.OXFORD102_K256_VGG16_PCA,
                                         output_dtype=np.float32)])
    shard = torch.empty((n, dim), dtype=torch.bfloat16, device=dev)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    chunk = 1024
    enc_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N.lib().pvs_launch_count_reset()
    with ClockSampler(local_rank) as clocks:
        for c0 in range(0, n, chunk):
            m = min(chunk, n - c0)
            cls = (torch.arange(lo + c0, lo + c0 + m, device=dev) % classes)
            x1 = (proto1[cls][:, None, :] + 2.0 * torch.randn((m, T1, 128), device=dev, generator=gen)).abs_()
            x1 = (x1 / (x1.sum(2, keepdim=True) + 1e-7)).sqrt_().reshape(m * T1, 128)
            x2 = (proto2[cls][:, None, :] + 3.0 * torch.randn((m, T2, 514), device=dev, generator=gen)).clamp_min_(0).reshape(m * T2, 514)
            o1, o2 = torch.arange(m + 1, dtype=torch.int64) * T1, torch.arange(m + 1, dtype=torch.int64) * T2
            torch.cuda.synchronize()
            e0.record()
            enc = pipe.encode_descriptors([x1, x2], [o1, o2])
            shard[c0:c0 + m] = retrieval.l2_normalize(enc, "bf16")
            e1.record()
            torch.cuda.synchronize()
            enc_ms += e0.elapsed_time(e1)
            del x1, x2, enc
        enc_launches = int(N.lib().pvs_launch_count())
        torch.cuda.empty_cache()
        # warm-up on a 512-row slice per rank (NCCL point-to-point channels, workspaces), not timed
        retrieval.all_pairs_topk_ring(shard[:min(n, 512)], k + 1, rank=rank, world=world, normalized=True, gather=False)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        s_, i_ = retrieval.all_pairs_topk_ring(shard, k + 1, rank=rank, world=world, normalized=True, gather=False)
        labels = (torch.arange(n_total, device=dev) % classes).to(torch.int32)
        hits, ap = retrieval.label_metrics(i_[:, 1:].contiguous(), labels, labels[lo:hi])
        agg = torch.stack([ap.double().sum(), (hits > 0).double().sum(),
                           (i_[:, 0] == torch.arange(lo, hi, device=dev)).double().sum()])
        if world > 1:
            dist.all_reduce(agg)
        r1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ret_ms = r0.elapsed_time(r1)
    t = torch.tensor([enc_ms, ret_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    enc_ms, ret_ms = float(t[0]), float(t[1])
    pk = peaks()
    flop = 2.0 * n_total * n_total * dim
    if rank == 0:
        print(json.dumps({
            "metric": "pipeline_retrieval_queries_per_s", "value": n_total / (ret_ms / 1e3), "unit": "queries/s", "n_gpus": world, "steps": 1, "warmup": 0,
            "ms_per_step": ret_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 encode, bf16 similarity",
            "data": "synthetic", "config": {"workload": desc, "images_total": n_total, "images_per_gpu": n, "dim": dim, "k": k, "classes": classes,
                                            "descriptors_per_image": {"rootsift128": T1, "vgg16_514": T2},
                                            "weights": "VLAD: random-init K-Means (no RootSIFT K-Means file is bundled); FV: bundled gmm_k256_deep_features_vgg16_pca + PCA",
                                            "parallelism": f"images sharded over {world} rank(s); shards of the bf16 encodings travel around a ring; mAP all-reduced"},
            "encode": {"images_per_s": n_total / (enc_ms / 1e3), "ms": enc_ms, "gpu_launches": enc_launches,
                       "note": "Pipeline.encode_descriptors + row normalisation, descriptors generated on the device chunk by chunk (generation not timed)"},
            "retrieval": {"ms": ret_ms, "pflops": flop / ret_ms / 1e12, "frac_of_bf16_sustained_peak": flop / ret_ms / 1e9 / world / pk["bf16_tflops_sustained"],
                          "shard_gb_per_rank": n * dim * 2 / 1e9},
            "quality": {"map_at_100": float(agg[0]) / n_total, "top100_accuracy": float(agg[1]) / n_total, "self_is_first": float(agg[2]) / n_total,
                        "chance_precision": per_class / n_total},
            "e2e": {"value": n_total / ((enc_ms + ret_ms) / 1e3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
                    "note": "encode + retrieval + mAP; descriptors are synthesised on the device (1.4 MB per image would be 1.5 TB of host traffic)"},
            "clocks": clocks.summary()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"], help="c2 (default) = the configuration the metric is quoted on")
    ap.add_argument("--images", type=int, default=0, help="override images per GPU (debug)")
    ap.add_argument("--images-per-call", type=int, default=0, help="0 = library default (4 per SM)")
    ap.add_argument("--e2e-images", type=int, default=0, help="images in the host-buffer leg (0 = all)")
    ap.add_argument("--cpu-sample", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the VLAD / similarity side measurements")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-stage CUDA-event timing")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c3":
        run_c3(args)
    elif args.workload == "c4":
        run_c4(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

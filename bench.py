#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its configuration, B200 path vs the reference path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

Workload at N=1 (and per rank at N>1, weak scaling): BASELINE.json configs[1] --
Fisher-vector encode, GMM K=256 (diag) on SIFT-PCA-64 (bundled
gmm_k256_sift_pca + pca_k256_sift_f2), 8,189 synthetic images x 2,000 SIFT-like 128-D fp32
descriptors.  One step = one pass of the encode path over that batch.

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
}


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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_img, T, d_in, desc = WORKLOADS["c2"]
    sample = 32
    rng = np.random.default_rng(0)
    descs = [sift_like_host(rng, T, d_in) for _ in range(sample)]
    for _ in range(args.warmup):
        oracle_fv_rate(descs[:4])
    times = []
    for _ in range(args.steps):
        _, dt, _ = oracle_fv_rate(descs)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms / 1e3)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "sample": f"{sample} images x {T} descriptors per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} of {n_img} images per step, NumPy/OpenBLAS default threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    xs = np.vstack([d1, d2])
    gold = O.kmeans_predict(xs, cen)
    bad = np.flatnonzero(lab != gold)
    gaps = []
    if bad.size:
        sc = O.kmeans_scores(xs[bad], cen)
        gaps = (np.abs(sc[np.arange(bad.size), lab[bad]] - sc[np.arange(bad.size), gold[bad]]) / np.abs(sc).max(axis=1)).tolist()
    out["quickstart_vlad_similarity_score"] = {"latency_ms_median": 1e3 * float(np.median(lat)), "latency_ms_min": 1e3 * float(np.min(lat)),
                                               "cpu_oracle_ms": 1e3 * cpu_s, "score": float(np.asarray(score).ravel()[0]),
                                               "abs_diff_vs_oracle": float(abs(np.asarray(score).ravel()[0] - np.asarray(ref).ravel()[0])),
                                               "label_flips_vs_fp32_oracle": int(bad.size), "rel_score_gap_of_flips": gaps,
                                               "descriptors": [2000, 1900]}
    del enc
    n, d, k = 16384, 32768, 100
    v = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for r in range(0, n, 4096):                           # VLAD-shaped rows: 256 unit blocks of 128, ~15 % empty
        b = torch.randn((4096, d // 128, 128), device=dev, generator=g)
        b = b / b.norm(dim=2, keepdim=True)
        b = b * (torch.rand((4096, d // 128, 1), device=dev, generator=g) > 0.15)
        b = b.reshape(4096, -1)
        v[r:r + 4096] = (b / b.norm(dim=1, keepdim=True).clamp_min(1e-30)).bfloat16()
    ms, _ = timed(lambda: retrieval.cosine_topk(v, v, k), 2)
    tf = 2.0 * n * n * d / ms / 1e9
    out["all_pairs_cosine_top100"] = {"tflops": tf, "queries_per_s": n / ms * 1e3, "n": n, "d": d, "k": k, "dtype": "bf16",
                                      "frac_of_bf16_sustained_peak": tf / pk["bf16_tflops_sustained"],
                                      "frac_of_bf16_burst_peak": tf / pk["bf16_tflops"]}
    return out

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
        dist.init_process_group("nccl", device_id=dev)

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

    # ---- roofline of the dominant kernel (live CUDA-event stage times) -------------------
    pk = peaks()
    dominant = max(stages.items(), key=lambda kv: kv[1][0]) if stages else (None, (0.0, 0))
    flops_per_image = {"pca_project": 2 * T * d_in * D, "gmm_logits": 2 * T * K * 2 * D,
                       "fv_stats": 2 * T * K * 2 * D, "tc_fv_posterior": 2 * T * K * 2 * D,
                       "tc_fv_stats": 2 * T * K * 2 * D, "tc_fv_project": 2 * T * d_in * D,
                       "tc_fv_poststats_fused": 4 * T * K * 2 * D}        # logits + statistics in one kernel
    # algorithmic HBM bytes per image of each kernel: what its interface makes it move
    bytes_per_image = {"gmm_softmax": 2 * T * K * 4,
                       "fv_finalize": K * 2 * D * 4 + 16 * K * 4 + out_dim * 4,  # S + zeroth-order partials in, encoding out
                       "tc_fv_project": T * (d_in + D) * 4,                    # X in, Y out
                       "tc_fv_posterior": T * (D + K) * 4,                     # Y in, Q out (fp16 hi + lo planes = 4 B)
                       "tc_fv_stats": T * (D + K) * 4 + K * 2 * D * 4}         # Q + Y in, S out
    # The fp16x2 kernels (three kind::f16 MMAs per product on power-of-two-scaled fp16 hi + lo operands) need
    # half the tensor time of 3xTF32, which puts posterior / statistics / projection on the HBM roofline:
    # they stream Q (1 KB per descriptor) through HBM.  The tensor fraction is reported beside it.
    hbm_bound = {"tc_fv_project", "tc_fv_posterior", "tc_fv_stats", "gmm_softmax", "fv_finalize"}
    # DRAM bytes per image of each kernel from the committed `ncu --set full` capture
    # (profiles/ncu_fv_r01c.txt: dram read + write per launch of 592 images)
    ncu_dram_bytes_per_image = {"tc_fv_project": (0.606304e9 + 0.272666e9) / 592, "tc_fv_posterior": NCU_POST_BYTES / 592,
                                "tc_fv_stats": (1.534512e9 + 0.063308e9) / 592}
    roofline = None
    if dominant[0]:
        name, (ms, n) = dominant
        per_launch_ms = ms / n
        imgs_per_launch = n_img * prof_steps / n
        traffic = ncu_dram_bytes_per_image.get(name)
        tflops = flops_per_image.get(name, 0) * imgs_per_launch / (per_launch_ms / 1e3) / 1e12
        if name in hbm_bound:
            ach = bytes_per_image.get(name, 0) * imgs_per_launch / (per_launch_ms / 1e3) / 1e9
            roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / pk["hbm_gbs"], "traffic": traffic * imgs_per_launch if traffic else None,
                        "peak_source": f"hbm copy, {pk['source']}",
                        "algorithmic_bytes_per_launch": bytes_per_image.get(name, 0) * imgs_per_launch}
            if tflops:
                roofline["tensor"] = {"achieved_tflops": tflops, "frac_of_bf16_sustained": tflops / pk["bf16_tflops_sustained"],
                                      "note": "fp16x2: three kind::f16 MMAs per algorithmic product (ceiling = peak / 3)"}
        else:
            peak = pk["bf16_tflops_sustained"]
            roofline = {"bound": "tensor", "kernel": name, "achieved": tflops, "peak": peak, "unit": "TFLOP/s",
                        "frac": tflops / peak, "traffic": traffic * imgs_per_launch if traffic else None,
                        "peak_source": f"bf16 dense sustained, {pk['source']} (kernel runs inside a long step)",
                        "note": "fp16x2: three kind::f16 MMAs per algorithmic product, so the ceiling of this kernel is "
                                "peak / 3; frac_of_fp16x2_ceiling = 3 * frac",
                        "frac_of_fp16x2_ceiling": 3 * tflops / peak}
        roofline["kernel_ms_per_launch"] = per_launch_ms
        roofline["kernel_share_of_step"] = ms / sum(v[0] for v in stages.values())
        roofline["timing"] = f"{prof_steps} extra single-stream step(s), CUDA events around each launch"
    # whole-path HBM roofline (algorithmic bytes per image: descriptors in + encoding out)
    alg_bytes = T * d_in * 4 + out_dim * 4
    path_gbs = alg_bytes * n_img / (ms_per_step / 1e3) / 1e9

    # ---- CPU baseline: oracle port on a bounded sample (rank 0, N=1 only) ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample
        descs = [x[i * T:(i + 1) * T].cpu().numpy() for i in range(sample)]
        rate, dt, ref_out = oracle_fv_rate(descs)
        err = float(np.linalg.norm(out[:sample].cpu().numpy().astype(np.float64) - ref_out) / np.linalg.norm(ref_out))
        cpu = {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {sample} of {n_img} images, {dt:.1f} s, NumPy/OpenBLAS default threads",
               "parity_rel_l2_vs_gpu": err}

    extra = None
    if rank == 0 and world == 1 and not args.no_extra:
        del x
        torch.cuda.empty_cache()
        extra = run_extras(dev, pk)

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
            "path_hbm": {"algorithmic_bytes_per_image": alg_bytes, "achieved_gbs": path_gbs,
                         "frac_of_hbm_peak": path_gbs / pk["hbm_gbs"]},
            "stages_ms": {k: round(v[0] / max(prof_steps, 1), 4) for k, v in stages.items()},
            # algorithmic bytes of each stage / its single-stream time, as a fraction of the measured HBM peak
            "stages_hbm_frac": {k: round(bytes_per_image[k] * n_img / (v[0] / max(prof_steps, 1) / 1e3) / 1e9 / pk["hbm_gbs"], 4)
                                for k, v in stages.items() if k in bytes_per_image and v[0] > 0},
            "cpu_baseline": cpu,
            "extra": extra,
            "clocks": clocks.summary(),
            "checksum": checksum,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# DRAM read + write bytes of one posterior launch (592 images) in the committed ncu capture
NCU_POST_BYTES = 0.321038e9 + 1.172946e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""CPU-only tests: the C-ABI library loads and exports every declared symbol, the Python
mirror validates like the reference, and the multi-rank plumbing works under gloo."""
import os
import re
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT

from pyvisim_b200 import _native as N
from pyvisim_b200.encoders import VLADEncoder, FisherVectorEncoder, Pipeline, GMMWeights, KMeansWeights
from pyvisim_b200.encoders._base_encoder import check_desired_output, kmeans_from_centers, pack_descriptors, _PCA
from pyvisim_b200.features import Descriptors, Lambda, RootSIFT, SIFT, FeatureExtractorBase
from pyvisim_b200.retrieval import shard_bounds


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pvs_b200.h")).read()
    declared = set(re.findall(r"\b(pvs_[a-z0-9_]+)\s*\(", header))
    declared -= {"pvs_status", "pvs_model", "pvs_dtype", "pvs_path", "pvs_model_kind"}
    assert declared == set(N.EXPORTS), declared ^ set(N.EXPORTS)
    lib = N.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.pvs_version() == 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(N.PvsError) as e:
        N.Model.kmeans(np.zeros((4, 8), np.float32))
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "python-visual-similarity_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pvs_oracle" not in src and "oracle/" not in src, f


def test_weight_enums_and_missing_files():
    assert [m.name for m in KMeansWeights] == [m.name for m in GMMWeights] == [
        "OXFORD102_K256_VGG16_PCA", "OXFORD102_K256_VGG16", "OXFORD102_K256_ROOTSIFT_PCA",
        "OXFORD102_K256_ROOTSIFT", "OXFORD102_K256_SIFT_PCA", "OXFORD102_K256_SIFT"]
    gm = GMMWeights.OXFORD102_K256_SIFT_PCA.load()
    assert type(gm).__name__ == "GaussianMixture" and gm.means_.shape == (256, 64) and gm.n_features_in_ == 64
    pca = _PCA.OXFORD102_PCA256_VGG16.load()
    assert pca.components_.shape == (257, 514) and pca.n_components == 257 and pca.n_features_in_ == 514
    with pytest.raises(FileNotFoundError):
        KMeansWeights.OXFORD102_K256_ROOTSIFT.load()          # stripped from the reference checkout too
    with pytest.raises(FileNotFoundError):
        GMMWeights.OXFORD102_K256_VGG16.load()


def test_constructor_validation_follows_reference():
    km = kmeans_from_centers(np.zeros((8, 128), np.float32))
    with pytest.raises(TypeError):
        VLADEncoder(feature_extractor="sift", kmeans_model=km)
    with pytest.raises(ValueError):
        VLADEncoder(kmeans_model=GMMWeights.OXFORD102_K256_SIFT.load())
    with pytest.raises(ValueError):
        VLADEncoder(weights=GMMWeights.OXFORD102_K256_SIFT)
    with pytest.raises(ValueError):
        FisherVectorEncoder(weights=KMeansWeights.OXFORD102_K256_SIFT)
    with pytest.raises(RuntimeError):                       # extractor 64-D vs model 128-D
        VLADEncoder(feature_extractor=Descriptors(64), kmeans_model=km)
    with pytest.raises(ValueError):                         # PCA input 128 vs extractor 514
        FisherVectorEncoder(feature_extractor=Descriptors(514), weights=GMMWeights.OXFORD102_K256_SIFT_PCA)
    enc = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA)
    assert enc.pca is not None and enc.power_norm_weight == 0.5 and enc.encoding_dim == 33024
    v = VLADEncoder(kmeans_model=km)
    assert isinstance(v.feature_extractor, RootSIFT) and v.power_norm_weight == 1 and v.encoding_dim == 8 * 128
    # incompatible clustering model resets the PCA (default) or raises
    with pytest.warns(UserWarning):
        enc.clustering_model = GMMWeights.OXFORD102_K256_SIFT.load()
    assert enc.pca is None
    enc2 = FisherVectorEncoder(feature_extractor=Descriptors(128), weights=GMMWeights.OXFORD102_K256_SIFT_PCA,
                               raise_error_when_pca_incompatible=True)
    with pytest.raises(RuntimeError):
        enc2.clustering_model = GMMWeights.OXFORD102_K256_SIFT.load()
    with pytest.raises(ValueError):
        Pipeline([enc, "nope"])
    import torch
    with pytest.raises(RuntimeError):
        v.encode(torch.zeros(2, 3))


def test_similarity_func_contract_q4():
    good = lambda a, b: a @ b.T
    assert check_desired_output(good, np.ones((3, 4)), np.ones((5, 4))) is good
    scalar = lambda a, b: float((a * b).sum())
    with pytest.warns(UserWarning):
        wrapped = check_desired_output(scalar, np.ones((3, 4)), np.ones((5, 4)))
    out = wrapped(np.ones((3, 4)), np.ones((5, 4)))
    assert out.shape == (3, 5) and out.dtype == np.float32 and np.all(out == 4)
    km = kmeans_from_centers(np.zeros((8, 128), np.float32))
    enc = VLADEncoder(kmeans_model=km, similarity_func=good)
    assert enc.similarity_func is good


def test_feature_extractors_contract():
    assert SIFT().output_dim == 128 and RootSIFT().output_dim == 128
    assert isinstance(Descriptors(5), FeatureExtractorBase)
    lam = Lambda(lambda im: np.ones((3, 7), np.float32), 7)
    assert lam(np.zeros((4, 4, 3), np.uint8)).shape == (3, 7)
    with pytest.raises(ValueError):
        Lambda(lambda im: np.ones((3, 6), np.float32), 7)(np.zeros((4, 4, 3), np.uint8))
    assert Lambda(lambda im: None, 7)(np.zeros((4, 4, 3), np.uint8)).shape == (0, 7)
    with pytest.raises(ValueError):
        Lambda(3, 7)
    img = (np.random.default_rng(0).random((96, 96, 3)) * 255).astype(np.uint8)
    d = RootSIFT()(img)
    assert d.ndim == 2 and d.shape[1] == 128
    if len(d):
        assert np.allclose((d ** 2).sum(1), 1.0, atol=1e-3)


def test_pack_and_shard_bounds():
    x, offs = pack_descriptors([np.ones((2, 3)), np.zeros((0, 3)), 2 * np.ones((4, 3))], 3)
    assert x.dtype == np.float32 and x.shape == (6, 3) and list(offs) == [0, 2, 2, 6]
    for n, w in [(10, 4), (262144, 8), (7, 8), (0, 2)]:
        b = [shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


GLOO_WORKER = """
import os, sys
sys.path.insert(0, {pkg!r})
import numpy as np, torch, torch.distributed as dist
from pyvisim_b200.retrieval import gather_topk, shard_bounds
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank, n, k = dist.get_rank(), 11, 4
lo, hi = shard_bounds(n, 2, rank)
full_s = torch.arange(n * k, dtype=torch.float32).reshape(n, k)
full_i = torch.arange(n * k, dtype=torch.int64).reshape(n, k) * 3
s, i = gather_topk(full_s[lo:hi].clone(), full_i[lo:hi].clone(), n)
assert torch.equal(s, full_s) and torch.equal(i, full_i), (rank, s.shape)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_gather_topk_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(GLOO_WORKER.format(pkg=os.path.join(ROOT, "python-visual-similarity_b200"), port=port)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


RING_WORKER = """
import os, sys
sys.path.insert(0, {pkg!r})
import numpy as np, torch, torch.distributed as dist
from pyvisim_b200 import retrieval as R
W = 3
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=W)
rank = dist.get_rank()
# CPU stand-ins for the two device calls (test infrastructure: the ring schedule is what is under test)
def topk_cpu(q, db, k, index_offset=0):
    s = q @ db.T
    i = torch.argsort(-s, dim=1, stable=True)[:, :k]
    return torch.gather(s, 1, i), i + index_offset
def merge_cpu(sc, ix, k):
    parts, nq, kk = sc.shape
    s = sc.permute(1, 0, 2).reshape(nq, parts * kk)
    i = ix.permute(1, 0, 2).reshape(nq, parts * kk)
    key = torch.argsort(i, dim=1, stable=True)                    # lowest index first on ties ...
    s, i = torch.gather(s, 1, key), torch.gather(i, 1, key)
    o = torch.argsort(-s, dim=1, stable=True)[:, :k]              # ... then score descending (stable)
    return torch.gather(s, 1, o), torch.gather(i, 1, o)
R.cosine_topk, R.merge_topk = topk_cpu, merge_cpu
g = torch.Generator().manual_seed(0)
n, d, k = 37, 16, 5
x = torch.randn((n, d), generator=g)
x = x / x.norm(dim=1, keepdim=True)
x[7] = x[3]                                                        # exact ties across shards
lo, hi = R.shard_bounds(n, W, rank)
s, i = R.all_pairs_topk_ring(x[lo:hi].clone(), k, rank=rank, world=W, normalized=True, gather=True)
rs, ri = topk_cpu(x, x, k)
assert torch.equal(i, ri), (rank, (i != ri).nonzero()[:4])
assert torch.allclose(s, rs)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_ring_pass_sharded_database_world_size_3_gloo(tmp_path):
    """all_pairs_topk_ring: every rank only ever holds two shards of the database; three ranks, uneven shards,
    exact ties across shards; result = a single pass over the whole database."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "ring_worker.py"
    script.write_text(textwrap.dedent(RING_WORKER.format(pkg=os.path.join(ROOT, "python-visual-similarity_b200"), port=port)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(3)]
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_allgather_topk_argument_checks_without_a_communicator():
    """C1 export: a NULL communicator is a bad argument (no NCCL, no device needed)."""
    from pyvisim_b200 import _native as N
    lib = N.lib()
    rc = lib.pvs_allgather_topk(None, None, None, 4, 3, None, None, None)
    assert rc == -1 and b"communicator" in lib.pvs_last_error()
    assert lib.pvs_comm_destroy(None) == 0

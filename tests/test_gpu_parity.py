"""GPU parity tests: the CUDA path (through the C ABI) against the golden vectors written by
the real reference and against the NumPy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): hard-assignment / top-k indices exact except at
near-ties (relative score gap < 1e-6, which are listed and bounded, not hidden);
encodings and fp32 similarity within 1e-4 relative (vector L2);
bf16 similarity within 1e-2.
"""
import numpy as np
import pytest

import pvs_oracle as O
from conftest import load_golden, load_weights, split, rel_l2

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

FV_CASES = [
    ("fv_sift_pca", "OXFORD102_K256_SIFT_PCA"),
    ("fv_sift_pca_gmmsampled", "OXFORD102_K256_SIFT_PCA"),
    ("fv_rootsift_pca", "OXFORD102_K256_ROOTSIFT_PCA"),
    ("fv_rootsift_nopca", "OXFORD102_K256_ROOTSIFT"),
    ("fv_sift_nopca", "OXFORD102_K256_SIFT"),
    ("fv_vgg_pca", "OXFORD102_K256_VGG16_PCA"),
    ("fv_vgg_pca_gmmsampled", "OXFORD102_K256_VGG16_PCA"),
]


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyvisim_b200 import encoders, features, retrieval, eval as ev, _native, _utils
    import types
    return types.SimpleNamespace(enc=encoders, feat=features, ret=retrieval, ev=ev, nat=_native, utils=_utils)


def vlad_encoder(api, centers, d_in, pca=None, **kw):
    from pyvisim_b200.encoders._base_encoder import kmeans_from_centers
    return api.enc.VLADEncoder(feature_extractor=api.feat.Descriptors(d_in), kmeans_model=kmeans_from_centers(centers),
                               pca=pca, **kw)


def assert_labels(labels, gold, desc, centers, max_near_ties=2):
    """Exact equality except near-ties: every mismatch must have a relative fp64 score gap
    below 1e-6 between the two candidates, and there may be only a handful."""
    bad = np.flatnonzero(labels != gold)
    if bad.size == 0:
        return
    s = O.kmeans_scores(desc[bad], centers)
    gap = np.abs(s[np.arange(bad.size), labels[bad]] - s[np.arange(bad.size), gold[bad]])
    scale = np.abs(s).max(axis=1)
    assert np.all(gap <= 1e-6 * scale), f"label mismatches that are not near-ties: rows {bad}, gaps {gap / scale}"
    assert bad.size <= max_near_ties, f"{bad.size} near-tie flips: {list(zip(bad, gap / scale))}"


# ---------------------------------------------------------------------------------------
# VLAD
# ---------------------------------------------------------------------------------------
def test_vlad_rootsift128_golden(api):
    g = load_golden("vlad_rootsift128")
    descs = split(g["desc"], g["offsets"])
    enc = vlad_encoder(api, g["centers"], 128)
    out, labels = enc.encode_descriptors(descs, return_labels=True)
    assert out.dtype == np.float32 and out.shape == (4, 256 * 128)
    assert_labels(labels, g["labels"], g["desc"], g["centers"])
    assert rel_l2(out, g["out"]) <= 1e-4
    # the same through encode() (image loop + extractor) and flatten=False (quirk Q2)
    assert np.array_equal(enc.encode(descs), out)
    enc.flatten = False
    nf = enc.encode(descs[:2])
    assert nf.shape == (2 * 256, 128) and rel_l2(nf, g["out_noflatten_first2"]) <= 1e-4
    enc.flatten = True
    # non-default power / norm order
    enc2 = vlad_encoder(api, g["centers"], 128, power_norm_weight=0.5, norm_order=1)
    assert rel_l2(enc2.encode_descriptors(descs), g["out_pow05_l1"]) <= 1e-4
    # similarity_score -> float32 (1, 4)
    sim = enc.similarity_score(descs[2:3], descs)
    assert sim.dtype == np.float32 and sim.shape == (1, 4)
    assert np.abs(sim - g["sim_0_vs_rest"]).max() <= 1e-4


def test_vlad_aggregation_order_matches_reference(api):
    """Members of a cluster are summed in descriptor order, like the reference's Python
    loop, so with identical labels the residual sums are identical; only the norm
    reduction order differs (NumPy pairwise vs warp tree): ~1 ulp, far below 1e-4."""
    g = load_golden("vlad_rootsift128")
    descs = split(g["desc"], g["offsets"])
    out, labels = vlad_encoder(api, g["centers"], 128).encode_descriptors(descs, return_labels=True)
    if np.array_equal(labels, g["labels"]):
        assert rel_l2(out, g["out"]) <= 3e-7
        assert np.array_equal(out == 0, g["out"] == 0)


def test_vlad_quirk_q1_and_empty_rows(api):
    g = load_golden("vlad_rootsift128")
    descs = split(g["desc"], g["offsets"])
    enc = vlad_encoder(api, g["centers"], 128)
    q1 = enc.encode([descs[1], np.zeros((0, 128), np.float32), descs[2]])
    assert q1.shape == (256 * 128,) and q1.dtype == np.float32 and not q1.any()
    # bulk entry: an image without descriptors is a zero row, neighbours unaffected
    out = enc.encode_descriptors([descs[1], np.zeros((0, 128), np.float32), descs[2]])
    assert not out[1].any()
    assert rel_l2(out[[0, 2]], g["out"][[1, 2]]) <= 1e-4


def test_vlad_pca64_golden(api):
    g = load_golden("vlad_rootsift_pca64")
    from pyvisim_b200.encoders._base_encoder import _PCA
    enc = vlad_encoder(api, g["centers"], 128, pca=_PCA.OXFORD102_PCA256_ROOTSIFT.load())
    out = enc.encode_descriptors(g["desc"], g["offsets"])
    assert out.shape == (2, 256 * 64)
    assert rel_l2(out, g["out"]) <= 1e-4


def test_vlad_vgg514_golden(api):
    g = load_golden("vlad_vgg514")
    enc = vlad_encoder(api, g["centers"], 514)
    out, labels = enc.encode_descriptors(g["desc"], g["offsets"], return_labels=True)
    assert out.shape == (2, 131584)
    assert_labels(labels, g["labels"], g["desc"], g["centers"])
    assert rel_l2(out, g["out"]) <= 1e-4


@pytest.mark.parametrize("k,d,images", [(256, 100, 3), (32, 33, 2), (200, 514, 2), (64, 64, 5), (100, 128, 3), (256, 6, 2)])
def test_vlad_tensor_assignment_odd_shapes_vs_oracle(api, k, d, images):
    """tcgen05 assignment (3xTF32) on shapes the named configs do not cover: d not a multiple of
    32 / 4 / 2 (zero-padded k-blocks, 8- and 4-byte row alignment), k < 128 (the second CTA of
    the pair holds only padding), k not a multiple of 128, ragged images incl. a 1-row one.
    Labels vs the fp64 oracle (exact except near-ties), encodings vs the oracle, and the
    tensor path against the CUDA-core path."""
    rng = np.random.default_rng(k * 1000 + d)
    centers = rng.standard_normal((k, d)).astype(np.float32)
    descs = [(centers[rng.integers(0, k, t)] + 0.7 * rng.standard_normal((t, d))).astype(np.float32)
             for t in ([1, 300, 77, 257, 513][:images])]
    enc = vlad_encoder(api, centers, d)
    api.nat.set_path(api.nat.PATH_TENSOR)
    try:
        out_tc, lab_tc = enc.encode_descriptors(descs, return_labels=True)
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)
    desc_all = np.vstack(descs)
    assert_labels(lab_tc, O.kmeans_predict(desc_all, centers), desc_all, centers, max_near_ties=3)
    assert rel_l2(out_tc, O.vlad_encode(descs, centers)) <= 1e-4
    api.nat.set_path(api.nat.PATH_SIMT)
    try:
        out_cc, lab_cc = enc.encode_descriptors(descs, return_labels=True)
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)
    assert (lab_tc != lab_cc).sum() <= 3
    if np.array_equal(lab_tc, lab_cc):
        assert np.array_equal(out_tc, out_cc)      # same aggregation kernel, same order


@pytest.mark.parametrize("d", [128, 514, 64, 33])
def test_vlad_fp16x2_assignment_and_range_guard(api, d):
    """Hard assignment runs on fp16 hi+lo operands scaled by a power of two taken from the centres; a
    descriptor far outside fp16's range raises the device-side flag and the 3xTF32 kernel redoes the call."""
    import os
    rng = np.random.default_rng(40 + d)
    k = 256
    centers = (3.0 * rng.standard_normal((k, d))).astype(np.float32)
    descs = [(centers[rng.integers(0, k, t)] + 1.5 * rng.standard_normal((t, d))).astype(np.float32) for t in (700, 1, 300)]
    enc = vlad_encoder(api, centers, d)

    def run(dd, no_fp16):
        if no_fp16:
            os.environ["PVS_VLAD_NO_FP16X2"] = "1"
        try:
            api.nat.set_path(api.nat.PATH_TENSOR)
            return enc.encode_descriptors(dd, return_labels=True)
        finally:
            api.nat.set_path(api.nat.PATH_AUTO)
            os.environ.pop("PVS_VLAD_NO_FP16X2", None)

    desc_all = np.vstack(descs)
    gold = O.kmeans_predict(desc_all, centers)
    (out_h, lab_h), (out_t, lab_t) = run(descs, False), run(descs, True)
    assert_labels(lab_h, gold, desc_all, centers, max_near_ties=3)
    assert_labels(lab_t, gold, desc_all, centers, max_near_ties=3)
    assert rel_l2(out_h, O.vlad_encode(descs, centers)) <= 1e-4
    # 1e7 * 2^-e is far beyond fp16's 65504: the guard has to hand the call to the 3xTF32 kernel
    wild = [x.copy() for x in descs]
    wild[0][5, 3] = 1.0e7
    wild_all = np.vstack(wild)
    (out_hw, lab_hw), (out_tw, lab_tw) = run(wild, False), run(wild, True)
    assert np.array_equal(lab_hw, lab_tw) and np.array_equal(out_hw, out_tw)
    assert_labels(lab_hw, O.kmeans_predict(wild_all, centers), wild_all, centers, max_near_ties=3)


def test_vlad_long_image_takes_the_label_scan_path(api):
    """An image with more descriptors than the shared-memory member list holds (here forced
    by a batch whose mean T is small) must give the same block as the sorted path."""
    rng = np.random.default_rng(5)
    centers = rng.standard_normal((256, 64)).astype(np.float32)
    big = (centers[rng.integers(0, 256, 9000)] + 0.5 * rng.standard_normal((9000, 64))).astype(np.float32)
    small = [big[i * 10:(i + 1) * 10] for i in range(40)]
    enc = vlad_encoder(api, centers, 64)
    mixed = enc.encode_descriptors(small + [big])          # mean T ~ 230 -> t_cap 1024 < 9000
    alone = enc.encode_descriptors([big])                  # t_cap 32768: sorted path
    assert np.array_equal(mixed[-1], alone[0])
    assert rel_l2(mixed, O.vlad_encode(small + [big], centers)) <= 1e-4


def test_vlad_device_resident_equals_host_path(api):
    g = load_golden("vlad_vgg514")
    enc = vlad_encoder(api, g["centers"], 514)
    host = enc.encode_descriptors(g["desc"], g["offsets"])
    dev = enc.encode_descriptors(torch.from_numpy(g["desc"]).cuda(), torch.from_numpy(g["offsets"]))
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)


def test_vlad_properties_at_c3_shape(api):
    """Size-independent properties on a C3-shaped batch (T=196, D=514): every cluster
    block has unit norm or is exactly zero, blocks with no member are zero, and the
    vector norm is sqrt(#non-empty clusters)."""
    rng = np.random.default_rng(5)
    n, t, d = 64, 196, 514
    x = rng.standard_normal((n * t, d)).astype(np.float32)
    centers = rng.standard_normal((256, d)).astype(np.float32)
    offs = np.arange(n + 1, dtype=np.int64) * t
    enc = vlad_encoder(api, centers, d)
    out, labels = enc.encode_descriptors(x, offs, return_labels=True)
    blocks = out.reshape(n, 256, d)
    norms = np.linalg.norm(blocks, axis=2)
    counts = np.stack([np.bincount(labels[i * t:(i + 1) * t], minlength=256) for i in range(n)])
    assert np.all(norms[counts == 0] == 0)
    assert np.allclose(norms[counts > 0], 1.0, atol=1e-5)
    assert np.allclose(np.linalg.norm(out, axis=1), np.sqrt((counts > 0).sum(1)), atol=1e-4)
    ref_labels = O.kmeans_predict(x, centers)
    assert_labels(labels, ref_labels, x, centers, max_near_ties=4)


# ---------------------------------------------------------------------------------------
# Fisher vectors
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,member", FV_CASES)
def test_fv_golden(api, case, member):
    g = load_golden(case)
    d_in = g["desc"].shape[1]
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(d_in),
                                      weights=api.enc.GMMWeights[member])
    descs = split(g["desc"], g["offsets"])
    out = enc.encode(descs)
    assert out.dtype == np.float64 and out.shape == g["out"].shape
    err = rel_l2(out, g["out"])
    assert err <= 1e-4, f"{case}: rel-L2 {err:.3e}"
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-5)
    enc2 = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(d_in),
                                       weights=api.enc.GMMWeights[member], power_norm_weight=1.0, norm_order=1)
    err2 = rel_l2(enc2.encode(descs[:1]), g["out_pow1_l1_img0"])
    assert err2 <= 1e-4, f"{case} (power 1, L1): rel-L2 {err2:.3e}"


def test_fv_posterior_matches_predict_proba(api):
    g = load_golden("fv_sift_pca_gmmsampled")
    w = load_weights("gmm_k256_sift_pca")
    p = load_weights("pca_k256_sift_f2")
    d0 = split(g["desc"], g["offsets"])[0]
    y0 = O.pca_transform(d0.astype(np.float32), p["components"], p["mean"])
    model = api.nat.Model.gmm(w["weights"], w["means"], w["covariances"], w["precisions_cholesky"])
    y = torch.from_numpy(y0).cuda()
    q = torch.empty((y.shape[0], 256), dtype=torch.float32, device="cuda")
    api.nat.check(api.nat.lib().pvs_gmm_posterior(model.handle, y.data_ptr(), y.shape[0], q.data_ptr(), None))
    torch.cuda.synchronize()
    q = q.cpu().numpy()
    gold = g["posterior_img0"]
    assert np.abs(q - gold).max() <= 2e-4
    assert np.allclose(q.sum(1), 1.0, atol=1e-5)
    flips = np.flatnonzero(q.argmax(1) != gold.argmax(1))
    for r in flips:                                    # only allowed at posterior near-ties
        top2 = np.sort(gold[r])[-2:]
        assert abs(top2[1] - top2[0]) <= 1e-3, f"arg-max flip at row {r} with gap {top2[1] - top2[0]}"


def test_fv_empty_image_raises_like_reference(api):
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    with pytest.raises(ValueError):
        enc.encode([np.zeros((0, 128), np.float32)])


def test_fv_device_resident_equals_host_and_oracle_c2_shape(api):
    """C2-shaped images (T=2000, SIFT-like, PCA-64) vs the fp64 oracle."""
    rng = np.random.default_rng(11)
    n, t = 6, 2000
    x = np.floor(np.clip(np.abs(rng.normal(0, 40, (n * t, 128))), 0, 255)).astype(np.float32)
    offs = np.arange(n + 1, dtype=np.int64) * t
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    host = enc.encode_descriptors(x, offs)
    dev = enc.encode_descriptors(torch.from_numpy(x).cuda(), torch.from_numpy(offs), images_per_call=4)
    assert np.array_equal(dev.cpu().numpy(), host)
    w = load_weights("gmm_k256_sift_pca")
    p = load_weights("pca_k256_sift_f2")
    ref = O.fv_encode(split(x, offs), w["weights"], w["means"], w["covariances"], w["precisions_cholesky"],
                      pca=(p["components"], p["mean"]))
    assert rel_l2(host, ref) <= 1e-4
    assert np.allclose(np.linalg.norm(host, axis=1), 1.0, atol=1e-5)


def test_fv_image_independence_across_chunks_and_batch_order(api):
    """Size-independent properties on a multi-chunk, ragged batch (fp16x2 path, two streams): unit norms;
    an image's encoding depends on nothing but its own descriptors -- the same images in reverse order and
    with a different chunking give bit-identical rows; a subset against the fp64 oracle."""
    rng = np.random.default_rng(77)
    n = 700
    ts = rng.integers(40, 400, n)
    ts[[0, 5, n - 1]] = [1, 16, 17]
    offs = np.concatenate([[0], np.cumsum(ts)]).astype(np.int64)
    x = np.floor(np.clip(np.abs(rng.normal(0, 40, (int(offs[-1]), 128))), 0, 255)).astype(np.float32)
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    xd = torch.from_numpy(x).cuda()
    a = enc.encode_descriptors(xd, torch.from_numpy(offs), images_per_call=256).cpu().numpy()
    assert np.allclose(np.linalg.norm(a, axis=1), 1.0, atol=1e-5)
    order = np.arange(n)[::-1]
    xr = np.concatenate([x[offs[i]:offs[i + 1]] for i in order])
    offr = np.concatenate([[0], np.cumsum(ts[order])]).astype(np.int64)
    b = enc.encode_descriptors(torch.from_numpy(xr).cuda(), torch.from_numpy(offr), images_per_call=97).cpu().numpy()
    assert np.array_equal(a, b[::-1])
    w = load_weights("gmm_k256_sift_pca")
    p = load_weights("pca_k256_sift_f2")
    pick = [0, 5, 123, n - 1]
    ref = O.fv_encode([x[offs[i]:offs[i + 1]] for i in pick], w["weights"], w["means"], w["covariances"],
                      w["precisions_cholesky"], pca=(p["components"], p["mean"]))
    for r, i in enumerate(pick):
        assert rel_l2(a[i], ref[r]) <= 1e-4, (i, rel_l2(a[i], ref[r]))


def test_fv_generic_k32(api):
    """A k=32 vocabulary (what getting_started.ipynb trains) takes the generic path:
    shape formula 2*32*64+32 = 4128 and parity with the oracle."""
    rng = np.random.default_rng(3)
    k, d = 32, 64
    w = rng.random(k); w /= w.sum()
    mu = rng.standard_normal((k, d))
    var = rng.random((k, d)) + 0.5
    pc = 1.0 / np.sqrt(var)
    from sklearn.mixture import GaussianMixture
    gm = GaussianMixture(n_components=k, covariance_type="diag")
    gm.weights_, gm.means_, gm.covariances_, gm.precisions_cholesky_ = w, mu, var, pc
    gm.n_features_in_ = d
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(d), gmm_model=gm)
    descs = [rng.standard_normal((t, d)).astype(np.float32) for t in (17, 200)]
    out = enc.encode(descs)
    assert out.shape == (2, 4128)
    assert rel_l2(out, O.fv_encode(descs, w, mu, var, pc)) <= 1e-4


# ---------------------------------------------------------------------------------------
# Pipeline, similarity, top-k
# ---------------------------------------------------------------------------------------
def test_pipeline_golden_and_hstack_identity(api):
    g = load_golden("pipeline_rootsift")
    descs = split(g["desc"], g["offsets"])
    vlad = vlad_encoder(api, g["centers"], 128)
    fv = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                     weights=api.enc.GMMWeights.OXFORD102_K256_ROOTSIFT_PCA)
    pipe = api.enc.Pipeline([vlad, fv])
    out = pipe.encode(descs)
    assert out.dtype == np.float64 and out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) <= 1e-4
    assert np.array_equal(out, np.hstack([vlad.encode(descs), fv.encode(descs)]))      # pipeline.ipynb:305/336
    sim = pipe.similarity_score(descs[:2], descs)
    assert sim.dtype == np.float32 and np.abs(sim - g["sim_first2_vs_all"]).max() <= 1e-4
    db = {f"img{i}": out[i] for i in range(6)}
    top = api.ev.retrieve_top_k_similar(descs[3][None], db, pipe, k=4)
    assert [int(p[3:]) for p, _ in top] == list(g["top4_names"])
    assert np.abs(np.array([s for _, s in top]) - g["top4_scores"]).max() <= 1e-4


def test_cosine_similarity_contract(api):
    g = load_golden("cosine_small")
    cs = api.utils.cosine_similarity
    s32 = cs(g["a"], g["b"])
    s64 = cs(g["a"].astype(np.float64), g["b"])
    assert s32.dtype == np.float32 and s64.dtype == np.float64 and s32.shape == (9, 13)
    assert np.abs(s32 - g["s32"]).max() <= 1e-5 and np.abs(s64 - g["s64"]).max() <= 1e-5
    assert not s32[2].any()
    assert cs(np.ones(4), np.ones((2, 4))).shape == (1, 2)
    with pytest.raises(ValueError):
        cs(np.ones((3, 1)), np.ones((2, 1)))
    x = np.random.default_rng(0).standard_normal((33, 700)).astype(np.float32)
    assert np.allclose(np.diag(cs(x, x)), 1.0, atol=1e-5)


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_topk_vs_oracle(api, dtype, tol):
    rng = np.random.default_rng(21)
    nq, ndb, d, k = 37, 5000, 256, 100
    db = rng.standard_normal((ndb, d)).astype(np.float32)
    q = db[:nq] + 0.5 * rng.standard_normal((nq, d)).astype(np.float32)
    s_ref, i_ref = O.cosine_topk(q, db, k)
    qn = api.ret.l2_normalize(torch.from_numpy(q).cuda(), dtype)
    dbn = api.ret.l2_normalize(torch.from_numpy(db).cuda(), dtype)
    s, i = api.ret.cosine_topk(qn, dbn, k)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert np.all(np.diff(s, axis=1) <= 0)
    assert np.abs(s - s_ref).max() <= tol
    full = O.cosine_similarity(q, db)
    if dtype == "fp32":
        mism = np.argwhere(i != i_ref)
        for r, c in mism:       # swaps only between scores closer than 1e-6
            assert abs(full[r, i[r, c]] - full[r, i_ref[r, c]]) <= 1e-6
        assert len(mism) <= 8
    else:                       # bf16: retrieved set must be near-optimal in true score
        kth = s_ref[:, -1]
        assert np.all(np.take_along_axis(full, i, 1).min(1) >= kth - 2e-2)


def test_topk_ties_lowest_index_first_and_small_db(api):
    d = 64
    base = np.random.default_rng(2).standard_normal((5, d)).astype(np.float32)
    db = np.tile(base, (40, 1))                       # 200 rows, every vector repeated 40x
    qn = api.ret.l2_normalize(torch.from_numpy(base[:2]).cuda())
    dbn = api.ret.l2_normalize(torch.from_numpy(db).cuda())
    s, i = api.ret.cosine_topk(qn, dbn, 10)
    i = i.cpu().numpy()
    assert np.array_equal(i[0], np.arange(0, 50, 5)) and np.array_equal(i[1], np.arange(1, 51, 5))
    s2, i2 = api.ev.topk_host(base[:1], db[:7], 100)       # k clipped to the database size
    assert i2.shape == (1, 7) and sorted(i2[0]) == list(range(7))


def test_topk_merge_and_label_metrics(api):
    rng = np.random.default_rng(9)
    nq, ndb, d, k, parts = 16, 1200, 96, 20, 3
    db = rng.standard_normal((ndb, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    qn = api.ret.l2_normalize(torch.from_numpy(q).cuda())
    dbn = api.ret.l2_normalize(torch.from_numpy(db).cuda())
    s_full, i_full = api.ret.cosine_topk(qn, dbn, k)
    ss, ii = [], []
    for p in range(parts):
        lo, hi = api.ret.shard_bounds(ndb, parts, p)
        s, i = api.ret.cosine_topk(qn, dbn[lo:hi], k, index_offset=lo)
        ss.append(s); ii.append(i)
    s_m, i_m = api.ret.merge_topk(torch.stack(ss), torch.stack(ii), k)
    assert torch.equal(i_m, i_full) and torch.equal(s_m, s_full)
    db_labels = torch.from_numpy(rng.integers(0, 10, ndb).astype(np.int32))
    q_labels = torch.from_numpy(rng.integers(0, 10, nq).astype(np.int32))
    hits, ap = api.ret.label_metrics(i_full, db_labels, q_labels)
    idx = i_full.cpu().numpy()
    assert np.isclose(hits.float().mean().item(), O.top_k_accuracy_from_lists(idx, db_labels.numpy(), q_labels.numpy()))
    assert np.isclose(ap.mean().item(), O.top_k_map_from_lists(idx, db_labels.numpy(), q_labels.numpy()), atol=1e-6)


def test_all_pairs_row_block_sharding_matches_single_pass(api):
    """Emulates W=4 ranks on one GPU: each 'rank' scores its own row block against the
    replicated database; concatenating the blocks equals the single-pass result exactly."""
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.standard_normal((1000, 512)).astype(np.float32)).cuda()
    s1, i1 = api.ret.all_pairs_topk(x, 10, dtype="fp32")
    blocks = [api.ret.all_pairs_topk(x, 10, dtype="fp32", rank=r, world=4, gather=False) for r in range(4)]
    assert torch.equal(torch.cat([b[1] for b in blocks]), i1)
    assert torch.equal(torch.cat([b[0] for b in blocks]), s1)
    assert torch.equal(i1[:, 0].cpu(), torch.arange(1000))          # every vector's best match is itself


def test_eval_functions(api):
    rng = np.random.default_rng(8)
    g = load_golden("vlad_rootsift128")
    enc = vlad_encoder(api, g["centers"], 128)
    descs = [np.sqrt(np.abs(rng.standard_normal((60, 128))).astype(np.float32) / 128) for _ in range(12)]
    vecs = enc.encode(descs)
    emap = {f"p{i}": vecs[i] for i in range(12)}
    labels = {f"p{i}": i % 3 for i in range(12)}
    qs, ql = descs[:5], [0, 1, 2, 0, 1]
    acc = api.ev.top_k_accuracy(qs, ql, emap, labels, enc, k=3)
    mp = api.ev.top_k_map(qs, ql, emap, labels, enc, k=5)
    _, idx = O.cosine_topk(vecs[:5], vecs, 5)
    dbl = np.array([i % 3 for i in range(12)])
    assert np.isclose(acc, O.top_k_accuracy_from_lists(idx[:, :3], dbl, np.array(ql)))
    assert np.isclose(mp, O.top_k_map_from_lists(idx, dbl, np.array(ql)))


# ---------------------------------------------------------------------------------------
# tensor-core (tcgen05, 3xTF32) path vs CUDA-core path
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case,member", [c for c in FV_CASES if c[0] in ("fv_sift_pca", "fv_sift_pca_gmmsampled", "fv_rootsift_pca")])
def test_fv_tensor_path_matches_golden_and_simt(api, case, member):
    g = load_golden(case)
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128), weights=api.enc.GMMWeights[member])
    descs = split(g["desc"], g["offsets"])
    try:
        api.nat.set_path(api.nat.PATH_TENSOR)
        out_tc = enc.encode(descs)
        api.nat.set_path(api.nat.PATH_SIMT)
        out_simt = enc.encode(descs)
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)
    e_tc, e_simt = rel_l2(out_tc, g["out"]), rel_l2(out_simt, g["out"])
    assert e_tc <= 1e-4 and e_simt <= 1e-4, (e_tc, e_simt)
    assert rel_l2(out_tc, out_simt) <= 1e-4


def test_fv_tensor_path_ragged_batch_vs_oracle(api):
    """Ragged T (1, 31, 32, 33, 128, 129, 517, 2000), GMM-sampled descriptors so the
    posteriors are spread over many components; tensor path forced."""
    w = load_weights("gmm_k256_sift_pca")
    p = load_weights("pca_k256_sift_f2")
    r = np.random.RandomState(5)
    descs = []
    for t in (1, 31, 32, 33, 128, 129, 517, 2000):
        comp = r.choice(256, size=t, p=w["weights"] / w["weights"].sum())
        y = w["means"][comp] + r.standard_normal((t, 64)) * np.sqrt(w["covariances"][comp])
        descs.append((y @ p["components"] + p["mean"]).astype(np.float32))
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    try:
        api.nat.set_path(api.nat.PATH_TENSOR)
        out = enc.encode(descs)
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)
    ref = O.fv_encode(descs, w["weights"], w["means"], w["covariances"], w["precisions_cholesky"],
                      pca=(p["components"], p["mean"]))
    errs = [rel_l2(out[i], ref[i]) for i in range(len(descs))]
    assert max(errs) <= 1e-4, errs


# fused kernels under test: PVS_FV_FUSED=1 (one CTA per SM), =2 (2-CTA clusters that split the components, statistics folded
# in segments like the default two-kernel path)
FUSED_MODES = ["1", "2"]


def test_fv_fp16x2_path_and_range_guard(api):
    """K=256 / D=64 with a PCA runs the posterior / statistics contractions on fp16 hi+lo operands
    (scaled by powers of two from the mixture model).  A descriptor far outside the model's range
    raises the device-side flag and the same call is served by the 3xTF32 kernels instead."""
    import os
    w = load_weights("gmm_k256_sift_pca")
    p = load_weights("pca_k256_sift_f2")
    rng = np.random.default_rng(21)
    descs = [np.floor(np.clip(np.abs(rng.normal(0, 40, (t, 128))), 0, 255)).astype(np.float32) for t in (700, 33, 2000)]
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    ref = O.fv_encode(descs, w["weights"], w["means"], w["covariances"], w["precisions_cholesky"],
                      pca=(p["components"], p["mean"]))

    def run(d, no_fp16):
        if no_fp16:
            os.environ["PVS_FV_NO_FP16X2"] = "1"
        try:
            api.nat.set_path(api.nat.PATH_TENSOR)
            return enc.encode(d)
        finally:
            api.nat.set_path(api.nat.PATH_AUTO)
            os.environ.pop("PVS_FV_NO_FP16X2", None)

    fast, slow = run(descs, False), run(descs, True)
    assert rel_l2(fast, ref) <= 1e-4 and rel_l2(slow, ref) <= 1e-4, (rel_l2(fast, ref), rel_l2(slow, ref))
    assert not np.array_equal(fast, slow), "the fp16x2 kernels did not run"
    assert rel_l2(fast, slow) <= 5e-5
    # one descriptor 200x out of range (|y| ~ 1e5 >> 255 * 2^3): the guard must hand the call to 3xTF32
    wild = [d.copy() for d in descs]
    wild[1][7] *= 200.0
    ref_w = O.fv_encode(wild, w["weights"], w["means"], w["covariances"], w["precisions_cholesky"],
                        pca=(p["components"], p["mean"]))
    fast_w, slow_w = run(wild, False), run(wild, True)
    assert np.isfinite(fast_w).all()
    assert np.array_equal(fast_w, slow_w), "range guard did not fall back to the 3xTF32 kernels"
    assert rel_l2(fast_w, ref_w) <= 1e-4


@pytest.mark.parametrize("mode", FUSED_MODES)
def test_fv_fused_posterior_statistics_kernel(api, mode):
    """Fused kernels (PVS_FV_FUSED=2 and =1: posterior + statistics in one kernel, the [y^2|y] tile serving as
    K-major operand of the logit MMA and as MN-major operand of the statistics MMA): ragged images incl.
    T = 1, 127, 128, 129 vs the fp64 oracle and vs the default (unfused) path."""
    import os
    w = load_weights("gmm_k256_sift_pca")
    p = load_weights("pca_k256_sift_f2")
    r = np.random.RandomState(9)
    descs = []
    for t in (1, 127, 128, 129, 300, 2000, 16, 512, 513, 1100):        # 512 / 513: one segment of four tiles / one tile more
        comp = r.choice(256, size=t, p=w["weights"] / w["weights"].sum())
        y = w["means"][comp] + r.standard_normal((t, 64)) * np.sqrt(w["covariances"][comp])
        descs.append((y @ p["components"] + p["mean"]).astype(np.float32))
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    os.environ["PVS_FV_FUSED"] = "0"                          # the two unfused kernels
    base = enc.encode(descs)
    os.environ["PVS_FV_FUSED"] = mode
    try:
        n0 = api.nat.lib().pvs_launch_count()
        fused = enc.encode(descs)
        launches = api.nat.lib().pvs_launch_count() - n0
    finally:
        os.environ.pop("PVS_FV_FUSED", None)
    ref = O.fv_encode(descs, w["weights"], w["means"], w["covariances"], w["precisions_cholesky"],
                      pca=(p["components"], p["mean"]))
    errs = [rel_l2(fused[i], ref[i]) for i in range(len(descs))]
    assert max(errs) <= 1e-4, errs
    assert rel_l2(fused, base) <= 5e-5
    assert not np.array_equal(fused, base), "the fused kernel did not run"
    assert launches == 6          # project, fused, segment-count kernel + two gated 3xTF32 kernels, finalize


@pytest.mark.parametrize("mode", FUSED_MODES)
def test_fv_fused_kernels_many_images_per_cta(api, mode):
    """More images than CTAs / clusters (several images per persistent CTA, statistics accumulator reused,
    image-end hand-over) and two chunkings: the fused kernels against the default path, bit-identical between
    the chunkings."""
    import os
    rng = np.random.default_rng(123)
    n = 400
    ts = rng.integers(1, 300, n)
    ts[:4] = [1, 128, 129, 256]
    offs = np.concatenate([[0], np.cumsum(ts)]).astype(np.int64)
    x = np.floor(np.clip(np.abs(rng.normal(0, 40, (int(offs[-1]), 128))), 0, 255)).astype(np.float32)
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT_PCA)
    xd, od = torch.from_numpy(x).cuda(), torch.from_numpy(offs)
    os.environ["PVS_FV_FUSED"] = "0"                          # the two unfused kernels
    base = enc.encode_descriptors(xd, od, images_per_call=400).cpu().numpy()
    os.environ["PVS_FV_FUSED"] = mode
    try:
        a = enc.encode_descriptors(xd, od, images_per_call=400).cpu().numpy()
        b = enc.encode_descriptors(xd, od, images_per_call=150).cpu().numpy()
    finally:
        os.environ.pop("PVS_FV_FUSED", None)
    assert np.array_equal(a, b)
    assert np.isfinite(a).all() and not np.array_equal(a, base)
    assert max(rel_l2(a[i], base[i]) for i in range(n)) <= 5e-5


def test_fv_tensor_path_without_pca_d64(api):
    """K=256, D=64 GMM fed 64-D descriptors directly (no PCA stage)."""
    w = load_weights("gmm_k256_root_sift_pca")
    from sklearn.mixture import GaussianMixture
    gm = GaussianMixture(n_components=256, covariance_type="diag")
    gm.weights_, gm.means_, gm.covariances_, gm.precisions_cholesky_ = (w["weights"], w["means"], w["covariances"],
                                                                        w["precisions_cholesky"])
    gm.n_features_in_ = 64
    r = np.random.RandomState(2)
    descs = []
    for t in (40, 300):
        comp = r.choice(256, size=t)
        descs.append((w["means"][comp] + r.standard_normal((t, 64)) * np.sqrt(w["covariances"][comp])).astype(np.float32))
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(64), gmm_model=gm)
    try:
        api.nat.set_path(api.nat.PATH_TENSOR)
        out = enc.encode(descs)
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)
    ref = O.fv_encode(descs, w["weights"], w["means"], w["covariances"], w["precisions_cholesky"])
    assert rel_l2(out, ref) <= 1e-4


def test_tensor_path_refuses_unsupported_shape(api):
    enc = api.enc.FisherVectorEncoder(feature_extractor=api.feat.Descriptors(128),
                                      weights=api.enc.GMMWeights.OXFORD102_K256_SIFT)      # D = 128
    try:
        api.nat.set_path(api.nat.PATH_TENSOR)
        with pytest.raises(api.nat.PvsError):
            enc.encode([np.ones((5, 128), np.float32)])
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)


@pytest.mark.parametrize("nq,ndb,d,k", [(37, 5000, 256, 100), (300, 3000, 1024, 10), (129, 700, 64, 1), (5, 257, 128, 300)])
def test_topk_tensor_path_bf16_matches_cuda_core_path(api, nq, ndb, d, k):
    """The tcgen05 similarity/top-k kernel and the CUDA-core path consume the same bf16
    operands; scores must agree to fp32 accumulation error and the index lists must be
    identical except where two scores differ by less than that error."""
    rng = np.random.default_rng(nq + ndb)
    db = rng.standard_normal((ndb, d)).astype(np.float32)
    q = db[rng.integers(0, ndb, nq)] + 0.7 * rng.standard_normal((nq, d)).astype(np.float32)
    qn = api.ret.l2_normalize(torch.from_numpy(q).cuda(), "bf16")
    dbn = api.ret.l2_normalize(torch.from_numpy(db).cuda(), "bf16")
    try:
        api.nat.set_path(api.nat.PATH_TENSOR)
        s_tc, i_tc = api.ret.cosine_topk(qn, dbn, k, index_offset=7)
        api.nat.set_path(api.nat.PATH_SIMT)
        s_cc, i_cc = api.ret.cosine_topk(qn, dbn, k, index_offset=7)
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)
    s_tc, i_tc, s_cc, i_cc = (t.cpu().numpy() for t in (s_tc, i_tc, s_cc, i_cc))
    kk = min(k, ndb)
    assert np.all(i_tc[:, kk:] == -1) and np.all(i_cc[:, kk:] == -1)
    assert np.all(np.diff(s_tc[:, :kk], axis=1) <= 0)
    # the tensor core's fp32 accumulation truncates, CUDA cores round: ~1e-6 apart at d = 1024
    assert np.abs(s_tc[:, :kk] - s_cc[:, :kk]).max() <= 1e-5
    full = (qn.float() @ dbn.float().T).cpu().numpy()
    for r, c in np.argwhere(i_tc[:, :kk] != i_cc[:, :kk]):
        assert abs(full[r, i_tc[r, c] - 7] - full[r, i_cc[r, c] - 7]) <= 1e-5
    # and against the fp64 oracle on the original vectors, bf16 tolerance
    s_ref, _ = O.cosine_topk(q, db, kk)
    assert np.abs(s_tc[:, :kk] - s_ref).max() <= 1e-2


def test_topk_tensor_path_exact_ties(api):
    base = np.random.default_rng(2).standard_normal((5, 128)).astype(np.float32)
    db = np.tile(base, (120, 1))                      # 600 rows: every vector 120 times
    qn = api.ret.l2_normalize(torch.from_numpy(base[:2]).cuda(), "bf16")
    dbn = api.ret.l2_normalize(torch.from_numpy(db).cuda(), "bf16")
    try:
        api.nat.set_path(api.nat.PATH_TENSOR)
        _, i = api.ret.cosine_topk(qn, dbn, 10)
    finally:
        api.nat.set_path(api.nat.PATH_AUTO)
    i = i.cpu().numpy()
    assert np.array_equal(i[0], np.arange(0, 50, 5)) and np.array_equal(i[1], np.arange(1, 51, 5))

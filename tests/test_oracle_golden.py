"""Pin the NumPy oracle against outputs of the real reference (tests/golden/*.npz,
written by tests/golden/make_golden.py from pyvisim + scikit-learn)."""
import numpy as np
import pytest

import pvs_oracle as O
from conftest import load_golden, load_weights, split, rel_l2

FV_CASES = [
    ("fv_sift_pca", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
    ("fv_sift_pca_gmmsampled", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
    ("fv_rootsift_pca", "gmm_k256_root_sift_pca", "pca_k256_root_sift_f2"),
    ("fv_rootsift_nopca", "gmm_k256_root_sift_no_pca", None),
    ("fv_sift_nopca", "gmm_k256_sift_no_pca", None),
    ("fv_vgg_pca", "gmm_k256_deep_features_vgg16_pca", "pca_k256_deep_features_vgg16_f2"),
    ("fv_vgg_pca_gmmsampled", "gmm_k256_deep_features_vgg16_pca", "pca_k256_deep_features_vgg16_f2"),
]


def gmm_args(g):
    return g["weights"], g["means"], g["covariances"], g["precisions_cholesky"]


def test_vlad_rootsift128_labels_and_encoding():
    g = load_golden("vlad_rootsift128")
    descs = split(g["desc"], g["offsets"])
    labels = np.concatenate([O.kmeans_predict(d, g["centers"]) for d in descs])
    assert np.array_equal(labels, g["labels"])
    out = O.vlad_encode(descs, g["centers"])
    assert out.dtype == np.float32 and out.shape == (4, 256 * 128)
    assert rel_l2(out, g["out"]) < 1e-6
    out_p = O.vlad_encode(descs, g["centers"], power_norm_weight=0.5, norm_order=1)
    assert rel_l2(out_p, g["out_pow05_l1"]) < 1e-6
    nf = O.vlad_encode(descs[:2], g["centers"], flatten=False)
    assert nf.shape == (2 * 256, 128)                       # quirk Q2
    assert rel_l2(nf, g["out_noflatten_first2"]) < 1e-6
    sim = O.similarity_score(O.vlad_encode(descs[2:3], g["centers"]), out)
    assert sim.dtype == np.float32 and np.allclose(sim, g["sim_0_vs_rest"], atol=1e-6)


def test_vlad_quirk_q1_empty_descriptor_set():
    g = load_golden("vlad_rootsift128")
    descs = split(g["desc"], g["offsets"])
    q1 = O.vlad_encode([descs[1], np.zeros((0, 128), np.float32), descs[2]], g["centers"])
    ref = load_golden("vlad_q1_empty")["out"]
    assert q1.shape == ref.shape == (256 * 128,) and not q1.any() and not ref.any()


def test_vlad_pca64_projection_and_encoding():
    g = load_golden("vlad_rootsift_pca64")
    p = load_weights("pca_k256_root_sift_f2")
    y = O.pca_transform(g["desc"], p["components"], p["mean"])
    assert y.dtype == np.float32
    assert rel_l2(y, g["projected"]) < 1e-6
    out = O.vlad_encode(split(g["desc"], g["offsets"]), g["centers"], pca=(p["components"], p["mean"]))
    assert rel_l2(out, g["out"]) < 1e-5


def test_vlad_vgg514():
    g = load_golden("vlad_vgg514")
    descs = split(g["desc"], g["offsets"])
    labels = np.concatenate([O.kmeans_predict(d, g["centers"]) for d in descs])
    assert np.array_equal(labels, g["labels"])
    out = O.vlad_encode(descs, g["centers"])
    assert out.shape == (2, 131584) and rel_l2(out, g["out"]) < 1e-6   # pipeline.ipynb:404


@pytest.mark.parametrize("case,gmm,pca", FV_CASES)
def test_fv_matches_reference(case, gmm, pca):
    g = load_golden(case)
    w = load_weights(gmm)
    p = load_weights(pca) if pca else None
    pca_pair = (p["components"], p["mean"]) if p else None
    descs = split(g["desc"], g["offsets"])
    out = O.fv_encode(descs, *gmm_args(w), pca=pca_pair)
    d = w["means"].shape[1]
    assert out.dtype == np.float64 and out.shape == (len(descs), 2 * 256 * d + 256)
    assert rel_l2(out, g["out"]) < 1e-9
    assert np.allclose(np.linalg.norm(out, axis=1), 1.0, atol=1e-6)
    out2 = O.fv_encode(descs[:1], *gmm_args(w), pca=pca_pair, power_norm_weight=1.0, norm_order=1)
    assert rel_l2(out2, g["out_pow1_l1_img0"]) < 1e-9
    if "posterior_img0" in g:
        y0 = O.pca_transform(descs[0].astype(np.float32), *pca_pair) if p else descs[0]
        q = O.gmm_predict_proba(y0, w["weights"], w["means"], w["precisions_cholesky"])
        assert np.abs(q - g["posterior_img0"]).max() < 1e-10
        assert np.array_equal(q.argmax(1), g["posterior_img0"].argmax(1))


def test_fv_shape_formulas_from_notebooks():
    # getting_started.ipynb:424,448 and pipeline.ipynb:214,404
    assert 32 * 64 == 2048 and 2 * 32 * 64 + 32 == 4128
    assert 2 * 256 * 257 + 256 == 131840 and 256 * 514 == 131584
    assert 131840 + 131584 == 263424


def test_pipeline_is_hstack_and_retrieval():
    g = load_golden("pipeline_rootsift")
    w = load_weights("gmm_k256_root_sift_pca")
    p = load_weights("pca_k256_root_sift_f2")
    descs = split(g["desc"], g["offsets"])
    v = O.vlad_encode(descs, g["centers"])
    f = O.fv_encode(descs, *gmm_args(w), pca=(p["components"], p["mean"]))
    out = O.pipeline_encode([v, f])
    assert out.dtype == np.float64 and out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) < 1e-7
    # pipeline.ipynb:305 vs :336 -- Pipeline.similarity_score == cosine of the hstack
    sim = O.similarity_score(out[:2], out)
    assert np.allclose(sim, g["sim_first2_vs_all"], atol=1e-6)
    scores, idx = O.cosine_topk(out[3:4], out, 4)
    assert np.array_equal(idx[0], g["top4_names"])
    assert np.allclose(scores[0], g["top4_scores"], atol=1e-9)


def test_cosine_dtype_rules_zero_rows_and_topk():
    g = load_golden("cosine_small")
    s32 = O.cosine_similarity(g["a"], g["b"])
    s64 = O.cosine_similarity(g["a"].astype(np.float64), g["b"])
    assert s32.dtype == np.float32 and s64.dtype == np.float64
    assert np.allclose(s32, g["s32"], atol=1e-6) and np.allclose(s64, g["s64"], atol=1e-12)
    assert not s32[2].any()                                        # zero row stays zero
    assert np.array_equal(np.stack([O.topk_indices(r, 5) for r in g["s32"]]), g["top5"])
    with pytest.raises(ValueError):
        O.cosine_similarity(np.ones((3, 1)), np.ones((2, 1)))
    assert O.cosine_similarity(np.ones(4), np.ones((2, 4))).shape == (1, 2)


def test_eval_label_logic():
    idx = np.array([[0, 1, 2], [3, 4, 5]])
    db = np.array([7, 8, 7, 1, 1, 2])
    assert O.top_k_accuracy_from_lists(idx, db, np.array([8, 9])) == 0.5
    # query0 label 7: relevant ranks 1,3 -> (1/1 + 2/3)/2 ; query1 label 2: rank 3 -> (1/3)/1
    assert np.isclose(O.top_k_map_from_lists(idx, db, np.array([7, 2])), ((1 + 2 / 3) / 2 + 1 / 3) / 2)

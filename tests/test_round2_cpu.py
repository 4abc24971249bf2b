"""CPU tests for the round-2 rows: the oracle's label logic and learn() restatements pinned on
fixtures written by the real reference (tests/golden/make_golden_r2.py), and the DeepConvFeature
mirror against the reference extractor's output (torch on the CPU: no product kernel involved)."""
import numpy as np
import pytest

import pvs_oracle as O
from conftest import load_golden


def test_oracle_label_logic_matches_reference_eval():
    g = load_golden("eval_labels")
    for k in (1, 10, 100, 1500, None):
        kk = g["db"].shape[0] if k is None else k
        _, idx = O.cosine_topk(g["q"], g["db"], kk)
        assert abs(O.top_k_map_from_lists(idx, g["db_labels"], g["q_labels"]) - float(g[f"map_k{k}"])) < 1e-12, k
    for k in (1, 5, 50):
        _, idx = O.cosine_topk(g["q"], g["db"], k)
        assert O.top_k_accuracy_from_lists(idx, g["db_labels"], g["q_labels"]) == float(g[f"acc_k{k}"])
    s, idx = O.cosine_topk(g["q"][3:4], g["db"], 7)
    assert np.array_equal(idx[0], g["top7_idx"]) and np.allclose(s[0], g["top7_scores"], atol=1e-6)


def test_oracle_kmeans_lloyd_matches_reference_learn():
    g = load_golden("learn_kmeans")
    c, labels, n_iter, inertia = O.kmeans_lloyd(g["x"], g["init"], 50, 1e-6)
    assert n_iter == int(g["n_iter"]) and np.array_equal(labels, g["labels"])
    assert np.abs(c - g["centers"]).max() <= 1e-5
    assert abs(inertia - float(g["inertia"])) <= 1e-5 * float(g["inertia"])


def test_oracle_gmm_em_matches_reference_learn():
    g = load_golden("learn_gmm")
    r = O.gmm_em_diag(g["x"], np.full(8, 1 / 8), g["means_init"], np.ones((8, 16)), 40, 1e-5)
    assert r["n_iter"] == int(g["n_iter"]) and r["converged"] == bool(g["converged"])
    for key in ("weights", "means", "covariances", "precisions_cholesky"):
        assert np.abs(r[key] - g[key]).max() <= 1e-10, key
    assert abs(r["lower_bound"] - float(g["lower_bound"])) <= 1e-10


def test_deepconv_feature_matches_reference_extractor_cpu():
    torch = pytest.importorskip("torch")
    tvm = pytest.importorskip("torchvision.models")
    from pyvisim_b200.features import DeepConvFeature
    g = load_golden("deepconv_vgg16")
    torch.manual_seed(0)
    model = tvm.vgg16(weights=None)                           # the fixture's weights: same seed, same torch build
    ext = DeepConvFeature(model=model, device="cpu")
    assert ext.output_dim == 514 and ext.selected_layer_name == "features.28"
    desc, offs = ext.extract_batch([g["img0"], g["img1"]])    # different sizes: resized per image, one forward pass
    assert desc.shape == (392, 514) and offs.tolist() == [0, 196, 392]
    d = desc.numpy()
    for i, key in enumerate(("desc0", "desc1")):
        blk = d[196 * i:196 * (i + 1)]
        assert np.array_equal(blk[:, 512:], g[key][:, 512:])                      # (x/W, y/H), raster order: exact
        assert np.linalg.norm(blk - g[key]) <= 1e-5 * np.linalg.norm(g[key])
    one = ext(g["img0"])
    assert one.shape == (196, 514) and np.linalg.norm(one - g["desc0"]) <= 1e-5 * np.linalg.norm(g["desc0"])
    ext2 = DeepConvFeature(model=model, device="cpu", layer_index=10, spatial_encoding=False)
    o2 = ext2(g["img0"])
    assert list(o2.shape) == g["layer10_shape"].tolist()
    assert np.linalg.norm(o2[::37] - g["layer10_rows"]) <= 1e-5 * np.linalg.norm(g["layer10_rows"])
    with pytest.raises(IndexError):
        DeepConvFeature(model=model, device="cpu", layer_index=99)
    with pytest.raises(AttributeError):
        DeepConvFeature(model=model, device="cpu", target_submodule="nope")
    assert len(DeepConvFeature(model=model, device="cpu", target_submodule="features").list_conv_layers()) == 13


def test_learn_keywords_are_sklearns():
    from pyvisim_b200.encoders import _learn
    with pytest.raises(TypeError):
        _learn._kwargs(_learn._KMEANS_KW, {"n_jobs": 2}, "KMeans")
    assert _learn._kwargs(_learn._GMM_KW, {"tol": 1e-5}, "GaussianMixture")["reg_covar"] == 1e-6

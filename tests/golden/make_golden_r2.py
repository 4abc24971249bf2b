#!/usr/bin/env python
"""Round-2 golden vectors, again produced by running the REAL reference (pyvisim +
scikit-learn + torchvision) in the build container; same shims as ``make_golden.py``.

    python tests/golden/make_golden_r2.py         # writes eval_labels / deepconv / learn_* .npz

* ``eval_labels``   : ``pyvisim.eval.top_k_map`` / ``top_k_accuracy`` / ``retrieve_top_k_similar``
                      (eval.py:13-145) on a labelled set of random encodings, k = None / 1 / 10 / 1500.
* ``deepconv_vgg16``: ``pyvisim.features.DeepConvFeature`` (_features.py:150-306) with a VGG16 whose
                      weights come from ``torch.manual_seed(0)`` (the ImageNet file is not available
                      offline; the init is reproducible for a given torch build), two uint8 images.
* ``learn_kmeans`` / ``learn_gmm`` : ``VLADEncoder.learn`` / ``FisherVectorEncoder.learn``
                      (_base_encoder.py:311-342) with explicit initial parameters passed through
                      ``**kwargs`` to ``KMeans`` / ``GaussianMixture``.
"""
import os
import shutil
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    scratch = import_reference()
    warnings.simplefilter("ignore")
    import torch
    from pyvisim.encoders import VLADEncoder, FisherVectorEncoder
    from pyvisim._base_classes import FeatureExtractorBase
    from pyvisim.eval import retrieve_top_k_similar, top_k_map, top_k_accuracy
    from pyvisim.features import DeepConvFeature
    import torchvision.models as tvm

    class PassThrough(FeatureExtractorBase):
        def __init__(self, dim):
            super().__init__()
            self._d = dim

        def __call__(self, image):
            return image[0] if image.ndim == 3 else image

        @property
        def output_dim(self):
            return self._d

    # ---------------- eval label logic ----------------------------------------------------
    class RowEncoder:
        """encoder.encode(img) returns the 'image' itself (already an encoding)."""
        def encode(self, img):
            return np.asarray(img)

    rng = np.random.default_rng(11)
    n_db, n_q, d, n_cls = 1600, 23, 48, 7
    cls_dirs = rng.standard_normal((n_cls, d)).astype(np.float32)
    db_lab = rng.integers(0, n_cls, n_db)
    db = (cls_dirs[db_lab] * 0.6 + rng.standard_normal((n_db, d))).astype(np.float32)
    q_lab = rng.integers(0, n_cls, n_q)
    q = (cls_dirs[q_lab] * 0.6 + rng.standard_normal((n_q, d))).astype(np.float32)
    paths = [f"p{i:04d}" for i in range(n_db)]
    emap = {p: v for p, v in zip(paths, db)}
    plab = {p: int(l) for p, l in zip(paths, db_lab)}
    enc = RowEncoder()
    res = {}
    for k in (None, 1, 10, 100, 1500):
        res[f"map_k{k}"] = top_k_map(list(q), list(q_lab), emap, plab, enc, k=k)
    for k in (1, 5, 50):
        res[f"acc_k{k}"] = top_k_accuracy(list(q), list(q_lab), emap, plab, enc, k=k)
    top = retrieve_top_k_similar(q[3], emap, enc, k=7)
    save("eval_labels", db=db, db_labels=db_lab.astype(np.int64), q=q, q_labels=q_lab.astype(np.int64),
         top7_idx=np.array([int(p[1:]) for p, _ in top], np.int64), top7_scores=np.array([s for _, s in top], np.float32),
         **{k: np.float64(v) for k, v in res.items()})

    # ---------------- DeepConvFeature ------------------------------------------------------
    torch.manual_seed(0)
    model = tvm.vgg16(weights=None)                          # patched by import_reference: no download
    ext = DeepConvFeature(model=model, device="cpu")
    r = np.random.default_rng(5)
    imgs = [r.integers(0, 256, (96, 128, 3), dtype=np.uint8), r.integers(0, 256, (150, 100, 3), dtype=np.uint8)]
    with torch.no_grad():
        outs = [ext(im) for im in imgs]
    assert outs[0].shape == (196, 514)
    # a second layer / no coordinates: conv index 4, spatial_encoding False (56 x 56 map of 256 channels is large:
    # keep only a checksum row subset)
    ext2 = DeepConvFeature(model=model, device="cpu", layer_index=10, spatial_encoding=False)
    with torch.no_grad():
        o2 = ext2(imgs[0])
    save("deepconv_vgg16", img0=imgs[0], img1=imgs[1], desc0=outs[0].astype(np.float32), desc1=outs[1].astype(np.float32),
         layer10_shape=np.array(o2.shape), layer10_rows=o2[::37].astype(np.float32))

    # ---------------- learn() ---------------------------------------------------------------
    r = np.random.default_rng(3)
    k, dd = 8, 16
    centres = (r.standard_normal((k, dd)) * 1.2).astype(np.float32)
    images = []
    for t in (400, 650, 300, 500, 550, 600):
        lab = r.integers(0, k, t)
        images.append((centres[lab] + r.standard_normal((t, dd)) * (0.5 + 0.1 * lab[:, None])).astype(np.float32))
    X = np.vstack(images)
    init = X[r.choice(len(X), k, replace=False)].copy()         # classical 'random points' start: many Lloyd / EM iterations
    v = VLADEncoder(feature_extractor=PassThrough(dd))
    v.learn(images, n_clusters=k, init=init, n_init=1, max_iter=50, tol=1e-6, algorithm="lloyd")
    km = v.clustering_model
    save("learn_kmeans", x=X, offsets=np.concatenate([[0], np.cumsum([len(i) for i in images])]).astype(np.int64),
         init=init, centers=km.cluster_centers_, labels=km.labels_.astype(np.int32), n_iter=np.int64(km.n_iter_),
         inertia=np.float64(km.inertia_))
    f = FisherVectorEncoder(feature_extractor=PassThrough(dd))
    w0 = np.full(k, 1.0 / k)
    p0 = np.ones((k, dd))
    f.learn(images, n_clusters=k, means_init=init.astype(np.float64), weights_init=w0, precisions_init=p0, max_iter=40, tol=1e-5)
    g = f.clustering_model
    save("learn_gmm", x=X, means_init=init.astype(np.float64), weights=g.weights_, means=g.means_, covariances=g.covariances_,
         precisions_cholesky=g.precisions_cholesky_, n_iter=np.int64(g.n_iter_), lower_bound=np.float64(g.lower_bound_),
         converged=np.bool_(g.converged_))
    # learn with PCA (dim_reduction_factor=2): only shapes are pinned (sklearn's randomised SVD sign conventions aside)
    v2 = VLADEncoder(feature_extractor=PassThrough(dd))
    v2.learn(images, n_clusters=k, dim_reduction_factor=2, n_init=1, random_state=0)
    print("learn+pca:", v2.pca.components_.shape, v2.clustering_model.cluster_centers_.shape)

    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()

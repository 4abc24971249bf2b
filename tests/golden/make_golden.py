#!/usr/bin/env python
"""Generate golden vectors by running the REAL reference (pyvisim + scikit-learn).

Runs only in the build container (needs ``/root/reference``).  The reference ships no
tests or golden vectors, so parity is pinned on outputs of the
reference itself: this script copies ``/root/reference/pyvisim`` to a scratch directory
(importing it in place would write ``res/logs``, ``pyvisim/_config.py:9-11``), installs
the four shims of SURVEY.md section 8(c), feeds seeded synthetic descriptors through
``VLADEncoder.encode`` / ``FisherVectorEncoder.encode`` / ``Pipeline`` /
``similarity_score`` / ``cosine_similarity`` / ``retrieve_top_k_similar`` and stores
inputs + outputs as small ``.npz`` fixtures next to this file.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import os
import shutil
import sys
import tempfile
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    scratch = tempfile.mkdtemp(prefix="pyvisim_ref_")
    shutil.copytree(os.path.join(REF, "pyvisim"), os.path.join(scratch, "pyvisim"))
    sys.dont_write_bytecode = True
    for m in ("h5py", "matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import torchvision.models as tvm
    orig = tvm.vgg16
    tvm.vgg16 = lambda *a, **k: orig(weights=None)          # no download at import
    sys.path.insert(0, scratch)
    warnings.simplefilter("ignore")
    import pyvisim  # noqa: F401
    return scratch


# ---- synthetic descriptor generators ----------------------------
def rootsift_like(rng, t, d=128):
    a = np.abs(rng.standard_normal((t, d))).astype(np.float32)
    a /= (a.sum(axis=1, keepdims=True) + 1e-7)
    return np.sqrt(a)


def sift_like(rng, t, d=128):
    return np.floor(np.clip(np.abs(rng.normal(0, 40, (t, d))), 0, 255)).astype(np.float32)


def vgg_like(rng, t=196, c=512, side=14):
    f = rng.standard_normal((t, c)).astype(np.float32)
    ys, xs = np.divmod(np.arange(t), side)
    coords = np.stack([xs / side, ys / side], axis=1).astype(np.float32)
    return np.hstack([f, coords])


def main():
    scratch = import_reference()
    from pyvisim.encoders import VLADEncoder, FisherVectorEncoder, Pipeline, GMMWeights
    from pyvisim.encoders._base_encoder import _PCA
    from pyvisim._base_classes import FeatureExtractorBase
    from pyvisim._utils import cosine_similarity
    from pyvisim.eval import retrieve_top_k_similar
    from sklearn.cluster import KMeans

    class PassThrough(FeatureExtractorBase):
        """Returns its argument: the 'image' already is the (T, D) descriptor matrix."""
        def __init__(self, dim):
            super().__init__()
            self._d = dim

        def __call__(self, image):
            # a (1, T, D) array plays the role of "one H x W x 3 image" (vlad.py:85-86)
            return image[0] if image.ndim == 3 else image

        @property
        def output_dim(self):
            return self._d

    def random_kmeans(train, seed=0):
        km = KMeans(256, n_init=1, max_iter=1, random_state=seed)
        km.fit(train.astype(np.float32))
        assert km.cluster_centers_.dtype == np.float32
        return km

    def save(name, **arrays):
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")

    def pack(desc_list):
        offs = np.zeros(len(desc_list) + 1, np.int64)
        offs[1:] = np.cumsum([len(d) for d in desc_list])
        return np.vstack(desc_list), offs

    # ---------------- VLAD -------------------------------------------------------------
    rng = np.random.default_rng(0)
    # G1: RootSIFT-128, random-init K-Means (bundled k-means pickles are missing)
    km128 = random_kmeans(rootsift_like(rng, 4096))
    descs = [rootsift_like(rng, t) for t in (1, 37, 300, 517)]
    enc = VLADEncoder(feature_extractor=PassThrough(128), kmeans_model=km128)
    out = enc.encode(descs)
    labels = np.concatenate([km128.predict(d) for d in descs]).astype(np.int32)
    enc_p = VLADEncoder(feature_extractor=PassThrough(128), kmeans_model=km128,
                        power_norm_weight=0.5, norm_order=1)
    out_p = enc_p.encode(descs)
    enc_nf = VLADEncoder(feature_extractor=PassThrough(128), kmeans_model=km128, flatten=False)
    out_nf = enc_nf.encode(descs[:2])
    x, offs = pack(descs)
    save("vlad_rootsift128", desc=x, offsets=offs, centers=km128.cluster_centers_,
         labels=labels, out=out, out_pow05_l1=out_p, out_noflatten_first2=out_nf,
         sim_0_vs_rest=enc.similarity_score(descs[2:3], descs))
    # quirk Q1: an empty descriptor set aborts the batch with one 1-D zero vector
    q1 = enc.encode([descs[1], np.zeros((0, 128), np.float32), descs[2]])
    save("vlad_q1_empty", out=q1)

    # G2: RootSIFT-128 -> PCA-64 (bundled PCA), random-init K-Means on the projected space
    pca_rs = _PCA.OXFORD102_PCA256_ROOTSIFT.load()
    km64 = random_kmeans(pca_rs.transform(rootsift_like(rng, 4096)))
    descs = [rootsift_like(rng, t) for t in (64, 333)]
    enc = VLADEncoder(feature_extractor=PassThrough(128), kmeans_model=km64, pca=pca_rs)
    x, offs = pack(descs)
    save("vlad_rootsift_pca64", desc=x, offsets=offs, centers=km64.cluster_centers_,
         projected=pca_rs.transform(x.astype(np.float32)),
         out=enc.encode(descs))

    # G3: VGG16-514 (config C3 shape), T=196
    km514 = random_kmeans(vgg_like(rng, 4096 // 196 * 196 + 196 * 3)[:4096])
    descs = [vgg_like(rng), vgg_like(rng)]
    enc = VLADEncoder(feature_extractor=PassThrough(514), kmeans_model=km514)
    x, offs = pack(descs)
    save("vlad_vgg514", desc=x, offsets=offs, centers=km514.cluster_centers_,
         labels=np.concatenate([km514.predict(d) for d in descs]).astype(np.int32),
         out=enc.encode(descs))

    # ---------------- Fisher vectors ---------------------------------------------------
    def fv_case(name, weights_enum, gen, dim_in, ts, with_posterior=False, gmm_sampled=False):
        enc = FisherVectorEncoder(feature_extractor=PassThrough(dim_in), weights=weights_enum)
        gmm, pca = enc.clustering_model, enc.pca
        if gmm_sampled:
            # realistic posteriors: sample from the GMM and back-project through PCA^T
            descs = []
            for i, t in enumerate(ts):
                r = np.random.RandomState(100 + i)
                comp = r.choice(256, size=t, p=gmm.weights_ / gmm.weights_.sum())
                y = gmm.means_[comp] + r.standard_normal((t, gmm.means_.shape[1])) * np.sqrt(gmm.covariances_[comp])
                xs = y @ pca.components_ + pca.mean_ if pca is not None else y
                descs.append(xs.astype(np.float32))
        else:
            descs = [gen(rng, t) for t in ts]
        out = enc.encode(descs)
        x, offs = pack(descs)
        extra = {}
        if with_posterior:
            d0 = descs[0]
            y0 = pca.transform(d0.astype(np.float32)) if pca is not None else d0
            extra["posterior_img0"] = gmm.predict_proba(y0)
        enc2 = FisherVectorEncoder(feature_extractor=PassThrough(dim_in), weights=weights_enum,
                                   power_norm_weight=1.0, norm_order=1)
        extra["out_pow1_l1_img0"] = enc2.encode(descs[:1])
        save(name, desc=x, offsets=offs, out=out, **extra)

    fv_case("fv_sift_pca", GMMWeights.OXFORD102_K256_SIFT_PCA, sift_like, 128, (50, 300, 700), with_posterior=True)
    fv_case("fv_sift_pca_gmmsampled", GMMWeights.OXFORD102_K256_SIFT_PCA, None, 128, (128, 500), with_posterior=True, gmm_sampled=True)
    fv_case("fv_rootsift_pca", GMMWeights.OXFORD102_K256_ROOTSIFT_PCA, rootsift_like, 128, (1, 200))
    fv_case("fv_rootsift_nopca", GMMWeights.OXFORD102_K256_ROOTSIFT, rootsift_like, 128, (90, 260))
    fv_case("fv_sift_nopca", GMMWeights.OXFORD102_K256_SIFT, sift_like, 128, (90, 260))
    fv_case("fv_vgg_pca", GMMWeights.OXFORD102_K256_VGG16_PCA, lambda r, t: vgg_like(r), 514, (196, 196))
    fv_case("fv_vgg_pca_gmmsampled", GMMWeights.OXFORD102_K256_VGG16_PCA, None, 514, (196,), gmm_sampled=True)

    # ---------------- Pipeline + similarity + top-k -----------------------------------
    descs = [rootsift_like(rng, t) for t in (120, 80, 200, 150, 60, 90)]
    vlad = VLADEncoder(feature_extractor=PassThrough(128), kmeans_model=km128)
    fv = FisherVectorEncoder(feature_extractor=PassThrough(128), weights=GMMWeights.OXFORD102_K256_ROOTSIFT_PCA)
    pipe = Pipeline([vlad, fv])
    pout = pipe.encode(descs)
    assert pout.shape == (6, 256 * 128 + 2 * 256 * 64 + 256)
    sim = pipe.similarity_score(descs[:2], descs)
    x, offs = pack(descs)
    # retrieval through eval.retrieve_top_k_similar (query = image 3, db = all six)
    db = {f"img{i}": pout[i] for i in range(6)}
    top = retrieve_top_k_similar(descs[3][None], db, pipe, k=4)
    save("pipeline_rootsift", desc=x, offsets=offs, centers=km128.cluster_centers_, out=pout,
         sim_first2_vs_all=sim, top4_names=np.array([int(p[3:]) for p, _ in top], np.int64),
         top4_scores=np.array([s for _, s in top]))

    # cosine_similarity on plain matrices: fp32/fp32 -> fp32, mixed -> fp64, zero rows
    r = np.random.default_rng(7)
    a = r.standard_normal((9, 40)).astype(np.float32)
    b = r.standard_normal((13, 40)).astype(np.float32)
    a[2] = 0
    s32 = cosine_similarity(a, b)
    s64 = cosine_similarity(a.astype(np.float64), b)
    order = np.stack([np.argsort(-row, kind="stable")[:5] for row in s32])
    save("cosine_small", a=a, b=b, s32=s32, s64=s64, top5=order)

    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()

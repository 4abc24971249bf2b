"""CPU oracle for pyvisim's encode-and-compare hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-NumPy restatement of the arithmetic the reference executes for the
path SURVEY.md section 8 names.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import it.  Nothing under ``python-visual-similarity_b200/`` imports it and the product
path has no CPU fallback.

Where the arithmetic lives
--------------------------
The reference (pyvisim 0.1.3, pure Python) delegates the numerics to **scikit-learn**
(dependency *unpinned* in ``setup.py:19-35``; the bundled pickles were written by 1.5.1;
this image has 1.9.0) and NumPy/OpenBLAS.  Each function below restates the published
algorithm of the sklearn routine the reference calls and cites both the reference call
site and the sklearn source it follows.

Pinning status: **pinned against the reference itself**.  The reference ships no tests
and no golden vectors, so ``tests/golden/make_golden.py`` imports
the real ``pyvisim`` package (scratch copy of ``/root/reference`` + the real scikit-learn
estimators) in the build container, runs it on seeded descriptors and commits the
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function in
this file against those vectors (exact label / index equality, <=1e-6 relative on fp32
encodings, <=1e-12 on fp64 encodings).

No sklearn import here on purpose: the oracle must also run when only NumPy is present.
"""
from __future__ import annotations

import math
from typing import Iterable, Optional, Sequence

import numpy as np

__all__ = [
    "pca_transform", "kmeans_predict", "vlad_encode_one", "vlad_encode",
    "gmm_predict_proba", "fv_encode_one", "fv_encode", "pipeline_encode",
    "cosine_similarity", "topk_indices", "cosine_topk", "similarity_score",
    "top_k_accuracy_from_lists", "top_k_map_from_lists", "kmeans_lloyd", "gmm_em_diag",
]


# --------------------------------------------------------------------------------------
# a1  PCA projection
# --------------------------------------------------------------------------------------
def pca_transform(x: np.ndarray, components: np.ndarray, mean: np.ndarray) -> np.ndarray:
    """``PCA.transform`` with ``whiten=False``.

    Reference call sites: ``pyvisim/encoders/vlad.py:89-90``,
    ``pyvisim/encoders/fisher_vector.py:91-92`` (always on ``descriptors.astype(float32)``).
    sklearn: ``decomposition/_base.py:151-159``  --  ``X @ C.T`` then
    ``-= mean.reshape(1,-1) @ C.T`` (centring is applied *after* the projection).
    All three bundled PCAs are float32, so the result is float32.
    """
    x = np.asarray(x)
    y = x @ components.T
    y -= mean.reshape(1, -1) @ components.T
    return y


# --------------------------------------------------------------------------------------
# a2  hard assignment
# --------------------------------------------------------------------------------------
def kmeans_predict(x: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """``KMeans.predict`` labels.

    Reference call site: ``pyvisim/encoders/vlad.py:95``.
    sklearn: ``cluster/_kmeans.py`` ``_labels_inertia`` -> Cython
    ``_k_means_lloyd.pyx:_update_chunk_dense``: for every sample the score of centre j is
    ``||c_j||^2 - 2 <x, c_j>`` (``||x||^2`` is never added), evaluated as one ``sgemm``
    with ``alpha=-2, beta=1`` onto a matrix pre-filled with the squared centre norms, all
    in the centres' dtype (float32 here); the arg-min scan uses strict ``<`` so the
    **lowest index wins exact ties**.  ``np.argmin`` has the same tie rule.
    """
    x = np.asarray(x, dtype=centers.dtype)
    c2 = np.einsum("ij,ij->i", centers, centers)          # row_norms(centers, squared=True)
    score = c2[None, :] + np.dot(x, centers.T) * centers.dtype.type(-2.0)
    return np.argmin(score, axis=1).astype(np.int32)


def kmeans_scores(x: np.ndarray, centers: np.ndarray) -> np.ndarray:
    """float64 scores ``||c||^2 - 2 x.c`` used by the tests to *classify* label
    mismatches as near-ties (gap < 1e-6 relative) or real errors."""
    x = np.asarray(x, dtype=np.float64)
    c = np.asarray(centers, dtype=np.float64)
    return (c * c).sum(1)[None, :] - 2.0 * (x @ c.T)


# --------------------------------------------------------------------------------------
# a3 + a4  VLAD aggregation and normalisation
# --------------------------------------------------------------------------------------
def _signed_power(v: np.ndarray, p) -> np.ndarray:
    # vlad.py:106 / fisher_vector.py:127
    return np.sign(v) * np.abs(v) ** p


def vlad_encode_one(desc: np.ndarray, centers: np.ndarray, *, pca=None,
                    power_norm_weight=1, norm_order=2, epsilon=1e-9,
                    flatten=True, return_labels=False):
    """One image, following ``pyvisim/encoders/vlad.py:88-111`` line by line.

    ``pca`` is ``None`` or a ``(components, mean)`` pair.  Residuals are accumulated
    *sequentially in descriptor order* into a float32 ``(K, D)`` matrix
    (``vlad.py:98-104``); ``np.add.at`` performs the same unbuffered in-order float32 adds
    as the reference's Python loop.  Normalisation is the signed power followed by a
    **per-cluster** ``ord``-norm (``axis=1`` of the K x D matrix, ``vlad.py:107-108``).
    """
    desc = np.asarray(desc)
    if pca is not None:
        desc = pca_transform(desc.astype(np.float32), pca[0], pca[1])
    k, dim = centers.shape[0], desc.shape[1]
    if desc.shape[0] == 0:                                   # quirk Q1, vlad.py:92-93
        out = np.zeros(k * dim, dtype=np.float32)
        return (out, np.zeros(0, np.int32)) if return_labels else out
    x = desc.astype(np.float32)
    labels = kmeans_predict(x, centers)
    v = np.zeros((k, dim), dtype=np.float32)
    np.add.at(v, labels, desc - centers[labels])
    v = _signed_power(v, power_norm_weight)
    norms = np.linalg.norm(v, axis=1, ord=norm_order, keepdims=True) + epsilon
    v = v / norms
    if flatten:
        v = v.flatten()
    return (v, labels) if return_labels else v


def vlad_encode(desc_list: Iterable[np.ndarray], centers: np.ndarray, **kw) -> np.ndarray:
    """``VLADEncoder.encode`` over a list of per-image descriptor matrices
    (``vlad.py:81-115``), including quirk Q1: the first image with zero descriptors makes
    the whole call *return* a single 1-D zero vector."""
    rows = []
    for d in desc_list:
        d = np.asarray(d)
        if d.shape[0] == 0:
            dim = kw["pca"][0].shape[0] if kw.get("pca") is not None else d.shape[1]
            return np.zeros(centers.shape[0] * dim, dtype=np.float32)
        rows.append(vlad_encode_one(d, centers, **kw))
    return np.vstack(rows)


# --------------------------------------------------------------------------------------
# a5  GMM posteriors
# --------------------------------------------------------------------------------------
def _logsumexp_rows(a: np.ndarray) -> np.ndarray:
    # scipy.special.logsumexp(a, axis=1) as used by sklearn mixture/_base.py:573
    m = np.max(a, axis=1, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    with np.errstate(under="ignore"):
        s = np.sum(np.exp(a - m), axis=1)
    return np.log(s) + m[:, 0]


def gmm_weighted_log_prob(x, weights, means, precisions_cholesky) -> np.ndarray:
    """sklearn ``mixture/_gaussian_mixture.py:536-553`` (diag) + ``log(weights_)``
    (``mixture/_base.py:513-524``).  Note ``precisions = precisions_cholesky_**2`` (not
    ``1/covariances_``) and that ``X**2`` is formed in X's own dtype (float32 for every
    bundled configuration) *before* NumPy promotes the product to float64."""
    x = np.asarray(x)
    n_features = x.shape[1]
    precisions = precisions_cholesky ** 2
    log_det = np.sum(np.log(precisions_cholesky), axis=1)
    log_prob = (np.sum(means ** 2 * precisions, axis=1)
                - 2.0 * (x @ (means * precisions).T)
                + (x ** 2 @ precisions.T))
    return -0.5 * (n_features * math.log(2 * math.pi) + log_prob) + log_det + np.log(weights)


def gmm_predict_proba(x, weights, means, precisions_cholesky) -> np.ndarray:
    """``GaussianMixture.predict_proba`` (call site ``fisher_vector.py:99``; sklearn
    ``mixture/_base.py:414-432,552-582``): ``exp(wlp - logsumexp_k(wlp))``."""
    wlp = gmm_weighted_log_prob(x, weights, means, precisions_cholesky)
    with np.errstate(under="ignore"):
        return np.exp(wlp - _logsumexp_rows(wlp)[:, None])


# --------------------------------------------------------------------------------------
# a6 + a7 + a8  Fisher vector
# --------------------------------------------------------------------------------------
def fv_encode_one(desc, weights, means, covariances, precisions_cholesky, *, pca=None,
                  power_norm_weight=0.5, norm_order=2, epsilon=1e-9, flatten=True):
    """One image, following ``pyvisim/encoders/fisher_vector.py:90-133``.

    Output layout ``[d_pi (K) | d_mu (K*D, cluster-major) | d_sigma (K*D, cluster-major)]``,
    float64, signed power (default 0.5) then one *global* ``ord``-norm.
    No guard for T == 0 in the reference (division by zero -> NaN); kept.
    """
    desc = np.asarray(desc)
    if pca is not None:
        desc = pca_transform(desc.astype(np.float32), pca[0], pca[1])
    t = len(desc)
    q = gmm_predict_proba(desc, weights, means, precisions_cholesky)     # :99
    with np.errstate(all="ignore"):
        pp_sum = q.mean(axis=0, keepdims=True).T                          # :102
        pp_x = q.T.dot(desc) / t                                          # :103
        pp_x_2 = q.T.dot(np.power(desc, 2)) / t                           # :104
        d_pi = pp_sum.squeeze() - weights                                 # :107
        d_mu = pp_x - pp_sum * means                                      # :109
        d_sigma = (-pp_x_2 - pp_sum * np.power(means, 2)
                   + pp_sum * covariances + 2 * pp_x * means)             # :111-114
        sw = np.sqrt(weights)                                             # :117
        d_pi = d_pi / sw
        d_mu = d_mu / (sw[:, None] * np.sqrt(covariances))
        d_sigma = d_sigma / (np.sqrt(2) * sw[:, None] * covariances)
        v = np.hstack((d_pi, d_mu.ravel(), d_sigma.ravel())).reshape(1, -1)   # :123-124
        v = np.sign(v) * np.power(np.abs(v), power_norm_weight)           # :127
        norm = np.linalg.norm(v, axis=1, ord=norm_order, keepdims=True) + epsilon
        v = v / norm
    return v.flatten() if flatten else v


def fv_encode(desc_list, weights, means, covariances, precisions_cholesky, **kw) -> np.ndarray:
    """``FisherVectorEncoder.encode`` (``fisher_vector.py:83-135``)."""
    return np.vstack([fv_encode_one(d, weights, means, covariances, precisions_cholesky, **kw)
                      for d in desc_list])


# --------------------------------------------------------------------------------------
# a9  Pipeline
# --------------------------------------------------------------------------------------
def pipeline_encode(encodings: Sequence[np.ndarray]) -> np.ndarray:
    """``Pipeline.encode`` (``pipeline.py:59-66``): ``np.hstack`` of the flattened
    per-encoder outputs, no re-normalisation; float32 (+) float64 promotes to float64."""
    return np.hstack(list(encodings))


# --------------------------------------------------------------------------------------
# a10  cosine similarity
# --------------------------------------------------------------------------------------
def _normalize_rows(a: np.ndarray) -> np.ndarray:
    # sklearn preprocessing.normalize(norm="l2"): norms = sqrt(einsum('ij,ij->i'));
    # zero norms are replaced by 1 so zero rows stay zero.
    norms = np.sqrt(np.einsum("ij,ij->i", a, a))
    norms[norms == 0.0] = 1.0
    return a / norms[:, None]


def cosine_similarity(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """``pyvisim/_utils.py:312-330`` -> sklearn ``metrics/pairwise.py:1742-1750``.

    1-D inputs become one row; fewer than two features raises; dtype is float32 only if
    *both* inputs are float32 (``check_pairwise_arrays``), otherwise float64.
    """
    x = np.asarray(x)
    y = np.asarray(y)
    x = x.reshape(1, -1) if x.ndim == 1 else x
    y = y.reshape(1, -1) if y.ndim == 1 else y
    if x.shape[-1] <= 1 or y.shape[-1] <= 1:
        raise ValueError(
            f"Cosine similarity requires at least 2 features. Got {x.shape[-1]} features "
            f"for x and {y.shape[-1]} features for y.")
    dt = np.float32 if (x.dtype == np.float32 and y.dtype == np.float32) else np.float64
    x = x.astype(dt, copy=False)
    y = y.astype(dt, copy=False)
    return _normalize_rows(x) @ _normalize_rows(y).T


def similarity_score(v1: np.ndarray, v2: np.ndarray) -> np.ndarray:
    """``_base_encoder.py:371-385`` / ``pipeline.py:92-103``: ``np.float32(cos(v1, v2))``."""
    return np.float32(cosine_similarity(v1, v2))


# --------------------------------------------------------------------------------------
# a11  top-k
# --------------------------------------------------------------------------------------
def topk_indices(scores: np.ndarray, k: int) -> np.ndarray:
    """``np.argsort(-scores)[:k]`` (``eval.py:40-43,78-80,132``).  The reference's sort is
    introsort, so the order of *exactly equal* scores is unspecified; the oracle (and the
    CUDA path) define it: **lowest index first** (stable sort)."""
    return np.argsort(-scores, kind="stable")[:k]


def cosine_topk(queries: np.ndarray, database: np.ndarray, k: int):
    """Brute-force retrieval of ``eval.py:37-43`` for a batch of queries: cosine of every
    query against the database, then the k best per row.  Returns (scores, indices)."""
    s = cosine_similarity(queries, database)
    idx = np.stack([topk_indices(r, k) for r in s])
    return np.take_along_axis(s, idx, axis=1), idx.astype(np.int64)


# --------------------------------------------------------------------------------------
# f1  label logic that consumes the top-k lists (next row)
# --------------------------------------------------------------------------------------
def top_k_accuracy_from_lists(topk_idx: np.ndarray, db_labels: np.ndarray, query_labels: np.ndarray) -> float:
    """``eval.py:126-145``: a query is correct when any of its k retrieved items has its
    label; accuracy = correct / number of queries."""
    hit = (db_labels[topk_idx] == np.asarray(query_labels)[:, None]).any(axis=1)
    return float(hit.sum() / len(query_labels))


def top_k_map_from_lists(topk_idx: np.ndarray, db_labels: np.ndarray, query_labels: np.ndarray) -> float:
    """``eval.py:69-100`` including quirk Q6: R is the number of relevant items *within
    the truncated list*, AP = sum_{relevant ranks} (relevant_so_far / rank) / R."""
    aps = []
    for row, lbl in zip(topk_idx, query_labels):
        rel = db_labels[row] == lbl
        r = int(rel.sum())
        if r == 0:
            aps.append(0.0)
            continue
        ranks = np.arange(1, len(row) + 1)
        aps.append(float((np.cumsum(rel)[rel] / ranks[rel]).sum() / r))
    return float(np.mean(aps))


# --------------------------------------------------------------------------------------
# f4  learn(): K-Means Lloyd and diagonal-GMM EM with given initial parameters
# --------------------------------------------------------------------------------------
def kmeans_lloyd(x: np.ndarray, init: np.ndarray, max_iter: int = 300, tol: float = 1e-4):
    """``KMeans(init=array, n_init=1, algorithm="lloyd").fit`` as reached from
    ``_base_encoder.py:333-334,341`` (sklearn ``cluster/_kmeans.py``: ``fit`` centres X on its
    mean, scales ``tol`` by the mean feature variance, ``_kmeans_single_lloyd`` alternates
    arg-min labels (lowest index on ties) and member means until the labels repeat or the
    squared centre shift is <= tol, then re-labels once and adds the mean back).  Clusters that
    lose all members keep their centre here (sklearn relocates them to far points; the fixtures
    never trigger it).  Returns (centers, labels, n_iter, inertia) in x's dtype."""
    x = np.array(x, copy=True)
    dt = x.dtype
    mean = x.mean(axis=0)
    x -= mean
    centers = np.array(init, dtype=dt, copy=True) - mean
    tol_abs = np.mean(np.var(x, axis=0)) * tol
    labels_old = np.full(x.shape[0], -1, np.int32)
    strict = False
    n_iter = 0
    for it in range(max_iter):
        n_iter = it + 1
        labels = kmeans_predict(x, centers)
        new = centers.copy()
        for j in range(centers.shape[0]):
            m = labels == j
            if m.any():
                new[j] = x[m].sum(axis=0, dtype=dt) / dt.type(m.sum())
        shift = ((new - centers) ** 2).sum()
        centers = new
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if shift <= tol_abs:
            break
        labels_old = labels
    if not strict:
        labels = kmeans_predict(x, centers)
    inertia = float(((x - centers[labels]) ** 2).sum(dtype=np.float64))
    return centers + mean, labels.astype(np.int32), n_iter, inertia


def gmm_em_diag(x: np.ndarray, weights_init, means_init, precisions_init, max_iter: int = 100, tol: float = 1e-3,
                reg_covar: float = 1e-6):
    """``GaussianMixture(covariance_type="diag", weights_init=, means_init=, precisions_init=).fit``
    as reached from ``_base_encoder.py:335-341`` (sklearn ``mixture/_base.py:fit_predict`` and
    ``_gaussian_mixture.py``: E-step = normalised log responsibilities, M-step = nk (+10 eps),
    means, ``avg_X2 - means**2 + reg_covar``, weights / sum; stop when the mean log-likelihood
    changes by less than tol).  Returns a dict with the fitted parameters."""
    x = np.asarray(x)
    dt = np.result_type(x.dtype, np.asarray(means_init).dtype)   # float32 data with float64 initial parameters runs in float64
    w = np.asarray(weights_init, dtype=dt)
    mu = np.asarray(means_init, dtype=dt)
    pc = np.sqrt(np.asarray(precisions_init, dtype=dt))
    cov = None
    lower = -np.inf
    converged = False
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        prev = lower
        wlp = gmm_weighted_log_prob(x, w, mu, pc)
        lpn = _logsumexp_rows(wlp)
        with np.errstate(under="ignore"):
            resp = np.exp(wlp - lpn[:, None])
        nk = resp.sum(axis=0) + 10 * np.finfo(resp.dtype).eps
        mu = (resp.T @ x) / nk[:, None]
        cov = (resp.T @ (x * x)) / nk[:, None] - mu ** 2 + reg_covar
        w = nk / x.shape[0]
        w = w / w.sum()
        pc = 1.0 / np.sqrt(cov)
        lower = float(np.mean(lpn))
        if abs(lower - prev) < tol:
            converged = True
            break
    return {"weights": w, "means": mu, "covariances": cov, "precisions_cholesky": pc, "n_iter": n_iter,
            "lower_bound": lower, "converged": converged}

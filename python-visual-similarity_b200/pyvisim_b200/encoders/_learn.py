"""``learn()`` on the device (reference ``pyvisim/encoders/_base_encoder.py:311-342``).

The reference hands the stacked training descriptors to scikit-learn's ``KMeans`` (VLAD) or
``GaussianMixture(covariance_type="diag")`` (Fisher vectors) and forwards ``**kwargs``.  Here
the iterations run on the GPU through the C ABI: one ``pvs_kmeans_lloyd_step`` /
``pvs_gmm_em_step`` per iteration = one pass of the encode path's own assignment / posterior
kernels over the descriptors plus the M-step accumulators (fp64).  What stays on the host is what
is O(k * d) per iteration -- the parameter update and the convergence test, written after
scikit-learn's (``cluster/_kmeans.py:_kmeans_single_lloyd``, ``mixture/_base.py:fit_predict``,
``mixture/_gaussian_mixture.py:_estimate_gaussian_parameters``) -- and the one-off initialisation
(``kmeans_plusplus`` seeding draws from a NumPy ``RandomState`` exactly like scikit-learn, so a
given ``random_state`` starts from the same centres).  The keyword names and defaults are
scikit-learn's; unknown keywords raise ``TypeError`` as the estimators' constructors would.
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from .. import _native as N

_KMEANS_KW = {"init": "k-means++", "n_init": "auto", "max_iter": 300, "tol": 1e-4, "verbose": 0,
              "random_state": None, "copy_x": True, "algorithm": "lloyd"}
_GMM_KW = {"tol": 1e-3, "reg_covar": 1e-6, "max_iter": 100, "n_init": 1, "init_params": "kmeans",
           "weights_init": None, "means_init": None, "precisions_init": None, "random_state": None,
           "warm_start": False, "verbose": 0, "verbose_interval": 10}


def _kwargs(defaults: dict, given: dict, who: str) -> dict:
    bad = set(given) - set(defaults)
    if bad:
        raise TypeError(f"{who}.__init__() got an unexpected keyword argument {sorted(bad)[0]!r}")
    return {**defaults, **given}


def _device():
    import torch
    if not torch.cuda.is_available():
        raise N.PvsError(-3, "learn() runs on the GPU: no CUDA device available (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream(dev):
    import torch
    return torch.cuda.current_stream(dev).cuda_stream


# ---------------------------------------------------------------------------------------------
# K-Means (Lloyd)
# ---------------------------------------------------------------------------------------------
def _lloyd_single(x_dev, x_host, init, max_iter, tol_abs, verbose):
    """One Lloyd run from ``init`` on centred data.  Returns (labels torch int32, inertia, centers, n_iter)."""
    import torch
    lib = N.lib()
    dev = x_dev.device
    rows, d = x_dev.shape
    k = init.shape[0]
    centers = np.ascontiguousarray(init, dtype=np.float32)
    labels = torch.empty((rows,), dtype=torch.int32, device=dev)
    labels_old = torch.full((rows,), -1, dtype=torch.int32, device=dev)
    sums = torch.empty((k, d), dtype=torch.float64, device=dev)
    counts = torch.empty((k,), dtype=torch.int64, device=dev)
    inertia = torch.empty((1,), dtype=torch.float64, device=dev)
    ws = None
    strict = False
    n_iter = 0

    def step(c, want_sums=True):
        nonlocal ws
        model = N.Model.kmeans(c)
        try:
            need = int(lib.pvs_kmeans_lloyd_workspace_bytes(model.handle, rows))
            if ws is None or ws.numel() < max(need, 1):
                ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
            N.check(lib.pvs_kmeans_lloyd_step(model.handle, x_dev.data_ptr(), rows, labels.data_ptr(), sums.data_ptr(),
                                              counts.data_ptr(), inertia.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
            torch.cuda.current_stream(dev).synchronize()       # the model block is freed below
        finally:
            model.close()

    for it in range(max_iter):
        n_iter = it + 1
        step(centers)
        cnt = counts.cpu().numpy()
        new = centers.copy()
        nz = cnt > 0
        new[nz] = (sums.cpu().numpy()[nz] / cnt[nz, None]).astype(np.float32)
        if not nz.all():
            # sklearn relocates empty clusters to the points farthest from their centres
            # (_k_means_common.pyx:_relocate_empty_clusters_dense); rare, so done on the host
            lab = labels.cpu().numpy()
            dist = ((x_host - centers[lab]) ** 2).sum(axis=1)
            empty = np.flatnonzero(~nz)
            far = np.argpartition(dist, -empty.size)[-empty.size:]
            s = sums.cpu().numpy().copy()
            c2 = cnt.astype(np.float64).copy()
            for e, f in zip(empty, far):
                old = lab[f]
                s[old] -= x_host[f]
                c2[old] -= 1
                s[e] = x_host[f]
                c2[e] = 1
            ok = c2 > 0
            new[ok] = (s[ok] / c2[ok, None]).astype(np.float32)
        shift = float(((new - centers) ** 2).sum(dtype=np.float64))
        centers = new
        if verbose:
            print(f"Iteration {it}, inertia {float(inertia.item())}.")
        if torch.equal(labels, labels_old):
            strict = True
            break
        if shift <= tol_abs:
            break
        labels_old.copy_(labels)
    step(centers)                                              # labels / inertia consistent with the final centres
    return labels, float(inertia.item()), centers, n_iter


def fit_kmeans(x: np.ndarray, n_clusters: int, **kwargs):
    """Device Lloyd with scikit-learn's ``KMeans`` keywords.  Returns a fitted ``sklearn.cluster.KMeans``
    object (attributes set from the device run) so ``clustering_model`` holds what the reference's holds."""
    import torch
    from sklearn.cluster import KMeans, kmeans_plusplus
    from sklearn.utils import check_random_state
    p = _kwargs(_KMEANS_KW, kwargs, "KMeans")
    x = np.ascontiguousarray(x)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    out_dtype = x.dtype
    if x.shape[0] < n_clusters:
        raise ValueError(f"n_samples={x.shape[0]} should be >= n_clusters={n_clusters}.")
    dev = _device()
    rs = check_random_state(p["random_state"])
    init = p["init"]
    init_is_array = not isinstance(init, str) and not callable(init)
    n_init = p["n_init"]
    if n_init == "auto":
        n_init = 1 if (init_is_array or init == "k-means++") else 10
    if init_is_array:
        init = np.array(init, dtype=np.float32)
        if init.shape != (n_clusters, x.shape[1]):
            raise ValueError(f"The shape of the initial centers {init.shape} does not match "
                             f"(n_clusters, n_features) = {(n_clusters, x.shape[1])}.")
        n_init = 1
    # sklearn centres X on its mean for the distance computations and scales tol by the mean variance
    xf = x.astype(np.float32, copy=True)
    mean = xf.mean(axis=0)
    tol_abs = float(np.mean(np.var(xf, axis=0)) * p["tol"])
    x_dev = torch.from_numpy(xf).to(dev)
    mean_dev = torch.from_numpy(mean).to(dev)
    with torch.cuda.device(dev):
        N.check(N.lib().pvs_rows_sub(x_dev.data_ptr(), x_dev.shape[0], x_dev.shape[1], mean_dev.data_ptr(), _stream(dev)))
    xc_host = xf - mean
    best = None
    with torch.cuda.device(dev):
        for _ in range(int(n_init)):
            if init_is_array:
                c0 = init - mean
            elif init == "k-means++":
                c0, _ = kmeans_plusplus(xc_host, n_clusters, random_state=rs)
            elif init == "random":
                c0 = xc_host[rs.permutation(x.shape[0])[:n_clusters]]
            else:
                raise ValueError(f"init should be 'k-means++', 'random' or an array, got {init!r}")
            labels, inertia, centers, n_iter = _lloyd_single(x_dev, xc_host, np.asarray(c0, np.float32), int(p["max_iter"]),
                                                             tol_abs, p["verbose"])
            if best is None or inertia < best[1]:
                best = (labels.cpu().numpy(), inertia, centers, n_iter)
    labels, inertia, centers, n_iter = best
    km = KMeans(n_clusters=n_clusters, **{k: v for k, v in p.items() if not (k == "init" and init_is_array)})
    if init_is_array:
        km.init = np.asarray(p["init"])
    km.cluster_centers_ = (centers + mean).astype(out_dtype)
    km.labels_ = labels.astype(np.int32)
    km.inertia_ = inertia
    km.n_iter_ = n_iter
    km.n_features_in_ = x.shape[1]
    km._n_features_out = n_clusters
    km._n_threads = 1
    if len(set(labels.tolist())) < n_clusters:
        from sklearn.exceptions import ConvergenceWarning
        warnings.warn(f"Number of distinct clusters ({len(set(labels.tolist()))}) found smaller than n_clusters "
                      f"({n_clusters}). Possibly due to duplicate points in X.", ConvergenceWarning, stacklevel=2)
    return km


# ---------------------------------------------------------------------------------------------
# diagonal GMM (EM)
# ---------------------------------------------------------------------------------------------
def _m_step(s0, s1, s2, n, reg_covar):
    nk = s0 + 10 * np.finfo(np.float64).eps
    means = s1 / nk[:, None]
    cov = s2 / nk[:, None] - means ** 2 + reg_covar
    w = nk / n
    return w / w.sum(), means, cov


def fit_gmm(x: np.ndarray, n_components: int, **kwargs):
    """Device EM with scikit-learn's ``GaussianMixture`` keywords (``covariance_type`` is "diag" as in the
    reference).  Returns a fitted ``sklearn.mixture.GaussianMixture`` object."""
    import torch
    from sklearn.mixture import GaussianMixture
    from sklearn.utils import check_random_state
    p = _kwargs(_GMM_KW, kwargs, "GaussianMixture")
    x = np.ascontiguousarray(x)
    if x.shape[0] < n_components:
        raise ValueError(f"Expected n_samples >= n_components but got n_components = {n_components}, "
                         f"n_samples = {x.shape[0]}")
    dev = _device()
    lib = N.lib()
    rs = check_random_state(p["random_state"])
    n, d = x.shape
    xf = x.astype(np.float32, copy=False)
    x_dev = torch.from_numpy(np.ascontiguousarray(xf)).to(dev)
    s0 = torch.empty((n_components,), dtype=torch.float64, device=dev)
    s1 = torch.empty((n_components, d), dtype=torch.float64, device=dev)
    s2 = torch.empty((n_components, d), dtype=torch.float64, device=dev)
    ll = torch.empty((1,), dtype=torch.float64, device=dev)
    ws = None
    reg = float(p["reg_covar"])

    def em_pass(w, mu, cov):
        nonlocal ws
        model = N.Model.gmm(w, mu, cov, 1.0 / np.sqrt(cov))
        try:
            need = int(lib.pvs_gmm_em_workspace_bytes(model.handle, n))
            if ws is None or ws.numel() < max(need, 1):
                ws = torch.empty((max(need, 1),), dtype=torch.uint8, device=dev)
            N.check(lib.pvs_gmm_em_step(model.handle, x_dev.data_ptr(), n, s0.data_ptr(), s1.data_ptr(), s2.data_ptr(),
                                        ll.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
            torch.cuda.current_stream(dev).synchronize()
        finally:
            model.close()
        return s0.cpu().numpy(), s1.cpu().numpy(), s2.cpu().numpy(), float(ll.item()) / n

    best = None
    with torch.cuda.device(dev):
        for _ in range(max(1, int(p["n_init"]))):
            # ---- initialisation (mixture/_base.py:_initialize_parameters) ----
            if p["means_init"] is not None and p["weights_init"] is not None and p["precisions_init"] is not None:
                resp_stats = None
            elif p["init_params"] == "kmeans":
                km = fit_kmeans(xf, n_components, n_init=1, random_state=rs)
                lab = km.labels_
                resp_stats = lab
            elif p["init_params"] in ("random_from_data", "k-means++", "random"):
                if p["init_params"] == "random":
                    resp = rs.uniform(size=(n, n_components))
                    lab = None
                    resp /= resp.sum(axis=1, keepdims=True)
                elif p["init_params"] == "random_from_data":
                    lab = None
                    resp = np.zeros((n, n_components))
                    resp[rs.choice(n, size=n_components, replace=False), np.arange(n_components)] = 1
                else:
                    from sklearn.cluster import kmeans_plusplus
                    lab = None
                    resp = np.zeros((n, n_components))
                    _, ind = kmeans_plusplus(xf, n_components, random_state=rs)
                    resp[ind, np.arange(n_components)] = 1
                resp_stats = resp
            else:
                raise ValueError(f"Unimplemented initialization method {p['init_params']!r}")
            if resp_stats is None:
                w = mu = cov = None
            else:
                x64 = xf.astype(np.float64)
                if isinstance(resp_stats, np.ndarray) and resp_stats.ndim == 1:
                    resp = np.zeros((n, n_components))
                    resp[np.arange(n), resp_stats] = 1
                else:
                    resp = resp_stats
                w, mu, cov = _m_step(resp.sum(axis=0), resp.T @ x64, resp.T @ (x64 * x64), n, reg)
            if p["weights_init"] is not None:
                w = np.asarray(p["weights_init"], np.float64)
            if p["means_init"] is not None:
                mu = np.asarray(p["means_init"], np.float64)
            if p["precisions_init"] is not None:
                cov = 1.0 / np.asarray(p["precisions_init"], np.float64)
            # ---- EM (mixture/_base.py:fit_predict) ----
            lower = -np.inf
            converged = False
            n_iter = 0
            bounds = []
            for n_iter in range(1, int(p["max_iter"]) + 1):
                prev = lower
                a0, a1, a2, lower = em_pass(w, mu, cov)
                w, mu, cov = _m_step(a0, a1, a2, n, reg)
                bounds.append(lower)
                if p["verbose"]:
                    print(f"  Iteration {n_iter}\t ll change {lower - prev:.5f}")
                if abs(lower - prev) < p["tol"]:
                    converged = True
                    break
            if best is None or lower > best["lower"]:
                best = {"w": w, "mu": mu, "cov": cov, "lower": lower, "n_iter": n_iter, "converged": converged,
                        "bounds": bounds}
    if not best["converged"] and int(p["max_iter"]) > 0:
        from sklearn.exceptions import ConvergenceWarning
        warnings.warn("Best performing initialization did not converge. Try different init parameters, or "
                      "increase max_iter, tol, or check for degenerate data.", ConvergenceWarning)
    g = GaussianMixture(n_components=n_components, covariance_type="diag", **p)
    g.weights_, g.means_, g.covariances_ = best["w"], best["mu"], best["cov"]
    g.precisions_cholesky_ = 1.0 / np.sqrt(best["cov"])
    g.precisions_ = g.precisions_cholesky_ ** 2
    g.converged_, g.n_iter_, g.lower_bound_, g.lower_bounds_ = best["converged"], best["n_iter"], best["lower"], best["bounds"]
    g.n_features_in_ = d
    return g

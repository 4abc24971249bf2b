#!/usr/bin/env python
"""One-time exporter: pyvisim's joblib pickles -> flat, versioned ``.npz`` weight files.

The reference loads fitted scikit-learn estimators from joblib pickles
(``pyvisim/encoders/_base_encoder.py:117-145``).  Pickles couple the serving path to the
scikit-learn version that wrote them (1.5.1) and cannot be read by the C side.  This tool
reads each pickle once (needs ``/root/reference`` and scikit-learn, so it only runs in the
build container) and writes the arrays the kernels need, with their original dtypes:

* GMM  : ``weights_`` (K,), ``means_`` (K,D), ``covariances_`` (K,D),
         ``precisions_cholesky_`` (K,D)            -- float64, diag covariance
* PCA  : ``components_`` (D,D_in), ``mean_`` (D_in,) -- float32, ``whiten`` False

Usage:  python tools/export_weights.py [/root/reference] [out_dir]
"""
import glob
import os
import sys
import warnings

import joblib
import numpy as np

FORMAT_VERSION = 1


def export(ref_root: str, out_dir: str) -> None:
    os.makedirs(out_dir, exist_ok=True)
    files = sorted(glob.glob(os.path.join(ref_root, "pyvisim/res/model_files/*.pkl")))
    if not files:
        raise SystemExit(f"no pickles under {ref_root}")
    for path in files:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = joblib.load(path)
        name = os.path.splitext(os.path.basename(path))[0]
        kind = type(model).__name__
        if kind == "GaussianMixture":
            assert model.covariance_type == "diag", model.covariance_type
            arrays = dict(
                kind="gmm_diag",
                weights=np.asarray(model.weights_),
                means=np.asarray(model.means_),
                covariances=np.asarray(model.covariances_),
                precisions_cholesky=np.asarray(model.precisions_cholesky_),
            )
        elif kind == "PCA":
            assert not model.whiten
            arrays = dict(
                kind="pca",
                components=np.asarray(model.components_),
                mean=np.asarray(model.mean_),
                explained_variance=np.asarray(model.explained_variance_),
            )
        elif kind == "KMeans":
            arrays = dict(kind="kmeans", cluster_centers=np.asarray(model.cluster_centers_))
        else:
            raise SystemExit(f"unexpected estimator {kind} in {path}")
        arrays["format_version"] = np.int64(FORMAT_VERSION)
        arrays["sklearn_version"] = str(getattr(model, "__getstate__", dict)().get("_sklearn_version", "?"))
        out = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(out, **arrays)
        print(f"{name}: {kind} -> {out} ({os.path.getsize(out)/1e6:.2f} MB)")


if __name__ == "__main__":
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    default_out = os.path.join(here, "..", "python-visual-similarity_b200", "pyvisim_b200", "res", "model_files")
    export(ref, sys.argv[2] if len(sys.argv) > 2 else default_out)

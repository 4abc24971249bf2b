"""Fisher-vector encoder on the B200 path (drop-in for ``pyvisim.encoders.FisherVectorEncoder``).

Reference behaviour: ``pyvisim/encoders/fisher_vector.py:41-135``.  Per image: optional
PCA, diagonal-GMM posteriors, 0th/1st/2nd-order statistics, gradients w.r.t. (pi, mu,
sigma) with the analytic normalisation, signed power (default 0.5), one global ``ord``-norm.
Output layout ``[d_pi (K) | d_mu (K*D) | d_sigma (K*D)]``.
"""
from __future__ import annotations

import warnings
from typing import Callable, Iterable

import numpy as np

from .. import _native as N
from .._base_classes import FeatureExtractorBase
from .._utils import cosine_similarity
from ..features import RootSIFT
from . import _device as D
from ._base_encoder import ImageEncoderBase


class FisherVectorEncoder(ImageEncoderBase):
    """``encode`` returns ``(N, 2*K*D + K)``.  The device computes in fp32; ``encode``
    returns float64 like the reference (``output_dtype`` can be set to ``np.float32`` to
    skip the widening copy)."""

    _native_kind = "gmm"

    def __init__(self, feature_extractor: FeatureExtractorBase = None, weights=None, gmm_model=None,
                 power_norm_weight: float = 0.5, norm_order: int = 2, epsilon: float = 1e-9, flatten: bool = True,
                 similarity_func: Callable = cosine_similarity, pca=None,
                 raise_error_when_pca_incompatible: bool = False, output_dtype=np.float64):
        from sklearn.mixture import GaussianMixture
        if feature_extractor is None:
            feature_extractor = RootSIFT()
        if gmm_model is not None:
            if not isinstance(gmm_model, GaussianMixture):
                raise ValueError(f"The clustering model must be an instance of GaussianMixture, not {type(gmm_model)}")
            gmm_model.covariance_type = "diag"
        if weights is not None and type(weights).__name__ != "GMMWeights":
            raise ValueError(f"You can only pass an instance of GMMWeights, not {type(weights).__name__}")
        self.output_dtype = np.dtype(output_dtype)
        super().__init__(feature_extractor, weights, gmm_model, similarity_func, power_norm_weight, norm_order,
                         epsilon, flatten, pca, raise_error_when_pca_incompatible)

    @property
    def clustering_model(self):
        return ImageEncoderBase.clustering_model.fget(self)

    @clustering_model.setter
    def clustering_model(self, model):
        from sklearn.mixture import GaussianMixture
        if not isinstance(model, GaussianMixture):
            raise ValueError(f"The clustering model must be an instance of GaussianMixture, not {type(model)}")
        if model.covariance_type != "diag":
            warnings.warn("Attribute 'covariance_type' of the clustering model is set to 'diag' because "
                          "training will take too long otherwise.")
            model.covariance_type = "diag"
        if np.ndim(model.covariances_) != 2:
            raise ValueError("only diagonal covariances (covariances_ of shape (K, D)) are supported")
        ImageEncoderBase.clustering_model.fset(self, model)

    @property
    def encoding_dim(self) -> int:
        k, d = self.clustering_model.means_.shape
        return int(2 * k * d + k)

    def encode(self, images: Iterable[np.ndarray] | np.ndarray) -> np.ndarray:
        on_dev = self._extract_on_device(images)
        if on_dev is not None:                                  # conv features stay on the device
            return self.encode_descriptors(*on_dev).cpu().numpy().astype(self.output_dtype, copy=False)
        descs = self._extract(images)
        for dsc in descs:
            if dsc is None or dsc.shape[0] == 0:
                # the reference has no guard: scikit-learn rejects the empty matrix
                raise ValueError(f"Found array with 0 sample(s) (shape={getattr(dsc, 'shape', None)}) while a "
                                 "minimum of 1 is required.")
        if not descs:
            raise ValueError("need at least one array to concatenate")
        out = self.encode_descriptors(descs)
        return out.astype(self.output_dtype, copy=False)       # flatten=False has the same 2-D shape

    def encode_descriptors(self, descriptors, offsets=None, *, out=None, chunk_rows: int = 0,
                           images_per_call: int = 0, n_streams: int = 2):
        """Bulk entry, see :meth:`VLADEncoder.encode_descriptors`.  Returns float32.

        ``images_per_call`` (device-resident input): images per library call; 0 = four per SM,
        which keeps the statistics kernel (one image per CTA at a time) evenly loaded.
        ``n_streams``: consecutive chunks alternate between this many side streams (see
        ``_device.run_device``); 1 = everything on the caller's stream."""
        if images_per_call <= 0:
            images_per_call = 4 * N.sm_count()
        cluster, pca = self._cluster_handle(), self._pca_handle()
        d_in = pca.d_in if pca else cluster.d
        x, offs, on_device = D.normalise_inputs(descriptors, offsets, d_in)
        D.check_offsets(offs, x.shape[0])
        n = offs.size - 1
        dim = 2 * cluster.k * cluster.d + cluster.k
        params = (float(self.power_norm_weight), float(self.norm_order), float(self.epsilon))
        if on_device:
            res, _ = D.run_device(N.lib().pvs_fv_encode, N.lib().pvs_fv_workspace_bytes, cluster, pca, x, offs, dim,
                                  params, images_per_call, False, out=out, n_streams=n_streams)
            return res
        if out is None:
            out = np.empty((n, dim), dtype=np.float32)
        elif out.dtype != np.float32 or out.shape != (n, dim) or not out.flags.c_contiguous:
            raise ValueError(f"out must be C-contiguous float32 of shape {(n, dim)}")
        host_fn = N.lib().pvs_fv_encode_host_u8 if x.dtype == np.uint8 else N.lib().pvs_fv_encode_host
        N.check(host_fn(cluster.handle, pca.handle if pca else None, x.ctypes.data,
                        offs.ctypes.data, n, *params, out.ctypes.data, int(chunk_rows)))
        return out

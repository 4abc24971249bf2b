"""VLAD encoder on the B200 path (drop-in for ``pyvisim.encoders.VLADEncoder``).

Reference behaviour: ``pyvisim/encoders/vlad.py:42-115``.  Per image: optional PCA, hard
assignment to the nearest of K centres, per-cluster sum of residuals (descriptor order),
signed power, per-cluster ``ord``-norm, flatten.  Here the whole batch is one packed
descriptor matrix and three kernels (projection, assignment, aggregation+normalisation).
"""
from __future__ import annotations

from typing import Callable, Iterable

import numpy as np

from .. import _native as N
from .._base_classes import FeatureExtractorBase
from .._utils import cosine_similarity
from ..features import RootSIFT
from . import _device as D
from ._base_encoder import ImageEncoderBase


class VLADEncoder(ImageEncoderBase):
    """Vector of Locally Aggregated Descriptors: ``encode`` returns float32 ``(N, K*D)``
    (or ``(N*K, D)`` with ``flatten=False``, quirk Q2)."""

    _native_kind = "kmeans"

    def __init__(self, feature_extractor: FeatureExtractorBase = None, weights=None, kmeans_model=None,
                 power_norm_weight: float = 1, norm_order: int = 2, epsilon: float = 1e-9, flatten: bool = True,
                 similarity_func: Callable = cosine_similarity, pca=None,
                 raise_error_when_pca_incompatible: bool = False) -> None:
        from sklearn.cluster import KMeans
        if feature_extractor is None:
            feature_extractor = RootSIFT()
        if kmeans_model is not None and not isinstance(kmeans_model, KMeans):
            raise ValueError(f"The clustering model must be an instance of KMeans, not {type(kmeans_model)}")
        if weights is not None and type(weights).__name__ != "KMeansWeights":
            raise ValueError(f"You can only pass an instance of KMeansWeights, not {type(weights).__name__}")
        super().__init__(feature_extractor, weights, kmeans_model, similarity_func, power_norm_weight, norm_order,
                         epsilon, flatten, pca, raise_error_when_pca_incompatible)

    @property
    def clustering_model(self):
        return ImageEncoderBase.clustering_model.fget(self)

    @clustering_model.setter
    def clustering_model(self, model):
        from sklearn.cluster import KMeans
        if not isinstance(model, KMeans):
            raise ValueError(f"The clustering model must be an instance of KMeans, not {type(model)}")
        ImageEncoderBase.clustering_model.fset(self, model)

    @property
    def encoding_dim(self) -> int:
        c = self.clustering_model.cluster_centers_
        return int(c.shape[0] * c.shape[1])

    def encode(self, images: Iterable[np.ndarray] | np.ndarray) -> np.ndarray:
        k, d = self.clustering_model.cluster_centers_.shape
        on_dev = self._extract_on_device(images)
        if on_dev is not None:                                  # conv features: never empty, stay on the device
            out = self.encode_descriptors(*on_dev).cpu().numpy()
            return out if self.flatten else out.reshape(out.shape[0] * k, d)
        descs = self._extract(images)
        for dsc in descs:
            if dsc is None or dsc.shape[0] == 0:
                if self.pca:
                    raise ValueError(f"Found array with 0 sample(s) (shape={dsc.shape}) while a minimum of 1 is "
                                     "required by PCA.")
                # quirk Q1 (vlad.py:92-93): the first empty image aborts the batch
                return np.zeros(k * dsc.shape[1], dtype=np.float32)
        if not descs:
            raise ValueError("need at least one array to concatenate")
        out = self.encode_descriptors(descs)
        return out if self.flatten else out.reshape(len(descs) * k, d)

    def encode_descriptors(self, descriptors, offsets=None, *, return_labels: bool = False, out=None,
                           chunk_rows: int = 0, images_per_call: int = 4096, n_streams: int = 2):
        """Bulk entry: descriptors of many images at once.

        ``descriptors``: list of ``(T_i, D_in)`` arrays, or a packed ``(sum T, D_in)`` NumPy
        array / CUDA tensor with int64 ``offsets`` (N+1).  Host inputs return NumPy
        float32 ``(N, K*D)`` (staged copies inside the call); a CUDA tensor returns a CUDA
        tensor and nothing crosses PCIe.  Images with no descriptors give zero rows here.
        """
        cluster, pca = self._cluster_handle(), self._pca_handle()
        d_in = pca.d_in if pca else cluster.d
        x, offs, on_device = D.normalise_inputs(descriptors, offsets, d_in)
        D.check_offsets(offs, x.shape[0])
        n = offs.size - 1
        dim = cluster.k * cluster.d
        params = (float(self.power_norm_weight), float(self.norm_order), float(self.epsilon))
        if on_device:
            res, labels = D.run_device(N.lib().pvs_vlad_encode, N.lib().pvs_vlad_workspace_bytes, cluster, pca, x,
                                       offs, dim, params, images_per_call, return_labels, out=out, n_streams=n_streams)
            return (res, labels) if return_labels else res
        if out is None:
            out = np.empty((n, dim), dtype=np.float32)
        elif out.dtype != np.float32 or out.shape != (n, dim) or not out.flags.c_contiguous:
            raise ValueError(f"out must be C-contiguous float32 of shape {(n, dim)}")
        labels = np.empty(x.shape[0], dtype=np.int32) if return_labels else None
        host_fn = N.lib().pvs_vlad_encode_host_u8 if x.dtype == np.uint8 else N.lib().pvs_vlad_encode_host
        N.check(host_fn(cluster.handle, pca.handle if pca else None, x.ctypes.data,
                        offs.ctypes.data, n, *params, out.ctypes.data,
                        labels.ctypes.data if return_labels else None, int(chunk_rows)))
        return (out, labels) if return_labels else out

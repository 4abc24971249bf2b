"""GPU-backed ``cosine_similarity`` (reference: ``pyvisim/_utils.py:312-330``)."""
from __future__ import annotations

import numpy as np

from . import _native as N
from ._errors import InvalidImageError

__all__ = ["cosine_similarity", "is_numpy_image"]


def is_numpy_image(image: np.ndarray, pos: int = 0) -> None:
    """Image validation the stock extractors run (reference ``_utils.py:34-53``)."""
    if image.ndim == 2:
        if not np.all(image == image.astype(np.int64)):
            raise InvalidImageError(f"Mask values must be integers. Got min={image.min()} and max={image.max()}.")
        return
    if image.shape[2] != 3:
        raise InvalidImageError(f"NumPy 3D images must have shape (H, W, 3). Got {image.shape}.")
    if image.min() < 0 or image.max() > 255:
        raise InvalidImageError(
            f"Image values must be in the range [0, 255]. Got min={image.min()} and max={image.max()} for position {pos}.")


def _as_matrix(a) -> np.ndarray:
    try:
        import torch
        if isinstance(a, torch.Tensor):
            a = a.detach().cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    a = np.asarray(a)
    return a.reshape(1, -1) if a.ndim == 1 else a


def cosine_similarity(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """N x M cosine similarity on the GPU (row-normalise, then one contraction).

    Same contract as the reference callable: NumPy (or torch) in, ``np.ndarray (N, M)``
    out, 1-D inputs become one row, fewer than two features raises ``ValueError``, and the
    result is float32 only if both inputs are float32 (otherwise float64, as
    scikit-learn's ``check_pairwise_arrays`` decides).  The arithmetic itself is fp32 on
    the device in both cases.
    """
    x, y = _as_matrix(x), _as_matrix(y)
    if x.shape[-1] <= 1 or y.shape[-1] <= 1:
        raise ValueError(
            f"Cosine similarity requires at least 2 features. Got {x.shape[-1]} features for x "
            f"and {y.shape[-1]} features for y.")
    if x.shape[1] != y.shape[1]:
        raise ValueError(f"Incompatible dimension for X and Y matrices: X.shape[1] == {x.shape[1]} "
                         f"while Y.shape[1] == {y.shape[1]}")
    out_dtype = np.float32 if (x.dtype == np.float32 and y.dtype == np.float32) else np.float64
    xf = np.ascontiguousarray(x, dtype=np.float32)
    yf = np.ascontiguousarray(y, dtype=np.float32)
    out = np.empty((xf.shape[0], yf.shape[0]), dtype=np.float32)
    N.check(N.lib().pvs_cosine_matrix_host(xf.ctypes.data, xf.shape[0], yf.ctypes.data, yf.shape[0],
                                           xf.shape[1], out.ctypes.data))
    return out if out_dtype == np.float32 else out.astype(np.float64)

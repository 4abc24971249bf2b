"""Local-descriptor producers (upstream of the hot path).

``SIFT`` / ``RootSIFT`` / ``Lambda`` keep the reference's names and contracts
(``pyvisim/features/_features.py:54-148``) and are thin host-side wrappers around OpenCV.
``Descriptors`` is new: a pass-through for callers that already hold ``(T, D)`` descriptor
matrices (synthetic benchmarks, cached features), so they can use ``encode()`` unchanged.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .._base_classes import FeatureExtractorBase
from .._utils import is_numpy_image


def _checked(extractor: FeatureExtractorBase, feats, image) -> np.ndarray:
    """Output contract of ``_check_output_shape`` (reference ``_features.py:24-51``)."""
    try:
        import torch
        if isinstance(image, torch.Tensor):
            raise TypeError("Currently, only Torch images are not supported yet. Please convert to NumPy.")
    except ImportError:  # pragma: no cover
        pass
    if feats is None:
        return np.zeros((0, extractor.output_dim), dtype=np.float32)
    if not isinstance(feats, np.ndarray):
        raise ValueError(f"Expected output to be a NumPy array, got {type(feats)} instead.")
    if feats.ndim != 2:
        raise ValueError(f"Feature extractor output must be 2D. Got shape {feats.shape}.")
    if feats.shape[1] != extractor.output_dim:
        raise ValueError(f"Expected feat_vecs.shape[1] == {extractor.output_dim}, but got {feats.shape[1]}.")
    return feats


class SIFT(FeatureExtractorBase):
    """OpenCV SIFT, 128-D float32 descriptors."""

    output_dim = 128

    def _detect(self, image):
        import cv2
        is_numpy_image(image, 0)
        return cv2.SIFT.create().detectAndCompute(image, None)[1]

    def __call__(self, image: np.ndarray, /) -> np.ndarray:
        return _checked(self, self._detect(image), image)

    def __repr__(self):
        return f"{type(self).__name__}(output_dim={self.output_dim})"


class RootSIFT(SIFT):
    """SIFT followed by the Hellinger map: L1-normalise (+1e-7) then square root
    (reference ``_features.py:113-114``)."""

    def __call__(self, image: np.ndarray, /) -> np.ndarray:
        d = self._detect(image)
        if d is not None:
            d = np.sqrt(d / (d.sum(axis=1, keepdims=True) + 1e-7))
        return _checked(self, d, image)


class Lambda(FeatureExtractorBase):
    """User function ``image -> (T, output_dim)`` descriptors."""

    def __init__(self, func: Callable, output_dim: int):
        super().__init__()
        if not callable(func):
            raise ValueError(f"Argument func must be a callable object, got {type(func)} instead")
        self.func = func
        self._output_dim = int(output_dim)

    @property
    def output_dim(self) -> int:
        return self._output_dim

    def __call__(self, image: np.ndarray, /) -> np.ndarray:
        is_numpy_image(image, 0)
        return _checked(self, self.func(image), image)


class Descriptors(FeatureExtractorBase):
    """Pass-through extractor: the "image" already is a ``(T, D)`` descriptor matrix
    (a ``(1, T, D)`` array is accepted as the single-image form)."""

    def __init__(self, output_dim: int):
        super().__init__()
        self._output_dim = int(output_dim)

    @property
    def output_dim(self) -> int:
        return self._output_dim

    def __call__(self, image: np.ndarray, /) -> np.ndarray:
        a = np.asarray(image)
        if a.ndim == 3 and a.shape[0] == 1:
            a = a[0]
        return _checked(self, a, image)


class DeepConvFeature(FeatureExtractorBase):
    """Conv-layer activations as local descriptors with (x/W, y/H) appended
    (reference ``_features.py:151-306``): C x H x W -> (H*W) x (C+2), raster order.

    Batched: ``extract_batch`` runs one forward pass for a stack of images and builds the
    ``(N*H*W, C+2)`` descriptor matrix on the device, ready for ``encode_descriptors``.
    """

    def __init__(self, model=None, layer_index: int = -1, spatial_encoding: bool = True, device=None):
        super().__init__()
        import torch
        self._torch = torch
        if model is None:
            from torchvision.models import vgg16
            model = vgg16(weights=None)
        self.model = model.eval()
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.model.to(self.device)
        convs = [m for m in self.model.modules() if isinstance(m, torch.nn.Conv2d)]
        if not convs:
            raise ValueError("model has no Conv2d layer to hook")
        self._layer = convs[layer_index]
        self.spatial_encoding = spatial_encoding
        self._feat = None
        self._layer.register_forward_hook(lambda mod, inp, out: setattr(self, "_feat", out))
        self._output_dim = self._layer.out_channels + (2 if spatial_encoding else 0)

    @property
    def output_dim(self) -> int:
        return self._output_dim

    def extract_batch(self, images: np.ndarray):
        """images: (N, H, W, 3) uint8/float in [0,255] -> torch (N*h*w, C+2) fp32 on device,
        plus offsets (N+1,) int64."""
        torch = self._torch
        x = torch.as_tensor(np.asarray(images), device=self.device).float().permute(0, 3, 1, 2) / 255.0
        mean = torch.tensor([0.485, 0.456, 0.406], device=self.device).view(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], device=self.device).view(1, 3, 1, 1)
        with torch.no_grad():
            self.model((x - mean) / std)
        f = self._feat                                             # N, C, h, w
        n, c, h, w = f.shape
        desc = f.permute(0, 2, 3, 1).reshape(n, h * w, c)
        if self.spatial_encoding:
            ys, xs = torch.meshgrid(torch.arange(h, device=self.device), torch.arange(w, device=self.device), indexing="ij")
            coords = torch.stack([xs.flatten() / w, ys.flatten() / h], dim=1).float()
            desc = torch.cat([desc, coords.unsqueeze(0).expand(n, -1, -1)], dim=2)
        offsets = torch.arange(n + 1, dtype=torch.int64) * (h * w)
        return desc.reshape(n * h * w, -1).contiguous(), offsets

    def __call__(self, image: np.ndarray, /) -> np.ndarray:
        is_numpy_image(image, 0)
        desc, _ = self.extract_batch(image[None])
        return _checked(self, desc.cpu().numpy(), image)

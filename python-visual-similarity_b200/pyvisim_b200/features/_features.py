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
    (reference ``_features.py:150-306``): C x H x W -> (H*W) x (C+2), raster order, hook on the
    Conv2d OUTPUT (pre-activation, ``:254-261``).

    Same constructor as the reference: ``model`` (default VGG16 with the ImageNet weights --
    if they cannot be loaded the constructor raises instead of silently running random weights,
    which would not match the bundled VGG16 vocabularies), ``target_submodule``, ``layer_index``,
    ``spatial_encoding``, ``device``, ``transform`` (default ``ToTensor`` + ``Resize((224, 224))``,
    no mean / std normalisation: ``:188-191``).

    New: ``extract_batch`` runs ONE forward pass for a stack of images and builds the
    ``(N*H*W, C+2)`` descriptor matrix on the device, ready for ``encode_descriptors`` -- the
    reference forwards one image at a time, copies the map to the host and appends the coordinates
    in a Python loop (``:263-300``).  With the default transform the resize runs batched on the
    device with the same kernel torchvision uses (bilinear, antialias).
    """

    def __init__(self, model=None, target_submodule: str = None, layer_index: int = -1, spatial_encoding: bool = True,
                 device=None, transform=None):
        super().__init__()
        import torch
        self._torch = torch
        if model is None:
            try:
                from torchvision.models import vgg16, VGG16_Weights
                model = vgg16(weights=VGG16_Weights.DEFAULT)
            except Exception as e:
                raise RuntimeError(
                    "DeepConvFeature: the default model is VGG16 with the ImageNet weights "
                    f"(reference _features.py:179) and they could not be loaded ({e}); pass model=... explicitly") from e
        if not isinstance(model, torch.nn.Module):
            raise TypeError(f"Currently, only torch.nn.Module is supported. Got {type(model)} instead.")
        self.model = model.eval()
        self.layer_index = layer_index
        self.spatial_encoding = spatial_encoding
        self.device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self.transform = transform                               # None = the reference default, applied batched
        self.model.to(self.device)
        if target_submodule is None:
            root = self.model
        elif hasattr(self.model, target_submodule):
            root = getattr(self.model, target_submodule)
        else:
            raise AttributeError(f"Model {self.model._get_name()} has no submodule named {target_submodule}.")
        self._conv_layers = [(i, n, m) for i, (n, m) in
                             enumerate((n, m) for n, m in root.named_modules() if isinstance(m, torch.nn.Conv2d))]
        if not self._conv_layers:
            raise ValueError(f"No convolutional layers found in model {self.model._get_name()}.")
        try:
            _, self.selected_layer_name, self.selected_layer_module = self._conv_layers[layer_index]
        except IndexError:
            info = "" if target_submodule is None else f" in submodule {root._get_name()}"
            raise IndexError(f"Model {self.model._get_name()} has only {len(self._conv_layers)} convolutional layers {info}"
                             f". Got layer_index={layer_index}.")
        self.buffer = None
        self.hook = self.selected_layer_module.register_forward_hook(
            lambda mod, inp, out: setattr(self, "buffer", out.detach()))
        self._output_dim = self.selected_layer_module.out_channels + (2 if spatial_encoding else 0)

    def list_conv_layers(self):
        return list(self._conv_layers)

    @property
    def output_dim(self) -> int:
        return self._output_dim

    def _to_batch(self, images):
        """-> float tensor (N, 3, 224, 224) on the device (default transform) or whatever the user
        transform produces, stacked."""
        torch = self._torch
        imgs = [images] if isinstance(images, np.ndarray) and images.ndim == 3 else list(images)
        if self.transform is not None:
            return torch.stack([self.transform(im) for im in imgs]).to(self.device)
        import torch.nn.functional as F
        out = []
        same = len({im.shape for im in imgs}) == 1 and len({im.dtype for im in imgs}) == 1
        groups = [imgs] if same else [[im] for im in imgs]
        for g in groups:
            a = np.stack(g)
            is_u8 = a.dtype == np.uint8
            x = torch.from_numpy(np.ascontiguousarray(a)).to(self.device).permute(0, 3, 1, 2)
            x = x.float().div(255) if is_u8 else x.float()      # ToTensor scales uint8 only
            # torchvision Resize on tensors: bilinear, antialias=True, align_corners=False
            out.append(F.interpolate(x, size=(224, 224), mode="bilinear", align_corners=False, antialias=True))
        return torch.cat(out)

    def extract_batch(self, images):
        """images: iterable / stack of H x W x 3 arrays -> (descriptors torch fp32 [N*h*w, C(+2)] on the
        device, offsets int64 [N+1]); image i owns rows ``offsets[i]:offsets[i+1]`` in raster order."""
        torch = self._torch
        x = self._to_batch(images)
        self.model.eval()
        self.buffer = None
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            self.model(x)                                        # only the hook's output matters
        if self.buffer is None:
            raise RuntimeError("Forward hook did not capture any features.")
        f = self.buffer.float()                                  # N, C, h, w
        n, c, h, w = f.shape
        desc = f.permute(0, 2, 3, 1).reshape(n, h * w, c)
        if self.spatial_encoding:
            # (x / Wf, y / Hf) in raster order, computed like the reference (:291-297): Python floats rounded once to
            # fp32 (a device-side divide by a scalar is a multiply by the rounded reciprocal and differs in the last bit)
            ys, xs = np.divmod(np.arange(h * w), w)
            coords = torch.from_numpy(np.stack([xs / w, ys / h], axis=1).astype(np.float32)).to(f.device)
            desc = torch.cat([desc, coords.unsqueeze(0).expand(n, -1, -1)], dim=2)
        offsets = torch.arange(n + 1, dtype=torch.int64) * (h * w)
        return desc.reshape(n * h * w, -1).contiguous(), offsets

    def __call__(self, image: np.ndarray, /) -> np.ndarray:
        is_numpy_image(image, 0)
        desc, _ = self.extract_batch(image)
        return _checked(self, desc.cpu().numpy(), image)

    def __repr__(self):
        return (f"DeepConvFeature(model={self.model._get_name()}, layer_index={self.layer_index}, "
                f"spatial_encoding={self.spatial_encoding}, device={self.device}, transform={self.transform}, "
                f"selected_layer_name={self.selected_layer_name}, selected_layer_module={self.selected_layer_module}, "
                f"output_dim={self.output_dim})")

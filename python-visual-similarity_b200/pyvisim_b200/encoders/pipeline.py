"""Concatenation of several encoders (drop-in for ``pyvisim.encoders.Pipeline``,
reference ``pyvisim/encoders/pipeline.py:29-103``)."""
from __future__ import annotations

import logging
from typing import Callable, Iterable

import numpy as np

from .._base_classes import SimilarityMetric
from .._utils import cosine_similarity
from ._base_encoder import ImageEncoderBase, check_desired_output


class Pipeline(SimilarityMetric):
    """Runs every encoder over the same images with ``flatten`` forced on and stacks the
    encodings side by side (no re-normalisation; float32 + float64 promotes to float64)."""

    _logger = logging.getLogger("Pipeline")

    def __init__(self, encoders: list[ImageEncoderBase], similarity_func: Callable = cosine_similarity):
        for enc in encoders:
            if not isinstance(enc, ImageEncoderBase):
                raise ValueError(f"Pipeline only accepts instances of ImageEncoderBase, not {type(enc)}")
        self.encoders = encoders
        self._similarity_func = similarity_func            # quirk Q5: not validated at construction

    @property
    def similarity_func(self):
        return self._similarity_func

    @similarity_func.setter
    def similarity_func(self, func: Callable):
        self._similarity_func = check_desired_output(func, np.random.rand(10, 10), np.random.rand(10, 10))

    def encode(self, images: Iterable[np.ndarray] | np.ndarray) -> np.ndarray:
        try:
            import torch
            if isinstance(images, torch.Tensor):
                raise RuntimeError("Torch images are not supported yet.")
        except ImportError:  # pragma: no cover
            pass
        if isinstance(images, np.ndarray) and images.ndim == 3:
            images = [images]
        images = list(images)              # the reference tees the iterable; every encoder sees all images
        parts = []
        for enc in self.encoders:
            saved, enc.flatten = enc.flatten, True
            try:
                parts.append(enc.encode(images))
            finally:
                enc.flatten = saved
        return np.hstack(parts)

    def encode_descriptors(self, per_encoder_descriptors, per_encoder_offsets=None):
        """Bulk form: one (descriptors, offsets) pair per encoder (they may use different
        extractors).  CUDA tensors in -> one CUDA tensor out (fp32)."""
        offs = per_encoder_offsets or [None] * len(self.encoders)
        parts = [e.encode_descriptors(x, o) for e, x, o in zip(self.encoders, per_encoder_descriptors, offs)]
        if isinstance(parts[0], np.ndarray):
            return np.hstack(parts)
        import torch
        return torch.cat(parts, dim=1)

    def generate_encoding_map(self, image_paths: Iterable[str]) -> dict[str, np.ndarray]:
        import cv2
        image_paths = list(image_paths)
        images = (cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB) for p in image_paths)
        return dict(zip(image_paths, self.encode(images)))

    def similarity_score(self, images1, images2):
        return np.float32(self.similarity_func(self.encode(images1), self.encode(images2)))

    def __repr__(self) -> str:
        inner = "\n".join(str(e) for e in self.encoders)
        name = getattr(self._similarity_func, "__name__", str(self._similarity_func))
        return f"Pipeline(\nencoders=[{inner}],\nsimilarity_func={name})"

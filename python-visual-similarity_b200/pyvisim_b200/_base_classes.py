"""Abstract interfaces kept from pyvisim so ``isinstance`` checks in user code hold
(reference: ``pyvisim/_base_classes.py:9-55``)."""
from __future__ import annotations

import abc
import logging

import numpy as np


class SimilarityMetric(abc.ABC):
    """Anything that can score the similarity of two (batches of) images."""

    _logger = logging.getLogger("Similarity_Metrics")

    @abc.abstractmethod
    def similarity_score(self, image1, image2):
        """Similarity of ``image1`` against ``image2`` (matrix for batches)."""


class FeatureExtractorBase(abc.ABC):
    """Maps one image (H x W x 3 ndarray) to a ``(T, output_dim)`` descriptor matrix."""

    _logger = logging.getLogger("Feature_Extractor")

    def __init__(self):
        pass

    @abc.abstractmethod
    def __call__(self, image: np.ndarray):
        ...

    @property
    @abc.abstractmethod
    def output_dim(self) -> int:
        """Descriptor dimensionality D (``shape[1]`` of what ``__call__`` returns)."""

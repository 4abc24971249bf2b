"""Retrieval evaluation (drop-in for ``pyvisim/eval.py``): same signatures and results,
but all queries are encoded in one batch, scored with one fused similarity / top-k pass and
the label logic (hit test, average precision with quirk Q6) runs on the device
(``pvs_topk_label_metrics``) instead of a Python loop with a full argsort per query."""
from __future__ import annotations

from typing import Iterable

import numpy as np

from . import _native as N
from ._utils import cosine_similarity  # noqa: F401  (re-exported like the reference's star import)

__all__ = ["retrieve_top_k_similar", "top_k_map", "top_k_accuracy", "topk_host"]


def _check_pair(queries, database):
    q = np.ascontiguousarray(queries, dtype=np.float32)
    db = np.ascontiguousarray(database, dtype=np.float32)
    q = q.reshape(1, -1) if q.ndim == 1 else q
    if q.shape[1] != db.shape[1]:
        raise ValueError("feature dimensions differ")
    if q.shape[1] <= 1:
        raise ValueError(f"Cosine similarity requires at least 2 features. Got {q.shape[1]}")
    return q, db


def topk_host(queries: np.ndarray, database: np.ndarray, k: int, use_bf16: bool = False):
    """(scores fp32 [nq,k], indices int64 [nq,k]) for raw (un-normalised) host matrices.
    Any ``0 <= k <= len(database)`` (fp32); beyond ``PVS_TOPK_MAX`` the library ranks in passes."""
    q, db = _check_pair(queries, database)
    k = max(0, min(int(k), db.shape[0]))
    scores = np.empty((q.shape[0], k), np.float32)
    idx = np.empty((q.shape[0], k), np.int64)
    if k == 0 or q.shape[0] == 0:
        return scores, idx
    N.check(N.lib().pvs_cosine_topk_host(q.ctypes.data, q.shape[0], db.ctypes.data, db.shape[0], q.shape[1], k,
                                         int(use_bf16), scores.ctypes.data, idx.ctypes.data))
    return scores, idx


def _as_list(images):
    if isinstance(images, np.ndarray) and images.ndim == 3:
        return [images]
    return list(images)


def retrieve_top_k_similar(uploaded_image: np.ndarray, dataset: dict[str, np.ndarray], encoder, k: int = 5
                           ) -> list[tuple[str, float]]:
    """Top-k ``(path, score)`` of ``dataset`` for one query image (``eval.py:13-46``)."""
    paths = list(dataset.keys())
    vectors = np.array(list(dataset.values()))
    q = encoder.encode(uploaded_image)
    q = q.reshape(1, -1) if q.ndim == 1 else q
    scores, idx = topk_host(q[:1], vectors, k)
    return [(paths[i], scores[0, j]) for j, i in enumerate(idx[0])]


def _device_metrics(images, image_labels, encoding_map, path_labels_dict, encoder, k):
    """Encode the queries, rank the database on the device and run the label kernel.
    Returns (hits int32 [nq], ap fp32 [nq]) as NumPy arrays, or None when the list is empty (k == 0)."""
    import torch
    from . import retrieval
    paths = list(encoding_map.keys())
    vectors = np.array(list(encoding_map.values()))
    q = encoder.encode(images)
    q, db = _check_pair(q, vectors)
    labels = list(image_labels)[:q.shape[0]]                  # zip() semantics of the reference loop
    q = q[:len(labels)]
    kk = len(paths) if k is None else max(0, min(int(k), len(paths)))
    if kk == 0 or q.shape[0] == 0:
        return None
    # labels may be any hashable: map them to dense int32 ids shared by queries and database
    ids: dict = {}
    db_ids = np.array([ids.setdefault(path_labels_dict[p], len(ids)) for p in paths], np.int32)
    q_ids = np.array([ids.setdefault(l, len(ids)) for l in labels], np.int32)
    dev = torch.device("cuda", torch.cuda.current_device())
    qn = retrieval.l2_normalize(torch.from_numpy(q).to(dev))
    dbn = retrieval.l2_normalize(torch.from_numpy(db).to(dev))
    _, idx = retrieval.cosine_topk(qn, dbn, kk)
    hits, ap = retrieval.label_metrics(idx, torch.from_numpy(db_ids), torch.from_numpy(q_ids))
    return hits.cpu().numpy(), ap.cpu().numpy()


def top_k_map(images: Iterable[np.ndarray], image_labels: Iterable[int], encoding_map: dict[str, np.ndarray],
              path_labels_dict: dict[str, int], encoder, k: int = None) -> float:
    """Mean average precision over the queries (``eval.py:49-100``).  ``k=None`` ranks the whole
    database like the reference; quirk Q6 is kept: the number of relevant items is counted
    inside the truncated list."""
    res = _device_metrics(_as_list(images), image_labels, encoding_map, path_labels_dict, encoder, k)
    if res is None:
        return 0.0                                           # an empty list has no relevant item: AP = 0
    return float(np.mean(res[1].astype(np.float64)))


def top_k_accuracy(images: Iterable[np.ndarray], image_labels: Iterable[int], encoding_map: dict[str, np.ndarray],
                   path_labels_dict: dict[str, int], encoder, k: int) -> float:
    """Fraction of queries with at least one same-label item among their k best
    (``eval.py:102-145``; quirk Q7: divides by ``len(images)``)."""
    images = _as_list(images)
    res = _device_metrics(images, image_labels, encoding_map, path_labels_dict, encoder, k)
    if res is None:
        return 0.0
    return float(int(res[0].sum()) / len(images))

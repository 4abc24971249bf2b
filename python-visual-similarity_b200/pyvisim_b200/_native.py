"""ctypes binding of ``libpvs_b200.so`` (the C ABI declared in ``include/pvs_b200.h``).

There is no CPU fallback: if the library cannot be loaded (or built) every entry point
raises :class:`NativeLibraryError`; if there is no CUDA device every compute call raises
:class:`PvsError` with the library's message.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _build

__all__ = ["lib", "check", "PvsError", "NativeLibraryError", "Model", "EXPORTS",
           "F32", "BF16", "F16X2", "PATH_AUTO", "PATH_SIMT", "PATH_TENSOR"]

F32, BF16, F16X2 = 0, 1, 2
PATH_AUTO, PATH_SIMT, PATH_TENSOR = 0, 1, 2
KIND_KMEANS, KIND_GMM, KIND_PCA = 1, 2, 3


class NativeLibraryError(RuntimeError):
    """libpvs_b200.so is missing and could not be built."""


class PvsError(RuntimeError):
    """A pvs_* call returned a negative status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[pvs status {status}] {message}")
        self.status = status


_i64, _i32, _f32, _vp, _sz = C.c_int64, C.c_int, C.c_float, C.c_void_p, C.c_size_t
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes): every symbol include/pvs_b200.h declares
EXPORTS = {
    "pvs_version": (_i32, []),
    "pvs_last_error": (C.c_char_p, []),
    "pvs_device_info": (_i32, [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_sz)]),
    "pvs_launch_count": (_i64, []),
    "pvs_launch_count_reset": (None, []),
    "pvs_set_path": (_i32, [_i32]),
    "pvs_profile_enable": (_i32, [_i32]),
    "pvs_profile_stage_count": (_i32, []),
    "pvs_profile_stage_name": (C.c_char_p, [_i32]),
    "pvs_profile_read": (_i32, [_i32, C.POINTER(C.c_double), C.POINTER(_i64)]),
    "pvs_kmeans_create": (_i32, [_vp, _i32, _i32, _pp]),
    "pvs_gmm_create": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _pp]),
    "pvs_pca_create": (_i32, [_vp, _vp, _i32, _i32, _pp]),
    "pvs_model_destroy": (_i32, [_vp]),
    "pvs_model_dims": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "pvs_pca_project": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "pvs_vlad_workspace_bytes": (_sz, [_vp, _vp, _i64, _i64]),
    "pvs_vlad_encode": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "pvs_fv_workspace_bytes": (_sz, [_vp, _vp, _i64, _i64]),
    "pvs_fv_encode": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _f32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "pvs_gmm_posterior": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "pvs_kmeans_assign_workspace_bytes": (_sz, [_vp, _i64]),
    "pvs_kmeans_assign": (_i32, [_vp, _vp, _i64, _vp, _vp, _sz, _vp]),
    "pvs_l2_normalize_rows": (_i32, [_vp, _i64, _i64, _vp, _i32, _vp]),
    "pvs_cosine_matrix_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "pvs_cosine_matrix": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _sz, _vp]),
    "pvs_cosine_topk_workspace_bytes": (_sz, [_i64, _i64, _i64, _i32, _i32]),
    "pvs_cosine_topk": (_i32, [_vp, _vp, _i32, _i64, _i64, _i64, _i32, _i64, _vp, _vp, _vp, _sz, _vp]),
    "pvs_cosine_topk_exact_stats": (_i32, [_vp, _i64, _i64, _i32, C.POINTER(_i64), C.POINTER(_i64), _vp]),
    "pvs_topk_merge": (_i32, [_vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp]),
    "pvs_topk_label_metrics": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "pvs_rows_sub": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "pvs_cluster_sums": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "pvs_kmeans_lloyd_workspace_bytes": (_sz, [_vp, _i64]),
    "pvs_kmeans_lloyd_step": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pvs_gmm_em_workspace_bytes": (_sz, [_vp, _i64]),
    "pvs_gmm_em_step": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pvs_nccl_load": (_i32, [C.c_char_p]),
    "pvs_comm_unique_id": (_i32, [_vp]),
    "pvs_comm_create": (_i32, [_vp, _i32, _i32, _pp]),
    "pvs_comm_destroy": (_i32, [_vp]),
    "pvs_allgather_topk": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "pvs_vlad_encode_host": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _vp, _vp, _i64]),
    "pvs_fv_encode_host": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _vp, _i64]),
    "pvs_vlad_encode_host_u8": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _vp, _vp, _i64]),
    "pvs_fv_encode_host_u8": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _vp, _i64]),
    "pvs_cosine_matrix_host": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp]),
    "pvs_cosine_topk_host": (_i32, [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _vp]),
    "pvs_debug_tc_gemm": (_i32, [_i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
}

_lock = threading.Lock()
_lib = None


def lib() -> C.CDLL:
    """Load (building first if the in-tree .so is missing or older than its sources)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        try:
            if _build.is_stale():
                _build.build()
        except Exception as e:  # no nvcc: use a prebuilt library if there is one
            if not os.path.exists(path):
                raise NativeLibraryError(
                    f"{path} is missing and could not be built ({e}); run "
                    "`python __graft_entry__.py build` -- there is no CPU fallback") from e
        try:
            handle = C.CDLL(path)
        except OSError as e:
            raise NativeLibraryError(f"cannot load {path}: {e}") from e
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise PvsError(status, lib().pvs_last_error().decode("utf-8", "replace"))


def ptr(a) -> int | None:
    """Device pointer of a torch tensor / host pointer of a NumPy array / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


class Model:
    """Owner of one ``pvs_model*`` (device-resident weights)."""

    def __init__(self, handle: int, kind: int, k: int, d: int, d_in: int):
        self.handle, self.kind, self.k, self.d, self.d_in = handle, kind, k, d, d_in

    @classmethod
    def _wrap(cls, out: C.c_void_p) -> "Model":
        kind, k, d, d_in = _i32(), _i32(), _i32(), _i32()
        check(lib().pvs_model_dims(out, C.byref(kind), C.byref(k), C.byref(d), C.byref(d_in)))
        return cls(out.value, kind.value, k.value, d.value, d_in.value)

    @classmethod
    def kmeans(cls, centers: np.ndarray) -> "Model":
        c = np.ascontiguousarray(centers, dtype=np.float32)
        out = C.c_void_p()
        check(lib().pvs_kmeans_create(c.ctypes.data, c.shape[0], c.shape[1], C.byref(out)))
        return cls._wrap(out)

    @classmethod
    def gmm(cls, weights, means, covariances, precisions_cholesky) -> "Model":
        w, m, v, p = (np.ascontiguousarray(a, dtype=np.float64) for a in
                      (weights, means, covariances, precisions_cholesky))
        if not (m.shape == v.shape == p.shape and m.ndim == 2 and w.shape == (m.shape[0],)):
            raise ValueError("GMM parameter shapes are inconsistent (diag covariance expected)")
        out = C.c_void_p()
        check(lib().pvs_gmm_create(w.ctypes.data, m.ctypes.data, v.ctypes.data, p.ctypes.data,
                                   m.shape[0], m.shape[1], C.byref(out)))
        return cls._wrap(out)

    @classmethod
    def pca(cls, components, mean) -> "Model":
        c = np.ascontiguousarray(components, dtype=np.float32)
        mu = np.ascontiguousarray(mean, dtype=np.float32)
        out = C.c_void_p()
        check(lib().pvs_pca_create(c.ctypes.data, mu.ctypes.data, c.shape[0], c.shape[1], C.byref(out)))
        return cls._wrap(out)

    def close(self) -> None:
        if self.handle:
            lib().pvs_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_sm_count_cache: dict = {}


def sm_count() -> int:
    """SM count of the current device (cached: the property query costs ~1 ms)."""
    import torch
    dev = torch.cuda.current_device()
    if dev not in _sm_count_cache:
        _sm_count_cache[dev] = device_info()["sm_count"]
    return _sm_count_cache[dev]


def device_info() -> dict:
    sm, major, minor, mem = _i32(), _i32(), _i32(), _sz()
    check(lib().pvs_device_info(C.byref(sm), C.byref(major), C.byref(minor), C.byref(mem)))
    return {"sm_count": sm.value, "cc": (major.value, minor.value), "total_mem": mem.value}


def profile_enable(on: bool) -> None:
    check(lib().pvs_profile_enable(int(on)))


def profile_read() -> dict:
    """{stage name: (total ms, launches)} accumulated since ``profile_enable(True)``."""
    out = {}
    for s in range(lib().pvs_profile_stage_count()):
        ms, n = C.c_double(), _i64()
        check(lib().pvs_profile_read(s, C.byref(ms), C.byref(n)))
        if n.value:
            out[lib().pvs_profile_stage_name(s).decode()] = (ms.value, n.value)
    return out


def set_path(path: int) -> None:
    """PATH_AUTO (tensor cores when the shape allows), PATH_SIMT, PATH_TENSOR (error if the
    shape is not supported by the tcgen05 kernels)."""
    check(lib().pvs_set_path(int(path)))

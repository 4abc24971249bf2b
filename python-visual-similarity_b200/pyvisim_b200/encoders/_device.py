"""Shared device/host dispatch for the two encoders' ``encode_descriptors``."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .. import _native as N
from ._base_encoder import pack_descriptors


def _is_torch(x) -> bool:
    try:
        import torch
        return isinstance(x, torch.Tensor)
    except ImportError:  # pragma: no cover
        return False


def normalise_inputs(descriptors, offsets, d_in: int):
    """-> (packed, offsets, on_device).  Accepts a list of (T_i, d_in) arrays, a packed
    NumPy matrix + offsets, or a packed CUDA torch tensor + offsets."""
    if _is_torch(descriptors):
        import torch
        if not descriptors.is_cuda:
            return normalise_inputs(descriptors.numpy(), None if offsets is None else np.asarray(offsets), d_in)
        x = descriptors.contiguous().float()
        if x.ndim != 2 or x.shape[1] != d_in:
            raise ValueError(f"descriptors must be (rows, {d_in}), got {tuple(x.shape)}")
        if offsets is None:
            offsets = torch.tensor([0, x.shape[0]], dtype=torch.int64)
        offs_host = offsets.detach().cpu().numpy().astype(np.int64) if _is_torch(offsets) else np.asarray(offsets, np.int64)
        return x, offs_host, True
    if isinstance(descriptors, np.ndarray) and descriptors.ndim == 2:
        if descriptors.shape[1] != d_in:
            raise ValueError(f"descriptors must be (rows, {d_in}), got {descriptors.shape}")
        # uint8 rows (integer-valued descriptors such as OpenCV SIFT's, as they are usually stored) keep their dtype: the
        # host entry points move them over PCIe as bytes and widen them on the device (bit-identical results)
        x = np.ascontiguousarray(descriptors) if descriptors.dtype == np.uint8 else np.ascontiguousarray(descriptors, dtype=np.float32)
        offs = np.array([0, x.shape[0]], np.int64) if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
        return x, offs, False
    descs = [np.asarray(d) for d in descriptors]
    for d in descs:
        if d.ndim != 2 or d.shape[1] != d_in:
            raise ValueError(f"every descriptor matrix must be (T, {d_in}), got {d.shape}")
    x, offs = pack_descriptors(descs, d_in)
    return x, offs, False


def check_offsets(offs: np.ndarray, rows: int) -> None:
    if offs.ndim != 1 or offs.size < 1 or offs[0] != 0 or offs[-1] != rows or np.any(np.diff(offs) < 0):
        raise ValueError("offsets must be non-decreasing int64 with offsets[0] == 0 and offsets[-1] == rows")


_side_streams: dict = {}


def _streams_for(dev, n: int):
    """Two (cached) side streams per device: consecutive image chunks alternate between them, so
    the ramp-down of one chunk's kernels overlaps the ramp-up of the next chunk's."""
    import torch
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), n)
    if key not in _side_streams:
        _side_streams[key] = [torch.cuda.Stream(dev) for _ in range(n)]
    return _side_streams[key]


def run_device(encode_fn, ws_fn, cluster: N.Model, pca: Optional[N.Model], x, offs_host: np.ndarray,
               out_dim: int, params, images_per_call: int, want_rows_i32: bool, out=None, n_streams: int = 2):
    """Device-resident path: loop over image chunks so the workspace stays bounded.  Chunks are
    independent; with more than one chunk they are issued round-robin on ``n_streams`` side
    streams (one workspace each) that fork from and join back into the caller's stream."""
    import torch
    n_images = offs_host.size - 1
    dev = x.device
    if out is None:
        out = torch.empty((n_images, out_dim), dtype=torch.float32, device=dev)
    elif (not isinstance(out, torch.Tensor) or out.device != dev or out.dtype != torch.float32
          or tuple(out.shape) != (n_images, out_dim) or not out.is_contiguous()):
        raise ValueError(f"out must be a contiguous float32 tensor of shape {(n_images, out_dim)} on {dev}")
    rows_i32 = torch.empty((x.shape[0],), dtype=torch.int32, device=dev) if want_rows_i32 else None
    # chunk-local CSR offsets of every chunk, built on the host and uploaded in ONE copy (a slice-and-subtract on the device
    # was one extra kernel launch per chunk)
    starts = list(range(0, n_images, images_per_call))
    local_host = np.concatenate([offs_host[i0:min(n_images, i0 + images_per_call) + 1] - offs_host[i0] for i0 in starts]) \
        if starts else np.zeros((1,), np.int64)
    local_dev = torch.as_tensor(np.ascontiguousarray(local_host, dtype=np.int64), device=dev)
    caller = torch.cuda.current_stream(dev)
    n_chunks = (n_images + images_per_call - 1) // images_per_call
    streams = [caller] if (n_chunks <= 1 or n_streams <= 1) else _streams_for(dev, n_streams)
    for s in streams:
        if s is not caller:
            s.wait_stream(caller)                       # inputs / offsets / out are ready on the caller's stream
    pca_h = pca.handle if pca else None
    ws = [None] * len(streams)
    power, order, eps = params
    lpos = 0
    with torch.cuda.device(dev):
        for c, i0 in enumerate(starts):
            si = c % len(streams)
            with torch.cuda.stream(streams[si]):
                i1 = min(n_images, i0 + images_per_call)
                r0, r1 = int(offs_host[i0]), int(offs_host[i1])
                need = ws_fn(cluster.handle, pca_h, r1 - r0, i1 - i0)
                if ws[si] is None or ws[si].numel() < need:
                    ws[si] = torch.empty((need,), dtype=torch.uint8, device=dev)
                local = local_dev[lpos:lpos + (i1 - i0) + 1]
                lpos += (i1 - i0) + 1
                N.check(encode_fn(cluster.handle, pca_h, x[r0:r1].data_ptr() if r1 > r0 else x.data_ptr(),
                                  local.data_ptr(), i1 - i0, r1 - r0, power, order, eps, out[i0:i1].data_ptr(),
                                  rows_i32[r0:].data_ptr() if (rows_i32 is not None and r1 > r0) else None,
                                  ws[si].data_ptr(), ws[si].numel(), streams[si].cuda_stream))
                del local
    for s in streams:
        if s is not caller:
            caller.wait_stream(s)
    return out, rows_i32

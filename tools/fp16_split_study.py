"""CPU emulation of the fp16 hi+lo ("fp16x2") operand split for the FV contractions.

kind::f16 MMAs run at twice the kind::tf32 rate; with x = hi + lo (both fp16, after a
power-of-two scale that keeps the values inside fp16's range) the product
hi*hi + hi*lo + lo*hi carries 22 mantissa bits -- the same class as 3xTF32 -- at half the
tensor time and half the shared-memory bytes.  This script measures the final FV rel-L2
error vs the fp64 reference for: 'f' exact fp32 products, '3' 3xTF32, 'h' fp16x2.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from tf32_split_study import split as split_tf32, mm, W  # noqa: E402


def split16(x):
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def gemm(a, b, mode):
    if mode == "f":
        return mm(a, b)
    if mode == "3":
        ah, al = split_tf32(a); bh, bl = split_tf32(b)
    else:
        ah, al = split16(a); bh, bl = split16(b)
        assert np.isfinite(ah).all() and np.isfinite(bh).all(), "fp16 overflow"
    return mm(ah, bh) + (mm(al, bh) + mm(ah, bl))


def model_exponent(mu, var):
    r = float(np.max(np.abs(mu) + 6.0 * np.sqrt(var)))
    return int(np.ceil(np.log2(r))) - 7          # r * 2^-e <= 128


def fv_emulated(descs, g, p, m_log, m_stat):
    w, mu, var, pc = g["weights"], g["means"], g["covariances"], g["precisions_cholesky"]
    k, d = mu.shape
    P = pc ** 2
    e = model_exponent(mu, var)
    sc = np.float32(2.0 ** -e)
    # interleaved (y^2, y) operand and matching weights, scaled by exact powers of two
    wq, wl = (-0.5 * P), (mu * P)
    cst = (-0.5 * (d * np.log(2 * np.pi) + (mu * mu * P).sum(1)) + np.log(pc).sum(1) + np.log(w)).astype(np.float32)
    outs, ymax = [], 0.0
    for x in descs:
        if p is not None:
            bias = -(p["mean"].reshape(1, -1) @ p["components"].T).astype(np.float32)
            y = mm(x, p["components"].T.astype(np.float32)) + bias
        else:
            y = x.astype(np.float32)
        ymax = max(ymax, float(np.abs(y).max()))
        ys = (y * sc).astype(np.float32)
        a = np.empty((len(y), 2 * d), np.float32); a[:, 0::2] = ys * ys; a[:, 1::2] = ys
        wc = np.empty((k, 2 * d), np.float32)
        wc[:, 0::2] = (wq * 2.0 ** (2 * e)).astype(np.float32); wc[:, 1::2] = (wl * 2.0 ** e).astype(np.float32)
        if m_log == "h":
            L = gemm(a, wc.T.copy(), "h") + cst
        else:
            a0 = np.empty_like(a); a0[:, 0::2] = y * y; a0[:, 1::2] = y
            w0 = np.empty_like(wc); w0[:, 0::2] = wq; w0[:, 1::2] = wl
            L = gemm(a0, w0.T.copy(), m_log) + cst
        mx = L.max(1, keepdims=True)
        ex = np.exp((L - mx).astype(np.float32))
        q = (ex / ex.sum(1, keepdims=True, dtype=np.float32)).astype(np.float32)
        if m_stat == "h":
            qs = q * np.float32(2.0 ** 14)
            z = np.concatenate([ys, ys * ys], axis=1).astype(np.float32)
            S = gemm(qs.T.copy(), z, "h")
            S[:, :d] *= np.float32(2.0 ** (e - 14)); S[:, d:] *= np.float32(2.0 ** (2 * e - 14))
            S = S / np.float32(len(x))
        else:
            z = np.concatenate([y, y * y], axis=1).astype(np.float32)
            S = gemm(q.T.copy(), z, m_stat) / np.float32(len(x))
        s0 = q.sum(0, dtype=np.float32) / np.float32(len(x))
        s1, s2 = S[:, :d], S[:, d:]
        mu32, var32, w32 = mu.astype(np.float32), var.astype(np.float32), w.astype(np.float32)
        sw = np.sqrt(w32)
        dpi = (s0 - w32) / sw
        dmu = (s1 - s0[:, None] * mu32) / (sw[:, None] * np.sqrt(var32))
        dsg = (-s2 - s0[:, None] * mu32 ** 2 + s0[:, None] * var32 + 2 * s1 * mu32) / (np.sqrt(np.float32(2)) * sw[:, None] * var32)
        v = np.hstack([dpi, dmu.ravel(), dsg.ravel()]).astype(np.float32)
        v = np.sign(v) * np.sqrt(np.abs(v))
        outs.append(v / (np.linalg.norm(v) + 1e-9))
    return np.vstack(outs), e, ymax


def main():
    cases = [
        ("fv_sift_pca", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
        ("fv_sift_pca_gmmsampled", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
        ("fv_rootsift_pca", "gmm_k256_root_sift_pca", "pca_k256_root_sift_f2"),
        ("fv_rootsift_nopca", "gmm_k256_root_sift_no_pca", None),
        ("fv_sift_nopca", "gmm_k256_sift_no_pca", None),
        ("fv_vgg_pca_gmmsampled", "gmm_k256_deep_features_vgg16_pca", "pca_k256_deep_features_vgg16_f2"),
    ]
    modes = [("f", "f"), ("3", "3"), ("3", "h"), ("h", "3"), ("h", "h")]
    print("case".ljust(26) + "".join(f"{'/'.join(m):>10}" for m in modes) + "   e  max|y|*2^-e")
    for case, gn, pn in cases:
        gold = dict(np.load(os.path.join(ROOT, "tests", "golden", case + ".npz")))
        g = dict(np.load(os.path.join(W, gn + ".npz")))
        p = dict(np.load(os.path.join(W, pn + ".npz"))) if pn else None
        offs = gold["offsets"]
        descs = [gold["desc"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
        ref = gold["out"]
        row = case.ljust(26)
        for m in modes:
            out, e, ymax = fv_emulated(descs, g, p, *m)
            err = np.linalg.norm(out.astype(np.float64) - ref) / np.linalg.norm(ref)
            row += f"{err:10.1e}"
        print(row + f"  {e:3d}  {ymax * 2.0 ** -e:8.2f}", flush=True)


if __name__ == "__main__":
    main()

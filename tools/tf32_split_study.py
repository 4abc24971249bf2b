"""CPU emulation of TF32 split strategies for the three FV contractions (design study).

Emulates tcgen05 kind::tf32 products (operands rounded to 10 mantissa bits, fp32
accumulate) and measures the final Fisher-vector rel-L2 error against the fp64 oracle for
different choices of which operand of which contraction is split into hi+lo parts.
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pvs_oracle as O

W = os.path.join(ROOT, "python-visual-similarity_b200", "pyvisim_b200", "res", "model_files")


def tf32(x):
    """round-to-nearest-even to 10 explicit mantissa bits"""
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32)


def tf32_trunc(x):
    x = np.ascontiguousarray(x, np.float32)
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = tf32(x)
    lo = tf32(x - hi)
    return hi, lo


def mm(a, b):
    """fp32-accumulated product of tf32-exact operands (emulated in fp64 then rounded)"""
    return (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32)


def gemm(a, b, mode):
    """a [m,k] @ b [k,n] with split mode: '1' = single pass, 'A' split a only (2 passes),
    'B' split b only, '3' = 3-pass (hi*hi + lo*hi + hi*lo), 'f' = exact fp32"""
    if mode == "f":
        return mm(a, b)
    ah, al = split(a)
    bh, bl = split(b)
    if mode == "1":
        return mm(ah, bh)
    if mode == "A":
        return mm(ah, bh) + mm(al, bh)
    if mode == "B":
        return mm(ah, bh) + mm(ah, bl)
    if mode == "3":
        return mm(ah, bh) + (mm(al, bh) + mm(ah, bl))
    raise ValueError(mode)


def fv_emulated(descs, g, p, m_pca, m_log, m_stat):
    w, mu, var, pc = g["weights"], g["means"], g["covariances"], g["precisions_cholesky"]
    k, d = mu.shape
    P = pc ** 2
    wcat = np.concatenate([-0.5 * P, mu * P], axis=1).astype(np.float32)          # [k, 2d]
    cst = (-0.5 * (d * np.log(2 * np.pi) + (mu * mu * P).sum(1)) + np.log(pc).sum(1) + np.log(w)).astype(np.float32)
    outs = []
    for x in descs:
        if p is not None:
            bias = -(p["mean"].reshape(1, -1) @ p["components"].T).astype(np.float32)
            y = gemm(x, p["components"].T, m_pca) + bias
        else:
            y = x.astype(np.float32)
        ycat = np.concatenate([y * y, y], axis=1).astype(np.float32)
        L = gemm(ycat, wcat.T, m_log) + cst
        mx = L.max(1, keepdims=True)
        e = np.exp((L - mx).astype(np.float32))
        q = (e / e.sum(1, keepdims=True, dtype=np.float32)).astype(np.float32)
        zcat = np.concatenate([y, y * y], axis=1).astype(np.float32)
        S = gemm(q.T.copy(), zcat, m_stat) / np.float32(len(x))
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
    return np.vstack(outs)


def main():
    cases = [
        ("fv_sift_pca", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
        ("fv_sift_pca_gmmsampled", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
        ("fv_rootsift_pca", "gmm_k256_root_sift_pca", "pca_k256_root_sift_f2"),
        ("fv_rootsift_nopca", "gmm_k256_root_sift_no_pca", None),
        ("fv_sift_nopca", "gmm_k256_sift_no_pca", None),
        ("fv_vgg_pca_gmmsampled", "gmm_k256_deep_features_vgg16_pca", "pca_k256_deep_features_vgg16_f2"),
    ]
    modes = [("f", "f", "f"), ("3", "3", "3"), ("3", "3", "A"), ("3", "3", "B"), ("3", "3", "1"),
             ("3", "A", "3"), ("3", "B", "3"), ("A", "3", "3"), ("B", "3", "3"), ("1", "3", "3"), ("3", "1", "3")]
    print("case".ljust(26) + "".join(f"{'/'.join(m):>10}" for m in modes))
    for case, gn, pn in cases:
        gold = dict(np.load(os.path.join(ROOT, "tests", "golden", case + ".npz")))
        g = dict(np.load(os.path.join(W, gn + ".npz")))
        p = dict(np.load(os.path.join(W, pn + ".npz"))) if pn else None
        offs = gold["offsets"]
        descs = [gold["desc"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
        ref = gold["out"]
        row = case.ljust(26)
        for m in modes:
            out = fv_emulated(descs, g, p, *m)
            err = np.linalg.norm(out.astype(np.float64) - ref) / np.linalg.norm(ref)
            row += f"{err:10.1e}"
        print(row, flush=True)


if __name__ == "__main__":
    main()

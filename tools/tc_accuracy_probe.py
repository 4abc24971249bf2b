"""GPU probe: how much Fisher-vector error do 3xTF32 tensor-core contractions introduce?

Runs the logits and the statistics contractions of the FV path through pvs_debug_tc_gemm
(3xTF32, the same MMA sequence the production kernels issue), does everything else in fp64
on the host, and reports the final FV rel-L2 against the reference golden vectors.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-visual-similarity_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pvs_oracle as O
from pyvisim_b200 import _native as N

W = os.path.join(ROOT, "python-visual-similarity_b200", "pyvisim_b200", "res", "model_files")


def tf32_round(x):
    u = x.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~0x1FFF
    u = torch.where(u >= 2 ** 31, u - 2 ** 32, u)
    return u.to(torch.int32).view(torch.float32)


def split(x):
    hi = tf32_round(x)
    return hi, tf32_round(x - hi)


def gemm(mode, a, b, m, n, k):
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    c = torch.empty((m, n), dtype=torch.float32, device="cuda")
    N.check(N.lib().pvs_debug_tc_gemm(mode, a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(),
                                      c.data_ptr(), m, n, k, 256, None))
    torch.cuda.synchronize()
    return c


def fv_from(descs, g, p, use_tc_logits, use_tc_stats):
    w, mu, var, pc = g["weights"], g["means"], g["covariances"], g["precisions_cholesky"]
    k, d = mu.shape
    P = pc ** 2
    wcat = np.concatenate([-0.5 * P, mu * P], axis=1)
    cst = -0.5 * (d * np.log(2 * np.pi) + (mu * mu * P).sum(1)) + np.log(pc).sum(1) + np.log(w)
    outs = []
    for x in descs:
        y = O.pca_transform(x.astype(np.float32), p["components"], p["mean"]) if p is not None else x.astype(np.float32)
        T = len(y)
        ycat = np.concatenate([y * y, y], axis=1).astype(np.float32)
        if use_tc_logits:
            Tp = (T + 127) // 128 * 128
            a = torch.zeros((Tp, 2 * d), device="cuda"); a[:T] = torch.from_numpy(ycat).cuda()
            L = gemm(1, a, torch.from_numpy(wcat.astype(np.float32)).cuda(), Tp, k, 2 * d)[:T].cpu().numpy().astype(np.float64) + cst
        else:
            L = ycat.astype(np.float64) @ wcat.T + cst
        L -= L.max(1, keepdims=True)
        q = np.exp(L); q /= q.sum(1, keepdims=True)
        if use_tc_stats:
            S = gemm(4, torch.from_numpy(ycat).cuda(), torch.from_numpy(q.astype(np.float32)).cuda(), 2 * d, k, T)
            S = S.cpu().numpy().astype(np.float64).T / T                     # [k, 2d] = [s2 | s1]
        else:
            S = q.T @ ycat.astype(np.float64) / T
        s2, s1, s0 = S[:, :d], S[:, d:], q.mean(0)
        sw = np.sqrt(w)
        dpi = (s0 - w) / sw
        dmu = (s1 - s0[:, None] * mu) / (sw[:, None] * np.sqrt(var))
        dsg = (-s2 - s0[:, None] * mu ** 2 + s0[:, None] * var + 2 * s1 * mu) / (np.sqrt(2) * sw[:, None] * var)
        v = np.hstack([dpi, dmu.ravel(), dsg.ravel()])
        v = np.sign(v) * np.sqrt(np.abs(v))
        outs.append(v / (np.linalg.norm(v) + 1e-9))
    return np.vstack(outs)


def main():
    cases = [("fv_sift_pca", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
             ("fv_sift_pca_gmmsampled", "gmm_k256_sift_pca", "pca_k256_sift_f2"),
             ("fv_rootsift_pca", "gmm_k256_root_sift_pca", "pca_k256_root_sift_f2")]
    print("case".ljust(26) + f"{'fp64':>10}{'tc-logits':>10}{'tc-stats':>10}{'tc-both':>10}")
    for case, gn, pn in cases:
        gold = dict(np.load(os.path.join(ROOT, "tests", "golden", case + ".npz")))
        g = dict(np.load(os.path.join(W, gn + ".npz")))
        p = dict(np.load(os.path.join(W, pn + ".npz")))
        offs = gold["offsets"]
        descs = [gold["desc"][offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
        row = case.ljust(26)
        for tl, ts in ((0, 0), (1, 0), (0, 1), (1, 1)):
            out = fv_from(descs, g, p, tl, ts)
            row += f"{np.linalg.norm(out - gold['out']) / np.linalg.norm(gold['out']):10.1e}"
        print(row, flush=True)


if __name__ == "__main__":
    main()

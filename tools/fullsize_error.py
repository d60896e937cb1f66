"""Achieved error of the surrogate VALUE S(x) at BASELINE's full size (cfg3: cubic RBF fitted on 4096 centers), the number
tests/test_gpu_fullsize.py bounds but does not report: the GPU evaluation (stand-alone kernel and the chain kernel's traced
log-posterior) and the float64 numpy oracle, each against an 80-bit long-double evaluation of the same expansion.
    python tools/fullsize_error.py > table.md     (needs a GPU)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from oracle import surrogate_rbf as ORB
from surrdamh_b200 import _lib as L


def main():
    w = bench.WORKLOADS["cfg3"]
    _, sur, _ = bench.build_models(w)
    SOL = sur.sol_numpy()
    coefs, centers = SOL[0], SOL[1]
    n, d = centers.shape
    X, _ = bench.training_set(w, 2048, 123)
    ld = np.longdouble
    cl, wl = centers.astype(ld), coefs.astype(ld)
    S_ref = np.empty(len(X), dtype=ld)
    S_abs = np.empty(len(X), dtype=ld)
    for i, x in enumerate(X):
        r2 = ((cl - x.astype(ld)) ** 2).sum(axis=1)
        phi = r2 * np.sqrt(r2)
        terms = wl[:n, 0] * phi
        tail = (wl[n:n + d, 0] * x.astype(ld)).sum() + wl[n + d, 0]
        S_ref[i] = terms.sum() + tail
        S_abs[i] = np.abs(terms).sum() + abs(tail)
    S_np = ORB.RbfApply(d, 1, kernel_type=1).apply(SOL, X)[:, 0]
    pts = L.dev(X)
    out = torch.empty((len(X), 1), dtype=torch.float64, device="cuda")
    L.check(L.lib().sdb_rbf_eval(L.ptr(pts), len(X), d, 1, C.byref(sur.struct), L.ptr(out), L.stream_handle()))
    S_gpu = out.cpu().numpy()[:, 0]
    cancel = (S_abs / np.abs(S_ref)).astype(float)

    def row(name, S):
        e = np.abs(S.astype(ld) - S_ref)
        rel = (e / np.abs(S_ref)).astype(float)
        sc = (e / S_abs).astype(float)
        return "| %s | %.2e | %.2e | %.2e | %.2e |" % (name, np.median(rel), rel.max(), np.median(sc) / 2.2e-16, sc.max() / 2.2e-16)

    print("cfg3 surrogate: %d centers, d = %d; %d prior-distributed points; cancellation sum|w phi| / |S|: median %.1e, max %.1e" %
          (n, d, len(X), np.median(cancel), cancel.max()))
    print()
    print("| evaluation | median |S - S_ref| / |S_ref| | max | median error in units of eps * sum|w phi| | max |")
    print("|---|---|---|---|---|")
    print(row("numpy float64 oracle (the reference's arithmetic)", S_np))
    print(row("GPU `sdb_rbf_eval`", S_gpu))
    print()
    print("GPU vs numpy directly: max |S_gpu - S_np| / |S_np| = %.2e" % float(np.max(np.abs(S_gpu - S_np) / np.abs(S_np))))


if __name__ == "__main__":
    main()

"""Stand-alone surrogate evaluation (Surrogate_apply.apply -- surrogate_rbf.py:20-40): time of
sdb_rbf_eval / sdb_rbf_eval_f32 for N points against n centers, as a fraction of the live FP64 / FP32 FMA peak.
    python tools/time_eval.py > table.md"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from surrdamh_b200 import _lib as L
from surrdamh_b200.device import DeviceModel


def main():
    dev = torch.device("cuda")
    p64, _ = bench.fp64_peaks(dev)
    p32 = bench.fp32_peak(dev)
    h = L.lib()
    print("| points | centers | d | m | FP64 ms | TFLOP/s | of the FP64 peak (%.1f) | FP32 ms | of the FP32 peak (%.1f) |" % (p64, p32))
    print("|---|---|---|---|---|---|---|---|---|")
    rs = np.random.RandomState(0)
    for N, n, d, m in ((1 << 20, 4096, 2, 1), (1 << 22, 4096, 2, 1), (1 << 20, 500, 4, 3), (1 << 20, 2000, 4, 3), (1 << 18, 16384, 2, 1)):
        centers = rs.standard_normal((n, d))
        coefs = rs.standard_normal((n + d + 1, m)) * 1e-3
        mo = DeviceModel.rbf(coefs, centers, 1)
        pts = torch.randn((N, d), dtype=torch.float64, device=dev)
        out = torch.empty((N, m), dtype=torch.float64, device=dev)
        flop = float(N) * (n * (3 * d + 3 + 2 * m) + m * (2 * d + 1))
        res = []
        for fn in (h.sdb_rbf_eval, h.sdb_rbf_eval_f32):
            try:
                for _ in range(2):
                    L.check(fn(L.ptr(pts), N, d, m, C.byref(mo.struct), L.ptr(out), L.stream_handle()))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    L.check(fn(L.ptr(pts), N, d, m, C.byref(mo.struct), L.ptr(out), L.stream_handle()))
                b.record()
                b.synchronize()
                res.append(a.elapsed_time(b) / 3)
            except RuntimeError:
                res.append(float("nan"))
        t64, t32 = res
        print("| %d | %d | %d | %d | %.3f | %.2f | %.3f | %.3f | %.3f |" % (N, n, d, m, t64, flop / t64 / 1e9, flop / t64 / 1e9 / p64, t32,
                                                                         flop / t32 / 1e9 / p32))


if __name__ == "__main__":
    main()

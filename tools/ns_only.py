"""Run the null-space fit alone (for ncu launch lists): python tools/ns_only.py [n] [reps] [d] [m]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surrdamh_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = int(sys.argv[3]) if len(sys.argv) > 3 else 2
m = int(sys.argv[4]) if len(sys.argv) > 4 else 1
rs = np.random.RandomState(0)
with torch.cuda.stream(torch.cuda.Stream()):
    X = L.dev(rs.standard_normal((n, d)) * 2 + 5)
    Y = L.dev(rs.standard_normal((n, m)))
    h = L.lib()
    CO = torch.empty((n + d + 1, m), dtype=torch.float64, device="cuda")
    wk = torch.empty((h.sdb_rbf_fit_nullspace_work(n, d, m),), dtype=torch.float64, device="cuda")
    info = torch.zeros((1,), dtype=torch.int32, device="cuda")
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.check(h.sdb_rbf_fit_nullspace(L.ptr(X), L.ptr(Y), n, d, m, 1, L.ptr(CO), L.ptr(wk), L.ptr(info), L.stream_handle()))
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    print("n=%d nullspace fit ms:" % n, ["%.3f" % t for t in ts], "info", int(info.item()))

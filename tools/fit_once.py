"""One null-space RBF fit (FIT_N points, d = 2, m = 1) inside cudaProfilerStart/Stop, for ncu launch lists:
    SDB_GRAPHS=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python tools/fit_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surrdamh_b200 import _lib as L
n, d, m = int(os.environ.get("FIT_N", "8192")), 2, 1
rs = np.random.RandomState(1)
Xn = rs.standard_normal((n, d)); Yn = np.sin(Xn).sum(axis=1, keepdims=True)
X, Y = L.dev(Xn), L.dev(Yn)
h = L.lib()
CO = torch.empty((n + d + 1, m), dtype=torch.float64, device="cuda")
info = torch.zeros((4,), dtype=torch.int32, device="cuda")
w1 = torch.empty((h.sdb_rbf_fit_nullspace_work(n, d, m),), dtype=torch.float64, device="cuda")
with torch.cuda.stream(torch.cuda.Stream()):
    for i in range(3):
        if i == 2: torch.cuda.profiler.start()
        L.check(h.sdb_rbf_fit_nullspace(L.ptr(X), L.ptr(Y), n, d, m, 1, L.ptr(CO), L.ptr(w1), L.ptr(info), L.stream_handle()))
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("info", int(info[0]))

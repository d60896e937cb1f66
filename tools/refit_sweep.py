"""BASELINE.json configs[3]: RBF surrogate refit sweep -- kernel-matrix build + dense solve on the GPU
(null-space Cholesky and LU with partial pivoting) against the oracle's np.linalg.solve on the host cores.
X ~ N(0, I_d), Y_k = sum_j sin((k+1) x_j), kernel_type 1, m = 3.   python tools/refit_sweep.py > table.md
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surrdamh_b200 import _lib as L


def ev_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def main():
    ns = [int(x) for x in os.environ.get("SWEEP_N", "1024,2048,4096,8192,16384,32768").split(",")]
    ds = [int(x) for x in os.environ.get("SWEEP_D", "4,8,16,32").split(",")]
    cpu_max = int(os.environ.get("SWEEP_CPU_MAX", "8192"))
    m = 3
    h = L.lib()
    print("| n | d | GPU null-space+Cholesky ms | GPU LU (partial pivoting) ms | backward error (ns / lu) | host np.linalg.solve s (%d threads) | speed-up vs host |" % len(os.sched_getaffinity(0)))
    print("|---|---|---|---|---|---|---|")
    for n in ns:
        for d in ds:
            rs = np.random.RandomState(n + d)
            Xn = rs.standard_normal((n, d))
            Yn = np.stack([np.sin((k + 1) * Xn).sum(axis=1) for k in range(m)], axis=1)
            X, Y = L.dev(Xn), L.dev(Yn)
            N = n + d + 1
            CO = torch.empty((N, m), dtype=torch.float64, device="cuda")
            info = torch.zeros((4,), dtype=torch.int32, device="cuda")
            w1 = torch.empty((h.sdb_rbf_fit_nullspace_work(n, d, m),), dtype=torch.float64, device="cuda")
            t_ns = ev_time(lambda: L.check(h.sdb_rbf_fit_nullspace(L.ptr(X), L.ptr(Y), n, d, m, 1, L.ptr(CO), L.ptr(w1), L.ptr(info), L.stream_handle())))
            inf_ns = int(info[0].item())
            co_ns = CO.cpu().numpy().copy()
            t_lu = ev_time(lambda: L.check(h.sdb_rbf_fit(L.ptr(X), L.ptr(Y), n, d, m, 1, L.ptr(CO), L.ptr(w1), L.ptr(info), L.stream_handle())))
            co_lu = CO.cpu().numpy().copy()
            del w1
            be_ns = be_lu = float("nan")
            t_cpu = float("nan")
            if n <= cpu_max:
                from oracle import surrogate_rbf as ORB
                up = ORB.RbfUpdate(d, m, kernel_type=1, solver_type="direct")
                up.add_arrays(Xn, Yn)
                t0 = time.perf_counter(); up.update(); t_cpu = time.perf_counter() - t0
                S, R = up.last_system, up.last_rhs
                nS = np.linalg.norm(S)
                be_ns = np.linalg.norm(S @ co_ns - R) / (nS * np.linalg.norm(co_ns))
                be_lu = np.linalg.norm(S @ co_lu - R) / (nS * np.linalg.norm(co_lu))
            print("| %d | %d | %.2f%s | %.2f | %.1e / %.1e | %.3f | %.0fx |" % (n, d, t_ns, "" if inf_ns == 0 else " (info %d)" % inf_ns, t_lu, be_ns, be_lu,
                                                                             t_cpu, (t_cpu * 1e3 / t_ns) if t_cpu == t_cpu else float("nan")), flush=True)


if __name__ == "__main__":
    with torch.cuda.stream(torch.cuda.Stream()):
        main()

"""Time the refit pieces on one GPU with CUDA events (after warm-up): thinning, system build, LU solve.
    python tools/time_refit.py [n ...]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surrdamh_b200 import _lib as L


def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    sizes = [int(x) for x in sys.argv[1:]] or [1024, 2048, 4096, 8192]
    d, m = 2, 1
    rs = np.random.RandomState(0)
    for n in sizes:
        X = L.dev(rs.standard_normal((n, d)) * 2 + 5)
        Y = L.dev(rs.standard_normal((n, m)))
        N = n + d + 1
        S = torch.empty((N, N), dtype=torch.float64, device="cuda")
        S0 = torch.empty((N, N), dtype=torch.float64, device="cuda")
        R = torch.empty((m, N), dtype=torch.float64, device="cuda")
        R0 = torch.empty((m, N), dtype=torch.float64, device="cuda")
        piv = torch.empty((N,), dtype=torch.int32, device="cuda")
        info = torch.zeros((1,), dtype=torch.int32, device="cuda")
        h = L.lib()
        t_sys = ev_time(lambda: L.check(h.sdb_rbf_system(L.ptr(X), L.ptr(Y), n, d, m, 1, L.ptr(S0), L.ptr(R0), L.stream_handle())))

        def lu():
            S.copy_(S0); R.copy_(R0)
            L.check(h.sdb_lu_solve(L.ptr(S), N, L.ptr(R), m, L.ptr(piv), L.ptr(info), L.stream_handle()))
        t_copy = ev_time(lambda: (S.copy_(S0), R.copy_(R0)))
        h.sdb_launch_count(1)
        t_lu = ev_time(lu)
        launches = h.sdb_launch_count(1) // 6
        x = R.cpu().numpy().T
        Sn, Rn = S0.cpu().numpy(), R0.cpu().numpy().T
        res = np.linalg.norm(Sn @ x - Rn) / (np.linalg.norm(Sn) * np.linalg.norm(x))
        # thinning: n + 256 -> n
        Xt = L.dev(rs.standard_normal((n + 256, d)) * 2 + 5)
        keep = torch.empty((n + 256,), dtype=torch.uint8, device="cuda")
        wk = torch.empty((h.sdb_rbf_thin_work(n + 256),), dtype=torch.float64, device="cuda")
        t_thin = ev_time(lambda: L.check(h.sdb_rbf_thin(L.ptr(Xt), n + 256, d, n, L.ptr(keep), L.ptr(wk), L.stream_handle())))
        # null-space + Cholesky
        CO = torch.empty((N, m), dtype=torch.float64, device="cuda")
        wk2 = torch.empty((h.sdb_rbf_fit_nullspace_work(n, d, m),), dtype=torch.float64, device="cuda")
        t_ns = ev_time(lambda: L.check(h.sdb_rbf_fit_nullspace(L.ptr(X), L.ptr(Y), n, d, m, 1, L.ptr(CO), L.ptr(wk2), L.ptr(info), L.stream_handle())))
        xns = CO.cpu().numpy()
        res_ns = np.linalg.norm(Sn @ xns - Rn) / (np.linalg.norm(Sn) * np.linalg.norm(xns))
        flops = 2.0 / 3.0 * N ** 3
        print("n=%6d  system %.3f ms  lu+solve %.3f ms (%.2f TF/s, %d launches, backward error %.1e)  nullspace-cholesky fit %.3f ms (info %d, backward error %.1e)  thin(+256) %.3f ms"
              % (n, t_sys[0], t_lu[0] - t_copy[0], flops / ((t_lu[0] - t_copy[0]) * 1e-3) / 1e12, launches, res, t_ns[0], int(info.item()), res_ns, t_thin[0]))


if __name__ == "__main__":
    # a non-default stream: the legacy default stream cannot be captured into the LU graph
    with torch.cuda.stream(torch.cuda.Stream()):
        main()

"""Cycle breakdown of the LU base kernel (library built with SDB_NVCC_EXTRA=-DSDB_LU_PROFILE)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from surrdamh_b200 import _lib as L
h = L.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4099
A = L.dev(np.random.RandomState(0).standard_normal((n, n)))
B = L.dev(np.random.RandomState(1).standard_normal((1, n)))
piv = torch.empty((n,), dtype=torch.int32, device="cuda"); info = torch.zeros((1,), dtype=torch.int32, device="cuda")
A0 = A.clone()
out = (C.c_longlong * 8)()
for rep in range(2):
    A.copy_(A0)
    h.sdb_lu_profile(None, 1)
    L.check(h.sdb_lu_solve(L.ptr(A), n, L.ptr(B), 1, L.ptr(piv), L.ptr(info), None))
    torch.cuda.synchronize()
    h.sdb_lu_profile(out, 0)
names = ["loop-top", "argmax+reduce", "publish", "cluster.sync", "pick+swap+eliminate"]
tot = sum(out[:5])
for k, nm in enumerate(names):
    print("%-22s %10.0f cycles/column  %5.1f%%" % (nm, out[k] / n, 100.0 * out[k] / tot))
print("total %.0f cycles/column" % (tot / n))

// Instantiations of the fused chain kernel (sdb_chain_kernel.cuh) with warp-owned chains (WT): a warp owns up to 32 chains and all its
// lanes split the centers of every evaluation (sdb_wt_points) -- the default launch shape for an RBF surrogate.  FP32 tables.
#include "sdb_chain_kernel.cuh"

typedef void (*chain_fn)(const sdb_chain_args, const int, const int);

chain_fn sdb_chain_pick_wt_f32(int d, int m) {
    if (d == 2 && m == 1) return chain_kernel<2, 1, 1, true>;
    if (d == 4 && m == 3) return chain_kernel<4, 3, 1, true>;
    return chain_kernel<0, 0, 1, true>;
}

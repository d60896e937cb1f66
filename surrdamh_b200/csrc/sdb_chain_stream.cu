// Instantiations of the fused chain kernel (sdb_chain_kernel.cuh) with the RBF surrogate streamed from L2 through a TMA ring (tables larger than shared memory);
// a translation unit of its own so that the three kernel families compile in parallel.
#include "sdb_chain_kernel.cuh"

typedef void (*chain_fn)(const sdb_chain_args, const int, const int);

chain_fn sdb_chain_pick_stream(int d, int m) {
    if (d == 2 && m == 1) return chain_kernel<2, 1, 2>;
    if (d == 4 && m == 3) return chain_kernel<4, 3, 2>;
    return chain_kernel<0, 0, 2>;
}

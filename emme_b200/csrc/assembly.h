// assembly.h -- host-side launcher interface of kernel 1 (assembly.cu).
#pragma once
#include <cuda_runtime.h>

#define EMME_MAX_PEERS 8

namespace emme {
struct RunConst;

// where kernel 1 stores its entries: p[0..n) are the same dim x dim matrix on n GPUs
struct PeerSet {
    double2* p[EMME_MAX_PEERS];
    int n;
};

// persistent grid size (SM count x resident CTAs) for the given Gauss-Kronrod order
int assembly_grid_blocks(int order, int device);
int assembly_groups_per_block(int order);
int assembly_stack_smem();

// Launch diagonal + quadrature kernels on `stream`.  `counter` (1 x u64), `stats` (8 x u64)
// and `spill` (grid_blocks*groups_per_block*spill_cap double2, may be null when
// spill_cap == 0) are device scratch owned by the handle.  refill_min: idle lanes of a warp
// refill together once at least this many are idle (1 = per lane, 32 = whole-warp batches).
// d_split: item order, diagonals d >= d_split first (0: d = 1, 2, ...; see decode_item).
cudaError_t launch_assembly(const RunConst& rc, const double* eta, const double* g,
                            const double* bi, const PeerSet& A, int shard_index, int shard_count,
                            unsigned long long* counter, void* spill, int spill_cap,
                            unsigned long long* stats, int grid_blocks, cudaStream_t stream,
                            unsigned long long* n_launches, int refill_min, void* node_table, int d_split);

// load the kernels of assembly.cu now (see dense_preload in dense.h)
cudaError_t assembly_preload();

// scratch for the table of node_const() of the first bisection levels; launch_assembly rebuilds
// it at the start of every assembly (its entries depend on omega and arc_coeff)
size_t assembly_node_table_bytes(int order);
}  // namespace emme

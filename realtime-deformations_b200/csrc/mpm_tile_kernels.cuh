// Block-tile P2G / G2P kernels (placeholder: forwards to the baseline kernels until the tile kernels land).
#pragma once
#include "mpm_kernels.cuh"
namespace mpm {
inline cudaError_t tile_kernels_init() { return cudaSuccess; }
template <int MODE>
cudaError_t launch_p2g_tile(Planes P, const int* sorted_ids, const int* blk_start, const int* blk_count, const int* pblock_list,
                            DevCounters* dc, float4* grid, GridDims gd, SimConst sc, float dt, int num_sms, int n_bound, cudaStream_t st) {
    (void)blk_start; (void)blk_count; (void)pblock_list; (void)num_sms;
    k_p2g_atomic<MODE><<<(n_bound + 127) / 128, 128, 0, st>>>(P, sorted_ids, dc, grid, gd, sc, dt);
    return cudaGetLastError();
}
template <int FLAGS>
cudaError_t launch_g2p_tile(Planes C, Planes N, const int* sorted_ids, const int* blk_start, const int* blk_count, const int* pblock_list,
                            DevCounters* dc, const float4* grid, GridDims gd, SimConst sc, float dt, int num_sms, int n_bound, cudaStream_t st) {
    k_g2p_direct<FLAGS><<<(n_bound + 127) / 128, 128, 0, st>>>(C, N, sorted_ids, dc, grid, gd, sc, dt);
    return cudaGetLastError();
}
}  // namespace mpm

// kernels.hpp — host-callable launchers of kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "plan.hpp"

namespace imm3 {

size_t blocks_kernel_smem_bytes(int npfor, int max_block_rows);
cudaError_t blocks_kernel_occupancy(size_t dyn_smem, int* blocks_per_sm);
cudaError_t filter_kernel_occupancy(size_t dyn_smem, int* blocks_per_sm);
cudaError_t emit_kernel_occupancy(bool general, int* blocks_per_sm);
cudaError_t launch_filter(const ScanPlan& plan, uint32_t* bitmap, uint32_t* span_cnt, uint32_t* tile_cnt, unsigned long long* tile_off,
                          ScanCtrl* ctrl, int grid, size_t dyn_smem, cudaStream_t stream);
cudaError_t launch_emit(const ScanPlan& plan, const uint32_t* bitmap, const uint32_t* span_cnt, const unsigned long long* tile_off,
                        int spans_per_tile, long long nspans, int grid, int dense_off, const ScanCtrl* ctrl, bool general, bool pdl,
                        cudaStream_t stream);
size_t emit_stream_smem_bytes(int stage_bytes, int ring);
int emit_stream_header_bytes();
cudaError_t emit_stream_occupancy(size_t dyn_smem, int* blocks_per_sm);
cudaError_t launch_emit_stream(const ScanPlan& plan, const uint32_t* bitmap, const uint32_t* span_cnt, const uint32_t* tile_cnt,
                               const unsigned long long* tile_off, long long nsub, int ring, int stage_bytes, int dense_mode, int grid,
                               size_t dyn_smem, ScanCtrl* ctrl, bool pdl, cudaStream_t stream);
cudaError_t launch_agg_init(AggEntry* table, uint32_t slots, const AggPlan& a, unsigned int* counters, cudaStream_t stream);
cudaError_t launch_agg(const AggPlan& a, const uint32_t* bitmap, const uint32_t* span_cnt, AggEntry* table, AggEntry* out, unsigned int* counters,
                       const ScanCtrl* ctrl, int num_sms, cudaStream_t stream);
long long scan_inline_max_tiles();
cudaError_t launch_offset_scan(const uint32_t* tile_cnt, unsigned long long* tile_off, long long ntiles, long long limit, uint32_t epoch,
                               unsigned long long* partials, ScanCtrl* ctrl, unsigned int* tile_list, cudaStream_t stream);
cudaError_t blocks_scan_emit_grid(int num_sms, long long ntiles8, int* grid);
cudaError_t launch_blocks_scan_emit(const ScanPlan& plan, const uint32_t* bitmapB, const uint32_t* blk_cnt, const uint32_t* tile_cnt,
                                    unsigned long long* tile_off, long long nblocks, uint32_t epoch, unsigned long long* partials, ScanCtrl* ctrl,
                                    unsigned int* tile_list, bool pdl, int grid, cudaStream_t stream, CtrlBlock* pub = nullptr,
                                    unsigned long long pub_seq = 0);
int blocks_group_emit_max_groups();
size_t blocks_group_sum_bytes();
cudaError_t blocks_group_emit_grid(int num_sms, long long nblocks, int* grid, int* ngroups);
cudaError_t launch_blocks_group_emit(const ScanPlan& plan, const uint32_t* bitmapB, const uint32_t* blk_cnt, uint32_t* grp_sum, long long nblocks,
                                     int ngroups, ScanCtrl* ctrl, bool pdl, int grid, cudaStream_t stream, CtrlBlock* pub = nullptr,
                                     unsigned long long pub_seq = 0);
size_t blocks_filter_smem_bytes(int nstaged, int tile_cap_bytes, int ring);
int blocks_filter_slot_bytes(int nstaged, int tile_cap_bytes);
size_t blocks_emit_smem_bytes(int npfor, int words_cap);
int blocks_filter_quad_slot_bytes(int tile_cap_bytes);
cudaError_t blocks_multi_occupancy(size_t filter_smem, size_t emit_smem, int* filter_blocks_per_sm, int* emit_blocks_per_sm, int mode, bool rowspace);  // mode 0: lane = mini-block, 1: quad, 2 | warps << 8: lane = block with that many warps per CTA
cudaError_t launch_blocks_filter(const ScanPlan& plan, uint32_t* bitmapB, uint32_t* blk_cnt, uint32_t* tile_cnt, unsigned long long* tile_off,
                                 ScanCtrl* ctrl, long long nblocks, int grid, size_t dyn_smem, int mode, const unsigned int* work,
                                 cudaStream_t stream, uint32_t* grp_sum = nullptr);  // grp_sum: blocks_group_emit_kernel follows (lane mode only)
cudaError_t launch_block_stats(const PforCol& pc, const uint64_t* row_start, long long nblocks, int words_cap, int num_sms, BlockStat* stats,
                               cudaStream_t stream);
cudaError_t launch_blocks_prune(const PrunePlan& q, const uint64_t* row_start, long long nblocks, long long ntiles8, uint32_t* blk_cnt,
                                uint32_t* tile_cnt, unsigned int* work, int num_sms, cudaStream_t stream, uint32_t* grp_sum = nullptr);
// rowspace: the bitmap / counts come from the dense filter kernel (row space), not from blocks_filter_kernel (block-local)
cudaError_t launch_blocks_emit(const ScanPlan& plan, const uint32_t* bitmap, const uint32_t* cnts, const unsigned int* tile_list, const unsigned long long* tile_off,
                               long long nblocks, const ScanCtrl* ctrl, bool rowspace, bool pdl, int grid, size_t dyn_smem,
                               cudaStream_t stream);
cudaError_t launch_count_exchange(const CommPlan& plan, const ScanCtrl* ctrl, CommOut* out, cudaStream_t stream);
cudaError_t launch_scan_blocks(const ScanPlan& plan, ScanCtrl* ctrl, unsigned long long* status, int grid, size_t dyn_smem,
                               cudaStream_t stream);

}  // namespace imm3

// k_blocks_prune.cuh - per-block min/max of the sorted-integer codec's columns, and the filter front end that uses them
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// Block pruning (the reference declares the hook - SegmentStats / check(), Segment.scala:18-30 - and leaves
// it a stub that always answers true).  block_stats_kernel computes the exact signed min / max of every block of an
// encoded INT column once, when the table is opened (one warp per block, full decode: exact for unsorted data too).
// blocks_prune_kernel then decides most blocks of a range query from 8 bytes instead of ~290: a block whose [min, max]
// misses the window of ANY predicate has no surviving row, one that lies inside EVERY window keeps all of its rows; what is
// left - the blocks a window edge cuts through, two per window on a sorted column - goes on a work list of tiles for the
// regular filter kernel, which then stages and decides only those tiles.  Results are identical with and without pruning
// (IMM3_NO_PRUNE=1 switches it off per query; the benchmark reports both).
// =============================================================================================
__global__ void __launch_bounds__(kComputeThreads) block_stats_kernel(PforCol pc, const uint64_t* __restrict__ row_start, long long nblocks,
                                                                     int words_cap, BlockStat* __restrict__ stats) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t* const Wb = reinterpret_cast<uint32_t*>(dyn_smem) + warp * (words_cap + kBlkVals);
    uint32_t* const vals = Wb + words_cap;
    for (long long blk = (long long)blockIdx.x * kComputeWarps + warp; blk < nblocks; blk += (long long)gridDim.x * kComputeWarps) {
        const long long R0 = (long long)row_start[blk];
        const int n = (int)((long long)row_start[blk + 1] - R0);
        const uint32_t w0 = __ldg(pc.word_off + blk), w1 = __ldg(pc.word_off + blk + 1);
        __syncwarp();
        const uint32_t base = pfor_decode_warp(pc.words, w0, w1, n, Wb, words_cap, vals, lane);
        const int nmini = n >> 5;
        int mn = INT_MAX, mx = INT_MIN;
        if (lane < nmini) {
#pragma unroll 8
            for (int j = 0; j < 32; j++) {
                const int v = (int)(vals[lane * kBlkLane + j] + base);
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
            }
        }
        const int tail = n - (nmini << 5);  // var-byte remainder: absolute values behind the last mini-block's row
        if (lane < tail) {
            const int v = (int)vals[nmini * kBlkLane + lane];
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
        }
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if (lane == 0) *reinterpret_cast<int2*>(stats + blk) = make_int2(mn, mx);
    }
}

// One lane per block, one warp per 32 consecutive blocks.  work[0] = number of listed tiles, work[1 ..] = their indices.
__global__ void __launch_bounds__(kComputeThreads) blocks_prune_kernel(const __grid_constant__ PrunePlan Q, const uint64_t* __restrict__ row_start,
                                                                      long long nblocks, long long ntiles8, uint32_t* __restrict__ blk_cnt,
                                                                      uint32_t* __restrict__ tile_cnt, unsigned int* __restrict__ work, uint32_t* __restrict__ grp_sum) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long ngroups32 = (nblocks + 31) >> 5;
    for (long long T = warp0; T < ngroups32; T += nwarps) {
        const long long b = T * 32 + lane;
        const bool valid = b < nblocks;
        unsigned n = 0;
        bool none = !valid, all = valid;
        if (valid) {
#pragma unroll
            for (int f = 0; f < kMaxFilterCols; f++) {
                if (f < Q.nfilter) {
                    const int2 st = __ldg(reinterpret_cast<const int2*>(Q.stats[f] + b));
                    none = none || st.y < Q.lo[f] || st.x > Q.hi[f];
                    all = all && st.x >= Q.lo[f] && st.y <= Q.hi[f];
                }
            }
        }
        const bool partial = valid && !none && !all;
        if (valid && !none && all) n = (unsigned)(row_start[b + 1] - row_start[b]);  // (row counts matter only for fully selected blocks)
        const unsigned cnt = n;
        const unsigned pmask = __ballot_sync(0xFFFFFFFFu, partial);
        bool listed;  // this lane's tile goes to the filter kernel, which decides (and writes) all of its blocks
        if (Q.group_shift == 5) {
            listed = pmask != 0u;
            if (listed && lane == 0) work[1 + atomicAdd(work, 1u)] = (unsigned)T;
        } else {
            // tiles of 8 blocks: lanes 0, 8, 16, 24 list their tile if it holds a partial block
            listed = ((pmask >> (lane & 24)) & 0xFFu) != 0u;
            if (listed && (lane & 7) == 0 && T * 4 + (lane >> 3) < ntiles8) work[1 + atomicAdd(work, 1u)] = (unsigned)(T * 4 + (lane >> 3));
        }
        if (valid && !listed) blk_cnt[b] = cnt;
        unsigned c8 = cnt;
        c8 += __shfl_xor_sync(0xFFFFFFFFu, c8, 1);
        c8 += __shfl_xor_sync(0xFFFFFFFFu, c8, 2);
        c8 += __shfl_xor_sync(0xFFFFFFFFu, c8, 4);
        if (!listed && (lane & 7) == 0 && T * 4 + (lane >> 3) < ntiles8) tile_cnt[T * 4 + (lane >> 3)] = c8;
        if (grp_sum && Q.group_shift == 5) {  // (blocks_group_emit_kernel follows; a listed tile is counted by the filter kernel)
            unsigned c32 = c8 + __shfl_xor_sync(0xFFFFFFFFu, c8, 8);
            c32 += __shfl_xor_sync(0xFFFFFFFFu, c32, 16);
            if (lane == 0 && !listed && c32 != 0u) atomicAdd(grp_sum + (T >> 5), c32);
        }
    }
}

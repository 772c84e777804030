// k_agg.cuh - count / min / max ... group by over the selection bitmap (ProjectAggOp, ProjectAggregate.scala:115-226)
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// The reference aggregates per segment worker into a LinkedHashMap keyed by the group values (ProjectAggregate.scala:
// 160-166) and merges the per-segment maps in ProjectAggregateQueueOp (ProjectAggregateQueue.scala:21-45).  Here:
//   filter_kernel            the same K1 as every query: selection bitmap + span / tile counts (nothing is projected)
//   agg_kernel               one warp per 1024-row span with selected rows: the span's selection vector (append_selection),
//                            then 32 selected rows at a time - a lane gathers its row's group cells (packed into a 64-bit
//                            key) and aggregate inputs, lanes with equal keys combine through MATCH.ANY + REDUX, one lane
//                            per distinct key updates the CTA's shared-memory hash table; when the CTA has no tile left,
//                            its groups are merged into the global table with atomics
//   agg_compact_kernel       occupied slots of the global table -> a dense array (any order; the host sorts the groups by
//                            the canonical ordinal of their first row, the order the reference reports them in with one
//                            worker) and clears the table for the next query
// Exact integer arithmetic: counts are 64-bit, min / max are taken on the raw int32 / int8 values (the reference widens to
// Double, which is exact for both types; the host formats them as Doubles).
// =============================================================================================
constexpr unsigned long long kAggEmpty = ~0ull;

__device__ __forceinline__ uint32_t agg_hash(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xFF51AFD7ED558CCDull;
    k ^= k >> 33;
    return (uint32_t)k;
}

struct AggShared {
    unsigned long long key[kAggSmemSlots];
    unsigned long long first_row[kAggSmemSlots];
    long long val[kMaxAggs][kAggSmemSlots];
};

// Find or claim the slot of `key` in a table of `slots` entries (power of two).  Returns the slot, or -1 if the probe
// sequence is exhausted (shared table full: the caller goes to the global table).
template <typename KeyPtr>
__device__ __forceinline__ int agg_find_slot(KeyPtr keys, uint32_t slots, unsigned long long key, int max_probe) {
    uint32_t h = agg_hash(key) & (slots - 1);
    for (int p = 0; p < max_probe; p++, h = (h + 1) & (slots - 1)) {
        unsigned long long cur = keys[h];
        if (cur == key) return (int)h;
        if (cur == kAggEmpty) {
            cur = atomicCAS(&keys[h], kAggEmpty, key);
            if (cur == kAggEmpty || cur == key) return (int)h;
        }
    }
    return -1;
}

// Slot of `key` in the global table (claimed if new), -1 = table full.
__device__ __forceinline__ int agg_global_slot(AggEntry* table, uint32_t slots, unsigned long long key) {
    uint32_t h = agg_hash(key) & (slots - 1);
    for (uint32_t p = 0; p < slots; p++, h = (h + 1) & (slots - 1)) {
        unsigned long long* kp = &table[h].key;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
        if (cur == kAggEmpty) {
            const unsigned long long old = atomicCAS(kp, kAggEmpty, key);
            cur = old == kAggEmpty ? key : old;
        }
        if (cur == key) return (int)h;
    }
    return -1;
}

__device__ __forceinline__ void agg_update(long long* v, int op, long long x) {
    if (op == kAggCount) atomicAdd(reinterpret_cast<unsigned long long*>(v), (unsigned long long)x);
    else if (op == kAggMin) atomicMin(v, x);
    else atomicMax(v, x);
}
__device__ __forceinline__ long long agg_identity(int op) { return op == kAggCount ? 0ll : (op == kAggMin ? LLONG_MAX : LLONG_MIN); }

__global__ void __launch_bounds__(kComputeThreads, 2) agg_kernel(const __grid_constant__ AggPlan A, const uint32_t* __restrict__ bitmap,
                                                                  const uint32_t* __restrict__ span_cnt, AggEntry* __restrict__ table,
                                                                  unsigned int* __restrict__ overflow, const ScanCtrl* ctrl) {
    extern __shared__ __align__(128) uint8_t agg_smem[];
    AggShared& S = *reinterpret_cast<AggShared*>(agg_smem);
    unsigned short* const sel_all = reinterpret_cast<unsigned short*>(agg_smem + sizeof(AggShared));
    __shared__ AggCol s_agg[kMaxAggs];
    __shared__ GroupCol s_group[kMaxGroupCols];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int i = 0; i < kMaxAggs; i++)
        if (tid == i) s_agg[i] = A.agg[i];
#pragma unroll
    for (int i = 0; i < kMaxGroupCols; i++)
        if (tid == 32 + i) s_group[i] = A.group[i];
    __syncthreads();
    for (int i = tid; i < kAggSmemSlots; i += kComputeThreads) {
        S.key[i] = kAggEmpty;
        S.first_row[i] = ~0ull;
        for (int a = 0; a < A.naggs; a++) S.val[a][i] = agg_identity(s_agg[a].op);
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the filter kernel's bitmap and counts are final
    if (__ldcg(&ctrl->total) != 0ull) {
        unsigned short* const sel_w = sel_all + warp * 1024;
        const long long nspans = A.ntiles * 8;
        for (long long span = (long long)blockIdx.x * kComputeWarps + warp; span < nspans; span += (long long)gridDim.x * kComputeWarps) {
            const unsigned n = __ldg(span_cnt + span);
            if (n == 0) continue;
            const uint32_t m = __ldg(bitmap + span * 32 + lane);
            __syncwarp();
            append_selection(m, lane, sel_w, 0u);
            __syncwarp();
            const long long row0 = span * 1024;
            for (unsigned i0 = 0; i0 < n; i0 += 32) {
                const bool live = i0 + lane < n;
                const long long row = row0 + (live ? sel_w[i0 + lane] : 0);
                unsigned long long key = 0;
                for (int gcol = 0; gcol < A.ngroup; gcol++) {
                    const GroupCol gc = s_group[gcol];
                    unsigned long long cell = 0;
                    const uint8_t* src = gc.base + row * gc.width;
                    for (int b = 0; b < gc.width; b++) cell |= (unsigned long long)__ldg(src + b) << (8 * b);
                    key |= cell << gc.key_shift;
                }
                long long x[kMaxAggs];
#pragma unroll
                for (int a = 0; a < kMaxAggs; a++) {
                    x[a] = 0;
                    if (a < A.naggs) {
                        const AggCol ac = s_agg[a];
                        if (ac.op == kAggCount) x[a] = 1;
                        else x[a] = ac.width == 4 ? (long long)(int)__ldg(reinterpret_cast<const uint32_t*>(ac.base) + row) : (long long)(signed char)__ldg(ac.base + row);
                    }
                }
                // lanes with the same key combine first: one table update per distinct key of the 32 rows
                const unsigned active = __ballot_sync(0xFFFFFFFFu, live);
                if (live) {
                    const unsigned peers = __match_any_sync(active, key);
                    const int leader = __ffs((int)peers) - 1;
                    unsigned long long fr = (unsigned long long)row;
                    // (rows of a selection vector ascend with the lane: the leader - lowest lane - holds the smallest row)
#pragma unroll
                    for (int a = 0; a < kMaxAggs; a++) {
                        if (a < A.naggs) {
                            const int op = s_agg[a].op;
                            if (op == kAggCount) x[a] = (long long)__popc(peers);
                            else {
                                // 64-bit values: reduce the (int32-range) payload as 32-bit
                                const int v32 = (int)x[a];
                                x[a] = op == kAggMin ? (long long)__reduce_min_sync(peers, v32) : (long long)__reduce_max_sync(peers, v32);
                            }
                        }
                    }
                    if (lane == leader) {
                        int slot = agg_find_slot(S.key, kAggSmemSlots, key, 64);
                        if (slot >= 0) {
                            atomicMin(&S.first_row[slot], fr);
                            for (int a = 0; a < A.naggs; a++) agg_update(&S.val[a][slot], s_agg[a].op, x[a]);
                        } else {
                            // the CTA's table is full: straight to the global table
                            const int gs = agg_global_slot(table, A.table_slots, key);
                            if (gs < 0) atomicExch(overflow, 1u);
                            else {
                                atomicMin(&table[gs].first_row, fr);
                                for (int a = 0; a < A.naggs; a++) agg_update(&table[gs].val[a], s_agg[a].op, x[a]);
                            }
                        }
                    }
                }
            }
        }
    }
    // merge the CTA's groups into the global table
    __syncthreads();
    for (int i = tid; i < kAggSmemSlots; i += kComputeThreads) {
        const unsigned long long key = S.key[i];
        if (key == kAggEmpty) continue;
        const int gs = agg_global_slot(table, A.table_slots, key);
        if (gs < 0) { atomicExch(overflow, 1u); continue; }
        atomicMin(&table[gs].first_row, S.first_row[i]);
        for (int a = 0; a < A.naggs; a++) agg_update(&table[gs].val[a], s_agg[a].op, S.val[a][i]);
    }
}

// Occupied slots -> out[0 .. *nout), in any order (the host sorts the groups by first_row).
__global__ void __launch_bounds__(kComputeThreads) agg_compact_kernel(const AggEntry* __restrict__ table, uint32_t slots, AggEntry* __restrict__ out,
                                                                     unsigned int* __restrict__ nout) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += gridDim.x * blockDim.x) {
        const AggEntry e = table[i];
        if (e.key != kAggEmpty) out[atomicAdd(nout, 1u)] = e;
    }
}

// An empty table for this query's aggregates (identities: 0 for COUNT, +inf / -inf for MIN / MAX).
struct AggOps {
    int32_t op[kMaxAggs];
    int32_t naggs;
};
__global__ void __launch_bounds__(kComputeThreads) agg_init_kernel(AggEntry* __restrict__ table, uint32_t slots, AggOps ops, unsigned int* __restrict__ counters) {
    if (blockIdx.x == 0 && threadIdx.x < 2) counters[threadIdx.x] = 0;  // [0] = groups compacted, [1] = overflow flag
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += gridDim.x * blockDim.x) {
        AggEntry e;
        e.key = kAggEmpty;
        e.first_row = ~0ull;
#pragma unroll
        for (int a = 0; a < kMaxAggs; a++) e.val[a] = a < ops.naggs ? agg_identity(ops.op[a]) : 0;
        table[i] = e;
    }
}

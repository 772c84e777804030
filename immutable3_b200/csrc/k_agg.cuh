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

// Per-WARP hash tables in shared memory: a group is updated by one lane per 32-row step (the leader of the lanes that hold
// its key), so inside a warp's own table the values need no atomics at all - only claiming a fresh slot does (two new keys
// may want the same slot in one step).  With one table per CTA, a low-cardinality group-by (51 states) serialised every
// warp of the SM on the same few shared-memory words: 5.2 ms for the reference's example on 100 M rows.
constexpr int kAggWarpSlots = 128;
struct AggWarpTable {
    unsigned long long key[kAggWarpSlots];
    unsigned long long first_row[kAggWarpSlots];
    int val[kMaxAggs][kAggWarpSlots];  // 32-bit: native shared-memory atomics (a warp counts far fewer than 2^32 rows; min / max inputs are int32 / int8)
};

// Slot of `key` in a shared-memory table (claimed if new), -1 if `max_probe` slots in a row are taken by other keys.
__device__ __forceinline__ int agg_smem_slot(unsigned long long* keys, unsigned long long key, int max_probe) {
    uint32_t h = agg_hash(key) & (kAggWarpSlots - 1);
    for (int p = 0; p < max_probe; p++, h = (h + 1) & (kAggWarpSlots - 1)) {
        unsigned long long cur = keys[h];
        if (cur == kAggEmpty) {
            const unsigned long long old = atomicCAS(&keys[h], kAggEmpty, key);
            cur = old == kAggEmpty ? key : old;
        }
        if (cur == key) return (int)h;
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

__device__ __forceinline__ void agg_to_global(AggEntry* table, uint32_t slots, unsigned int* overflow, unsigned long long key, unsigned long long fr,
                                              const long long* x, const AggCol* aggs, int naggs) {
    const int gs = agg_global_slot(table, slots, key);
    if (gs < 0) {
        atomicExch(overflow, 1u);
        return;
    }
    atomicMin(&table[gs].first_row, fr);
    for (int a = 0; a < naggs; a++) agg_update(&table[gs].val[a], aggs[a].op, x[a]);
}

__global__ void __launch_bounds__(kComputeThreads, 2) agg_kernel(const __grid_constant__ AggPlan A, const uint32_t* __restrict__ bitmap,
                                                                  const uint32_t* __restrict__ span_cnt, AggEntry* __restrict__ table,
                                                                  unsigned int* __restrict__ overflow, const ScanCtrl* ctrl) {
    extern __shared__ __align__(128) uint8_t agg_smem[];
    AggWarpTable* const tables = reinterpret_cast<AggWarpTable*>(agg_smem);
    unsigned short* const sel_all = reinterpret_cast<unsigned short*>(agg_smem + kComputeWarps * sizeof(AggWarpTable));
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
    AggWarpTable& T = tables[warp];
    for (int i = lane; i < kAggWarpSlots; i += 32) {
        T.key[i] = kAggEmpty;
        T.first_row[i] = ~0ull;
        for (int a = 0; a < A.naggs; a++) T.val[a][i] = s_agg[a].op == kAggCount ? 0 : (s_agg[a].op == kAggMin ? INT_MAX : INT_MIN);
    }
    __syncwarp();
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
                    if (gc.width == 2) cell = __ldg(reinterpret_cast<const unsigned short*>(src));
                    else if (gc.width == 4) cell = __ldg(reinterpret_cast<const uint32_t*>(src));
                    else
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
                // every lane updates its group in the warp's table with native 32-bit shared-memory atomics.  (Combining the lanes
                // of a key first - MATCH.ANY + REDUX per distinct key - was 10x slower: with ~20 distinct keys per 32 rows the
                // partial-mask reductions run one after the other.)
                if (live) {
                    const int slot = agg_smem_slot(T.key, key, 16);
                    if (slot >= 0) {
                        // rows reach a warp in ascending order: only the step that creates the group can lower first_row
                        if ((unsigned long long)row < *reinterpret_cast<volatile unsigned long long*>(&T.first_row[slot])) atomicMin(&T.first_row[slot], (unsigned long long)row);
#pragma unroll
                        for (int a = 0; a < kMaxAggs; a++) {
                            if (a < A.naggs) {
                                const int op = s_agg[a].op;
                                if (op == kAggCount) atomicAdd(reinterpret_cast<unsigned int*>(&T.val[a][slot]), 1u);
                                else if (op == kAggMin) atomicMin(&T.val[a][slot], (int)x[a]);
                                else atomicMax(&T.val[a][slot], (int)x[a]);
                            }
                        }
                    } else {
                        agg_to_global(table, A.table_slots, overflow, key, (unsigned long long)row, x, s_agg, A.naggs);  // the warp's table is full
                    }
                }
            }
        }
    }
    // the CTA's warps fold their tables into warp 0's (shared-memory atomics now: several writers), which goes to the global table
    __syncthreads();
    AggWarpTable& T0 = tables[0];
    if (warp > 0) {
        for (int i = lane; i < kAggWarpSlots; i += 32) {
            const unsigned long long key = T.key[i];
            if (key == kAggEmpty) continue;
            long long x[kMaxAggs];
            for (int a = 0; a < A.naggs; a++) x[a] = s_agg[a].op == kAggCount ? (long long)(unsigned int)T.val[a][i] : (long long)T.val[a][i];
            const int slot = agg_smem_slot(T0.key, key, kAggWarpSlots);
            if (slot >= 0) {
                atomicMin(&T0.first_row[slot], T.first_row[i]);
                for (int a = 0; a < A.naggs; a++) {
                    const int op = s_agg[a].op;
                    if (op == kAggCount) atomicAdd(reinterpret_cast<unsigned int*>(&T0.val[a][slot]), (unsigned int)T.val[a][i]);  // (a CTA counts < 2^32 rows)
                    else if (op == kAggMin) atomicMin(&T0.val[a][slot], T.val[a][i]);
                    else atomicMax(&T0.val[a][slot], T.val[a][i]);
                }
            } else {
                agg_to_global(table, A.table_slots, overflow, key, T.first_row[i], x, s_agg, A.naggs);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < kAggWarpSlots; i += kComputeThreads) {
        const unsigned long long key = T0.key[i];
        if (key == kAggEmpty) continue;
        long long x[kMaxAggs];
        for (int a = 0; a < A.naggs; a++) x[a] = s_agg[a].op == kAggCount ? (long long)(unsigned int)T0.val[a][i] : (long long)T0.val[a][i];
        agg_to_global(table, A.table_slots, overflow, key, T0.first_row[i], x, s_agg, A.naggs);
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

// k_agg.cuh - count / min / max ... group by over the selection bitmap (ProjectAggOp, ProjectAggregate.scala:115-226)
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// The reference aggregates per segment worker into a LinkedHashMap keyed by the group values (ProjectAggregate.scala:
// 160-166) and merges the per-segment maps in ProjectAggregateQueueOp (ProjectAggregateQueue.scala:21-45).  Here:
//   filter_kernel            the same K1 as every query: selection bitmap + span / tile counts (nothing is projected)
//   agg_kernel               one warp per 1024-row span with selected rows: the span's selection vector (append_selection),
//                            then 64 selected rows per iteration (two steps of 32, the gathers of both issued before the
//                            first update) - a lane gathers its row's group cells (packed into a 64-bit key) and aggregate
//                            inputs and updates its group in the WARP's own shared-memory hash table: the slot is found
//                            with one load when the group exists, MIN / MAX read the slot first and only issue an atomic
//                            when the row improves it (after the first rows of a group almost never), COUNT is one native
//                            32-bit shared-memory atomic; warps fold into one table per CTA, CTAs into the global table
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

// Per-WARP hash tables in shared memory.  With one table per CTA, a low-cardinality group-by (51 states) serialised every
// warp of the SM on the same few shared-memory words: 5.2 ms for the reference's example on 100 M rows.
// MIN and MAX share one update: a MIN column is kept as the MAX of the complemented inputs (~x reverses the order of int32
// without overflow), so both are "read the slot, atomicMax only if the row beats it"; the value is complemented back when
// the table leaves shared memory.
constexpr int kAggWarpSlots = 128;
constexpr int kAggWarpSlotBits = 7;
template <int NA>
struct AggWarpTableT {
    unsigned long long key[kAggWarpSlots];
    unsigned long long first_row[kAggWarpSlots];
    int val[NA][kAggWarpSlots];  // 32-bit: native shared-memory atomics (a CTA counts far fewer than 2^32 rows; min / max inputs are int32 / int8)
};
// Home slot in a shared-memory table: two 32-bit multiplies, the top bits of their XOR.
__device__ __forceinline__ uint32_t agg_home_slot(unsigned long long k) {
    return (((uint32_t)k * 0x9E3779B1u) ^ ((uint32_t)(k >> 32) * 0x85EBCA6Bu)) >> (32 - kAggWarpSlotBits);
}
// Slot of `key` in a shared-memory table (linear probing from its home slot; bit 8 of the result = the slot was claimed by
// this call), -1 if `max_probe` slots in a row are taken by other keys.
__device__ __forceinline__ int agg_smem_slot(unsigned long long* keys, unsigned long long key, int max_probe) {
    uint32_t h = agg_home_slot(key);
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&keys[h]);
    if (cur == key) return (int)h;  // (nearly every row after a group's first)
    for (int p = 0; p < max_probe; p++) {
        if (cur == kAggEmpty) {
            const unsigned long long old = atomicCAS(&keys[h], kAggEmpty, key);
            if (old == kAggEmpty) return (int)h | 0x100;
            cur = old;
        }
        if (cur == key) return (int)h;
        h = (h + 1) & (kAggWarpSlots - 1);
        cur = *reinterpret_cast<volatile unsigned long long*>(&keys[h]);
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

// One group's partial (x[a]: a count, or an extreme as int64) into the global table.
__device__ __noinline__ void agg_to_global(AggEntry* table, uint32_t slots, unsigned int* overflow, unsigned long long key, unsigned long long fr,
                                           const long long* x, const AggCol* aggs, int naggs) {
    const int gs = agg_global_slot(table, slots, key);
    if (gs < 0) {
        atomicExch(overflow, 1u);
        return;
    }
    atomicMin(&table[gs].first_row, fr);
    for (int a = 0; a < naggs; a++) agg_update(&table[gs].val[a], aggs[a].op, x[a]);
}
// A shared-memory value (COUNT: the count; MIN: max of ~x; MAX: max of x) as the int64 the global table holds.
__device__ __forceinline__ long long agg_smem_value(int op, int v) { return op == kAggCount ? (long long)(unsigned int)v : (long long)(op == kAggMin ? ~v : v); }

// NG = group-by columns (0, 1, 2 exact; 4 = A.ngroup of them, up to 4), NA = aggregates (1 .. 4 exact; 8 = A.naggs of them, up to 8):
// the loops over the plan are unrolled with compile-time indices into the kernel's parameter block, so the per-column
// descriptors are constant-bank operands - nothing about the query is decoded per row and nothing of it lives in registers.
template <int NG, int NA>
__global__ void __launch_bounds__(kComputeThreads, 3) agg_kernel(const __grid_constant__ AggPlan A, const uint32_t* __restrict__ bitmap,
                                                                  const uint32_t* __restrict__ span_cnt, AggEntry* __restrict__ table,
                                                                  unsigned int* __restrict__ overflow, const ScanCtrl* ctrl) {
    using WarpTable = AggWarpTableT<NA>;
    extern __shared__ __align__(128) uint8_t agg_smem[];
    WarpTable* const tables = reinterpret_cast<WarpTable*>(agg_smem);
    unsigned short* const sel_all = reinterpret_cast<unsigned short*>(agg_smem + kComputeWarps * sizeof(WarpTable));
    __shared__ AggCol s_agg[kMaxAggs];  // (for the cold paths only: a row or a table entry that goes to the global table)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int i = 0; i < kMaxAggs; i++)
        if (tid == i) s_agg[i] = A.agg[i];
    const int naggs = NA == 8 ? A.naggs : NA, ngroup = NG == 4 ? A.ngroup : NG;
    WarpTable& T = tables[warp];
    for (int i = lane; i < kAggWarpSlots; i += 32) {
        T.key[i] = kAggEmpty;
        T.first_row[i] = ~0ull;
#pragma unroll
        for (int a = 0; a < NA; a++) T.val[a][i] = A.agg[a].op == kAggCount ? 0 : INT_MIN;
    }
    __syncthreads();

    // the group key and the aggregate inputs of one row (MIN inputs complemented)
    auto gather = [&](long long row, unsigned long long& key, int (&x)[NA]) {
        key = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            if (NG == 4 && g >= ngroup) break;
            const int w = A.group[g].width;
            const uint8_t* src = A.group[g].base + row * w;
            unsigned long long cell = 0;
            if (w == 2) cell = __ldg(reinterpret_cast<const unsigned short*>(src));
            else if (w == 1) cell = __ldg(src);
            else if (w == 4) cell = __ldg(reinterpret_cast<const uint32_t*>(src));
            else
                for (int b = 0; b < w; b++) cell |= (unsigned long long)__ldg(src + b) << (8 * b);
            key |= cell << A.group[g].key_shift;
        }
#pragma unroll
        for (int a = 0; a < NA; a++) {
            x[a] = 0;
            if (NA == 8 && a >= naggs) break;
            const int op = A.agg[a].op;
            if (op != kAggCount) {
                const int v = A.agg[a].width == 4 ? (int)__ldg(reinterpret_cast<const uint32_t*>(A.agg[a].base) + row) : (int)(signed char)__ldg(A.agg[a].base + row);
                x[a] = op == kAggMin ? ~v : v;
            }
        }
    };
    // one row into its group of the warp's table (all 32 lanes call this; `live` lanes carry a row)
    auto apply = [&](bool live, long long row, unsigned long long key, const int (&x)[NA]) {
        int slot = -1;
        if (live) slot = agg_smem_slot(T.key, key, 16);
        // rows reach a warp in ascending order: a group's first row is in the step that creates its slot (maybe on another
        // lane than the one that claimed it) - only such a step touches first_row
        if (__any_sync(0xFFFFFFFFu, slot >= 0x100)) {
            if (slot >= 0) {
                slot &= 0xFF;
                atomicMin(&T.first_row[slot], (unsigned long long)row);
            }
        }
        if (!live) return;
        if (slot >= 0) {
#pragma unroll
            for (int a = 0; a < NA; a++) {
                if (NA == 8 && a >= naggs) break;
                int* const v = &T.val[a][slot];
                if (A.agg[a].op == kAggCount) atomicAdd(reinterpret_cast<unsigned int*>(v), 1u);
                else if (x[a] > *reinterpret_cast<volatile int*>(v)) atomicMax(v, x[a]);  // (a stale read is only ever too low: at worst one atomic too many)
            }
        } else {  // the warp's table is full
            long long xl[kMaxAggs];
#pragma unroll
            for (int a = 0; a < NA; a++) xl[a] = A.agg[a].op == kAggCount ? 1ll : (long long)(A.agg[a].op == kAggMin ? ~x[a] : x[a]);
            agg_to_global(table, A.table_slots, overflow, key, (unsigned long long)row, xl, s_agg, naggs);
        }
    };

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
            for (unsigned i0 = 0; i0 < n; i0 += 64) {  // two steps of 32 rows: the gathers of both are in flight before the first update
                const bool live0 = i0 + lane < n, live1 = i0 + 32 + lane < n;
                const long long r0 = row0 + (live0 ? sel_w[i0 + lane] : 0), r1 = row0 + (live1 ? sel_w[i0 + 32 + lane] : 0);
                unsigned long long k0 = 0, k1 = 0;
                int x0[NA], x1[NA];
                if (live0) gather(r0, k0, x0);
                if (live1) gather(r1, k1, x1);
                apply(live0, r0, k0, x0);
                if (i0 + 32 < n) apply(live1, r1, k1, x1);
            }
        }
    }
    // the CTA's warps fold their tables into warp 0's (several writers now: atomics throughout), which goes to the global table
    __syncthreads();
    WarpTable& T0 = tables[0];
    if (warp > 0) {
        for (int i = lane; i < kAggWarpSlots; i += 32) {
            const unsigned long long key = T.key[i];
            if (key == kAggEmpty) continue;
            int slot = agg_smem_slot(T0.key, key, kAggWarpSlots);
            if (slot >= 0) {
                slot &= 0xFF;
                atomicMin(&T0.first_row[slot], T.first_row[i]);
#pragma unroll
                for (int a = 0; a < NA; a++) {
                    if (NA == 8 && a >= naggs) break;
                    if (A.agg[a].op == kAggCount) atomicAdd(reinterpret_cast<unsigned int*>(&T0.val[a][slot]), (unsigned int)T.val[a][i]);
                    else atomicMax(&T0.val[a][slot], T.val[a][i]);
                }
            } else {
                long long xl[kMaxAggs];
#pragma unroll
                for (int a = 0; a < NA; a++) xl[a] = agg_smem_value(A.agg[a].op, T.val[a][i]);
                agg_to_global(table, A.table_slots, overflow, key, T.first_row[i], xl, s_agg, naggs);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < kAggWarpSlots; i += kComputeThreads) {
        const unsigned long long key = T0.key[i];
        if (key == kAggEmpty) continue;
        long long xl[kMaxAggs];
#pragma unroll
        for (int a = 0; a < NA; a++) xl[a] = agg_smem_value(A.agg[a].op, T0.val[a][i]);
        agg_to_global(table, A.table_slots, overflow, key, T0.first_row[i], xl, s_agg, naggs);
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

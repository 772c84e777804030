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
//                            inputs, finds the group's slot in the CTA's key table through a hash -> slot cache (two loads,
//                            no probing) and updates the WARP's own partial aggregates of that slot: MIN / MAX read the
//                            value first and only issue an atomic when the row improves it (after the first rows of a
//                            group almost never), COUNT is one native 32-bit shared-memory atomic; at the end the CTA
//                            folds its warps' values per slot and adds the groups to the global table
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

// Shared-memory state of a CTA (dynamic shared memory, in this order):
//   dict   2^14 bytes   hash(key) -> slot: a CACHE in front of the key table (any byte value is a valid guess: the key table
//                       confirms it), so a row finds its group with two loads and no probing; written without atomics
//   keys   256 x u64    the CTA's groups: open addressing, a slot is claimed with one CAS and never moves
//   vals   per WARP (NA + 1) x 256 x int32: the warp's own partial aggregates of every slot - a low-cardinality group-by
//                       (51 states) on one table per CTA serialised every warp of the SM on the same few words (5.2 ms for the
//                       reference's example on 100 M rows).  Value NA is the group's first row as the warp saw it.
//   sel    per warp 1024 x u16: the selection vector of the span in hand
// MIN, MAX and the first row share one update - "read the slot, atomicMax only if the row beats it": a MIN column is kept
// as the MAX of the complemented inputs (~x reverses the order of int32 without overflow) and complemented back when the
// table leaves shared memory; after a group's first rows the atomics all but disappear.  COUNT is one native 32-bit
// shared-memory atomic per row (a warp counts far fewer than 2^32 rows).
constexpr int kAggSlots = 256;
constexpr int kAggDictBits = 14;
constexpr size_t agg_smem_bytes(int na) {
    return ((size_t)1 << kAggDictBits) + (size_t)kAggSlots * 8 + (size_t)kComputeWarps * (na + 1) * kAggSlots * 4 + (size_t)kComputeWarps * 1024 * 2;
}
__device__ __forceinline__ uint32_t agg_hash32(unsigned long long k) {
    return ((uint32_t)k * 0x9E3779B1u) ^ ((uint32_t)(k >> 32) * 0x85EBCA6Bu);
}
// Slot of `key` in the CTA's key table (claimed if new), -1 if the table has no room near the key's home slot.
__device__ __forceinline__ int agg_cta_slot(uint8_t* dict, unsigned long long* keys, unsigned long long key) {
    const uint32_t hh = agg_hash32(key);
    volatile uint8_t* const guess = dict + (hh >> (32 - kAggDictBits));
    uint32_t h = *guess;
    if (*reinterpret_cast<volatile unsigned long long*>(&keys[h]) == key) return (int)h;  // (nearly every row after a group's first)
    h = hh >> 24;
    for (int p = 0; p < 16; p++, h = (h + 1) & (kAggSlots - 1)) {  // (16 taken slots in a row: the table is as good as full)
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&keys[h]);
        if (cur == kAggEmpty) {
            const unsigned long long old = atomicCAS(&keys[h], kAggEmpty, key);
            cur = old == kAggEmpty ? key : old;
        }
        if (cur == key) {
            *guess = (uint8_t)h;
            return (int)h;
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
// NG = group-by columns (0, 1, 2 exact; 4 = A.ngroup of them, up to 4), NA = aggregates (1 .. 4 exact; 8 = A.naggs of them, up to 8):
// the loops over the plan are unrolled with compile-time indices into the kernel's parameter block, so the per-column
// descriptors are constant-bank operands - nothing about the query is decoded per row and nothing of it lives in registers.
template <int NG, int NA>
__global__ void __launch_bounds__(kComputeThreads, 3) agg_kernel(const __grid_constant__ AggPlan A, const uint32_t* __restrict__ bitmap,
                                                                  const uint32_t* __restrict__ span_cnt, AggEntry* __restrict__ table,
                                                                  unsigned int* __restrict__ overflow, const ScanCtrl* ctrl) {
    extern __shared__ __align__(128) uint8_t agg_smem[];
    uint8_t* const dict = agg_smem;
    unsigned long long* const keys = reinterpret_cast<unsigned long long*>(agg_smem + (1u << kAggDictBits));
    int* const vals = reinterpret_cast<int*>(agg_smem + (1u << kAggDictBits) + kAggSlots * 8);
    unsigned short* const sel_all = reinterpret_cast<unsigned short*>(vals + kComputeWarps * (NA + 1) * kAggSlots);
    __shared__ AggCol s_agg[kMaxAggs];  // (for the cold paths only: a row or a group that goes to the global table)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int i = 0; i < kMaxAggs; i++)
        if (tid == i) s_agg[i] = A.agg[i];
    const int naggs = NA == 8 ? A.naggs : NA, ngroup = NG == 4 ? A.ngroup : NG;
    for (int i = tid; i < (1 << kAggDictBits) / 4; i += kComputeThreads) reinterpret_cast<uint32_t*>(dict)[i] = 0u;
    for (int i = tid; i < kAggSlots; i += kComputeThreads) keys[i] = kAggEmpty;
    int* const V = vals + warp * (NA + 1) * kAggSlots;  // this warp's values: V[a * kAggSlots + slot]
    for (int i = lane; i < kAggSlots; i += 32) {
#pragma unroll
        for (int a = 0; a < NA; a++) V[a * kAggSlots + i] = A.agg[a].op == kAggCount ? 0 : INT_MIN;
        V[NA * kAggSlots + i] = INT_MIN;
    }
    __syncthreads();

    // The group key and the aggregate inputs of row `r` of the span in hand, in two halves so that the loads of several rows
    // are in flight together: fetch_raw ISSUES the loads (cells of 1, 2 and 4 bytes: the aligned 32-bit word that holds them -
    // the arenas are padded to whole tiles - one load type, no branch on the width, 32-bit offsets) and consumes nothing;
    // decode turns the words into the key and the inputs (MIN inputs complemented).  gp / ap = the columns at the span's
    // first row.  Group cells of other widths are gathered byte-wise in decode.
    const uint8_t* gp[NG > 0 ? NG : 1];
    const uint8_t* ap[NA];
    auto fetch_raw = [&](unsigned r, uint32_t (&gw)[NG > 0 ? NG : 1], uint32_t (&aw)[NA]) {
#pragma unroll
        for (int g = 0; g < NG; g++) {
            gw[g] = 0;
            if (NG == 4 && g >= ngroup) break;
            const unsigned w = (unsigned)A.group[g].width;
            if (w <= 4u && (w & (w - 1u)) == 0u) gw[g] = __ldg(reinterpret_cast<const uint32_t*>(gp[g] + ((r * w) & ~3u)));
        }
#pragma unroll
        for (int a = 0; a < NA; a++) {
            aw[a] = 0;
            if (NA == 8 && a >= naggs) break;
            if (A.agg[a].op != kAggCount) aw[a] = __ldg(reinterpret_cast<const uint32_t*>(ap[a] + ((r * (unsigned)A.agg[a].width) & ~3u)));
        }
    };
    auto decode = [&](unsigned r, const uint32_t (&gw)[NG > 0 ? NG : 1], const uint32_t (&aw)[NA], unsigned long long& key, int (&x)[NA]) {
        key = 0;
#pragma unroll
        for (int g = 0; g < NG; g++) {
            if (NG == 4 && g >= ngroup) break;
            const unsigned w = (unsigned)A.group[g].width;
            const unsigned off = r * w;
            unsigned long long cell = 0;
            if (w <= 4u && (w & (w - 1u)) == 0u) {
                cell = (gw[g] >> ((off & 3u) * 8u)) & (0xFFFFFFFFu >> (32u - 8u * w));
            } else {
                for (unsigned b = 0; b < w; b++) cell |= (unsigned long long)__ldg(gp[g] + off + b) << (8u * b);
            }
            key |= cell << A.group[g].key_shift;
        }
#pragma unroll
        for (int a = 0; a < NA; a++) {
            x[a] = 0;
            if (NA == 8 && a >= naggs) break;
            const int op = A.agg[a].op;
            if (op != kAggCount) {
                const unsigned w = (unsigned)A.agg[a].width;  // 4 (INT) or 1 (TINYINT, sign-extended)
                const int v = (int)(aw[a] << ((4u - w - ((r * w) & 3u)) * 8u)) >> ((4u - w) * 8u);
                x[a] = v ^ (op == kAggMin ? -1 : 0);
            }
        }
    };
    // one row into its group: `ord` = the row's ordinal in this warp's own sequence of rows (ascending: a group's first row
    // is the one that creates its entry in the warp's values)
    auto apply = [&](long long row, unsigned ord, unsigned long long key, const int (&x)[NA]) {
        const int slot = agg_cta_slot(dict, keys, key);
        if (slot >= 0) {
#pragma unroll
            for (int a = 0; a < NA; a++) {
                if (NA == 8 && a >= naggs) break;
                int* const v = V + a * kAggSlots + slot;
                if (A.agg[a].op == kAggCount) atomicAdd(reinterpret_cast<unsigned int*>(v), 1u);
                else if (x[a] > *reinterpret_cast<volatile int*>(v)) atomicMax(v, x[a]);  // (a stale read is only ever too low: at worst one atomic too many)
            }
            int* const f = V + NA * kAggSlots + slot;
            const int fo = ~(int)ord;
            if (fo > *reinterpret_cast<volatile int*>(f)) atomicMax(f, fo);
        } else {  // the CTA's table is full
            long long xl[kMaxAggs];
#pragma unroll
            for (int a = 0; a < NA; a++) xl[a] = A.agg[a].op == kAggCount ? 1ll : (long long)(A.agg[a].op == kAggMin ? ~x[a] : x[a]);
            agg_to_global(table, A.table_slots, overflow, key, (unsigned long long)row, xl, s_agg, naggs);
        }
    };

    asm volatile("griddepcontrol.wait;" ::: "memory");  // the filter kernel's bitmap and counts are final
    const long long span_stride = (long long)gridDim.x * kComputeWarps, span_first = (long long)blockIdx.x * kComputeWarps + warp;
    if (__ldcg(&ctrl->total) != 0ull) {
        unsigned short* const sel_w = sel_all + warp * 1024;
        const long long nspans = A.ntiles * 8;
        unsigned ord0 = 0;  // 1024 x the number of spans this warp has been through (< 2^31: a warp sees every (8 x grid)-th span)
        for (long long span = span_first; span < nspans; span += span_stride, ord0 += 1024u) {
            const unsigned n = __ldg(span_cnt + span);
            if (n == 0) continue;
            const uint32_t m = __ldg(bitmap + span * 32 + lane);
            __syncwarp();
            append_selection(m, lane, sel_w, 0u);
            __syncwarp();
            const long long row0 = span * 1024;
#pragma unroll
            for (int g = 0; g < NG; g++) gp[g] = A.group[g].base + row0 * A.group[g].width;
#pragma unroll
            for (int a = 0; a < NA; a++) ap[a] = A.agg[a].base + row0 * A.agg[a].width;
            // 64 selected rows per iteration (two steps of 32), software-pipelined: the loads of the NEXT iteration are issued
            // before this iteration's updates and consumed after them, so the global-memory latency hides behind the
            // shared-memory work
            constexpr int NGW = NG > 0 ? NG : 1;
            unsigned s0 = lane < n ? sel_w[lane] : 0u, s1 = 32u + lane < n ? sel_w[32 + lane] : 0u;
            uint32_t g0[NGW], g1[NGW], a0[NA], a1[NA];
            fetch_raw(s0, g0, a0);
            fetch_raw(s1, g1, a1);
            for (unsigned i0 = 0; i0 < n; i0 += 64) {
                const bool live0 = i0 + lane < n, live1 = i0 + 32 + lane < n;
                unsigned long long k0, k1;
                int x0[NA], x1[NA];
                decode(s0, g0, a0, k0, x0);
                decode(s1, g1, a1, k1, x1);
                const unsigned c0 = s0, c1 = s1;
                if (i0 + 64 < n) {  // (rows past the selection vector's end fetch row 0 of the span: harmless, never applied)
                    s0 = i0 + 64 + lane < n ? sel_w[i0 + 64 + lane] : 0u;
                    s1 = i0 + 96 + lane < n ? sel_w[i0 + 96 + lane] : 0u;
                    fetch_raw(s0, g0, a0);
                    fetch_raw(s1, g1, a1);
                }
                if (live0) apply(row0 + c0, ord0 + c0, k0, x0);
                if (live1) apply(row0 + c1, ord0 + c1, k1, x1);
            }
        }
    }
    // thread t folds slot t over the CTA's warps and hands the group to the global table
    __syncthreads();
    for (int i = tid; i < kAggSlots; i += kComputeThreads) {
        const unsigned long long key = keys[i];
        if (key == kAggEmpty) continue;
        unsigned long long first = ~0ull;
        int acc[NA];
#pragma unroll
        for (int a = 0; a < NA; a++) acc[a] = A.agg[a].op == kAggCount ? 0 : INT_MIN;
        for (int w = 0; w < kComputeWarps; w++) {
            const int* const W = vals + w * (NA + 1) * kAggSlots;
            const int fo = W[NA * kAggSlots + i];
            if (fo == INT_MIN) continue;  // warp w never saw this group
            const unsigned ord = (unsigned)~fo;
            const unsigned long long r = ((unsigned long long)((long long)blockIdx.x * kComputeWarps + w + (long long)(ord >> 10) * span_stride) << 10) | (ord & 1023u);
            first = r < first ? r : first;
#pragma unroll
            for (int a = 0; a < NA; a++) {
                if (NA == 8 && a >= naggs) break;
                const int v = W[a * kAggSlots + i];
                acc[a] = A.agg[a].op == kAggCount ? (int)((unsigned)acc[a] + (unsigned)v) : (v > acc[a] ? v : acc[a]);  // (a CTA counts < 2^32 rows)
            }
        }
        if (first == ~0ull) continue;
        long long xl[kMaxAggs];
#pragma unroll
        for (int a = 0; a < NA; a++) xl[a] = A.agg[a].op == kAggCount ? (long long)(unsigned int)acc[a] : (long long)(A.agg[a].op == kAggMin ? ~acc[a] : acc[a]);
        agg_to_global(table, A.table_slots, overflow, key, first, xl, s_agg, naggs);
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

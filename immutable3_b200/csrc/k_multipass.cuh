// k_multipass.cuh - the default dense pipeline: filter_kernel (K1) + offset scan, emit_kernel / emit_general_kernel (K3g), emit_stream_kernel (K3s), select_all_kernel
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// Multi-pass pipeline for unlimited (or large-LIMIT) queries on dense tables
//
//   K1 filter_kernel : persistent CTAs stream the filter columns through a deep TMA ring (tiles are
//                      statically strided over the CTAs: no ordering, no tickets), evaluate the
//                      conjunction and write the selection BITMAP (one word per lane, coalesced) plus
//                      the match count of every 1024-row span and of every tile.
//   K2 (tail of K1)  : the last CTA to finish K1 turns the tile counts into device-wide exclusive offsets
//                      (LIMIT clamp, total) - no separate launch.
//   K3 emit_kernel   : one warp per group of eight 1024-row spans, no inter-warp dependency at all: offset =
//                      tile offset + the counts of the earlier spans of the tile; popc/scan compaction of
//                      the bitmap words into a warp-private selection vector; cooperative, coalesced
//                      Project gather of the select-list columns.
// Every stage is embarrassingly parallel, so none of them can be held up by a slow CTA the way a
// chained single-pass scan is; the price is the bitmap round trip (1 bit/row written + read).
// =============================================================================================
struct FilterShared {
    unsigned long long mbar_full[kMaxFilterStages];
    unsigned long long mbar_empty[kMaxFilterStages];
    unsigned int tile_acc[kMaxFilterStages];  // per ring slot: [31:20] warps arrived, [19:0] rows selected
    unsigned int tile_id[kMaxFilterStages];   // tile held by a ring slot (kNoMoreTiles = the CTA is done)
    unsigned long long scan_warp[2 * kComputeWarps];  // double-buffered per scan round
    unsigned int is_last;
    uint8_t lits[kLitPoolBytes];  // MATCH literals of the plan
    FilterCol filter[kMaxFilterCols];
    ProjCol proj[kMaxProjCols];
};

// Exclusive scan of the tile counts by one CTA of kComputeThreads threads, 4096 counts per round.  Warp w owns 512
// consecutive counts of the round; lane l handles the count PAIRS l, l+32, ..., l+224 of them, so every load and every
// offset store of a warp is one fully coalesced access (256 B / 512 B) - with 16 consecutive counts per thread the
// 128-byte-strided stores cost ~1 us of LSU wavefronts per round.  Eight warp scans chain the pairs, one block-level
// exchange per round (double-buffered, one barrier) chains the warps; the next round's counts are in flight meanwhile.
// Also applies the LIMIT clamp to the total and sums the rows that live in dense tiles (emit-kernel choice).
__device__ __forceinline__ uint2 ldcg_v2_here(const uint32_t* p) {
    uint2 v;
    asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

__device__ void scan_tile_counts(FilterShared& S, const uint32_t* tile_cnt, unsigned long long* tile_off, long long ntiles,
                                 long long limit, ScanCtrl* ctrl, int dense_tile_rows = kDenseTileRowsPerWord,
                                 unsigned long long* dbg = nullptr) {
    constexpr int kRound = kComputeThreads * 16;  // counts per round
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long running = 0;
    unsigned long long dense = 0;  // rows selected in dense tiles (>= kDenseTileMinRows of 8192; this thread's share)
    const unsigned dense_min = dense_tile_rows == kDenseTileRowsPerWord ? (unsigned)kDenseTileMinRows : ((unsigned)dense_tile_rows + 31u) / 32u;
    const int my0 = warp * 512 + lane * 2;  // first count of this lane's pair 0 inside a round; pair j is 64 counts further
    uint2 nx[8];                            // the next round's counts, in flight while this round is scanned
#pragma unroll
    for (int j = 0; j < 8; j++) nx[j] = ldcg_v2_here(tile_cnt + my0 + 64 * j);
    unsigned round = 0;
    for (long long base = 0; base < ntiles; base += kRound, round++) {
        uint2 c[8];
#pragma unroll
        for (int j = 0; j < 8; j++) c[j] = nx[j];
        if (base + kRound < ntiles) {
#pragma unroll
            for (int j = 0; j < 8; j++) nx[j] = ldcg_v2_here(tile_cnt + base + kRound + my0 + 64 * j);
        }
        // pair sums -> exclusive position of every pair inside the warp's 512 counts
        unsigned excl_pair[8], carry = 0, dsum = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const long long i = base + my0 + 64 * j;  // padding holds stale counts
            if (i >= ntiles) c[j].x = 0;
            if (i + 1 >= ntiles) c[j].y = 0;
            dsum += (c[j].x >= dense_min ? c[j].x : 0u) + (c[j].y >= dense_min ? c[j].y : 0u);
            const unsigned ps = c[j].x + c[j].y;
            unsigned incl = ps;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += nb;
            }
            excl_pair[j] = carry + incl - ps;
            carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        dense += dsum;
        if (lane == 0) S.scan_warp[warp + 8 * (round & 1u)] = carry;  // the warp's 512 counts; double-buffered: one barrier per round
        bar_sync(1, kComputeThreads);
        if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[round & 7] = t; }
        unsigned long long wbase = running, total = 0;
#pragma unroll
        for (int w = 0; w < kComputeWarps; w++) {
            const unsigned long long ws = S.scan_warp[w + 8 * (round & 1u)];
            if (w < warp) wbase += ws;
            total += ws;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const long long i = base + my0 + 64 * j;
            if (i < ntiles) {  // (pairs: the arrays are padded, the entry after the last tile is rewritten below)
                ulonglong2 o;
                o.x = wbase + excl_pair[j];
                o.y = o.x + c[j].x;
                *reinterpret_cast<ulonglong2*>(tile_off + i) = o;
            }
        }
        running += total;
    }
    if (tid == 0) ctrl->dense_rows = 0;
    bar_sync(1, kComputeThreads);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dense += __shfl_xor_sync(0xFFFFFFFFu, dense, o);  // one atomic per warp, not per thread
    if (lane == 0 && dense) atomicAdd(&ctrl->dense_rows, dense);
    if (tid == 0) {
        tile_off[ntiles] = running;
        ctrl->total = running < (unsigned long long)limit ? running : (unsigned long long)limit;
    }
}

// The same exclusive scan as a kernel of its own, one CTA per 4096 counts: at 1 B rows there are 122 K tile counts and the
// last CTA of the filter kernel needs ~50 us for them (30 serial rounds on one SM) - a quarter of a C4 query.  Here every
// CTA scans one round's worth, publishes its sum as an epoch-tagged word and adds up the sums of the chunks before it
// (chunks are handed out by an atomic ticket, so a CTA only ever waits for CTAs that are already running).  ~5 us.
// Launched between the filter and the emit kernel as a programmatic dependent of the former when the table has more than
// kScanInlineMaxTiles tiles; the CTA of the last chunk writes the total, the LIMIT clamp and the dense-tile row count.
constexpr long long kScanInlineMaxTiles = 4 * kComputeThreads * 16;

struct ScanShared {
    unsigned long long s_sum[kComputeWarps], s_dense[kComputeWarps], s_prev[kComputeWarps], s_prevd[kComputeWarps];
    unsigned int chunk, s_nz[kComputeWarps], lbase;
};
// One chunk (4096 tile counts) of the offset scan, by a CTA of kComputeThreads threads (all of them call; contains barriers).
__device__ __forceinline__ void offset_scan_chunk(ScanShared& SS, long long chunk, const uint32_t* __restrict__ tile_cnt,
                                                  unsigned long long* __restrict__ tile_off, long long ntiles, long long limit, uint32_t epoch,
                                                  unsigned long long* __restrict__ partials, ScanCtrl* ctrl, unsigned int* __restrict__ tile_list) {
    constexpr int kRound = kComputeThreads * 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nchunks = (ntiles + kRound - 1) / kRound, base_i = chunk * kRound;
    const int my0 = warp * 512 + lane * 2;  // (same pair layout as scan_tile_counts: every access of a warp is coalesced)
    uint2 c[8];
#pragma unroll
    for (int j = 0; j < 8; j++) c[j] = ldcg_v2_here(tile_cnt + base_i + my0 + 64 * j);  // (the arrays are padded to whole rounds)
    unsigned excl_pair[8], carry = 0, dsum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const long long i = base_i + my0 + 64 * j;
        if (i >= ntiles) c[j].x = 0;
        if (i + 1 >= ntiles) c[j].y = 0;
        dsum += (c[j].x >= (unsigned)kDenseTileMinRows ? c[j].x : 0u) + (c[j].y >= (unsigned)kDenseTileMinRows ? c[j].y : 0u);
        const unsigned ps = c[j].x + c[j].y;
        unsigned incl = ps;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += nb;
        }
        excl_pair[j] = carry + incl - ps;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    dsum = __reduce_add_sync(0xFFFFFFFFu, dsum);
    // (block pipeline) the non-empty tiles of this chunk go on a list - in any order: a tile's place in the result comes from
    // tile_off - so that the emit kernel need not look at 977 K block counts to find the 1 % that have rows
    unsigned nz = 0, nz_before = 0;
    if (tile_list) {
#pragma unroll
        for (int j = 0; j < 8; j++) nz += (c[j].x != 0u) + (c[j].y != 0u);
        unsigned incl = nz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += nb;
        }
        nz_before = incl - nz;
        if (lane == 31) SS.s_nz[warp] = incl;
    }
    if (lane == 0) {
        SS.s_sum[warp] = carry;
        SS.s_dense[warp] = dsum;
    }
    bar_sync(2, kComputeThreads);
    if (tile_list && tid == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int w = 0; w < kComputeWarps; w++) tot += SS.s_nz[w];
        SS.lbase = tot ? atomicAdd(&ctrl->ticket2, tot) : 0u;
    }
    unsigned long long wbase = 0, chunk_total = 0, chunk_dense = 0;
#pragma unroll
    for (int w = 0; w < kComputeWarps; w++) {
        if (w < warp) wbase += SS.s_sum[w];
        chunk_total += SS.s_sum[w];
        chunk_dense += SS.s_dense[w];
    }
    constexpr unsigned long long kVal = (1ull << 40) - 1ull;
    const unsigned long long tag = (unsigned long long)(epoch & 0xFFFFFFu) << 40;
    if (tid == 0) {
        st_relaxed_u64(partials + 2 * chunk, tag | (chunk_total & kVal));
        st_relaxed_u64(partials + 2 * chunk + 1, tag | (chunk_dense & kVal));
    }
    // sums of the chunks before this one (tickets: they are all running or done); the final chunk also needs their dense rows
    const bool is_final = chunk == nchunks - 1;
    unsigned long long prev = 0, prevd = 0;
    for (int word = 0; word < (is_final ? 2 : 1); word++) {
        for (long long p = tid; p < chunk; p += kComputeThreads) {
            unsigned long long v = ld_relaxed_u64(partials + 2 * p + word);
            if ((v >> 40) != (tag >> 40)) {
                const uint64_t t0 = globaltimer_ns();
                unsigned spins = 0;
                while (((v = ld_relaxed_u64(partials + 2 * p + word)) >> 40) != (tag >> 40)) {
                    __nanosleep(20);
                    if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > kWatchdogNs) watchdog_trap(ctrl, 3);
                }
            }
            if (word == 0) prev += v & kVal;
            else prevd += v & kVal;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        prev += __shfl_xor_sync(0xFFFFFFFFu, prev, o);
        prevd += __shfl_xor_sync(0xFFFFFFFFu, prevd, o);
    }
    if (lane == 0) {
        SS.s_prev[warp] = prev;
        SS.s_prevd[warp] = prevd;
    }
    bar_sync(2, kComputeThreads);
    unsigned long long base = 0, based = 0;
#pragma unroll
    for (int w = 0; w < kComputeWarps; w++) {
        base += SS.s_prev[w];
        based += SS.s_prevd[w];
    }
    if (tile_list) {  // (SS.lbase was written before the barrier above)
        unsigned pos = SS.lbase + nz_before;
#pragma unroll
        for (int w = 0; w < kComputeWarps; w++) pos += w < warp ? SS.s_nz[w] : 0u;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const unsigned i = (unsigned)(base_i + my0 + 64 * j);
            IMM3_CHECK(ctrl, (long long)pos + 2 <= ntiles + 2, 9);  // the list holds at most every tile
            if (c[j].x != 0u) tile_list[pos++] = i;
            if (c[j].y != 0u) tile_list[pos++] = i + 1u;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const long long i = base_i + my0 + 64 * j;
        if (i < ntiles) {
            ulonglong2 o;
            o.x = base + wbase + excl_pair[j];
            o.y = o.x + c[j].x;
            *reinterpret_cast<ulonglong2*>(tile_off + i) = o;
        }
    }
    bar_sync(2, kComputeThreads);
    if (is_final && tid == 0) {
        const unsigned long long total = base + chunk_total;
        tile_off[ntiles] = total;
        ctrl->total = total < (unsigned long long)limit ? total : (unsigned long long)limit;
        ctrl->dense_rows = based + chunk_dense;
    }
}

__global__ void __launch_bounds__(kComputeThreads) offset_scan_kernel(const uint32_t* __restrict__ tile_cnt, unsigned long long* __restrict__ tile_off,
                                                                     long long ntiles, long long limit, uint32_t epoch,
                                                                     unsigned long long* __restrict__ partials, ScanCtrl* ctrl,
                                                                     unsigned int* __restrict__ tile_list) {
    __shared__ ScanShared SS;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the filter kernel's tile counts are final
    const int tid = threadIdx.x;
    if (tid == 0) SS.chunk = atomicAdd(&ctrl->ticket, 1u);
    __syncthreads();
    offset_scan_chunk(SS, (long long)SS.chunk, tile_cnt, tile_off, ntiles, limit, epoch, partials, ctrl, tile_list);
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&ctrl->exited, 1u);
        if (done == gridDim.x - 1) {  // last CTA out: the counters are the next query's again
            ctrl->exited = 0;
            ctrl->ticket = 0;
        }
    }
}

// K1: tile = 8192 rows = 8 spans, one per compute warp.  A producer warp streams the tiles of this CTA (statically
// strided: no ordering, no tickets) through a TMA ring `ring` tiles deep; the compute warps never synchronise with each
// other - each evaluates the conjunction on its span, stores its bitmap word and span count, and adds the count to the
// tile's total in shared memory; the warp that completes a tile writes the tile count.  The kernel is issue-bound, so
// everything that does not depend on the tile is hoisted out of the loop (the single-filter-column case keeps the
// whole predicate descriptor in registers) and the row-count mask is only built for the table's last tile.
template <bool STAGED>
__global__ void __launch_bounds__(kComputeThreads + 32, 4) filter_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmap,
                                                                           uint32_t* __restrict__ span_cnt, uint32_t* __restrict__ tile_cnt,
                                                                           unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl) {
    constexpr int kTile = kDenseTileRowsPerWord;
    __shared__ FilterShared S;
    const uint32_t ring_addr = smem_u32(dyn_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ring = STAGED ? P.stages : 2;
    const long long ntiles = P.ntiles;
    if (tid == 0) phase_stamp(P, 0);
    // Programmatic dependent launch: the emit kernel's CTAs may take over SMs as this grid's CTAs retire and run their
    // prologue; they block in griddepcontrol.wait until this whole grid (including the offset scan) has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    for (int i = tid; i < P.lit_bytes; i += kComputeThreads + 32) S.lits[i] = P.lits[i];  // (nothing to copy unless a MATCH predicate exists)
    copy_plan_tables(P, S.filter, S.proj, tid, kComputeThreads + 32);
    if (tid == 0) {
        for (int s = 0; s < kMaxFilterStages; s++) {
            mbar_init(smem_u32(&S.mbar_full[s]), 1);
            mbar_init(smem_u32(&S.mbar_empty[s]), kComputeWarps);
            S.tile_acc[s] = 0;
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kComputeWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint64_t pol_keep = l2_policy_evict_last();
            // Tiles are drawn from an atomic ticket (the CTAs finish within one tile of each other instead of one in twenty);
            // the next ticket is already in flight while this tile's copies are issued.
            unsigned t_next = atomicAdd(&ctrl->ticket, 1u);
            RingPos rp;
            for (;; rp.advance(ring)) {
                const long long tile = t_next;
                if (tile < ntiles) t_next = atomicAdd(&ctrl->ticket, 1u);
                const int slot = rp.slot;
                const unsigned use = rp.use;
                if (use > 0) mbar_wait(smem_u32(&S.mbar_empty[slot]), (use - 1) & 1u, nullptr);
                const uint32_t bar = smem_u32(&S.mbar_full[slot]);
                S.tile_id[slot] = tile < ntiles ? (unsigned)tile : kNoMoreTiles;
                if (tile >= ntiles) {
                    mbar_arrive(bar);
                    break;
                }
                if (STAGED) {
                    mbar_arrive_expect_tx(bar, (uint32_t)P.stage_bytes);
#pragma unroll 1
                    for (int i = 0; i < P.nfilter; i++) {
                        const FilterCol& f = S.filter[i];
                        const uint32_t bytes = (uint32_t)(kTile * f.width);
                        const uint32_t dst = ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes + (uint32_t)f.smem_off;
                        if ((P.debug & 8u) || !f.keep_l2) tma_load_1d(dst, f.base + tile * bytes, bytes, bar);
                        else tma_load_1d_hint(dst, f.base + tile * bytes, bytes, bar, pol_keep);
                    }
                } else {
                    mbar_arrive(bar);
                }
            }
        }
    } else {
        // ---------------- compute warps: warp w = span w of every tile ----------------
        const int nf = P.nfilter;
        const FilterCol f0 = S.filter[0];                          // the (very common) single-column predicate lives in registers
        const int cell = (warp * 1024 + lane * 32) * f0.width;     // this lane's 32 rows inside a tile of column 0
        const uint8_t* const lits0 = S.lits + f0.lit_off;
        const long long full_tiles = P.nrows / kTile;              // tiles below this index have no rows past the end
        uint32_t* const bm_w0 = bitmap + warp * 32 + lane;
        uint32_t* const sc_w0 = span_cnt + warp;
        // The loop is instantiated once per "filter program": the single-predicate forms that dominate in practice (range
        // on a TINYINT column in its three sign modes, range on an INT column) have the predicate inlined - no call, no
        // dispatch on kind per span; everything else takes the general body.
        auto consume = [&](auto prog_tag) {
        constexpr int PROG = decltype(prog_tag)::value;
        const int hi0 = f0.lo + (int)f0.span;
        for (RingPos rp;; rp.advance(ring)) {
            const int slot = rp.slot;
            mbar_wait(smem_u32(&S.mbar_full[slot]), rp.use & 1u, nullptr);
            const unsigned tile_u = S.tile_id[slot];
            if (tile_u == kNoMoreTiles) break;
            const long long tile = tile_u;
            uint32_t* const bm_w = bm_w0 + tile * (kComputeWarps * 32);
            uint32_t* const sc_w = sc_w0 + tile * kComputeWarps;
            const uint32_t stage_addr = ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes;
            uint32_t m = 0xFFFFFFFFu;
            if (tile >= full_tiles) {
                const long long left = P.nrows - (tile * kTile + warp * 1024 + lane * 32);
                m = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
            }
            if (PROG >= 1 && PROG <= 3) {
                m &= eval_i8<STAGED, PROG - 1>(stage_addr + (uint32_t)f0.smem_off + (uint32_t)cell, f0.base + tile * (kTile * f0.width) + cell, lane,
                                               f0.lo, hi0);
            } else if (PROG == 4) {
                const uint32_t cs = stage_addr + (uint32_t)f0.smem_off + (uint32_t)cell;
                const uint8_t* cg = f0.base + tile * (kTile * f0.width) + cell;
                uint32_t mask = 0;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int q = (c + lane) & 7;
                    const uint4 v = ld16<STAGED>(cs + 16u * q, cg + 16 * q);
                    mask |= range_i32_chunk(v, (uint32_t)f0.lo, f0.span) << (4 * q);
                }
                m &= mask;
            } else if (P.debug & 2u) {
                m = 0;  // timing experiment: stream the tiles, skip the predicate
            } else if (nf == 1) {
                m &= eval_filter_span<STAGED>(stage_addr + (uint32_t)f0.smem_off + (uint32_t)cell, f0.base + tile * (kTile * f0.width) + cell, f0.kind,
                                              f0.width, f0.lo, f0.span, f0.nlit, lits0, lane);
            } else {
#pragma unroll 1
                for (int i = 0; i < nf; i++) dense_eval_filter<1, STAGED>(S.filter[i], S.lits, stage_addr, tile * kTile, warp * 1024, lane, &m);
            }
            if (P.or_accumulate) m |= __ldcg(bm_w);  // real OR (imm3_query_begin_dnf): this term's rows join those of the terms before it
            *bm_w = m;
            const unsigned c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(m));
            if (lane == 0) {
                *sc_w = c;
                // tile total: [31:20] warps arrived, [19:0] rows selected; the eighth arrival publishes and clears
                const unsigned old = atomicAdd(&S.tile_acc[slot], c + (1u << 20));
                if ((old >> 20) == kComputeWarps - 1) {
                    tile_cnt[tile] = (old & 0xFFFFFu) + c;
                    S.tile_acc[slot] = 0;  // (nobody touches it again before this warp's arrival on `empty` below)
                }
                mbar_arrive(smem_u32(&S.mbar_empty[slot]));  // this warp is done with the slot's bytes
            }
            __syncwarp();
        }
        };
        int prog = 0;
        if (nf == 1 && !(P.debug & 2u)) {
            if (f0.kind == kFilterI8Range) prog = f0.lo >= 0 ? 1 : (f0.lo + (int)f0.span < 0 ? 2 : 3);
            else if (f0.kind == kFilterI32Range) prog = 4;
        }
        switch (prog) {
            case 1: consume(std::integral_constant<int, 1>{}); break;
            case 2: consume(std::integral_constant<int, 2>{}); break;
            case 3: consume(std::integral_constant<int, 3>{}); break;
            case 4: consume(std::integral_constant<int, 4>{}); break;
            default: consume(std::integral_constant<int, 0>{}); break;
        }
    }

    // The last CTA to finish turns the tile counts into device-wide offsets (saves a launch).  (Letting every emit CTA
    // derive the offsets of its own tiles instead was measured: slower, 12-18 us of dependent L2 round trips per CTA.)
    __syncthreads();
    if (tid == 0) {
        phase_stamp(P, 1);
        __threadfence();
        const unsigned prev = atomicAdd(&ctrl->exited, 1u);
        S.is_last = prev == gridDim.x - 1;
        if (S.is_last) {  // everybody has drawn its last ticket: reset the counters for the emit kernel and the next query
            ctrl->exited = 0;
            ctrl->ticket = 0;
            ctrl->ticket2 = 0;
        }
    }
    __syncthreads();
    if (S.is_last && warp < kComputeWarps && P.scan_inline) {  // (large tables: offset_scan_kernel takes over)
        __threadfence();
        if (tid == 0) phase_stamp(P, 2);
        scan_tile_counts(S, tile_cnt, tile_off, ntiles, P.limit, ctrl, kDenseTileRowsPerWord, ((P.debug & 16u) && P.trace) ? P.trace + 32 : nullptr);
        if (tid == 0) phase_stamp(P, 3);
    }
}

// A short selection vector (a few rows): every lane fetches ALL columns of its row before the first store, so the rows
// cost one global round trip instead of one per column.  Up to 4 columns of width 1, 2 or 4 (the caller checks).
__device__ __noinline__ void emit_rows_fused(const ProjCol* proj, int nproj, const unsigned short* sel_w, int n, int lane, long long row0,
                                                long long g0) {
    const uint32_t sel_addr = smem_u32(sel_w);
    for (int i0 = lane; i0 < n; i0 += 32) {
        const long long row = row0 + lds_cell<uint16_t>(sel_addr + 2u * (uint32_t)i0);
        uint32_t v[4];
#pragma unroll
        for (int pc = 0; pc < 4; pc++) {
            if (pc < nproj) {
                const int w = proj[pc].width;
                const uint8_t* src = proj[pc].base + row * w;
                v[pc] = w == 4 ? __ldg(reinterpret_cast<const uint32_t*>(src))
                               : (w == 2 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(src)) : (uint32_t)__ldg(src));
            }
        }
#pragma unroll
        for (int pc = 0; pc < 4; pc++) {
            if (pc < nproj) {
                const int w = proj[pc].width;
                uint8_t* dst = proj[pc].out + (g0 + i0) * w;
                if (w == 4) *reinterpret_cast<uint32_t*>(dst) = v[pc];
                else if (w == 2) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)v[pc];
                else *dst = (uint8_t)v[pc];
            }
        }
    }
}
__device__ __forceinline__ bool can_emit_fused(const ProjCol* proj, int nproj) {
    bool ok = nproj <= 4;
    for (int pc = 0; pc < nproj && pc < 4; pc++) ok = ok && (proj[pc].width == 4 || proj[pc].width == 2 || proj[pc].width == 1);
    return ok;
}

// Result class decided by K1's offset scan: dense = at least half of the selected rows live in tiles with at least
// kDenseTileMinRows selected rows (those tiles are streamed); otherwise the rows are thinly spread and the gather kernel is the better fit.
__device__ __forceinline__ int emit_class_dense(const ScanCtrl* ctrl) {
    const unsigned long long total = __ldcg(&ctrl->total), dense = __ldcg(&ctrl->dense_rows);
    return (total > 0 && dense * 2ull >= total) ? 1 : 0;
}

// K3, sparse results: one warp per group of 8 spans = one 8192-row tile.  Two kernels, chosen on the host by the select
// list:
//   emit_kernel          up to 4 columns of width 1/2/4 (can_emit_fused).
//     * The next group's metadata is always in flight: its span counts and tile offset in registers, its 256 bitmap
//       words on their way into a warp-private shared-memory buffer (LDGSTS, double-buffered), so a group exposes ONE
//       global round trip - its gathers.
//     * The eight spans' lane counts are scanned together (two 16-bit counts per register: 20 shuffles per group, not
//       40); surviving rows are appended span by span to one warp-private selection vector (flushed when the next span
//       would not fit) and gathered 128 at a time - each lane issues the loads of 4 rows x all columns before its first
//       store.
//   emit_general_kernel  any select list: span by span, full spans copied straight, long vectors column by column.
constexpr int kEmitWarpSmemBytes = 2048 + 2 * 1024;  // selection vector (1024 x u16) + two buffers of 256 bitmap words

__global__ void __launch_bounds__(kComputeThreads, IMM3_EMIT_MIN_BLOCKS) emit_kernel(const __grid_constant__ ScanPlan P, const uint32_t* __restrict__ bitmap,
                                                                 const uint32_t* __restrict__ span_cnt,
                                                                 const unsigned long long* __restrict__ tile_off, int spans_per_tile,
                                                                 long long nspans, int dense_off, const ScanCtrl* ctrl) {
    __shared__ struct { FilterCol filter[kMaxFilterCols]; ProjCol proj[kMaxProjCols]; } SE;
    copy_plan_tables(P, SE.filter, SE.proj, threadIdx.x, kComputeThreads);
    __syncthreads();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    (void)spans_per_tile;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // (a programmatic dependent of the filter kernel when it is the only emit kernel)
    if (__ldcg(&ctrl->total) == 0ull) return;            // nothing survived the predicates
    if (dense_off && emit_class_dense(ctrl)) return;     // the streaming emit kernel takes dense results
    const long long warp0 = (long long)blockIdx.x * kComputeWarps + warp, nwarps = (long long)gridDim.x * kComputeWarps;
    const long long ngroups = (nspans + 7) >> 3;
    const uint32_t sel_addr = smem_u32(dyn_smem + warp * kEmitWarpSmemBytes), bm_addr = sel_addr + 2048u;

    unsigned c_n = 0;
    unsigned long long toff_n = 0;
    auto load_group = [&](long long u, int buf) {  // independent loads, pinned in place
        const long long p0 = u * 8;
        c_n = 0;
        if (lane < 8 && p0 + lane < nspans) asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(c_n) : "l"(span_cnt + p0 + lane) : "memory");
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(toff_n) : "l"(tile_off + u) : "memory");
        // lane l copies words [8l, 8l+8) of the group = a quarter of span l/4
        const uint32_t dst = bm_addr + (uint32_t)buf * 1024u + (uint32_t)lane * 32u;
        if (p0 + (lane >> 2) < nspans) {
            const uint32_t* src = bitmap + p0 * 32 + lane * 8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u), "l"(src + 4) : "memory");
        } else {
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst + 16u), "r"(0u) : "memory");
        }
    };
    int buf = 0;
    if (warp0 < ngroups) load_group(warp0, 0);
#pragma unroll 1
    for (long long u = warp0; u < ngroups; u += nwarps, buf ^= 1) {
        const long long p0 = u * 8;  // first span of the group (a group = one 8192-row tile; spans_per_tile is 8)
        const unsigned c = c_n;
        const unsigned long long toff = toff_n;
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const uint32_t bm_cur = bm_addr + (uint32_t)buf * 1024u + (uint32_t)lane * 4u;  // word k of this lane: + 128 k
        if (u + nwarps < ngroups) load_group(u + nwarps, buf ^ 1);
        const unsigned in_group = __reduce_add_sync(0xFFFFFFFFu, lane < 8 ? c : 0u);
        if (in_group == 0) continue;
        long long g0 = (long long)toff;  // ordinal of the first surviving row not emitted yet
        if (g0 >= P.limit) continue;
        // ---- lane offsets of all eight spans in one go ----
        uint32_t pk[4], inc[4];
#pragma unroll
        for (int q = 0; q < 4; q++)
            inc[q] = pk[q] = (uint32_t)__popc(lds_cell<uint32_t>(bm_cur + 256u * q)) | ((uint32_t)__popc(lds_cell<uint32_t>(bm_cur + 256u * q + 128u)) << 16);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc[q], o);
                if (lane >= o) inc[q] += t;
            }
        }
        const long long row0 = p0 * 1024;
        uint32_t fill = 0;  // rows in the selection vector; the first of them is global ordinal g0
#pragma unroll 1
        for (int k = 0; k <= 8; k++) {
            uint32_t excl = 0, n_k = 0;
            if (k < 8) {
                const int q = k >> 1, sh = 16 * (k & 1);
                const uint32_t iq = q == 0 ? inc[0] : (q == 1 ? inc[1] : (q == 2 ? inc[2] : inc[3]));
                const uint32_t pq = q == 0 ? pk[0] : (q == 1 ? pk[1] : (q == 2 ? pk[2] : pk[3]));
                excl = ((iq - pq) >> sh) & 0xFFFFu;
                n_k = (__shfl_sync(0xFFFFFFFFu, iq, 31) >> sh) & 0xFFFFu;
                if (n_k == 0) continue;
            }
            if (k == 8 || fill + n_k > 1024u) {
                // ---- gather what the vector holds: 128 rows per round, all loads of a round before its first store ----
                __syncwarp();
                const int nn = (int)(P.limit - g0 < (long long)fill ? P.limit - g0 : (long long)fill);
#pragma unroll 1
                for (int b0 = 0; b0 < nn; b0 += 128) {
                    int idx[4];  // row within the group, -1 = no row
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const int i = b0 + lane + 32 * r;
                        idx[r] = i < nn ? (int)lds_cell<uint16_t>(sel_addr + 2u * (uint32_t)i) : -1;
                    }
                    uint32_t v[4][4];
#pragma unroll
                    for (int pc = 0; pc < 4; pc++) {
                        if (pc < P.nproj) {
                            const int w = SE.proj[pc].width;
                            const uint8_t* cbase = SE.proj[pc].base + row0 * w;
                            if (w == 4) {
#pragma unroll
                                for (int r = 0; r < 4; r++) v[r][pc] = idx[r] >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(cbase) + idx[r]) : 0u;
                            } else if (w == 1) {
#pragma unroll
                                for (int r = 0; r < 4; r++) v[r][pc] = idx[r] >= 0 ? (uint32_t)__ldg(cbase + idx[r]) : 0u;
                            } else {
#pragma unroll
                                for (int r = 0; r < 4; r++)
                                    v[r][pc] = idx[r] >= 0 ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(cbase) + idx[r]) : 0u;
                            }
                        }
                    }
#pragma unroll
                    for (int pc = 0; pc < 4; pc++) {
                        if (pc < P.nproj) {
                            const int w = SE.proj[pc].width;
                            uint8_t* obase = SE.proj[pc].out + (g0 + b0 + lane) * w;
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                if (idx[r] >= 0) {
                                    if (w == 4) reinterpret_cast<uint32_t*>(obase)[32 * r] = v[r][pc];
                                    else if (w == 1) obase[32 * r] = (uint8_t)v[r][pc];
                                    else reinterpret_cast<uint16_t*>(obase)[32 * r] = (uint16_t)v[r][pc];
                                }
                            }
                        }
                    }
                }
                __syncwarp();
                g0 += fill;
                fill = 0;
                if (k == 8 || g0 >= P.limit) break;
            }
            // ---- append span k ----
            uint32_t addr = sel_addr + 2u * (fill + excl);
            const uint32_t base = (uint32_t)k * 1024u + (uint32_t)lane * 32u;
            uint32_t rm = __brev(lds_cell<uint32_t>(bm_cur + 128u * (uint32_t)k));  // leading zeros = index of the lowest set bit of the word
            while (rm) {
                const int b = __clz((int)rm);
                sts_u16(addr, base + (uint32_t)b);
                addr += 2u;
                rm &= ~(0x80000000u >> b);
            }
            fill += n_k;
        }
    }
}

__global__ void __launch_bounds__(kComputeThreads, 2) emit_general_kernel(const __grid_constant__ ScanPlan P, const uint32_t* __restrict__ bitmap,
                                                                 const uint32_t* __restrict__ span_cnt,
                                                                 const unsigned long long* __restrict__ tile_off, int spans_per_tile,
                                                                 long long nspans, int dense_off, const ScanCtrl* ctrl) {
    __shared__ struct { FilterCol filter[kMaxFilterCols]; ProjCol proj[kMaxProjCols]; } SE;
    copy_plan_tables(P, SE.filter, SE.proj, threadIdx.x, kComputeThreads);
    __syncthreads();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned short* sel_w = reinterpret_cast<unsigned short*>(dyn_smem) + warp * 1024;
    (void)spans_per_tile;
    asm volatile("griddepcontrol.wait;" ::: "memory");  // (a programmatic dependent of the filter kernel when it is the only emit kernel)
    if (__ldcg(&ctrl->total) == 0ull) return;            // nothing survived the predicates
    if (dense_off && emit_class_dense(ctrl)) return;  // the streaming emit kernel takes dense results
    const long long warp0 = (long long)blockIdx.x * kComputeWarps + warp, nwarps = (long long)gridDim.x * kComputeWarps;

    // ---------------- one warp per group of 8 spans ----------------
    const bool fused_ok = can_emit_fused(SE.proj, P.nproj);
    const long long ngroups = (nspans + 7) >> 3;
    // The next group's metadata (8 span counts, tile offset, 8 bitmap words per lane - all independent loads, pinned in
    // place) is in flight while this group is emitted: a group costs one exposed round trip (its gathers), not three.
    auto load_u32 = [](const uint32_t* p) {
        uint32_t v;
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        return v;
    };
    unsigned c_n = 0;
    unsigned long long toff_n = 0;
    uint32_t mw_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    auto load_group = [&](long long u) {
        const long long p0 = u * 8;
        c_n = (lane < 8 && p0 + lane < nspans) ? load_u32(span_cnt + p0 + lane) : 0u;
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(toff_n) : "l"(tile_off + u) : "memory");
#pragma unroll
        for (int k = 0; k < 8; k++) mw_n[k] = p0 + k < nspans ? load_u32(bitmap + (p0 + k) * 32 + lane) : 0u;
    };
    if (warp0 < ngroups) load_group(warp0);
    for (long long u = warp0; u < ngroups; u += nwarps) {
        const long long p0 = u * 8;  // first span of the group (a group = one 8192-row tile; spans_per_tile is 8)
        const int k0 = 0;
        const unsigned c = c_n;
        const unsigned long long toff = toff_n;
        uint32_t mw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) mw[k] = mw_n[k];
        if (u + nwarps < ngroups) load_group(u + nwarps);
        const unsigned in_group = __reduce_add_sync(0xFFFFFFFFu, lane < 8 ? c : 0u);
        if (in_group == 0) continue;
        long long g0 = (long long)toff;  // ordinal of the group's first surviving row
        if (g0 >= P.limit) continue;
        int fill = 0;          // rows in the selection vector, first of them is global ordinal g0
        auto emit_group = [&](int nfill) {
            const int nn = (int)(P.limit - g0 < (long long)nfill ? P.limit - g0 : (long long)nfill);
            if (fused_ok && nn <= 128) emit_rows_fused(SE.proj, P.nproj, sel_w, nn, lane, p0 * 1024, g0);
            else emit_span_all(SE.proj, P.nproj, nullptr, sel_w, nn, lane, false, 0u, 0, p0 * 1024, g0);
        };
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int n = (int)__shfl_sync(0xFFFFFFFFu, c, k0 + k);
            if (n == 0) continue;
            if (n == 1024 || fill + n > 1024) {  // flush what has been gathered so far
                __syncwarp();
                if (fill && g0 < P.limit) emit_group(fill);
                __syncwarp();
                g0 += fill;
                fill = 0;
            }
            if (n == 1024) {
                if (g0 < P.limit) emit_span_full(SE.proj, P.nproj, lane, (p0 + k) * 1024, g0, (int)(P.limit - g0 < 1024 ? P.limit - g0 : 1024));
                g0 += 1024;
                continue;
            }
            uint32_t mm = mw[k];
            const unsigned cnt = (unsigned)__popc(mm);
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += nb;
            }
            unsigned o = (unsigned)fill + incl - cnt;
            const unsigned iters = __reduce_max_sync(0xFFFFFFFFu, cnt);
            for (unsigned it = 0; it < iters; it++) {
                if (mm) {
                    sel_w[o++] = (unsigned short)(k * 1024 + lane * 32 + __ffs(mm) - 1);
                    mm &= mm - 1u;
                }
            }
            fill += n;
        }
        __syncwarp();
        if (fill && g0 < P.limit) emit_group(fill);
        __syncwarp();
    }
}

// K3, dense results: a persistent TMA-ring kernel.  Tile = 8192 rows (8 spans, one per compute warp).  The producer
// warp knows every tile's match count and offset before it starts (K1 finished), so it prefetches, `ring` tiles
// ahead, exactly what the tile needs: its 256 bitmap words and 8 span counts, plus - for a tile with at least
// kDenseTileMinRows (4.9 %) selected rows - the tile of every projected column as TMA bulk copies (whole DRAM pages
// instead of one request per selected row; at that density most 128-byte lines would be fetched anyway).
// Sparse tiles gather their few rows straight from global memory.  Empty tiles cost one count load.
// The compute warps never wait on a global load of their own for a dense tile and never talk to each other.
// Straight copy of `nbytes` staged bytes (shared address sb, 4-byte aligned) to an arbitrarily aligned global address:
// the body goes out as aligned 32-bit words, assembled from two shared words when source and destination disagree.
__device__ __forceinline__ void copy_smem_to_global(uint32_t sb, uint8_t* dst, int nbytes, int lane) {
    const int head = (int)((4u - ((unsigned)(uintptr_t)dst & 3u)) & 3u);  // bytes before the first aligned word of dst
    if (lane < head && lane < nbytes) dst[lane] = (uint8_t)lds_u8(sb + (uint32_t)lane);
    const int nwords = nbytes > head ? (nbytes - head) >> 2 : 0;
    uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + head);
    const uint32_t sh = (uint32_t)head * 8u;
    for (int i = lane; i < nwords; i += 32) {
        const uint32_t a = sb + (uint32_t)i * 4u;  // source bytes [head + 4i, head + 4i + 4)
        uint32_t v = lds_cell<uint32_t>(a);
        if (head) v = __funnelshift_r(v, lds_cell<uint32_t>(a + 4u), sh);
        d4[i] = v;
    }
    const int done = head + nwords * 4;
    if (lane < nbytes - done) dst[done + lane] = (uint8_t)lds_u8(sb + (uint32_t)(done + lane));
}

constexpr int kMaxEmitStages = 4;
constexpr int kEmitHdrBytes = 1024 + 128;  // bitmap words + span counts (padded)

struct EmitShared {
    unsigned long long mbar_full[kMaxEmitStages];
    unsigned long long mbar_empty[kMaxEmitStages];
    long long off[kMaxEmitStages];     // result ordinal of the tile's first selected row
    unsigned int tile[kMaxEmitStages];
    unsigned int mode[kMaxEmitStages];  // 0 = no more tiles, 1 = gather from global, 2 = projected columns staged
    FilterCol filter[kMaxFilterCols];
    ProjCol proj[kMaxProjCols];
};

__global__ void __launch_bounds__(kComputeThreads + 32, 2) emit_stream_kernel(const __grid_constant__ ScanPlan P, const uint32_t* __restrict__ bitmap,
                                                                                const uint32_t* __restrict__ span_cnt,
                                                                                const uint32_t* __restrict__ tile_cnt,
                                                                                const unsigned long long* __restrict__ tile_off, long long nsub,
                                                                                int ring, int stage_bytes, int dense_mode, ScanCtrl* ctrl) {
    __shared__ EmitShared S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // One of the two emit kernels does the work, picked on the device from the match count (no host round trip).
    if (tid == 0) phase_stamp(P, 4);
    const uint32_t ring_addr = smem_u32(dyn_smem) + kComputeWarps * 1024 * 2;  // after the selection vectors
    copy_plan_tables(P, S.filter, S.proj, tid, kComputeThreads + 32);
    if (tid == 0) {
        for (int s = 0; s < kMaxEmitStages; s++) {
            mbar_init(smem_u32(&S.mbar_full[s]), 1);
            mbar_init(smem_u32(&S.mbar_empty[s]), kComputeWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // Everything above overlapped the tail of the filter kernel (programmatic dependent launch); its outputs - counts,
    // offsets, bitmap, result class - may only be read from here on.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (dense_mode >= 0 && emit_class_dense(ctrl) != dense_mode) return;  // (-1: take every result)
    const bool can_stage = stage_bytes > kEmitHdrBytes;

    if (warp == kComputeWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint64_t pol_stream = l2_policy_evict_first();
            RingPos rp;
            unsigned t_next = atomicAdd(&ctrl->ticket2, 1u);  // dynamic tile assignment; the next ticket is always in flight
            for (;;) {
                const long long tile = t_next;
                if (tile < nsub) t_next = atomicAdd(&ctrl->ticket2, 1u);
                unsigned mode = 0, c = 0;
                long long off = 0;
                if (tile < nsub) {
                    c = __ldcg(tile_cnt + tile);
                    off = (long long)__ldcg(tile_off + tile);
                    if (c == 0) continue;  // nothing selected: the compute warps never hear of this tile
                    if (off < P.limit) mode = (can_stage && c >= (unsigned)kDenseTileMinRows) ? 2u : 1u;  // else: LIMIT reached, stop
                }
                const int slot = rp.slot;
                const unsigned use = rp.use;
                if (use > 0) mbar_wait(smem_u32(&S.mbar_empty[slot]), (use - 1) & 1u, nullptr);
                S.off[slot] = off;
                S.tile[slot] = (unsigned)tile;
                S.mode[slot] = mode;
                const uint32_t bar = smem_u32(&S.mbar_full[slot]);
                if (mode == 0) {
                    mbar_arrive(bar);
                    break;
                }
                const uint32_t dst = ring_addr + (uint32_t)slot * (uint32_t)stage_bytes;
                mbar_arrive_expect_tx(bar, mode == 2 ? (uint32_t)(stage_bytes - 96) : 1024u + 32u);
                tma_load_1d(dst, bitmap + tile * 256, 1024u, bar);
                tma_load_1d(dst + 1024u, span_cnt + tile * 8, 32u, bar);
                if (mode == 2) {
#pragma unroll 1
                    for (int pc = 0; pc < P.nproj; pc++) {
                        const ProjCol& pj = S.proj[pc];
                        const uint32_t bytes = (uint32_t)(kDenseTileRowsPerWord * pj.width);
                        if ((P.debug & 8u) || pj.filter_idx >= 0)
                            tma_load_1d(dst + (uint32_t)kEmitHdrBytes + 8u * (uint32_t)pj.stage_off, pj.base + tile * bytes, bytes, bar);
                        else
                            tma_load_1d_hint(dst + (uint32_t)kEmitHdrBytes + 8u * (uint32_t)pj.stage_off, pj.base + tile * bytes, bytes, bar, pol_stream);
                    }
                }
                rp.advance(ring);
            }
        }
    } else {
        // ---------------- compute warps: warp w = span w of every tile ----------------
        unsigned short* sel_w = reinterpret_cast<unsigned short*>(dyn_smem) + warp * 1024;
        const bool fast_sparse = can_emit_fused(S.proj, P.nproj);  // sparse tiles: fused multi-column gather
        for (RingPos rp;; rp.advance(ring)) {
            const int slot = rp.slot;
            mbar_wait(smem_u32(&S.mbar_full[slot]), rp.use & 1u, nullptr);
            const unsigned mode = S.mode[slot];
            if (tid == 0 && rp.use == 0 && rp.slot == 0) phase_stamp(P, 5);
            if (mode == 0) break;
            const uint32_t stage = ring_addr + (uint32_t)slot * (uint32_t)stage_bytes;
            const long long tile_row0 = (long long)S.tile[slot] * kDenseTileRowsPerWord;
            uint32_t m, c;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(m) : "r"(stage + (uint32_t)(warp * 32 + lane) * 4u));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c) : "r"(stage + 1024u + (uint32_t)(lane & 7) * 4u));
            const unsigned n = __shfl_sync(0xFFFFFFFFu, c, warp);
            const unsigned before = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? c : 0u);
            const long long g0 = S.off[slot] + before;
            const long long room = P.limit - g0;
            if (n > 0 && room > 0) {
                const int nn = room < (long long)n ? (int)room : (int)n;
                const int span_row = warp * 1024;
                if (n == 1024u && mode != 2) {
                    emit_span_full(S.proj, P.nproj, lane, tile_row0 + span_row, g0, nn);
                } else if (n == 1024u) {
#pragma unroll 1
                    for (int pc = 0; pc < P.nproj; pc++) {
                        const int w = S.proj[pc].width;
                        copy_smem_to_global(stage + (uint32_t)kEmitHdrBytes + 8u * (uint32_t)S.proj[pc].stage_off + (uint32_t)(span_row * w),
                                            S.proj[pc].out + g0 * w, nn * w, lane);
                    }
                } else {
                    append_selection(m, lane, sel_w, 0u);
                    __syncwarp();
                    if (mode == 2) {
                        // staged tile: entry i of the vector -> out[g0 + i], four entries per lane per round, every column in the
                        // same round (one read of the vector)
                        const uint32_t sel_addr = smem_u32(sel_w);
                        const uint32_t cols = stage + (uint32_t)kEmitHdrBytes;
                        for (int i0 = lane; i0 < nn; i0 += 128) {
                            uint32_t r[4];
#pragma unroll
                            for (int k = 0; k < 4; k++) r[k] = i0 + 32 * k < nn ? lds_cell<uint16_t>(sel_addr + 2u * (uint32_t)(i0 + 32 * k)) : 0xFFFFFFFFu;
#pragma unroll 1
                            for (int pc = 0; pc < P.nproj; pc++) {
                                const int w = S.proj[pc].width;
                                const uint32_t sb = cols + 8u * (uint32_t)S.proj[pc].stage_off + (uint32_t)(span_row * w);
                                uint8_t* const ob = S.proj[pc].out + (g0 + i0) * w;
                                if (w == 4) {
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        if (r[k] != 0xFFFFFFFFu) reinterpret_cast<uint32_t*>(ob)[32 * k] = lds_cell<uint32_t>(sb + r[k] * 4u);
                                } else if (w == 1) {
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        if (r[k] != 0xFFFFFFFFu) ob[32 * k] = lds_cell<uint8_t>(sb + r[k]);
                                } else if (w == 2) {
#pragma unroll
                                    for (int k = 0; k < 4; k++)
                                        if (r[k] != 0xFFFFFFFFu) reinterpret_cast<uint16_t*>(ob)[32 * k] = lds_cell<uint16_t>(sb + r[k] * 2u);
                                } else {
                                    for (int k = 0; k < 4; k++)
                                        if (r[k] != 0xFFFFFFFFu)
                                            for (int b = 0; b < w; b++) ob[(32 * k) * w + b] = (uint8_t)lds_u8(sb + r[k] * (uint32_t)w + (uint32_t)b);
                                }
                            }
                        }
                    } else if (fast_sparse) {
                        emit_rows_fused(S.proj, P.nproj, sel_w, nn, lane, tile_row0 + span_row, g0);  // sparse tile: one round trip for all columns
                    } else {
#pragma unroll 1
                        for (int pc = 0; pc < P.nproj; pc++) {
                            const ProjCol& pj = S.proj[pc];
                            const int w = pj.width;
                            emit_col(sel_w, nn, lane, w, false, 0u, pj.base + (tile_row0 + span_row) * w, pj.out + g0 * w);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&S.mbar_empty[slot]));
        }
        if (tid == 0) phase_stamp(P, 6);
    }
}

// K1 of a query without predicates: every row is selected, so the bitmap (ones, masked past the last row), the span and
// tile counts and the tile offsets are known without reading a byte of the table.
__global__ void __launch_bounds__(kComputeThreads) select_all_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmap,
                                                                    uint32_t* __restrict__ span_cnt, uint32_t* __restrict__ tile_cnt,
                                                                    unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    constexpr int kTile = kDenseTileRowsPerWord;
    for (long long tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        const long long row0 = tile * kTile + (long long)tid * 32;
        const long long left = P.nrows - row0;
        bitmap[tile * (kTile / 32) + tid] = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
        const long long in_tile = P.nrows - tile * kTile < kTile ? P.nrows - tile * kTile : kTile;
        if (tid < 8) {
            const long long s = in_tile - tid * 1024;
            span_cnt[tile * 8 + tid] = (uint32_t)(s >= 1024 ? 1024 : (s <= 0 ? 0 : s));
        }
        if (tid == 8) tile_cnt[tile] = (uint32_t)in_tile;
        if (tid == 9) tile_off[tile] = (unsigned long long)(tile * kTile);
    }
    if (blockIdx.x == 0 && tid == 0) {
        ctrl->total = (unsigned long long)(P.nrows < P.limit ? P.nrows : P.limit);
        ctrl->dense_rows = (unsigned long long)P.nrows;
        ctrl->ticket = 0;
        ctrl->ticket2 = 0;
    }
}


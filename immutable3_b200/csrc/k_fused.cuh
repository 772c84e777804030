// k_fused.cuh - the fused single-pass dense kernel (scan_dense_kernel) and its scanner warp; behind IMM3_PATH=fused
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// Dense kernel (fused single pass)
//
// CTA = 8 compute warps + a producer warp + a scanner-candidate warp, several CTAs per SM.  A tile is
// 8192*W consecutive rows (W = 1, 2 or 4 bitmap words per lane): compute warp w owns rows
// [w*1024*W, (w+1)*1024*W) of the tile, split into W spans of 1024 rows in which lane l owns rows
// [32*l, 32*l+32) = one 32-bit word of the selection bitmap.
//
//   producer warp : draws tiles from an atomic ticket `ring` tiles ahead and starts their 1-D TMA bulk
//                   copies into the CTA's shared-memory ring (full/empty mbarriers per slot).
//   compute warps : per tile
//     1. decode + conjunctive filter: 128-bit shared-memory loads, SIMD-within-a-register compares
//        -> W bitmap words per lane in registers; popc + warp reduce -> tile count, PUBLISHED at once
//     2. bitmap word -> warp-private selection vector (popc scan), overlapping the offset hand-off
//     3. Project: entry i of the selection vector is gathered (filter columns from the staged tile,
//        other columns from global memory, four independent gathers per lane) and stored at
//        offset + rank: coalesced stores in canonical row order, LIMIT = clamp on the offset.
//   scanner warp  : ONE warp of the whole grid (elected by an atomic) turns the published tile counts into
//                   exclusive offsets, 256 tiles per round, and owns the LIMIT cut and the total.
// =============================================================================================
// ---- the scanner: one warp of the whole grid turns tile counts into exclusive offsets ----------
// Workers publish agg[tile] = count as soon as a tile is filtered; the scanner walks the tiles in order,
// 32*K status words per round (K consecutive tiles per lane, warp scan of the lane sums), and writes
// pre[tile] = rows selected in all earlier tiles.  A worker therefore waits one hand-off (its own word),
// however many tiles are in flight - a chained look-back would have every tile of a generation wait for
// the prefix to ripple through all of them.  The scanner also owns the LIMIT cut (`done`, Project.scala:73-77)
// and the total.
constexpr int kScanK = 8;  // tiles per lane per round: one 64-byte aligned group, four 128-bit loads

__device__ __forceinline__ void trace_stamp(const ScanPlan& P, long long tile, int ev) {
    if (P.trace) P.trace[tile * 8 + ev] = globaltimer_ns();
}
__device__ __forceinline__ void ld_relaxed_v2u64(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ void st_relaxed_v2u64(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

// The status arrays are padded to a whole round, so the scanner needs no bounds checks: words past the
// last tile are never published and are treated as empty tiles.  Loads are 64-byte aligned groups (four
// 128-bit loads per lane, all issued before the first use); progress is per TILE - a CTA may hold two
// tiles of one group (one in work, one prefetched), so waiting for whole groups could deadlock.
__device__ __noinline__ void scanner_loop(const ScanPlan& P, ScanCtrl* ctrl, const unsigned long long* agg, unsigned long long* pre,
                                          int lane) {
    const unsigned ep = P.epoch & 0x3FFFFFu;
    const unsigned long long want = ((unsigned long long)ep << 2) | kStateAggregate;
    const long long ntiles = P.ntiles;
    long long pos = 0;  // first tile without an offset yet
    unsigned long long running = 0;  // rows selected in tiles [0, pos)
    uint64_t t0 = 0;
    unsigned spins = 0;
    while (pos < ntiles) {
        const long long idx0 = (pos & ~(long long)(kScanK - 1)) + lane * kScanK;
        unsigned long long st[kScanK];
#pragma unroll
        for (int k = 0; k < kScanK; k += 2) ld_relaxed_v2u64(agg + idx0 + k, st[k], st[k + 1]);
        // leading entries of the lane that are settled: already scanned (< pos), published, or past the end
        unsigned open = 1, lane_valid = 0, lane_sum = 0;
        unsigned cnt[kScanK];
#pragma unroll
        for (int k = 0; k < kScanK; k++) {
            const long long idx = idx0 + k;
            const bool counted = idx >= pos && idx < ntiles;
            open &= (!counted || (st[k] & 0xFFFFFFull) == want) ? 1u : 0u;
            cnt[k] = (counted && open) ? (unsigned)(st[k] >> 24) : 0u;
            lane_valid += open;
            lane_sum += cnt[k];
        }
        const unsigned full = __ballot_sync(0xFFFFFFFFu, lane_valid == (unsigned)kScanK);
        const int fl = full == 0xFFFFFFFFu ? 32 : __ffs((int)~full) - 1;  // first lane with an unpublished tile
        const long long new_pos = fl == 32 ? idx0 - lane * kScanK + 32 * kScanK
                                           : idx0 - lane * kScanK + fl * kScanK + (long long)__shfl_sync(0xFFFFFFFFu, lane_valid, fl & 31);
        if (new_pos <= pos) {
            if (spins == 0) t0 = globaltimer_ns();
            if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > kWatchdogNs) watchdog_trap(ctrl, 2);
            __nanosleep(20);
            continue;
        }
        spins = 0;
        if (lane > fl) lane_sum = 0;  // (lane fl: cnt[] is already zero from its first unpublished tile on)
        unsigned incl = lane_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += nb;
        }
        if (lane <= fl) {
            unsigned long long base = running + (incl - lane_sum);
#pragma unroll
            for (int k = 0; k < kScanK; k++) {
                const long long idx = idx0 + k;
                if (idx >= pos && idx < new_pos && idx < ntiles) {
                    st_relaxed_u64(pre + idx, pack_status(ep, kStatePrefix, base));
                    trace_stamp(P, idx, 5);
                }
                base += cnt[k];
            }
        }
        running += __shfl_sync(0xFFFFFFFFu, incl, 31);
        pos = new_pos;
        if (running >= (unsigned long long)P.limit) break;  // every tile up to the LIMIT cut has its offset
    }
    if (lane == 0) {
        const bool cut = running >= (unsigned long long)P.limit;
        ctrl->total = cut ? (unsigned long long)P.limit : running;
        __threadfence();
        if (cut) atomicExch(&ctrl->done, 1u);  // after the offsets: a worker that sees `done` and no offset is beyond the cut
    }
}

// Exclusive offset of `tile` (one thread).  -1 = the LIMIT was reached before this tile.
__device__ __forceinline__ long long wait_prefix(const unsigned long long* pre, unsigned tile, uint32_t epoch, ScanCtrl* ctrl) {
    const unsigned ep = epoch & 0x3FFFFFu;
    uint64_t t0 = 0;
    unsigned spins = 0;
    for (;;) {
        const unsigned long long s = ld_relaxed_u64(pre + tile);
        const unsigned d = ld_relaxed_u32(&ctrl->done);
        if ((((s >> 2) & 0x3FFFFFu) == ep) && ((unsigned)(s & 3u) == kStatePrefix)) return (long long)(s >> 24);
        if (d) {
            __threadfence();
            const unsigned long long s2 = ld_relaxed_u64(pre + tile);
            if ((((s2 >> 2) & 0x3FFFFFu) == ep) && ((unsigned)(s2 & 3u) == kStatePrefix)) return (long long)(s2 >> 24);
            return -1;
        }
        if (spins == 0) t0 = globaltimer_ns();
        if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > kWatchdogNs) watchdog_trap(ctrl, 3);
        __nanosleep(32);
    }
}

template <bool STAGED>
__global__ void __launch_bounds__(kDenseThreads, IMM3_DENSE_MIN_BLOCKS) scan_dense_kernel(const __grid_constant__ ScanPlan P, ScanCtrl* ctrl,
                                                                                             unsigned long long* status) {
    constexpr int kSub = kDenseTileRowsPerWord;  // rows per sub-tile (one ring slot): 8 warps x 32 lanes x 32 rows
    __shared__ DenseShared S;
    // dynamic shared memory: [8 selection vectors of 1024 uint16][2 x NS x 256 bitmap words]
    //                        [8 warp-private spans of the projected columns][TMA ring of filter-column sub-tiles]
    const int NS = P.subtiles;          // sub-tiles per tile
    const int tile_rows = NS * kSub;    // rows per tile
    unsigned short* const sel_all = reinterpret_cast<unsigned short*>(dyn_smem);
    uint32_t* const bm_all = reinterpret_cast<uint32_t*>(dyn_smem + kComputeWarps * 1024 * 2);
    const uint32_t pstage_addr = smem_u32(dyn_smem) + kComputeWarps * 1024 * 2 + 2u * (uint32_t)NS * 256u * 4u;  // 8 x proj_stage_bytes
    const uint32_t ring_addr = pstage_addr + (uint32_t)kComputeWarps * (uint32_t)P.proj_stage_bytes;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ring = STAGED ? P.stages : 2;
    const unsigned ntiles = (unsigned)P.ntiles;
    const unsigned long long* agg = status;
    unsigned long long* pre = status + status_round_up(P.ntiles);
    const bool want_offsets = !P.bitmap && P.nproj > 0;

    for (int i = tid; i < P.lit_bytes; i += kDenseThreads) S.lits[i] = P.lits[i];  // (nothing to copy unless a MATCH predicate exists)
    copy_plan_tables(P, S.filter, S.proj, tid, kDenseThreads);
    if (tid == 0) {
        for (int s = 0; s < kMaxStages; s++) {
            mbar_init(smem_u32(&S.mbar_full[s]), 1);
            mbar_init(smem_u32(&S.mbar_empty[s]), kComputeWarps);
        }
        for (int w = 0; w < kComputeWarps; w++) mbar_init(smem_u32(&S.mbar_warp[w]), 1);
        fence_mbar_init();
        // Scanner election: the first CTA to get here.  On a full-size grid the scanner gets its SM to itself
        // (its own compute warps and the other CTAs of that SM retire at once): every tile of the grid waits
        // on this one warp, so it must not queue for issue slots behind two dozen ALU-bound warps.
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned old = atomicCAS(&ctrl->scanner, 0u, smid + 1u);
        const bool dedicate = gridDim.x >= 64u;
        S.role = old == 0u ? (dedicate ? kRoleScannerOnly : kRoleScannerAndWorker) : ((dedicate && old == smid + 1u) ? kRoleIdle : kRoleWorker);
    }
    __syncthreads();
    const unsigned role = S.role;

    if (role == kRoleIdle || (role == kRoleScannerOnly && warp != kComputeWarps + 1)) {
        // nothing to do: leave the SM to the scanner warp
    } else if (warp == kComputeWarps) {
        // ---------------- producer: tickets + TMA, `ring` sub-tiles ahead of the compute warps ----------------
        if (lane == 0) {
            RingPos rp;
            for (bool more = true; more;) {
                unsigned t = kNoMoreTiles;
                if (!ld_relaxed_u32(&ctrl->done)) t = atomicAdd(&ctrl->ticket, 1u);  // after LIMIT: stop drawing tiles
                if (t < ntiles) trace_stamp(P, t, 0);
                for (int sub = 0; sub < NS && more; sub++, rp.advance(ring)) {
                    const int slot = rp.slot;
                    const unsigned use = rp.use;
                    if (use > 0) mbar_wait(smem_u32(&S.mbar_empty[slot]), (use - 1) & 1u, ctrl);
                    S.tile[slot] = t;
                    const uint32_t bar = smem_u32(&S.mbar_full[slot]);
                    const long long row0 = (long long)t * tile_rows + (long long)sub * kSub;
                    if (t >= ntiles) {
                        mbar_arrive(bar);
                        more = false;
                    } else if (STAGED && row0 < P.nrows && !(P.debug & 4u)) {
                        mbar_arrive_expect_tx(bar, (uint32_t)P.stage_bytes);
#pragma unroll 1
                        for (int i = 0; i < P.nfilter; i++) {
                            const FilterCol& f = S.filter[i];
                            const uint32_t bytes = (uint32_t)(kSub * f.width);
                            tma_load_1d(ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes + (uint32_t)f.smem_off, f.base + row0 * f.width,
                                        bytes, bar);
                        }
                    } else {
                        mbar_arrive(bar);  // direct loads, or a sub-tile past the last row: nothing to stage
                    }
                }
            }
        }
    } else if (warp == kComputeWarps + 1) {
        // ---------------- scanner warp of the elected CTA: serves the whole grid ----------------
        if (role == kRoleScannerOnly || role == kRoleScannerAndWorker) scanner_loop(P, ctrl, agg, pre, lane);
    } else {
        // ---------------- compute warps ----------------
        unsigned short* sel_w = sel_all + warp * 1024;
        RingPos rp;
        uint32_t wparity = 0;  // phase of this warp's projected-span barrier
        for (unsigned j = 0;; j++) {
            const int e = (int)(j & 1u);
            uint32_t* bm = bm_all + e * NS * 256;
            unsigned tile = kNoMoreTiles;

            // ---- phase 1: stream the tile's sub-tiles: decode + conjunctive filter -> bitmap words + span counts ----
            for (int sub = 0; sub < NS; sub++, rp.advance(ring)) {
                const int slot = rp.slot;
                mbar_wait(smem_u32(&S.mbar_full[slot]), rp.use & 1u, ctrl);
                tile = S.tile[slot];
                if (tile >= ntiles) break;  // CTA-uniform; only ever at sub == 0
                if (tid == 0 && sub == 0) trace_stamp(P, tile, 1);
                const long long sub_row0 = (long long)tile * tile_rows + (long long)sub * kSub;
                const uint32_t stage_addr = ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes;
                const long long left = P.nrows - (sub_row0 + warp * 1024 + lane * 32);
                uint32_t m = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
                if (!(P.debug & 2u) && sub_row0 < P.nrows) {  // (a sub-tile past the last row has nothing staged)
                    #pragma unroll 1
                    for (int i = 0; i < P.nfilter; i++) dense_eval_filter<1, STAGED>(S.filter[i], S.lits, stage_addr, sub_row0, warp * 1024, lane, &m);
                } else if (P.debug & 2u) {
                    m = 0;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&S.mbar_empty[slot]));  // this warp is done with the slot's bytes
                if (P.bitmap) P.bitmap[((sub_row0 + warp * 1024) >> 5) + lane] = m;
                bm[(sub * kComputeWarps + warp) * 32 + lane] = m;  // span sub*8 + warp of the tile
                const unsigned c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(m));
                if (lane == 0) S.span_cnt[e][sub * kComputeWarps + warp] = c;
            }
            if (tile >= ntiles) break;
            bar_sync(1, kComputeThreads);

            // Warp w owns the tile's spans [w*NS, (w+1)*NS) = NS*1024 consecutive rows from here on.
            const int nspans = NS * kComputeWarps;  // <= 64
            const unsigned c0 = lane < nspans ? S.span_cnt[e][lane] : 0u;
            const unsigned c1 = lane + 32 < nspans ? S.span_cnt[e][lane + 32] : 0u;
            const int first = warp * NS;
            const unsigned tile_count = __reduce_add_sync(0xFFFFFFFFu, c0 + c1);
            const unsigned warp_base = __reduce_add_sync(0xFFFFFFFFu, (lane < first ? c0 : 0u) + (lane + 32 < first ? c1 : 0u));
            const bool mine0 = lane >= first && lane < first + NS, mine1 = lane + 32 >= first && lane + 32 < first + NS;
            const unsigned warp_total = __reduce_add_sync(0xFFFFFFFFu, (mine0 ? c0 : 0u) + (mine1 ? c1 : 0u));
            const bool any_full = __any_sync(0xFFFFFFFFu, (mine0 && c0 == 1024u) || (mine1 && c1 == 1024u));
            if (tid == 0) {
                st_relaxed_u64(status + tile, pack_status(P.epoch, kStateAggregate, tile_count));
                trace_stamp(P, tile, 2);
            }
            if (!want_offsets) continue;

            // ---- phase 2: Project ----
            // Sparse spans: their selected rows are appended to one selection vector and gathered straight from
            // global memory.  Dense spans (>= stream_min_cnt rows of 1024): nearly every sector of the span would
            // be touched anyway, so the span of every projected column is streamed into the warp's shared-memory
            // buffer with TMA bulk copies (full DRAM pages, no per-row requests) and gathered from there.
            const long long tile_row0 = (long long)tile * tile_rows;
            const int warp_row = first * 1024;
            const bool project = warp_total > 0;  // warp-uniform
            const unsigned stream_min = P.proj_stage_bytes > 0 ? (unsigned)P.stream_min_cnt : 1024u;
            const bool any_dense = __any_sync(0xFFFFFFFFu, (mine0 && c0 >= stream_min) || (mine1 && c1 >= stream_min));
            const bool prebuilt = project && warp_total <= 1024u && !any_dense;  // one vector for the warp's rows, built while the
            if (prebuilt) {                                                        // scanner resolves the tile's offset
                unsigned fill = 0;
                for (int s = 0; s < NS; s++) {
                    const unsigned cnt = S.span_cnt[e][first + s];
                    if (cnt) append_selection(bm[(first + s) * 32 + lane], lane, sel_w + fill, (unsigned)(s * 1024));
                    fill += cnt;
                }
                __syncwarp();
            }
            if (tid == 0) {
                S.excl[e] = (P.debug & 1u) ? (long long)tile * 1800 : wait_prefix(pre, tile, P.epoch, ctrl);
                trace_stamp(P, tile, 3);
            }
            bar_sync(1, kComputeThreads);
            const long long excl = S.excl[e];
            if (project && excl >= 0 && excl + warp_base < P.limit) {
                long long g0 = excl + warp_base;  // result ordinal of the first row of the vector
                auto flush = [&](unsigned fill) {
                    const long long room = P.limit - g0;
                    if (fill && room > 0)
                        emit_span_all(S.proj, P.nproj, S.filter, sel_w, room < (long long)fill ? (int)room : (int)fill, lane, false, 0u, warp_row, tile_row0, g0);
                    g0 += fill;
                };
                if (prebuilt) {
                    flush(warp_total);
                } else {
                    const uint32_t wbuf = pstage_addr + (uint32_t)warp * (uint32_t)P.proj_stage_bytes;
                    const uint32_t wbar = smem_u32(&S.mbar_warp[warp]);
                    unsigned fill = 0;
                    for (int s = 0; s < NS; s++) {
                        const unsigned cnt = S.span_cnt[e][first + s];
                        if (cnt == 0) continue;  // warp-uniform
                        const bool dense = cnt >= stream_min;
                        if (dense || fill + cnt > 1024u) {
                            __syncwarp();
                            flush(fill);
                            __syncwarp();  // the gathers are done with the vector
                            fill = 0;
                        }
                        const long long room = P.limit - g0;
                        if (room <= 0) break;
                        const long long span_row0 = tile_row0 + warp_row + s * 1024;
                        if (cnt == 1024u) {
                            emit_span_full(S.proj, P.nproj, lane, span_row0, g0, room < 1024 ? (int)room : 1024);
                            g0 += 1024;
                        } else if (dense) {
                            if (lane == 0) {
                                mbar_arrive_expect_tx(wbar, (uint32_t)P.proj_stage_bytes);
#pragma unroll 1
                                for (int pc = 0; pc < P.nproj; pc++) {
                                    const ProjCol& pj = S.proj[pc];
                                    tma_load_1d(wbuf + (uint32_t)pj.stage_off, pj.base + span_row0 * pj.width, (uint32_t)(1024 * pj.width), wbar);
                                }
                            }
                            append_selection(bm[(first + s) * 32 + lane], lane, sel_w, 0u);
                            __syncwarp();
                            mbar_wait(wbar, wparity, ctrl);
                            wparity ^= 1u;
                            const int n = room < (long long)cnt ? (int)room : (int)cnt;
#pragma unroll 1
                            for (int pc = 0; pc < P.nproj; pc++) {
                                const ProjCol& pj = S.proj[pc];
                                emit_col(sel_w, n, lane, pj.width, true, wbuf + (uint32_t)pj.stage_off, nullptr, pj.out + g0 * pj.width);
                            }
                            g0 += cnt;
                            __syncwarp();  // the gathers are done with the vector and with the buffer
                        } else {
                            append_selection(bm[(first + s) * 32 + lane], lane, sel_w + fill, (unsigned)(s * 1024));
                            fill += cnt;
                        }
                    }
                    __syncwarp();
                    flush(fill);
                }
            }
            __syncwarp();
            if (tid == 0) trace_stamp(P, tile, 4);
        }
    }
    __syncthreads();
    cta_exit(ctrl);
}


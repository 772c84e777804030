// k_blocks_filter.cuh - K1b: range predicates on the sorted-integer codec WITHOUT materialising the values
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// blocks_filter_kernel (round 2).  One warp per reference block of <= 1024 rows, lane m = mini-block m (32 values).
//
// Round 1 unpacked every value of every block (717 warp-instructions per block, 0.05 of the HBM roofline).  A range
// predicate does not need the values: inside a mini-block of width b < 32 the deltas are unsigned, so the values rise
// monotonically (mod 2^32) from  first = carry + d0  to  last = carry + d0 + rest,  where d0 is the first delta and rest
// the sum of the other 31.  `rest` comes straight from the packed words (b = 1: one POPC; b = 2, 4, 8: SWAR + IDP4A; any
// other width: a per-field sum, still without prefix sums or stores), a warp scan of the totals gives every mini-block's
// carry, and a mini-block whose [first, last] lies wholly inside or wholly outside the predicate's window gets its 32
// selection bits at once.  Only a mini-block that STRADDLES a window edge (at most two per window on a sorted column),
// a raw b = 32 mini-block (unsorted data) or one whose arithmetic could wrap is unpacked, cooperatively: a value per lane
// + a warp scan.  All comparisons are modular ((v - lo) <= span, uint32), exactly like the row-space kernels; the
// wrap-around cases fall back to the exact unpack, so the result is bit-identical to decode-then-compare.
//
// The encoded bytes of a tile (8 consecutive blocks = one contiguous byte range of the arena) and its metadata (9 row
// ordinals, 9 word offsets) are staged by a producer warp through a TMA ring: the compute warps never issue a global load.
// Super-block headers are walked speculatively in parallel (lane s reads header s at the position it would have if all
// earlier super-blocks had the shape of super-block 1; one vote validates the guess) with the serial walk as fallback.
// Outputs: per-block match count; the block's 32 bitmap words ONLY if 0 < count < rows (the emit kernel synthesises the
// words of fully selected blocks and never looks at empty ones); tile counts accumulated in shared memory (no global
// atomics, no memset); the last CTA turns them into offsets (same scan as the dense pipeline).
// =============================================================================================
constexpr int kBlkHdrBytes = 96 + 48 * kMaxPforCols + 32;  // row ordinals (80 B) + word offsets (48 B per encoded column), padded

// (bswap32, k_blocks_single.cuh: putInt is big-endian, PFORCodec.scala:22)

// Field `lane` (width b < 32) of the mini-block whose packed words start at W[pos] (raw big-endian words in shared memory).
__device__ __forceinline__ uint32_t pfor_field_of_lane(const uint32_t* W, int pos, int b, int lane) {
    const uint32_t off = (uint32_t)(lane * b);
    const uint32_t* p = W + pos + (off >> 5);
    return __funnelshift_r(bswap32(p[0]), bswap32(p[1]), off) & ((1u << b) - 1u);  // (shift taken mod 32; b == 0 -> mask 0)
}

// Selection word of lane m (rows 32m .. 32m+31 of the block, before the row-count mask) for  (v - lo) <= span.
// W: the block's words (count word first), nw of them without the 8 pad bytes; n: rows of the block.
__device__ __forceinline__ uint32_t pfor_range_word(const uint32_t* __restrict__ W, int nw, int n, uint32_t lo, uint32_t span, int lane) {
    const int packed = n & ~31, nmini = packed >> 5, nsuper = packed >> 7;
    const int q = lane & 3, k = lane >> 2;
    const uint32_t before = q == 0 ? 0u : (0x01010100u << (8 * (3 - q)));  // selects the widths of the mini-blocks ahead of q
    int ip = 1, mypos = 0, mybits = 0;
    // ---------------- header walk ----------------
    bool spec_ok = false;
    if (nsuper >= 2) {
        const uint32_t h0 = bswap32(W[1]);
        const int p1 = 2 + (int)__dp4a(h0, 0x01010101u, 0u);
        const uint32_t h1 = bswap32(W[p1 < nw ? p1 : nw]);
        const int S1 = (int)__dp4a(h1, 0x01010101u, 0u);
        int ps = lane == 0 ? 1 : p1 + (lane - 1) * (1 + S1);
        const bool inside = ps < nw;
        ps = inside ? ps : nw;
        const uint32_t hs = lane == 0 ? h0 : bswap32(W[ps]);
        const int Ss = (int)__dp4a(hs, 0x01010101u, 0u);
        // lane s's position is right if super-blocks 1 .. s-1 all have S1 payload words; lanes 1 .. nsuper-2 vouch for the next one
        spec_ok = __all_sync(0xFFFFFFFFu, lane >= nsuper || (inside && (lane == 0 || lane == nsuper - 1 || Ss == S1)));
        if (spec_ok) {
            const uint32_t myh = __shfl_sync(0xFFFFFFFFu, hs, k);
            const int pk = __shfl_sync(0xFFFFFFFFu, ps, k);
            mypos = pk + 1 + (int)__dp4a(myh, before, 0u);
            mybits = (int)((myh >> (24 - 8 * q)) & 0xFFu);
            ip = __shfl_sync(0xFFFFFFFFu, ps + 1 + Ss, nsuper - 1);
        }
    }
    if (!spec_ok) {
        uint32_t myh = 0;
#pragma unroll 1
        for (int s = 0; s < nsuper; s++) {
            const uint32_t h = bswap32(W[ip]);
            const int pos = ip + 1 + (int)__dp4a(h, before, 0u);
            mypos = k == s ? pos : mypos;
            myh = k == s ? h : myh;
            ip += 1 + (int)__dp4a(h, 0x01010101u, 0u);
        }
        mybits = (int)((myh >> (24 - 8 * q)) & 0xFFu);
    }
    for (int m = nsuper * 4; m < nmini; m++) {  // left-over mini-blocks carry their own header word
        const int b = (int)bswap32(W[ip++]);
        if (m == lane) { mypos = ip; mybits = b; }
        ip += b;
    }
    // ---------------- d0 / rest of every mini-block ----------------
    const bool active = lane < nmini;
    const bool raw = active && mybits >= 32;
    uint32_t d0 = 0, rest = 0;  // raw mini-blocks: d0 = their last value (the carry they hand on), rest unused
    bool unsafe = false;        // the sum of the deltas might not fit 32 bits: classify by unpacking
    const unsigned same = __match_any_sync(0xFFFFFFFFu, active ? mybits : -1 - lane);
    const unsigned vote = __reduce_max_sync(0xFFFFFFFFu, active ? ((unsigned)__popc(same) << 8) | (unsigned)mybits : 0u);
    const int bmode = (int)(vote & 0xFFu);
    const unsigned odd = __ballot_sync(0xFFFFFFFFu, active && mybits != bmode);
    const bool swar = bmode <= 2 || bmode == 4 || bmode == 8;
    if (nmini > 0 && __popc(odd) <= 4 && (swar || bmode >= 32)) {
        // the usual shape of a sorted column: one width everywhere except a few mini-blocks (the first one carries the
        // block's absolute start value in its first delta).  The odd ones are summed cooperatively, a field per lane.
        for (unsigned todo = odd; todo; todo &= todo - 1u) {
            const int m = __ffs((int)todo) - 1;
            const int bm = __shfl_sync(0xFFFFFFFFu, mybits, m), pm = __shfl_sync(0xFFFFFFFFu, mypos, m);
            if (bm >= 32) {
                if (lane == m) d0 = bswap32(W[pm + 31]);
            } else {
                const uint32_t f = pfor_field_of_lane(W, pm, bm, lane);
                const uint32_t fr = lane ? f : 0u;
                const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, fr), mx = __reduce_max_sync(0xFFFFFFFFu, fr);
                const uint32_t f0 = __shfl_sync(0xFFFFFFFFu, f, 0);
                if (lane == m) { d0 = f0; rest = sum; unsafe = mx >= (1u << 26); }
            }
        }
        if (active && mybits == bmode) {
            const uint32_t* wp = W + mypos;  // fields of these widths never straddle a byte: no byte swap needed for sums
            if (bmode == 1) {
                const uint32_t w = wp[0];
                d0 = (w >> 24) & 1u;
                rest = (uint32_t)__popc(w) - d0;
            } else if (bmode == 2) {
                const uint32_t w0 = wp[0], w1 = wp[1];
                d0 = (w0 >> 24) & 3u;
                rest = (uint32_t)(__popc(w0 & 0x55555555u) + __popc(w1 & 0x55555555u)) + 2u * (uint32_t)(__popc(w0 & 0xAAAAAAAAu) + __popc(w1 & 0xAAAAAAAAu)) - d0;
            } else if (bmode == 4) {
                uint32_t acc = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t w = wp[i];
                    acc = __dp4a((w & 0x0F0F0F0Fu) + ((w >> 4) & 0x0F0F0F0Fu), 0x01010101u, acc);
                }
                d0 = (wp[0] >> 24) & 15u;
                rest = acc - d0;
            } else if (bmode == 8) {
                uint32_t acc = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) acc = __dp4a(wp[i], 0x01010101u, acc);
                d0 = wp[0] >> 24;
                rest = acc - d0;
            } else if (bmode >= 32) {
                d0 = bswap32(wp[31]);
            }  // bmode == 0: every value equals the carry
        }
    } else if (active) {
        // any other shape: a per-lane walk over the 32 fields (sums only - still no prefix sums, no stores)
        if (raw) {
            d0 = bswap32(W[mypos + 31]);
        } else {
            const uint32_t mask = (1u << mybits) - 1u;
            uint32_t off = 0, mx = 0;
#pragma unroll 4
            for (int j = 0; j < 32; j++, off += (uint32_t)mybits) {
                const uint32_t* p = W + mypos + (off >> 5);
                const uint32_t f = __funnelshift_r(bswap32(p[0]), bswap32(p[1]), off) & mask;
                if (j == 0) d0 = f;
                else { rest += f; mx = f > mx ? f : mx; }
            }
            unsafe = mx >= (1u << 26);
        }
    }
    // ---------------- chain the mini-blocks: carry(m) = raw ? last raw value : carry(m-1) + d0 + rest ----------------
    uint32_t v = active ? (raw ? d0 : d0 + rest) : 0u;
    const unsigned any_raw = __ballot_sync(0xFFFFFFFFu, raw);
    if (any_raw) {
        unsigned f = raw ? 1u : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t pv = __shfl_up_sync(0xFFFFFFFFu, v, o);
            const unsigned pf = __shfl_up_sync(0xFFFFFFFFu, f, o);
            if (lane >= o) {
                if (!f) v += pv;
                f |= pf;
            }
        }
    } else {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t pv = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (lane >= o) v += pv;
        }
    }
    uint32_t base = __shfl_up_sync(0xFFFFFFFFu, v, 1);  // carry handed to this mini-block
    if (lane == 0) base = 0;                             // initvalue = 0 at every block
    // ---------------- classify ----------------
    uint32_t word = 0;
    bool need = false;
    if (active) {
        if (raw || unsafe) {
            need = true;
        } else {
            const uint32_t uf = base + d0 - lo, ul = uf + rest;  // first / last value in the window's frame
            if (ul < uf) need = true;                            // the run wraps past 2^32 in that frame
            else if (ul <= span) word = 0xFFFFFFFFu;             // uf <= ul <= span: all 32 rows pass
            else if (uf > span) word = 0u;                       // span < uf <= ul: none passes
            else need = true;                                    // a window edge falls inside the run
        }
    }
    for (unsigned todo = __ballot_sync(0xFFFFFFFFu, need); todo; todo &= todo - 1u) {
        const int m = __ffs((int)todo) - 1;
        const int bm = __shfl_sync(0xFFFFFFFFu, mybits, m), pm = __shfl_sync(0xFFFFFFFFu, mypos, m);
        const uint32_t cm = __shfl_sync(0xFFFFFFFFu, base, m);
        uint32_t val;
        if (bm >= 32) {
            val = bswap32(W[pm + lane]);  // raw: the values themselves
        } else {
            val = pfor_field_of_lane(W, pm, bm, lane);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, val, o);
                if (lane >= o) val += t;
            }
            val += cm;
        }
        const uint32_t wd = __ballot_sync(0xFFFFFFFFu, (val - lo) <= span);
        if (lane == m) word = wd;
    }
    // ---------------- var-byte remainder (n % 32 values): 7-bit groups, low first, the last byte of a value has bit 7 set ----------------
    if (n > packed) {
        uint32_t carry = __shfl_sync(0xFFFFFFFFu, v, (nmini + 31) & 31);
        if (nmini == 0) carry = 0;
        uint32_t tw = 0;
        if (lane == 0) {
            int wpos = ip, shb = 0, shift = 0;
            uint32_t acc = 0, cur = carry;
            for (int i = packed; i < n;) {
                const uint32_t c = bswap32(W[wpos]) >> shb;
                shb += 8;
                wpos += shb >> 5;
                shb &= 31;
                acc += (c & 127u) << shift;
                if (c & 128u) {
                    cur += acc;
                    if ((cur - lo) <= span) tw |= 1u << (i - packed);
                    i++;
                    acc = 0;
                    shift = 0;
                } else {
                    shift += 7;
                }
            }
        }
        tw = __shfl_sync(0xFFFFFFFFu, tw, 0);
        if (lane == nmini) word = tw;
    }
    return word;
}

__host__ __device__ constexpr int blk_filter_slot_bytes(int nstaged, int tile_cap_bytes) { return kBlkHdrBytes + nstaged * tile_cap_bytes; }

__global__ void __launch_bounds__(kComputeThreads + 32, 4) blocks_filter_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmapB,
                                                                                  uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ tile_cnt,
                                                                                  unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl,
                                                                                  long long nblocks, const unsigned int* __restrict__ work) {
    __shared__ FilterShared S;
    __shared__ PforCol s_pfor[kMaxPforCols];
    __shared__ uint32_t s_base[kMaxFilterStages][kMaxPforCols];  // per ring slot and encoded column: arena word that sits at the slot's data offset
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the emit kernel may become resident and set itself up
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < P.lit_bytes; i += kComputeThreads + 32) S.lits[i] = P.lits[i];  // (nothing to copy unless a MATCH predicate exists)
    copy_plan_tables(P, S.filter, S.proj, tid, kComputeThreads + 32);
    if (tid < kMaxPforCols) {
#pragma unroll
        for (int i = 0; i < kMaxPforCols; i++)
            if (tid == i) s_pfor[i] = P.pfor[i];
    }
    const int ring = P.stages;
    if (tid == 0) {
        for (int s = 0; s < kMaxFilterStages; s++) {
            mbar_init(smem_u32(&S.mbar_full[s]), 1);
            mbar_init(smem_u32(&S.mbar_empty[s]), kComputeWarps);
            S.tile_acc[s] = 0;
        }
        fence_mbar_init();
    }
    __syncthreads();
    const long long ntiles = P.ntiles;  // tiles of 8 blocks
    const int cap = P.blk_tile_bytes;   // bytes reserved per staged column in a slot
    const int slot_bytes = P.stage_bytes;
    // staged columns: the encoded columns that carry a predicate, in slot order
    const unsigned fmask = P.pfor_filter_mask;

    if (warp == kComputeWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            RingPos rp;
            const long long nwork = work ? (long long)work[0] : ntiles;  // pruned query: only the tiles blocks_prune_kernel listed
            for (long long k = blockIdx.x;; k += gridDim.x, rp.advance(ring)) {
                const int slot = rp.slot;
                if (rp.use > 0) mbar_wait(smem_u32(&S.mbar_empty[slot]), (rp.use - 1) & 1u, nullptr);
                const uint32_t bar = smem_u32(&S.mbar_full[slot]);
                const long long tile = k < nwork ? (work ? (long long)work[1 + k] : k) : ntiles;
                if (tile >= ntiles) {
                    S.tile_id[slot] = kNoMoreTiles;
                    mbar_arrive(bar);
                    break;
                }
                S.tile_id[slot] = (unsigned)tile;
                const uint32_t dst = smem_u32(dyn_smem) + (uint32_t)slot * (uint32_t)slot_bytes;
                const long long b0 = tile * kComputeWarps;
                const long long b8 = b0 + kComputeWarps < nblocks ? b0 + kComputeWarps : nblocks;
                uint32_t tx = 80u;
                uint32_t base_w[kMaxPforCols], nbytes[kMaxPforCols];
#pragma unroll
                for (int s = 0; s < kMaxPforCols; s++) {
                    base_w[s] = nbytes[s] = 0;
                    if (s < P.npfor && ((fmask >> s) & 1u)) {
                        const uint32_t wo0 = __ldg(s_pfor[s].word_off + b0), wo8 = __ldg(s_pfor[s].word_off + b8);
                        base_w[s] = wo0 & ~3u;  // 16-byte aligned source
                        uint32_t nb = ((wo8 - base_w[s]) * 4u + 15u) & ~15u;
                        nbytes[s] = nb < (uint32_t)cap ? nb : (uint32_t)cap;
                        s_base[slot][s] = base_w[s];
                        tx += 48u + nbytes[s];
                    }
                }
                mbar_arrive_expect_tx(bar, tx);
                tma_load_1d(dst, P.row_start + b0, 80u, bar);
                int at = 0;
#pragma unroll
                for (int s = 0; s < kMaxPforCols; s++) {
                    if (s < P.npfor && ((fmask >> s) & 1u)) {
                        tma_load_1d(dst + 96u + 48u * (uint32_t)s, s_pfor[s].word_off + b0, 48u, bar);
                        if (nbytes[s]) tma_load_1d(dst + (uint32_t)kBlkHdrBytes + (uint32_t)(at * cap), s_pfor[s].words + base_w[s], nbytes[s], bar);
                        at++;
                    }
                }
            }
        }
    } else {
        // ---------------- compute warps: warp w = block 8 * tile + w ----------------
        for (RingPos rp;; rp.advance(ring)) {
            const int slot = rp.slot;
            mbar_wait(smem_u32(&S.mbar_full[slot]), rp.use & 1u, nullptr);
            const unsigned tile_u = S.tile_id[slot];
            if (tile_u == kNoMoreTiles) break;
            const long long tile = tile_u;
            const long long blk = tile * kComputeWarps + warp;
            const uint8_t* const sl = dyn_smem + (size_t)slot * (size_t)slot_bytes;
            unsigned cnt = 0;
            if (blk < nblocks) {
                const unsigned long long* rs = reinterpret_cast<const unsigned long long*>(sl);
                const long long R0 = (long long)rs[warp];
                const int n = (int)((long long)rs[warp + 1] - R0);
                const int nwords = (n + 31) >> 5;
                uint32_t myword;  // lane w keeps bitmap word w of the block
                {
                    const int left = n - lane * 32;
                    myword = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
                }
#pragma unroll 1
                for (int fi = 0; fi < P.nfilter; fi++) {
                    const FilterCol f = S.filter[fi];
                    if (f.kind == kFilterI32Range && f.pfor_slot >= 0) {
                        const int s = f.pfor_slot;
                        const int at = __popc(fmask & ((1u << s) - 1u));
                        const uint32_t* wo = reinterpret_cast<const uint32_t*>(sl + 96 + 48 * s);
                        const uint32_t w0 = wo[warp], w1 = wo[warp + 1];
                        const uint32_t* W = reinterpret_cast<const uint32_t*>(sl + kBlkHdrBytes + at * cap) + (w0 - s_base[slot][s]);
                        myword &= pfor_range_word(W, (int)(w1 - w0) - 2, n, (uint32_t)f.lo, f.span, lane);
                    } else {
                        // dense column: row per lane (coalesced), one ballot per 32 rows
                        for (int w = 0; w < nwords; w++) {
                            const int i = w * 32 + lane;
                            bool pass = i < n;
                            if (f.kind == kFilterI32Range) {
                                const uint32_t v = pass ? __ldg(reinterpret_cast<const uint32_t*>(f.base) + R0 + i) : 0u;
                                pass = pass && ((v - (uint32_t)f.lo) <= f.span);
                            } else if (f.kind == kFilterI8Range) {
                                const int v = pass ? (int)(signed char)__ldg(f.base + R0 + i) : 0;
                                pass = pass && ((uint32_t)(v - f.lo) <= f.span);
                            } else {
                                bool hit = false;
                                if (pass) {
                                    const uint8_t* cell = f.base + (R0 + i) * f.width;
                                    for (int l = 0; l < f.nlit && !hit; l++) {
                                        bool eq = true;
                                        for (int bb = 0; bb < f.width; bb++) eq = eq && (__ldg(cell + bb) == S.lits[f.lit_off + l * f.width + bb]);
                                        hit = eq;
                                    }
                                }
                                pass = hit;
                            }
                            const uint32_t word = __ballot_sync(0xFFFFFFFFu, pass);
                            if (lane == w) myword &= word;
                        }
                    }
                }
                cnt = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(myword));
                if (cnt != 0u && cnt != (unsigned)n) bitmapB[blk * 32 + lane] = myword;  // (all / none: the count says it all)
                if (lane == 0) blk_cnt[blk] = cnt;
            }
            __syncwarp();
            if (lane == 0) {
                // tile total: [31:20] warps arrived, [19:0] rows selected; the eighth arrival publishes and clears
                const unsigned old = atomicAdd(&S.tile_acc[slot], cnt + (1u << 20));
                if ((old >> 20) == kComputeWarps - 1) {
                    tile_cnt[tile] = (old & 0xFFFFFu) + cnt;
                    S.tile_acc[slot] = 0;
                }
                mbar_arrive(smem_u32(&S.mbar_empty[slot]));
            }
        }
    }

    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&ctrl->exited, 1u);
        S.is_last = prev == gridDim.x - 1;
        if (S.is_last) {
            ctrl->exited = 0;
            ctrl->ticket2 = 0;  // (offset_scan_kernel counts the non-empty tiles of this query here)
        }
    }
    __syncthreads();
    if (S.is_last && warp < kComputeWarps && P.scan_inline) {
        __threadfence();
        scan_tile_counts(S, tile_cnt, tile_off, ntiles, P.limit, ctrl);
    }
}

// =============================================================================================
// blocks_filter_quad_kernel: the same decision procedure, mapped  lane = (block g of 4, super-block s of 8).
//
// A query whose only predicate is a range on ONE encoded column (C4) spends most of blocks_filter_kernel's ~350
// warp-instructions per block on work that is uniform across the warp (header walk, votes, scans, bookkeeping).  Here a warp
// takes FOUR blocks at a time: lane (g, s) owns super-block s (128 values = 4 mini-blocks) of block g, so every uniform step
// serves four blocks, each lane reads its own super-block header (no broadcast), the carry chain is a 3-step scan over 8
// lanes, and a lane leaves with four selection words = one 16-byte store.  Mini-blocks whose width is not 0/1/2/4/8 (the
// first one of every block of a sorted column: its first delta is the block's absolute start value) are summed by the 8
// lanes of their block together, 4 fields each; straddling / raw / possibly-wrapping mini-blocks are unpacked by the
// whole warp (rare).  A quad in which any block is irregular (rows not a multiple of 128: the 1-row tail block of a
// segment, odd block sizes) or whose header speculation fails is handed, block by block, to pfor_range_word above.
// CTA tile = 32 blocks (8 warps x 4); the 8-block tile counts of the offset scan are summed by warp pairs in shared memory.
// =============================================================================================
#ifndef IMM3_QUAD_MIN_BLOCKS
#define IMM3_QUAD_MIN_BLOCKS 4
#endif
constexpr int kQuadTileBlocks = 4 * kComputeWarps;           // blocks per CTA tile
constexpr int kQuadHdrBytes = 272 + 16 + 144 + 16;           // 34 row ordinals, 36 word offsets (both padded to 16 bytes)
constexpr int kQuadWoOff = 288;

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// Selection words of lane (g, s): word[q] = rows of mini-block 4s+q of block g.  wa: shared address of the block's word 0,
// nw: its words without the pad, nsuper: its super-blocks (rows / 128, >= 1).  Returns false (nothing decided) if the
// header speculation fails in any of the four blocks.
__device__ __forceinline__ uint32_t widths_before(int q) { return q == 0 ? 0u : (0x01010100u << (8 * (3 - q))); }  // IDP4A selector: widths of the mini-blocks ahead of q

__device__ __forceinline__ bool pfor_range_quad(uint32_t wa, int nw, int nsuper, uint32_t lo, uint32_t span, int lane, uint32_t (&word)[4]) {
    const int s = lane & 7;
    // ---------------- headers: lane s reads header s where it sits if super-blocks 1 .. s-1 have the shape of super-block 1 ----------------
    const uint32_t h0 = bswap32(lds32(wa + 4u));
    const int p1 = 2 + (int)__dp4a(h0, 0x01010101u, 0u);
    const uint32_t h1 = bswap32(lds32(wa + 4u * (uint32_t)(p1 < nw ? p1 : nw)));
    const int S1 = (int)__dp4a(h1, 0x01010101u, 0u);
    int ps = s == 0 ? 1 : p1 + (s - 1) * (1 + S1);
    const bool inside = ps < nw;
    ps = inside ? ps : nw;
    uint32_t hs = bswap32(lds32(wa + 4u * (uint32_t)ps));
    const int Ss = (int)__dp4a(hs, 0x01010101u, 0u);
    if (!__all_sync(0xFFFFFFFFu, s >= nsuper || (inside && (s == 0 || s == nsuper - 1 || Ss == S1)))) return false;
    const bool active = s < nsuper;
    if (!active) hs = 0;  // (lanes past the block's last super-block: no widths, no rows)
    // ---------------- d0 / rest of the lane's four mini-blocks.  Widths 0 and 1 - a sorted, dense column - inline and branch-free ----------------
    uint32_t d0[4], rest[4];
    unsigned slow = 0;
    {
        uint32_t a = wa + 4u * (uint32_t)(ps + 1);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t b = (hs >> (24 - 8 * q)) & 0xFFu;
            const uint32_t w = lds32(a);  // (read even when unused: the address stays inside the staged tile)
            const bool one = b == 1u;
            d0[q] = one ? ((w >> 24) & 1u) : 0u;
            rest[q] = one ? (uint32_t)__popc(w) - d0[q] : 0u;
            if (b > 1u) slow |= 1u << q;
            a += 4u * b;
        }
    }
    // ---------------- any other width: per lane (2, 4, 8: SWAR; 32: raw) or left to the block's 8 lanes together ----------------
    unsigned rawm = 0, pend = 0, unsafe = 0;  // bit q: raw (d0 = last value) / needs the cooperative sum / sum may not fit 32 bits
    while (slow) {
        const int q = __ffs((int)slow) - 1;
        slow &= slow - 1u;
        const int b = (int)((hs >> (24 - 8 * q)) & 0xFFu);
        if (b > 8 && b < 32) { pend |= 1u << q; continue; }
        const uint32_t a = wa + 4u * (uint32_t)(ps + 1 + (int)__dp4a(hs, widths_before(q), 0u));
        uint32_t x0 = 0, xr = 0;
        if (b >= 32) {
            x0 = bswap32(lds32(a + 4u * 31u));
            rawm |= 1u << q;
        } else if (b == 2) {  // fields of width 2, 4, 8 never straddle a byte: no byte swap needed for sums
            const uint32_t w0 = lds32(a), w1 = lds32(a + 4u);
            x0 = (w0 >> 24) & 3u;
            xr = (uint32_t)(__popc(w0 & 0x55555555u) + __popc(w1 & 0x55555555u)) + 2u * (uint32_t)(__popc(w0 & 0xAAAAAAAAu) + __popc(w1 & 0xAAAAAAAAu)) - x0;
        } else if (b == 4) {
            uint32_t acc = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t w = lds32(a + 4u * i);
                acc = __dp4a((w & 0x0F0F0F0Fu) + ((w >> 4) & 0x0F0F0F0Fu), 0x01010101u, acc);
            }
            x0 = (lds32(a) >> 24) & 15u;
            xr = acc - x0;
        } else if (b == 8) {
            uint32_t acc = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) acc = __dp4a(lds32(a + 4u * i), 0x01010101u, acc);
            x0 = lds32(a) >> 24;
            xr = acc - x0;
        } else {
            pend |= 1u << q;  // 3, 5, 6, 7
            continue;
        }
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k == q) { d0[k] = x0; rest[k] = xr; }
    }
    // ---------------- the 8 lanes of a block sum one such mini-block together, 4 fields each ----------------
    const unsigned gmask = 0xFFu << (lane & 24);
    for (;;) {
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, pend != 0u);
        if (!bal) break;
        const unsigned mine = bal & gmask;
        const int m = mine ? __ffs((int)mine) - 1 : lane;  // the lane of this block that is served in this round
        const int myq = pend ? __ffs((int)pend) - 1 : 0;   // every lane offers its lowest pending mini-block
        const int ob = (int)((hs >> (24 - 8 * myq)) & 0xFFu), op = ps + 1 + (int)__dp4a(hs, widths_before(myq), 0u);
        const int bm = __shfl_sync(0xFFFFFFFFu, ob, m), pm = __shfl_sync(0xFFFFFFFFu, op, m);
        uint32_t sum = 0, f0 = 0;
        bool big = false;
        if (mine) {
            const uint32_t mask = (1u << bm) - 1u;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t off = (uint32_t)((4 * s + i) * bm);
                const uint32_t a = wa + 4u * ((uint32_t)pm + (off >> 5));
                const uint32_t f = __funnelshift_r(bswap32(lds32(a)), bswap32(lds32(a + 4u)), off) & mask;
                if (i == 0) f0 = f;
                if (i == 0 && s == 0) continue;  // the first delta is not part of `rest`
                sum += f;
                big = big || f >= (1u << 26);
            }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        f0 = __shfl_sync(0xFFFFFFFFu, f0, lane & 24);  // field 0 sits in the block's first lane
        const unsigned bigs = __ballot_sync(0xFFFFFFFFu, big) & gmask;
        if (mine && lane == m) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k == myq) { d0[k] = f0; rest[k] = sum; }
            if (bigs) unsafe |= 1u << myq;
            pend &= pend - 1u;
        }
    }
    // ---------------- carry chain: within the lane, then a segmented scan over the 8 lanes of the block ----------------
    uint32_t c = 0;
    unsigned absf = 0;  // the lane's span contains a raw mini-block: its outgoing carry does not depend on the incoming one
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const bool r = (rawm >> q) & 1u;
        c = r ? d0[q] : c + d0[q] + rest[q];
        absf |= r ? 1u : 0u;
    }
    if (__any_sync(0xFFFFFFFFu, rawm != 0u)) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t pc = __shfl_up_sync(0xFFFFFFFFu, c, o, 8);
            const unsigned pa = __shfl_up_sync(0xFFFFFFFFu, absf, o, 8);
            if (s >= o) {
                if (!absf) c += pc;
                absf |= pa;
            }
        }
    } else {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t pc = __shfl_up_sync(0xFFFFFFFFu, c, o, 8);
            if (s >= o) c += pc;
        }
    }
    uint32_t base = __shfl_up_sync(0xFFFFFFFFu, c, 1, 8);
    if (s == 0) base = 0;  // initvalue = 0 at every block
    // ---------------- classify (branch-free): all rows / no row / has to be unpacked ----------------
    unsigned need = 0;
    const unsigned hard = rawm | unsafe;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t uf = base + d0[q] - lo, ul = uf + rest[q];  // first / last value of the run in the window's frame
        const bool mono = ul >= uf && !((hard >> q) & 1u);         // (no wrap past 2^32 in that frame)
        const bool all = mono && ul <= span, none = mono && uf > span;
        word[q] = all ? 0xFFFFFFFFu : 0u;
        if (!all && !none) need |= 1u << q;
        base = ((rawm >> q) & 1u) ? d0[q] : base + d0[q] + rest[q];
    }
    if (!active) {
        need = 0;
#pragma unroll
        for (int q = 0; q < 4; q++) word[q] = 0;
    }
    // ---------------- the few mini-blocks that have to be unpacked: the whole warp, a value per lane ----------------
    for (;;) {
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, need != 0u);
        if (!bal) break;
        const int m = __ffs((int)bal) - 1;
        const int myq = need ? __ffs((int)need) - 1 : 0;
        const int ob = (int)((hs >> (24 - 8 * myq)) & 0xFFu), op = ps + 1 + (int)__dp4a(hs, widths_before(myq), 0u);
        // carry handed to mini-block myq: the lane's incoming carry plus its earlier mini-blocks (a raw one restarts the chain)
        uint32_t ocarry = __shfl_up_sync(0xFFFFFFFFu, c, 1, 8);
        if (s == 0) ocarry = 0;
#pragma unroll
        for (int q = 0; q < 3; q++)
            if (q < myq) ocarry = ((rawm >> q) & 1u) ? d0[q] : ocarry + d0[q] + rest[q];
        const int bm = __shfl_sync(0xFFFFFFFFu, ob, m), pm = __shfl_sync(0xFFFFFFFFu, op, m);
        const uint32_t cm = __shfl_sync(0xFFFFFFFFu, ocarry, m), wam = __shfl_sync(0xFFFFFFFFu, wa, m);
        uint32_t val;
        if (bm >= 32) {
            val = bswap32(lds32(wam + 4u * (uint32_t)(pm + lane)));
        } else {
            const uint32_t off = (uint32_t)(lane * bm);
            const uint32_t a = wam + 4u * ((uint32_t)pm + (off >> 5));
            val = __funnelshift_r(bswap32(lds32(a)), bswap32(lds32(a + 4u)), off) & ((1u << bm) - 1u);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, val, o);
                if (lane >= o) val += t;
            }
            val += cm;
        }
        const uint32_t wd = __ballot_sync(0xFFFFFFFFu, (val - lo) <= span);
        if (lane == m) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (q == myq) word[q] = wd;
            need &= need - 1u;
        }
    }
    return true;
}

// The four blocks bi0 .. bi0 + 3 of the staged 32-block tile at shared address sl (global blocks blk0 ..): lane (g, s) decides
// super-block s of block g with pfor_range_quad; an irregular quad goes block by block through pfor_range_word.  Writes the
// blocks' counts and - for partially selected blocks - bitmap words; returns the rows selected in the quad (every lane).
__device__ __forceinline__ unsigned quad_decide(uint32_t sl, uint32_t ring_addr, uint32_t base_w, int bi0, long long blk0, long long nblocks,
                                                uint32_t lo, uint32_t span, int lane, uint32_t* __restrict__ bitmapB, uint32_t* __restrict__ blk_cnt) {
    const int g = lane >> 3, s = lane & 7;
    const int bi = bi0 + g;  // my block inside the staged tile
    const long long blk = blk0 + g;
    const bool exists = blk < nblocks;
    int n = 0, nw = 0;
    uint32_t wa = sl + (uint32_t)kQuadHdrBytes;
    if (exists) {
        const unsigned long long r0 = lds_cell<unsigned long long>(sl + 8u * (uint32_t)bi), r1 = lds_cell<unsigned long long>(sl + 8u * (uint32_t)bi + 8u);
        const uint32_t w0 = lds32(sl + (uint32_t)kQuadWoOff + 4u * (uint32_t)bi), w1 = lds32(sl + (uint32_t)kQuadWoOff + 4u * (uint32_t)bi + 4u);
        n = (int)(r1 - r0);
        nw = (int)(w1 - w0) - 2;
        wa += 4u * (w0 - base_w);
    }
    unsigned quad_cnt = 0;
    uint32_t word[4];
    bool fast = __all_sync(0xFFFFFFFFu, exists && n > 0 && (n & 127) == 0);
    if (fast) fast = pfor_range_quad(wa, nw, n >> 7, lo, span, lane, word);
    if (fast) {
        unsigned c = (unsigned)(__popc(word[0]) + __popc(word[1]) + __popc(word[2]) + __popc(word[3]));
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);  // rows selected in my block
        if (c != 0u && c != (unsigned)n)  // (lanes past the block's last super-block write zeros: the emit kernel reads all 32 words)
            *reinterpret_cast<uint4*>(bitmapB + blk * 32 + 4 * s) = make_uint4(word[0], word[1], word[2], word[3]);
        if (s == 0) blk_cnt[blk] = c;
        quad_cnt = c + __shfl_xor_sync(0xFFFFFFFFu, c, 8);
        quad_cnt += __shfl_xor_sync(0xFFFFFFFFu, quad_cnt, 16);
    } else {
        // irregular quad: block by block, lane m = mini-block m (pfor_range_word)
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            const int nk = __shfl_sync(0xFFFFFFFFu, n, 8 * k), nwk = __shfl_sync(0xFFFFFFFFu, nw, 8 * k);
            const uint32_t wak = __shfl_sync(0xFFFFFFFFu, wa, 8 * k);
            if (blk0 + k >= nblocks) break;
            const uint32_t* W = reinterpret_cast<const uint32_t*>(dyn_smem + (wak - ring_addr));
            const int left = nk - lane * 32;
            uint32_t mw = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
            mw &= pfor_range_word(W, nwk, nk, lo, span, lane);
            const unsigned c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(mw));
            if (c != 0u && c != (unsigned)nk) bitmapB[(blk0 + k) * 32 + lane] = mw;
            if (lane == 0) blk_cnt[blk0 + k] = c;
            quad_cnt += c;
        }
    }
    return quad_cnt;
}

__global__ void __launch_bounds__(kComputeThreads + 32, IMM3_QUAD_MIN_BLOCKS) blocks_filter_quad_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmapB,
                                                                                       uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ tile_cnt,
                                                                                       unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl,
                                                                                       long long nblocks, const unsigned int* __restrict__ work) {
    __shared__ FilterShared S;
    __shared__ uint32_t s_base[kMaxFilterStages];       // per ring slot: arena word that sits at the slot's data offset
    __shared__ unsigned int s_tacc[kMaxFilterStages][4];  // per ring slot: the four 8-block tile counts ([31:20] warps arrived)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ring = P.stages;
    if (tid == 0) {
        for (int s = 0; s < kMaxFilterStages; s++) {
            mbar_init(smem_u32(&S.mbar_full[s]), 1);
            mbar_init(smem_u32(&S.mbar_empty[s]), kComputeWarps);
            for (int k = 0; k < 4; k++) s_tacc[s][k] = 0;
        }
        fence_mbar_init();
    }
    __syncthreads();
    const long long ntiles8 = P.ntiles;                                         // 8-block tiles (offset scan, emit kernel)
    const long long nct = (nblocks + kQuadTileBlocks - 1) / kQuadTileBlocks;     // CTA tiles of 32 blocks
    const int slot_bytes = P.stage_bytes;
    const FilterCol f0 = P.filter[0];
    const PforCol pc = P.pfor[f0.pfor_slot < 0 ? 0 : (f0.pfor_slot == 0 ? 0 : (f0.pfor_slot == 1 ? 1 : (f0.pfor_slot == 2 ? 2 : 3)))];
    const uint32_t ring_addr = smem_u32(dyn_smem);

    if (warp == kComputeWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            RingPos rp;
            const long long nwork = work ? (long long)work[0] : nct;  // pruned query: only the tiles blocks_prune_kernel listed
            for (long long k = blockIdx.x;; k += gridDim.x, rp.advance(ring)) {
                const int slot = rp.slot;
                if (rp.use > 0) mbar_wait(smem_u32(&S.mbar_empty[slot]), (rp.use - 1) & 1u, nullptr);
                const uint32_t bar = smem_u32(&S.mbar_full[slot]);
                const long long ct = k < nwork ? (work ? (long long)work[1 + k] : k) : nct;
                if (ct >= nct) {
                    S.tile_id[slot] = kNoMoreTiles;
                    mbar_arrive(bar);
                    break;
                }
                S.tile_id[slot] = (unsigned)ct;
                const uint32_t dst = ring_addr + (uint32_t)slot * (uint32_t)slot_bytes;
                const long long b0 = ct * kQuadTileBlocks;
                const long long b1 = b0 + kQuadTileBlocks < nblocks ? b0 + kQuadTileBlocks : nblocks;
                const uint32_t wo0 = __ldg(pc.word_off + b0), wo1 = __ldg(pc.word_off + b1);
                const uint32_t base_w = wo0 & ~3u;  // 16-byte aligned source
                uint32_t nb = ((wo1 - base_w) * 4u + 15u) & ~15u;
                if (nb > (uint32_t)P.blk_tile_bytes) nb = (uint32_t)P.blk_tile_bytes;
                s_base[slot] = base_w;
                mbar_arrive_expect_tx(bar, 272u + 144u + nb);
                tma_load_1d(dst, P.row_start + b0, 272u, bar);
                tma_load_1d(dst + (uint32_t)kQuadWoOff, pc.word_off + b0, 144u, bar);
                tma_load_1d(dst + (uint32_t)kQuadHdrBytes, pc.words + base_w, nb, bar);
            }
        }
    } else {
        // ---------------- compute warps: warp w = blocks 32 * ct + 4 w .. + 3 ----------------
        const uint32_t lo = (uint32_t)f0.lo, span = f0.span;
        for (RingPos rp;; rp.advance(ring)) {
            const int slot = rp.slot;
            mbar_wait(smem_u32(&S.mbar_full[slot]), rp.use & 1u, nullptr);
            const unsigned ct_u = S.tile_id[slot];
            if (ct_u == kNoMoreTiles) break;
            const long long blk0 = (long long)ct_u * kQuadTileBlocks + 4 * warp;  // first block of this warp's quad
            const uint32_t sl = ring_addr + (uint32_t)slot * (uint32_t)slot_bytes;
            const unsigned quad_cnt = quad_decide(sl, ring_addr, s_base[slot], 4 * warp, blk0, nblocks, lo, span, lane, bitmapB, blk_cnt);
            __syncwarp();
            if (lane == 0) {
                // 8-block tile = this warp's quad + its neighbour's: the second arrival publishes and clears
                const long long t8 = blk0 >> 3;
                const unsigned old = atomicAdd(&s_tacc[slot][warp >> 1], quad_cnt + (1u << 20));
                if ((old >> 20) == 1u) {
                    if (t8 < ntiles8) tile_cnt[t8] = (old & 0xFFFFFu) + quad_cnt;
                    s_tacc[slot][warp >> 1] = 0;
                }
                mbar_arrive(smem_u32(&S.mbar_empty[slot]));
            }
        }
    }

    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&ctrl->exited, 1u);
        S.is_last = prev == gridDim.x - 1;
        if (S.is_last) {
            ctrl->exited = 0;
            ctrl->ticket2 = 0;  // (offset_scan_kernel counts the non-empty tiles of this query here)
        }
    }
    __syncthreads();
    if (S.is_last && warp < kComputeWarps && P.scan_inline) {
        __threadfence();
        scan_tile_counts(S, tile_cnt, tile_off, ntiles8, P.limit, ctrl);
    }
}

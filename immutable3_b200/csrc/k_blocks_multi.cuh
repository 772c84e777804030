// k_blocks_multi.cuh - the block pipeline of the sorted-integer codec: warp-per-block decode (pfor_decode_warp), blocks_emit_kernel (K3b)
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// Block-mode multi-pass pipeline (tables with blocks of <= 1024 rows; the sorted-integer codec's normal case)
//
//   K1b blocks_filter_kernel : one WARP per reference block (tile = 8 consecutive blocks = one CTA iteration).  The warp
//                              decodes the block of every PFOR_INT filter column in shared memory (lane m unpacks
//                              mini-block m: 32 funnel-shift extractions with a running delta sum; a segmented warp scan
//                              chains the mini-blocks, b = 32 mini-blocks restart the chain), evaluates the conjunction
//                              row-per-lane (ballot = one bitmap word per 32 rows) and stores the block's 32 bitmap
//                              words (block-local alignment), its match count, and the tile count.  The last CTA turns
//                              tile counts into offsets (same scan as the dense pipeline).
//   K3b blocks_emit_kernel   : one warp per NON-EMPTY block: offset = tile offset + counts of the tile's earlier blocks;
//                              PFOR columns of the select list are decoded again (only for blocks with matches - 1 % of
//                              them for C4), rows are emitted word by word: rank = popc(word & lanemask_lt).
// No look-back chain: the old single-pass block kernel spent 13 ns per block on it (97.6 K blocks per 100 M rows).
// =============================================================================================
constexpr int kBlkRows = 1024;        // largest block this pipeline takes
constexpr int kBlkLane = 36;          // decoded values: mini-block m lives at vals[36 m .. 36 m + 32): 16-byte aligned rows, so a lane
constexpr int kBlkVals = 32 * kBlkLane;  // moves its mini-block with 128-bit accesses (conflict-free per quarter warp), and the
                                      // row-per-lane view (emit) reads consecutive words
// per warp, in words: the emit kernel keeps the block's byte-swapped words, every decoded column of the select list, their
// mini-block bases and a 1024-entry selection vector  (the filter kernel, k_blocks_filter.cuh, stages raw tiles instead)
__host__ __device__ constexpr int blk_emit_warp_words(int npfor, int words_cap) { return words_cap + npfor * (kBlkVals + 32) + 512; }

// One mini-block of 32 B-bit deltas, B known at compile time: every word index and shift folds to an immediate.
template <int B>
__device__ __forceinline__ uint32_t unpack_fixed(const uint32_t* __restrict__ wp, uint32_t* __restrict__ vp) {
    uint32_t w[B > 0 ? B : 1];
#pragma unroll
    for (int i = 0; i < B; i++) w[i] = wp[i];
    uint32_t total = 0;
    uint32_t t4[4];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if (B > 0) {
            const int bit = j * B, wi = bit >> 5, sh = bit & 31;
            uint32_t d;
            if (sh + B <= 32) d = w[wi] >> sh;
            else d = __funnelshift_r(w[wi], w[wi + 1 < B ? wi + 1 : wi], sh);
            if (sh + B != 32) d &= (1u << B) - 1u;
            total += d;
        }
        t4[j & 3] = total;
        if ((j & 3) == 3) reinterpret_cast<uint4*>(vp)[j >> 2] = make_uint4(t4[0], t4[1], t4[2], t4[3]);  // (rows are 16-byte aligned)
    }
    return total;
}

// Decode one PFOR_INT block (n <= 1024 values) by one warp.  vals[kBlkLane m + j] + base(m) = value 32m+j,
// where base(m) is returned in lane m (mini-block-local prefix sums are stored; raw b = 32 mini-blocks and the
// var-byte tail store absolute values with base 0).
__device__ __forceinline__ uint32_t pfor_decode_warp(const uint32_t* __restrict__ words, uint32_t w0, uint32_t w1, int n, uint32_t* Wb,
                                                     int words_cap, uint32_t* vals, int lane) {
    int nw = (int)(w1 - w0) - 2;  // PFORCodecInt.encode appends 8 zero bytes (PFORCodec.scala:20)
    if (nw > words_cap - 2) nw = words_cap - 2;
    for (int i = lane; i < nw; i += 32) Wb[i] = __byte_perm(__ldg(words + w0 + i), 0, 0x0123);  // putInt is big-endian
    __syncwarp();
    const int packed = n & ~31, nmini = packed >> 5, nsuper = packed >> 7;
    // header walk: one word per 128-value super-block (four 8-bit widths, first mini-block in the top byte), then one word
    // per left-over mini-block.  Lane m picks up mini-block m: its width and where its words start (byte sums by IDP.4A).
    int ip = 1, mypos = 0, mybits = 0;
    {
        const int q = lane & 3, k = lane >> 2;
        const uint32_t before = q == 0 ? 0u : (0x01010100u << (8 * (3 - q)));  // selects the widths of the mini-blocks ahead of q
        uint32_t myh = 0;
#pragma unroll 1
        for (int s = 0; s < nsuper; s++) {
            const uint32_t h = Wb[ip];
            const int pos = ip + 1 + (int)__dp4a(h, before, 0u);
            mypos = k == s ? pos : mypos;  // (selects, not branches)
            myh = k == s ? h : myh;
            ip += 1 + (int)__dp4a(h, 0x01010101u, 0u);
        }
        mybits = (int)((myh >> (24 - 8 * q)) & 0xFFu);
        for (int m = nsuper * 4; m < nmini; m++) {
            const int b = (int)Wb[ip++];
            if (m == lane) { mypos = ip; mybits = b; }
            ip += b;
        }
    }
    // mini-block `lane`: 32 values
    uint32_t total = 0;
    const bool raw = mybits >= 32;
    const uint32_t* wp = Wb + mypos;
    uint32_t* vp = vals + lane * kBlkLane;
    // The usual shape of a sorted column's block: one width for (nearly) every mini-block, except the first one, whose
    // first delta carries the block's absolute start value.  The mini-blocks of the majority width (<= 16 bits) take the
    // fully specialised unpack; the few odd ones are decoded cooperatively first (a value per lane + a warp scan);
    // anything less regular takes the generic per-lane loop.
    const unsigned same = __match_any_sync(0xFFFFFFFFu, lane < nmini ? mybits : -1 - lane);
    const unsigned vote = __reduce_max_sync(0xFFFFFFFFu, lane < nmini ? ((unsigned)__popc(same) << 8) | (unsigned)mybits : 0u);
    const int bmode = (int)(vote & 0xFFu);
    const unsigned odd = __ballot_sync(0xFFFFFFFFu, lane < nmini && mybits != bmode);
    if (nmini > 0 && bmode <= 16 && __popc(odd) <= 4) {
        for (unsigned rest = odd; rest; rest &= rest - 1u) {
            const int m = __ffs((int)rest) - 1;
            const int bm = __shfl_sync(0xFFFFFFFFu, mybits, m), pm = __shfl_sync(0xFFFFFFFFu, mypos, m);
            uint32_t v;
            if (bm >= 32) {
                v = Wb[pm + lane];  // raw: the values themselves
            } else {
                const uint32_t off = (uint32_t)(lane * bm);
                const uint32_t* p = Wb + pm + (off >> 5);
                v = __funnelshift_r(p[0], p[1], off) & ((1u << bm) - 1u);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
                    if (lane >= o) v += t;
                }
            }
            vals[m * kBlkLane + lane] = v;
            const uint32_t last = __shfl_sync(0xFFFFFFFFu, v, 31);
            if (lane == m) total = last;
        }
        if (lane < nmini && mybits == bmode) {
            switch (bmode) {
#define IMM3_UNPACK_CASE(B) case B: total = unpack_fixed<B>(wp, vp); break;
                IMM3_UNPACK_CASE(0) IMM3_UNPACK_CASE(1) IMM3_UNPACK_CASE(2) IMM3_UNPACK_CASE(3) IMM3_UNPACK_CASE(4)
                IMM3_UNPACK_CASE(5) IMM3_UNPACK_CASE(6) IMM3_UNPACK_CASE(7) IMM3_UNPACK_CASE(8) IMM3_UNPACK_CASE(9)
                IMM3_UNPACK_CASE(10) IMM3_UNPACK_CASE(11) IMM3_UNPACK_CASE(12) IMM3_UNPACK_CASE(13) IMM3_UNPACK_CASE(14)
                IMM3_UNPACK_CASE(15) IMM3_UNPACK_CASE(16)
#undef IMM3_UNPACK_CASE
                default: break;
            }
        }
    } else if (lane < nmini) {
        if (raw) {
#pragma unroll
            for (int j = 0; j < 32; j++) vp[j] = total = wp[j];
        } else {
            const uint32_t mask = (1u << mybits) - 1u;
            uint32_t off = 0;
#pragma unroll
            for (int j = 0; j < 32; j++, off += (uint32_t)mybits) {
                const uint32_t* p = wp + (off >> 5);
                total += __funnelshift_r(p[0], p[1], off) & mask;  // (the shift amount is taken mod 32)
                vp[j] = total;
            }
        }
    }
    // chain the mini-blocks: carry(m) = raw ? last raw value : carry(m-1) + total   (segmented inclusive scan)
    uint32_t v = lane < nmini ? total : 0u;
    unsigned f = (lane < nmini && raw) ? 1u : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t pv = __shfl_up_sync(0xFFFFFFFFu, v, o);
        const unsigned pf = __shfl_up_sync(0xFFFFFFFFu, f, o);
        if (lane >= o) {
            if (!f) v += pv;
            f |= pf;
        }
    }
    uint32_t base = __shfl_up_sync(0xFFFFFFFFu, v, 1);  // carry after the previous mini-block
    if (lane == 0) base = 0;                             // initvalue = 0 at every block
    uint32_t carry_tail = __shfl_sync(0xFFFFFFFFu, v, (nmini + 31) & 31);  // carry after the last packed mini-block
    if (nmini == 0) carry_tail = 0;
    if (raw || lane >= nmini) base = 0;
    // var-byte remainder (n % 32 values): 7-bit groups, low first, the last byte of a value has bit 7 set
    if (n > packed && lane == 0) {
        int wpos = ip, shb = 0, shift = 0;
        uint32_t acc = 0, cur = carry_tail;
        for (int k = packed; k < n;) {
            const uint32_t c = Wb[wpos] >> shb;
            shb += 8;
            wpos += shb >> 5;
            shb &= 31;
            acc += (c & 127u) << shift;
            if (c & 128u) {
                cur += acc;
                vals[nmini * kBlkLane + (k - packed)] = cur;
                k++;
                acc = 0;
                shift = 0;
            } else {
                shift += 7;
            }
        }
    }
    __syncwarp();
    return base;
}

// Emit kernel of the block pipelines: one warp per reference block with at least one surviving row.
//   ROWSPACE = false : bitmap written by blocks_filter_kernel (32 words per block, block-local alignment); a block's first
//                      ordinal = tile offset + counts of the tile's earlier blocks.
//   ROWSPACE = true  : bitmap written by the DENSE filter kernel over the table's row space (no predicate touches an
//                      encoded column, so K1 never decodes anything): the block's bits start at bit R0 of that bitmap
//                      (funnel shift of two words per lane); first ordinal = tile offset + span counts + popc of the
//                      words of R0's span below R0.
// The block's surviving rows go to a warp-private selection vector; encoded columns of the select list are decoded once
// into shared memory; rows are emitted 128 per round, each lane fetching 4 rows x all columns before its first store.
template <bool ROWSPACE>
__global__ void __launch_bounds__(kComputeThreads, 3) blocks_emit_kernel(const __grid_constant__ ScanPlan P, const uint32_t* __restrict__ bitmap,
                                                                           const uint32_t* __restrict__ cnts,
                                                                           const unsigned long long* __restrict__ tile_off, long long nblocks,
                                                                           const ScanCtrl* ctrl) {
    __shared__ ProjCol s_proj[kMaxProjCols];
    __shared__ FilterCol s_filter[kMaxFilterCols];
    __shared__ PforCol s_pfor[kMaxPforCols];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    copy_plan_tables(P, s_filter, s_proj, tid, kComputeThreads);
    if (tid < kMaxPforCols) {
#pragma unroll
        for (int i = 0; i < kMaxPforCols; i++)
            if (tid == i) s_pfor[i] = P.pfor[i];
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // launched as a programmatic dependent of the filter kernel: its results are final now
    if (__ldcg(&ctrl->total) == 0ull) return;  // nothing survived the predicates
    uint32_t* const Wb = reinterpret_cast<uint32_t*>(dyn_smem) + warp * blk_emit_warp_words(P.npfor, P.blk_words_cap);
    uint32_t* const vals0 = Wb + P.blk_words_cap;
    uint32_t* const bases = vals0 + P.npfor * kBlkVals;  // [slot][mini-block]: what to add to the stored prefix sums
    unsigned short* const sel_w = reinterpret_cast<unsigned short*>(bases + P.npfor * 32);
    const uint32_t sel_addr = smem_u32(sel_w);
    const long long warp0 = (long long)blockIdx.x * kComputeWarps + warp, nwarps = (long long)gridDim.x * kComputeWarps;
    unsigned used_slots = 0;  // encoded columns of the select list
    bool fused_ok = P.nproj <= 4;
    for (int pc = 0; pc < P.nproj; pc++) {
        if (s_proj[pc].pfor_slot >= 0) used_slots |= 1u << s_proj[pc].pfor_slot;
        fused_ok = fused_ok && (s_proj[pc].width == 4 || s_proj[pc].width == 2 || s_proj[pc].width == 1);
    }
    // block metadata in ONE round trip: lanes 0,1 = row ordinals, lanes 2+2s, 3+2s = word offsets of encoded column s
    auto load_meta = [&](long long b) -> unsigned long long {
        unsigned long long m = 0;
        if (b < nblocks) {
            if (lane < 2) m = P.row_start[b + lane];
            else if (lane < 2 + 2 * P.npfor) m = s_pfor[(lane - 2) >> 1].word_off[b + (lane & 1)];
        }
        return m;
    };
    // A warp visits blocks warp0, warp0 + nwarps, ... (round robin, so that a clustered result spreads over all warps).
    // Row space: the next block's metadata (its bits are found through R0) is in flight while this one is handled.
    // Block-local bitmap: the counts of the warp's next 32 blocks are fetched in ONE round trip (a lane each) and only the
    // non-empty ones (1 % of them for C4) are visited; those alone fetch their metadata and their tile's counts.
    unsigned long long meta_n = ROWSPACE ? load_meta(warp0) : 0ull;
#pragma unroll 1
    for (long long it = 0;; it++) {
        unsigned todo;  // blocks of this iteration still to handle (row space: bit 0)
        const long long first = ROWSPACE ? warp0 + it * nwarps : warp0 + it * 32 * nwarps;  // block of lane 0 / of bit 0
        if (first >= nblocks) break;
        unsigned long long meta_it = 0;
        if (ROWSPACE) {
            meta_it = meta_n;
            meta_n = load_meta(first + nwarps);
            todo = 1u;
        } else {
            const long long b = first + lane * nwarps;
            todo = __ballot_sync(0xFFFFFFFFu, b < nblocks && __ldg(cnts + b) != 0u);
        }
#pragma unroll 1
        while (todo) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const long long blk = first + src * nwarps;
        const unsigned long long meta = ROWSPACE ? meta_it : load_meta(blk);
        unsigned tile_c = 0;                 // block-local: counts of the tile's blocks (lanes 0-7)
        unsigned long long tile_o = 0;
        if (!ROWSPACE) {
            const long long t8 = blk & ~7ll;
            tile_c = (lane < 8 && t8 + lane < nblocks) ? __ldg(cnts + t8 + lane) : 0u;
            tile_o = __ldg(tile_off + (blk >> 3));
        }
        const long long R0 = (long long)__shfl_sync(0xFFFFFFFFu, meta, 0);
        const int n = (int)((long long)__shfl_sync(0xFFFFFFFFu, meta, 1) - R0);
        uint32_t myword;  // lane w: rows [32w, 32w+32) of the block
        long long g;      // ordinal of the block's first surviving row
        if (ROWSPACE) {
            const long long bit0 = R0 + 32 * lane;
            const uint32_t lo = __ldg(bitmap + (bit0 >> 5)), hi = __ldg(bitmap + (bit0 >> 5) + 1);
            const long long span = R0 >> 10;
            const uint32_t sw = __ldg(bitmap + span * 32 + lane);                        // R0's span, word `lane`
            const unsigned sc = lane < (int)(span & 7) ? __ldg(cnts + (span & ~7ll) + lane) : 0u;  // earlier spans of the tile
            const unsigned long long toff = __ldg(tile_off + (span >> 3));
            myword = __funnelshift_r(lo, hi, (uint32_t)(bit0 & 31));
            const int left = n - lane * 32;
            myword &= left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
            if (__ballot_sync(0xFFFFFFFFu, myword != 0u) == 0u) continue;
            const long long wrow0 = (span << 10) + 32 * lane;  // first row of span word `lane`
            const unsigned below = wrow0 + 32 <= R0 ? (unsigned)__popc(sw) : (wrow0 < R0 ? (unsigned)__popc(sw & ((1u << (int)(R0 - wrow0)) - 1u)) : 0u);
            g = (long long)toff + __reduce_add_sync(0xFFFFFFFFu, sc + below);
        } else {
            g = (long long)tile_o + __reduce_add_sync(0xFFFFFFFFu, lane < (int)(blk & 7) ? tile_c : 0u);
            // the filter kernel stores the 32 words of a block only if SOME of its rows survive; all of them: the count says so
            const unsigned mycnt = __shfl_sync(0xFFFFFFFFu, tile_c, (int)(blk & 7));
            if (mycnt == (unsigned)n) {
                const int left = n - lane * 32;
                myword = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
            } else {
                myword = __ldg(bitmap + blk * 32 + lane);
            }
        }
        if (g >= P.limit) continue;
        __syncwarp();  // (the previous block's readers are done with the scratch)
        // decode the encoded columns of the select list
#pragma unroll 1
        for (int s = 0; s < P.npfor; s++) {
            if (!((used_slots >> s) & 1u)) continue;
            const uint32_t wo0 = (uint32_t)__shfl_sync(0xFFFFFFFFu, meta, 2 + 2 * s);
            const uint32_t wo1 = (uint32_t)__shfl_sync(0xFFFFFFFFu, meta, 3 + 2 * s);
            const uint32_t b = pfor_decode_warp(s_pfor[s].words, wo0, wo1, n, Wb, P.blk_words_cap, vals0 + s * kBlkVals, lane);
            bases[s * 32 + lane] = b;
            __syncwarp();
        }
        // selection vector of the block (ascending rows)
        const int cnt = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(myword));
        const int nn = (int)(P.limit - g < (long long)cnt ? P.limit - g : (long long)cnt);
        if (cnt == n) {
            // every row of the block survives (the inside of a window on a sorted column): no selection vector - decoded
            // columns go out row by row (a coalesced store per 32 rows), dense columns are copied straight
#pragma unroll 1
            for (int pc = 0; pc < P.nproj; pc++) {
                const int w = s_proj[pc].width, slot = s_proj[pc].pfor_slot;
                if (slot >= 0) {
                    const uint32_t* vs = vals0 + slot * kBlkVals;
                    const uint32_t* bs = bases + slot * 32;
                    uint32_t* o = reinterpret_cast<uint32_t*>(s_proj[pc].out) + g;
#pragma unroll 4
                    for (int i = lane; i < nn; i += 32) o[i] = vs[(i >> 5) * kBlkLane + lane] + bs[i >> 5];
                } else {
                    copy_rows(s_proj[pc].base + R0 * w, s_proj[pc].out + g * w, nn * w, lane);
                }
            }
            continue;
        }
        append_selection(myword, lane, sel_w, 0u);
        __syncwarp();
        if (fused_ok) {
#pragma unroll 1
            for (int b0 = 0; b0 < nn; b0 += 128) {
                int idx[4];  // row within the block, -1 = no row
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int i = b0 + lane + 32 * r;
                    idx[r] = i < nn ? (int)lds_cell<uint16_t>(sel_addr + 2u * (uint32_t)i) : -1;
                }
                uint32_t v[4][4];
#pragma unroll
                for (int pc = 0; pc < 4; pc++) {
                    if (pc < P.nproj) {
                        const int w = s_proj[pc].width, slot = s_proj[pc].pfor_slot;
                        if (slot >= 0) {
                            const uint32_t* vs = vals0 + slot * kBlkVals;
                            const uint32_t* bs = bases + slot * 32;
#pragma unroll
                            for (int r = 0; r < 4; r++) v[r][pc] = idx[r] >= 0 ? vs[(idx[r] >> 5) * kBlkLane + (idx[r] & 31)] + bs[idx[r] >> 5] : 0u;
                        } else {
                            const uint8_t* cbase = s_proj[pc].base + R0 * w;
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
                }
#pragma unroll
                for (int pc = 0; pc < 4; pc++) {
                    if (pc < P.nproj) {
                        const int w = s_proj[pc].width;
                        uint8_t* obase = s_proj[pc].out + (g + b0 + lane) * w;
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
        } else {
            // any select list: column by column, a row per lane
#pragma unroll 1
            for (int pc = 0; pc < P.nproj; pc++) {
                const int w = s_proj[pc].width, slot = s_proj[pc].pfor_slot;
                for (int i = lane; i < nn; i += 32) {
                    const int row = (int)lds_cell<uint16_t>(sel_addr + 2u * (uint32_t)i);
                    uint8_t* o = s_proj[pc].out + (g + i) * w;
                    if (slot >= 0) {
                        *reinterpret_cast<uint32_t*>(o) = vals0[slot * kBlkVals + (row >> 5) * kBlkLane + (row & 31)] + bases[slot * 32 + (row >> 5)];
                    } else {
                        const uint8_t* src = s_proj[pc].base + (R0 + row) * w;
                        for (int b = 0; b < w; b++) o[b] = __ldg(src + b);
                    }
                }
            }
        }
        }
    }
}


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

// A fully selected 1024-row block of the dense sorted shape - widths (B, 1, 1, 1), (1, 1, 1, 1) x 7, the shape the lane filter
// kernel takes (k_blocks_lane.cuh) - decoded straight from global memory into the result column, no shared-memory staging:
// four independent loads per lane (its header, its one-bit mini-block, the two words holding field `lane` of the wide
// mini-block), two warp scans (the wide mini-block's values, the mini-blocks' carries), then row 32 m + lane =
// carry(m) + popc(word(m) & low bits) - one coalesced store per mini-block.  B follows from the block's length and is
// checked against header 0.  out: the block's first result row; scratch: 64 warp-private words of shared memory.
struct DenseRegs {
    uint32_t hraw, nraw, x0, x1;
    int B;  // width of the wide mini-block if the block can have the dense shape, else -1
};
// ... the loads (issued as early as the block's word offsets are known: the caller overlaps them with the previous block)
__device__ __forceinline__ DenseRegs dense_issue(const uint32_t* __restrict__ words, uint32_t wo0, uint32_t wo1, int n, int lane) {
    DenseRegs r;
    r.hraw = r.nraw = r.x0 = r.x1 = 0;
    r.B = (int)(wo1 - wo0) - 42;
    if (n != 1024 || r.B < 0 || r.B > 31) {
        r.B = -1;
        return r;
    }
    const int B = r.B;
    const uint32_t* W = words + wo0;
    const int hidx = lane == 0 ? 1 : 5 + B + 5 * ((lane & 7) - 1 < 0 ? 0 : (lane & 7) - 1);
    const int nidx = 1 + B + 5 * (lane >> 2) + (lane & 3);  // (lane 0: unused)
    const uint32_t off = (uint32_t)(lane * B);
    r.hraw = __ldg(W + hidx);
    r.nraw = __ldg(W + nidx);
    r.x0 = __ldg(W + 2 + (off >> 5));
    r.x1 = __ldg(W + 3 + (off >> 5));
    return r;
}
// ... and the rest.  Returns false (nothing written) if the block does not have the dense shape.
__device__ __forceinline__ bool dense_finish(const DenseRegs& r, uint32_t* __restrict__ out, int nn, int lane, uint32_t* scratch, uint32_t dbg = 0) {
    if (r.B < 0) return false;
    const int B = r.B;
    const uint32_t off = (uint32_t)(lane * B);
    const uint32_t hexp = lane == 0 ? (0x01010100u | (uint32_t)B) : 0x01010101u;
    if (!__all_sync(0xFFFFFFFFu, lane >= 8 || r.hraw == hexp)) return false;
    uint32_t v = __funnelshift_r(__byte_perm(r.x0, 0, 0x0123), __byte_perm(r.x1, 0, 0x0123), off) & ((1u << B) - 1u);  // field `lane`
    const uint32_t X = __byte_perm(r.nraw, 0, 0x0123);
    const uint32_t wide_total = __reduce_add_sync(0xFFFFFFFFu, v);       // the wide mini-block's last value
    const uint32_t tot = lane == 0 ? wide_total : (uint32_t)__popc(X);  // what mini-block `lane` adds to the running value
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // (two independent scans, interleaved: the wide mini-block's values, the mini-blocks' carries)
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o), u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) {
            v += t;
            incl += u;
        }
    }
    const uint32_t carry = incl - tot;  // value of the row before this lane's mini-block
    const uint32_t low = (2u << lane) - 1u;
    uint2* const pair = reinterpret_cast<uint2*>(scratch);  // (word, carry) of every mini-block: one broadcast LDS.64 per 32 rows
    pair[lane] = make_uint2(X, carry);
    __syncwarp();
    uint32_t* const o = out + lane;
    if (dbg & 64u) return true;  // timing experiment (wrong results): no stores
    if (nn >= 1024) {
        o[0] = v;
#pragma unroll
        for (int m = 1; m < 32; m++) {
            const uint2 xc = pair[m];
            o[32 * m] = xc.y + (uint32_t)__popc(xc.x & low);
        }
    } else {
        if (lane < nn) o[0] = v;
#pragma unroll 4
        for (int m = 1; m < 32; m++) {
            const uint2 xc = pair[m];
            if (32 * m + lane < nn) o[32 * m] = xc.y + (uint32_t)__popc(xc.x & low);
        }
    }
    __syncwarp();
    return true;
}

// ... or kept in a form that gives ANY row's value in O(1), for blocks of which only some rows are wanted (a scattered
// predicate on another column: 10 % of the rows of every block): area[0 .. 31] = the wide mini-block's values, area[32 + 2m],
// area[33 + 2m] = mini-block m's word and the value before it.  No unpacking of 1024 values, no 4.6 KB of stores.
__device__ __forceinline__ bool dense_prepare(const DenseRegs& r, uint32_t* area, int lane) {
    if (r.B < 0) return false;
    const int B = r.B;
    const uint32_t off = (uint32_t)(lane * B);
    const uint32_t hexp = lane == 0 ? (0x01010100u | (uint32_t)B) : 0x01010101u;
    if (!__all_sync(0xFFFFFFFFu, lane >= 8 || r.hraw == hexp)) return false;
    uint32_t v = __funnelshift_r(__byte_perm(r.x0, 0, 0x0123), __byte_perm(r.x1, 0, 0x0123), off) & ((1u << B) - 1u);  // field `lane`
    const uint32_t X = __byte_perm(r.nraw, 0, 0x0123);
    const uint32_t wide_total = __reduce_add_sync(0xFFFFFFFFu, v);
    const uint32_t tot = lane == 0 ? wide_total : (uint32_t)__popc(X);
    uint32_t incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o), u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) {
            v += t;
            incl += u;
        }
    }
    area[lane] = v;
    reinterpret_cast<uint2*>(area + 32)[lane] = make_uint2(X, incl - tot);
    __syncwarp();
    return true;
}
__device__ __forceinline__ uint32_t dense_value(const uint32_t* area, int row) {
    const int m = row >> 5, j = row & 31;
    const uint2 xc = reinterpret_cast<const uint2*>(area + 32)[m];
    const uint32_t v = xc.y + (uint32_t)__popc(xc.x & ((2u << j) - 1u));
    return m == 0 ? area[j] : v;
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
__global__ void __launch_bounds__(kComputeThreads, ROWSPACE ? 3 : 2) blocks_emit_kernel(const __grid_constant__ ScanPlan P, const uint32_t* __restrict__ bitmap,
                                                                           const uint32_t* __restrict__ cnts, const unsigned int* __restrict__ tile_list,
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
    if (lane == 0) phase_stamp(P, 8);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // launched as a programmatic dependent of the filter kernel: its results are final now
    if (lane == 0) phase_stamp(P, 9);
    if (__ldcg(&ctrl->total) == 0ull) return;  // nothing survived the predicates
    if (P.debug & 128u) return;                // timing experiment (no results): launch + prologue alone
    uint32_t* const Wb = reinterpret_cast<uint32_t*>(dyn_smem) + warp * blk_emit_warp_words(P.npfor, P.blk_words_cap);
    uint32_t* const vals0 = Wb + P.blk_words_cap;
    uint32_t* const bases = vals0 + P.npfor * kBlkVals;  // [slot][mini-block]: what to add to the stored prefix sums
    unsigned short* const sel_w = reinterpret_cast<unsigned short*>(bases + P.npfor * 32);
    const uint32_t sel_addr = smem_u32(sel_w);
    const long long warp0 = (long long)blockIdx.x * kComputeWarps + warp, nwarps = (long long)gridDim.x * kComputeWarps;
    unsigned used_slots = 0;  // encoded columns of the select list
    unsigned ent0 = 0, ent1 = 0, ent2 = 0, ent3 = 0;  // per encoded column: the select-list entries that project it
    bool fused_ok = P.nproj <= 4;
    for (int pc = 0; pc < P.nproj; pc++) {
        if (s_proj[pc].pfor_slot >= 0) used_slots |= 1u << s_proj[pc].pfor_slot;
        const int sl = s_proj[pc].pfor_slot;
        ent0 |= sl == 0 ? 1u << pc : 0u;
        ent1 |= sl == 1 ? 1u << pc : 0u;
        ent2 |= sl == 2 ? 1u << pc : 0u;
        ent3 |= sl == 3 ? 1u << pc : 0u;
        fused_ok = fused_ok && (s_proj[pc].width == 4 || s_proj[pc].width == 2 || s_proj[pc].width == 1);
    }
    // One block with surviving rows: R0 / n = its first row / rows, myword = lane w's selection word (rows 32w ..), g = ordinal of
    // its first surviving row, wo(slot, k) = first (k = 0) / end (k = 1) word of the block in encoded column `slot`.
    auto emit_block = [&](long long R0, int n, uint32_t myword, long long g, auto wo, int prim, const DenseRegs& pre) {
        if (g >= P.limit) return;
        __syncwarp();  // (the previous block's readers are done with the scratch)
        // selection vector of the block (ascending rows)
        const int cnt = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(myword));
        const int nn = (int)(P.limit - g < (long long)cnt ? P.limit - g : (long long)cnt);
        IMM3_CHECK(ctrl, nn >= 0 && (unsigned long long)(g + nn) <= __ldcg(&ctrl->total) && n > 0 && n <= kBlkRows, 7);  // the block's rows fit the result
        // decode the encoded columns of the select list (every row survives: only those emit_dense_block does not take)
        unsigned direct = 0;       // encoded columns already written by emit_dense_block
        unsigned dense_slots = 0;  // encoded columns held in dense_prepare's form (partially selected blocks)
        if (cnt == n) {
#pragma unroll 1
            for (int pc = 0; pc < P.nproj; pc++) {
                const int slot = s_proj[pc].pfor_slot;
                if (slot < 0) continue;
                bool done;
                if (pc == prim) {  // (its loads were issued while the previous block was being written)
                    done = dense_finish(pre, reinterpret_cast<uint32_t*>(s_proj[pc].out) + g, nn, lane, Wb, P.debug);
                } else {
                    const uint32_t wo0 = wo(slot, 0), wo1 = wo(slot, 1);
                    const DenseRegs r = dense_issue(s_pfor[slot].words, wo0, wo1, n, lane);
                    done = dense_finish(r, reinterpret_cast<uint32_t*>(s_proj[pc].out) + g, nn, lane, Wb);
                }
                if (done) direct |= 1u << pc;
            }
        }
#pragma unroll 1
        for (int s = 0; s < P.npfor; s++) {
            if (!((used_slots >> s) & 1u)) continue;
            {
                const unsigned ent = s == 0 ? ent0 : (s == 1 ? ent1 : (s == 2 ? ent2 : ent3));
                if (cnt == n && (ent & ~direct) == 0u) continue;  // every select-list entry of this column was written directly
            }  // every select-list entry of this column was written directly
            const uint32_t wo0 = wo(s, 0), wo1 = wo(s, 1);
            if (cnt != n) {  // some rows only: the dense sorted shape needs no unpacking at all (dense_value)
                const DenseRegs dr = dense_issue(s_pfor[s].words, wo0, wo1, n, lane);
                if (dense_prepare(dr, vals0 + s * kBlkVals, lane)) {
                    dense_slots |= 1u << s;
                    continue;
                }
            }
            IMM3_CHECK(ctrl, wo1 >= wo0 + 3u && (int)(wo1 - wo0) <= P.blk_words_cap, 8);  // the block's words fit the decode scratch
            const uint32_t b = pfor_decode_warp(s_pfor[s].words, wo0, wo1, n, Wb, P.blk_words_cap, vals0 + s * kBlkVals, lane);
            bases[s * 32 + lane] = b;
            __syncwarp();
        }
        if (cnt == n) {
            // every row of the block survives (the inside of a window on a sorted column): no selection vector - decoded
            // columns go out row by row (a coalesced store per 32 rows), dense columns are copied straight
#pragma unroll 1
            for (int pc = 0; pc < P.nproj; pc++) {
                const int w = s_proj[pc].width, slot = s_proj[pc].pfor_slot;
                if ((direct >> pc) & 1u) continue;
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
            return;
        }
        append_selection(myword, lane, sel_w, 0u);
        __syncwarp();
        if (fused_ok) {
            // R rows per lane and round; a block with few surviving rows (a scattered 1 % predicate leaves ~10 of 1024) takes the
            // R = 1 instantiation: a quarter of the (predicated-off) gathers and stores of the 4-row form
            auto emit_rows = [&](auto rtag, int b0) {
                constexpr int R = decltype(rtag)::value;
                int idx[R];  // row within the block, -1 = no row
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int i = b0 + lane + 32 * r;
                    idx[r] = i < nn ? (int)lds_cell<uint16_t>(sel_addr + 2u * (uint32_t)i) : -1;
                }
                uint32_t v[R][4];
#pragma unroll
                for (int pc = 0; pc < 4; pc++) {
                    if (pc < P.nproj) {
                        const int w = s_proj[pc].width, slot = s_proj[pc].pfor_slot;
                        if (slot >= 0) {
                            const uint32_t* vs = vals0 + slot * kBlkVals;
                            const uint32_t* bs = bases + slot * 32;
                            if ((dense_slots >> slot) & 1u) {
#pragma unroll
                                for (int r = 0; r < R; r++) v[r][pc] = idx[r] >= 0 ? dense_value(vs, idx[r]) : 0u;
                            } else {
#pragma unroll
                                for (int r = 0; r < R; r++) v[r][pc] = idx[r] >= 0 ? vs[(idx[r] >> 5) * kBlkLane + (idx[r] & 31)] + bs[idx[r] >> 5] : 0u;
                            }
                        } else {
                            const uint8_t* cbase = s_proj[pc].base + R0 * w;
                            if (w == 4) {
#pragma unroll
                                for (int r = 0; r < R; r++) v[r][pc] = idx[r] >= 0 ? __ldg(reinterpret_cast<const uint32_t*>(cbase) + idx[r]) : 0u;
                            } else if (w == 1) {
#pragma unroll
                                for (int r = 0; r < R; r++) v[r][pc] = idx[r] >= 0 ? (uint32_t)__ldg(cbase + idx[r]) : 0u;
                            } else {
#pragma unroll
                                for (int r = 0; r < R; r++)
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
                        for (int r = 0; r < R; r++) {
                            if (idx[r] >= 0) {
                                if (w == 4) reinterpret_cast<uint32_t*>(obase)[32 * r] = v[r][pc];
                                else if (w == 1) obase[32 * r] = (uint8_t)v[r][pc];
                                else reinterpret_cast<uint16_t*>(obase)[32 * r] = (uint16_t)v[r][pc];
                            }
                        }
                    }
                }
            };
#pragma unroll 1
            for (int b0 = 0; b0 < nn; b0 += 128) {
                if (nn - b0 <= 32) emit_rows(std::integral_constant<int, 1>{}, b0);
                else emit_rows(std::integral_constant<int, 4>{}, b0);
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
                        *reinterpret_cast<uint32_t*>(o) = ((dense_slots >> slot) & 1u) ? dense_value(vals0 + slot * kBlkVals, row)
                                                                                      : vals0[slot * kBlkVals + (row >> 5) * kBlkLane + (row & 31)] + bases[slot * 32 + (row >> 5)];
                    } else {
                        const uint8_t* src = s_proj[pc].base + (R0 + row) * w;
                        for (int b = 0; b < w; b++) o[b] = __ldg(src + b);
                    }
                }
            }
        }
    };

    if (ROWSPACE) {
        // block metadata in ONE round trip: lanes 0,1 = row ordinals, lanes 2+2s, 3+2s = word offsets of encoded column s
        auto load_meta = [&](long long b) -> unsigned long long {
            unsigned long long m = 0;
            if (b < nblocks) {
                if (lane < 2) m = P.row_start[b + lane];
                else if (lane < 2 + 2 * P.npfor) m = s_pfor[(lane - 2) >> 1].word_off[b + (lane & 1)];
            }
            return m;
        };
        // One block (its metadata in `meta`): the bits of its rows, its first result ordinal, emit_block.
        auto handle = [&](unsigned long long meta) {
            const long long R0 = (long long)__shfl_sync(0xFFFFFFFFu, meta, 0);
            const int n = (int)((long long)__shfl_sync(0xFFFFFFFFu, meta, 1) - R0);
            const long long bit0 = R0 + 32 * lane;
            const uint32_t lo = __ldg(bitmap + (bit0 >> 5)), hi = __ldg(bitmap + (bit0 >> 5) + 1);
            const long long span = R0 >> 10;
            const uint32_t sw = __ldg(bitmap + span * 32 + lane);                        // R0's span, word `lane`
            const unsigned sc = lane < (int)(span & 7) ? __ldg(cnts + (span & ~7ll) + lane) : 0u;  // earlier spans of the tile
            const unsigned long long toff = __ldg(tile_off + (span >> 3));
            uint32_t myword = __funnelshift_r(lo, hi, (uint32_t)(bit0 & 31));
            const int left = n - lane * 32;
            myword &= left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
            if (__ballot_sync(0xFFFFFFFFu, myword != 0u) == 0u) return;
            const long long wrow0 = (span << 10) + 32 * lane;  // first row of span word `lane`
            const unsigned below = wrow0 + 32 <= R0 ? (unsigned)__popc(sw) : (wrow0 < R0 ? (unsigned)__popc(sw & ((1u << (int)(R0 - wrow0)) - 1u)) : 0u);
            const long long g = (long long)toff + __reduce_add_sync(0xFFFFFFFFu, sc + below);
            emit_block(R0, n, myword, g, [&](int slot, int k) -> uint32_t { return (uint32_t)__shfl_sync(0xFFFFFFFFu, meta, 2 + 2 * slot + k); }, -1, DenseRegs{0, 0, 0, 0, -1});
        };
        // A warp visits blocks warp0, warp0 + nwarps, ... (round robin, so that a clustered result spreads over all warps).
        if (__ldcg(&ctrl->total) < (unsigned long long)nblocks && !(P.debug & 2048u)) {
            // Fewer surviving rows than blocks: most blocks are empty.  32 blocks at a time, a lane each, are tested through the
            // counts of the (at most two) 1024-row spans they overlap - two loads per lane instead of a warp-wide visit per
            // block; the warp then handles the blocks that may hold rows, the next one's metadata in flight.
#pragma unroll 1
            for (long long b0 = warp0; b0 < nblocks; b0 += 32 * nwarps) {
                const long long b = b0 + lane * nwarps;
                bool cand = b < nblocks;
                if (cand) {
                    const unsigned long long r0 = P.row_start[b], r1 = P.row_start[b + 1];
                    const long long s0 = (long long)(r0 >> 10), s1 = (long long)((r1 - 1) >> 10);
                    const unsigned c = __ldg(cnts + s0) + (s1 != s0 ? __ldg(cnts + s1) : 0u);
                    cand = c != 0u;
                }
                unsigned todo = __ballot_sync(0xFFFFFFFFu, cand);
                if (!todo) continue;
                int nsrc = __ffs((int)todo) - 1;
                unsigned long long meta_n = load_meta(__shfl_sync(0xFFFFFFFFu, b, nsrc));
#pragma unroll 1
                while (todo) {
                    todo &= todo - 1u;
                    const unsigned long long meta = meta_n;
                    if (todo) {
                        nsrc = __ffs((int)todo) - 1;
                        meta_n = load_meta(__shfl_sync(0xFFFFFFFFu, b, nsrc));
                    }
                    handle(meta);
                }
            }
        } else {
            // ... every block in turn, the next block's metadata (its bits are found through R0) in flight while this one is handled
            unsigned long long meta_n = load_meta(warp0);
#pragma unroll 1
            for (long long blk = warp0; blk < nblocks; blk += nwarps) {
                const unsigned long long meta = meta_n;
                meta_n = load_meta(blk + nwarps);
                handle(meta);
            }
        }
    } else {
        // Block-local bitmap.  A warp visits blocks warp0, warp0 + nwarps, ... (round robin, so that a clustered result spreads
        // over all warps) in groups of 32, a lane each:
        //   round trip 0: the counts of 8 such groups (eight loads in flight per lane) - or, on large tables where that scan of
        //                 every block count would cost a million sector requests, entries of the non-empty-tile list;
        //   round trip 1: every lane with a non-empty block fetches that block's metadata itself (row ordinals, word offsets,
        //                 its tile's counts and offset) - one round trip for all the non-empty blocks of the group;
        //   then the blocks one after the other, the loads of the next block's encoded words (dense_issue) in flight while
        //   this one is decoded and written.
        int prim = -1;  // the select-list entry whose loads are issued ahead: the first encoded column
        for (int pc = P.nproj - 1; pc >= 0; pc--) prim = s_proj[pc].pfor_slot >= 0 ? pc : prim;
        const int pslot = prim >= 0 ? s_proj[prim].pfor_slot : 0;
        const uint32_t* const pwords = s_pfor[pslot].words;
        // my block b (if cand): metadata in one round trip, then the group's non-empty blocks one after the other
        auto process_group = [&](long long b, bool cand) {
            // ---- round trip 1: my block's metadata ----
            unsigned long long r0 = 0, r1 = 0, g = 0;
            uint32_t w0a = 0, w1a = 0, w0b = 0, w1b = 0, w0c = 0, w1c = 0, w0d = 0, w1d = 0, mycnt = 0;
            if (cand) {
                r0 = P.row_start[b];
                r1 = P.row_start[b + 1];
                if (P.npfor > 0) { w0a = __ldg(s_pfor[0].word_off + b); w1a = __ldg(s_pfor[0].word_off + b + 1); }
                if (P.npfor > 1) { w0b = __ldg(s_pfor[1].word_off + b); w1b = __ldg(s_pfor[1].word_off + b + 1); }
                if (P.npfor > 2) { w0c = __ldg(s_pfor[2].word_off + b); w1c = __ldg(s_pfor[2].word_off + b + 1); }
                if (P.npfor > 3) { w0d = __ldg(s_pfor[3].word_off + b); w1d = __ldg(s_pfor[3].word_off + b + 1); }
                const uint4* tc = reinterpret_cast<const uint4*>(cnts + (b & ~7ll));
                const uint4 ca = __ldg(tc), cb = __ldg(tc + 1);
                g = __ldg(tile_off + (b >> 3));
                const int j = (int)(b & 7);
                const uint32_t c8[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    g += i < j ? c8[i] : 0u;
                    mycnt = i == j ? c8[i] : mycnt;
                }
            }
            const uint32_t pw0 = pslot == 0 ? w0a : (pslot == 1 ? w0b : (pslot == 2 ? w0c : w0d));
            const uint32_t pw1 = pslot == 0 ? w1a : (pslot == 1 ? w1b : (pslot == 2 ? w1c : w1d));
            const int myn = (int)(r1 - r0);
            const bool ahead_ok = prim >= 0 && mycnt == (uint32_t)myn && (long long)g < P.limit;  // (per lane: my block can take the fast path)
            // ... and nothing but that: the query projects this one encoded column and the LIMIT does not cut the block
            const bool direct_ok = ahead_ok && P.nproj == 1 && (long long)g + myn <= P.limit;
            if (lane == 0) phase_stamp(P, 11);
            unsigned todo = __ballot_sync(0xFFFFFFFFu, cand && mycnt != 0u);
            if (!todo) return;
            // ---- the blocks, one after the other ----
            int nsrc = __ffs((int)todo) - 1;
            DenseRegs nx = {0, 0, 0, 0, -1};
            if (__shfl_sync(0xFFFFFFFFu, (int)ahead_ok, nsrc))
                nx = dense_issue(pwords, __shfl_sync(0xFFFFFFFFu, pw0, nsrc), __shfl_sync(0xFFFFFFFFu, pw1, nsrc), __shfl_sync(0xFFFFFFFFu, myn, nsrc), lane);
#pragma unroll 1
            while (todo) {
                const int src = nsrc;
                todo &= todo - 1u;
                const DenseRegs cur = nx;
                nx.B = -1;
                if (todo) {
                    nsrc = __ffs((int)todo) - 1;
                    if (__shfl_sync(0xFFFFFFFFu, (int)ahead_ok, nsrc))
                        nx = dense_issue(pwords, __shfl_sync(0xFFFFFFFFu, pw0, nsrc), __shfl_sync(0xFFFFFFFFu, pw1, nsrc), __shfl_sync(0xFFFFFFFFu, myn, nsrc), lane);
                }
                if (__shfl_sync(0xFFFFFFFFu, (int)direct_ok, src)) {
                    const long long gd = (long long)__shfl_sync(0xFFFFFFFFu, g, src);
                    __syncwarp();
                    if (dense_finish(cur, reinterpret_cast<uint32_t*>(s_proj[0].out) + gd, 1024, lane, Wb, P.debug)) continue;
                }
                const long long blk = __shfl_sync(0xFFFFFFFFu, b, src);
                const long long R0 = (long long)__shfl_sync(0xFFFFFFFFu, r0, src);
                const int n = __shfl_sync(0xFFFFFFFFu, myn, src);
                const long long gb = (long long)__shfl_sync(0xFFFFFFFFu, g, src);
                // the filter kernel stores the 32 words of a block only if SOME of its rows survive; all of them: the count says so
                uint32_t myword;
                if (__shfl_sync(0xFFFFFFFFu, mycnt, src) == (unsigned)n) {
                    const int left = n - lane * 32;
                    myword = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
                } else {
                    myword = __ldg(bitmap + blk * 32 + lane);
                }
                if (lane == 0) phase_stamp(P, 12);
                emit_block(R0, n, myword, gb, [&](int slot, int k) -> uint32_t {
                    const uint32_t x0 = slot == 0 ? w0a : (slot == 1 ? w0b : (slot == 2 ? w0c : w0d));
                    const uint32_t x1 = slot == 0 ? w1a : (slot == 1 ? w1b : (slot == 2 ? w1c : w1d));
                    return __shfl_sync(0xFFFFFFFFu, k ? x1 : x0, src);
                }, cur.B >= 0 ? prim : -1, cur);
                if (lane == 0) phase_stamp(P, 13);
            }
        };
        const unsigned nlist = tile_list ? __ldcg(&ctrl->ticket2) : 0u;
        if (tile_list) {
            // large tables: offset_scan_kernel left the list of non-empty TILES; entry e = (list position e / 8, block e % 8 of
            // that tile), entries warp0, warp0 + nwarps, ... are mine, 32 of them (a lane each) per step
            const long long nent = (long long)nlist * 8;
#pragma unroll 1
            for (long long e0 = warp0; e0 < nent; e0 += 32 * nwarps) {
                const long long e = e0 + lane * nwarps;
                long long b = 0;
                bool cand = e < nent;
                if (cand) {
                    b = (long long)__ldg(tile_list + (e >> 3)) * 8 + (e & 7);
                    cand = b < nblocks;
                }
                if (lane == 0) phase_stamp(P, 10);
                process_group(b, cand);
            }
        } else {
#pragma unroll 1
            for (long long it0 = 0;; it0 += 8) {
                if (warp0 + it0 * 32 * nwarps >= nblocks) break;
                unsigned mymask = 0;  // lane u: the non-empty blocks of group it0 + u
                {
                    unsigned c[8];
                    const uint32_t* cp = cnts + warp0 + (it0 * 32 + lane) * nwarps;
                    const long long left = nblocks - (warp0 + (it0 * 32 + lane) * nwarps);  // blocks from this lane's first one on
                    const long long stride = 32 * nwarps;
#pragma unroll
                    for (int u = 0; u < 8; u++) c[u] = (long long)u * stride < left ? __ldg(cp + u * stride) : 0u;
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const unsigned bal = __ballot_sync(0xFFFFFFFFu, c[u] != 0u);
                        if (lane == u) mymask = bal;
                    }
                }
                if (lane == 0) phase_stamp(P, 10);
#pragma unroll 1
                for (int u = 0; u < 8; u++) {
                    const unsigned m = __shfl_sync(0xFFFFFFFFu, mymask, u);
                    if (!m) continue;
                    process_group(warp0 + ((it0 + u) * 32 + lane) * nwarps, (m >> lane) & 1u);
                }
            }
        }
        if (lane == 0) phase_stamp(P, 14);
    }
}

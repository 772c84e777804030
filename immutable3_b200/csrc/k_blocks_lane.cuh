// k_blocks_lane.cuh - K1b for a sorted column: one LANE per block, whole blocks decided from two numbers
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// blocks_filter_lane_kernel (round 2, third mapping of the same decision procedure; one range predicate on one encoded column).
//
// blocks_filter_quad_kernel decides every MINI-BLOCK of every block (134 warp-instructions per 1024-row block, issue-bound at
// 0.27 of the HBM roofline).  On a sorted column almost every block lies wholly inside or wholly outside the window, and
// that follows from two numbers: the block's first value d0 (first delta of its first mini-block: initvalue = 0 at every
// block) and the sum T of its other 1023 deltas.  All widths < 32 means all deltas are non-negative, so if T cannot wrap the
// values rise from d0 to d0 + T and the block is decided at once by  uf = d0 - lo, ul = uf + T  (modular, as everywhere):
// ul >= uf && ul <= span -> every row passes;  ul >= uf && uf > span -> none does.  Sums commute, so ONE LANE can add up a
// whole block with no cross-lane traffic at all: a warp takes 32 blocks, and a block costs ~20 warp-instructions.
//
//   * Shape the fast path takes (checked per lane, against the bytes - nothing is assumed): rows = n0 (the tile's first block,
//     a multiple of 128), first mini-block of any width B < 32 (it carries the block's absolute start value), every other
//     mini-block of the width k in {0, 1, 2, 4, 8} that the tile's first block uses - i.e. header 0 = (B, k, k, k), headers
//     1.. = (k, k, k, k) and the block's word count equals what that shape implies.  k-bit fields never straddle a byte, so
//     their sums need no byte swap (k = 1: one POPC per mini-block; 2, 4, 8: SWAR + IDP4A); the B-bit mini-block is summed
//     field by field (funnel shift), its fields OR-ed to prove T < 2^31.
//   * A lane whose block has another shape, whose T might wrap, or whose [d0, d0 + T] straddles a window edge hands the
//     block to the exact per-block routine (pfor_range_word, whole warp, lane = mini-block); if more than four lanes do -
//     unsorted or irregular data - the tile goes, quad by quad, through pfor_range_quad like in the quad kernel.  The result
//     is bit-identical to decode-then-compare in every case.
//   * Bank conflicts: lane l reads its own block, blocks are S words apart, so word j of every block sits in bank
//     (S l + j) mod 32 - an 8-way conflict for the 72-word blocks of ids >= 2^29.  Sums commute: lane l starts its walk
//     rho(l) = l / (32 / gcd(S, 32)) super-blocks (fields, for the wide mini-block) further on, which spreads the lanes of
//     one bank class over distinct banks.
//   * Every warp runs its own TMA ring (lane 0 issues the bulk copies of the tile after next, the warp waits on its own
//     mbarriers): no producer warp, no CTA-wide synchronisation, no shared counters; tile counts leave from registers.
// =============================================================================================
constexpr int kLaneWarpsMax = 16;  // warps of a CTA (blockDim.x / 32 says how many it has)
constexpr int kLaneStages = 4;  // most ring slots a warp can have (ScanPlan::stages says how many it has: 2 .. 4)

template <int K>
__device__ __forceinline__ uint32_t narrow_sum(uint32_t a) {
    if (K == 1) return (uint32_t)__popc(lds32(a));
    if (K == 2) {
        const uint32_t w0 = lds32(a), w1 = lds32(a + 4u);
        return (uint32_t)(__popc(w0 & 0x55555555u) + __popc(w1 & 0x55555555u)) + 2u * (uint32_t)(__popc(w0 & 0xAAAAAAAAu) + __popc(w1 & 0xAAAAAAAAu));
    }
    if (K == 4) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t w = lds32(a + 4u * i);
            acc = __dp4a((w & 0x0F0F0F0Fu) + ((w >> 4) & 0x0F0F0F0Fu), 0x01010101u, acc);
        }
        return acc;
    }
    if (K == 8) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) acc = __dp4a(lds32(a + 4u * i), 0x01010101u, acc);
        return acc;
    }
    return 0u;
}

// Sum of the narrow mini-blocks of the lane's block + header check.  wa: shared address of the block's word 0; B: width of
// its first mini-block; rs: rotation of the super-block walk.  Returns the sum; bad != 0 if a header is not (k, k, k, k).
template <int K>
__device__ __forceinline__ uint32_t lane_narrow(uint32_t wa, int nsuper, uint32_t B, int rs, uint32_t& bad) {
    const uint32_t P1 = 2u + B + 3u * K;  // position of header 1
    uint32_t sum = 0;
#pragma unroll 2
    for (int si = 0; si < nsuper; si++) {
        int s = si + rs;
        s = s >= nsuper ? s - nsuper : s;
        const bool first = s == 0;
        const uint32_t ph = first ? 1u : P1 + (uint32_t)(s - 1) * (1u + 4u * K);
        const uint32_t hraw = lds32(wa + 4u * ph);
        bad |= first ? 0u : (hraw ^ (K * 0x01010101u));  // (header 0 was checked by the caller)
        const uint32_t a = wa + 4u * (first ? 2u + B - K : ph + 1u);  // mini-block q of this super-block sits at a + 4 q K
        if (K > 0) {
            const uint32_t m0 = narrow_sum<K>(first ? a + 4u * K : a);  // (first: mini-block 0 is the wide one - read mini-block 1 twice, count it once)
            const uint32_t m1 = narrow_sum<K>(a + 4u * K), m2 = narrow_sum<K>(a + 8u * K), m3 = narrow_sum<K>(a + 12u * K);
            sum += (first ? 0u : m0) + m1 + m2 + m3;
        }
    }
    return sum;
}

// The dense sorted column's block - 1024 rows, widths (B, 1, 1, 1), (1, 1, 1, 1) x 7 - with B known at compile time.
// lane_wide: the B words of the wide mini-block go to registers and every field's word index, shift and mask folds to an
// immediate (3 instructions per field instead of 12).
//   M < 0  : LDS.32 at the lane's own word addresses (block sizes B + 42 that are not a multiple of 4 words put the lanes'
//            blocks in different banks: at most a 2-way conflict)
//   M 0..3 : blocks of B + 42 = 0 mod 4 words all start at the same word M of a 16-byte chunk (and would meet in one bank
//            eight at a time): LDS.128 of whole chunks, block word j = chunk word j + M
// Returns d0; wsum / wor = sum / OR of the other 31 fields.
template <int B, int M>
__device__ __forceinline__ uint32_t lane_wide(uint32_t ww, uint32_t& wsum, uint32_t& wor) {
    uint32_t y[B];  // the mini-block's words (block words 2 .. B + 1), byte-swapped
    if (M < 0) {
#pragma unroll
        for (int j = 0; j < B; j++) y[j] = bswap32(lds32(ww + 4u * (2 + j)));
    } else {
        constexpr int MM = M < 0 ? 0 : M;
        constexpr int C0 = (2 + MM) / 4, C1 = (B + 1 + MM) / 4;  // first / last chunk that holds a word of the mini-block
        const uint32_t cb = ww - 4u * MM;
#pragma unroll
        for (int c = C0; c <= C1; c++) {
            const uint4 v = lds128(cb + 16u * c);
            if (4 * c + 0 - MM - 2 >= 0 && 4 * c + 0 - MM - 2 < B) y[4 * c + 0 - MM - 2 < 0 ? 0 : 4 * c + 0 - MM - 2] = bswap32(v.x);
            if (4 * c + 1 - MM - 2 >= 0 && 4 * c + 1 - MM - 2 < B) y[4 * c + 1 - MM - 2 < 0 ? 0 : 4 * c + 1 - MM - 2] = bswap32(v.y);
            if (4 * c + 2 - MM - 2 >= 0 && 4 * c + 2 - MM - 2 < B) y[4 * c + 2 - MM - 2 < 0 ? 0 : 4 * c + 2 - MM - 2] = bswap32(v.z);
            if (4 * c + 3 - MM - 2 >= 0 && 4 * c + 3 - MM - 2 < B) y[4 * c + 3 - MM - 2 < 0 ? 0 : 4 * c + 3 - MM - 2] = bswap32(v.w);
        }
    }
    constexpr uint32_t mask = (1u << B) - 1u;
    uint32_t sum = 0, orv = 0;
#pragma unroll
    for (int i = 1; i < 32; i++) {
        const int bit = i * B, wi = bit >> 5, sh = bit & 31;
        uint32_t f;
        if (sh + B <= 32) f = y[wi] >> sh;
        else f = __funnelshift_r(y[wi], y[wi + 1 < B ? wi + 1 : wi], sh);
        if (sh + B != 32) f &= mask;
        sum += f;
        orv |= f;
    }
    wsum = sum;
    wor = orv;
    return y[0] & mask;
}

// ... and its 31 one-bit mini-blocks: a POPC each.  The seven regular super-blocks (header + 4 words) are walked in a
// loop that starts rho7 super-blocks further on in lane l (bank conflicts, see above).  bad != 0: a header is not (1, 1, 1, 1).
__device__ __forceinline__ uint32_t lane_narrow_dense(uint32_t ww, uint32_t B, int rho7, uint32_t& bad) {
    const uint32_t n0a = ww + 4u * (2u + B);  // the three narrow mini-blocks of super-block 0
    uint32_t sum = (uint32_t)(__popc(lds32(n0a)) + __popc(lds32(n0a + 4u)) + __popc(lds32(n0a + 8u)));
    const uint32_t p0 = n0a + 12u, pend = p0 + 140u;  // header 1 .. behind super-block 7
    uint32_t p = p0 + 20u * (uint32_t)rho7, hb = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        const uint32_t h = lds32(p), a = lds32(p + 4u), b = lds32(p + 8u), c = lds32(p + 12u), d = lds32(p + 16u);
        hb |= h ^ 0x01010101u;
        sum += (uint32_t)(__popc(a) + __popc(b)) + (uint32_t)(__popc(c) + __popc(d));
        p += 20u;
        p = p == pend ? p0 : p;
    }
    bad = hb;
    return sum;
}

#define IMM3_LANE_WIDE_CASE(BB, MMODE) case BB: d0 = lane_wide<BB, MMODE>(ww, wsum, wor); break;
#define IMM3_LANE_WIDE_ALL(MMODE)                                                                                                   \
    IMM3_LANE_WIDE_CASE(17, MMODE) IMM3_LANE_WIDE_CASE(18, MMODE) IMM3_LANE_WIDE_CASE(19, MMODE) IMM3_LANE_WIDE_CASE(20, MMODE) \
    IMM3_LANE_WIDE_CASE(21, MMODE) IMM3_LANE_WIDE_CASE(22, MMODE) IMM3_LANE_WIDE_CASE(23, MMODE) IMM3_LANE_WIDE_CASE(24, MMODE) \
    IMM3_LANE_WIDE_CASE(25, MMODE) IMM3_LANE_WIDE_CASE(26, MMODE) IMM3_LANE_WIDE_CASE(27, MMODE) IMM3_LANE_WIDE_CASE(28, MMODE) \
    IMM3_LANE_WIDE_CASE(29, MMODE) IMM3_LANE_WIDE_CASE(30, MMODE) IMM3_LANE_WIDE_CASE(31, MMODE)
#define IMM3_LANE_WIDE_CHUNKS(MMODE) \
    IMM3_LANE_WIDE_CASE(18, MMODE) IMM3_LANE_WIDE_CASE(22, MMODE) IMM3_LANE_WIDE_CASE(26, MMODE) IMM3_LANE_WIDE_CASE(30, MMODE)
constexpr int kLaneFixedMinB = 17;

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(kLaneWarpsMax * 32, 1) blocks_filter_lane_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmapB,
                                                                                uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ tile_cnt,
                                                                                unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl,
                                                                                long long nblocks, const unsigned int* __restrict__ work, uint32_t* __restrict__ grp_sum) {
    __shared__ FilterShared S;
    __shared__ unsigned long long s_bar[kLaneWarpsMax][kLaneStages];
    __shared__ __align__(16) int s_meta[kLaneWarpsMax][8][4];  // per warp, a ring of work items: {32-block tile (-1: none), its first word offset, its end, -}
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) phase_stamp(P, 0);
    if (tid < kLaneWarpsMax * kLaneStages) mbar_init(smem_u32(&s_bar[0][0]) + 8u * (uint32_t)tid, 1);
    if (tid == 0) fence_mbar_init();
    __syncthreads();
    const long long ntiles8 = P.ntiles;           // 8-block tiles (offset scan, emit kernel)
    const long long nct = (nblocks + 31) >> 5;    // 32-block tiles
    const int slot_bytes = P.stage_bytes;
    const FilterCol f0 = P.filter[0];
    const PforCol pc = P.pfor[f0.pfor_slot < 0 ? 0 : (f0.pfor_slot == 0 ? 0 : (f0.pfor_slot == 1 ? 1 : (f0.pfor_slot == 2 ? 2 : 3)))];
    const uint32_t ring_addr = smem_u32(dyn_smem);
    const int nstages = P.stages;
    const uint32_t my_ring = ring_addr + (uint32_t)(warp * nstages) * (uint32_t)slot_bytes;
    const uint32_t lo = (uint32_t)f0.lo, span = f0.span;
    const int nwork = work ? (int)__ldg(work) : (int)nct;  // pruned query: only the tiles blocks_prune_kernel listed
    const int cta_warps = (int)(blockDim.x >> 5);
    const int gw = (int)blockIdx.x * cta_warps + warp, nwarps = (int)gridDim.x * cta_warps;
    const uint32_t meta_addr = smem_u32(&s_meta[warp][0][0]);

    // Work items q = 0, 1, ... of this warp are the work-list entries (or tiles) gw + q nwarps.  Lane 0 keeps three things going
    // for the items ahead, one step each per tile decided, none of which it ever waits for: S1 fetch the item's tile id
    // (pruned queries: from the work list), S2 fetch the tile's first / end word offset, S3 issue its bulk copies.  The
    // fetches are 4-byte cp.async into s_meta - a plain load would stall the warp at the first touch of its register.
    const uint64_t pol_stream = l2_policy_evict_first();  // the column is streamed exactly once (IMM3_DEBUG bit 256: default policy)
    auto S1 = [&](int q) {
        const long long k = (long long)gw + (long long)q * nwarps;
        const uint32_t e = meta_addr + 16u * (uint32_t)(q & 7);
        if (k < nwork) {
            if (work) cp_async4(e, work + 1 + k);
            else s_meta[warp][q & 7][0] = (int)k;
        } else {
            s_meta[warp][q & 7][0] = -1;
        }
    };
    auto S2 = [&](int q) {
        const int t = *reinterpret_cast<volatile int*>(&s_meta[warp][q & 7][0]);
        if (t >= 0) {
            const long long b0 = (long long)t * 32, b1 = b0 + 32 < nblocks ? b0 + 32 : nblocks;
            const uint32_t e = meta_addr + 16u * (uint32_t)(q & 7);
            cp_async4(e + 4u, pc.word_off + b0);
            cp_async4(e + 8u, pc.word_off + b1);
        }
    };
    auto S3 = [&](int q, int slot) {
        const int t = *reinterpret_cast<volatile int*>(&s_meta[warp][q & 7][0]);
        if (t >= 0) {
            const uint32_t wo0 = *reinterpret_cast<volatile uint32_t*>(&s_meta[warp][q & 7][1]);
            const uint32_t wo1 = *reinterpret_cast<volatile uint32_t*>(&s_meta[warp][q & 7][2]);
            const uint32_t bar = smem_u32(&s_bar[warp][slot]);
            const uint32_t dst = my_ring + (uint32_t)slot * (uint32_t)slot_bytes;
            const long long b0 = (long long)t * 32;
            const uint32_t base_w = wo0 & ~3u;  // 16-byte aligned source
            uint32_t nb = ((wo1 - base_w) * 4u + 15u) & ~15u;
            if (nb > (uint32_t)P.blk_tile_bytes) nb = (uint32_t)P.blk_tile_bytes;
            mbar_arrive_expect_tx(bar, 272u + 144u + nb);
            tma_load_1d(dst, P.row_start + b0, 272u, bar);
            tma_load_1d(dst + (uint32_t)kQuadWoOff, pc.word_off + b0, 144u, bar);
            if (!(P.debug & 256u)) tma_load_1d_hint(dst + (uint32_t)kQuadHdrBytes, pc.words + base_w, nb, bar, pol_stream);
            else tma_load_1d(dst + (uint32_t)kQuadHdrBytes, pc.words + base_w, nb, bar);
        }
    };
    if (lane == 0) {
        for (int q = 0; q <= nstages; q++) S1(q);
        cp_async_commit_wait_all();
        for (int q = 0; q < nstages; q++) S2(q);
        cp_async_commit_wait_all();
        for (int q = 0; q < nstages - 1; q++) S3(q, q);
    }
    __syncwarp();

    int slot = 0, pslot = nstages - 1;
    unsigned use = 0;
#pragma unroll 1
    for (int it = 0;; it++) {
        if (lane == 0) {
            cp_async_wait_all();           // (issued one step ago)
            S3(it + nstages - 1, pslot);   // into the slot decided in the previous step: every lane is past its last read of it
            S2(it + nstages);
            S1(it + nstages + 1);
            cp_async_commit();
        }
        __syncwarp();
        const int T = *reinterpret_cast<volatile int*>(&s_meta[warp][it & 7][0]);
        if (T < 0) break;
        mbar_wait(smem_u32(&s_bar[warp][slot]), use & 1u, nullptr);
        const uint32_t sl = my_ring + (uint32_t)slot * (uint32_t)slot_bytes;
        const uint32_t base_w = *reinterpret_cast<volatile uint32_t*>(&s_meta[warp][it & 7][1]) & ~3u;
        const long long blk0 = (long long)T * 32, blk = blk0 + lane;
        if (P.debug & 32u) {  // timing experiment (wrong results): the data path alone
            __syncwarp();
            pslot = slot;
            if (++slot == nstages) { slot = 0; use++; }
            continue;
        }
        // ---------------- my block ----------------
        const bool exists = blk < nblocks;
        int n = 0, nw = 0;
        uint32_t wa = sl + (uint32_t)kQuadHdrBytes, stride = 0, gw0 = 0;
        if (exists) {
            const unsigned long long r0 = lds_cell<unsigned long long>(sl + 8u * (uint32_t)lane), r1 = lds_cell<unsigned long long>(sl + 8u * (uint32_t)lane + 8u);
            const uint32_t w0 = lds32(sl + (uint32_t)kQuadWoOff + 4u * (uint32_t)lane), w1 = lds32(sl + (uint32_t)kQuadWoOff + 4u * (uint32_t)lane + 4u);
            n = (int)(r1 - r0);
            stride = w1 - w0;
            gw0 = w0;
            nw = (int)stride - 2;
            wa += 4u * (w0 - base_w);
            IMM3_CHECK(ctrl, n >= 0 && nw >= 1 && wa + 4u * stride <= sl + (uint32_t)slot_bytes + 16u, 1);  // the block lies inside the ring slot
        }
        IMM3_CHECK(ctrl, (long long)T < nct, 2);
        // the tile's shape is taken from its first block of whole super-blocks (NOT simply from lane 0: the 1-row tail block that
        // ends every segment sits there in one tile of 32, and sending that whole tile through the quad routine cost its warp
        // 25 us - the slow CTAs of every run)
        const unsigned whole = __ballot_sync(0xFFFFFFFFu, exists && n > 0 && (n & 127) == 0 && nw >= 2);
        const int lane0 = whole ? __ffs((int)whole) - 1 : 0;
        const int n0 = whole ? __shfl_sync(0xFFFFFFFFu, n, lane0) : 0;
        const uint32_t h0raw = (exists && nw >= 2) ? lds32(wa + 4u) : 0xFFFFFFFFu;  // raw big-endian header 0: byte 0 = B, bytes 1..3 = the other widths
        const uint32_t k = __shfl_sync(0xFFFFFFFFu, h0raw, lane0) >> 24;
        const int nsuper = n0 >> 7;
        unsigned cnt = 0;           // rows selected in my block
        unsigned c8_fallback = 0;   // (quad fallback: per 8-block tile counts are written there)
        bool counted = false;
        const bool tile_ok = n0 > 0 && (n0 & 127) == 0 && n0 <= 1024 && (k <= 2u || k == 4u || k == 8u);
        unsigned hard = 0xFFFFFFFFu, good = 0;
        uint32_t B = 0;
        if (tile_ok) {
            B = h0raw & 0xFFu;
            const bool shape = exists && n == n0 && B < 32u && (h0raw >> 8) == k * 0x010101u &&
                               nw == 1 + nsuper + (int)B + (4 * nsuper - 1) * (int)k;
            good = __ballot_sync(0xFFFFFFFFu, shape);
            hard = __ballot_sync(0xFFFFFFFFu, exists && !shape);
        }
        if (tile_ok && good != 0u && __popc(hard) <= 4) {
            const bool mine = (good >> lane) & 1u;
            const int src = __ffs((int)good) - 1;
            const uint32_t wa_src = __shfl_sync(0xFFFFFFFFu, wa, src);
            const uint32_t ww = mine ? wa : wa_src;  // (a lane without a block of its own walks a good one with B = 0: every load stays in bounds)
            if (!mine) B = 0;
            uint32_t d0 = 0, rest = 0, wor = 0, bad = 0;
            // the dense sorted column's shape with one B for the whole tile: everything from registers (lane_fixed)
            const uint32_t B0 = __shfl_sync(0xFFFFFFFFu, B, src);
            const bool fixed = n0 == 1024 && k == 1u && B0 >= (uint32_t)kLaneFixedMinB && B0 < 32u && __all_sync(0xFFFFFFFFu, !mine || B == B0);
            if (fixed) {
                const uint32_t S0 = B0 + 42u;  // words per block
                const int rho = lane >> (6 - __ffs((int)(S0 | 32u)));  // lane / (32 / gcd(S0, 32)): 0 .. gcd - 1
                const int rho7 = rho - 7 * ((rho * 37) >> 8);          // rho % 7
                const uint32_t Mu = (wa_src >> 2) & 3u;
                const bool chunks = (S0 & 3u) == 0u && __all_sync(0xFFFFFFFFu, ((ww >> 2) & 3u) == Mu);
                uint32_t wsum = 0;
                if (chunks) {
                    switch (Mu) {
                        case 0: switch (B0) { IMM3_LANE_WIDE_CHUNKS(0) default: break; } break;
                        case 1: switch (B0) { IMM3_LANE_WIDE_CHUNKS(1) default: break; } break;
                        case 2: switch (B0) { IMM3_LANE_WIDE_CHUNKS(2) default: break; } break;
                        default: switch (B0) { IMM3_LANE_WIDE_CHUNKS(3) default: break; } break;
                    }
                } else {
                    switch (B0) {
                        IMM3_LANE_WIDE_ALL(-1)
                        default: break;
                    }
                }
                rest = wsum + lane_narrow_dense(ww, B0, rho7, bad);
            } else {
                // rotation: lanes whose blocks start in the same bank start their walk at different places
                const uint32_t S0 = __shfl_sync(0xFFFFFFFFu, stride, src);
                const int rho = lane >> (6 - __ffs((int)(S0 | 32u)));  // lane / (32 / gcd(S0, 32)): 0 .. gcd - 1
                const int rs = (nsuper & (nsuper - 1)) == 0 ? (rho & (nsuper - 1)) : rho % nsuper;
                // ---------------- wide mini-block: d0 + the sum of its other 31 fields ----------------
                const uint32_t wb = ww + 8u, mask = (1u << B) - 1u;
                uint32_t d0w, wsum = 0, wor_w = 0;
                {
                    const uint32_t x0 = bswap32(lds32(wb));
                    d0w = x0 & mask;
                    int i = rho >= 31 ? 1 : 1 + rho;
#pragma unroll 8
                    for (int fi = 0; fi < 31; fi++) {
                        const uint32_t off = (uint32_t)i * B;
                        const uint32_t a = wb + 4u * (off >> 5);
                        const uint32_t f = __funnelshift_r(bswap32(lds32(a)), bswap32(lds32(a + 4u)), off) & mask;
                        wsum += f;
                        wor_w |= f;
                        i = i == 31 ? 1 : i + 1;
                    }
                }
                // ---------------- narrow mini-blocks ----------------
                uint32_t nsum;
                switch (k) {
                    case 0: nsum = lane_narrow<0>(ww, nsuper, B, rs, bad); break;
                    case 1: nsum = lane_narrow<1>(ww, nsuper, B, rs, bad); break;
                    case 2: nsum = lane_narrow<2>(ww, nsuper, B, rs, bad); break;
                    case 4: nsum = lane_narrow<4>(ww, nsuper, B, rs, bad); break;
                    default: nsum = lane_narrow<8>(ww, nsuper, B, rs, bad); break;
                }
                d0 = d0w;
                rest = wsum + nsum;
                wor = wor_w;
            }
            // ---------------- decide ----------------
            // (rest < 31 * 2^21 + 1023 * 255 if wor < 2^21)
            const uint32_t uf = d0 - lo, ul = uf + rest;           // first / last value of the block in the window's frame
            const bool mono = ul >= uf && wor < (1u << 21) && bad == 0u;
            const bool all = mono && ul <= span, none = mono && uf > span;
            cnt = (mine && all) ? (unsigned)n : 0u;
            hard |= __ballot_sync(0xFFFFFFFFu, mine && !all && !none);
            // the few blocks a window edge cuts through (or that are not what they seemed): exact, the whole warp per block
            if (hard) {
                for (unsigned todo = hard; todo; todo &= todo - 1u) {
                    const int m = __ffs((int)todo) - 1;
                    const int nm = __shfl_sync(0xFFFFFFFFu, n, m), nwm = __shfl_sync(0xFFFFFFFFu, nw, m);
                    const uint32_t wam = __shfl_sync(0xFFFFFFFFu, wa, m);
                    const uint32_t* W = reinterpret_cast<const uint32_t*>(dyn_smem + (wam - ring_addr));
                    const int left = nm - lane * 32;
                    uint32_t mw = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
                    mw &= pfor_range_word(W, nwm, nm, lo, span, lane);
                    const unsigned c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(mw));
                    if (c != 0u && c != (unsigned)nm) bitmapB[(blk0 + m) * 32 + lane] = mw;
                    if (lane == m) cnt = c;
                }
            }
            if (exists) blk_cnt[blk] = cnt;
            counted = true;
        } else {
            // irregular tile: the quad routine (which writes counts and bitmap words itself), eight quads
#pragma unroll 1
            for (int qd = 0; qd < 8; qd++) {
                if (blk0 + 4 * qd >= nblocks) break;
                const unsigned qc = quad_decide(sl, ring_addr, base_w, 4 * qd, blk0 + 4 * qd, nblocks, lo, span, lane, bitmapB, blk_cnt);
                if ((lane >> 3) == (qd >> 1)) c8_fallback += qc;
            }
        }
        // ---------------- counts of the four 8-block tiles ----------------
        unsigned c8 = cnt;
        c8 += __shfl_xor_sync(0xFFFFFFFFu, c8, 1);
        c8 += __shfl_xor_sync(0xFFFFFFFFu, c8, 2);
        c8 += __shfl_xor_sync(0xFFFFFFFFu, c8, 4);
        if (!counted) c8 = c8_fallback;
        const long long t8 = (blk0 >> 3) + (lane >> 3);
        if ((lane & 7) == 0 && t8 < ntiles8) tile_cnt[t8] = c8;
        if (grp_sum) {  // blocks_group_emit_kernel follows: this tile's rows join the sum of its group of 1024 blocks (fire and forget)
            unsigned c32 = c8 + __shfl_xor_sync(0xFFFFFFFFu, c8, 8);
            c32 += __shfl_xor_sync(0xFFFFFFFFu, c32, 16);
            if (lane == 0 && c32 != 0u) atomicAdd(grp_sum + (blk0 >> 10), c32);
            // ... and what the emit kernel will read of it - the selected blocks' words, the tile's row ordinals and word offsets - is
            // touched again so that it outlives the rest of this kernel's stream in L2 (the stream itself is marked evict-first)
            if (!(P.debug & 512u) && c32 != 0u && counted) {  // (IMM3_DEBUG bit 512 switches it off)
                if (cnt != 0u) {
                    const char* a = reinterpret_cast<const char*>(pc.words + gw0);
                    for (uint32_t o = 0; o < 4u * stride; o += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(a + o));
                }
                if (lane < 3) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(P.row_start + blk0) + 128 * lane));
                else if (lane < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(pc.word_off + blk0) + 128 * (lane - 3)));
            }
        }
        __syncwarp();
        pslot = slot;
        if (++slot == nstages) { slot = 0; use++; }
    }

    if (lane == 0) phase_stamp(P, 2);
    __syncthreads();
    if (tid == 0) {
        phase_stamp(P, 1);
        if ((P.debug & 16u) && P.trace && blockIdx.x < 512) {  // debugging: when this CTA was done, and on which SM
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            P.trace[64 + 2 * blockIdx.x] = globaltimer_ns();
            P.trace[64 + 2 * blockIdx.x + 1] = (unsigned long long)smid << 32;
        }
        if (grp_sum) {
            S.is_last = 0;  // blocks_group_emit_kernel follows: it needs nothing from a last CTA, and nobody waits here for a fence + an atomic
        } else {
            __threadfence();
            const unsigned prev = atomicAdd(&ctrl->exited, 1u);
            S.is_last = prev == gridDim.x - 1;
            if (S.is_last) {
                ctrl->exited = 0;
                ctrl->ticket2 = 0;  // (offset_scan_kernel counts the non-empty tiles of this query here)
            }
        }
    }
    __syncthreads();
    if (S.is_last && warp < kComputeWarps && P.scan_inline) {
        __threadfence();
        scan_tile_counts(S, tile_cnt, tile_off, ntiles8, P.limit, ctrl);
    }
}

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

__global__ void __launch_bounds__(kComputeThreads, 1) blocks_filter_lane_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmapB,
                                                                                uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ tile_cnt,
                                                                                unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl,
                                                                                long long nblocks, const unsigned int* __restrict__ work) {
    __shared__ FilterShared S;
    __shared__ unsigned long long s_bar[kComputeWarps][kLaneStages];
    __shared__ uint32_t s_basew[kComputeWarps][kLaneStages];  // arena word that sits at the slot's data offset
    __shared__ int s_tile[kComputeWarps][kLaneStages];        // 32-block tile held by the slot (-1: no more work)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < kComputeWarps * kLaneStages) mbar_init(smem_u32(&s_bar[0][0]) + 8u * (uint32_t)tid, 1);
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
    const long long nwork = work ? (long long)__ldg(work) : nct;  // pruned query: only the tiles blocks_prune_kernel listed
    const long long gw = (long long)blockIdx.x * kComputeWarps + warp, nwarps = (long long)gridDim.x * kComputeWarps;

    // Work items gw, gw + nwarps, ...; the copies of item i + stages - 1 are issued while item i is decided.  The tile id of an item is
    // fetched two steps, its first / last word offset one step before its copies are issued (no global load is waited for).
    auto tile_of = [&](long long k) -> long long { return k < nwork ? (work ? (long long)__ldg(work + 1 + k) : k) : -1ll; };
    long long kq = gw;           // next item whose tile id gets fetched
    long long tA = tile_of(kq);  // tile id fetched, word offsets not yet
    kq += nwarps;
    long long tB = -1;           // tile id + word offsets fetched: ready to issue
    uint32_t woB0 = 0, woB1 = 0;
    auto advance = [&]() {
        tB = tA;
        if (tB >= 0) {
            const long long b0 = tB * 32, b1 = b0 + 32 < nblocks ? b0 + 32 : nblocks;
            woB0 = __ldg(pc.word_off + b0);
            woB1 = __ldg(pc.word_off + b1);
        }
        tA = tile_of(kq);
        kq += nwarps;
    };
    auto issue = [&](int slot) {  // copies of item (tB, woB0, woB1) into `slot`
        if (lane == 0) {
            s_tile[warp][slot] = (int)tB;
            if (tB >= 0) {
                const uint32_t bar = smem_u32(&s_bar[warp][slot]);
                const uint32_t dst = my_ring + (uint32_t)slot * (uint32_t)slot_bytes;
                const long long b0 = tB * 32;
                const uint32_t base_w = woB0 & ~3u;  // 16-byte aligned source
                uint32_t nb = ((woB1 - base_w) * 4u + 15u) & ~15u;
                if (nb > (uint32_t)P.blk_tile_bytes) nb = (uint32_t)P.blk_tile_bytes;
                s_basew[warp][slot] = base_w;
                mbar_arrive_expect_tx(bar, 272u + 144u + nb);
                tma_load_1d(dst, P.row_start + b0, 272u, bar);
                tma_load_1d(dst + (uint32_t)kQuadWoOff, pc.word_off + b0, 144u, bar);
                tma_load_1d(dst + (uint32_t)kQuadHdrBytes, pc.words + base_w, nb, bar);
            }
        }
    };
    advance();
#pragma unroll 1
    for (int j = 0; j < nstages - 1; j++) {
        issue(j);
        advance();
    }
    __syncwarp();

    int slot = 0, pslot = nstages - 1;
    unsigned use = 0;
#pragma unroll 1
    for (;;) {
        issue(pslot);  // (the slot decided in the previous step: every lane is past its last read of it)
        advance();
        __syncwarp();
        const int T = *reinterpret_cast<volatile int*>(&s_tile[warp][slot]);
        if (T < 0) break;
        mbar_wait(smem_u32(&s_bar[warp][slot]), use & 1u, nullptr);
        const uint32_t sl = my_ring + (uint32_t)slot * (uint32_t)slot_bytes;
        const uint32_t base_w = *reinterpret_cast<volatile uint32_t*>(&s_basew[warp][slot]);
        const long long blk0 = (long long)T * 32, blk = blk0 + lane;
        // ---------------- my block ----------------
        const bool exists = blk < nblocks;
        int n = 0, nw = 0;
        uint32_t wa = sl + (uint32_t)kQuadHdrBytes, stride = 0;
        if (exists) {
            const unsigned long long r0 = lds_cell<unsigned long long>(sl + 8u * (uint32_t)lane), r1 = lds_cell<unsigned long long>(sl + 8u * (uint32_t)lane + 8u);
            const uint32_t w0 = lds32(sl + (uint32_t)kQuadWoOff + 4u * (uint32_t)lane), w1 = lds32(sl + (uint32_t)kQuadWoOff + 4u * (uint32_t)lane + 4u);
            n = (int)(r1 - r0);
            stride = w1 - w0;
            nw = (int)stride - 2;
            wa += 4u * (w0 - base_w);
        }
        const int n0 = __shfl_sync(0xFFFFFFFFu, n, 0);
        const uint32_t h0raw = (exists && nw >= 2) ? lds32(wa + 4u) : 0xFFFFFFFFu;  // raw big-endian header 0: byte 0 = B, bytes 1..3 = the other widths
        const uint32_t k = __shfl_sync(0xFFFFFFFFu, h0raw, 0) >> 24;
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
            // rotation: lanes whose blocks start in the same bank start their walk at different places
            const uint32_t S0 = __shfl_sync(0xFFFFFFFFu, stride, src);
            const int rho = lane >> (6 - __ffs((int)(S0 | 32u)));  // lane / (32 / gcd(S0, 32)): 0 .. gcd - 1
            const int rs = (nsuper & (nsuper - 1)) == 0 ? (rho & (nsuper - 1)) : rho % nsuper;
            // ---------------- wide mini-block: d0 + the sum of its other 31 fields ----------------
            const uint32_t wb = ww + 8u, mask = (1u << B) - 1u;
            uint32_t d0, wsum = 0, wor = 0;
            {
                const uint32_t x0 = bswap32(lds32(wb));
                d0 = x0 & mask;
                int i = rho >= 31 ? 1 : 1 + rho;
#pragma unroll 8
                for (int fi = 0; fi < 31; fi++) {
                    const uint32_t off = (uint32_t)i * B;
                    const uint32_t a = wb + 4u * (off >> 5);
                    const uint32_t f = __funnelshift_r(bswap32(lds32(a)), bswap32(lds32(a + 4u)), off) & mask;
                    wsum += f;
                    wor |= f;
                    i = i == 31 ? 1 : i + 1;
                }
            }
            // ---------------- narrow mini-blocks ----------------
            uint32_t bad = 0, nsum;
            switch (k) {
                case 0: nsum = lane_narrow<0>(ww, nsuper, B, rs, bad); break;
                case 1: nsum = lane_narrow<1>(ww, nsuper, B, rs, bad); break;
                case 2: nsum = lane_narrow<2>(ww, nsuper, B, rs, bad); break;
                case 4: nsum = lane_narrow<4>(ww, nsuper, B, rs, bad); break;
                default: nsum = lane_narrow<8>(ww, nsuper, B, rs, bad); break;
            }
            // ---------------- decide ----------------
            const uint32_t rest = wsum + nsum;                     // < 31 * 2^21 + 1023 * 255 if wor < 2^21
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
        __syncwarp();
        pslot = slot;
        if (++slot == nstages) { slot = 0; use++; }
    }

    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&ctrl->exited, 1u);
        S.is_last = prev == gridDim.x - 1;
        if (S.is_last) ctrl->exited = 0;
    }
    __syncthreads();
    if (S.is_last && P.scan_inline) {
        __threadfence();
        scan_tile_counts(S, tile_cnt, tile_off, ntiles8, P.limit, ctrl);
    }
}

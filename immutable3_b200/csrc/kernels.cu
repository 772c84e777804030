// kernels.cu — sm_100a kernels of the scan -> filter -> project path.
//
// Two persistent, single-pass kernels.  Both fuse, per tile:
//   decode (ScanOp / DenseCodec*.decode / sorted-int codec; Scan.scala:28-70, DenseCodec.scala:34-74,
//           PFORCodec.scala:12-28)
//   -> conjunctive RangeFilter / MatchFilter into a selection mask (Select.scala:14-165)
//   -> warp popc/ballot stream compaction into a tile-local selection vector
//   -> device-wide exclusive prefix over tiles (decoupled look-back, so canonical row order and an
//      exact LIMIT cut need no second pass; Project.scala:37-80)
//   -> Project gather of the select-list columns into column-major result buffers.
//
//  * scan_dense_kernel  — tables whose touched columns are all DENSE_*.  Columns live in HBM as flat
//    arrays in canonical row order, a tile is kTileRows consecutive rows, filter columns are staged
//    tile-by-tile into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier) through a
//    multi-stage ring, each lane owns 32 consecutive rows (= one 32-bit word of the selection
//    bitmap) and evaluates its predicate with SIMD-within-a-register compares.
//  * scan_blocks_kernel — anything touching a PFOR_INT column (or forced for cross-checking): a tile
//    is one reference block; the sorted-int codec is unpacked in shared memory (bit-unpack + warp
//    scans of the deltas), dense columns of the same rows are read row-per-lane.
//
// Tiles are handed out by an atomic ticket so a tile only ever waits on tiles that are already
// running (forward progress without relying on block scheduling order).  No spin is unbounded: a
// watchdog traps instead of hanging the GPU.
#include <cuda_runtime.h>
#include <type_traits>

#include <cstdint>

#include "kernels.hpp"
#include "plan.hpp"

namespace imm3 {

// =============================================================================================
// PTX helpers
// =============================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// Debugging (IMM3_DEBUG bit 4 + IMM3_TRACE): phase stamps of the multi-pass kernels, min and max over CTAs per event.
__device__ __forceinline__ void phase_stamp(const ScanPlan& P, int ev) {
    if ((P.debug & 16u) && P.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        atomicMin(P.trace + 2 * ev, t);
        atomicMax(P.trace + 2 * ev + 1, t);
    }
}

// L2 cache policies for bulk copies: a column that a later kernel reads again is kept (evict_last), a column that
// is streamed exactly once goes first (evict_first) so that it does not push the former out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 ldg128(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

constexpr unsigned long long kWatchdogNs = 4000000000ull;  // 4 s: far beyond any legitimate wait

__device__ __noinline__ void watchdog_trap(ScanCtrl* ctrl, unsigned code) {
    if (ctrl) atomicExch(&ctrl->error, code);
    __threadfence_system();
    __trap();
}

// try_wait with a suspend-time hint: the hardware parks the warp instead of having it spin through issue slots
// that the working warps of the SM need.
__device__ __forceinline__ bool mbar_try_wait_park(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, ScanCtrl* ctrl) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    unsigned spins = 0;
    while (!mbar_try_wait_park(bar, parity)) {
        if ((++spins & 63u) == 0) {  // the watchdog clock is read once per 64 parked waits
            const uint64_t now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kWatchdogNs) watchdog_trap(ctrl, 1);
        }
    }
}

// =============================================================================================
// Tile status words for the decoupled look-back: [63:24] value, [23:2] epoch, [1:0] state
// =============================================================================================
constexpr unsigned kStateNone = 0, kStateAggregate = 1, kStatePrefix = 2;


__device__ __forceinline__ unsigned long long pack_status(uint32_t epoch, unsigned state, unsigned long long value) {
    return (value << 24) | ((unsigned long long)(epoch & 0x3FFFFFu) << 2) | state;
}

// Exclusive prefix of tile `tile` (sum of the selected-row counts of all earlier tiles), computed
// by one full warp polling a window of 128 predecessor status words at a time (4 per lane).
// Returns -1 if the LIMIT was reached while waiting (the tile is then dead: the tile that set
// `done` had already seen every earlier tile published, so a tile still waiting on an unpublished
// predecessor lies beyond the cut).
__device__ long long lookback_exclusive(const unsigned long long* status, long long tile, uint32_t epoch, ScanCtrl* ctrl,
                                        int lane) {
    long long running = 0;
    long long pos = tile - 1;
    uint64_t t0 = 0;
    unsigned spins = 0;
    const unsigned ep = epoch & 0x3FFFFFu;
    for (;;) {
        unsigned long long st[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long idx = pos - 4 * lane - k;  // lane 0 / k 0 is the nearest predecessor
            st[k] = idx >= 0 ? ld_relaxed_u64(status + idx) : pack_status(epoch, kStatePrefix, 0);  // virtual tile -1: prefix 0
        }
        unsigned long long lsum = 0;
        bool lpre = false, linv = false;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned state = (unsigned)(st[k] & 3u);
            if (((st[k] >> 2) & 0x3FFFFFu) != ep) state = kStateNone;
            if (!lpre && !linv) {
                if (state == kStateNone) linv = true;
                else {
                    lsum += st[k] >> 24;
                    lpre = state == kStatePrefix;
                }
            }
        }
        const unsigned inv = __ballot_sync(0xFFFFFFFFu, linv);
        const unsigned pre = __ballot_sync(0xFFFFFFFFu, lpre);
        const int p = pre ? (__ffs(pre) - 1) : 32;
        const unsigned need = (p >= 31) ? 0xFFFFFFFFu : ((2u << p) - 1u);  // lanes 0..p
        if (inv & need) {
            if (__any_sync(0xFFFFFFFFu, ld_relaxed_u32(&ctrl->done) != 0u)) return -1;
            if (spins == 0) t0 = globaltimer_ns();
            if ((++spins & 255u) == 0 && globaltimer_ns() - t0 > kWatchdogNs) watchdog_trap(ctrl, 2);
            __nanosleep(20);
            continue;
        }
        unsigned long long c = ((need >> lane) & 1u) ? lsum : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
        running += (long long)c;
        if (p < 32) return running;
        pos -= 128;
    }
}

// Publish this tile's count, resolve its exclusive prefix, publish the inclusive prefix, and do
// the LIMIT / total bookkeeping.  Called by warp 0 only; returns the exclusive prefix (-1 = dead).
__device__ long long resolve_tile(const ScanPlan& P, ScanCtrl* ctrl, unsigned long long* status, long long tile,
                                  unsigned tile_count, int lane) {
    long long excl = 0;
    if (tile == 0) {
        if (lane == 0) st_relaxed_u64(status + tile, pack_status(P.epoch, kStatePrefix, tile_count));
    } else {
        if (lane == 0) st_relaxed_u64(status + tile, pack_status(P.epoch, kStateAggregate, tile_count));
        excl = (P.debug & 1u) ? (long long)tile * 1800 : lookback_exclusive(status, tile, P.epoch, ctrl, lane);
        if (excl < 0) return -1;
        if (lane == 0) st_relaxed_u64(status + tile, pack_status(P.epoch, kStatePrefix, (unsigned long long)excl + tile_count));
    }
    if (lane == 0) {
        const long long incl = excl + (long long)tile_count;
        if (excl < P.limit && incl >= P.limit) {  // this tile crosses the LIMIT (Project.scala:73-77)
            ctrl->total = (unsigned long long)P.limit;
            __threadfence();
            atomicExch(&ctrl->done, 1u);
        } else if (tile == P.ntiles - 1 && incl < P.limit) {
            ctrl->total = (unsigned long long)incl;
        }
    }
    return excl;
}

// Last CTA out resets the control block for the next launch on this stream.
__device__ __forceinline__ void cta_exit(ScanCtrl* ctrl) {
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&ctrl->exited, 1u);
        if (prev == gridDim.x - 1) {
            ctrl->ticket = 0;
            ctrl->done = 0;
            ctrl->exited = 0;
            ctrl->scanner = 0;
        }
    }
}

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
#ifndef IMM3_EMIT_MIN_BLOCKS
#define IMM3_EMIT_MIN_BLOCKS 4
#endif
#ifndef IMM3_DENSE_MIN_BLOCKS
#define IMM3_DENSE_MIN_BLOCKS 3  // register budget: 3 CTAs (24 compute warps) per SM
#endif
constexpr int kComputeThreads = 256;
constexpr int kComputeWarps = kComputeThreads / 32;
constexpr unsigned kNoMoreTiles = 0xFFFFFFFFu;
static_assert(kDenseThreads == kComputeThreads + 64, "dense kernel: 8 compute warps + producer warp + scanner warp");

struct DenseShared {
    unsigned long long mbar_full[kMaxStages];   // producer -> compute warps: tile id valid, TMA bytes landed
    unsigned long long mbar_empty[kMaxStages];  // compute warps -> producer: slot free again
    unsigned long long mbar_warp[kComputeWarps];  // per compute warp: its projected-column span has landed
    unsigned int tile[kMaxStages];              // tile held by a ring slot
    unsigned int span_cnt[2][kMaxSubtiles * kComputeWarps];  // selected rows of every 1024-row span of the tile
    long long excl[2];
    unsigned int role;
    uint8_t lits[kLitPoolBytes];  // MATCH literals of the plan
    FilterCol filter[kMaxFilterCols];  // the plan's tables (loop-indexed, so not read from the parameter bank)
    ProjCol proj[kMaxProjCols];
};
constexpr unsigned kRoleWorker = 0, kRoleScannerAndWorker = 1, kRoleScannerOnly = 2, kRoleIdle = 3;

__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---- SIMD-within-a-register predicates ----------------------------------------------------------
// TINYINT inclusive range [lo, hi] on 16 raw (two's-complement) bytes -> 16 selection bits.  The low
// seven bits of every byte are range-tested with carry-free byte-wise arithmetic (bit 7 of xl + c1 says
// xl >= t1, bit 7 of c2 - xl says xl <= t2), the sign bit picks which tests apply:
//   MODE 0: 0 <= lo        rows must be non-negative, t1 = lo,       t2 = hi
//   MODE 1: hi < 0         rows must be negative,     t1 = lo + 128, t2 = hi + 128
//   MODE 2: lo < 0 <= hi   negative rows: xl >= lo + 128; non-negative rows: xl <= hi
// The four flag bits of a word (bits 7, 15, 23, 31) are gathered into a nibble by one multiply.
template <int MODE>
__device__ __forceinline__ uint32_t i8_flags(uint32_t x, uint32_t c1, uint32_t c2) {  // bit 7 of every byte: row selected
    const uint32_t xl = x & 0x7F7F7F7Fu;
    const uint32_t g = xl + c1, l = c2 - xl;
    if (MODE == 0) return (g & l & 0x80808080u) & ~x;
    if (MODE == 1) return (g & l & 0x80808080u) & x;
    return ((x & g) | (~x & l)) & 0x80808080u;
}
// The flag bytes (0x80 / 0x00) of two words -> one byte of selection bits, scaled by 128: a byte-wise dot product with
// the weights 1,2,4,8 | 16,32,64,128 (IDP4A accumulates, so a pair costs two instructions off the ALU pipe).
__device__ __forceinline__ uint32_t flags_pair(uint32_t m_lo, uint32_t m_hi) {
    return __dp4a(m_lo, 0x08040201u, __dp4a(m_hi, 0x80402010u, 0u));
}
__device__ __forceinline__ uint32_t range_i32_chunk(const uint4& v, uint32_t lo, uint32_t span) {
    return (uint32_t)((v.x - lo) <= span) | ((uint32_t)((v.y - lo) <= span) << 1) | ((uint32_t)((v.z - lo) <= span) << 2) |
           ((uint32_t)((v.w - lo) <= span) << 3);
}
// Outer perfect shuffle: bit i of the low half goes to bit 2i, bit i of the high half to bit 2i+1.
__device__ __forceinline__ uint32_t zip16(uint32_t x) {
    uint32_t t;
    t = (x ^ (x >> 8)) & 0x0000FF00u; x ^= t ^ (t << 8);
    t = (x ^ (x >> 4)) & 0x00F000F0u; x ^= t ^ (t << 4);
    t = (x ^ (x >> 2)) & 0x0C0C0C0Cu; x ^= t ^ (t << 2);
    t = (x ^ (x >> 1)) & 0x22222222u; x ^= t ^ (t << 1);
    return x;
}

template <bool STAGED>
__device__ __forceinline__ uint4 ld16(uint32_t saddr, const uint8_t* gaddr) {
    if constexpr (STAGED) return lds128(saddr);
    else return ldg128(gaddr);
}

// Selection word of the lane for one filter column: its 32 consecutive rows start at shared address `cell_s`
// (staged tile) or global address `cell_g` (direct loads).  Out of line, everything passed by value: one copy
// of every predicate loop per kernel, and its registers are not the caller's problem.
template <bool STAGED, int MODE>
__device__ __forceinline__ uint32_t eval_i8(uint32_t cell_s, const uint8_t* cell_g, int lane, int lo, int hi) {
    const int t1 = MODE == 0 ? lo : lo + 128;
    const int t2 = MODE == 1 ? hi + 128 : hi;
    const uint32_t c1 = (uint32_t)(128 - t1) * 0x01010101u;
    const uint32_t c2 = (uint32_t)(128 + t2) * 0x01010101u;
    // the lane's two 16-byte chunks, fetched in rotated order so that the 8 lanes of a quarter-warp hit distinct banks
    const int q0 = lane & 1;
    const uint4 a = ld16<STAGED>(cell_s + 16u * q0, cell_g + 16 * q0);
    const uint4 b = ld16<STAGED>(cell_s + 16u * (q0 ^ 1), cell_g + 16 * (q0 ^ 1));
    const uint32_t b0 = flags_pair(i8_flags<MODE>(a.x, c1, c2), i8_flags<MODE>(a.y, c1, c2));
    const uint32_t b1 = flags_pair(i8_flags<MODE>(a.z, c1, c2), i8_flags<MODE>(a.w, c1, c2));
    const uint32_t b2 = flags_pair(i8_flags<MODE>(b.x, c1, c2), i8_flags<MODE>(b.y, c1, c2));
    const uint32_t b3 = flags_pair(i8_flags<MODE>(b.z, c1, c2), i8_flags<MODE>(b.w, c1, c2));
    const uint32_t r = (b0 >> 7) + b1 * 2u + b2 * 512u + b3 * 131072u;  // chunk a = bits 0..15, chunk b = bits 16..31
    return __funnelshift_l(r, r, 16 * q0);                              // un-rotate
}

template <bool STAGED>
__device__ __noinline__ uint32_t eval_filter_span(uint32_t cell_s, const uint8_t* cell_g, int kind, int width, int lo, uint32_t span,
                                                  int nlit, const uint8_t* lits, int lane) {
    if (kind == kFilterI8Range) {
        const int hi = lo + (int)span;
        if (lo >= 0) return eval_i8<STAGED, 0>(cell_s, cell_g, lane, lo, hi);
        if (hi < 0) return eval_i8<STAGED, 1>(cell_s, cell_g, lane, lo, hi);
        return eval_i8<STAGED, 2>(cell_s, cell_g, lane, lo, hi);
    }
    if (kind == kFilterI32Range) {
        uint32_t mask = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const int q = (c + lane) & 7;
            const uint4 v = ld16<STAGED>(cell_s + 16u * q, cell_g + 16 * q);
            mask |= range_i32_chunk(v, (uint32_t)lo, span) << (4 * q);
        }
        return mask;
    }
    if (width == 2) {
        // Two 2-byte cells per word: min(cell ^ literal, 1) is the cell's MISMATCH flag (bits 0 and 16).  The
        // flags of the 16 words of a lane are accumulated as  even rows -> bits 0..15, odd rows -> bits
        // 16..31  and interleaved once at the end.
        uint32_t miss_all = 0xFFFFFFFFu;
        for (int l = 0; l < nlit; l++) {
            const uint32_t ll = ((uint32_t)lits[2 * l] | ((uint32_t)lits[2 * l + 1] << 8)) * 0x00010001u;
            uint32_t miss = 0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int q = (c + lane) & 3;
                const uint4 v = ld16<STAGED>(cell_s + 16u * q, cell_g + 16 * q);
                const uint32_t a = __vminu2(v.x ^ ll, 0x00010001u) + (__vminu2(v.y ^ ll, 0x00010001u) << 1) +
                                   (__vminu2(v.z ^ ll, 0x00010001u) << 2) + (__vminu2(v.w ^ ll, 0x00010001u) << 3);
                miss += a << (4 * q);
            }
            miss_all &= miss;
        }
        return zip16(~miss_all);
    }
    // Generic k-byte cells: row-per-lane compare, ballot gives the bitmap word of rows 32j..32j+31, which lane
    // j keeps.  (cell_s / cell_g are this lane's; the warp's span starts 32*lane rows earlier.)
    const int k = width;
    const uint32_t wbase_s = cell_s - (uint32_t)(lane * 32 * k);
    const uint8_t* wbase_g = cell_g - lane * 32 * k;
    uint32_t mask = 0;
    for (int j = 0; j < 32; j++) {
        const int r = j * 32 + lane;
        bool hit = false;
        for (int l = 0; l < nlit && !hit; l++) {
            bool eq = true;
            for (int b = 0; b < k; b++) {
                const uint32_t cell = STAGED ? lds_u8(wbase_s + (uint32_t)(r * k + b)) : (uint32_t)__ldg(wbase_g + r * k + b);
                eq = eq && (cell == (uint32_t)lits[l * k + b]);
            }
            hit = eq;
        }
        const uint32_t w = __ballot_sync(0xFFFFFFFFu, hit);
        if (lane == j) mask = w;
    }
    return mask;
}

// Selection words of the lane for one filter column: word s covers the lane's 32 rows of the 1024-row
// span starting at tile-relative row `warp_row + s*1024`.  masks[s] is AND-ed in place.  `lits` = the plan's
// literal pool copied to shared memory.
template <int W, bool STAGED>
__device__ __forceinline__ void dense_eval_filter(const FilterCol& f, const uint8_t* lits, uint32_t stage_addr, long long tile_row0,
                                                  int warp_row, int lane, uint32_t* masks) {
    const uint32_t col_s = stage_addr + (uint32_t)f.smem_off;
    const uint8_t* col_g = f.base + tile_row0 * f.width;
#pragma unroll
    for (int s = 0; s < W; s++) {
        const int off = (warp_row + s * 1024 + lane * 32) * f.width;
        masks[s] &= eval_filter_span<STAGED>(col_s + (uint32_t)off, col_g + off, f.kind, f.width, f.lo, f.span, f.nlit, lits + f.lit_off, lane);
    }
}

// Typed shared-memory loads for the projected cells of staged columns.
template <typename T> __device__ __forceinline__ T lds_cell(uint32_t addr);
template <> __device__ __forceinline__ uint8_t lds_cell<uint8_t>(uint32_t addr) { return (uint8_t)lds_u8(addr); }
template <> __device__ __forceinline__ uint16_t lds_cell<uint16_t>(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
template <> __device__ __forceinline__ uint32_t lds_cell<uint32_t>(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
template <> __device__ __forceinline__ unsigned long long lds_cell<unsigned long long>(uint32_t addr) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}

// Cooperative emission of one 1024-row span by one warp: entry i of the warp-private selection
// vector (row index inside the span) goes to out[g0 + i].  Lanes take consecutive entries, so stores are
// coalesced and every lane carries four independent gathers.
template <typename T, bool FROM_SMEM>
__device__ __forceinline__ void emit_span(const unsigned short* sel_w, int n, int lane, uint32_t sbase, const T* __restrict__ gbase,
                                          T* __restrict__ out) {
    constexpr int U = FROM_SMEM ? 4 : 8;  // independent gathers per lane
    for (int i0 = 0; i0 < n; i0 += 32 * U) {
        T v[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            const int i = i0 + k * 32 + lane;
            if (i < n) {
                const uint32_t r = sel_w[i];
                v[k] = FROM_SMEM ? lds_cell<T>(sbase + r * (uint32_t)sizeof(T)) : __ldg(gbase + r);
            }
        }
#pragma unroll
        for (int k = 0; k < U; k++) {
            const int i = i0 + k * 32 + lane;
            if (i < n) out[i] = v[k];
        }
    }
}
// Any other cell width: byte-wise.
__device__ __forceinline__ void emit_span_bytes(const unsigned short* sel_w, int n, int lane, bool from_smem, uint32_t sbase,
                                                const uint8_t* __restrict__ gbase, uint8_t* __restrict__ out, int w) {
    for (int i = lane; i < n; i += 32) {
        const uint32_t r = sel_w[i];
        for (int b = 0; b < w; b++)
            out[(long long)i * w + b] = (uint8_t)(from_smem ? lds_u8(sbase + r * (uint32_t)w + b) : (uint32_t)__ldg(gbase + (long long)r * w + b));
    }
}

// One projected column of a selection vector: n selected rows (already clamped to the LIMIT), entry i goes to
// out[i].  Out of line (one copy of the width dispatch per kernel) with everything passed in registers: the
// plan lives in the kernel's parameter bank and must not be dereferenced through a pointer here.
__device__ __noinline__ void emit_col(const unsigned short* sel_w, int n, int lane, int w, bool from_smem, uint32_t sbase,
                                      const uint8_t* __restrict__ gbase, uint8_t* __restrict__ out) {
    if (w == 4) {
        if (from_smem) emit_span<uint32_t, true>(sel_w, n, lane, sbase, nullptr, (uint32_t*)out);
        else emit_span<uint32_t, false>(sel_w, n, lane, 0u, (const uint32_t*)gbase, (uint32_t*)out);
    } else if (w == 1) {
        if (from_smem) emit_span<uint8_t, true>(sel_w, n, lane, sbase, nullptr, out);
        else emit_span<uint8_t, false>(sel_w, n, lane, 0u, gbase, out);
    } else if (w == 2) {
        if (from_smem) emit_span<uint16_t, true>(sel_w, n, lane, sbase, nullptr, (uint16_t*)out);
        else emit_span<uint16_t, false>(sel_w, n, lane, 0u, (const uint16_t*)gbase, (uint16_t*)out);
    } else if (w == 8) {
        if (from_smem) emit_span<unsigned long long, true>(sel_w, n, lane, sbase, nullptr, (unsigned long long*)out);
        else emit_span<unsigned long long, false>(sel_w, n, lane, 0u, (const unsigned long long*)gbase, (unsigned long long*)out);
    } else {
        emit_span_bytes(sel_w, n, lane, from_smem, sbase, gbase, out, w);
    }
}

// All projected columns of a selection vector whose rows are relative to tile row `span_row`; the first entry
// goes to result ordinal g0.
// (`proj` / `filter` = the plan's tables copied to shared memory: indexing the kernel's parameter bank with a
// loop variable would make the compiler unroll or spill the whole plan.)
__device__ __forceinline__ void emit_span_all(const ProjCol* proj, int nproj, const FilterCol* filter, const unsigned short* sel_w, int n,
                                              int lane, bool staged, uint32_t stage_addr, int span_row, long long tile_row0, long long g0) {
#pragma unroll 1
    for (int pc = 0; pc < nproj; pc++) {
        const ProjCol& pj = proj[pc];
        const int w = pj.width;
        const bool from_smem = staged && pj.filter_idx >= 0;
        const uint32_t sbase = stage_addr + (from_smem ? (uint32_t)filter[pj.filter_idx].smem_off : 0u) + (uint32_t)(span_row * w);
        emit_col(sel_w, n, lane, w, from_smem, sbase, pj.base + (tile_row0 + span_row) * w, pj.out + g0 * w);
    }
}

// A span whose 1024 rows all survive: straight coalesced copy of n <= 1024 rows, no selection vector.
__device__ __noinline__ void copy_rows(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int bytes, int lane) {
    if ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)bytes) & 3u) == 0) {
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d4 = reinterpret_cast<uint32_t*>(dst);
        for (int i = lane; i < bytes / 4; i += 32) d4[i] = __ldg(s4 + i);
    } else {
        for (int i = lane; i < bytes; i += 32) dst[i] = __ldg(src + i);
    }
}
__device__ __forceinline__ void emit_span_full(const ProjCol* proj, int nproj, int lane, long long row0, long long g0, int n) {
#pragma unroll 1
    for (int pc = 0; pc < nproj; pc++) {
        const ProjCol& pj = proj[pc];
        const int w = pj.width;
        copy_rows(pj.base + row0 * w, pj.out + g0 * w, n * w, lane);
    }
}

// Bitmap word of the lane -> entries appended to the warp's selection vector (ascending row order): the
// rows of word `mm` are row_base + 32*lane + bit.
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void append_selection(uint32_t mm, int lane, unsigned short* sel_at, unsigned row_base) {
    const unsigned cnt = (unsigned)__popc(mm);
    unsigned incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += nb;
    }
    uint32_t addr = smem_u32(sel_at) + 2u * (incl - cnt);
    const uint32_t base = row_base + (unsigned)lane * 32u;
    uint32_t rm = __brev(mm);  // leading zeros of rm = index of the lowest set bit of mm
    while (rm) {
        const int b = __clz((int)rm);
        sts_u16(addr, base + (uint32_t)b);
        addr += 2u;
        rm &= ~(0x80000000u >> b);
    }
}

// Position in a ring of `ring` slots, advanced without integer division.
struct RingPos {
    int slot = 0;
    unsigned use = 0;  // how many times the ring has wrapped
    __device__ __forceinline__ void advance(int ring) {
        if (++slot == ring) {
            slot = 0;
            use++;
        }
    }
};

extern __shared__ __align__(128) uint8_t dyn_smem[];

// The plan's per-column tables -> shared memory.  Every access to P uses a compile-time index (fully unrolled
// selects), so the parameter bank is never indexed dynamically.
__device__ __forceinline__ void copy_plan_tables(const ScanPlan& P, FilterCol* filter, ProjCol* proj, int tid, int nthreads) {
    (void)nthreads;
    if (tid < P.nfilter) {
#pragma unroll
        for (int i = 0; i < kMaxFilterCols; i++)
            if (tid == i) filter[i] = P.filter[i];
    } else if (tid >= 32 && tid < 32 + P.nproj) {
#pragma unroll
        for (int i = 0; i < kMaxProjCols; i++)
            if (tid - 32 == i) proj[i] = P.proj[i];
    }
}

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

// =============================================================================================
// Block-mode kernel (sorted-integer codec and general fallback)
// =============================================================================================
struct BlockShared {
    unsigned int ticket;
    unsigned int done;
    unsigned int warp_cnt[kBlockThreads / 32];
    long long tile_excl;
    unsigned int vb_start;
};

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// Decode one PFOR_INT block (PFORCodecInt.encode's inverse) into vals[0..n).
// W: scratch for the byte-swapped words; mb_pos/mb_bits/mb_tot/mb_base: per-mini-block scratch.
__device__ void pfor_decode_block(const PforCol& pc, long long blk, int n, uint32_t* W, uint32_t* vals,
                                  unsigned short* mb_pos, unsigned char* mb_bits, uint32_t* mb_tot, uint32_t* mb_base,
                                  BlockShared& S) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t w0 = pc.word_off[blk], w1 = pc.word_off[blk + 1];
    const int nw = (int)(w1 - w0) - 2;  // PFORCodecInt.encode appends 8 zero bytes (PFORCodec.scala:20)
    for (int i = tid; i < nw; i += kBlockThreads) W[i] = bswap32(__ldg(pc.words + w0 + i));  // putInt is big-endian
    __syncthreads();
    const int packed = n & ~31, nmini = packed >> 5;
    if (tid == 0) {  // walk the headers: one word per 128-value super-block, then one per left-over mini-block
        int ip = 1, m = 0, s = 0;
        for (; s + 128 <= packed; s += 128) {
            const uint32_t h = W[ip++];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int b = (int)((h >> (24 - 8 * q)) & 0xFFu);
                mb_pos[m] = (unsigned short)ip;
                mb_bits[m] = (unsigned char)b;
                ip += b;
                m++;
            }
        }
        for (; s < packed; s += 32) {
            const int b = (int)W[ip++];
            mb_pos[m] = (unsigned short)ip;
            mb_bits[m] = (unsigned char)b;
            ip += b;
            m++;
        }
        S.vb_start = (unsigned)ip;
    }
    __syncthreads();
    for (int m = warp; m < nmini; m += kBlockThreads / 32) {
        const int b = mb_bits[m];
        const int p = mb_pos[m];
        uint32_t d;
        if (b == 32) {
            d = W[p + lane];  // raw values, not deltas
        } else if (b == 0) {
            d = 0;
        } else {
            const int off = lane * b, wi = p + (off >> 5), sh = off & 31;
            const uint32_t lo = W[wi];
            const uint32_t hi = (sh + b > 32) ? W[wi + 1] : 0u;
            d = __funnelshift_r(lo, hi, sh) & ((1u << b) - 1u);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {  // running sum of the deltas inside the mini-block
                const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, d, o);
                if (lane >= o) d += nb;
            }
        }
        vals[m * 32 + lane] = d;
        if (lane == 31) mb_tot[m] = d;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t base = 0;  // initvalue = 0 at every block
        for (int m = 0; m < nmini; m++) {
            mb_base[m] = base;
            base = (mb_bits[m] == 32) ? mb_tot[m] : base + mb_tot[m];
        }
        // var-byte remainder (n % 32 values): 7-bit groups, low first, last byte has bit 7 set
        int ip = (int)S.vb_start, sh = 0, shift = 0;
        uint32_t v = 0;
        for (int k = packed; k < n;) {
            const uint32_t c = W[ip] >> sh;
            sh += 8;
            ip += sh >> 5;
            sh &= 31;
            v += (c & 127u) << shift;
            if (c & 128u) {
                base += v;
                vals[k++] = base;
                v = 0;
                shift = 0;
            } else {
                shift += 7;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < packed; i += kBlockThreads) {
        const int m = i >> 5;
        if (mb_bits[m] != 32) vals[i] += mb_base[m];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kBlockThreads) scan_blocks_kernel(const __grid_constant__ ScanPlan P, ScanCtrl* ctrl,
                                                                      unsigned long long* status) {
    __shared__ BlockShared S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kWarps = kBlockThreads / 32;

    // carve the dynamic shared memory
    const int maxb = (P.max_block_rows + 31) & ~31;
    const int nmb = maxb / 32;
    uint32_t* vals = reinterpret_cast<uint32_t*>(dyn_smem);  // [npfor][maxb]
    uint32_t* W = vals + (size_t)(P.npfor > 0 ? P.npfor : 0) * maxb;
    const int wcap = P.npfor > 0 ? (maxb + nmb + 64) : 0;
    uint32_t* bm = W + wcap;          // [nmb]   selection bitmap words of the block
    uint32_t* woff = bm + nmb;        // [nmb+1] exclusive popcount prefix
    uint32_t* mb_tot = woff + nmb + 1;
    uint32_t* mb_base = mb_tot + nmb;
    unsigned short* mb_pos = reinterpret_cast<unsigned short*>(mb_base + nmb);
    unsigned char* mb_bits = reinterpret_cast<unsigned char*>(mb_pos + nmb);

    const unsigned ntiles = (unsigned)P.ntiles;
    for (;;) {
        if (tid == 0) {
            S.ticket = atomicAdd(&ctrl->ticket, 1u);
            S.done = ld_relaxed_u32(&ctrl->done);
        }
        __syncthreads();
        const unsigned blk = S.ticket;
        if (blk >= ntiles || S.done) break;
        const long long R0 = (long long)P.row_start[blk];
        const int n = (int)((long long)P.row_start[blk + 1] - R0);
        const int nwords = (n + 31) >> 5;

        for (int s = 0; s < P.npfor; s++)
            pfor_decode_block(P.pfor[s], blk, n, W, vals + (size_t)s * maxb, mb_pos, mb_bits, mb_tot, mb_base, S);

        // ---- conjunctive filter, row per lane; ballot builds the block's bitmap words ----
        for (int wd = warp; wd < nwords; wd += kWarps) {
            const int i = wd * 32 + lane;
            bool pass = i < n;
            for (int fi = 0; fi < P.nfilter; fi++) {
                const FilterCol& f = P.filter[fi];
                if (f.kind == kFilterI32Range) {
                    uint32_t v = 0;
                    if (pass) v = f.pfor_slot >= 0 ? vals[(size_t)f.pfor_slot * maxb + i]
                                                   : __ldg(reinterpret_cast<const uint32_t*>(f.base) + R0 + i);
                    pass = pass && ((v - (uint32_t)f.lo) <= f.span);
                } else if (f.kind == kFilterI8Range) {
                    int v = 0;
                    if (pass) v = (int)(signed char)__ldg(f.base + R0 + i);
                    pass = pass && ((uint32_t)(v - f.lo) <= f.span);
                } else {
                    bool hit = false;
                    if (pass) {
                        const uint8_t* cell = f.base + (R0 + i) * f.width;
                        for (int l = 0; l < f.nlit && !hit; l++) {
                            bool eq = true;
                            for (int b = 0; b < f.width; b++) eq = eq && (__ldg(cell + b) == P.lits[f.lit_off + l * f.width + b]);
                            hit = eq;
                        }
                    }
                    pass = pass && hit;
                }
            }
            const uint32_t word = __ballot_sync(0xFFFFFFFFu, pass);
            if (lane == 0) bm[wd] = word;
        }
        __syncthreads();

        // ---- exclusive scan of the word popcounts (nwords <= kBlockThreads) ----
        const unsigned cnt = tid < nwords ? __popc(bm[tid]) : 0u;
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += nb;
        }
        if (lane == 31) S.warp_cnt[warp] = incl;
        __syncthreads();
        unsigned warp_base = 0, tile_count = 0;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const unsigned c = S.warp_cnt[w];
            if (w < warp) warp_base += c;
            tile_count += c;
        }
        if (tid < nwords) woff[tid] = warp_base + incl - cnt;

        if (warp == 0) {
            const long long excl = resolve_tile(P, ctrl, status, blk, tile_count, lane);
            if (lane == 0) S.tile_excl = excl;
        }
        if (P.bitmap) {  // blocks start at arbitrary bit positions of the global bitmap
            for (int wd = tid; wd < nwords; wd += kBlockThreads) {
                const uint32_t word = bm[wd];
                if (!word) continue;
                const long long g = R0 + (long long)wd * 32;
                const int sh = (int)(g & 31);
                atomicOr(&P.bitmap[g >> 5], word << sh);
                if (sh) atomicOr(&P.bitmap[(g >> 5) + 1], word >> (32 - sh));
            }
        }
        __syncthreads();

        // ---- Project ----
        const long long excl = S.tile_excl;
        if (!P.bitmap && excl >= 0 && excl < P.limit) {
            const long long room = P.limit - excl;
            const unsigned n_emit = room < (long long)tile_count ? (unsigned)room : tile_count;
            for (int wd = warp; wd < nwords; wd += kWarps) {
                const uint32_t word = bm[wd];
                if (!((word >> lane) & 1u)) continue;
                const unsigned rank = woff[wd] + __popc(word & ((1u << lane) - 1u));
                if (rank >= n_emit) continue;
                const int i = wd * 32 + lane;
                for (int pc = 0; pc < P.nproj; pc++) {
                    const ProjCol& pj = P.proj[pc];
                    uint8_t* dst = pj.out + (excl + rank) * pj.width;
                    if (pj.pfor_slot >= 0) {
                        *reinterpret_cast<uint32_t*>(dst) = vals[(size_t)pj.pfor_slot * maxb + i];
                    } else if (pj.width == 4) {
                        *reinterpret_cast<uint32_t*>(dst) = __ldg(reinterpret_cast<const uint32_t*>(pj.base) + R0 + i);
                    } else {
                        const uint8_t* src = pj.base + (R0 + i) * pj.width;
                        for (int b = 0; b < pj.width; b++) dst[b] = __ldg(src + b);
                    }
                }
            }
        }
        __syncthreads();
    }
    cta_exit(ctrl);
}

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
    if (S.is_last && warp < kComputeWarps) {
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
// per warp, in words: the filter kernel keeps the block's byte-swapped words and ONE decoded column; the emit kernel keeps
// the words, every decoded column of the select list, their mini-block bases and a 1024-entry selection vector
__host__ __device__ constexpr int blk_warp_smem_words(int npfor, int words_cap) { return words_cap + (npfor > 0 ? kBlkVals : 0); }
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

__global__ void __launch_bounds__(kComputeThreads, 4) blocks_filter_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmapB,
                                                                             uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ tile_cnt,
                                                                             unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl,
                                                                             long long nblocks) {
    __shared__ FilterShared S;
    __shared__ PforCol s_pfor[kMaxPforCols];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < P.lit_bytes; i += kComputeThreads) S.lits[i] = P.lits[i];  // (nothing to copy unless a MATCH predicate exists)
    copy_plan_tables(P, S.filter, S.proj, tid, kComputeThreads);
    if (tid < kMaxPforCols) {
#pragma unroll
        for (int i = 0; i < kMaxPforCols; i++)
            if (tid == i) s_pfor[i] = P.pfor[i];
    }
    __syncthreads();
    uint32_t* const Wb = reinterpret_cast<uint32_t*>(dyn_smem) + warp * blk_warp_smem_words(P.npfor, P.blk_words_cap);
    uint32_t* const vals0 = Wb + P.blk_words_cap;
    const long long ntiles = P.ntiles;  // tiles of 8 blocks

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long blk = tile * kComputeWarps + warp;
        uint32_t myword = 0;  // lane w keeps bitmap word w of the block
        unsigned cnt = 0;
        if (blk < nblocks) {
            // block metadata in ONE round trip: lanes 0,1 = row ordinals, lanes 2+2s, 3+2s = word offsets of PFOR slot s
            unsigned long long meta = 0;
            if (lane < 2) meta = P.row_start[blk + lane];
            else if (lane < 2 + 2 * P.npfor) meta = s_pfor[(lane - 2) >> 1].word_off[blk + (lane & 1)];
            const long long R0 = (long long)__shfl_sync(0xFFFFFFFFu, meta, 0);
            const int n = (int)((long long)__shfl_sync(0xFFFFFFFFu, meta, 1) - R0);
            const int nwords = (n + 31) >> 5;
            {   // rows of the block that exist: lane w owns word w = rows [32w, 32w+32)
                const int left = n - lane * 32;
                myword = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
            }
            int decoded = -1;   // PFOR slot whose block sits in vals0 (one buffer: the filter kernel keeps no decoded column)
            uint32_t base = 0;
#pragma unroll 1
            for (int fi = 0; fi < P.nfilter; fi++) {
                const FilterCol f = S.filter[fi];
                if (f.kind == kFilterI32Range && f.pfor_slot >= 0) {
                    // decoded column: lane m tests its own mini-block (the values it just unpacked) - no ballots
                    if (f.pfor_slot != decoded) {  // (two predicates on one column share the decode)
                        const uint32_t wo0 = (uint32_t)__shfl_sync(0xFFFFFFFFu, meta, 2 + 2 * f.pfor_slot);
                        const uint32_t wo1 = (uint32_t)__shfl_sync(0xFFFFFFFFu, meta, 3 + 2 * f.pfor_slot);
                        base = pfor_decode_warp(s_pfor[f.pfor_slot].words, wo0, wo1, n, Wb, P.blk_words_cap, vals0, lane);
                        decoded = f.pfor_slot;
                    }
                    const uint32_t* vp = vals0 + lane * kBlkLane;
                    const uint32_t lo = (uint32_t)f.lo - base;
                    uint32_t word = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const uint4 v = reinterpret_cast<const uint4*>(vp)[q];
                        word |= ((uint32_t)((v.x - lo) <= f.span) | ((uint32_t)((v.y - lo) <= f.span) << 1) | ((uint32_t)((v.z - lo) <= f.span) << 2) |
                                 ((uint32_t)((v.w - lo) <= f.span) << 3))
                                << (4 * q);
                    }
                    myword &= word;
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
            bitmapB[blk * 32 + lane] = myword;
            cnt = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(myword));
            if (lane == 0) blk_cnt[blk] = cnt;
        }
        if (lane == 0 && cnt) atomicAdd(tile_cnt + tile, cnt);  // (the host zeroes the tile counts before the launch)
    }

    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&ctrl->exited, 1u);
        S.is_last = prev == gridDim.x - 1;
        if (S.is_last) ctrl->exited = 0;
    }
    __syncthreads();
    if (S.is_last) {
        __threadfence();
        scan_tile_counts(S, tile_cnt, tile_off, ntiles, P.limit, ctrl);
    }
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
            myword = __ldg(bitmap + blk * 32 + lane);
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
        append_selection(myword, lane, sel_w, 0u);
        __syncwarp();
        const int nn = (int)(P.limit - g < (long long)cnt ? P.limit - g : (long long)cnt);
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

// =============================================================================================
// Launchers
// =============================================================================================
size_t blocks_kernel_smem_bytes(int npfor, int max_block_rows) {
    const size_t maxb = (size_t)((max_block_rows + 31) & ~31), nmb = maxb / 32;
    size_t words = (size_t)npfor * maxb + (npfor > 0 ? maxb + nmb + 64 : 0) + nmb + (nmb + 1) + nmb + nmb;
    return words * 4 + nmb * 2 + nmb + 64;
}

static cudaError_t configure_once() {
    static cudaError_t rc = [] {
        cudaError_t e = cudaSuccess;
#define IMM3_SET_SMEM(K) if ((e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e
        IMM3_SET_SMEM(emit_stream_kernel);
        IMM3_SET_SMEM(blocks_filter_kernel);
        IMM3_SET_SMEM(blocks_emit_kernel<true>);
        IMM3_SET_SMEM(blocks_emit_kernel<false>);
        IMM3_SET_SMEM(emit_general_kernel);
        IMM3_SET_SMEM((scan_dense_kernel<true>));
        IMM3_SET_SMEM((scan_dense_kernel<false>));
        IMM3_SET_SMEM(filter_kernel<true>);
        IMM3_SET_SMEM(filter_kernel<false>);
#undef IMM3_SET_SMEM
        return cudaFuncSetAttribute(scan_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }();
    return rc;
}

cudaError_t dense_kernel_occupancy(bool staged, size_t dyn_smem, int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (staged) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, scan_dense_kernel<true>, kDenseThreads, dyn_smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, scan_dense_kernel<false>, kDenseThreads, dyn_smem);
}
cudaError_t blocks_kernel_occupancy(size_t dyn_smem, int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, scan_blocks_kernel, kBlockThreads, dyn_smem);
}

cudaError_t launch_scan_dense(const ScanPlan& plan, ScanCtrl* ctrl, unsigned long long* status, int grid, size_t dyn_smem,
                              cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (plan.stages > 0) scan_dense_kernel<true><<<grid, kDenseThreads, dyn_smem, stream>>>(plan, ctrl, status);
    else scan_dense_kernel<false><<<grid, kDenseThreads, dyn_smem, stream>>>(plan, ctrl, status);
    return cudaGetLastError();
}
cudaError_t filter_kernel_occupancy(size_t dyn_smem, int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (dyn_smem > 0) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, filter_kernel<true>, kComputeThreads + 32, dyn_smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, filter_kernel<false>, kComputeThreads + 32, dyn_smem);
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

cudaError_t emit_kernel_occupancy(bool general, int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (general) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, emit_general_kernel, kComputeThreads, kComputeWarps * 1024 * 2);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, emit_kernel, kComputeThreads, kComputeWarps * kEmitWarpSmemBytes);
}
cudaError_t launch_filter(const ScanPlan& plan, uint32_t* bitmap, uint32_t* span_cnt, uint32_t* tile_cnt, unsigned long long* tile_off,
                          ScanCtrl* ctrl, int grid, size_t dyn_smem, cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (plan.nfilter == 0) {  // no predicate: nothing to read - the selection metadata is written analytically
        select_all_kernel<<<grid, kComputeThreads, 0, stream>>>(plan, bitmap, span_cnt, tile_cnt, tile_off, ctrl);
        return cudaGetLastError();
    }
    if (plan.stages > 0) filter_kernel<true><<<grid, kComputeThreads + 32, dyn_smem, stream>>>(plan, bitmap, span_cnt, tile_cnt, tile_off, ctrl);
    else filter_kernel<false><<<grid, kComputeThreads + 32, dyn_smem, stream>>>(plan, bitmap, span_cnt, tile_cnt, tile_off, ctrl);
    return cudaGetLastError();
}
cudaError_t launch_emit(const ScanPlan& plan, const uint32_t* bitmap, const uint32_t* span_cnt, const unsigned long long* tile_off,
                        int spans_per_tile, long long nspans, int grid, int dense_off, const ScanCtrl* ctrl, bool general, bool pdl,
                        cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kComputeThreads);
    cfg.dynamicSmemBytes = kComputeWarps * kEmitWarpSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (general) {
        cfg.dynamicSmemBytes = kComputeWarps * 1024 * 2;
        return cudaLaunchKernelEx(&cfg, emit_general_kernel, plan, bitmap, span_cnt, tile_off, spans_per_tile, nspans, dense_off, ctrl);
    }
    return cudaLaunchKernelEx(&cfg, emit_kernel, plan, bitmap, span_cnt, tile_off, spans_per_tile, nspans, dense_off, ctrl);
}
size_t emit_stream_smem_bytes(int stage_bytes, int ring) { return (size_t)kComputeWarps * 1024 * 2 + (size_t)ring * (size_t)stage_bytes + 16; }
int emit_stream_header_bytes() { return kEmitHdrBytes; }
cudaError_t emit_stream_occupancy(size_t dyn_smem, int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, emit_stream_kernel, kComputeThreads + 32, dyn_smem);
}
cudaError_t launch_emit_stream(const ScanPlan& plan, const uint32_t* bitmap, const uint32_t* span_cnt, const uint32_t* tile_cnt,
                               const unsigned long long* tile_off, long long nsub, int ring, int stage_bytes, int dense_mode, int grid,
                               size_t dyn_smem, ScanCtrl* ctrl, bool pdl, cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kComputeThreads + 32);
    cfg.dynamicSmemBytes = dyn_smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, emit_stream_kernel, plan, bitmap, span_cnt, tile_cnt, tile_off, nsub, ring, stage_bytes, dense_mode, ctrl);
}

size_t blocks_multi_smem_bytes(int npfor, int words_cap) { return (size_t)kComputeWarps * blk_warp_smem_words(npfor, words_cap) * 4; }
size_t blocks_emit_smem_bytes(int npfor, int words_cap) { return (size_t)kComputeWarps * blk_emit_warp_words(npfor, words_cap) * 4; }
cudaError_t blocks_multi_occupancy(size_t filter_smem, size_t emit_smem, int* filter_blocks_per_sm, int* emit_blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (filter_blocks_per_sm) {
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(filter_blocks_per_sm, blocks_filter_kernel, kComputeThreads, filter_smem);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(emit_blocks_per_sm, blocks_emit_kernel<true>, kComputeThreads, emit_smem);
}
cudaError_t launch_blocks_filter(const ScanPlan& plan, uint32_t* bitmapB, uint32_t* blk_cnt, uint32_t* tile_cnt, unsigned long long* tile_off,
                                 ScanCtrl* ctrl, long long nblocks, int grid, size_t dyn_smem, cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    blocks_filter_kernel<<<grid, kComputeThreads, dyn_smem, stream>>>(plan, bitmapB, blk_cnt, tile_cnt, tile_off, ctrl, nblocks);
    return cudaGetLastError();
}
cudaError_t launch_blocks_emit(const ScanPlan& plan, const uint32_t* bitmap, const uint32_t* cnts, const unsigned long long* tile_off,
                               long long nblocks, const ScanCtrl* ctrl, bool rowspace, int grid, size_t dyn_smem, cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (rowspace) blocks_emit_kernel<true><<<grid, kComputeThreads, dyn_smem, stream>>>(plan, bitmap, cnts, tile_off, nblocks, ctrl);
    else blocks_emit_kernel<false><<<grid, kComputeThreads, dyn_smem, stream>>>(plan, bitmap, cnts, tile_off, nblocks, ctrl);
    return cudaGetLastError();
}

cudaError_t launch_scan_blocks(const ScanPlan& plan, ScanCtrl* ctrl, unsigned long long* status, int grid, size_t dyn_smem,
                               cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    scan_blocks_kernel<<<grid, kBlockThreads, dyn_smem, stream>>>(plan, ctrl, status);
    return cudaGetLastError();
}

}  // namespace imm3

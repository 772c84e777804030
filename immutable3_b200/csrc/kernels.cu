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

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, ScanCtrl* ctrl) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = globaltimer_ns();
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > kWatchdogNs) watchdog_trap(ctrl, 1);
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
        }
    }
}

// =============================================================================================
// Dense kernel
//
// CTA = 8 compute warps + 1 control warp.  A tile is 8192*W consecutive rows (W = 1, 2 or 4 bitmap
// words per lane): warp w owns rows [w*1024*W, (w+1)*1024*W) of the tile, split into W sub-spans of
// 1024 rows in which lane l owns rows [32*l, 32*l+32) = one 32-bit word of the selection bitmap.
//
// The tiles of a CTA are software-pipelined so that nothing waits on the device-wide prefix:
//
//   compute warps :  F(0) | F(1) E(0) | F(2) E(1) | ...
//       F(j) "filter" : wait for tile j's bytes (TMA -> shared memory, mbarrier), evaluate the conjunction
//                       with SWAR compares, popc + warp/CTA scan -> every lane knows the tile-local rank
//                       of its first selected row; PUBLISH the tile count.  The lane keeps its bitmap
//                       words and ranks in registers.
//       E(j) "emit"   : pick up the tile's global offset (resolved while F(j+1) ran), walk the set bits
//                       four at a time: gather the projected cells (staged filter columns from shared
//                       memory, other columns from global memory) and store them at offset + rank.
//                       Then recycle the tile's ring slot (ticket + TMA bulk copies).
//   control warp  :  for each tile in order: take its count, publish it, run the decoupled look-back,
//                    hand the exclusive offset back.  It overlaps F(j+1)/E(j) of the compute warps.
//
// Publishing count(j+1) BEFORE emitting tile j matters: tickets are drawn ahead (for the TMA ring), so
// a CTA busy emitting would otherwise sit on an un-counted earlier tile that every later tile in the
// grid has to wait for.  Hand-offs are two-entry rings guarded by mbarriers (count_full, excl_full).
// =============================================================================================
#ifndef IMM3_EMIT_MIN_BLOCKS
#define IMM3_EMIT_MIN_BLOCKS 2
#endif
#ifndef IMM3_DENSE_MIN_BLOCKS
#define IMM3_DENSE_MIN_BLOCKS 3  // register budget: 3 CTAs (27 warps) per SM
#endif
constexpr int kComputeThreads = 256;
constexpr int kComputeWarps = kComputeThreads / 32;
constexpr unsigned kNoMoreTiles = 0xFFFFFFFFu;

struct DenseShared {
    unsigned long long mbar_full[kMaxStages];  // TMA bytes of a ring slot have landed
    unsigned long long mbar_count[2];          // compute -> control: tile_count / tile_id valid
    unsigned long long mbar_excl[2];           // control -> compute: tile_excl valid
    unsigned int ticket[kMaxStages];
    int issued[kMaxStages];
    unsigned int warp_cnt[2][kComputeWarps];
    unsigned int tile_count[2];
    unsigned int tile_id[2];
    long long tile_excl[2];
};

__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// TINYINT range over the lane's 32 consecutive rows (two 16-byte chunks).  Values are biased to
// unsigned order (x ^ 0x80) and tested in 16-bit SWAR lanes:  bit 8 of (e + 256 - lo) says e >= lo,
// bit 8 of ((256 | hi) - e) says e <= hi.
__device__ __forceinline__ uint32_t swar_i8_word(uint32_t x, uint32_t c1, uint32_t c2) {
    const uint32_t y = x ^ 0x80808080u;
    const uint32_t e = y & 0x00FF00FFu;
    const uint32_t o = (y >> 8) & 0x00FF00FFu;
    const uint32_t pe = (e + c1) & (c2 - e);
    const uint32_t po = (o + c1) & (c2 - o);
    const uint32_t z = ((pe >> 8) & 0x00010001u) | ((po >> 7) & 0x00020002u);
    return (z | (z >> 14)) & 0xFu;
}
__device__ __forceinline__ uint32_t swar_i8_chunk(const uint4& v, uint32_t c1, uint32_t c2) {
    return swar_i8_word(v.x, c1, c2) | (swar_i8_word(v.y, c1, c2) << 4) | (swar_i8_word(v.z, c1, c2) << 8) |
           (swar_i8_word(v.w, c1, c2) << 12);
}
__device__ __forceinline__ uint32_t range_i32_chunk(const uint4& v, uint32_t lo, uint32_t span) {
    return (uint32_t)((v.x - lo) <= span) | ((uint32_t)((v.y - lo) <= span) << 1) | ((uint32_t)((v.z - lo) <= span) << 2) |
           ((uint32_t)((v.w - lo) <= span) << 3);
}
// Two 2-byte cells per word; a halfword of t is zero iff bit 15/31 of the result is set.
__device__ __forceinline__ uint32_t zero_halfwords(uint32_t t) {
    return ~(((t & 0x7FFF7FFFu) + 0x7FFF7FFFu) | t) & 0x80008000u;
}

// Selection words of the lane for one filter column: word s covers the lane's 32 rows of the 1024-row
// sub-span starting at tile-relative row `warp_row + s*1024`.  masks[s] is AND-ed in place.  One
// out-of-line call per (tile, filter column) keeps the kernels small and amortises the setup.
template <int W>
__device__ __noinline__ void dense_eval_filter(const ScanPlan& P, const FilterCol& f, bool staged, uint32_t stage_addr,
                                               long long tile_row0, int warp_row, int lane, uint32_t* masks) {
    const uint32_t col_s = stage_addr + (uint32_t)f.smem_off;
    const uint8_t* col_g = f.base + tile_row0 * f.width;
    if (f.kind == kFilterI8Range) {
        const uint32_t lo_b = ((uint32_t)f.lo ^ 0x80u) & 0xFFu;
        const uint32_t hi_b = lo_b + f.span;
        const uint32_t c1 = (0x100u - lo_b) * 0x00010001u;
        const uint32_t c2 = (0x100u | hi_b) * 0x00010001u;
#pragma unroll
        for (int s = 0; s < W; s++) {
            const uint32_t off = (uint32_t)(warp_row + s * 1024 + lane * 32);
            uint32_t mask = 0;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int q = (c + lane) & 1;  // rotate so the 8 lanes of a quarter-warp hit distinct banks
                const uint4 v = staged ? lds128(col_s + off + 16u * q) : ldg128(col_g + off + 16 * q);
                mask |= swar_i8_chunk(v, c1, c2) << (16 * q);
            }
            masks[s] &= mask;
        }
    } else if (f.kind == kFilterI32Range) {
        const uint32_t lo = (uint32_t)f.lo, span = f.span;
#pragma unroll
        for (int s = 0; s < W; s++) {
            const uint32_t off = (uint32_t)(warp_row + s * 1024 + lane * 32) * 4u;
            uint32_t mask = 0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const int q = (c + lane) & 7;
                const uint4 v = staged ? lds128(col_s + off + 16u * q) : ldg128(col_g + off + 16 * q);
                mask |= range_i32_chunk(v, lo, span) << (4 * q);
            }
            masks[s] &= mask;
        }
    } else if (f.width == 2) {
#pragma unroll
        for (int s = 0; s < W; s++) {
            const uint32_t off = (uint32_t)(warp_row + s * 1024 + lane * 32) * 2u;
            uint32_t mask = 0;
            for (int l = 0; l < f.nlit; l++) {
                const uint32_t lit = (uint32_t)P.lits[f.lit_off + 2 * l] | ((uint32_t)P.lits[f.lit_off + 2 * l + 1] << 8);
                const uint32_t ll = lit * 0x00010001u;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int q = (c + lane) & 3;
                    const uint4 v = staged ? lds128(col_s + off + 16u * q) : ldg128(col_g + off + 16 * q);
                    const uint32_t h0 = zero_halfwords(v.x ^ ll), h1 = zero_halfwords(v.y ^ ll);
                    const uint32_t h2 = zero_halfwords(v.z ^ ll), h3 = zero_halfwords(v.w ^ ll);
                    // bits 15/31 of h* -> two row bits each
                    const uint32_t b01 = (h0 >> 15) | (h1 >> 13), b23 = (h2 >> 11) | (h3 >> 9);
                    const uint32_t lo4 = (b01 & 0x5u) | ((b01 >> 15) & 0xAu);
                    const uint32_t hi4 = ((b23 >> 4) & 0x5u) | ((b23 >> 19) & 0xAu);
                    mask |= (lo4 | (hi4 << 4)) << (8 * q);
                }
            }
            masks[s] &= mask;
        }
    } else {
        // Generic k-byte cells: row-per-lane compare, ballot gives the bitmap word of rows 32j..32j+31,
        // which lane j keeps.
        const int k = f.width;
        for (int s = 0; s < W; s++) {
            const uint32_t wbase_s = col_s + (uint32_t)((warp_row + s * 1024) * k);
            const uint8_t* wbase_g = col_g + (long long)(warp_row + s * 1024) * k;
            uint32_t mask = 0;
            for (int j = 0; j < 32; j++) {
                const int r = j * 32 + lane;
                bool hit = false;
                for (int l = 0; l < f.nlit && !hit; l++) {
                    bool eq = true;
                    for (int b = 0; b < k; b++) {
                        const uint32_t cell = staged ? lds_u8(wbase_s + (uint32_t)(r * k + b)) : (uint32_t)__ldg(wbase_g + r * k + b);
                        eq = eq && (cell == (uint32_t)P.lits[f.lit_off + l * k + b]);
                    }
                    hit = eq;
                }
                const uint32_t w = __ballot_sync(0xFFFFFFFFu, hit);
                if (lane == j) mask = w;
            }
            masks[s] &= mask;
        }
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

// Cooperative emission of one 1024-row word-span by one warp: entry i of the warp-private selection
// list (row index inside the span) goes to out[g0 + i].  Lanes take consecutive entries, so stores are
// coalesced and every lane carries four independent gathers.
template <typename T, bool FROM_SMEM>
__device__ __forceinline__ void emit_span(const unsigned short* sel_w, int n, int lane, uint32_t sbase, const T* __restrict__ gbase,
                                          T* __restrict__ out, long long g0, long long limit) {
    const long long room = limit - g0;
    if (room <= 0) return;
    if (room < (long long)n) n = (int)room;
    for (int i0 = 0; i0 < n; i0 += 128) {
        T v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * 32 + lane;
            if (i < n) {
                const uint32_t r = sel_w[i];
                v[k] = FROM_SMEM ? lds_cell<T>(sbase + r * (uint32_t)sizeof(T)) : __ldg(gbase + r);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = i0 + k * 32 + lane;
            if (i < n) out[g0 + i] = v[k];
        }
    }
}
// Any other cell width: byte-wise.
__device__ __forceinline__ void emit_span_bytes(const unsigned short* sel_w, int n, int lane, bool from_smem, uint32_t sbase,
                                                const uint8_t* __restrict__ gbase, uint8_t* __restrict__ out, int w, long long g0,
                                                long long limit) {
    const long long room = limit - g0;
    if (room <= 0) return;
    if (room < (long long)n) n = (int)room;
    for (int i = lane; i < n; i += 32) {
        const uint32_t r = sel_w[i];
        for (int b = 0; b < w; b++)
            out[(g0 + i) * w + b] = (uint8_t)(from_smem ? lds_u8(sbase + r * (uint32_t)w + b) : (uint32_t)__ldg(gbase + (long long)r * w + b));
    }
}

// All projected columns of one span (kept out of line: the kernel body stays small and lean on registers).
__device__ __noinline__ void emit_span_all(const ScanPlan& P, const unsigned short* sel_w, int n, int lane, bool staged,
                                           uint32_t stage_addr, int span_row, long long tile_row0, long long g0) {
    for (int pc = 0; pc < P.nproj; pc++) {
        const ProjCol& pj = P.proj[pc];
        const int w = pj.width;
        const bool from_smem = staged && pj.filter_idx >= 0;
        const uint32_t sbase = stage_addr + (from_smem ? (uint32_t)P.filter[pj.filter_idx].smem_off : 0u) + (uint32_t)(span_row * w);
        const uint8_t* gbase = pj.base + (tile_row0 + span_row) * w;
        if (w == 4) {
            if (from_smem) emit_span<uint32_t, true>(sel_w, n, lane, sbase, nullptr, (uint32_t*)pj.out, g0, P.limit);
            else emit_span<uint32_t, false>(sel_w, n, lane, 0u, (const uint32_t*)gbase, (uint32_t*)pj.out, g0, P.limit);
        } else if (w == 1) {
            if (from_smem) emit_span<uint8_t, true>(sel_w, n, lane, sbase, nullptr, pj.out, g0, P.limit);
            else emit_span<uint8_t, false>(sel_w, n, lane, 0u, gbase, pj.out, g0, P.limit);
        } else if (w == 2) {
            if (from_smem) emit_span<uint16_t, true>(sel_w, n, lane, sbase, nullptr, (uint16_t*)pj.out, g0, P.limit);
            else emit_span<uint16_t, false>(sel_w, n, lane, 0u, (const uint16_t*)gbase, (uint16_t*)pj.out, g0, P.limit);
        } else if (w == 8) {
            if (from_smem) emit_span<unsigned long long, true>(sel_w, n, lane, sbase, nullptr, (unsigned long long*)pj.out, g0, P.limit);
            else emit_span<unsigned long long, false>(sel_w, n, lane, 0u, (const unsigned long long*)gbase, (unsigned long long*)pj.out, g0, P.limit);
        } else {
            emit_span_bytes(sel_w, n, lane, from_smem, sbase, gbase, pj.out, w, g0, P.limit);
        }
    }
}

extern __shared__ __align__(128) uint8_t dyn_smem[];

template <int W>
__global__ void __launch_bounds__(kDenseThreads, IMM3_DENSE_MIN_BLOCKS) scan_dense_kernel(const __grid_constant__ ScanPlan P, ScanCtrl* ctrl,
                                                                     unsigned long long* status) {
    constexpr int kTile = kDenseTileRowsPerWord * W;  // rows per tile
    constexpr int kWarpSpan = 1024 * W;               // rows per compute warp
    __shared__ DenseShared S;
    // dynamic shared memory: [8 warp-private selection lists of 1024 uint16][TMA ring]
    const uint32_t ring_addr = smem_u32(dyn_smem) + kComputeWarps * 1024 * 2;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_ctrl = warp == kComputeWarps;
    const bool staged = P.stages > 0;
    const int ring = staged ? P.stages : 3;
    const unsigned ntiles = (unsigned)P.ntiles;

    if (tid == 0) {
        for (int s = 0; s < kMaxStages; s++) mbar_init(smem_u32(&S.mbar_full[s]), 1);
        for (int e = 0; e < 2; e++) {
            mbar_init(smem_u32(&S.mbar_count[e]), 1);
            mbar_init(smem_u32(&S.mbar_excl[e]), 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (is_ctrl) {
        // ---------------- control warp: counts in, exclusive offsets out ----------------
        for (unsigned j = 0;; j++) {
            const int e = (int)(j & 1u);
            mbar_wait(smem_u32(&S.mbar_count[e]), (j >> 1) & 1u, ctrl);
            const unsigned cnt = S.tile_count[e];
            if (cnt == kNoMoreTiles) break;
            const long long excl = resolve_tile(P, ctrl, status, S.tile_id[e], cnt, lane);
            if (lane == 0) {
                S.tile_excl[e] = excl;
                mbar_arrive(smem_u32(&S.mbar_excl[e]));
            }
            __syncwarp();
        }
    } else {
        // ---------------- compute warps ----------------
        // Hand out the next tile to ring slot `slot` and, if staging, start its bulk copies (thread 0).
        auto refill = [&](int slot) {
            unsigned t = kNoMoreTiles;
            if (!ld_relaxed_u32(&ctrl->done)) t = atomicAdd(&ctrl->ticket, 1u);  // after LIMIT: stop drawing tiles
            S.ticket[slot] = t;
            S.issued[slot] = 0;
            if (staged && t < ntiles) {
                const uint32_t bar = smem_u32(&S.mbar_full[slot]);
                mbar_arrive_expect_tx(bar, (uint32_t)P.stage_bytes);
                for (int i = 0; i < P.nfilter; i++) {
                    const FilterCol& f = P.filter[i];
                    const uint32_t bytes = (uint32_t)(kTile * f.width);
                    tma_load_1d(ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes + (uint32_t)f.smem_off,
                                f.base + (long long)t * bytes, bytes, bar);
                }
                S.issued[slot] = 1;
            }
        };
        if (tid == 0)
            for (int s = 0; s < ring; s++) refill(s);
        bar_sync(1, kComputeThreads);

        // Per-lane state of the tile being filtered (cur) and of the tile waiting to be emitted (prev): the
        // bitmap words and the tile-local rank of the warp's first selected row.  The in-warp ranks are
        // recomputed at emit time (a 5-step shuffle scan) rather than carried in registers.
        uint32_t m_cur[W], m_prev[W];
        unsigned wbase_cur = 0, wbase_prev = 0;  // tile-local rank of the warp's first selected row
        long long row0_cur = 0, row0_prev = 0;
        int slot_cur = 0, slot_prev = 0;
        int slotF = 0;
        uint32_t parF = 0;
        unsigned short* sel_w = reinterpret_cast<unsigned short*>(dyn_smem) + warp * 1024;

        // F(j): returns false when the CTA has run out of tiles.
        auto filter_tile = [&](unsigned j) -> bool {
            const int e = (int)(j & 1u);
            const int slot = slotF;
            const unsigned tile = S.ticket[slot];
            if (tile >= ntiles) {
                if (tid == 0) {
                    S.tile_count[e] = kNoMoreTiles;
                    mbar_arrive(smem_u32(&S.mbar_count[e]));
                }
                return false;
            }
            const long long tile_row0 = (long long)tile * kTile;
            const uint32_t stage_addr = ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes;
            if (S.issued[slot]) mbar_wait(smem_u32(&S.mbar_full[slot]), parF, ctrl);
            if (++slotF == ring) { slotF = 0; parF ^= 1u; }

            // decode + conjunctive filter: W bitmap words per lane, then ranks
            unsigned lane_total = 0;
#pragma unroll
            for (int s = 0; s < W; s++) {
                const long long left = P.nrows - (tile_row0 + warp * kWarpSpan + s * 1024 + lane * 32);
                m_cur[s] = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
            }
            for (int i = 0; i < P.nfilter; i++) dense_eval_filter<W>(P, P.filter[i], staged, stage_addr, tile_row0, warp * kWarpSpan, lane, m_cur);
#pragma unroll
            for (int s = 0; s < W; s++) {
                if (P.bitmap) P.bitmap[((tile_row0 + warp * kWarpSpan + s * 1024) >> 5) + lane] = m_cur[s];
                lane_total += __popc(m_cur[s]);
            }
            const unsigned warp_total = __reduce_add_sync(0xFFFFFFFFu, lane_total);
            if (lane == 0) S.warp_cnt[e][warp] = warp_total;
            bar_sync(1, kComputeThreads);
            unsigned warp_base = 0, tile_count = 0;
#pragma unroll
            for (int w = 0; w < kComputeWarps; w++) {
                const unsigned c = S.warp_cnt[e][w];
                if (w < warp) warp_base += c;
                tile_count += c;
            }
            if (tid == 0) {  // publish the count as early as possible: the control warp starts the look-back
                S.tile_count[e] = tile_count;
                S.tile_id[e] = tile;
                mbar_arrive(smem_u32(&S.mbar_count[e]));
            }
            wbase_cur = warp_base;
            row0_cur = tile_row0;
            slot_cur = slot;
            return true;
        };

        // E(j): emit the rows of the tile held in *_prev and recycle its ring slot.
        auto emit_tile = [&](unsigned j) {
            const int e = (int)(j & 1u);
            mbar_wait(smem_u32(&S.mbar_excl[e]), (j >> 1) & 1u, ctrl);
            const long long excl = S.tile_excl[e];
            const uint32_t stage_addr = ring_addr + (uint32_t)slot_prev * (uint32_t)P.stage_bytes;
            if (!P.bitmap && excl >= 0 && excl < P.limit) {
                unsigned wrank = wbase_prev;
#pragma unroll
                for (int s = 0; s < W; s++) {
                    uint32_t mm = m_prev[s];
                    const unsigned cnt = (unsigned)__popc(mm);
                    unsigned incl = cnt;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += nb;
                    }
                    const int n = (int)__shfl_sync(0xFFFFFFFFu, incl, 31);
                    const long long g0 = excl + wrank;
                    wrank += (unsigned)n;
                    if (n == 0) continue;  // warp-uniform
                    // warp-private selection list of this 1024-row span, from the register bitmap words
                    {
                        unsigned o = incl - cnt;
                        const unsigned iters = __reduce_max_sync(0xFFFFFFFFu, cnt);
                        for (unsigned it = 0; it < iters; it++) {
                            if (mm) {
                                sel_w[o++] = (unsigned short)(lane * 32 + __ffs(mm) - 1);
                                mm &= mm - 1u;
                            }
                        }
                    }
                    __syncwarp();
                    const int span_row = warp * kWarpSpan + s * 1024;
                    emit_span_all(P, sel_w, n, lane, staged, stage_addr, span_row, row0_prev, g0);
                    __syncwarp();  // the list is rebuilt for the next span
                }
            }
            bar_sync(1, kComputeThreads);  // every warp is done with the slot's bytes
            if (tid == 0) refill(slot_prev);
        };

        unsigned j = 0;
        bool more = filter_tile(0);
        while (more) {
#pragma unroll
            for (int s = 0; s < W; s++) m_prev[s] = m_cur[s];
            wbase_prev = wbase_cur;
            row0_prev = row0_cur;
            slot_prev = slot_cur;
            const bool next = filter_tile(j + 1);
            emit_tile(j);
            more = next;
            j++;
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
    unsigned int warp_cnt[2][kComputeWarps];
    unsigned long long scan_warp[kComputeWarps];
    unsigned int is_last;
};

// Exclusive scan of the tile counts by one CTA of kComputeThreads threads (each thread owns a contiguous
// chunk: two passes over L2-resident counts, one block-wide scan), LIMIT clamp of the total.
__device__ void scan_tile_counts(FilterShared& S, const uint32_t* tile_cnt, unsigned long long* tile_off, long long ntiles,
                                 long long limit, ScanCtrl* ctrl) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long chunk = (ntiles + kComputeThreads - 1) / kComputeThreads;
    const long long i0 = (long long)tid * chunk, i1 = i0 + chunk < ntiles ? i0 + chunk : ntiles;
    unsigned long long run = 0;
    for (long long i = i0; i < i1; i++) run += __ldcg(tile_cnt + i);
    unsigned long long incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += nb;
    }
    if (lane == 31) S.scan_warp[warp] = incl;
    __syncthreads();
    unsigned long long excl = incl - run, total = 0;
#pragma unroll
    for (int w = 0; w < kComputeWarps; w++) {
        const unsigned long long ws = S.scan_warp[w];
        if (w < warp) excl += ws;
        total += ws;
    }
    for (long long i = i0; i < i1; i++) {
        tile_off[i] = excl;
        excl += __ldcg(tile_cnt + i);
    }
    if (tid == 0) {
        tile_off[ntiles] = total;
        ctrl->total = total < (unsigned long long)limit ? total : (unsigned long long)limit;
    }
}

template <int W>
__global__ void __launch_bounds__(kComputeThreads) filter_kernel(const __grid_constant__ ScanPlan P, uint32_t* __restrict__ bitmap,
                                                                   uint32_t* __restrict__ span_cnt, uint32_t* __restrict__ tile_cnt,
                                                                   unsigned long long* __restrict__ tile_off, ScanCtrl* ctrl) {
    constexpr int kTile = kDenseTileRowsPerWord * W;
    constexpr int kWarpSpan = 1024 * W;
    __shared__ FilterShared S;
    const uint32_t ring_addr = smem_u32(dyn_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool staged = P.stages > 0;
    const int ring = staged ? P.stages : 1;
    const long long ntiles = P.ntiles;

    auto issue = [&](long long tile, int slot) {  // thread 0
        const uint32_t bar = smem_u32(&S.mbar_full[slot]);
        mbar_arrive_expect_tx(bar, (uint32_t)P.stage_bytes);
        for (int i = 0; i < P.nfilter; i++) {
            const FilterCol& f = P.filter[i];
            const uint32_t bytes = (uint32_t)(kTile * f.width);
            tma_load_1d(ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes + (uint32_t)f.smem_off, f.base + tile * bytes, bytes, bar);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < kMaxFilterStages; s++) mbar_init(smem_u32(&S.mbar_full[s]), 1);
        fence_mbar_init();
        if (staged)
            for (int s = 0; s < ring; s++) {
                const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
                if (t < ntiles) issue(t, s);
            }
    }
    __syncthreads();

    int slot = 0;
    uint32_t parity = 0;
    unsigned it = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const long long tile_row0 = tile * kTile;
        const uint32_t stage_addr = ring_addr + (uint32_t)slot * (uint32_t)P.stage_bytes;
        if (staged) mbar_wait(smem_u32(&S.mbar_full[slot]), parity, nullptr);
        uint32_t m[W];
#pragma unroll
        for (int s = 0; s < W; s++) {
            const long long left = P.nrows - (tile_row0 + warp * kWarpSpan + s * 1024 + lane * 32);
            m[s] = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << (int)left) - 1u));
        }
        for (int i = 0; i < P.nfilter; i++) dense_eval_filter<W>(P, P.filter[i], staged, stage_addr, tile_row0, warp * kWarpSpan, lane, m);
        unsigned warp_total = 0;
#pragma unroll
        for (int s = 0; s < W; s++) {
            const long long span_row0 = tile_row0 + warp * kWarpSpan + s * 1024;
            bitmap[(span_row0 >> 5) + lane] = m[s];
            const unsigned c = __reduce_add_sync(0xFFFFFFFFu, (unsigned)__popc(m[s]));
            if (lane == 0) span_cnt[span_row0 >> 10] = c;
            warp_total += c;
        }
        const int e = (int)(it & 1u);
        if (lane == 0) S.warp_cnt[e][warp] = warp_total;
        __syncthreads();  // every warp has consumed the slot's bytes
        if (tid == 0) {
            unsigned c = 0;
#pragma unroll
            for (int w = 0; w < kComputeWarps; w++) c += S.warp_cnt[e][w];
            tile_cnt[tile] = c;
            const long long next = tile + (long long)ring * gridDim.x;
            if (staged && next < ntiles) issue(next, slot);
        }
        if (++slot == ring) { slot = 0; parity ^= 1u; }
    }

    // The last CTA to finish turns the tile counts into device-wide offsets (saves a launch).
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

// A span whose 1024 rows all survive: straight coalesced copy, no selection vector.
__device__ __noinline__ void emit_span_full(const ScanPlan& P, int lane, long long row0, long long g0) {
    long long n = P.limit - g0;
    if (n <= 0) return;
    if (n > 1024) n = 1024;
    for (int pc = 0; pc < P.nproj; pc++) {
        const ProjCol& pj = P.proj[pc];
        const int w = pj.width;
        if (w == 4) {
            const uint32_t* src = reinterpret_cast<const uint32_t*>(pj.base) + row0;
            uint32_t* dst = reinterpret_cast<uint32_t*>(pj.out) + g0;
            for (int i = lane; i < n; i += 32) dst[i] = __ldg(src + i);
        } else {
            const uint8_t* src = pj.base + row0 * w;
            uint8_t* dst = pj.out + g0 * w;
            for (int i = lane; i < (int)n * w; i += 32) dst[i] = __ldg(src + i);
        }
    }
}

// K3.  Two mappings, chosen on the device from the total match count:
//  * dense results (>= 32 surviving rows per span on average): one warp per 1024-row span - maximum
//    parallelism, every span's three metadata loads issued together;
//  * sparse results: one warp per group of 8 spans - the group's bitmap words are fetched up front, the
//    surviving rows of consecutive spans are appended to ONE warp-private selection vector and emitted
//    together, so a group pays one gather latency instead of eight.
__global__ void __launch_bounds__(kComputeThreads, IMM3_EMIT_MIN_BLOCKS) emit_kernel(const __grid_constant__ ScanPlan P, const uint32_t* __restrict__ bitmap,
                                                                 const uint32_t* __restrict__ span_cnt,
                                                                 const unsigned long long* __restrict__ tile_off, int spans_per_tile,
                                                                 long long nspans) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned short* sel_w = reinterpret_cast<unsigned short*>(dyn_smem) + warp * 1024;
    const long long ntiles = (nspans + spans_per_tile - 1) / spans_per_tile;
    const unsigned long long total = __ldg(tile_off + ntiles);
    const long long warp0 = (long long)blockIdx.x * kComputeWarps + warp, nwarps = (long long)gridDim.x * kComputeWarps;

    if (total >= (unsigned long long)nspans * 32ull) {
        // ---------------- one warp per span ----------------
        for (long long p = warp0; p < nspans; p += nwarps) {
            const long long t = p / spans_per_tile;
            const int k = (int)(p - t * spans_per_tile);
            // three independent loads: the tile's span counts (lane i holds span i), its offset, my bitmap word
            const unsigned c = lane < spans_per_tile && t * spans_per_tile + lane < nspans ? __ldg(span_cnt + t * spans_per_tile + lane) : 0u;
            const unsigned long long toff = __ldg(tile_off + t);
            uint32_t mm = __ldg(bitmap + p * 32 + lane);
            const int n = (int)__shfl_sync(0xFFFFFFFFu, c, k);
            if (n == 0) continue;
            const long long g0 = (long long)toff + __reduce_add_sync(0xFFFFFFFFu, lane < k ? c : 0u);
            if (g0 >= P.limit) continue;
            if (n == 1024) {
                emit_span_full(P, lane, p * 1024, g0);
                continue;
            }
            const unsigned cnt = (unsigned)__popc(mm);
            unsigned incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += nb;
            }
            unsigned o = incl - cnt;
            const unsigned iters = __reduce_max_sync(0xFFFFFFFFu, cnt);
            for (unsigned it = 0; it < iters; it++) {
                if (mm) {
                    sel_w[o++] = (unsigned short)(lane * 32 + __ffs(mm) - 1);
                    mm &= mm - 1u;
                }
            }
            __syncwarp();
            emit_span_all(P, sel_w, n, lane, false, 0u, 0, p * 1024, g0);
            __syncwarp();
        }
        return;
    }

    // ---------------- one warp per group of 8 spans ----------------
    const long long ngroups = (nspans + 7) >> 3;
    for (long long u = warp0; u < ngroups; u += nwarps) {
        const long long p0 = u * 8;                   // first span of the group
        const long long t = p0 / spans_per_tile;      // groups never straddle tiles (spans_per_tile is 8, 16 or 32)
        const int k0 = (int)(p0 - t * spans_per_tile);
        const unsigned c = lane < spans_per_tile && t * spans_per_tile + lane < nspans ? __ldg(span_cnt + t * spans_per_tile + lane) : 0u;
        const unsigned long long toff = __ldg(tile_off + t);
        const unsigned in_group = __reduce_add_sync(0xFFFFFFFFu, (lane >= k0 && lane < k0 + 8) ? c : 0u);
        if (in_group == 0) continue;
        long long g0 = (long long)toff + __reduce_add_sync(0xFFFFFFFFu, lane < k0 ? c : 0u);  // ordinal of the group's first surviving row
        if (g0 >= P.limit) continue;
        uint32_t mw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned ck = __shfl_sync(0xFFFFFFFFu, c, k0 + k);
            mw[k] = (ck && ck != 1024u) ? __ldg(bitmap + (p0 + k) * 32 + lane) : 0u;
        }
        int fill = 0;          // rows in the selection vector, first of them is global ordinal g0
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int n = (int)__shfl_sync(0xFFFFFFFFu, c, k0 + k);
            if (n == 0) continue;
            if (n == 1024 || fill + n > 1024) {  // flush what has been gathered so far
                __syncwarp();
                if (fill) emit_span_all(P, sel_w, fill, lane, false, 0u, 0, p0 * 1024, g0);
                __syncwarp();
                g0 += fill;
                fill = 0;
            }
            if (n == 1024) {
                emit_span_full(P, lane, (p0 + k) * 1024, g0);
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
        if (fill) emit_span_all(P, sel_w, fill, lane, false, 0u, 0, p0 * 1024, g0);
        __syncwarp();
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
        cudaError_t e = cudaFuncSetAttribute(scan_dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(scan_dense_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(scan_dense_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(filter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(filter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(filter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(scan_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }();
    return rc;
}

cudaError_t dense_kernel_occupancy(int words_per_lane, size_t dyn_smem, int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (words_per_lane == 4) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, scan_dense_kernel<4>, kDenseThreads, dyn_smem);
    if (words_per_lane == 2) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, scan_dense_kernel<2>, kDenseThreads, dyn_smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, scan_dense_kernel<1>, kDenseThreads, dyn_smem);
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
    if (plan.words_per_lane == 4) scan_dense_kernel<4><<<grid, kDenseThreads, dyn_smem, stream>>>(plan, ctrl, status);
    else if (plan.words_per_lane == 2) scan_dense_kernel<2><<<grid, kDenseThreads, dyn_smem, stream>>>(plan, ctrl, status);
    else scan_dense_kernel<1><<<grid, kDenseThreads, dyn_smem, stream>>>(plan, ctrl, status);
    return cudaGetLastError();
}
cudaError_t filter_kernel_occupancy(int words_per_lane, size_t dyn_smem, int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (words_per_lane == 4) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, filter_kernel<4>, kComputeThreads, dyn_smem);
    if (words_per_lane == 2) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, filter_kernel<2>, kComputeThreads, dyn_smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, filter_kernel<1>, kComputeThreads, dyn_smem);
}
cudaError_t emit_kernel_occupancy(int* blocks_per_sm) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, emit_kernel, kComputeThreads, kComputeWarps * 1024 * 2);
}
cudaError_t launch_filter(const ScanPlan& plan, uint32_t* bitmap, uint32_t* span_cnt, uint32_t* tile_cnt, unsigned long long* tile_off,
                          ScanCtrl* ctrl, int grid, size_t dyn_smem, cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    if (plan.words_per_lane == 4) filter_kernel<4><<<grid, kComputeThreads, dyn_smem, stream>>>(plan, bitmap, span_cnt, tile_cnt, tile_off, ctrl);
    else if (plan.words_per_lane == 2) filter_kernel<2><<<grid, kComputeThreads, dyn_smem, stream>>>(plan, bitmap, span_cnt, tile_cnt, tile_off, ctrl);
    else filter_kernel<1><<<grid, kComputeThreads, dyn_smem, stream>>>(plan, bitmap, span_cnt, tile_cnt, tile_off, ctrl);
    return cudaGetLastError();
}
cudaError_t launch_emit(const ScanPlan& plan, const uint32_t* bitmap, const uint32_t* span_cnt, const unsigned long long* tile_off,
                        int spans_per_tile, long long nspans, int grid, cudaStream_t stream) {
    cudaError_t e = configure_once();
    if (e != cudaSuccess) return e;
    emit_kernel<<<grid, kComputeThreads, kComputeWarps * 1024 * 2, stream>>>(plan, bitmap, span_cnt, tile_off, spans_per_tile, nspans);
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

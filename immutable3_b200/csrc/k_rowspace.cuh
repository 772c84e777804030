// k_rowspace.cuh - what every row-space kernel shares: launch constants, the plan tables in shared memory, SIMD-within-a-register predicates, selection vectors, gathers
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

#ifndef IMM3_EMIT_MIN_BLOCKS
#define IMM3_EMIT_MIN_BLOCKS 4
#endif
constexpr int kComputeThreads = 256;
constexpr int kComputeWarps = kComputeThreads / 32;
constexpr unsigned kNoMoreTiles = 0xFFFFFFFFu;
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


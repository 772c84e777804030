// k_blocks_single.cuh - single-pass block kernel with decoupled look-back (scan_blocks_kernel): blocks of more than 1024 rows, imm3_filter_bitmap on block tables
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

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


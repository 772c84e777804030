// k_blocks_scanemit.cuh - offset scan + emit of the block pipeline in ONE small-footprint kernel (single encoded column projected)
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// blocks_scan_emit_kernel (round 2).  `select id from t where id in a window` (C4) ran filter -> offset_scan_kernel ->
// blocks_emit_kernel; traced with %globaltimer, the emit kernel's first CTA started 7 us after the scan kernel's last one had
// left: it needs 90 KB of shared memory per CTA (decode scratch for any select list), so - although a programmatic
// dependent - it cannot become resident next to the filter kernel's 225 KB rings, and the whole launch latency sits
// between the kernels.  This kernel needs 64 registers and 3 KB of shared memory per CTA: its CTAs are resident, set up and
// parked in griddepcontrol.wait while the filter kernel is still running.
//
//   1. CTAs 0 .. nchunks-1 run one chunk of the offset scan each (offset_scan_chunk, the same code as offset_scan_kernel; the
//      chunk CTAs have the lowest block indices, so a chunk only ever waits for chunks that are running) and count
//      themselves off; every CTA waits for that count.
//   2. Emit, a warp per block with surviving rows, found through the list of non-empty tiles the scan leaves: entry e =
//      (list position e / 8, block e % 8 of that tile), a lane per entry, so one round trip brings the metadata of up to 32
//      blocks; then the blocks one after the other, the next one's encoded words (dense_issue) in flight meanwhile.
//        dense shape (k_blocks_multi.cuh: emit_dense_block), every row selected -> dense_finish
//        dense shape, some rows selected (the blocks a window edge cuts)       -> dense_finish_sel: same, rows compacted
//        any other shape (the 1-row block that ends a segment, unsorted data)   -> emit_any_block: lane = mini-block, straight
//                                                                                  from global memory, two passes (totals, values)
// The general blocks_emit_kernel keeps every other select list (dense columns, several columns, row-space bitmaps).
// =============================================================================================
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ldg_be(const uint32_t* __restrict__ W, int i) { return __byte_perm(__ldg(W + i), 0, 0x0123); }  // (putInt is big-endian)

// Dense-shape block of which SOME rows are selected: S = lane m's selection word (rows 32 m ..), out = the block's first
// result row.  Returns false (nothing written) if the block does not have the dense shape.  scratch: 128 warp-private words.
__device__ __forceinline__ bool dense_finish_sel(const DenseRegs& r, uint32_t S, uint32_t* __restrict__ out, int nn, int lane, uint32_t* scratch) {
    if (r.B < 0) return false;
    const int B = r.B;
    const uint32_t off = (uint32_t)(lane * B);
    const uint32_t hexp = lane == 0 ? (0x01010100u | (uint32_t)B) : 0x01010101u;
    if (!__all_sync(0xFFFFFFFFu, lane >= 8 || r.hraw == hexp)) return false;
    uint32_t v = __funnelshift_r(__byte_perm(r.x0, 0, 0x0123), __byte_perm(r.x1, 0, 0x0123), off) & ((1u << B) - 1u);  // field `lane`
    const uint32_t X = __byte_perm(r.nraw, 0, 0x0123);
    const uint32_t wide_total = __reduce_add_sync(0xFFFFFFFFu, v);
    const uint32_t tot = lane == 0 ? wide_total : (uint32_t)__popc(X);
    const uint32_t ns = (uint32_t)__popc(S);
    uint32_t incl = tot, pre = ns;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {  // three independent scans: the wide mini-block's values, the carries, the selected rows before a mini-block
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o), u = __shfl_up_sync(0xFFFFFFFFu, incl, o), w = __shfl_up_sync(0xFFFFFFFFu, pre, o);
        if (lane >= o) {
            v += t;
            incl += u;
            pre += w;
        }
    }
    uint2* const pair = reinterpret_cast<uint2*>(scratch);
    pair[lane] = make_uint2(X, incl - tot);
    pair[32 + lane] = make_uint2(S, pre - ns);
    __syncwarp();
    const uint32_t low = (2u << lane) - 1u, below = (1u << lane) - 1u;
    {
        const uint2 sp = pair[32];
        const uint32_t rk = (uint32_t)__popc(sp.x & below);
        if (((sp.x >> lane) & 1u) && rk < (uint32_t)nn) out[rk] = v;
    }
#pragma unroll 4
    for (int m = 1; m < 32; m++) {
        const uint2 xc = pair[m], sp = pair[32 + m];
        const uint32_t rk = sp.y + (uint32_t)__popc(sp.x & below);
        if (((sp.x >> lane) & 1u) && rk < (uint32_t)nn) out[rk] = xc.y + (uint32_t)__popc(xc.x & low);
    }
    __syncwarp();
    return true;
}

// A block of ANY shape (n <= 1024 rows), decoded by one warp straight from global memory - lane = mini-block, no shared
// memory: the header walk, then every lane sums its 32 deltas (pass 1), a segmented warp scan chains the mini-blocks (raw
// b = 32 mini-blocks restart the chain), and every lane walks its deltas again and stores the selected values at their ranks
// (pass 2).  W: the block's words (count word first), S: lane w's selection word (rows 32 w ..).
__device__ __noinline__ void emit_any_block(const uint32_t* __restrict__ W, int n, uint32_t S, uint32_t* __restrict__ out, int nn, int lane) {
    const int packed = n & ~31, nmini = packed >> 5, nsuper = packed >> 7;
    const int q = lane & 3, k = lane >> 2;
    const uint32_t before = q == 0 ? 0u : (0x01010100u << (8 * (3 - q)));  // selects the widths of the mini-blocks ahead of q
    int ip = 1, mypos = 0, mybits = 0;
    {
        uint32_t myh = 0;
#pragma unroll 1
        for (int s = 0; s < nsuper; s++) {
            const uint32_t h = ldg_be(W, ip);
            const int pos = ip + 1 + (int)__dp4a(h, before, 0u);
            mypos = k == s ? pos : mypos;
            myh = k == s ? h : myh;
            ip += 1 + (int)__dp4a(h, 0x01010101u, 0u);
        }
        mybits = (int)((myh >> (24 - 8 * q)) & 0xFFu);
#pragma unroll 1
        for (int m = nsuper * 4; m < nmini; m++) {  // left-over mini-blocks carry their own header word
            const int b = (int)ldg_be(W, ip++);
            if (m == lane) { mypos = ip; mybits = b; }
            ip += b;
        }
    }
    const bool active = lane < nmini, raw = active && mybits >= 32;
    const uint32_t mask = mybits >= 32 ? 0xFFFFFFFFu : ((1u << mybits) - 1u);
    uint32_t total = 0;
    if (active) {
        if (raw) {
            total = ldg_be(W, mypos + 31);
        } else if (mybits > 0) {
            uint32_t off = 0;
#pragma unroll 4
            for (int j = 0; j < 32; j++, off += (uint32_t)mybits) {
                const int wi = mypos + (int)(off >> 5);
                total += __funnelshift_r(ldg_be(W, wi), ldg_be(W, wi + 1), off) & mask;  // (the word behind the block's last one exists: 8 pad bytes)
            }
        }
    }
    uint32_t v = active ? total : 0u;
    unsigned f = raw ? 1u : 0u;
    uint32_t pre = (uint32_t)__popc(S);
    const uint32_t ns = pre;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t pv = __shfl_up_sync(0xFFFFFFFFu, v, o), pp = __shfl_up_sync(0xFFFFFFFFu, pre, o);
        const unsigned pf = __shfl_up_sync(0xFFFFFFFFu, f, o);
        if (lane >= o) {
            if (!f) v += pv;
            f |= pf;
            pre += pp;
        }
    }
    uint32_t base = __shfl_up_sync(0xFFFFFFFFu, v, 1);  // the value before my mini-block
    if (lane == 0) base = 0;                             // initvalue = 0 at every block
    uint32_t carry_tail = __shfl_sync(0xFFFFFFFFu, v, (nmini + 31) & 31);
    if (nmini == 0) carry_tail = 0;
    pre -= ns;  // selected rows before my 32
    if (active) {
        uint32_t run = base, off = 0;
#pragma unroll 4
        for (int j = 0; j < 32; j++, off += (uint32_t)mybits) {
            uint32_t val;
            if (raw) {
                val = ldg_be(W, mypos + j);
            } else {
                if (mybits > 0) {
                    const int wi = mypos + (int)(off >> 5);
                    run += __funnelshift_r(ldg_be(W, wi), ldg_be(W, wi + 1), off) & mask;
                }
                val = run;
            }
            if ((S >> j) & 1u) {
                const uint32_t rk = pre + (uint32_t)__popc(S & ((1u << j) - 1u));
                if (rk < (uint32_t)nn) out[rk] = val;
            }
        }
    }
    // var-byte remainder (n % 32 values): 7-bit groups, low first, the last byte of a value has bit 7 set; selection word nmini
    if (n > packed && lane == nmini) {
        int wpos = ip, shb = 0, shift = 0;
        uint32_t acc = 0, cur = carry_tail;
        for (int i = 0; i < n - packed;) {
            const uint32_t c = ldg_be(W, wpos) >> shb;
            shb += 8;
            wpos += shb >> 5;
            shb &= 31;
            acc += (c & 127u) << shift;
            if (c & 128u) {
                cur += acc;
                if ((S >> i) & 1u) {
                    const uint32_t rk = pre + (uint32_t)__popc(S & ((1u << i) - 1u));
                    if (rk < (uint32_t)nn) out[rk] = cur;
                }
                i++;
                acc = 0;
                shift = 0;
            } else {
                shift += 7;
            }
        }
    }
}

// The emit phase: warps warp0, warp0 + nwarps, ... of `nwarps` take the entries of the non-empty-tile list.  scratch: 128
// warp-private words of shared memory.  (Offsets, total and list were written earlier in the SAME kernel, by the
// chunk CTAs: they are read through L2, never through the non-coherent path.)
__device__ __forceinline__ void lean_emit(const LeanPlan& P, const uint32_t* __restrict__ bitmapB, const uint32_t* __restrict__ blk_cnt,
                                          const unsigned long long* __restrict__ tile_off, long long nblocks, const ScanCtrl* ctrl,
                                          const unsigned int* __restrict__ tile_list, uint32_t* scratch, long long warp0, long long nwarps, int lane) {
    struct { const uint32_t* words; const uint32_t* word_off; } pc = {P.words, P.word_off};
    uint32_t* const outc = reinterpret_cast<uint32_t*>(P.out);
    const long long nent = (long long)__ldcg(&ctrl->ticket2) * 8;
#pragma unroll 1
    for (long long e0 = warp0; e0 < nent; e0 += 32 * nwarps) {
        // ---- my entry's block: metadata in one round trip ----
        const long long e = e0 + lane * nwarps;
        long long b = 0;
        bool cand = e < nent;
        if (cand) {
            const unsigned t8 = __ldcg(tile_list + (e >> 3));
            IMM3_CHECK(ctrl, (long long)t8 < P.ntiles, 4);  // a list entry is a tile of this table
            b = (long long)t8 * 8 + (e & 7);
            cand = b < nblocks;
        }
        unsigned long long r0 = 0, r1 = 0, g = 0;
        uint32_t w0 = 0, w1 = 0, mycnt = 0;
        if (cand) {
            r0 = P.row_start[b];
            r1 = P.row_start[b + 1];
            w0 = __ldg(pc.word_off + b);
            w1 = __ldg(pc.word_off + b + 1);
            const uint4* tc = reinterpret_cast<const uint4*>(blk_cnt + (b & ~7ll));
            const uint4 ca = __ldcg(tc), cb = __ldcg(tc + 1);
            g = __ldcg(tile_off + (b >> 3));
            const int j = (int)(b & 7);
            const uint32_t c8[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
                g += i < j ? c8[i] : 0u;
                mycnt = i == j ? c8[i] : mycnt;
            }
        }
        const int myn = (int)(r1 - r0);
        unsigned todo = __ballot_sync(0xFFFFFFFFu, cand && mycnt != 0u && (long long)g < P.limit);
        if (!todo) continue;
        // ---- the blocks, one after the other; the next one's words are in flight ----
        int nsrc = __ffs((int)todo) - 1;
        DenseRegs nx = dense_issue(pc.words, __shfl_sync(0xFFFFFFFFu, w0, nsrc), __shfl_sync(0xFFFFFFFFu, w1, nsrc), __shfl_sync(0xFFFFFFFFu, myn, nsrc), lane);
#pragma unroll 1
        while (todo) {
            const int src = nsrc;
            todo &= todo - 1u;
            const DenseRegs cur = nx;
            if (todo) {
                nsrc = __ffs((int)todo) - 1;
                nx = dense_issue(pc.words, __shfl_sync(0xFFFFFFFFu, w0, nsrc), __shfl_sync(0xFFFFFFFFu, w1, nsrc), __shfl_sync(0xFFFFFFFFu, myn, nsrc), lane);
            }
            const int n = __shfl_sync(0xFFFFFFFFu, myn, src);
            const unsigned cnt = __shfl_sync(0xFFFFFFFFu, mycnt, src);
            const long long gb = (long long)__shfl_sync(0xFFFFFFFFu, g, src);
            const int nn = (int)(P.limit - gb < (long long)cnt ? P.limit - gb : (long long)cnt);
            uint32_t* const o = outc + gb;
            IMM3_CHECK(ctrl, nn >= 0 && (unsigned long long)(gb + nn) <= __ldcg(&ctrl->total) && n > 0 && n <= 1024, 5);  // the block's rows fit the result
            // the filter kernel stores the 32 words of a block only if SOME of its rows survive; all of them: the count says so
            uint32_t S;
            if (cnt == (unsigned)n) {
                const int left = n - lane * 32;
                S = left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
            } else {
                S = __ldcg(bitmapB + __shfl_sync(0xFFFFFFFFu, b, src) * 32 + lane);
            }
            __syncwarp();
            bool done = false;
            if (cur.B >= 0) done = cnt == (unsigned)n ? dense_finish(cur, o, nn, lane, scratch) : dense_finish_sel(cur, S, o, nn, lane, scratch);
            if (!done) emit_any_block(pc.words + __shfl_sync(0xFFFFFFFFu, w0, src), n, S, o, nn, lane);
        }
    }
}

__global__ void __launch_bounds__(kComputeThreads, 4) blocks_scan_emit_kernel(const __grid_constant__ LeanPlan P, const uint32_t* __restrict__ bitmapB,
                                                                                const uint32_t* __restrict__ blk_cnt, const uint32_t* __restrict__ tile_cnt,
                                                                                unsigned long long* __restrict__ tile_off, long long nblocks, uint32_t epoch,
                                                                                unsigned long long* __restrict__ partials, ScanCtrl* ctrl,
                                                                                unsigned int* __restrict__ tile_list, CtrlBlock* pub, unsigned long long pub_seq) {
    __shared__ ScanShared SS;
    __shared__ uint32_t s_scratch[kComputeWarps][128];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (the count exchange of a sharded table rides behind)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long ntiles8 = P.ntiles;
    const unsigned nchunks = (unsigned)((ntiles8 + kComputeThreads * 16 - 1) / (kComputeThreads * 16));
    // chunks done: a counter on a cache line of its own behind the chunk sums (hundreds of CTAs poll it while the chunk CTAs
    // work on the control block's line)
    unsigned int* const done_ctr = reinterpret_cast<unsigned int*>(partials + 2 * (size_t)nchunks + 16);
    if (lane == 0) phase_stamp(P, 8);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the filter kernel's counts are final
    if (lane == 0) phase_stamp(P, 9);
    // ---------------- 1. the offset scan ----------------
    if (blockIdx.x < nchunks) {
        offset_scan_chunk(SS, (long long)blockIdx.x, tile_cnt, tile_off, ntiles8, P.limit, epoch, partials, ctrl, tile_list);
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(done_ctr, 1u);
        }
    }
    if (tid == 0) {
        uint64_t t0 = 0;
        unsigned spins = 0;
        while (ld_acquire_u32(done_ctr) < nchunks) {
            __nanosleep(400);
            if ((++spins & 255u) == 0) {
                const uint64_t now = globaltimer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kWatchdogNs) watchdog_trap(ctrl, 4);
            }
        }
    }
    __syncthreads();
    if (lane == 0) phase_stamp(P, 10);
    // ---------------- 2. emit ----------------
    if (__ldcg(&ctrl->total) != 0ull) {
        lean_emit(P, bitmapB, blk_cnt, tile_off, nblocks, ctrl, tile_list, s_scratch[warp], (long long)blockIdx.x * kComputeWarps + warp,
                  (long long)gridDim.x * kComputeWarps, lane);
    }
    if (lane == 0) phase_stamp(P, 14);
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&ctrl->exited, 1u);
        if (done == gridDim.x - 1) {  // last CTA out: the counters are the next query's again
            ctrl->exited = 0;
            ctrl->ticket = 0;
            *done_ctr = 0;
            if (pub) {  // publish (plan.hpp): no kernel follows - the host is polling its pinned copy of the control block.
                // ONE posted 8-byte store carries everything it needs - [63:41] sequence, [40] watchdog fired, [39:0] rows - so the
                // kernel does not have to wait for a system-scope fence before it ends
                const unsigned long long word = ((pub_seq & 0x7FFFFFull) << 41) | (__ldcg(&ctrl->error) ? (1ull << 40) : 0ull) |
                                                (__ldcg(&ctrl->total) & ((1ull << 40) - 1ull));
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&pub->pub_seq), "l"(word) : "memory");
            }
        }
    }
}

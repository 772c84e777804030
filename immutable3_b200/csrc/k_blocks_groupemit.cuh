// k_blocks_groupemit.cuh - emit of the block pipeline WITHOUT an offset scan phase (single encoded column projected)
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

// =============================================================================================
// blocks_group_emit_kernel (round 2, after blocks_scan_emit_kernel).  Traced with %globaltimer on C4 at 1 B rows, the
// scan-emit kernel's CTAs were resident and parked in griddepcontrol.wait 8 us into the filter kernel, saw its counts 1.9 us
// after its last CTA - and then spent 12 us in the chunked offset scan (counts -> chunk sums published -> earlier chunks'
// sums awaited -> offsets + tile list written -> a device-wide count of finished chunks polled) before the first row moved.
//
// Here nothing is scanned device-wide and nobody waits for anybody:
//   * the filter (and prune) kernel adds every 32-block tile's match count to the sum of its GROUP (1024 blocks) with one
//     fire-and-forget RED - only for tiles that have matches;
//   * every CTA of this kernel reads all group sums (4 KB at 1 B rows) and scans them itself: total T, and the result is cut
//     into one contiguous range of rows per CTA, [c T / C, (c + 1) T / C): a block belongs to the CTA whose range holds its
//     first result row;
//   * a CTA walks the groups its range touches: the group's 1024 block counts (one coalesced 4 KB read, four per thread)
//     scanned by the CTA give every block's first result row; its blocks go on a queue in shared memory, the warps share
//     them out (a lane per block fetches the metadata in one round trip) and emit them one after the other with the next
//     one's encoded words in flight (the per-block routines of blocks_scan_emit_kernel: dense_finish / dense_finish_sel /
//     emit_any_block).  (A first version did all of this per warp: 3 500 warps reading the same few 4 KB of counts cost more
//     in L2 than the scan had.)
// Critical path behind the filter kernel: group sums -> block counts -> block metadata -> encoded words -> stores: four
// round trips instead of the scan's nine or so, and no tile offsets, tile list or chunk sums are written at all.
// The last CTA out clears the group sums for the next query, writes the total and publishes it (plan.hpp: CtrlBlock).
// Tables with more than kGrpMaxGroups groups (4 M blocks) keep blocks_scan_emit_kernel.
// =============================================================================================
constexpr int kGrpShift = 10;          // a group = 1024 blocks = 32 tiles of the lane kernel
constexpr int kGrpMaxGroups = 4096;    // 16 group sums per thread of a 256-thread CTA
constexpr int kGrpSmemGroups = 1024;   // group sums kept in shared memory (1 M blocks = 1 B rows in blocks of 1024)
constexpr int kGrpSumWords = kGrpMaxGroups + 64;  // (whole 16-byte loads and 32-lane look-aheads past the last group read zeros)

struct GroupEmitShared {
    unsigned long long warp_sum[kComputeWarps];
    unsigned long long start_base;
    unsigned int start_grp;
    unsigned int blk_sum[kComputeWarps];   // per warp: rows of its 128 blocks of the current group
    unsigned int mine_cnt[kComputeWarps];  // per warp: how many of them are this CTA's
    unsigned int is_last;
    __align__(16) unsigned int sums[kGrpSmemGroups];  // the group sums (tables of up to kGrpSmemGroups groups: no second trip to L2 for them)
};

// Emit the queued blocks of one warp: lane q < qn holds entry q = (block, its first result row).
__device__ __forceinline__ void emit_queued_blocks(const LeanPlan& P, const uint32_t* __restrict__ bitmapB, const uint32_t* __restrict__ blk_cnt,
                                                   const ScanCtrl* ctrl, int qn, long long b, unsigned long long g, unsigned long long total,
                                                   uint32_t* scratch, int lane, bool stamp) {
    const bool cand = lane < qn;
    unsigned long long r0 = 0, r1 = 0;
    uint32_t w0 = 0, w1 = 0, mycnt = 0;
    if (cand) {  // the metadata of up to 32 blocks in one round trip
        r0 = P.row_start[b];
        r1 = P.row_start[b + 1];
        w0 = __ldg(P.word_off + b);
        w1 = __ldg(P.word_off + b + 1);
        mycnt = __ldcg(blk_cnt + b);
    }
    const int myn = (int)(r1 - r0);
    uint32_t* const outc = reinterpret_cast<uint32_t*>(P.out);
    unsigned todo = __ballot_sync(0xFFFFFFFFu, cand);
    if (!todo) return;
    int nsrc = __ffs((int)todo) - 1;
    DenseRegs nx = dense_issue(P.words, __shfl_sync(0xFFFFFFFFu, w0, nsrc), __shfl_sync(0xFFFFFFFFu, w1, nsrc), __shfl_sync(0xFFFFFFFFu, myn, nsrc), lane);
    if (stamp && lane == 0) phase_stamp(P, 12);  // (the metadata has arrived)
#pragma unroll 1
    while (todo) {
        const int src = nsrc;
        todo &= todo - 1u;
        const DenseRegs cur = nx;
        if (todo) {
            nsrc = __ffs((int)todo) - 1;
            nx = dense_issue(P.words, __shfl_sync(0xFFFFFFFFu, w0, nsrc), __shfl_sync(0xFFFFFFFFu, w1, nsrc), __shfl_sync(0xFFFFFFFFu, myn, nsrc), lane);
        }
        const int n = __shfl_sync(0xFFFFFFFFu, myn, src);
        const unsigned cnt = __shfl_sync(0xFFFFFFFFu, mycnt, src);
        if (stamp && lane == 0 && (P.debug & 16u) && P.trace) {  // (debugging: when the block's words are here)
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t) : "r"(cur.hraw + cur.x0 + cur.nraw) : "memory");
            atomicMin(P.trace + 30, t);
            atomicMax(P.trace + 31, t);
        }
        const long long gb = (long long)__shfl_sync(0xFFFFFFFFu, g, src);
        const int nn = (int)(P.limit - gb < (long long)cnt ? P.limit - gb : (long long)cnt);
        uint32_t* const o = outc + gb;
        IMM3_CHECK(ctrl, nn > 0 && (unsigned long long)(gb + nn) <= total && n > 0 && n <= 1024 && cnt <= (unsigned)n, 5);  // the block's rows fit the result
        (void)total;
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
        if (!done) emit_any_block(P.words + __shfl_sync(0xFFFFFFFFu, w0, src), n, S, o, nn, lane);
        if (stamp && lane == 0) phase_stamp(P, 13);
    }
}

__global__ void __launch_bounds__(kComputeThreads, 4) blocks_group_emit_kernel(const __grid_constant__ LeanPlan P, const uint32_t* __restrict__ bitmapB,
                                                                                 const uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ grp_sum,
                                                                                 long long nblocks, int ngroups, ScanCtrl* ctrl, CtrlBlock* pub,
                                                                                 unsigned long long pub_seq) {
    __shared__ GroupEmitShared GS;
    __shared__ __align__(16) uint32_t s_scratch[kComputeWarps][128];  // per warp: dense_finish's words; between rounds: the CTA's queue
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // (the count exchange of a sharded table rides behind)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) GS.start_grp = 0xFFFFFFFFu;
    if (tid == 0) phase_stamp(P, 8);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the filter kernel's counts and group sums are final
    if (tid == 0) phase_stamp(P, 9);

    // ---------------- 1. the group sums, their total, this CTA's rows of the result and the group they start in ----------------
    // (all threads: done by one warp per CTA - a fifth fewer instructions in the kernel - this phase took 4 - 7 us instead of 2 - 3)
    const int per_thread = ((ngroups + kComputeThreads - 1) / kComputeThreads + 3) & ~3;  // group sums per thread: 4, 8, 12 or 16
    uint32_t gs[16];
    {
        const uint4* src = reinterpret_cast<const uint4*>(grp_sum + (size_t)tid * (size_t)per_thread);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (4 * i < per_thread) v = __ldcg(src + i);  // (entries behind the last group are zero)
            gs[4 * i] = v.x;
            gs[4 * i + 1] = v.y;
            gs[4 * i + 2] = v.z;
            gs[4 * i + 3] = v.w;
        }
    }
    const bool sums_in_smem = ngroups <= kGrpSmemGroups;  // (then per_thread = 4)
    if (sums_in_smem) reinterpret_cast<uint4*>(GS.sums)[tid] = make_uint4(gs[0], gs[1], gs[2], gs[3]);
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) tsum += gs[i];  // (<= 16 * 2^20)
    uint32_t incl = tsum;                         // (a warp's 512 groups: <= 2^29)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += nb;
    }
    if (lane == 31) GS.warp_sum[warp] = incl;
    __syncthreads();
    unsigned long long tbase = incl - tsum, total = 0;
#pragma unroll
    for (int w = 0; w < kComputeWarps; w++) {
        const unsigned long long ws = GS.warp_sum[w];
        if (w < warp) tbase += ws;
        total += ws;
    }
    const unsigned long long want = total < (unsigned long long)P.limit ? total : (unsigned long long)P.limit;  // rows of the result
    // CTA c takes rows [floor(c want / C), floor((c + 1) want / C)): one 64-bit division, the rest in 32 bits (c r < C^2 < 2^32)
    const unsigned long long rq = want / gridDim.x;
    const uint32_t rr = (uint32_t)(want - rq * gridDim.x);
    const unsigned long long R0 = (unsigned long long)blockIdx.x * rq + (unsigned long long)(blockIdx.x * rr / gridDim.x),
                             R1 = ((unsigned long long)blockIdx.x + 1ull) * rq + (unsigned long long)((blockIdx.x + 1u) * rr / gridDim.x);
    if (tsum != 0u && R0 < R1 && R0 >= tbase && R0 < tbase + tsum) {  // the group that holds the CTA's first row lives in exactly one thread
        unsigned long long base = tbase;
        int k = 0;
        bool go = true;  // stop at the first group whose rows reach beyond R0: that one holds it
#pragma unroll
        for (int i = 0; i < 16; i++) {
            go = go && base + gs[i] <= R0;
            if (go) {
                base += gs[i];
                k = i + 1;
            }
        }
        GS.start_grp = (unsigned)(tid * per_thread + k);
        GS.start_base = base;
    }
    __syncthreads();
    if (tid == 0) phase_stamp(P, 10);

    // ---------------- 2. the CTA's rows: a block is this CTA's if its first result row lies in [R0, R1) ----------------
    if (R0 < R1 && GS.start_grp != 0xFFFFFFFFu) {
        long long k = (long long)GS.start_grp;
        unsigned long long base = GS.start_base;
        IMM3_CHECK(ctrl, k < (long long)ngroups && base <= R0, 4);  // the start group exists and begins at or before the CTA's first row
        uint2* const queue = reinterpret_cast<uint2*>(&s_scratch[0][0]);  // 256 entries: (block, first result row - R0)
        uint32_t qfill = 0;  // entries waiting in the queue (the blocks of several groups are emitted together)
        // entry 8 q + w goes to lane q of warp w: every warp gets its share however few there are
        auto flush = [&](uint32_t count) {
            __syncthreads();
            const int qn = (int)((count + (uint32_t)(kComputeWarps - 1 - warp)) / (uint32_t)kComputeWarps);
            uint2 ent = make_uint2(0u, 0u);
            if (lane < qn) ent = queue[lane * kComputeWarps + warp];
            __syncthreads();  // (the queue lives in the warps' scratch words)
            emit_queued_blocks(P, bitmapB, blk_cnt, ctrl, qn, (long long)ent.x, R0 + ent.y, want, s_scratch[warp], lane, warp == 0);
            __syncthreads();
        };
#pragma unroll 1
        while (base < R1 && k < (long long)ngroups) {
            // the next 32 groups (every warp looks for itself: no barrier): which of them hold rows of this CTA
            uint32_t s = 0;
            if (sums_in_smem) s = k + lane < kGrpSmemGroups ? GS.sums[k + lane] : 0u;
            else s = __ldcg(grp_sum + k + lane);  // (zeros behind the last group)
            uint32_t gi = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, gi, o);
                if (lane >= o) gi += nb;
            }
            const uint32_t gex = gi - s;
            unsigned gmask = __ballot_sync(0xFFFFFFFFu, s != 0u && base + gex < R1 && base + gi > R0);
#pragma unroll 1
            for (; gmask; gmask &= gmask - 1u) {
                const int j = __ffs((int)gmask) - 1;
                const unsigned long long gbase = base + __shfl_sync(0xFFFFFFFFu, gex, j);
                // ---- the group's 1024 block counts, four per thread, scanned by the CTA ----
                const long long bb = ((k + j) << kGrpShift) + 4 * tid;
                uint4 c = make_uint4(0u, 0u, 0u, 0u);
                if (bb < nblocks) c = __ldcg(reinterpret_cast<const uint4*>(blk_cnt + bb));  // (the array is padded: whole 16-byte loads)
                if (bb + 1 >= nblocks) c.y = 0u;
                if (bb + 2 >= nblocks) c.z = 0u;
                if (bb + 3 >= nblocks) c.w = 0u;
                const uint32_t s4 = c.x + c.y + c.z + c.w;
                uint32_t in4 = s4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, in4, o);
                    if (lane >= o) in4 += nb;
                }
                if (lane == 31) GS.blk_sum[warp] = in4;
                __syncthreads();
                uint32_t ex = in4 - s4, gsum = 0;
#pragma unroll
                for (int w = 0; w < kComputeWarps; w++) {
                    const uint32_t ws = GS.blk_sum[w];
                    if (w < warp) ex += ws;
                    gsum += ws;
                }
                IMM3_CHECK(ctrl, gsum == __shfl_sync(0xFFFFFFFFu, s, j), 6);  // the block counts add up to the group's sum
                (void)gsum;
                // ---- this CTA's blocks of it, ranked, then queued (256 at a time) ----
                const uint32_t cc[4] = {c.x, c.y, c.z, c.w};
                unsigned long long g4[4];
                unsigned mine = 0;
                {
                    uint32_t run = ex;
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        g4[e] = gbase + run;
                        run += cc[e];
                        if (cc[e] != 0u && g4[e] >= R0 && g4[e] < R1) mine |= 1u << e;
                    }
                }
                const uint32_t nmine = (uint32_t)__popc(mine);
                uint32_t rincl = nmine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t nb = __shfl_up_sync(0xFFFFFFFFu, rincl, o);
                    if (lane >= o) rincl += nb;
                }
                if (lane == 31) GS.mine_cnt[warp] = rincl;
                __syncthreads();
                if (tid == 0) phase_stamp(P, 11);
                uint32_t rank = rincl - nmine, nfound = 0;
#pragma unroll
                for (int w = 0; w < kComputeWarps; w++) {
                    const uint32_t wc = GS.mine_cnt[w];
                    if (w < warp) rank += wc;
                    nfound += wc;
                }
#pragma unroll 1
                for (uint32_t p0 = 0; p0 < nfound;) {
                    const uint32_t room = (uint32_t)kComputeThreads - qfill, take = nfound - p0 < room ? nfound - p0 : room;
                    uint32_t rk = rank - p0;  // (mod 2^32: entries of other rounds land at or above `take`)
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        if ((mine >> e) & 1u) {
                            if (rk < take) queue[qfill + rk] = make_uint2((uint32_t)(bb + e), (uint32_t)(g4[e] - R0));
                            rk++;
                        }
                    }
                    qfill += take;
                    p0 += take;
                    if (qfill == (uint32_t)kComputeThreads) {
                        flush(qfill);
                        qfill = 0;
                    }
                }
            }
            base += __shfl_sync(0xFFFFFFFFu, gi, 31);
            k += 32;
        }
        if (qfill) flush(qfill);
    }
    if (lane == 0) phase_stamp(P, 14);

    // ---------------- 3. last CTA out (warp 0 alone: the other warps leave) ----------------
    __syncthreads();
    if (warp == 0) {
        unsigned last = 0;
        if (lane == 0) {  // release: this CTA's rows are written (the barrier above makes that cumulative); acquire: so are everybody else's
            unsigned done;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(done) : "l"(&ctrl->exited) : "memory");
            last = done == gridDim.x - 1 ? 1u : 0u;
        }
        last = __shfl_sync(0xFFFFFFFFu, last, 0);
        if (last) {  // every CTA has read the group sums: they are the next query's again
            for (int i = lane; i < ngroups; i += 32) grp_sum[i] = 0u;
            if (lane == 0) {
                ctrl->exited = 0;
                ctrl->ticket = 0;
                ctrl->ticket2 = 0;
                ctrl->total = want;
                ctrl->dense_rows = 0;
                if (pub) {  // publish (plan.hpp): no kernel follows - the host is polling its pinned copy of the control block
                    const unsigned long long word = ((pub_seq & 0x7FFFFFull) << 41) | (__ldcg(&ctrl->error) ? (1ull << 40) : 0ull) | (want & ((1ull << 40) - 1ull));
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(&pub->pub_seq), "l"(word) : "memory");
                }
            }
        }
    }
}

// k_ptx.cuh - PTX helpers (mbarrier, TMA bulk copies, L2 policies, relaxed/volatile loads) and the tile status words of the decoupled look-back
// Fragment of kernels.cu (one translation unit, included inside namespace imm3 in the order listed there).
#pragma once

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
template <class PlanT>
__device__ __forceinline__ void phase_stamp(const PlanT& P, int ev) {
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
// Tile status words of the single-pass look-back are published with release and polled with acquire semantics: a tile
// that reads a predecessor's Prefix also observes everything that predecessor had read before publishing it.
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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

// Device-side bounds checks of the block pipeline, compiled in with -DIMM3_BOUNDS=1 (tests/test_bounds_build.py builds that
// variant and runs queries through it): compute-sanitizer is not available on this pool, so the kernels check the indices
// they derive from table metadata themselves - a failed check records 0x100 + its number in ScanCtrl::error and traps.
#ifdef IMM3_BOUNDS
#define IMM3_CHECK(ctrl, cond, code)                                              \
    do {                                                                          \
        if (!(cond)) watchdog_trap(const_cast<ScanCtrl*>(ctrl), 0x100u + (code)); \
    } while (0)
#else
#define IMM3_CHECK(ctrl, cond, code) \
    do {                             \
    } while (0)
#endif

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
    bool saw_done = false;
    const unsigned ep = epoch & 0x3FFFFFu;
    for (;;) {
        unsigned long long st[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long idx = pos - 4 * lane - k;  // lane 0 / k 0 is the nearest predecessor
            st[k] = idx >= 0 ? ld_acquire_u64(status + idx) : pack_status(epoch, kStatePrefix, 0);  // virtual tile -1: prefix 0
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
            // LIMIT reached somewhere.  The tile that crossed it had every predecessor's status in hand before it raised
            // `done`, so a tile BEFORE it finds all of its own predecessors published once it re-reads them behind a fence
            // (it must go on and emit its rows); only a tile AFTER the crossing can still see an unpublished predecessor
            // (a CTA that drew its ticket and left): that one gives up.  (Giving up at the first sight of `done` dropped
            // the rows of earlier tiles whose status reads were a moment older than the flag.)
            if (__any_sync(0xFFFFFFFFu, ld_relaxed_u32(&ctrl->done) != 0u)) {
                if (saw_done) return -1;
                saw_done = true;
                __threadfence();
                continue;
            }
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
        if (lane == 0) st_release_u64(status + tile, pack_status(P.epoch, kStatePrefix, tile_count));
    } else {
        if (lane == 0) st_release_u64(status + tile, pack_status(P.epoch, kStateAggregate, tile_count));
        excl = (P.debug & 1u) ? (long long)tile * 1800 : lookback_exclusive(status, tile, P.epoch, ctrl, lane);
        if (excl < 0) return -1;
        if (lane == 0) st_release_u64(status + tile, pack_status(P.epoch, kStatePrefix, (unsigned long long)excl + tile_count));
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


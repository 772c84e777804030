// encode.cu — GPU form of the sorted-integer codec's ENCODER (writer side).
//
// Reference: PFORCodecInt.encode (PFORCodec.scala:17-28) = JavaFastPFOR IntegratedIntCompressor.compress over one block
// of the SegmentWriter (Segment.scala:99-151; every block restarts the delta chain at 0), written as big-endian words
// plus 8 trailing zero bytes.  The word format is the one restated in oracle/oracle.c (orc_iic_compress; PARITY
// UNPINNED against the real library, see its header) and decoded by kernels.cu:
//   word 0 = n; per 128 values one header word (four 8-bit widths, first mini-block in the top byte) followed by the
//   four mini-blocks; left-over mini-blocks of 32 get a header word each; a mini-block of width b is b words holding
//   its 32 deltas LSB-first (b = 32: the 32 VALUES, raw); the last n % 32 values are delta var-bytes (7 data bits,
//   low group first, last byte of a value has bit 7 set), packed little-endian into words.
//
// One warp per block (<= 1024 values), lane m owns mini-block m: width = bits(OR of its 32 deltas), position = warp
// exclusive scan of the widths + header words before it.  Two launches: sizes, then (after an exclusive scan of the
// block sizes on the host - one int per block) the words.  Results are bit-exact with imm3_pfor_encode (tests).
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "common.hpp"

using namespace imm3;

#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            release();                                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? IMM3_ERR_OOM : IMM3_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                       \
        }                                                                                                  \
    } while (0)

namespace {

constexpr int kEncWarps = 8;

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// Var-byte length of one delta (IntegratedVariableByte: 7 data bits per byte).
__device__ __forceinline__ int varbyte_len(uint32_t v) { return v < (1u << 7) ? 1 : (v < (1u << 14) ? 2 : (v < (1u << 21) ? 3 : (v < (1u << 28) ? 4 : 5))); }

// Encode block `blk` (values in[0..n)).  WRITE = false: only the word count is produced.
// `out` points at the block's first output word (big-endian words are stored); returns the number of words (all lanes).
template <bool WRITE>
__device__ __forceinline__ int encode_block(const int32_t* __restrict__ in, int n, uint32_t* __restrict__ out, int lane) {
    const int packed = n & ~31, nmini = packed >> 5, nsuper = packed >> 7;
    // ---- lane m: width of mini-block m ----
    int b = 0;
    uint32_t prev0 = 0;
    if (lane < nmini) {
        prev0 = lane == 0 ? 0u : (uint32_t)__ldg(in + 32 * lane - 1);
        uint32_t prev = prev0, mask = 0;
        const int4* p4 = reinterpret_cast<const int4*>(in + 32 * lane);  // (blocks start 128-byte aligned: see the launcher)
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int4 v = __ldg(p4 + q);
            mask |= (uint32_t)v.x - prev;
            mask |= (uint32_t)v.y - (uint32_t)v.x;
            mask |= (uint32_t)v.z - (uint32_t)v.y;
            mask |= (uint32_t)v.w - (uint32_t)v.z;
            prev = (uint32_t)v.w;
        }
        b = mask ? 32 - __clz((int)mask) : 0;
    }
    // ---- word position of every mini-block ----
    int incl = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    const int sum_b = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const int headers_before = lane < 4 * nsuper ? (lane >> 2) + 1 : nsuper + (lane - 4 * nsuper + 1);
    const int pos = 1 + headers_before + (incl - b);  // first word of mini-block `lane`
    const int nheaders = nsuper + (nmini - 4 * nsuper);
    int nw = 1 + nheaders + sum_b;
    // ---- var-byte tail: n % 32 values, lane k owns value packed + k ----
    const int rem = n - packed;
    uint32_t tail_delta = 0;
    int tail_len = 0;
    if (lane < rem) {
        const uint32_t pv = packed + lane == 0 ? 0u : (uint32_t)__ldg(in + packed + lane - 1);
        tail_delta = (uint32_t)__ldg(in + packed + lane) - pv;
        tail_len = varbyte_len(tail_delta);
    }
    int tail_incl = tail_len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, tail_incl, o);
        if (lane >= o) tail_incl += t;
    }
    const int tail_bytes = __shfl_sync(0xFFFFFFFFu, tail_incl, 31);
    const int tail_words = (tail_bytes + 3) >> 2;
    const int tail_pos = nw;
    nw += tail_words;
    if (!WRITE) return nw;

    if (lane == 0) out[0] = bswap32((uint32_t)n);
    // ---- headers ----
    {
        const int b1 = __shfl_down_sync(0xFFFFFFFFu, b, 1), b2 = __shfl_down_sync(0xFFFFFFFFu, b, 2), b3 = __shfl_down_sync(0xFFFFFFFFu, b, 3);
        if (lane < 4 * nsuper) {
            if ((lane & 3) == 0) out[pos - 1] = bswap32(((uint32_t)b << 24) | ((uint32_t)b1 << 16) | ((uint32_t)b2 << 8) | (uint32_t)b3);
        } else if (lane < nmini) {
            out[pos - 1] = bswap32((uint32_t)b);
        }
    }
    // ---- mini-block words: 32 deltas of b bits, LSB first (b = 32: the values themselves) ----
    if (lane < nmini && b > 0) {
        const int32_t* src = in + 32 * lane;
        if (b == 32) {
#pragma unroll 4
            for (int k = 0; k < 32; k++) out[pos + k] = bswap32((uint32_t)__ldg(src + k));
        } else {
            uint32_t prev = prev0, acc = 0;
            int fill = 0, w = 0;  // `fill` bits of `acc` are valid
            for (int k = 0; k < 32; k++) {
                const uint32_t v = (uint32_t)__ldg(src + k);
                const uint32_t d = v - prev;
                prev = v;
                acc |= d << fill;
                if (fill + b >= 32) {
                    out[pos + w++] = bswap32(acc);
                    acc = fill + b > 32 ? d >> (32 - fill) : 0u;
                    fill = fill + b - 32;
                } else {
                    fill += b;
                }
            }
        }
    }
    // ---- tail bytes, little-endian inside each word, zero padded ----
    if (rem > 0) {
        // every lane writes its bytes into the word-sized slots through byte stores (output is big-endian per WORD, so
        // little-endian byte i of word w lands at byte offset 4w + (3 - i))
        uint8_t* ob = reinterpret_cast<uint8_t*>(out + tail_pos);
        for (int i = lane; i < tail_words * 4; i += 32) ob[i] = 0;  // padding (and everything else, overwritten below)
        __syncwarp();
        if (lane < rem) {
            int at = tail_incl - tail_len;
            uint32_t v = tail_delta;
            for (int i = 0; i < tail_len; i++, at++) {
                const uint32_t byte = (v & 0x7Fu) | (i == tail_len - 1 ? 0x80u : 0u);
                v >>= 7;
                ob[(at & ~3) + (3 - (at & 3))] = (uint8_t)byte;
            }
        }
    }
    return nw;
}

__global__ void __launch_bounds__(kEncWarps * 32) pfor_block_sizes_kernel(const int32_t* __restrict__ values, long long n, int block_rows,
                                                                          long long nblocks, int* __restrict__ nwords) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kEncWarps + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kEncWarps;
    for (long long blk = warp0; blk < nblocks; blk += nwarps) {
        const long long r0 = blk * block_rows;
        const int nb = (int)(n - r0 < block_rows ? n - r0 : block_rows);
        const int nw = encode_block<false>(values + r0, nb, nullptr, lane);
        if (lane == 0) nwords[blk] = nw;
    }
}

__global__ void __launch_bounds__(kEncWarps * 32) pfor_encode_blocks_kernel(const int32_t* __restrict__ values, long long n, int block_rows,
                                                                            long long nblocks, const long long* __restrict__ byte_off,
                                                                            uint8_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * kEncWarps + (threadIdx.x >> 5), nwarps = (long long)gridDim.x * kEncWarps;
    for (long long blk = warp0; blk < nblocks; blk += nwarps) {
        const long long r0 = blk * block_rows;
        const int nb = (int)(n - r0 < block_rows ? n - r0 : block_rows);
        uint32_t* ow = reinterpret_cast<uint32_t*>(out + byte_off[blk]);
        const int nw = encode_block<true>(values + r0, nb, ow, lane);
        if (lane < 2) ow[nw + lane] = 0u;  // ByteBuffer.allocate(words * 4 + 8): eight zero bytes trail (PFORCodec.scala:20)
    }
}

}  // namespace

extern "C" int64_t imm3_pfor_encode_blocks_gpu(int device, const int32_t* values, int64_t n, int32_t block_rows, uint8_t* out,
                                                int64_t out_cap, int64_t* block_off) {
    if (n < 0 || (n > 0 && !values) || block_rows < 1) return fail(IMM3_ERR_INVALID_ARG, "imm3_pfor_encode_blocks_gpu: bad arguments");
    if (block_rows > 1024 || (block_rows % 32 && n > block_rows))
        return fail(IMM3_ERR_UNSUPPORTED, "imm3_pfor_encode_blocks_gpu: block_rows must be a multiple of 32, at most 1024 (got %d)", block_rows);
    const int64_t nblocks = (n + block_rows - 1) / block_rows;
    if (nblocks == 0) {
        if (block_off) block_off[0] = 0;
        return 0;
    }
    int32_t* d_values = nullptr;
    int* d_nwords = nullptr;
    long long* d_off = nullptr;
    uint8_t* d_out = nullptr;
    bool own_values = false, own_out = false;
    auto release = [&]() {
        if (own_values) cudaFree(d_values);
        cudaFree(d_nwords);
        cudaFree(d_off);
        if (own_out) cudaFree(d_out);
        d_values = nullptr, d_nwords = nullptr, d_off = nullptr, d_out = nullptr;
    };
    CUDA_TRY(cudaSetDevice(device));
    int num_sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    // `values` / `out` may live in device memory (tables generated or loaded on the device): then nothing is staged.
    auto on_device = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
    };
    if (on_device(values)) {
        if ((reinterpret_cast<uintptr_t>(values) & 15u) != 0) return fail(IMM3_ERR_INVALID_ARG, "imm3_pfor_encode_blocks_gpu: device values must be 16-byte aligned");
        d_values = const_cast<int32_t*>(values);
    } else {
        own_values = true;
        CUDA_TRY(cudaMalloc(&d_values, (size_t)(n + 32) * 4));
        CUDA_TRY(cudaMemcpy(d_values, values, (size_t)n * 4, cudaMemcpyHostToDevice));
    }
    CUDA_TRY(cudaMalloc(&d_nwords, (size_t)nblocks * sizeof(int)));
    const int grid = (int)std::min<int64_t>((nblocks + kEncWarps - 1) / kEncWarps, (int64_t)num_sms * 8);
    pfor_block_sizes_kernel<<<grid, kEncWarps * 32>>>(d_values, n, block_rows, nblocks, d_nwords);
    CUDA_TRY(cudaGetLastError());
    std::vector<int> nwords((size_t)nblocks);
    CUDA_TRY(cudaMemcpy(nwords.data(), d_nwords, (size_t)nblocks * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<long long> off((size_t)nblocks + 1);
    off[0] = 0;
    for (int64_t i = 0; i < nblocks; i++) off[(size_t)i + 1] = off[(size_t)i] + (long long)nwords[(size_t)i] * 4 + 8;
    const int64_t total = off[(size_t)nblocks];
    if (block_off)
        for (int64_t i = 0; i <= nblocks; i++) block_off[i] = off[(size_t)i];
    if (!out) {  // sizing call
        release();
        return total;
    }
    if (out_cap < total) {
        release();
        return fail(IMM3_ERR_INVALID_ARG, "imm3_pfor_encode_blocks_gpu: need %lld bytes", (long long)total);
    }
    CUDA_TRY(cudaMalloc(&d_off, ((size_t)nblocks + 1) * sizeof(long long)));
    CUDA_TRY(cudaMemcpy(d_off, off.data(), ((size_t)nblocks + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    if (on_device(out)) {
        if ((reinterpret_cast<uintptr_t>(out) & 3u) != 0) { release(); return fail(IMM3_ERR_INVALID_ARG, "imm3_pfor_encode_blocks_gpu: device out must be 4-byte aligned"); }
        d_out = out;
    } else {
        own_out = true;
        CUDA_TRY(cudaMalloc(&d_out, (size_t)total));
    }
    pfor_encode_blocks_kernel<<<grid, kEncWarps * 32>>>(d_values, n, block_rows, nblocks, d_off, d_out);
    CUDA_TRY(cudaGetLastError());
    if (own_out) CUDA_TRY(cudaMemcpy(out, d_out, (size_t)total, cudaMemcpyDeviceToHost));
    else CUDA_TRY(cudaDeviceSynchronize());
    release();
    return total;
}

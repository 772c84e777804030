// tools/micro/pdl_residency.cu - when does a programmatic dependent become resident next to its primary?
// primary: 148 CTAs x T threads, D bytes of dynamic shared memory, spins ~50 us, launch_dependents at its first instruction.
// secondary: G CTAs x 256 threads, stamps %globaltimer at entry (before griddepcontrol.wait) and after the wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o pdl_residency pdl_residency.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
extern __shared__ unsigned char dyn[];
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(512, 1) primary(unsigned long long* st, int spin_us, int early, int has_dyn) {
    if (early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = gt();
    if (threadIdx.x == 0) { if (has_dyn) dyn[0] = 1; st[blockIdx.x] = t0; }
    while (gt() - t0 < (unsigned long long)spin_us * 1000ull + (blockIdx.x % 16) * 500ull) __nanosleep(200);
    if (threadIdx.x == 0) st[1024 + blockIdx.x] = gt();
}
__global__ void __launch_bounds__(256, 4) secondary(unsigned long long* st, int nowait) {
    __shared__ volatile unsigned int s[1376];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) { st[2048 + blockIdx.x] = gt(); }
    if (!nowait) asm volatile("griddepcontrol.wait;" ::: "memory");
    s[threadIdx.x + 1] = threadIdx.x; __syncthreads(); if (threadIdx.x == 0) st[4096 + blockIdx.x] = gt() + (s[200] == 7777u);
}
int main(int argc, char** argv) {
    unsigned long long* d; cudaMalloc(&d, 8192 * 8);
    unsigned long long h[8192];
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int dyns[] = {0, 40 * 1024, 100 * 1024, 160 * 1024, 200 * 1024, 213 * 1024};
    const int grids[] = {30, 148, 592};
    const int thr[] = {352, 256};
    for (int carve = 0; carve < 2; carve++)
    for (int ti = 0; ti < 2; ti++)
    for (int di = 0; di < 6; di++)
    for (int gi = 0; gi < 3; gi++)
    for (int early = 1; early >= 0; early--) {
        if (carve) { cudaFuncSetAttribute(secondary, cudaFuncAttributePreferredSharedMemoryCarveout, 100); cudaFuncSetAttribute(primary, cudaFuncAttributePreferredSharedMemoryCarveout, 100); }
        double best_entry = 1e9, best_ready = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaMemsetAsync(d, 0, 8192 * 8, s);
            cudaLaunchConfig_t c1 = {}; c1.gridDim = 148; c1.blockDim = thr[ti]; c1.dynamicSmemBytes = dyns[di]; c1.stream = s;
            cudaLaunchKernelEx(&c1, primary, d, 50, early, dyns[di] > 0 ? 1 : 0);
            cudaLaunchConfig_t c2 = {}; c2.gridDim = grids[gi]; c2.blockDim = 256; c2.stream = s;
            cudaLaunchAttribute a; a.id = cudaLaunchAttributeProgrammaticStreamSerialization; a.val.programmaticStreamSerializationAllowed = 1;
            c2.attrs = &a; c2.numAttrs = 1;
            cudaLaunchKernelEx(&c2, secondary, d, 0);
            cudaStreamSynchronize(s);
            cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
            unsigned long long pend = 0, e0 = ~0ull, r0 = ~0ull;
            for (int i = 0; i < 148; i++) pend = std::max(pend, h[1024 + i]);
            for (int i = 0; i < grids[gi]; i++) { e0 = std::min(e0, h[2048 + i]); r0 = std::min(r0, h[4096 + i]); }
            if (rep) { best_entry = std::min(best_entry, ((double)e0 - (double)pend) / 1000.0); best_ready = std::min(best_ready, ((double)r0 - (double)pend) / 1000.0); }
        }
        printf("carve=%d threads=%d dyn=%3dKB grid=%3d early=%d : first entry %+7.2f us, first past wait %+7.2f us (relative to the primary's last CTA end)\n", carve, thr[ti], dyns[di] / 1024, grids[gi], early, best_entry, best_ready);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}

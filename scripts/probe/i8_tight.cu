// tcgen05 kind::i8 instruction cost vs N with a warp-uniform, fully unrolled issue loop (one elected lane, no per-lane
// waterfall): cycles per instruction for N = 64 ... 256 (M = 128, K = 32), accumulating into one TMEM range or cycling
// over several, A operand fixed or alternating.  Settles whether small-N instructions have a floor of their own.
#include <cstdio>
#include <cstdlib>
#include "../../qex_b200/csrc/tc05.cuh"
using namespace qexxc::tc05;

template <int N, int NREG, int NA>
__global__ void __launch_bounds__(128, 1) tight_kernel(int iters) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < (6 * 16384 + 256 * 128) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x01010101u * (uint32_t)(i % 3);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    const int warp = warp_uniform_idx();
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (warp == 1) {
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t alo = ((smem_u32(base) >> 4) & 0x3FFFu) | (1u << 16), blo = ((smem_u32(base + 6 * 16384) >> 4) & 0x3FFFu) | (1u << 16);
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j)  // 8 instruction groups of 4 k-steps
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        mma_i8_lohi(tm + (uint32_t)((j % NREG) * N), alo + (uint32_t)(j % NA) * 1024u + k * 2u, blo + k * 2u, DESC_HI, idesc_i8(128, N), 1u);
                if ((it & 7) == 7 || it == iters - 1) mma_commit(&bar);
            }
            __syncwarp();
            if ((it & 7) == 7 || it == iters - 1) { mbar_wait(&bar, phase); phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

template <int N, int NREG, int NA>
void run(const char* name, int nsm) {
    const int smem = 6 * 16384 + 256 * 128 + 1024, iters = 4000;
    cudaFuncSetAttribute(tight_kernel<N, NREG, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    tight_kernel<N, NREG, NA><<<nsm, 128, smem>>>(200);
    cudaEventRecord(e0);
    tight_kernel<N, NREG, NA><<<nsm, 128, smem>>>(iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cyc = ms * 1e-3 * 1.965e9 / (iters * 32.0);
    printf("{\"case\": \"%s\", \"N\": %d, \"acc_ranges\": %d, \"a_tiles\": %d, \"cycles_per_instr\": %.1f, \"cycles_per_256col\": %.1f, \"err\": \"%s\"}\n", name, N, NREG, NA, cyc,
           cyc * 256.0 / N, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    run<256, 1, 1>("n256_same", nsm);
    run<256, 2, 2>("n256_2acc_2a", nsm);
    run<192, 1, 1>("n192_same", nsm);
    run<128, 1, 1>("n128_same", nsm);
    run<128, 3, 1>("n128_3acc", nsm);
    run<128, 3, 3>("n128_3acc_3a", nsm);
    run<64, 1, 1>("n64_same", nsm);
    run<64, 6, 1>("n64_6acc", nsm);
    run<64, 1, 6>("n64_same_6a", nsm);
    run<64, 6, 6>("n64_6acc_6a", nsm);
    run<32, 1, 1>("n32_same", nsm);
    return 0;
}

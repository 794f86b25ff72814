// INT8 tcgen05 throughput probe for the Ozaki-split study: how fast can one SM run
// kind::i8 MMAs (M = 128, N = 256, K = 32 per instruction, INT32 accumulators in tensor memory) on operands that are
// already in shared memory?  That is the ceiling of an exact-INT8 emulation of the FP64 contractions; dividing it by
// the 21 slice products that 1e-10 needs (tests/studies/ozaki_study.py) gives the FP64-equivalent rate to hold against
// the measured 35.5 TFLOP/s of the DMMA path.  Also checks one accumulator against the CPU.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../qex_b200/csrc/tc05.cuh"
using namespace qexxc::tc05;

__global__ void __launch_bounds__(128, 1) i8_kernel(const signed char* A, const signed char* B, int* D, int iters, int nkind, int N) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    unsigned char* As = base;              // [128 rows][128 bytes] K-major, SWIZZLE_128B
    unsigned char* Bs = base + 128 * 128;  // [256 rows][128 bytes]
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 128; i += 128) {
        const int r = i >> 7, c = i & 127;
        As[(r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 4) ^ (r & 7)) << 4) | (c & 15))] = (unsigned char)A[i];
    }
    for (int i = tid; i < 256 * 128; i += 128) {
        const int r = i >> 7, c = i & 127;
        Bs[(r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 4) ^ (r & 7)) << 4) | (c & 15))] = (unsigned char)B[i];
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (tid == 0) {
        const uint32_t sa = smem_u32(As), sb = smem_u32(Bs);
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            // 4 k-steps of 32 bytes = one 128-byte row; alternate two accumulators like a pipelined tile loop would
            const uint32_t d = tm + (it & 1) * 256;
            if (nkind == 0) {
                const uint32_t id = idesc_i8(128, N);
                for (int k = 0; k < 4; ++k) mma_i8(d, smem_desc(sa + k * 32, 16, 1024), smem_desc(sb + k * 32, 16, 1024), id, (it > 1) | (k > 0));
            } else {
                const uint32_t id = idesc_bf16(128, N, 0, 0);
                for (int k = 0; k < 4; ++k) mma_bf16(d, smem_desc(sa + k * 32, 16, 1024), smem_desc(sb + k * 32, 16, 1024), id, (it > 1) | (k > 0));
            }
            if ((it & 63) == 63 || it == iters - 1) {  // bound the number of MMAs in flight
                mma_commit(&bar);
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
    }
    __syncthreads();
    tc_fence_after();
    if (blockIdx.x == 0 && nkind == 0 && iters == 1) {
        float v[16];
        tmem_ld16(tmem_addr(tm, 32 * warp, 0), v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[tid * 16 + j] = __float_as_int(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

int main() {
    std::vector<signed char> A(128 * 128), B(256 * 128);
    srand(1);
    for (auto& x : A) x = (signed char)((rand() % 127) - 63);
    for (auto& x : B) x = (signed char)((rand() % 127) - 63);
    signed char *dA, *dB; int* dD;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * 16 * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    const int smem = 128 * 128 + 256 * 128 + 1024;
    cudaFuncSetAttribute(i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // correctness: one K = 128 product
    i8_kernel<<<1, 128, smem>>>(dA, dB, dD, 1, 0, 256);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    std::vector<int> D(128 * 16);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n) {
        int ref = 0;
        for (int k = 0; k < 128; ++k) ref += (int)A[m * 128 + k] * (int)B[n * 128 + k];
        if (ref != D[m * 16 + n]) bad++;
    }
    printf("i8 check: %d mismatches of 2048 (D[0][0] = %d)\n", bad, D[0]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    for (int N = 256; N >= 32; N >>= 1)
    for (int kind = 0; kind < 2; ++kind) {
        const int iters = 20000;
        i8_kernel<<<nsm, 128, smem>>>(dA, dB, dD, 2000, kind, N);
        cudaEventRecord(e0);
        i8_kernel<<<nsm, 128, smem>>>(dA, dB, dD, iters, kind, N);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double kk = kind == 0 ? 128.0 : 64.0;  // K elements per 4 k-steps
        const double mac = (double)nsm * iters * 128.0 * N * kk;
        printf("{\"kind\": \"%s\", \"N\": %d, \"sms\": %d, \"ms\": %.3f, \"tera_ops_per_s\": %.1f, \"cycles_per_mma\": %.1f}\n", kind == 0 ? "i8" : "bf16", N, nsm,
               ms, 2 * mac / ms / 1e9, ms * 1e-3 * 1.965e9 / (iters * 4.0));
    }
    return 0;
}

// Hardware probe for the tcgen05 building blocks of the FP32 network path (run under gpurun; no GPU here).
// Checks, with small-integer operands (exact in TF32), that the descriptors in csrc/tc05.cuh address the
// swizzled planes the way the kernels assume:
//   T1  D[m][n]  = sum_k A[m][k] Bp[n][k]        A, B K-major             (back-propagation product)
//   T2  D[m][n]  = sum_k A[m][k] W[k][n]          A K-major, B MN-major    (forward product, W stored [in][out])
//   T3  D[i][j]  = sum_p H[p][i] Z[p][j]          A, B MN-major, K = 128   (weight-gradient product, M = 128
//                                                 spans the hi and lo planes of H)
//   T4  3xTF32 accuracy of T2 on random FP32 data.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../qex_b200/csrc/tc05.cuh"
using namespace qexxc::tc05;

struct Args { const float *A, *B; float* D; int test; };

__global__ void __launch_bounds__(128, 1) probe_kernel(Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    float* Ahi = (float*)base;                 // [128][64] plane  (32 KB)
    float* Alo = (float*)(base + 32768);       // [128][64] plane
    float* Bhi = (float*)(base + 65536);       // [64 or 128][64] plane
    float* Blo = (float*)(base + 65536 + 32768);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int test = a.test;
    const int RB = test == 3 ? 128 : 64;
    // stage operands into planes (hi/lo split)
    for (int i = tid; i < 128 * 64; i += 128) {
        const int r = i >> 6, c = i & 63;
        float hi, lo; split_tf32(a.A[i], hi, lo);
        *(float*)((unsigned char*)Ahi + plane_off(r, c, 128)) = hi;
        *(float*)((unsigned char*)Alo + plane_off(r, c, 128)) = lo;
    }
    for (int i = tid; i < RB * 64; i += 128) {
        const int r = i >> 6, c = i & 63;
        float hi, lo; split_tf32(a.B[i], hi, lo);
        *(float*)((unsigned char*)Bhi + plane_off(r, c, RB)) = hi;
        *(float*)((unsigned char*)Blo + plane_off(r, c, RB)) = lo;
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 128);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (tid == 0) {
        const uint32_t sa_hi = smem_u32(Ahi), sa_lo = smem_u32(Alo), sb_hi = smem_u32(Bhi), sb_lo = smem_u32(Blo);
        if (test == 1) {
            const uint32_t id = idesc_tf32(128, 64, 0, 0);
            for (int k = 0; k < 8; ++k) mma_tf32(tm, desc_kmajor(sa_hi, 128, k), desc_kmajor(sb_hi, 64, k), id, k > 0);
        } else if (test == 2) {
            const uint32_t id = idesc_tf32(128, 64, 0, 1);
            for (int k = 0; k < 8; ++k) mma_tf32(tm, desc_kmajor(sa_hi, 128, k), desc_mnmajor(sb_hi, 64, k), id, k > 0);
        } else if (test == 3) {
            // A = [Hhi | Hlo] as one 128-column MN-major operand (atoms at stride 16 KB: hi plane then lo plane)
            const uint32_t id = idesc_tf32(128, 64, 1, 1);
            for (int k = 0; k < 16; ++k) mma_tf32(tm, desc_mnmajor(sa_hi, 128, k), desc_mnmajor(sb_hi, 128, k), id, k > 0);
        } else {
            const uint32_t id = idesc_tf32(128, 64, 0, 1);
            for (int k = 0; k < 8; ++k) mma_tf32(tm, desc_kmajor(sa_hi, 128, k), desc_mnmajor(sb_hi, 64, k), id, k > 0);
            for (int k = 0; k < 8; ++k) mma_tf32(tm, desc_kmajor(sa_lo, 128, k), desc_mnmajor(sb_hi, 64, k), id, 1);
            for (int k = 0; k < 8; ++k) mma_tf32(tm, desc_kmajor(sa_hi, 128, k), desc_mnmajor(sb_lo, 64, k), id, 1);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_addr(tm, 32 * warp, c0), v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) a.D[(size_t)tid * 64 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 128);
}

static int run(int test) {
    const int RB = test == 3 ? 128 : 64;
    std::vector<float> A(128 * 64), B(RB * 64), D(128 * 64);
    std::vector<double> ref(128 * 64, 0.0);
    srand(test);
    auto rnd = [&](bool integer) { return integer ? (float)((rand() % 9) - 4) : (float)((rand() / (double)RAND_MAX) * 2 - 1); };
    for (auto& x : A) x = rnd(test != 4);
    for (auto& x : B) x = rnd(test != 4);
    if (test == 1) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) ref[m * 64 + n] += (double)A[m * 64 + k] * B[n * 64 + k];
    } else if (test == 2 || test == 4) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) ref[m * 64 + n] += (double)A[m * 64 + k] * B[k * 64 + n];
    } else {
        // rows 0..63 of D: Hhi^T Z ; rows 64..127: Hlo^T Z (= 0 for integer data)
        for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) for (int p = 0; p < 128; ++p) ref[i * 64 + j] += (double)A[p * 64 + i] * B[p * 64 + j];
    }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    const int smem = 4 * 32768 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Args a{dA, dB, dD, test};
    probe_kernel<<<1, 128, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T%d: CUDA error %s\n", test, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int bad = 0;
    for (int i = 0; i < 128 * 64; ++i) {
        const double err = fabs((double)D[i] - ref[i]);
        if (!(err <= 1e30)) { bad++; continue; }
        if (err > maxerr) maxerr = err;
        if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
    }
    printf("T%d: max|err| = %.3e  max|ref| = %.3e  nan/inf = %d   D[0][0..3] = %g %g %g %g  ref = %g %g %g %g   D[64][0..1] = %g %g\n", test, maxerr,
           maxref, bad, D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3], D[64 * 64], D[64 * 64 + 1]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}

int main() {
    for (int t = 1; t <= 4; ++t) if (run(t)) return 1;
    return 0;
}

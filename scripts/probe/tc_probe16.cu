// Hardware probe for the bf16 (kind::f16) tcgen05 building blocks: one [R x 64] bf16 plane read as a K-major and as
// an MN-major operand, and the 6-product bf16x3 split against FP64.
//   T1  D[m][n] = sum_k A[m][k] Bp[n][k]     A, B K-major
//   T2  D[m][n] = sum_k A[m][k] W[k][n]      A K-major, B MN-major (W stored [in][out])
//   T3  D[i][j] = sum_p H[p][i] Z[p][j]      A, B MN-major, K = 128 points; M = 128 spans two consecutive planes
//   T4  bf16x3 accuracy of T2 on random FP32 data (6 products)
//   T5  bf16x3 accuracy of T3 on random data (M = 128 over planes 1|2, plus M = 64 for plane 3)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../qex_b200/csrc/tc05.cuh"
using namespace qexxc::tc05;

struct Args { const float *A, *B; float* D; int test; };
constexpr uint32_t PL = 128 * 128;  // bytes of a [128 x 64] bf16 plane

__device__ void stage(unsigned char* base, const float* src, int R, int tid) {
    for (int i = tid; i < R * 64; i += 128) {
        const int r = i >> 6, c = i & 63;
        uint32_t b1, b2, b3;
        split_bf16x3(src[i], b1, b2, b3);
        const uint32_t o = plane16_off(r, c);
        *(unsigned short*)(base + o) = (unsigned short)(b1 >> 16);
        *(unsigned short*)(base + PL + o) = (unsigned short)(b2 >> 16);
        *(unsigned short*)(base + 2 * PL + o) = (unsigned short)(b3 >> 16);
    }
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Args a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    unsigned char* Ap = base;            // 3 planes [128 x 64]
    unsigned char* Bp = base + 3 * PL;   // 3 planes (R = 64 or 128; plane stride stays PL)
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int test = a.test;
    const int RB = (test == 3 || test == 5) ? 128 : 64;
    stage(Ap, a.A, 128, tid);
    stage(Bp, a.B, RB, tid);
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tslot, 128);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (tid == 0) {
        const uint32_t sa = smem_u32(Ap), sb = smem_u32(Bp);
        if (test == 1) {
            const uint32_t id = idesc_bf16(128, 64, 0, 0);
            for (int k = 0; k < 4; ++k) mma_bf16(tm, desc16_k(sa, k), desc16_k(sb, k), id, k > 0);
        } else if (test == 2) {
            const uint32_t id = idesc_bf16(128, 64, 0, 1);
            for (int k = 0; k < 4; ++k) mma_bf16(tm, desc16_k(sa, k), desc16_mn(sb, 64, k), id, k > 0);
        } else if (test == 3) {
            const uint32_t id = idesc_bf16(128, 64, 1, 1);
            for (int k = 0; k < 8; ++k) mma_bf16(tm, desc16_mn(sa, 128, k), desc16_mn(sb, 128, k), id, k > 0);
        } else if (test == 4) {
            const uint32_t id = idesc_bf16(128, 64, 0, 1);
            const int pa[6] = {0, 0, 1, 1, 0, 2}, pb[6] = {0, 1, 0, 1, 2, 0};
            for (int t = 5; t >= 0; --t)  // smallest terms first
                for (int k = 0; k < 4; ++k)
                    mma_bf16(tm, desc16_k(sa + pa[t] * PL, k), desc16_mn(sb + pb[t] * PL, 64, k), id, !(t == 5 && k == 0));
        } else {
            // rows 0..63: A1^T (B1+B2+B3), rows 64..127: A2^T (B1+B2); second accumulator (cols 64..127): A3^T B1 (M = 64)
            const uint32_t id = idesc_bf16(128, 64, 1, 1), id64 = idesc_bf16(64, 64, 1, 1);
            for (int pbi = 2; pbi >= 0; --pbi)
                for (int k = 0; k < 8; ++k)
                    mma_bf16(tm, desc16_mn(sa, 128, k), desc16_mn(sb + pbi * PL, 128, k), id, !(pbi == 2 && k == 0));
            for (int k = 0; k < 8; ++k) mma_bf16(tm + 64, desc16_mn(sa + 2 * PL, 128, k), desc16_mn(sb, 128, k), id64, k > 0);
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < 128; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_addr(tm, 32 * warp, c0), v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) a.D[(size_t)tid * 128 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 128);
}

static int run(int test) {
    const int RB = (test == 3 || test == 5) ? 128 : 64;
    std::vector<float> A(128 * 64), B(RB * 64), D(128 * 128);
    std::vector<double> ref(128 * 64, 0.0);
    srand(test);
    const bool integer = test <= 3;
    auto rnd = [&]() { return integer ? (float)((rand() % 9) - 4) : (float)((rand() / (double)RAND_MAX) * 2 - 1); };
    for (auto& x : A) x = rnd();
    for (auto& x : B) x = rnd();
    if (test == 1) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) ref[m * 64 + n] += (double)A[m * 64 + k] * B[n * 64 + k];
    } else if (test == 2 || test == 4) {
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) ref[m * 64 + n] += (double)A[m * 64 + k] * B[k * 64 + n];
    } else {
        for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) for (int p = 0; p < 128; ++p) ref[i * 64 + j] += (double)A[p * 64 + i] * B[p * 64 + j];
    }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    const int smem = 6 * PL + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    Args a{dA, dB, dD, test};
    probe_kernel<<<1, 128, smem>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T%d: CUDA error %s\n", test, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0; int bad = 0;
    const int rows = (test == 3 || test == 5) ? 64 : 128;
    for (int r = 0; r < rows; ++r) for (int c = 0; c < 64; ++c) {
        double got = D[r * 128 + c];
        if (test == 5) {  // M = 64 accumulator layout probe: print where row r of the second accumulator landed
            got += D[(64 + r) * 128 + c];
            // try the two candidate placements of the M = 64 result: lanes r (dense) or lanes 32*(r/16) + r%16
            const int lane64 = 32 * (r / 16) + (r % 16);
            got += D[lane64 * 128 + 64 + c];
        }
        const double err = fabs(got - ref[r * 64 + c]);
        if (!(err <= 1e30)) { bad++; continue; }
        if (err > maxerr) maxerr = err;
        if (fabs(ref[r * 64 + c]) > maxref) maxref = fabs(ref[r * 64 + c]);
    }
    printf("T%d: max|err| = %.3e  max|ref| = %.3e  nan/inf = %d   D[0][0..3] = %g %g %g %g  ref = %g %g %g %g\n", test, maxerr, maxref, bad,
           D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
    if (test == 5) {
        printf("    M=64 accumulator: D[lane][64] for lanes 0,1,15,16,17,31,32,33,48,64,96: ");
        const int ls[11] = {0, 1, 15, 16, 17, 31, 32, 33, 48, 64, 96};
        for (int i = 0; i < 11; ++i) printf("%g ", D[ls[i] * 128 + 64]);
        printf("\n");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}

int main() {
    for (int t = 1; t <= 5; ++t) if (run(t)) return 1;
    return 0;
}

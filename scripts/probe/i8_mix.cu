// Issue-rate probe for the instruction mixes of the Ozaki-split contractions (contract_i8.cu): one elected thread issues a
// fixed "program" of kind::i8 MMAs (M = 128, K = 32, various N and accumulator columns) per k-step, operands resident in
// shared memory, and reports cycles per 128-byte k-chunk (4 k-steps).  Answers: is the ~105-cycle floor of small-N
// instructions a dependency on the accumulator (then independent column ranges would hide it) or an issue cost?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../qex_b200/csrc/tc05.cuh"
using namespace qexxc::tc05;

struct Prog { int n; int col[16]; int N[16]; int brow[16]; int korder; int commit_mask; int wait_mask; };

__global__ void __launch_bounds__(128, 1) mix_kernel(Prog p, int iters) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    unsigned char* As = base;              // 4 x [128 rows][128 bytes]
    unsigned char* Bs = base + 4 * 16384;  // [512 rows][128 bytes]
    __shared__ uint64_t bar, dummy, done;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (4 * 16384 + 512 * 128) / 4; i += 128) ((uint32_t*)base)[i] = 0x01010101u * (i % 3);
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&dummy, 1); mbar_init(&done, 1); mbar_arrive(&done); mbar_fence_init(); }
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
            if (p.korder == 0) {
                for (int i = 0; i < p.n; ++i) {
                    if ((p.wait_mask >> i) & 1) { mbar_wait(&done, 0); tc_fence_after(); }  // already complete: the cost of a successful wait
                    for (int k = 0; k < 4; ++k)
                        mma_i8(tm + p.col[i], smem_desc(sa + (i & 3) * 16384 + k * 32, 16, 1024), smem_desc(sb + p.brow[i] * 128 + k * 32, 16, 1024),
                               idesc_i8(128, p.N[i]), 1);
                    if ((p.commit_mask >> i) & 1) mma_commit(&dummy);
                }
            } else {
                for (int k = 0; k < 4; ++k)
                    for (int i = 0; i < p.n; ++i)
                        mma_i8(tm + p.col[i], smem_desc(sa + (i & 3) * 16384 + k * 32, 16, 1024), smem_desc(sb + p.brow[i] * 128 + k * 32, 16, 1024),
                               idesc_i8(128, p.N[i]), 1);
            }
            if ((it & 15) == 15 || it == iters - 1) {
                mma_commit(&bar);
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

static Prog mk(std::vector<std::pair<int, int>> v, int korder, int cm = 0, int wm = 0) {
    Prog p{};
    p.n = (int)v.size();
    for (int i = 0; i < p.n; ++i) { p.col[i] = v[i].first; p.N[i] = v[i].second; p.brow[i] = v[i].first % 128; }
    p.korder = korder;
    p.commit_mask = cm;
    p.wait_mask = wm;
    return p;
}

int main() {
    const int smem = 4 * 16384 + 512 * 128 + 1024;
    cudaFuncSetAttribute(mix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    struct Case { const char* name; Prog p; int cols; };
    std::vector<Case> cases = {
        {"cur64_s_outer", mk({{0,256},{256,128},{64,256},{320,64},{128,256},{192,192},{256,128},{320,64}}, 0), 64},
        {"cur64_commit_per_s", mk({{0,256},{256,128},{64,256},{320,64},{128,256},{192,192},{256,128},{320,64}}, 0, 0xEA, 0), 64},
        {"cur64_commit_wait_per_s", mk({{0,256},{256,128},{64,256},{320,64},{128,256},{192,192},{256,128},{320,64}}, 0, 0xEA, 0x75), 64},
        {"cur64_wait_per_s", mk({{0,256},{256,128},{64,256},{320,64},{128,256},{192,192},{256,128},{320,64}}, 0, 0, 0x75), 64},
        {"cur64_k_outer", mk({{0,256},{256,128},{64,256},{320,64},{128,256},{192,192},{256,128},{320,64}}, 1), 64},
        {"cur64_192split", mk({{0,192},{192,192},{64,192},{256,128},{128,256},{192,192},{256,128},{320,64}}, 0), 64},
        {"in80", mk({{0,240},{240,240},{80,208},{288,192},{160,160},{320,160},{240,240},{320,160},{400,80}}, 0), 80},
        {"in80_b", mk({{0,256},{256,224},{80,256},{336,144},{160,256},{416,64},{240,240},{320,160},{400,80}}, 0), 80},
        {"n64_same", mk({{0,64},{0,64},{0,64},{0,64},{0,64},{0,64},{0,64},{0,64}}, 0), 0},
        {"n64_8regions_k_outer", mk({{0,64},{64,64},{128,64},{192,64},{256,64},{320,64},{384,64},{448,64}}, 1), 0},
        {"n128_same", mk({{0,128},{0,128},{0,128},{0,128}}, 0), 0},
        {"n128_4regions_k_outer", mk({{0,128},{128,128},{256,128},{384,128}}, 1), 0},
        {"n256_same", mk({{0,256},{0,256}}, 0), 0},
        {"n256_2regions_k_outer", mk({{0,256},{256,256}}, 1), 0},
        {"n192_same", mk({{0,192},{0,192}}, 0), 0},
        {"n32x6_two_tiles_interleaved", mk({{0,192},{192,192},{32,160},{224,160},{64,128},{256,128},{96,96},{288,96},{128,64},{320,64},{160,32},{352,32}}, 1), 64},
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (auto& c : cases) {
        const int iters = 4000;
        mix_kernel<<<nsm, 128, smem>>>(c.p, 200);
        cudaEventRecord(e0);
        mix_kernel<<<nsm, 128, smem>>>(c.p, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        if (cudaGetLastError() != cudaSuccess) { printf("CUDA error in %s\n", c.name); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double cyc = ms * 1e-3 * 1.965e9 / iters;  // per k-chunk (4 k-steps of the whole program)
        long ncols = 0; for (int i = 0; i < c.p.n; ++i) ncols += c.p.N[i];
        printf("{\"case\": \"%s\", \"instr_per_kstep\": %d, \"cycles_per_kchunk\": %.0f, \"cycles_per_instr\": %.1f, \"cycles_per_256col\": %.1f%s", c.name, c.p.n, cyc,
               cyc / (4.0 * c.p.n), cyc / 4.0 / (ncols / 256.0), c.cols ? "" : "}\n");
        if (c.cols) printf(", \"cycles_per_kchunk_per_64_output_cols\": %.0f}\n", cyc * 64.0 / c.cols);
    }
    return 0;
}

// Stage 2 / 4-transposed on the INT8 tensor cores: an EXACT integer split (Ozaki scheme) of the FP64 product
//     q_c[g] = fac_c * sum_ij ao_c[g,i] S[i,j] ao_0[g,j]                      (rowquad of contract.cu)
// The FP64 DMMA path saturates the FP64 pipe (0.99 of cuBLAS DGEMM); this path does not use it for the N^2 work:
//
//   ao_0[g,:] = 2^(ea[g]-48) * sum_k A_k[g,:] 256^k,   S[:,j] = 2^(eb[j]-48) * sum_l B_l[:,j] 256^l,   A_k, B_l in [-128, 127]
//   (ao_0 S)[g,j] ~= 2^(ea[g]+eb[j]-56) * sum_{d=0..5} 256^(5-d) * sum_{s+t=d} (A^(s) B^(t))[g,j]     (s = 5-k, t = 5-l)
//
// six balanced 8-bit digits per operand (46 bits relative to the row / column maximum), the 21 digit products with
// s + t <= 5, each an exact INT8 x INT8 -> INT32 GEMM on tcgen05 (kind::i8): 14 bits per product + log2(N <= 1024) and at
// most six products per diagonal stay below 2^27.  Measured on the c5 operands (tests/studies/ozaki_study.py): rho to 1.6e-11 of its
// largest element, i.e. inside the 1e-10 bar of the FP64 path; the digits of a diagonal share one TMEM accumulator, the
// six accumulators are combined in 64-bit integers (exact) and converted to FP64 twice per element (high / low half).
//
// tcgen05.mma has a floor of ~105 cycles per instruction for N <= 128 (scripts/probe/i8_rate.cu: N = 64 runs at a third of
// the N = 256 rate), so the products of one A digit plane with ALL the B planes it meets are ONE instruction: the B
// planes of a k-chunk sit in shared memory as a single [6 x 64 rows] K-major tile and the accumulators of consecutive
// diagonals are adjacent TMEM column ranges, so A^(s) x [B^(0); ...; B^(5-s)] lands in accumulators s ... 5 (N = 64 (6 - s),
// split at 256): 8 instructions per k-step instead of 21.
//
// Pipeline (one CTA per 128 grid rows, 6 warps): warp 0 streams digit tiles with 1-D bulk copies (the slicing kernels
// write them in the 128-byte-swizzled K-major tile layout, so no tensor maps are needed) -- the six B tiles of a
// 128-wide k-chunk once, the six A tiles through a 4-deep ring; warp 1 issues the tcgen05.mma stream; warps 2..5 drain
// the accumulators (tcgen05.ld), recombine, scale and row-dot with the FP64 ao rows.
//
// Opt-in (QEXXC_I8=1): the slices cost 6 bytes per AO value of extra memory and one pass over the AO tensor per geometry.
#include "common.cuh"
#include "tc05.cuh"

namespace qexxc {
namespace {
using namespace tc05;

constexpr int ND = 6;            // digits per operand
constexpr int IM = 128;          // grid rows per CTA
constexpr int IN = 64;           // columns per tile (6 accumulators x 64 columns = 384 of the 512 TMEM columns)
constexpr int KC = 128;          // k-chunk: 128 int8 = one 128-byte swizzled row
constexpr uint32_t ATILE = IM * KC;  // 16 KB
constexpr uint32_t BTILE = IN * KC;  // 8 KB
constexpr int NA = 4;            // A-tile ring depth
constexpr int NEW = 8;           // epilogue warps: 4 TMEM lane quadrants x 2 column halves
constexpr int I8_THREADS = 64 + 32 * NEW;

__device__ __forceinline__ uint32_t tile_off(int r, int c) {  // byte offset of (row r, k c) inside a swizzled tile
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((((c >> 4) ^ (r & 7)) << 4) | (c & 15));
}

// balanced base-256 digits of v (|v| < 2^46), least significant first
__device__ __forceinline__ void digits6(long long v, int (&d)[ND]) {
#pragma unroll
    for (int k = 0; k < ND; ++k) {
        const int low = (int)(((v + 128) & 255) - 128);
        d[k] = low;
        v = (v - low) >> 8;
    }
}

// ---- slicing kernels ----------------------------------------------------------------------------------------------
// A8[s][row tile][k-chunk][16 KB tile], s = 0 most significant; ea[g]: x = 2^(ea-48) * sum digits.  One warp per row.
__global__ void __launch_bounds__(256) slice_ao_kernel(const double* __restrict__ ao, int Npad, int nkc, long Gpad,
                                                       signed char* __restrict__ A8, float* __restrict__ sa) {
    const long g = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= Gpad) return;
    const double* row = ao + g * Npad;
    const long tile = g >> 7;
    const int r = (int)(g & 127);
    const long nt = Gpad >> 7;
    // pass 1: row maximum
    double mx = 0.0;
    for (int c = lane * 2; c < Npad; c += 64) {
        const double2 v = *reinterpret_cast<const double2*>(row + c);
        mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    int e = 0;
    if (mx > 0.0) {
        (void)frexp(mx, &e);  // mx = m 2^e, m in [0.5, 1)
        e += 2;               // |x| / 2^e < 0.25: the top balanced digit stays below 65 after carries
    }
    if (lane == 0) sa[g] = (float)e;
    const double scale = ldexp(1.0, 48 - e);
    // pass 2: lane owns 16 consecutive columns (one 16-byte chunk) per step
    for (int c0 = lane * 16; c0 < nkc * KC; c0 += 512) {
        unsigned w[ND][4];
#pragma unroll
        for (int s = 0; s < ND; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int c = c0 + j;
            const double x = c < Npad ? row[c] : 0.0;
            int d[ND];
            digits6(__double2ll_rn(x * scale), d);
#pragma unroll
            for (int s = 0; s < ND; ++s) w[s][j >> 2] |= (unsigned)(d[ND - 1 - s] & 0xff) << (8 * (j & 3));
        }
        const int kc = c0 >> 7, cc = c0 & 127;
#pragma unroll
        for (int s = 0; s < ND; ++s) {
            signed char* t = A8 + (((long)s * nt + tile) * nkc + kc) * ATILE + tile_off(r, cc);
            *reinterpret_cast<uint4*>(t) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
        }
    }
}

// B8[t][column tile (64)][k-chunk][8 KB tile]: B[n = j][k = i] = S[i][j]; sb[j] = 2^(eb[j]-56).  One block per column j.
__global__ void __launch_bounds__(256) slice_s_kernel(const double* __restrict__ S, int ldS, int Nc, int nkc, int nct,
                                                      signed char* __restrict__ B8, double* __restrict__ sb) {
    const int j = blockIdx.x;
    __shared__ double red[256];
    double mx = 0.0;
    if (j < Nc)
        for (int i = threadIdx.x; i < Nc; i += blockDim.x) mx = fmax(mx, fabs(S[(long)i * ldS + j]));
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    mx = red[0];
    int e = 0;
    if (mx > 0.0) {
        (void)frexp(mx, &e);
        e += 2;
    }
    if (threadIdx.x == 0) sb[j] = ldexp(1.0, e - 56);
    const double scale = ldexp(1.0, 48 - e);
    const int ct = j / IN, r = j % IN;
    for (int i = threadIdx.x; i < nkc * KC; i += blockDim.x) {
        const double x = (j < Nc && i < Nc) ? S[(long)i * ldS + j] : 0.0;
        int d[ND];
        digits6(__double2ll_rn(x * scale), d);
        const int kc = i >> 7, c = i & 127;
#pragma unroll
        for (int t = 0; t < ND; ++t)
            B8[(((long)t * nct + ct) * nkc + kc) * BTILE + tile_off(r, c)] = (signed char)d[ND - 1 - t];
    }
}

// ---- the contraction ----------------------------------------------------------------------------------------------
struct I8Args {
    const signed char* A8;
    const signed char* B8;
    const float* sa;
    const double* sb;
    const double* ao;  // FP64 ao rows for the row-dot epilogue
    double* q;
    long ao_cstride, q_cstride;
    int Npad, Nc, nkc, nct, nctB, ncomp, tri;  // nct: column tiles that hold data; nctB: column tiles of the B8 layout
    long ntiles;
    double f[4];
};

__device__ __forceinline__ int i8_kend(const I8Args& a, int ct) {
    if (!a.tri) return (a.Nc + KC - 1) / KC;
    const int last = min(a.Nc, (ct + 1) * IN);  // S is upper triangular: rows i <= last column only
    return (last + KC - 1) / KC;
}

__global__ void __launch_bounds__(I8_THREADS, 1) rowquad_i8_kernel(const I8Args a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* Bs = base;                        // [2][ND][BTILE]
    unsigned char* As = base + 2 * ND * BTILE;       // [NA][ATILE]
    uint64_t* bars = (uint64_t*)(As + NA * ATILE);
    uint64_t *fullB = bars, *emptyB = bars + 2, *fullA = bars + 4, *emptyA = bars + 4 + NA, *tfull = bars + 4 + 2 * NA,
             *tempty = bars + 5 + 2 * NA;
    uint32_t* tslot = (uint32_t*)(bars + 6 + 2 * NA);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long tile = blockIdx.x;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(fullB + i, 1);
            mbar_init(emptyB + i, 1);
        }
        for (int i = 0; i < NA; ++i) {
            mbar_init(fullA + i, 1);
            mbar_init(emptyA + i, 1);
        }
        mbar_init(tfull, 1);
        mbar_init(tempty, NEW);  // one arrival per epilogue warp
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tslot;

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int itB = 0, itA = 0;
            for (int ct = 0; ct < a.nct; ++ct) {
                const int kend = i8_kend(a, ct);
                for (int kc = 0; kc < kend; ++kc, ++itB) {
                    const int kb = itB & 1;
                    mbar_wait(emptyB + kb, ((itB >> 1) & 1) ^ 1);
                    mbar_expect_tx(fullB + kb, ND * BTILE);
                    for (int t = 0; t < ND; ++t)
                        bulk_g2s(Bs + ((size_t)kb * ND + t) * BTILE, a.B8 + (((long)t * a.nctB + ct) * a.nkc + kc) * BTILE, BTILE,
                                 fullB + kb);
                    for (int s = 0; s < ND; ++s, ++itA) {
                        const int sl = itA % NA;
                        mbar_wait(emptyA + sl, ((itA / NA) & 1) ^ 1);
                        mbar_expect_tx(fullA + sl, ATILE);
                        bulk_g2s(As + (size_t)sl * ATILE, a.A8 + (((long)s * a.ntiles + tile) * a.nkc + kc) * ATILE, ATILE,
                                 fullA + sl);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t sA = smem_u32(As), sB = smem_u32(Bs);
            int itB = 0, itA = 0;
            for (int ct = 0; ct < a.nct; ++ct) {
                const int kend = i8_kend(a, ct);
                mbar_wait(tempty, (ct & 1) ^ 1);  // the epilogue has drained the accumulators of the previous column tile
                tc_fence_after();
                for (int kc = 0; kc < kend; ++kc, ++itB) {
                    const int kb = itB & 1;
                    mbar_wait(fullB + kb, (itB >> 1) & 1);
                    for (int s = 0; s < ND; ++s, ++itA) {
                        const int sl = itA % NA;
                        mbar_wait(fullA + sl, (itA / NA) & 1);
                        tc_fence_after();
                        // A^(s) x [B^(0); ...; B^(5-s)] -> accumulators s ... 5 (adjacent TMEM columns), N <= 256 per instruction
                        const int ntot = (ND - s) * IN;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t acc = (uint32_t)(kc != 0 || s != 0 || k != 0);
                            const uint64_t da = smem_desc(sA + sl * ATILE + k * 32, 16, 1024);
                            const int n0 = ntot > 256 ? 256 : ntot;
                            mma_i8(tm + (uint32_t)s * IN, da, smem_desc(sB + kb * ND * BTILE + k * 32, 16, 1024), idesc_i8(IM, n0), acc);
                            if (ntot > 256)
                                mma_i8(tm + (uint32_t)s * IN + 256, da, smem_desc(sB + kb * ND * BTILE + 4 * BTILE + k * 32, 16, 1024),
                                       idesc_i8(IM, ntot - 256), acc);
                        }
                        mma_commit(emptyA + sl);  // the slot is free once these MMAs have read it
                    }
                    mma_commit(emptyB + kb);
                }
                mma_commit(tfull);
            }
        }
    } else {
        // ---------------- epilogue: 8 warps; thread = one grid row x one half of the tile's columns ----------------
        const int qd = warp & 3;           // TMEM lane quadrant of this warp
        const int half = (warp - 2) >> 2;  // columns [32 half, 32 half + 32) of every 64-column tile
        const int r = qd * 32 + lane;
        const long g = tile * IM + r;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int ct = 0; ct < a.nct; ++ct) {
            mbar_wait(tfull, ct & 1);
            __syncwarp();
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 32 * half; c0 < 32 * half + 32; c0 += 16) {
                long long hi[16], lo[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) hi[j] = lo[j] = 0;
#pragma unroll
                for (int d = 0; d < ND; ++d) {
                    float v[16];
                    tmem_ld16(tmem_addr(tm, 32 * qd, d * IN + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const long long p = (long long)__float_as_int(v[j]);
                        if (d < 3) hi[j] += p << (8 * (2 - d));
                        else lo[j] += p << (8 * (5 - d));
                    }
                }
                const int col = ct * IN + c0;
                if (col < a.Nc) {
                    const double* sbp = a.sb + col;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // sum_d P_d 256^(5-d) = hi 2^24 + lo, each half exact in FP64 (< 2^45); columns beyond Nc carry zero digits
                        const double t = fma((double)hi[j], 16777216.0, (double)lo[j]) * sbp[j];
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < a.ncomp) acc[c] = fma(t, a.ao[(long)c * a.ao_cstride + g * a.Npad + col + j], acc[c]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty);
        }
        // the two column halves of a row: fixed-order sum through shared memory (the B tiles are free now)
        double* part = reinterpret_cast<double*>(Bs);
        if (half == 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) part[c * IM + r] = acc[c];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * NEW) : "memory");
        if (half == 0) {
            const double rs = ldexp(1.0, (int)a.sa[g]);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < a.ncomp) a.q[(long)c * a.q_cstride + g] = a.f[c] * rs * (acc[c] + part[c * IM + r]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_free(tm, 512);
}

}  // namespace

bool rowquad_i8_enabled(const qexxc_ctx* c) {
    const char* e = getenv("QEXXC_I8");
    return e && atoi(e) != 0 && c->B == 1 && !c->ao_shared;
}

size_t rowquad_i8_smem() { return 2 * ND * BTILE + NA * ATILE + 32 * 8 + 1024; }

// (re)build the digit tiles of AO component 0 for the current grid
int launch_slice_ao(qexxc_ctx* c, cudaStream_t st) {
    const int nkc = (c->Npad + KC - 1) / KC;
    const size_t need = (size_t)ND * c->GpadMax * nkc * KC;
    if (!c->i8_A) {
        QX_CUDA(cudaMalloc(&c->i8_A, need));
        QX_CUDA(cudaMalloc(&c->i8_sa, sizeof(float) * c->GpadMax));
        const int nct = (nkc * KC) / IN;
        QX_CUDA(cudaMalloc(&c->i8_B, (size_t)ND * nct * nkc * BTILE));
        QX_CUDA(cudaMalloc(&c->i8_sb, sizeof(double) * nkc * KC));
        c->allocs.push_back(c->i8_A);
        c->allocs.push_back(c->i8_sa);
        c->allocs.push_back(c->i8_B);
        c->allocs.push_back(c->i8_sb);
        c->bytes += need + (size_t)ND * nct * nkc * BTILE;
    }
    const long warps = c->Gpad;
    slice_ao_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(c->ao, c->Npad, nkc, c->Gpad, (signed char*)c->i8_A,
                                                                          c->i8_sa);
    QX_LAUNCH_CHECK(c);
    c->i8_valid = true;
    return QEXXC_OK;
}

int launch_rowquad_i8(qexxc_ctx* c, int ncomp, int tri, const double* fac4, double* q, long q_cstride, cudaStream_t st) {
    const int nkc = (c->Npad + KC - 1) / KC, nct = (nkc * KC) / IN;
    if (!c->i8_valid) {
        ProfScope prof(c, QEXXC_PROF_EVAL_AO, st);
        QX_TRY(launch_slice_ao(c, st));
    }
    ProfScope prof(c, QEXXC_PROF_ROWQUAD, st);
    slice_s_kernel<<<nkc * KC, 256, 0, st>>>(c->S, c->Npad, c->Nc, nkc, nct, (signed char*)c->i8_B, c->i8_sb);
    QX_LAUNCH_CHECK(c);
    I8Args a{};
    a.A8 = (const signed char*)c->i8_A;
    a.B8 = (const signed char*)c->i8_B;
    a.sa = c->i8_sa;
    a.sb = c->i8_sb;
    a.ao = c->ao;
    a.q = q;
    a.ao_cstride = (long)c->GpadMax * c->Npad;
    a.q_cstride = q_cstride;
    a.Npad = c->Npad;
    a.Nc = c->Nc;
    a.nkc = nkc;
    a.nct = (c->Nc + IN - 1) / IN;  // column tiles that hold data
    a.nctB = nct;
    a.ncomp = ncomp;
    a.tri = tri;
    a.ntiles = c->Gpad / IM;
    a.f[0] = (tri ? 2.0 : 1.0) * fac4[0];
    a.f[1] = fac4[1];
    a.f[2] = fac4[2];
    a.f[3] = fac4[3];
    const size_t sm = rowquad_i8_smem();
    QX_CUDA(cudaFuncSetAttribute(rowquad_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    rowquad_i8_kernel<<<(unsigned)a.ntiles, I8_THREADS, sm, st>>>(a);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

// FP64 tensor-core contractions of the XC grid integration (stages 2 and 4 and their VJPs).
//
//   rowquad : q_c[g] = fac_c * sum_ij ao_c[g,i] S[i,j] ao_0[g,j]        (S symmetric, N x N)
//             = eval_rho            numint_legacy.py:351-410,469-481   (S = sym(dm))
//             = adjoint of V_xc     SURVEY a12: wv_bar_c = rowdot(ao_c, ao_0 (V_bar+V_bar^T))
//   wsyrk   : H = ao_0^T diag(s) ao_0   or   ao_0^T Bsrc ;  out = scale * (H + H^T)
//             = _scale_ao + _dot_ao_ao + vmat+vmat.T   numint_legacy.py:308-309,336-337,432-456
//             = adjoint of eval_rho w.r.t. dm (D_bar = ao_0^T t)
//
// Both kernels are warp-specialised: one producer warp streams operand tiles into a 3-stage
// shared-memory ring with 1-D bulk async copies (TMA engine, SASS UBLKCP) completing on
// mbarriers; eight consumer warps issue DMMA.8x8x4 with register-blocked accumulators.  Rows
// in shared memory are padded by 4 doubles so every fragment load is bank-conflict free.  The
// split-G reduction of wsyrk goes through a workspace and a fixed-order reduce kernel: no
// atomics, run-to-run deterministic.
#include "common.cuh"
#include "dmma.cuh"

namespace qexxc {

namespace {

constexpr int BM = 128;     // grid rows per CTA in rowquad
constexpr int BK = 32;      // reduction-dimension slab per pipeline stage
constexpr int NSTAGE = 3;
constexpr int NCONS = 8;    // consumer warps (4 x 2)
constexpr int NTHREADS = (NCONS + 1) * 32;

template <int BN>
struct RowquadCfg {
    static constexpr int LDA = BK + 4;
    static constexpr int LDB = BN + 4;
    static constexpr int A_ELEMS = BM * LDA;
    static constexpr int B_ELEMS = BK * LDB;
    static constexpr int STAGE = A_ELEMS + B_ELEMS;
    static constexpr int RED = 4 * 2 * BM;
    static constexpr size_t SMEM = (size_t)(NSTAGE * STAGE + RED) * 8 + 2 * NSTAGE * 8;
    static constexpr uint32_t TX = (BM * BK + BK * BN) * 8;
    static constexpr int NB = BN / 16;  // n8 blocks per warp (warp tile 32 x BN/2)
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
rowquad_kernel(const double* __restrict__ ao, const double* __restrict__ S, double* __restrict__ q,
               int Npad, long ao_cstride, long ao_bstride, long S_bstride, long q_cstride,
               long q_bstride, int ncomp, int tri, double f0, double f1, double f2, double f3) {
    // tri != 0: S holds only its upper triangle (diagonal halved), so column tile nt needs the
    // reduction rows k < (nt+1)*BN only; the caller folds the factor 2 into f0.
    using Cfg = RowquadCfg<BN>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    double* red = sm + NSTAGE * Cfg::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + Cfg::RED);
    uint64_t* empty = full + NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const long g0 = (long)blockIdx.x * BM;
    const double* ao_b = ao + (long)b * ao_bstride;
    const double* A0 = ao_b + g0 * Npad;
    const double* S_b = S + (long)b * S_bstride;
    const int KB = Npad / BK, NT = Npad / BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NCONS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCONS) {
        // ---------------- producer warp ----------------
        int it = 0;
        for (int nt = 0; nt < NT; ++nt) {
            const int kend = tri ? min(KB, (nt + 1) * (BN / BK)) : KB;
            for (int kb = 0; kb < kend; ++kb, ++it) {
                const int s = it % NSTAGE;
                const uint32_t u = (uint32_t)(it / NSTAGE);
                mbar_wait(empty + s, (u & 1) ^ 1);
                double* As = sm + s * Cfg::STAGE;
                double* Bs = As + Cfg::A_ELEMS;
                if (lane == 0) mbar_expect_tx(full + s, Cfg::TX);
                __syncwarp();
#pragma unroll
                for (int r = lane; r < BM; r += 32)
                    bulk_g2s(As + r * Cfg::LDA, A0 + (long)r * Npad + kb * BK, BK * 8, full + s);
                {
                    const int r = lane;  // BK == 32 rows of S
                    bulk_g2s(Bs + r * Cfg::LDB, S_b + (long)(kb * BK + r) * Npad + nt * BN, BN * 8,
                             full + s);
                }
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, qd = lane & 3;
    constexpr int NB = Cfg::NB;
    double rp[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) rp[c][mi] = 0.0;

    int it = 0;
    for (int nt = 0; nt < NT; ++nt) {
        double acc[4][NB][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int nj = 0; nj < NB; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;

        const int kend = tri ? min(KB, (nt + 1) * (BN / BK)) : KB;
        for (int kb = 0; kb < kend; ++kb, ++it) {
            const int s = it % NSTAGE;
            const uint32_t u = (uint32_t)(it / NSTAGE);
            mbar_wait(full + s, u & 1);
            const double* As = sm + s * Cfg::STAGE + (wm * 32 + g) * Cfg::LDA + qd;
            const double* Bs = sm + s * Cfg::STAGE + Cfg::A_ELEMS + qd * Cfg::LDB + wn * (BN / 2) + g;
#pragma unroll
            for (int kk = 0; kk < BK / 4; ++kk) {
                double a[4], bf[NB];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = As[mi * 8 * Cfg::LDA + kk * 4];
#pragma unroll
                for (int nj = 0; nj < NB; ++nj) bf[nj] = Bs[kk * 4 * Cfg::LDB + nj * 8];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int nj = 0; nj < NB; ++nj) dmma884(acc[mi][nj], a[mi], bf[nj]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        // epilogue of this column tile: row-dot the (ao_0 S) tile with each AO component
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < ncomp) {
                const double* P = ao_b + (long)c * ao_cstride + (g0 + wm * 32 + g) * Npad + nt * BN +
                                  wn * (BN / 2) + 2 * qd;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    double sum = 0.0;
#pragma unroll
                    for (int nj = 0; nj < NB; ++nj) {
                        const double2 v = *reinterpret_cast<const double2*>(P + (long)mi * 8 * Npad + nj * 8);
                        sum = fma(acc[mi][nj][0], v.x, sum);
                        sum = fma(acc[mi][nj][1], v.y, sum);
                    }
                    rp[c][mi] += sum;
                }
            }
        }
    }
    // fixed-order reduction: quad lanes, then the two column-warps through shared memory
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (c < ncomp) {
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                double v = rp[c][mi];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (qd == 0) red[(c * 2 + wn) * BM + wm * 32 + mi * 8 + g] = v;
            }
        }
    }
    named_bar_sync(1, NCONS * 32);
    if (threadIdx.x < BM) {
        const double fac[4] = {f0, f1, f2, f3};
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < ncomp)
                q[(long)b * q_bstride + (long)c * q_cstride + g0 + threadIdx.x] =
                    fac[c] * (red[(c * 2 + 0) * BM + threadIdx.x] + red[(c * 2 + 1) * BM + threadIdx.x]);
    }
}

template <int BN>
struct WsyrkCfg {
    static constexpr int LD = BN + 4;
    static constexpr int T_ELEMS = BK * LD;
    static constexpr int STAGE = 2 * T_ELEMS + BK;  // A slab, B slab, scale slab
    static constexpr size_t SMEM = (size_t)(NSTAGE * STAGE) * 8 + 2 * NSTAGE * 8;
    static constexpr int MB = BN / 32;  // warp tile (BN/4) x (BN/2)
    static constexpr int NB = BN / 16;
};

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
wsyrk_kernel(const double* __restrict__ ao0, const double* __restrict__ Bsrc,
             const double* __restrict__ sc, double* __restrict__ part, int Npad, int Gpad,
             int rows_per_split, int sym, long ao_bstride, long B_bstride, long s_bstride,
             long part_bstride) {
    using Cfg = WsyrkCfg<BN>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + NSTAGE * Cfg::STAGE);
    uint64_t* empty = full + NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NT = Npad / BN;
    int ti, tj;
    if (sym) {  // upper-triangular tile pairs, row by row
        int x = blockIdx.x;
        ti = 0;
        while (x >= NT - ti) {
            x -= NT - ti;
            ++ti;
        }
        tj = ti + x;
    } else {
        ti = blockIdx.x / NT;
        tj = blockIdx.x % NT;
    }
    const int split = blockIdx.y, b = blockIdx.z;
    const long gbeg = (long)split * rows_per_split;
    long gend = gbeg + rows_per_split;
    if (gend > Gpad) gend = Gpad;
    const int KB = gend > gbeg ? (int)((gend - gbeg) / BK) : 0;
    const double* Ag = ao0 + (long)b * ao_bstride + gbeg * Npad + ti * BN;
    const double* Bg = Bsrc + (long)b * B_bstride + gbeg * Npad + tj * BN;
    const double* sg = sc ? sc + (long)b * s_bstride + gbeg : nullptr;
    const uint32_t tx = (uint32_t)((2 * BK * BN + (sc ? BK : 0)) * 8);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NCONS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == NCONS) {
        for (int it = 0; it < KB; ++it) {
            const int s = it % NSTAGE;
            const uint32_t u = (uint32_t)(it / NSTAGE);
            mbar_wait(empty + s, (u & 1) ^ 1);
            double* As = sm + s * Cfg::STAGE;
            double* Bs = As + Cfg::T_ELEMS;
            double* Ss = Bs + Cfg::T_ELEMS;
            if (lane == 0) {
                mbar_expect_tx(full + s, tx);
                if (sg) bulk_g2s(Ss, sg + (long)it * BK, BK * 8, full + s);
            }
            __syncwarp();
            const long row = (long)it * BK + lane;  // BK == 32: one slab row per lane
            bulk_g2s(As + lane * Cfg::LD, Ag + row * Npad, BN * 8, full + s);
            bulk_g2s(Bs + lane * Cfg::LD, Bg + row * Npad, BN * 8, full + s);
        }
        return;
    }

    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, qd = lane & 3;
    constexpr int MB = Cfg::MB, NB = Cfg::NB;
    double acc[MB][NB][2];
#pragma unroll
    for (int mi = 0; mi < MB; ++mi)
#pragma unroll
        for (int nj = 0; nj < NB; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;

    for (int it = 0; it < KB; ++it) {
        const int s = it % NSTAGE;
        const uint32_t u = (uint32_t)(it / NSTAGE);
        mbar_wait(full + s, u & 1);
        const double* As = sm + s * Cfg::STAGE + qd * Cfg::LD + wm * (BN / 4) + g;
        const double* Bs = sm + s * Cfg::STAGE + Cfg::T_ELEMS + qd * Cfg::LD + wn * (BN / 2) + g;
        const double* Ss = sm + s * Cfg::STAGE + 2 * Cfg::T_ELEMS + qd;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MB], bf[NB];
            const double sv = sc ? Ss[kk * 4] : 1.0;
#pragma unroll
            for (int mi = 0; mi < MB; ++mi) a[mi] = As[kk * 4 * Cfg::LD + mi * 8] * sv;
#pragma unroll
            for (int nj = 0; nj < NB; ++nj) bf[nj] = Bs[kk * 4 * Cfg::LD + nj * 8];
#pragma unroll
            for (int mi = 0; mi < MB; ++mi)
#pragma unroll
                for (int nj = 0; nj < NB; ++nj) dmma884(acc[mi][nj], a[mi], bf[nj]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    double* P = part + (long)b * part_bstride + (long)split * Npad * Npad +
                (long)(ti * BN + wm * (BN / 4) + g) * Npad + tj * BN + wn * (BN / 2) + 2 * qd;
#pragma unroll
    for (int mi = 0; mi < MB; ++mi)
#pragma unroll
        for (int nj = 0; nj < NB; ++nj)
            *reinterpret_cast<double2*>(P + (long)mi * 8 * Npad + nj * 8) =
                make_double2(acc[mi][nj][0], acc[mi][nj][1]);
}

// out[i][j] = scale * (H[i][j] + H[j][i]), H = sum over splits (fixed order) of the partial tiles.
// With sym != 0 only tile pairs ti <= tj were computed; H is symmetric there by construction.
__global__ void wsyrk_reduce_kernel(const double* __restrict__ part, double* __restrict__ out, int N,
                                    int Npad, int BN, int nsplit, int sym, double scale, int tadd,
                                    long part_bstride, long out_bstride) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    const int b = blockIdx.z;
    if (j >= N) return;
    const double* P = part + (long)b * part_bstride;
    const long NN = (long)Npad * Npad;
    double hij = 0.0, hji = 0.0;
    const int ti = i / BN, tj = j / BN;
    const bool ij_ok = !sym || ti <= tj, ji_ok = !sym || tj <= ti;
    for (int s = 0; s < nsplit; ++s) {
        if (ij_ok) hij += P[s * NN + (long)i * Npad + j];
        if (ji_ok) hji += P[s * NN + (long)j * Npad + i];
    }
    if (!ij_ok) hij = hji;
    if (!ji_ok) hji = hij;
    out[(long)b * out_bstride + (long)i * N + j] = scale * (tadd ? hij + hji : hij);
}

// S[b][i][j] (Npad x Npad, zero padded) from src[b][N][N]: mode 0 (a+a^T)/2, 1 a, 2 a+a^T.
// tri != 0 keeps the upper triangle only, diagonal halved:  x^T S x = 2 x^T triu'(S) x.
__global__ void pad_sym_kernel(const double* __restrict__ src, double* __restrict__ S, int N, int Npad,
                               int mode, int tri) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (j >= Npad) return;
    double v = 0.0;
    if (i < N && j < N) {
        const double* a = src + (long)b * N * N;
        const double x = a[(long)i * N + j], y = a[(long)j * N + i];
        v = mode == 0 ? 0.5 * (x + y) : (mode == 1 ? x : x + y);
        if (tri) v = i < j ? v : (i == j ? 0.5 * v : 0.0);
    }
    S[(long)b * Npad * Npad + (long)i * Npad + j] = v;
}

// aow[b][g][n] = sum_c f_c wv[b][c][g] ao[b][c][g][n]   (_scale_ao, numint_legacy.py:432-442)
__global__ void build_aow_kernel(const double* __restrict__ ao, const double* __restrict__ wv,
                                 double* __restrict__ aow, int Npad, long ao_cstride, long ao_bstride,
                                 long wv_cstride, long wv_bstride, long aow_bstride, long total2,
                                 double f0, double f1, double f2, double f3) {
    const int b = blockIdx.y;
    const int half = Npad >> 1;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total2;
         idx += (long)gridDim.x * blockDim.x) {
        const long gi = idx / half;
        const long off = idx * 2;
        const double* w = wv + (long)b * wv_bstride + gi;
        const double* a = ao + (long)b * ao_bstride + off;
        const double w0 = f0 * w[0], w1 = f1 * w[wv_cstride], w2 = f2 * w[2 * wv_cstride],
                     w3 = f3 * w[3 * wv_cstride];
        const double2 a0 = *reinterpret_cast<const double2*>(a);
        const double2 a1 = *reinterpret_cast<const double2*>(a + ao_cstride);
        const double2 a2 = *reinterpret_cast<const double2*>(a + 2 * ao_cstride);
        const double2 a3 = *reinterpret_cast<const double2*>(a + 3 * ao_cstride);
        double2 r;
        r.x = w0 * a0.x + w1 * a1.x + w2 * a2.x + w3 * a3.x;
        r.y = w0 * a0.y + w1 * a1.y + w2 * a2.y + w3 * a3.y;
        *reinterpret_cast<double2*>(aow + (long)b * aow_bstride + off) = r;
    }
}

int pick_bn(int Npad) { return Npad % 128 == 0 ? 128 : (Npad % 64 == 0 ? 64 : 32); }

template <typename K>
int set_smem(K kernel, size_t bytes) {
    QX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return QEXXC_OK;
}

}  // namespace

int wsyrk_pick_nsplit(int num_sms, int Npad, int Gpad, int B, bool sym) {
    const int BN = Npad % 128 == 0 ? 128 : (Npad % 64 == 0 ? 64 : 32);
    const int NT = Npad / BN;
    const long tiles = (long)(sym ? NT * (NT + 1) / 2 : NT * NT) * B;
    const int cap = Gpad / 256 > 0 ? Gpad / 256 : 1;
    int best = 1;
    double best_eff = 0.0;
    for (int w = 1; w <= 4; ++w) {
        long ns = (long)num_sms * w / tiles;
        if (ns < 1) ns = 1;
        if (ns > cap) ns = cap;
        const long ctas = tiles * ns;
        const long waves = (ctas + num_sms - 1) / num_sms;
        const double eff = (double)ctas / (double)(waves * num_sms);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best = (int)ns;
        }
        if (eff >= 0.9) break;
    }
    return best;
}

int launch_pad_sym(qexxc_ctx* c, const double* src, int mode, int tri, cudaStream_t st) {
    dim3 grid((c->Npad + 127) / 128, c->Npad, c->B);
    pad_sym_kernel<<<grid, 128, 0, st>>>(src, c->S, c->N, c->Npad, mode, tri);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_rowquad(qexxc_ctx* c, int ncomp, int tri, const double* fac4, double* q, long q_bstride,
                   long q_cstride, cudaStream_t st) {
    const int BN = pick_bn(c->Npad);
    dim3 grid(c->Gpad / BM, c->B);
    const long ao_cs = (long)c->GpadMax * c->Npad, ao_bs = ao_cs * c->C, S_bs = (long)c->Npad * c->Npad;
#define QX_RQ(BNV)                                                                               \
    do {                                                                                         \
        QX_TRY(set_smem(rowquad_kernel<BNV>, RowquadCfg<BNV>::SMEM));                            \
        rowquad_kernel<BNV><<<grid, NTHREADS, RowquadCfg<BNV>::SMEM, st>>>(                      \
            c->ao, c->S, q, c->Npad, ao_cs, ao_bs, S_bs, q_cstride, q_bstride, ncomp, tri,       \
            (tri ? 2.0 : 1.0) * fac4[0], fac4[1], fac4[2], fac4[3]);                                                          \
    } while (0)
    ProfScope prof(c, QEXXC_PROF_ROWQUAD, st);
    if (BN == 128) QX_RQ(128);
    else if (BN == 64) QX_RQ(64);
    else QX_RQ(32);
#undef QX_RQ
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_wsyrk(qexxc_ctx* c, const double* s, long s_bstride, const double* Bsrc, double scale, int tadd,
                 double* out, long out_bstride, cudaStream_t st) {
    const int BN = pick_bn(c->Npad);
    const int NT = c->Npad / BN;
    const bool sym = (Bsrc == nullptr);
    int nsplit = wsyrk_pick_nsplit(c->num_sms, c->Npad, c->Gpad, c->B, sym);
    if (nsplit > c->nsplit_max) nsplit = c->nsplit_max;
    int rps = round_up((c->Gpad + nsplit - 1) / nsplit, BK);
    const long ao_cs = (long)c->GpadMax * c->Npad, ao_bs = ao_cs * c->C;
    const long part_bs = (long)c->nsplit_max * c->Npad * c->Npad;
    const double* Bp = sym ? c->ao : Bsrc;
    const long B_bs = sym ? ao_bs : (long)c->GpadMax * c->Npad;
    dim3 grid(sym ? NT * (NT + 1) / 2 : NT * NT, nsplit, c->B);
#define QX_WS(BNV)                                                                               \
    do {                                                                                         \
        QX_TRY(set_smem(wsyrk_kernel<BNV>, WsyrkCfg<BNV>::SMEM));                                \
        wsyrk_kernel<BNV><<<grid, NTHREADS, WsyrkCfg<BNV>::SMEM, st>>>(                          \
            c->ao, Bp, s, c->part, c->Npad, c->Gpad, rps, sym ? 1 : 0, ao_bs, B_bs, s_bstride,   \
            part_bs);                                                                            \
    } while (0)
    {
        ProfScope prof(c, QEXXC_PROF_WSYRK, st);
        if (BN == 128) QX_WS(128);
        else if (BN == 64) QX_WS(64);
        else QX_WS(32);
    }
#undef QX_WS
    QX_LAUNCH_CHECK(c);
    dim3 rgrid((c->N + 127) / 128, c->N, c->B);
    wsyrk_reduce_kernel<<<rgrid, 128, 0, st>>>(c->part, out, c->N, c->Npad, BN, nsplit, sym ? 1 : 0,
                                              scale, tadd, part_bs, out_bstride);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_build_aow(qexxc_ctx* c, const double* wv, long wv_bstride, long wv_cstride,
                     const double* fac4, cudaStream_t st) {
    const long total2 = (long)c->Gpad * c->Npad / 2;
    const long ao_cs = (long)c->GpadMax * c->Npad, ao_bs = ao_cs * c->C;
    long blocks = (total2 + 255) / 256;
    if (blocks > (long)c->num_sms * 16) blocks = (long)c->num_sms * 16;
    dim3 grid((unsigned)blocks, c->B);
    build_aow_kernel<<<grid, 256, 0, st>>>(c->ao, wv, c->aow, c->Npad, ao_cs, ao_bs, wv_cstride,
                                           wv_bstride, (long)c->GpadMax * c->Npad, total2, fac4[0],
                                           fac4[1], fac4[2], fac4[3]);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

// FP64 tensor-core contractions of the XC grid integration (stages 2 and 4 and their VJPs).
//
//   rowquad : q_c[g] = fac_c * sum_ij ao_c[g,i] S[i,j] ao_0[g,j]        (S symmetric, N x N)
//             = eval_rho            numint_legacy.py:351-410,469-481   (S = sym(dm))
//             = adjoint of V_xc     SURVEY a12: wv_bar_c = rowdot(ao_c, ao_0 (V_bar+V_bar^T))
//   wsyrk   : H = ao_0^T diag(s) ao_0   or   ao_0^T Bsrc ;  out = scale * (H + H^T)
//             = _scale_ao + _dot_ao_ao + vmat+vmat.T   numint_legacy.py:308-309,336-337,432-456
//             = adjoint of eval_rho w.r.t. dm (D_bar = ao_0^T t)
//
// Both kernels are warp-specialised: one producer warp streams operand slabs into a 3-stage
// shared-memory ring with 1-D bulk async copies (TMA engine, SASS UBLKCP) completing on
// mbarriers; eight consumer warps issue DMMA.8x8x4 with register-blocked accumulators.  Rows in
// shared memory are padded by 4 doubles so every fragment load is bank-conflict free.
//
// Work that is structurally zero is never issued, at 8x8-block granularity:
//   * AO columns are padded to a multiple of 8 only (N = 1000 stays 1000): ragged last tiles run
//     with fewer n8/m8 blocks and fewer k4 steps;
//   * symmetric operands: rowquad (one AO component) multiplies by the upper triangle of S only
//     (x^T S x = 2 x^T triu'(S) x), wsyrk computes blocks on or above the diagonal only;
//   * the (wm, wn) sub-tile of a warp is chosen so that the two warps sharing an SM sub-partition
//     (warp % 4) carry equal DMMA counts in the triangular tiles.
// wsyrk is a persistent kernel over a cost-aware static schedule of (tile, grid-chunk) items
// built on the host (greedy list scheduling, chunk-major so concurrent CTAs share AO rows in
// L2); every item owns a partial-tile slot and a fixed-order reduce kernel adds the chunks and
// the transpose: no atomics, bitwise run-to-run deterministic.
#include <algorithm>

#include "common.cuh"
#include "dmma.cuh"

namespace qexxc {

namespace {

constexpr int BM = 128;     // grid rows per CTA in rowquad
constexpr int BK = 32;      // reduction-dimension slab per pipeline stage
constexpr int NSTAGE = 3;
constexpr int NCONS = 8;    // consumer warps (4 x 2)
constexpr int NTHREADS = (NCONS + 4) * 32;  // 2 consumer warpgroups + 1 producer warpgroup (1 active warp)

// Register re-partitioning between warpgroups (setmaxnreg): the kernels are compiled for 384 threads
// (168 registers each); the producer warpgroup drops to 40 and the consumers grow to 232, which removes
// the accumulator spills and lets the compiler software-pipeline fragment loads across k4 steps.
// The two instructions sit on code paths that never re-join (the producer warpgroup returns).
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n"); }
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n"); }

__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// warp -> sub-tile.  wn = warp >> 2, so warps w and w+4 (same SM sub-partition) hold the two column
// halves; for wn = 1 the row quarter is permuted so that triangular tiles balance per sub-partition.
__host__ __device__ inline void warp_tile(int warp, int& wm, int& wn) {
    wn = warp >> 2;
    wm = wn == 0 ? (warp & 3) : ((0x1023 >> (4 * (warp & 3))) & 0xF);
}

template <int BN>
struct RowquadCfg {
    static constexpr int LDA = BK + 4;
    static constexpr int LDB = BN + 4;
    static constexpr int A_ELEMS = BM * LDA;
    static constexpr int B_ELEMS = BK * LDB;
    static constexpr int STAGE = A_ELEMS + B_ELEMS;
    static constexpr int RED = 4 * 2 * BM;
    static constexpr size_t SMEM = (size_t)(NSTAGE * STAGE + RED) * 8 + 2 * NSTAGE * 8;
    static constexpr int NB = BN / 16;  // n8 blocks per warp (warp tile 32 x BN/2)
};

// chunks of the reduction dimension needed by column tile nt (tri: rows k <= last column only)
__host__ __device__ inline int rq_kend(int tri, int Nc, int BN, int nt) {
    const int KB = (Nc + BK - 1) / BK;
    if (!tri) return KB;
    const int nw = imin(BN, Nc - nt * BN);
    return imin(KB, (nt * BN + nw + BK - 1) / BK);
}

// ---- consumer-side view of the smem ring -------------------------------------------------------
struct Ring {
    double* sm;
    uint64_t* full;
    uint64_t* empty;
    int it;
};
__device__ __forceinline__ int ring_wait(const Ring& r) {
    const int s = r.it % NSTAGE;
    mbar_wait(r.full + s, (uint32_t)(r.it / NSTAGE) & 1);
    return s;
}
__device__ __forceinline__ void ring_release(Ring& r, int s, int lane) {
    // The stage was read through the generic proxy (LDS) and will be overwritten through the async proxy
    // (cp.async.bulk): the mbarrier release/acquire chain alone does not order the two proxies.  Without this
    // fence a refill could overtake the last fragment loads of a slab -- seen as run-to-run differences of a
    // few rows of rho in one build variant (scripts/det_probe.py), never as a parity failure.
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(r.empty + s);
    ++r.it;
}
__device__ __forceinline__ void ring_skip(Ring& r, int n, int lane) {
    for (int k = 0; k < n; ++k) {
        const int s = ring_wait(r);
        ring_release(r, s, lane);
    }
}

// One 32-deep slab of the rowquad product for a warp: acc[4][NV] += A(32 x 32) * B(32 x 8*NV).
// REL < 0: dense.  REL = 0 / 32: the slab starts REL rows below the first column of the warp's
// sub-tile inside the upper-triangular S, so column block nj only sees k4 steps with
// (REL + 4*kk) >> 3 <= nj.  Every condition folds at compile time after unrolling: the skip
// patterns cost no predicates or branches in the DMMA stream.
// The two column-warps of a row quarter own the n8 blocks of the tile INTERLEAVED: warp wn holds
// blocks 2*j + wn.  In the triangular region both warps then carry (almost) the same DMMA count in
// every slab, so the sub-partition they share never idles on one of them.
// SD < 0: dense.  SD = 0..3: the slab is the SD-th 32-row slab of the diagonal block of this column
// tile; tile block nb only sees k4 steps with 4*SD + (kk >> 1) <= nb.
template <int BN, int NV, int SD, int WN>
__device__ __forceinline__ void rq_slab(double (&acc)[4][BN / 16][2], const double* __restrict__ As,
                                        const double* __restrict__ Bs) {
    using Cfg = RowquadCfg<BN>;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
        const int nbmin = SD < 0 ? 0 : 4 * SD + (kk >> 1);  // first tile block that sees rows of this k4 step
        if (nbmin <= 2 * (NV - 1) + WN) {
            // all fragment loads of the k4 step first, then the DMMAs (A-outer): keeps the shared-memory
            // latency off the tensor pipe's critical path
            double a[4], bf[NV];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) a[mi] = As[mi * 8 * Cfg::LDA + kk * 4];
#pragma unroll
            for (int j = 0; j < NV; ++j)
                if (2 * j + WN >= nbmin) bf[j] = Bs[kk * 4 * Cfg::LDB + j * 16];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    if (2 * j + WN >= nbmin) dmma884(acc[mi][j], a[mi], bf[j]);
        }
    }
}
template <int BN, int NV, int SD, int WN>
__device__ __forceinline__ void rq_run(double (&acc)[4][BN / 16][2], Ring& r, int nslabs, int aoff, int boff,
                                       int lane) {
    using Cfg = RowquadCfg<BN>;
    for (int k = 0; k < nslabs; ++k) {
        const int s = ring_wait(r);
        const double* st = r.sm + s * Cfg::STAGE;
        rq_slab<BN, NV, SD, WN>(acc, st + aoff, st + Cfg::A_ELEMS + boff);
        ring_release(r, s, lane);
    }
}
// dense slabs with a run-time number of valid column blocks -> compile-time variant
template <int BN, int NV>
__device__ __forceinline__ void rq_dense(double (&acc)[4][BN / 16][2], Ring& r, int nslabs, int nvalid, int aoff,
                                         int boff, int lane) {
    if (nvalid == NV) rq_run<BN, NV, -1, 0>(acc, r, nslabs, aoff, boff, lane);
    else if constexpr (NV > 1) rq_dense<BN, NV - 1>(acc, r, nslabs, nvalid, aoff, boff, lane);
    else ring_skip(r, nslabs, lane);
}
// one full-width diagonal slab: run-time (sd, wn) -> compile-time variant
template <int BN, int SD>
__device__ __forceinline__ void rq_diag(double (&acc)[4][BN / 16][2], Ring& r, int sd, int wn, int aoff, int boff,
                                        int lane) {
    constexpr int NB = BN / 16;
    if (sd == SD) {
        if (wn == 0) rq_run<BN, NB, SD, 0>(acc, r, 1, aoff, boff, lane);
        else rq_run<BN, NB, SD, 1>(acc, r, 1, aoff, boff, lane);
    } else if constexpr (SD > 0) {
        rq_diag<BN, SD - 1>(acc, r, sd, wn, aoff, boff, lane);
    }
}
template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
rowquad_kernel(const double* __restrict__ ao, const double* __restrict__ S, double* __restrict__ q,
               int Npad, int Nc, long ao_cstride, long ao_bstride, long S_bstride, long q_cstride,
               long q_bstride, int ncomp, int tri, double f0, double f1, double f2, double f3, int ldS, int Sc,
               const double* __restrict__ sgn, int nbulk, int ntail, double* __restrict__ qpart, int npair,
               unsigned mask0, unsigned mask1, double* __restrict__ qpair) {
    // tri != 0: S holds only its upper triangle (diagonal halved); the caller folds the factor 2
    // into f0.  Npad = storage pitch (multiple of 32, pad columns are zeros), Nc = compute extent
    // (multiple of 8): slabs are always full, column blocks beyond Nc are never issued.
    // S is [Npad rows][ldS pitch] with Sc compute columns: the square symmetric operand (ldS = Npad,
    // Sc = Nc) or, when sgn != nullptr, the occupation-scaled MO coefficients L = C sqrt|occ| of pyscf's
    // eval_rho2 (numint_legacy.py:527-545): then q[g] = f0 * sum_k sgn_k ((ao L)[g,k])^2.
    // Grid: blocks [0, nbulk) own one 128-row tile each and sweep all column tiles.  The last `ntail` row
    // tiles -- the partial wave that would leave SMs idle for a whole tile time -- are cut into one block per
    // (row tile, column tile), issued heaviest column tile first (list scheduling by the hardware block
    // dispatcher); their per-column-tile results go to qpart and rowquad_tail_kernel adds them in order.
    using Cfg = RowquadCfg<BN>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    double* red = sm + NSTAGE * Cfg::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + Cfg::RED);
    uint64_t* empty = full + NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const int NT = (Sc + BN - 1) / BN;
    // Pair mode (npair != 0, wide N): every bulk row tile is shared by TWO adjacent blocks that take complementary,
    // equally heavy sets of column tiles (mask0 / mask1).  A block then re-streams its 128 ao rows half as often and,
    // more to the point, only num_sms / 2 distinct row tiles are in flight, so the rows (8 KB each at N = 1000) stay
    // in L2 between the sweeps instead of being re-read from HBM; the two partial row sums are added in a fixed
    // order by rowquad_pair_kernel.
    int tile = blockIdx.x, nt_lo = 0, nt_hi = NT, half = -1;
    unsigned mask = 0xffffffffu;
    const int nb2 = npair ? 2 * nbulk : nbulk;
    const bool split = (int)blockIdx.x >= nb2;
    if (split) {
        const int idx = blockIdx.x - nb2;
        nt_lo = NT - 1 - idx / ntail;
        nt_hi = nt_lo + 1;
        tile = nbulk + idx % ntail;
    } else if (npair) {
        tile = blockIdx.x >> 1;
        half = blockIdx.x & 1;
        mask = half ? mask1 : mask0;
    }
    const long g0 = (long)tile * BM;
    const double* ao_b = ao + (long)b * ao_bstride;
    const double* A0 = ao_b + g0 * Npad;
    const double* S_b = S + (long)b * S_bstride;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NCONS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= NCONS) {
        reg_dec();
        // ---------------- producer warpgroup: 4 warps share the copy issue ----------------
        // (the A slab is 128 row copies of 256 B: one warp alone cannot issue them fast enough to keep
        // the cheap slabs of the triangular region fed).  Warp pw copies A rows 32*pw..32*pw+31; warp 0
        // also arms the barrier with the slab's byte count and copies the S rows.
        const int pw = warp - NCONS;
        int it = 0;
        for (int nt = nt_lo; nt < nt_hi; ++nt) {
            if (!((mask >> nt) & 1u)) continue;
            const int nw = imin(BN, ldS - nt * BN);  // copy width: storage columns (zeros beyond Sc)
            const int kend = rq_kend(tri, Nc, BN, nt);
            for (int kb = 0; kb < kend; ++kb, ++it) {
                const int s = it % NSTAGE;
                const uint32_t u = (uint32_t)(it / NSTAGE);
                mbar_wait(empty + s, (u & 1) ^ 1);
                double* As = sm + s * Cfg::STAGE;
                double* Bs = As + Cfg::A_ELEMS;
                if (pw == 0 && lane == 0) mbar_expect_tx(full + s, (uint32_t)((BM * BK + BK * nw) * 8));
                const int r = pw * 32 + lane;
                bulk_g2s(As + r * Cfg::LDA, A0 + (long)r * Npad + kb * BK, BK * 8, full + s);
                if (pw == 0)
                    bulk_g2s(Bs + lane * Cfg::LDB, S_b + (long)(kb * BK + lane) * ldS + nt * BN, nw * 8, full + s);
            }
        }
        return;
    }
    reg_inc();

    // ---------------- consumer warps ----------------
    int wm, wn;
    warp_tile(warp, wm, wn);
    const int g = lane >> 2, qd = lane & 3;
    constexpr int NB = Cfg::NB;
    double rp[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) rp[c][mi] = 0.0;

    Ring ring{sm, full, empty, 0};
    const int aoff = (wm * 32 + g) * Cfg::LDA + qd;
    const int boff = qd * Cfg::LDB + wn * 8 + g;  // this warp's blocks are 2*j + wn: 16 doubles apart
    for (int nt = nt_lo; nt < nt_hi; ++nt) {
        if (!((mask >> nt) & 1u)) continue;
        const int nw = imin(BN, Sc - nt * BN);
        const int kend = rq_kend(tri, Nc, BN, nt);
        const int nbv = nw >> 3;                             // valid n8 blocks of this tile
        const int nvalid = clampi((nbv - wn + 1) >> 1, 0, NB);  // ... of which this warp owns 2*j + wn < nbv
        const int c0 = nt * BN;                              // first column of the tile
        double acc[4][NB][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int nj = 0; nj < NB; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;

        // leading slabs entirely above the diagonal block of the triangular S: dense
        const int kd = tri ? imin(kend, c0 / BK) : kend;
        rq_dense<BN, NB>(acc, ring, kd, nvalid, aoff, boff, lane);
        // slabs of the diagonal block (tri only).  A ragged last tile runs the full-tile variants: its
        // extra column blocks read zeros / stale data and are ignored by the epilogue.
        for (int kb = kd; kb < kend; ++kb) {
            if (nvalid == 0) ring_skip(ring, 1, lane);
            else rq_diag<BN, BN / BK - 1>(acc, ring, (kb * BK - c0) / BK, wn, aoff, boff, lane);
        }
        if (sgn != nullptr) {
            // MO form: signed sum of squares of the (ao L) tile
            const double* sg = sgn + (long)b * ldS + c0 + wn * 8 + 2 * qd;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                double sum = 0.0;
#pragma unroll
                for (int nj = 0; nj < NB; ++nj) {
                    if (nj < nvalid) {
                        const double2 sv = *reinterpret_cast<const double2*>(sg + nj * 16);
                        sum = fma(acc[mi][nj][0] * acc[mi][nj][0], sv.x, sum);
                        sum = fma(acc[mi][nj][1] * acc[mi][nj][1], sv.y, sum);
                    }
                }
                rp[0][mi] += sum;
            }
            continue;
        }
        // epilogue of this column tile: row-dot the (ao_0 S) tile with each AO component
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < ncomp) {
                const double* P = ao_b + (long)c * ao_cstride + (g0 + wm * 32 + g) * Npad + c0 + wn * 8 + 2 * qd;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    double sum = 0.0;
#pragma unroll
                    for (int nj = 0; nj < NB; ++nj) {
                        if (nj < nvalid) {
                            const double2 v = *reinterpret_cast<const double2*>(P + (long)mi * 8 * Npad + nj * 16);
                            sum = fma(acc[mi][nj][0], v.x, sum);
                            sum = fma(acc[mi][nj][1], v.y, sum);
                        }
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
            if (c < ncomp) {
                const double v = fac[c] * (red[(c * 2 + 0) * BM + threadIdx.x] + red[(c * 2 + 1) * BM + threadIdx.x]);
                if (split) qpart[((long)(nt_lo * 4 + c) * ntail + (tile - nbulk)) * BM + threadIdx.x] = v;
                else if (half >= 0) qpair[((long)(half * 4 + c) * nbulk + tile) * BM + threadIdx.x] = v;
                else q[(long)b * q_bstride + (long)c * q_cstride + g0 + threadIdx.x] = v;
            }
    }
}

// q[c][row] = half 0 + half 1 of the paired bulk tiles
__global__ void rowquad_pair_kernel(const double* __restrict__ qpair, double* __restrict__ q, long q_cstride, int ncomp,
                                    int nbulk) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (long)nbulk * BM) return;
    for (int c = 0; c < ncomp; ++c)
        q[(long)c * q_cstride + r] = qpair[(long)c * nbulk * BM + r] + qpair[(long)(4 + c) * nbulk * BM + r];
}

// q[c][row] = sum over the column tiles (in order) of the split tail tiles' partial results
__global__ void rowquad_tail_kernel(const double* __restrict__ qpart, double* __restrict__ q, long q_cstride,
                                    int ncomp, int NT, int nbulk, int ntail) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ntail * BM) return;
    for (int c = 0; c < ncomp; ++c) {
        double s = 0.0;
        for (int nt = 0; nt < NT; ++nt) s += qpart[((long)(nt * 4 + c) * ntail) * BM + r];
        q[(long)c * q_cstride + (long)nbulk * BM + r] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// wsyrk
// ---------------------------------------------------------------------------------------------
struct WsItem {
    int b, ti, tj;   // batch element, output tile (row, column)
    int g0, kb;      // first grid row and number of 32-row slabs of this item
    int slot;        // partial-tile slot
    int diag;        // symmetric diagonal tile: only blocks on or above the diagonal
    int pad;
};

template <int BN>
struct WsyrkCfg {
    static constexpr int LD = BN + 4;
    static constexpr int T_ELEMS = BK * LD;
    static constexpr int STAGE = 2 * T_ELEMS + BK;  // A slab, B slab, scale slab
    static constexpr size_t SMEM = (size_t)(NSTAGE * STAGE) * 8 + 2 * NSTAGE * 8;
    static constexpr int MB = BN / 32;  // warp tile (BN/4) x (BN/2)
    static constexpr int NB = BN / 16;
};

// One 32-row slab of the wsyrk product for a warp: acc[MB][NV] += (A .* s)^T(8*MB x 32) * B(32 x 8*NV).
// DS is the triangular shift of a diagonal tile: block (mi, nj) is needed iff nj >= DS + mi
// (DS <= -(MB-1): dense).  Conditions fold at compile time.
template <int BN, int NV, int DS>
__device__ __forceinline__ void ws_slab(double (&acc)[BN / 32][BN / 16][2], const double* __restrict__ As,
                                        const double* __restrict__ Bs, const double* __restrict__ Ss, bool scaled) {
    using Cfg = WsyrkCfg<BN>;
    constexpr int MB = Cfg::MB;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
        double a[MB];
        const double sv = scaled ? Ss[kk * 4] : 1.0;
#pragma unroll
        for (int mi = 0; mi < MB; ++mi) a[mi] = As[kk * 4 * Cfg::LD + mi * 8] * sv;
        double bf[NV];
#pragma unroll
        for (int nj = 0; nj < NV; ++nj)
            if (nj >= DS) bf[nj] = Bs[kk * 4 * Cfg::LD + nj * 8];
#pragma unroll
        for (int mi = 0; mi < MB; ++mi)
#pragma unroll
            for (int nj = 0; nj < NV; ++nj)
                if (nj >= DS + mi) dmma884(acc[mi][nj], a[mi], bf[nj]);
    }
}
template <int BN, int NV, int DS>
__device__ __forceinline__ void ws_run(double (&acc)[BN / 32][BN / 16][2], Ring& r, int nslabs, int aoff, int boff,
                                       int soff, bool scaled, int lane) {
    using Cfg = WsyrkCfg<BN>;
    for (int k = 0; k < nslabs; ++k) {
        const int s = ring_wait(r);
        const double* st = r.sm + s * Cfg::STAGE;
        ws_slab<BN, NV, DS>(acc, st + aoff, st + Cfg::T_ELEMS + boff, st + 2 * Cfg::T_ELEMS + soff, scaled);
        ring_release(r, s, lane);
    }
}
template <int BN, int NV>
__device__ __forceinline__ void ws_dense(double (&acc)[BN / 32][BN / 16][2], Ring& r, int nslabs, int nvalid,
                                         int aoff, int boff, int soff, bool scaled, int lane) {
    if (nvalid == NV) ws_run<BN, NV, -100>(acc, r, nslabs, aoff, boff, soff, scaled, lane);
    else if constexpr (NV > 1) ws_dense<BN, NV - 1>(acc, r, nslabs, nvalid, aoff, boff, soff, scaled, lane);
    else ring_skip(r, nslabs, lane);
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
wsyrk_kernel(const double* __restrict__ ao0, const double* __restrict__ Bsrc,
             const double* __restrict__ sc, double* __restrict__ part, int Npad, int Nc,
             const WsItem* __restrict__ items, const int* __restrict__ cta_start, long ao_bstride,
             long B_bstride, long s_bstride) {
    using Cfg = WsyrkCfg<BN>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + NSTAGE * Cfg::STAGE);
    uint64_t* empty = full + NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = cta_start[blockIdx.x], last = cta_start[blockIdx.x + 1];

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NCONS);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= NCONS) {
        reg_dec();
        // producer warpgroup: the 64 row copies of a slab are spread over 4 warps (warp pw: slab rows
        // 8*pw..8*pw+7, lanes 0-7 the A rows, lanes 8-15 the B rows); warp 0 arms the barrier
        const int pw = warp - NCONS;
        int it = 0;
        for (int w = first; w < last; ++w) {
            const WsItem im = items[w];
            const int iw = imin(BN, Npad - im.ti * BN), jw = imin(BN, Npad - im.tj * BN);
            const double* Ag = ao0 + (long)im.b * ao_bstride + (long)im.g0 * Npad + im.ti * BN;
            const double* Bg = Bsrc + (long)im.b * B_bstride + (long)im.g0 * Npad + im.tj * BN;
            const double* sg = sc ? sc + (long)im.b * s_bstride + im.g0 : nullptr;
            const uint32_t tx = (uint32_t)((BK * (iw + jw) + (sc ? BK : 0)) * 8);
            const int r = pw * 8 + (lane & 7);
            for (int kc = 0; kc < im.kb; ++kc, ++it) {
                const int s = it % NSTAGE;
                const uint32_t u = (uint32_t)(it / NSTAGE);
                mbar_wait(empty + s, (u & 1) ^ 1);
                double* As = sm + s * Cfg::STAGE;
                double* Bs = As + Cfg::T_ELEMS;
                double* Ss = Bs + Cfg::T_ELEMS;
                if (pw == 0 && lane == 0) {
                    mbar_expect_tx(full + s, tx);
                    if (sg) bulk_g2s(Ss, sg + (long)kc * BK, BK * 8, full + s);
                }
                const long row = (long)kc * BK + r;
                if (lane < 8) bulk_g2s(As + r * Cfg::LD, Ag + row * Npad, iw * 8, full + s);
                else if (lane < 16) bulk_g2s(Bs + r * Cfg::LD, Bg + row * Npad, jw * 8, full + s);
            }
        }
        return;
    }
    reg_inc();

    int wm, wn;
    warp_tile(warp, wm, wn);
    const int g = lane >> 2, qd = lane & 3;
    constexpr int MB = Cfg::MB, NB = Cfg::NB;
    Ring ring{sm, full, empty, 0};
    const int aoff = qd * Cfg::LD + wm * (BN / 4) + g;
    const int boff = qd * Cfg::LD + wn * (BN / 2) + g;
    const int soff = qd;
    const bool scaled = sc != nullptr;
    for (int w = first; w < last; ++w) {
        const WsItem im = items[w];
        const int iw = imin(BN, Nc - im.ti * BN), jw = imin(BN, Nc - im.tj * BN);  // compute extents
        const int mvalid = clampi((iw - wm * (BN / 4)) / 8, 0, MB);
        const int nvalid = clampi((jw - wn * (BN / 2)) / 8, 0, NB);
        double acc[MB][NB][2];
#pragma unroll
        for (int mi = 0; mi < MB; ++mi)
#pragma unroll
            for (int nj = 0; nj < NB; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;

        // Variant of the slab loop for this (item, warp).  Ragged rows of the tile are computed on
        // stale shared-memory data and ignored downstream (separate accumulators, never read).
        const int ds = wm * MB - wn * NB;  // diagonal tile: block (mi, nj) needed iff nj >= ds + mi
        if (!im.diag) {
            ws_dense<BN, NB>(acc, ring, im.kb, mvalid > 0 ? nvalid : 0, aoff, boff, soff, scaled, lane);
        } else {
            // diagonal tile (ragged or not): compile-time triangular variants; blocks beyond Nc are
            // computed on zeros / stale data and ignored by the reduce kernel
            if (ds <= -(MB - 1)) ws_run<BN, NB, -100>(acc, ring, im.kb, aoff, boff, soff, scaled, lane);
            else if (ds >= NB) ring_skip(ring, im.kb, lane);
            else if (ds == 0) ws_run<BN, NB, 0>(acc, ring, im.kb, aoff, boff, soff, scaled, lane);
            else ws_run<BN, NB, MB>(acc, ring, im.kb, aoff, boff, soff, scaled, lane);  // ds == MB
        }
        // compact partial tile [BN][BN] of this item
        double* P = part + (long)im.slot * BN * BN + (long)(wm * (BN / 4) + g) * BN + wn * (BN / 2) + 2 * qd;
#pragma unroll
        for (int mi = 0; mi < MB; ++mi)
#pragma unroll
            for (int nj = 0; nj < NB; ++nj)
                *reinterpret_cast<double2*>(P + (long)mi * 8 * BN + nj * 8) = make_double2(acc[mi][nj][0], acc[mi][nj][1]);
    }
}

// out[i][j] = scale * (H[i][j] + (tadd ? H[j][i] : 0)), H = sum of the partial tiles of a tile (slots
// slot_start[tile] .. slot_start[tile+1], in grid order).
// With sym != 0 only 8x8 blocks on or above the diagonal were computed (H symmetric).  A block is
// 32 columns x RK chunk groups: group kg adds chunks kg, kg+RK, ... and the RK sub-sums are combined
// through shared memory in a fixed order (bit-reproducible, and the dependent-load chain is RK x shorter).
constexpr int RK = 8;
__global__ void __launch_bounds__(32 * RK)
wsyrk_reduce_kernel(const double* __restrict__ part, double* __restrict__ out, int N, int BN, int NT,
                    const int* __restrict__ slot_start, int ntile, int sym, double scale, int tadd, long out_bstride) {
    __shared__ double red[2][RK][32];
    const int lane = threadIdx.x & 31, kg = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + lane;
    const int i = blockIdx.y;
    const int b = blockIdx.z;
    const long tsz = (long)BN * BN;
    auto H = [&](int r, int c) {
        const int tr = r / BN, tc = c / BN;
        const int tile = sym ? tr * NT - tr * (tr - 1) / 2 + (tc - tr) : tr * NT + tc;
        const int s0 = slot_start[b * ntile + tile], cnt = slot_start[b * ntile + tile + 1] - s0;
        const double* P = part + (long)s0 * tsz + (long)(r - tr * BN) * BN + (c - tc * BN);
        double h = 0.0;
        for (int k = kg; k < cnt; k += RK) h += P[(long)k * tsz];
        return h;
    };
    // Symmetric case: only elements in 8x8 blocks on or above the diagonal are summed (coalesced rows of the
    // partial tiles); the result is also written to the mirrored position, so no thread ever walks a
    // partial tile column-wise.  Inside a diagonal block both H(i,j) and H(j,i) exist and are both read.
    if (sym && blockIdx.x * 32 + 31 < (i & ~7)) return;  // whole block below the diagonal blocks
    const bool up = (j >> 3) >= (i >> 3), dg = (j >> 3) == (i >> 3);
    double h0 = 0.0, h1 = 0.0;  // sub-sums of H(i,j) and of the transposed term H(j,i)
    if (j < N) {
        if (sym) {
            if (up) h0 = H(i, j);
            if (tadd && dg) h1 = H(j, i);
        } else {
            h0 = H(i, j);
            if (tadd) h1 = H(j, i);
        }
    }
    red[0][kg][lane] = h0;
    red[1][kg][lane] = h1;
    __syncthreads();
    if (kg != 0 || j >= N || (sym && !up)) return;
    double a = 0.0, t = 0.0;
#pragma unroll
    for (int k = 0; k < RK; ++k) {
        a += red[0][k][lane];
        t += red[1][k][lane];
    }
    if (sym && !dg) t = a;  // H(j,i) == H(i,j) was not computed separately
    const double v = scale * (tadd ? a + t : a);
    out[(long)b * out_bstride + (long)i * N + j] = v;
    if (sym && !dg) out[(long)b * out_bstride + (long)j * N + i] = v;
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

// L[b][i][k] = C[b][i][k] * sqrt(|occ[b][k]|) for |occ| > 1e-12 (pyscf OCCDROP), zero padded; sgn = sign(occ)
__global__ void pack_mo_kernel(const double* __restrict__ C, const double* __restrict__ occ, double* __restrict__ L,
                               double* __restrict__ sgn, int N, int nmo, int Npad, int ldL) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, b = blockIdx.z;
    if (k >= ldL) return;
    double v = 0.0, s = 0.0;
    if (k < nmo) {
        const double o = occ[(long)b * nmo + k];
        if (fabs(o) > 1e-12) {
            s = o > 0 ? 1.0 : -1.0;
            if (i < N) v = C[((long)b * N + i) * nmo + k] * sqrt(fabs(o));
        }
    }
    L[((long)b * Npad + i) * ldL + k] = v;
    if (i == 0) sgn[(long)b * ldL + k] = s;
}

int pick_bn(int Npad) { return Npad > 64 ? 128 : (Npad > 32 ? 64 : 32); }

template <typename K>
int set_smem(K kernel, size_t bytes) {
    QX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return QEXXC_OK;
}

// ---- host-side cost model (the same predicates as the kernels) -------------------------------------
// number of 8x8x4 DMMAs per k4 step issued by consumer warp `w` for a wsyrk tile
int ws_warp_blocks(int BN, int Nc, int ti, int tj, int diag, int w) {
    int wm, wn;
    warp_tile(w, wm, wn);
    const int MB = BN / 32, NB = BN / 16;
    const int iw = std::min(BN, Nc - ti * BN), jw = std::min(BN, Nc - tj * BN);
    const int mvalid = clampi((iw - wm * (BN / 4)) / 8, 0, MB), nvalid = clampi((jw - wn * (BN / 2)) / 8, 0, NB);
    if (!diag) return mvalid > 0 ? MB * nvalid : 0;  // ragged rows run dense (results ignored)
    int n = 0;  // diagonal tiles always run the full-tile triangular variants
    for (int mi = 0; mi < MB; ++mi)
        for (int nj = 0; nj < NB; ++nj)
            if (nj >= wm * MB - wn * NB + mi) ++n;
    return n;
}
// cost of a tile per 32-row slab: the busiest SM sub-partition (warps w and w+4 share one)
int ws_tile_cost(int BN, int Nc, int ti, int tj, int diag, long* total_blocks) {
    int worst = 0;
    long tot = 0;
    for (int sp = 0; sp < 4; ++sp) {
        const int a = ws_warp_blocks(BN, Nc, ti, tj, diag, sp), b = ws_warp_blocks(BN, Nc, ti, tj, diag, sp + 4);
        worst = std::max(worst, a + b);
        tot += a + b;
    }
    if (total_blocks) *total_blocks = tot;
    return std::max(worst, 1);
}

struct WsPlan {
    int BN = 0, NT = 0, ntile = 0, nchunk = 0, chunk_rows = 0, nitems = 0, nctas = 0;
};

// Grid-chunk boundaries: uniform chunks of `rows`, except that the last one or two chunks' worth of
// rows is cut into eight short chunks, so the items scheduled last are small and the greedy
// assignment ends within ~1% of perfect balance.  Returns nchunk + 1 starts (multiples of 32).
std::vector<int> ws_chunk_starts(int Gpad, int rows) {
    std::vector<int> st;
    const int nfull = Gpad / rows;
    const int nmain = nfull >= 3 ? nfull - 1 : 0;  // small problems: no tail refinement
    for (int k = 0; k < nmain; ++k) st.push_back(k * rows);
    int g = nmain * rows;
    if (nmain == 0) {
        for (; g < Gpad; g += rows) st.push_back(g);
    } else {
        const int tail = Gpad - g;
        int piece = ((tail + 7) / 8 + 31) / 32 * 32;
        if (piece < 32) piece = 32;
        for (; g < Gpad; g += piece) st.push_back(g);
    }
    st.push_back(Gpad);
    return st;
}

// Layout of the wsyrk work: tiles x grid chunks (x batch), and the partial-slot count.
void ws_shape(int num_sms, int Nc, int Gpad, int B, bool sym, WsPlan& p) {
    p.BN = pick_bn(Nc);
    p.NT = (Nc + p.BN - 1) / p.BN;
    p.ntile = sym ? p.NT * (p.NT + 1) / 2 : p.NT * p.NT;
    // items per CTA: >= 16 keeps greedy scheduling within a few percent of perfect balance; 64 makes the grid chunks
    // small enough (15 MB of ao rows at c5) that the ~4 chunks in flight stay in L2 while their 36 tile pairs re-read
    // them: measured DRAM traffic per launch 25.0 -> 11.0 GB (8.2 algorithmic) at unchanged kernel time
    // (profiles/r02/ws_items_sweep.log); the price is 0.9 GB more partial-tile workspace at c5
    static const long per_cta = getenv("QEXXC_WS_ITEMS") ? atol(getenv("QEXXC_WS_ITEMS")) : 64L;
    long rows = ((long)Gpad * p.ntile * B + per_cta * num_sms - 1) / (per_cta * num_sms);
    rows = ((rows + 255) / 256) * 256;
    if (rows < 256) rows = 256;
    if (rows > Gpad) rows = Gpad;
    p.chunk_rows = (int)rows;
    p.nchunk = (int)ws_chunk_starts(Gpad, p.chunk_rows).size() - 1;
    p.nitems = p.ntile * p.nchunk * B;
    p.nctas = std::min(num_sms, p.nitems);
}

// Build (or reuse) the static schedule for the current (Npad, Gpad, B, sym) and upload it.
int ws_schedule(qexxc_ctx* c, bool sym, WsPlan& plan, cudaStream_t st) {
    ws_shape(c->num_sms, c->Nc, c->Gpad, c->B, sym, plan);
    const int sl = sym ? 1 : 0;
    const long key = ((long)c->Gpad << 20) ^ ((long)c->Npad << 4) ^ (long)c->B;
    if (c->ws_key[sl] == key) return QEXXC_OK;
    const int BN = plan.BN, NT = plan.NT;
    std::vector<int> tcost(plan.ntile);
    std::vector<std::pair<int, int>> tiles(plan.ntile);
    {
        int t = 0;
        for (int ti = 0; ti < NT; ++ti)
            for (int tj = sym ? ti : 0; tj < NT; ++tj, ++t) {
                tiles[t] = {ti, tj};
                tcost[t] = ws_tile_cost(BN, c->Nc, ti, tj, sym && ti == tj, nullptr);
            }
    }
    // Chunk-major sequence of (chunk, batch, tile) items, heavier tiles first within a chunk.
    //  phase 1: each item goes whole to the least-loaded CTA, as long as that CTA stays within the average
    //           load -- CTAs therefore walk the sequence together and share the chunk's AO rows in L2;
    //  phase 2: the remaining tail of the sequence is split by grid rows (whole 32-row slabs, pieces of at
    //           least kMinSlabs): the least-loaded CTA is filled exactly to the average, again and again, so
    //           all CTAs finish within a fraction of a percent of each other.
    // Every piece owns one partial-tile slot; the slots of a tile are numbered in grid order.
    constexpr int kMinSlabs = 8;
    struct Piece { int b, t, g0, slabs; };
    std::vector<int> order(plan.ntile);
    for (int t = 0; t < plan.ntile; ++t) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return tcost[a] > tcost[b]; });
    const std::vector<int> starts_g = ws_chunk_starts(c->Gpad, plan.chunk_rows);
    std::vector<Piece> seq;
    seq.reserve(plan.nitems);
    double total = 0.0;
    for (int ch = 0; ch < plan.nchunk; ++ch)
        for (int b = 0; b < c->B; ++b)
            for (int oi = 0; oi < plan.ntile; ++oi) {
                const int t = order[oi], slabs = (starts_g[ch + 1] - starts_g[ch]) / BK;
                seq.push_back({b, t, starts_g[ch], slabs});
                total += (double)tcost[t] * slabs;
            }
    const double per = total / plan.nctas;
    std::vector<std::vector<Piece>> lists(plan.nctas);
    std::vector<double> load(plan.nctas, 0.0);
    size_t pos = 0;
    for (; pos < seq.size(); ++pos) {
        int best = 0;
        for (int k = 1; k < plan.nctas; ++k)
            if (load[k] < load[best]) best = k;
        const double cost = (double)tcost[seq[pos].t] * seq[pos].slabs;
        if (load[best] + cost > per) break;
        load[best] += cost;
        lists[best].push_back(seq[pos]);
    }
    for (; pos < seq.size(); ++pos) {
        Piece rest = seq[pos];
        while (rest.slabs > 0) {
            int k = 0;
            for (int q = 1; q < plan.nctas; ++q)
                if (load[q] < load[k]) k = q;
            // fill the least-loaded CTA up to the average; when even it has no room for a minimal piece
            // (coarse items on a small grid) the rest is dealt out in minimal pieces, least-loaded first
            long take = (long)std::floor((per - load[k]) / tcost[rest.t] + 0.5);
            if (take < kMinSlabs) take = kMinSlabs;
            if (take > rest.slabs || rest.slabs - take < kMinSlabs) take = rest.slabs;
            lists[k].push_back({rest.b, rest.t, rest.g0, (int)take});
            load[k] += (double)tcost[rest.t] * take;
            rest.g0 += (int)take * BK;
            rest.slabs -= (int)take;
        }
    }
    std::vector<WsItem> flat;
    flat.reserve(plan.nitems + plan.nctas);
    std::vector<int> piece_tile;
    std::vector<int> start(plan.nctas + 1, 0);
    for (int k = 0; k < plan.nctas; ++k) {
        start[k] = (int)flat.size();
        for (const Piece& pc : lists[k]) {
            WsItem im;
            im.b = pc.b;
            im.ti = tiles[pc.t].first;
            im.tj = tiles[pc.t].second;
            im.g0 = pc.g0;
            im.kb = pc.slabs;
            im.slot = 0;
            im.diag = (sym && im.ti == im.tj) ? 1 : 0;
            im.pad = 0;
            flat.push_back(im);
            piece_tile.push_back(pc.b * plan.ntile + pc.t);
        }
    }
    start[plan.nctas] = (int)flat.size();
    // slots: the pieces of one tile, numbered by first grid row (deterministic reduction order)
    const int ntb = plan.ntile * c->B;
    std::vector<int> slot_start(ntb + 1, 0);
    for (int id : piece_tile) slot_start[id + 1]++;
    for (int q = 0; q < ntb; ++q) slot_start[q + 1] += slot_start[q];
    {
        std::vector<int> idx(flat.size());
        for (size_t q = 0; q < flat.size(); ++q) idx[q] = (int)q;
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
            return piece_tile[a] != piece_tile[b] ? piece_tile[a] < piece_tile[b] : flat[a].g0 < flat[b].g0;
        });
        for (size_t q = 0; q < idx.size(); ++q) flat[idx[q]].slot = (int)q;  // sorted position == slot id
    }
    if (getenv("QEXXC_DEBUG")) {
        double mx = 0, sum = 0;
        for (double l : load) {
            mx = std::max(mx, l);
            sum += l;
        }
        fprintf(stderr, "[qexxc] wsyrk schedule: sym=%d BN=%d tiles=%d chunks=%d (rows %d) pieces=%zu ctas=%d imbalance=%.4f\n",
                (int)sym, BN, plan.ntile, plan.nchunk, plan.chunk_rows, flat.size(), plan.nctas,
                mx * plan.nctas / std::max(sum, 1.0));
    }
    start.insert(start.end(), slot_start.begin(), slot_start.end());  // device table: [nctas + 1 | ntile * B + 1]
    if (flat.size() * sizeof(WsItem) > c->ws_items_bytes || start.size() > c->ws_start_cap) {
        set_error("internal: wsyrk schedule exceeds its workspace (%zu items)", flat.size());
        return QEXXC_ERR_STATE;
    }
    // (re)built only when the problem shape changes; make sure no kernel still reads the old tables
    QX_CUDA(cudaStreamSynchronize(st));
    QX_CUDA(cudaMemcpy(c->ws_items[sl], flat.data(), flat.size() * sizeof(WsItem), cudaMemcpyHostToDevice));
    QX_CUDA(cudaMemcpy(c->ws_start[sl], start.data(), start.size() * sizeof(int), cudaMemcpyHostToDevice));
    c->ws_key[sl] = key;
    return QEXXC_OK;
}

}  // namespace

void wsyrk_workspace(int num_sms, int Nc, int GpadMax, int B, bool general, size_t* part_doubles,
                     size_t* item_bytes, size_t* start_ints) {
    WsPlan a, b;
    ws_shape(num_sms, Nc, GpadMax, B, true, a);
    // the schedule may split up to one item per CTA boundary
    size_t it = (size_t)a.nitems + 2 * num_sms, n = it * a.BN * a.BN, nt = (size_t)a.ntile * B;
    if (general) {
        ws_shape(num_sms, Nc, GpadMax, B, false, b);
        const size_t itb = (size_t)b.nitems + 2 * num_sms;
        n = std::max(n, itb * b.BN * b.BN);
        it = std::max(it, itb);
        nt = std::max(nt, (size_t)b.ntile * B);
    }
    *part_doubles = n;
    *item_bytes = it * sizeof(WsItem);
    *start_ints = (size_t)num_sms + 1 + nt + 1;
}

double rowquad_executed_flops(const qexxc_ctx* c, int tri) {
    const int BN = pick_bn(c->Nc), NB = BN / 16, NT = (c->Nc + BN - 1) / BN;
    double blocks = 0.0;  // DMMAs per 8-row block of grid points
    for (int nt = 0; nt < NT; ++nt) {
        const int nw = std::min(BN, c->Nc - nt * BN), kend = rq_kend(tri, c->Nc, BN, nt), nbv = nw >> 3;
        const int kd = tri ? std::min(kend, nt * BN / BK) : kend;
        for (int wn = 0; wn < 2; ++wn) {
            const int nvalid = clampi((nbv - wn + 1) >> 1, 0, NB);
            blocks += (double)kd * 8 * nvalid;  // dense slabs: 8 k4 steps x nvalid blocks
            if (nvalid == 0) continue;
            for (int kb = kd; kb < kend; ++kb)  // diagonal slabs run the full-tile triangular variants
                for (int kk = 0; kk < 8; ++kk) {
                    const int nbmin = 4 * ((kb * BK - nt * BN) / BK) + (kk >> 1);
                    for (int j = 0; j < NB; ++j)
                        if (2 * j + wn >= nbmin) blocks += 1.0;
                }
        }
    }
    return blocks * 512.0 * (c->Gpad / 8) * c->B;
}

double wsyrk_executed_flops(const qexxc_ctx* c, bool sym) {
    WsPlan p;
    ws_shape(c->num_sms, c->Nc, c->Gpad, c->B, sym, p);
    double fl = 0.0;
    for (int ti = 0; ti < p.NT; ++ti)
        for (int tj = sym ? ti : 0; tj < p.NT; ++tj) {
            long tot = 0;
            ws_tile_cost(p.BN, c->Nc, ti, tj, sym && ti == tj, &tot);
            fl += (double)tot * 512.0 * (c->Gpad / 4) * c->B;
        }
    return fl;
}

int launch_pack_mo(qexxc_ctx* c, const double* C, const double* occ, int nmo, double* L, double* sgn, int ldL,
                   cudaStream_t st) {
    dim3 grid((ldL + 127) / 128, c->Npad, c->B);
    pack_mo_kernel<<<grid, 128, 0, st>>>(C, occ, L, sgn, c->N, nmo, c->Npad, ldL);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_pad_sym(qexxc_ctx* c, const double* src, int mode, int tri, cudaStream_t st) {
    dim3 grid((c->Npad + 127) / 128, c->Npad, c->B);
    pad_sym_kernel<<<grid, 128, 0, st>>>(src, c->S, c->N, c->Npad, mode, tri);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

// MO form of the density (pyscf eval_rho2): L [B][Npad][ldL] = C sqrt|occ| (zero padded), sgn [B][ldL] = sign of
// the occupation (0 for padding); q[b][g] = sum_k sgn_k ((ao L)[g,k])^2 with nk compute columns.
int launch_rowquad_mo(qexxc_ctx* c, const double* L, int ldL, int nk, const double* sgn, double* q, long q_bstride,
                      cudaStream_t st) {
    if (i8_enabled(c)) return launch_rowquad_mo_i8(c, L, ldL, nk, sgn, q, q_bstride, st);
    const int Sc = round_up(nk > 0 ? nk : 1, kNBlock);
    const int BN = pick_bn(Sc);
    dim3 grid(c->Gpad / BM, c->B);
    const long ao_cs = (long)c->GpadMax * c->Npad, ao_bs = c->ao_shared ? 0 : ao_cs * c->C, L_bs = (long)c->Npad * ldL;
#define QX_RQM(BNV)                                                                              \
    do {                                                                                         \
        QX_TRY(set_smem(rowquad_kernel<BNV>, RowquadCfg<BNV>::SMEM));                            \
        rowquad_kernel<BNV><<<grid, NTHREADS, RowquadCfg<BNV>::SMEM, st>>>(                      \
            c->ao, L, q, c->Npad, c->Nc, ao_cs, ao_bs, L_bs, 0, q_bstride, 1, 0, 1.0, 0.0, 0.0,  \
            0.0, ldL, Sc, sgn, (int)grid.x, 0, nullptr, 0, 0u, 0u, nullptr);                     \
    } while (0)
    ProfScope prof(c, QEXXC_PROF_ROWQUAD, st);
    if (BN == 128) QX_RQM(128);
    else if (BN == 64) QX_RQM(64);
    else QX_RQM(32);
#undef QX_RQM
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

// Split the last partial wave of row tiles into per-column-tile blocks?  Unsplit it costs one full tile time;
// split it costs about (its share of the SMs) + (the heaviest column tile), in units of a tile time.
int rowquad_tail_tiles(const qexxc_ctx* c, int tri) {
    if (c->B != 1 || c->rq_part == nullptr || getenv("QEXXC_NO_TAIL_SPLIT")) return 0;
    const int BN = pick_bn(c->Nc), NT = (c->Nc + BN - 1) / BN, T = c->Gpad / BM;
    const int rem = T % c->num_sms;
    if (rem == 0 || NT < 2) return 0;
    const double heaviest = tri ? 2.0 / (NT + 1) : 1.0 / NT;
    return ((double)rem / c->num_sms + heaviest < 0.95) ? rem : 0;
}

// Pair mode pays when the 128-row ao tiles of all resident blocks no longer fit in L2 (wide N) and there are enough row
// tiles to keep every SM busy with half-tiles; masks: column tiles dealt heaviest-first to the lighter half.
static bool rowquad_pair_masks(const qexxc_ctx* c, int tri, int nbulk, unsigned* m0, unsigned* m1) {
    const int BN = pick_bn(c->Nc), NT = (c->Nc + BN - 1) / BN;
    if (c->B != 1 || c->rq_pair == nullptr || NT < 2 || NT > 32 || getenv("QEXXC_NO_PAIR")) return false;
    if ((long)c->Npad * BM * 8 * c->num_sms < (96L << 20) || nbulk < 2 * c->num_sms) return false;
    long w0 = 0, w1 = 0;
    *m0 = *m1 = 0;
    for (int nt = NT - 1; nt >= 0; --nt) {  // rq_kend is non-decreasing in nt: this is heaviest first
        const long w = rq_kend(tri, c->Nc, BN, nt) + 1;
        if (w0 <= w1) {
            *m0 |= 1u << nt;
            w0 += w;
        } else {
            *m1 |= 1u << nt;
            w1 += w;
        }
    }
    return true;
}

int launch_rowquad(qexxc_ctx* c, int ncomp, int tri, const double* fac4, double* q, long q_bstride,
                   long q_cstride, cudaStream_t st) {
    if (i8_enabled(c)) return launch_rowquad_i8(c, ncomp, tri, fac4, q, q_bstride, q_cstride, st);
    const int BN = pick_bn(c->Nc), NT = (c->Nc + BN - 1) / BN, T = c->Gpad / BM;
    const int ntail = rowquad_tail_tiles(c, tri), nbulk = T - ntail;
    unsigned m0 = 0, m1 = 0;
    const int npair = rowquad_pair_masks(c, tri, nbulk, &m0, &m1) ? 1 : 0;
    dim3 grid((npair ? 2 * nbulk : nbulk) + ntail * NT, c->B);
    const long ao_cs = (long)c->GpadMax * c->Npad, ao_bs = c->ao_shared ? 0 : ao_cs * c->C, S_bs = (long)c->Npad * c->Npad;
#define QX_RQ(BNV)                                                                               \
    do {                                                                                         \
        QX_TRY(set_smem(rowquad_kernel<BNV>, RowquadCfg<BNV>::SMEM));                            \
        rowquad_kernel<BNV><<<grid, NTHREADS, RowquadCfg<BNV>::SMEM, st>>>(                      \
            c->ao, c->S, q, c->Npad, c->Nc, ao_cs, ao_bs, S_bs, q_cstride, q_bstride, ncomp, tri, \
            (tri ? 2.0 : 1.0) * fac4[0], fac4[1], fac4[2], fac4[3], c->Npad, c->Nc, nullptr,     \
            nbulk, ntail, c->rq_part, npair, m0, m1, c->rq_pair);                                \
    } while (0)
    ProfScope prof(c, QEXXC_PROF_ROWQUAD, st);
    if (BN == 128) QX_RQ(128);
    else if (BN == 64) QX_RQ(64);
    else QX_RQ(32);
#undef QX_RQ
    QX_LAUNCH_CHECK(c);
    if (npair) {
        rowquad_pair_kernel<<<(unsigned)(((long)nbulk * BM + 255) / 256), 256, 0, st>>>(c->rq_pair, q, q_cstride, ncomp, nbulk);
        QX_LAUNCH_CHECK(c);
    }
    if (ntail > 0) {
        rowquad_tail_kernel<<<(ntail * BM + 255) / 256, 256, 0, st>>>(c->rq_part, q, q_cstride, ncomp, NT, nbulk, ntail);
        QX_LAUNCH_CHECK(c);
    }
    return QEXXC_OK;
}

int launch_wsyrk(qexxc_ctx* c, const double* s, long s_bstride, const double* Bsrc, double scale, int tadd,
                 double* out, long out_bstride, cudaStream_t st) {
    if (i8_enabled(c)) return launch_wsyrk_i8(c, s, s_bstride, Bsrc, scale, tadd, out, out_bstride, st);
    const bool sym = (Bsrc == nullptr);
    WsPlan plan;
    QX_TRY(ws_schedule(c, sym, plan, st));
    const long ao_cs = (long)c->GpadMax * c->Npad, ao_bs = c->ao_shared ? 0 : ao_cs * c->C;
    const double* Bp = sym ? c->ao : Bsrc;
    const long B_bs = sym ? ao_bs : (long)c->GpadMax * c->Npad;
    const WsItem* items = reinterpret_cast<const WsItem*>(c->ws_items[sym ? 1 : 0]);
    const int* starts = c->ws_start[sym ? 1 : 0];
#define QX_WS(BNV)                                                                               \
    do {                                                                                         \
        QX_TRY(set_smem(wsyrk_kernel<BNV>, WsyrkCfg<BNV>::SMEM));                                \
        wsyrk_kernel<BNV><<<plan.nctas, NTHREADS, WsyrkCfg<BNV>::SMEM, st>>>(                    \
            c->ao, Bp, s, c->part, c->Npad, c->Nc, items, starts, ao_bs, B_bs, s_bstride);  \
    } while (0)
    {
        ProfScope prof(c, QEXXC_PROF_WSYRK, st);
        if (plan.BN == 128) QX_WS(128);
        else if (plan.BN == 64) QX_WS(64);
        else QX_WS(32);
    }
#undef QX_WS
    QX_LAUNCH_CHECK(c);
    dim3 rgrid((c->N + 31) / 32, c->N, c->B);
    wsyrk_reduce_kernel<<<rgrid, 32 * RK, 0, st>>>(c->part, out, c->N, plan.BN, plan.NT, starts + plan.nctas + 1,
                                                  plan.ntile, sym ? 1 : 0, scale, tadd, out_bstride);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_build_aow(qexxc_ctx* c, const double* wv, long wv_bstride, long wv_cstride,
                     const double* fac4, cudaStream_t st) {
    if (i8_enabled(c) && c->B == 1) {  // INT8 wsyrk follows: aow and its column maxima in one pass
        const int rc = launch_build_aow_i8(c, wv, wv_cstride, fac4, st);
        if (rc != QEXXC_ERR_UNSUPPORTED) return rc;
    }
    const long total2 = (long)c->Gpad * c->Npad / 2;
    const long ao_cs = (long)c->GpadMax * c->Npad, ao_bs = c->ao_shared ? 0 : ao_cs * c->C;
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

// Stages 2 and 4 on the INT8 tensor cores: an EXACT integer split (Ozaki scheme) of the two FP64 contractions
//     rowquad:  q_c[g]  = fac_c * sum_ij ao_c[g,i] S[i,j] ao_0[g,j]                (numint_legacy.py:351-410, eval_rho)
//     wsyrk:    H[i,j]  = sum_g ao_0[g,i] s[g] ao_0[g,j]   or   sum_g ao_0[g,i] B[g,j] (numint_legacy.py:432-456, V_xc)
// The FP64 DMMA kernels of contract.cu run at 0.96-0.99 of the FP64 tensor pipe; this path does not use that pipe for
// the N^2 G work.  Both operands of a product are written in fixed point relative to a per-row / per-column power of
// two and cut into six balanced base-256 digits,
//
//   x[m,k] = 2^(ea[m]-46) * sum_s A^(s)[m,k] 256^(5-s),   y[n,k] = 2^(eb[n]-46) * sum_t B^(t)[n,k] 256^(5-t),   digits in [-128, 127]
//   (x y^T)[m,n] ~= 2^(ea[m]+eb[n]-52) * sum_{d=0..5} 256^(5-d) * sum_{s+t=d} (A^(s) B^(t)^T)[m,n]
//
// the 21 digit products with s + t <= 5 are exact INT8 x INT8 -> INT32 GEMMs on tcgen05 (kind::i8); the digits of one
// diagonal d share a TMEM accumulator, the six accumulators are recombined exactly in 64-bit integers and rounded to FP64
// once.  What is dropped (s + t >= 6) is below 2^-45 of (row maximum) x (column maximum).
//
// tcgen05.mma has a floor of ~105 cycles per instruction for N <= 128 (scripts/probe/i8_rate.cu), so the products of one
// A digit plane with ALL the B planes it meets are ONE instruction: the B planes of a k-chunk sit in shared memory as a
// single [6 x 64 rows] K-major tile and the accumulators of consecutive diagonals are adjacent TMEM column ranges, so
// A^(s) x [B^(0); ...; B^(5-s)] lands in accumulators s ... 5 (N = 64 (6 - s), split at 256).
//
// One pipeline serves both contractions (one CTA = one 128 x 64 output tile, 10 warps): warp 0 streams digit tiles with
// 1-D bulk copies (the slicing kernels write them in the 128-byte-swizzled K-major tile layout, so no tensor maps are
// needed), warp 1 issues the tcgen05.mma stream, warps 2..9 drain the accumulators (tcgen05.ld).
//   rowquad: M = 128 grid rows (row-scaled digits of ao_0), N = 64 columns of S, K = AO index; the epilogue row-dots the
//            recombined (ao_0 S) tile with the FP64 ao_c rows.
//   wsyrk:   M = 128 AO rows i, N = 64 AO columns j, K = grid index.  A = the SAME row-scaled planes as rowquad, read MN-major
//            (a [128 grid rows x 128 AO bytes] tile is the [AO x grid] operand transposed; exact integers relative to each
//            row's maximum); B = s[g] 2^ea[g] ao_0[g,j] sliced per call with fixed-point exponents per COLUMN and per block of
//            4096 grid rows (block floating point along K); the INT32 accumulators are drained into FP64 registers at every
//            block boundary, where the block's scale is applied.
//
// Policy: QEXXC_I8=1 forces this path, QEXXC_I8=0 the FP64 DMMA path; unset = this path when nao >= 256 (single-molecule
// contexts).  Cost: 12 bytes of digit planes per AO value of component 0.
#include "common.cuh"
#include "tc05.cuh"

namespace qexxc {
namespace {
using namespace tc05;

constexpr int ND = 6;            // digits per operand
constexpr int IM = 128;          // output rows per CTA
constexpr int IN = 64;           // output columns per CTA (6 accumulators x 64 columns = 384 of the 512 TMEM columns)
constexpr int KC = 128;          // k-chunk: 128 int8 = one 128-byte swizzled row
constexpr uint32_t ATILE = IM * KC;  // 16 KB
constexpr uint32_t BTILE = IN * KC;  // 8 KB
constexpr int NA = 6;            // A-tile ring depth = one slot per digit plane (slot index is a compile-time constant)
constexpr int NEW = 16;          // epilogue warps: 4 TMEM lane quadrants x 4 column groups
constexpr int CW = IN / (NEW / 4);  // columns of a tile per epilogue warp
constexpr int I8_THREADS = 64 + 32 * NEW;
constexpr int KDC = 32;          // wsyrk: k-chunks per exponent block (4096 grid rows between accumulator drains; INT32 is safe to 16384)

// x * scale rounded to an integer v (|v| <= 2^46), returned as w = v + 0x808080808080: byte k of w, XOR 0x80, is the
// balanced base-256 digit k of v (least significant first)
__device__ __forceinline__ void to_w48(double x, double scale, uint32_t& lo, uint32_t& hi) {
    const double d = fma(x, scale, 6755399441055744.0);  // 1.5 * 2^52: the integer sits in the low mantissa bits
    const unsigned long long w = (unsigned long long)__double_as_longlong(d) + (0x0000808080808080ull - 0x4338000000000000ull);
    lo = (uint32_t)w;
    hi = (uint32_t)(w >> 32);
}
// byte b (0..3) of four words -> one word
__device__ __forceinline__ uint32_t pack_byte(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, int b) {
    const uint32_t sel = (uint32_t)b | ((uint32_t)(4 + b) << 4);
    return __byte_perm(__byte_perm(a0, a1, sel), __byte_perm(a2, a3, sel), 0x5410);
}
// digit plane s (0 = most significant) of four consecutive values
__device__ __forceinline__ uint32_t plane_word(const uint32_t (&lo)[4], const uint32_t (&hi)[4], int s) {
    const int b = ND - 1 - s;
    const uint32_t w = b < 4 ? pack_byte(lo[0], lo[1], lo[2], lo[3], b) : pack_byte(hi[0], hi[1], hi[2], hi[3], b - 4);
    return w ^ 0x80808080u;
}
// four consecutive doubles with one 256-bit load that does not allocate in L1: a thread of the rowquad epilogue reads 128
// contiguous bytes of its own AO row, so narrower loads fetch every 32-byte sector twice and thrash the small L1 left
// beside 193 KB of shared memory
__device__ __forceinline__ void ldg256(const double* p, double (&v)[4], int wide = 1) {
    if (wide) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    } else {
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
    }
}
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

__device__ __forceinline__ uint32_t tile_off(int r, int c) {  // byte offset of (row r, k c) inside a swizzled tile
    return (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((((c >> 4) ^ (r & 7)) << 4) | (c & 15));
}

// ---- slicing kernels ----------------------------------------------------------------------------------------------
// Row-scaled planes of ao_0 (rowquad A operand): A8[s][row tile][k-chunk][16 KB tile]; sa[g] = exponent e, |ao[g,:]| < 2^e.
// One CTA per 128-row tile, one warp per grid row (16 rows per warp, one after the other); the NKC k-chunks of a row live in
// registers between the maximum and the digit pass.  The same pass takes the column maxima of the tile (cmax[tile][col],
// rounded up to float), which the column-scaled planes of wsyrk need: the AO tensor is read once for both.
template <int NKC>
__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ ao, int Npad, int nkc, long Gpad,
                                                         signed char* __restrict__ A8, float* __restrict__ sa, int NpadK,
                                                         float* __restrict__ cmax) {
    __shared__ __align__(16) uint32_t stage[8][ND][32];
    __shared__ unsigned int cm[NKC * KC];  // column maxima of the tile (non-negative floats order like their bit patterns)
    for (int c = threadIdx.x; c < NKC * KC; c += 256) cm[c] = 0u;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long tile = blockIdx.x, nt = Gpad >> 7;
    float cmx[NKC][4];
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc) cmx[kc][0] = cmx[kc][1] = cmx[kc][2] = cmx[kc][3] = 0.0f;
#pragma unroll 1
    for (int rr = 0; rr < 16; ++rr) {
        const int r = warp * 16 + rr;
        const long g = tile * 128 + r;
        const double* row = ao + g * Npad;
        double x[NKC][4];
        double mx = 0.0;
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
            const int c = kc * KC + lane * 4;
            if (c < Npad) {  // Npad is a multiple of 32: the four columns are inside or outside together
                ldg256(row + c, x[kc]);  // one 256-bit streaming load per lane: 1 KB per warp and k-chunk
            } else {
                x[kc][0] = x[kc][1] = x[kc][2] = x[kc][3] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double ax = fabs(x[kc][j]);
                mx = fmax(mx, ax);
                cmx[kc][j] = fmaxf(cmx[kc][j], __double2float_ru(ax));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        int e = 0;
        if (mx > 0.0) (void)frexp(mx, &e);  // mx = m 2^e, m in [0.5, 1)
        if (e < -900) e = -900;
        if (lane == 0) sa[g] = (float)e;
        const double scale = pow2(46 - e);
#pragma unroll
        for (int kc = 0; kc < NKC; ++kc) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) to_w48(x[kc][j], scale, lo[j], hi[j]);
#pragma unroll
            for (int s = 0; s < ND; ++s) stage[warp][s][(((lane >> 2) ^ (r & 7)) << 2) | (lane & 3)] = plane_word(lo, hi, s);
            __syncwarp();
            for (int idx = lane; idx < ND * 8; idx += 32) {
                const int s = idx >> 3, ch = idx & 7;
                const uint4 v = *reinterpret_cast<const uint4*>(&stage[warp][s][ch * 4]);
                signed char* t = A8 + (((long)s * nt + tile) * nkc + kc) * ATILE + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + ch * 16;
                *reinterpret_cast<uint4*>(t) = v;
            }
            __syncwarp();
        }
    }
#pragma unroll
    for (int kc = 0; kc < NKC; ++kc)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicMax(&cm[kc * KC + lane * 4 + j], __float_as_uint(cmx[kc][j]));
    __syncthreads();
    for (int c = threadIdx.x; c < NKC * KC; c += 256) cmax[tile * NpadK + c] = __uint_as_float(cm[c]);
}

// Column-scaled planes of S (rowquad B operand): B8[t][column tile (64)][k-chunk][8 KB tile], B[n = j][k = i] = S[i][j];
// sb[j] = 2^(eb[j]-52).  One block per column j; S is [nrow x ncol] (square for rowquad, [nao x n_occ] for the MO form).
__global__ void __launch_bounds__(256) slice_s_kernel(const double* __restrict__ S, int ldS, int nrow, int ncol, int nkc, int nct,
                                                      signed char* __restrict__ B8, double* __restrict__ sb) {
    const int j = blockIdx.x;
    __shared__ double red[256];
    double mx = 0.0;
    if (j < ncol)
        for (int i = threadIdx.x; i < nrow; i += blockDim.x) mx = fmax(mx, fabs(S[(long)i * ldS + j]));
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    mx = red[0];
    int e = 0;
    if (mx > 0.0) (void)frexp(mx, &e);
    if (e < -900) e = -900;
    if (threadIdx.x == 0) sb[j] = pow2(e - 52);
    const double scale = pow2(46 - e);
    const int ct = j / IN, r = j % IN;
    for (int i = threadIdx.x; i < nkc * KC; i += blockDim.x) {
        const double x = (j < ncol && i < nrow) ? S[(long)i * ldS + j] : 0.0;
        uint32_t lo, hi;
        to_w48(x, scale, lo, hi);
        const unsigned long long w = (((unsigned long long)hi << 32) | lo) ^ 0x0000808080808080ull;
        const int kc = i >> 7, c = i & 127;
#pragma unroll
        for (int t = 0; t < ND; ++t)
            B8[(((long)t * nct + ct) * nkc + kc) * BTILE + tile_off(r, c)] = (signed char)(w >> (8 * (ND - 1 - t)));
    }
}

// cmax[sub][col] >= max over the 128 grid rows of sub-block `sub` of |x[g][col]| 2^rexp[g] (rounded up to float)
__global__ void __launch_bounds__(256) colmax_kernel(const double* __restrict__ x, int ld, int NpadK, const float* __restrict__ rexp,
                                                     float* __restrict__ cmax) {
    const long g0 = (long)blockIdx.x * 128;
    __shared__ double rf[128];
    if (threadIdx.x < 128) rf[threadIdx.x] = rexp ? pow2((int)rexp[g0 + threadIdx.x]) : 1.0;
    __syncthreads();
    for (int c = 4 * threadIdx.x; c < NpadK; c += 1024) {
        double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
        if (c < ld) {
            const double* p = x + g0 * ld + c;
#pragma unroll 8
            for (int r = 0; r < 128; ++r) {
                const double2 u = *reinterpret_cast<const double2*>(p + (long)r * ld), v = *reinterpret_cast<const double2*>(p + (long)r * ld + 2);
                const double f = rf[r];
                m0 = fmax(m0, f * fabs(u.x)), m1 = fmax(m1, f * fabs(u.y)), m2 = fmax(m2, f * fabs(v.x)), m3 = fmax(m3, f * fabs(v.y));
            }
        }
        *reinterpret_cast<float4*>(cmax + (long)blockIdx.x * NpadK + c) =
            make_float4(__double2float_ru(m0), __double2float_ru(m1), __double2float_ru(m2), __double2float_ru(m3));
    }
}

// GGA: aow[g][n] = sum_c f_c wv[c][g] ao[c][g][n] (_scale_ao, numint_legacy.py:432-442) AND, in the same pass, the column
// maxima of aow 2^rexp[g] per 128-row sub-block (what colmax_kernel would compute from a second read of aow)
__global__ void __launch_bounds__(256) build_aow_cmax_kernel(const double* __restrict__ ao, long ao_cstride, const double* __restrict__ wv,
                                                             long wv_cstride, double f0, double f1, double f2, double f3, int ld,
                                                             int NpadK, const float* __restrict__ rexp, double* __restrict__ aow,
                                                             float* __restrict__ cmax) {
    const long g0 = (long)blockIdx.x * 128;
    __shared__ double rf[128], w4[4][128];
    if (threadIdx.x < 128) {
        const long g = g0 + threadIdx.x;
        rf[threadIdx.x] = pow2((int)rexp[g]);
        w4[0][threadIdx.x] = f0 * wv[g];
        w4[1][threadIdx.x] = f1 * wv[wv_cstride + g];
        w4[2][threadIdx.x] = f2 * wv[2 * wv_cstride + g];
        w4[3][threadIdx.x] = f3 * wv[3 * wv_cstride + g];
    }
    __syncthreads();
    for (int c = 4 * threadIdx.x; c < NpadK; c += 1024) {
        double m[4] = {0.0, 0.0, 0.0, 0.0};
        if (c < ld) {
#pragma unroll 4
            for (int r = 0; r < 128; ++r) {
                const double* p = ao + (g0 + r) * ld + c;
                double a0[4], a1[4], a2[4], a3[4], o[4];
                ldg256(p, a0);
                ldg256(p + ao_cstride, a1);
                ldg256(p + 2 * ao_cstride, a2);
                ldg256(p + 3 * ao_cstride, a3);
                const double w0 = w4[0][r], w1 = w4[1][r], w2 = w4[2][r], w3 = w4[3][r], f = rf[r];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    o[j] = w0 * a0[j] + w1 * a1[j] + w2 * a2[j] + w3 * a3[j];
                    m[j] = fmax(m[j], f * fabs(o[j]));
                }
                double* q = aow + (g0 + r) * ld + c;
                *reinterpret_cast<double2*>(q) = make_double2(o[0], o[1]);
                *reinterpret_cast<double2*>(q + 2) = make_double2(o[2], o[3]);
            }
        }
        *reinterpret_cast<float4*>(cmax + (long)blockIdx.x * NpadK + c) =
            make_float4(__double2float_ru(m[0]), __double2float_ru(m[1]), __double2float_ru(m[2]), __double2float_ru(m[3]));
    }
}

// Block exponents of the wsyrk B operand: eexp[blk][col] = e with |s[g] 2^rexp[g] x[g][col]| < 2^e for the (up to) 4096 grid
// rows of block blk.  Exact (up to float rounding of cmax) without weights; with weights it is the bound
// max_sub (max_{g in sub} |s[g]| 2^rexp[g]) * cmax[sub][col] over the block's 128-row sub-blocks (tight up to the variation of the
// row factor inside 128 consecutive grid points).
__global__ void __launch_bounds__(256) blk_exp_kernel(const float* __restrict__ cmax, const double* __restrict__ s, const float* __restrict__ rexp,
                                                      int ngc, int NpadK, int* __restrict__ eexp) {
    __shared__ double smax[KDC];
    const int blk = blockIdx.x, gc0 = blk * KDC, nsub = min(KDC, ngc - gc0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int sub = warp; sub < nsub; sub += 8) {
        double d = 1.0;
        if (s) {
            const long g = ((long)(gc0 + sub) * 128) + lane * 4;
            d = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) d = fmax(d, fabs(s[g + k]) * (rexp ? pow2((int)rexp[g + k]) : 1.0));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
        }
        if (lane == 0) smax[sub] = d * (1.0 + 1e-6);  // slack for the roundings of the products formed later
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NpadK; c += 256) {
        double m = 0.0;
        for (int sub = 0; sub < nsub; ++sub) m = fmax(m, smax[sub] * (double)cmax[(long)(gc0 + sub) * NpadK + c]);
        int e = 0;
        if (m > 0.0) (void)frexp(m, &e);
        e = max(-900, min(900, e));  // inf / denormal bounds: keep the shifts finite
        eexp[(long)blk * NpadK + c] = e;
    }
}

// Column-scaled planes of x[g][col] (* s[g]) (* 2^rexp[g]) for wsyrk, K = grid index:  P[plane][g-chunk][column tile (64)][8 KB tile],
// tile row = column, tile byte = grid row inside the 128-row chunk.  One CTA per (g-chunk, 128 columns); a thread owns
// one column and 16 consecutive grid rows = one 16-byte chunk per plane, staged in shared memory in the final layout.
__global__ void __launch_bounds__(256) slice_cols_kernel(const double* __restrict__ x, int ld, const double* __restrict__ s,
                                                         const float* __restrict__ rexp, const int* __restrict__ eexp, int NpadK, int njt, long plane_stride,
                                                         signed char* __restrict__ P) {
    extern __shared__ __align__(1024) unsigned char stage[];  // [ND][128 columns][128 bytes]
    const int gc = blockIdx.x, cg = blockIdx.y, blk = gc / KDC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long g0 = (long)gc * 128 + 16 * warp;
    double sv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) sv[k] = (s ? s[g0 + k] : 1.0) * (rexp ? pow2((int)rexp[g0 + k]) : 1.0);
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
        const int r = q * 32 + lane, col = cg * 128 + r;
        const double scale = pow2(46 - eexp[(long)blk * NpadK + col]);
        double v[16];
        if (col < ld) {
            const double* p = x + g0 * ld + col;
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = p[(long)k * ld];
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) v[k] = 0.0;
        }
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) to_w48(v[k] * sv[k], scale, lo[k], hi[k]);
#pragma unroll
        for (int pl = 0; pl < ND; ++pl) {
            uint4 w;
            {
                const uint32_t l0[4] = {lo[0], lo[1], lo[2], lo[3]}, h0[4] = {hi[0], hi[1], hi[2], hi[3]};
                w.x = plane_word(l0, h0, pl);
                const uint32_t l1[4] = {lo[4], lo[5], lo[6], lo[7]}, h1[4] = {hi[4], hi[5], hi[6], hi[7]};
                w.y = plane_word(l1, h1, pl);
                const uint32_t l2[4] = {lo[8], lo[9], lo[10], lo[11]}, h2[4] = {hi[8], hi[9], hi[10], hi[11]};
                w.z = plane_word(l2, h2, pl);
                const uint32_t l3[4] = {lo[12], lo[13], lo[14], lo[15]}, h3[4] = {hi[12], hi[13], hi[14], hi[15]};
                w.w = plane_word(l3, h3, pl);
            }
            *reinterpret_cast<uint4*>(stage + pl * 16384 + (r >> 3) * 1024 + (r & 7) * 128 + ((warp ^ (r & 7)) << 4)) = w;
        }
    }
    __syncthreads();
    // 128 columns = two consecutive 8 KB column tiles = 16 KB contiguous per plane
    for (int pl = 0; pl < ND; ++pl) {
        uint4* dst = reinterpret_cast<uint4*>(P + pl * plane_stride + ((long)gc * njt + cg * 2) * BTILE);
        const uint4* src = reinterpret_cast<const uint4*>(stage + pl * 16384);
        for (int i = threadIdx.x; i < 1024; i += 256) dst[i] = src[i];
    }
}

// ---- the shared pipeline --------------------------------------------------------------------------------------------
struct Pipe {
    unsigned char *As, *Bs;
    uint64_t *fullB, *emptyB, *fullA, *emptyA, *tfull, *tempty;
};

// A "schedule" names the output units of a CTA (column tiles of a row tile / exponent blocks of an output tile) and,
// per unit, the k-chunk range and where its digit tiles are.
struct RqSched {
    static constexpr bool AMN = false;
    const signed char *A, *B;
    long a_plane, b_plane, tile;
    int nkc, nct, Nc, tri, P, p;  // this CTA takes the column tiles ct = p, p + P, ... of its row tile
    __device__ int nunits() const { return (nct - p + P - 1) / P; }
    __device__ int ct_of(int u) const { return u * P + p; }
    __device__ int kbeg(int) const { return 0; }
    __device__ int kend(int u) const {
        if (!tri) return (Nc + KC - 1) / KC;
        const int last = min(Nc, (ct_of(u) + 1) * IN);  // S is upper triangular: rows i <= last column only
        return (last + KC - 1) / KC;
    }
    // 32-byte k-steps of chunk kc that can hold non-zero products: an even column tile ct = 2m ends at column 128m + 63,
    // so with the upper-triangular S only rows i < 128m + 64 of its last k-chunk (kc = m) contribute
    __device__ int ksteps(int u, int kc) const { return (tri && !(ct_of(u) & 1) && kc == (ct_of(u) >> 1)) ? 2 : 4; }
    __device__ long aoff(int, int kc) const { return (tile * nkc + kc) * (long)ATILE; }
    __device__ long boff(int u, int kc) const { return ((long)ct_of(u) * nkc + kc) * (long)BTILE; }
};
// AROWS: the A operand is the ROW-scaled plane set of rowquad (tile [128 grid rows x 128 AO bytes] of row tile gc, k-chunk it),
// read MN-major (M = AO index along the 128-byte rows, K = grid row); otherwise the column-scaled planes T, K-major.
template <bool AROWS>
struct WsSched {
    static constexpr bool AMN = AROWS;
    const signed char *A, *B;
    long a_plane, b_plane;
    int it, jt, njt, ngc, blk0, blk1, nkc;
    __device__ int nunits() const { return blk1 - blk0; }
    __device__ int kbeg(int u) const { return (blk0 + u) * KDC; }
    __device__ int kend(int u) const { return min(ngc, (blk0 + u + 1) * KDC); }
    __device__ int ksteps(int, int) const { return 4; }
    __device__ long aoff(int, int gc) const {
        return AROWS ? ((long)gc * nkc + it) * (long)ATILE : ((long)gc * njt + 2 * it) * (long)BTILE;
    }
    __device__ long boff(int, int gc) const { return ((long)gc * njt + jt) * (long)BTILE; }
};

__device__ __forceinline__ Pipe pipe_setup(unsigned char* smem_raw, uint32_t** tslot) {
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    Pipe p;
    p.Bs = base;                   // [2][ND][BTILE]
    p.As = base + 2 * ND * BTILE;  // [NA][ATILE]
    uint64_t* bars = (uint64_t*)(p.As + NA * ATILE);
    p.fullB = bars, p.emptyB = bars + 2, p.fullA = bars + 4, p.emptyA = bars + 4 + NA, p.tfull = bars + 4 + 2 * NA,
    p.tempty = bars + 5 + 2 * NA;
    *tslot = (uint32_t*)(bars + 6 + 2 * NA);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(p.fullB + i, 1);
            mbar_init(p.emptyB + i, 1);
        }
        for (int i = 0; i < NA; ++i) {
            mbar_init(p.fullA + i, 1);
            mbar_init(p.emptyA + i, 1);
        }
        mbar_init(p.tfull, 1);
        mbar_init(p.tempty, NEW);  // one arrival per epilogue warp
        mbar_fence_init();
    }
    return p;
}

template <class S>
__device__ __forceinline__ void i8_produce(const S& sc, const Pipe& p) {
    int itB = 0;
    const int nu = sc.nunits();
    for (int u = 0; u < nu; ++u) {
        const int k1 = sc.kend(u);
        for (int kc = sc.kbeg(u); kc < k1; ++kc, ++itB) {
            const int kb = itB & 1;
            mbar_wait(p.emptyB + kb, ((itB >> 1) & 1) ^ 1);
            const signed char* bsrc = sc.B + sc.boff(u, kc);
            mbar_expect_tx(p.fullB + kb, ND * BTILE);
            for (int t = 0; t < ND; ++t) bulk_g2s(p.Bs + ((size_t)kb * ND + t) * BTILE, bsrc + t * sc.b_plane, BTILE, p.fullB + kb);
            const signed char* asrc = sc.A + sc.aoff(u, kc);
            for (int s = 0; s < ND; ++s) {
                mbar_wait(p.emptyA + s, (itB & 1) ^ 1);
                mbar_expect_tx(p.fullA + s, ATILE);
                bulk_g2s(p.As + (size_t)s * ATILE, asrc + s * sc.a_plane, ATILE, p.fullA + s);
            }
        }
    }
}

// The MMA stream, executed by one whole (converged) warp so that every operand is warp-uniform; the instructions themselves
// are issued by one elected lane.  Fully unrolled over the six A planes and the four 32-byte k-steps of a k-chunk.
template <class S>
__device__ __forceinline__ void i8_issue(const S& sc, const Pipe& p, uint32_t tm) {
    constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, descriptor version 1, SWIZZLE_128B
    const uint32_t sA = smem_u32(p.As), sB = smem_u32(p.Bs);
    const uint32_t alo0 = ((sA >> 4) & 0x3FFFu) | (1u << 16), blo0 = ((sB >> 4) & 0x3FFFu) | (1u << 16);
    int itB = 0;
    const int nu = sc.nunits();
    for (int u = 0; u < nu; ++u) {
        mbar_wait(p.tempty, (u & 1) ^ 1);  // the epilogue has drained the accumulators of the previous unit
        tc_fence_after();
        const int k0 = sc.kbeg(u), k1 = sc.kend(u);
        for (int kc = k0; kc < k1; ++kc, ++itB) {
            const int kb = itB & 1;
            mbar_wait(p.fullB + kb, (itB >> 1) & 1);
            const uint32_t blo = blo0 + (uint32_t)kb * (ND * BTILE >> 4);
            const uint32_t first = (uint32_t)(kc != k0);
            const int nks = sc.ksteps(u, kc);
#pragma unroll
            for (int s = 0; s < ND; ++s) {  // A ring slot == plane (NA == ND)
                mbar_wait(p.fullA + s, itB & 1);
                tc_fence_after();
                // A^(s) x [B^(0); ...; B^(5-s)] -> accumulators s ... 5 (adjacent TMEM columns), N <= 256 per instruction
                constexpr int dummy = 0;
                (void)dummy;
                const int ntot = (ND - s) * IN;
                const int n0 = ntot > 256 ? 256 : ntot;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k >= nks) break;
                        const uint32_t acc = (s != 0 || k != 0) ? 1u : first;
                        // K-major A: a k-step is 32 bytes along the 128-byte rows; MN-major A: 32 rows = four 1024-byte atoms
                        const uint32_t alo = alo0 + (uint32_t)s * (ATILE >> 4) + (uint32_t)k * (S::AMN ? 256u : 2u);
                        mma_i8_lohi(tm + (uint32_t)s * IN, alo, blo + (uint32_t)k * 2u, DESC_HI, idesc_i8(IM, n0, S::AMN ? 1 : 0), acc);
                        if (ntot > 256)
                            mma_i8_lohi(tm + (uint32_t)s * IN + 256, alo, blo + (4 * BTILE >> 4) + (uint32_t)k * 2u, DESC_HI,
                                        idesc_i8(IM, ntot - 256, S::AMN ? 1 : 0), acc);
                    }
                    mma_commit(p.emptyA + s);  // the slot is free once these MMAs have read it
                    if (s == ND - 1) mma_commit(p.emptyB + kb);
                }
                __syncwarp();
            }
        }
        if (elect_one()) mma_commit(p.tfull);
        __syncwarp();
    }
}

// eight columns (c0 ...) of the six diagonal accumulators of this thread's TMEM lane -> T[j] = sum_d P_d 256^(5-d), exact in
// two 64-bit halves, rounded to FP64 once
__device__ __forceinline__ void drain8(uint32_t tm, int qd, int c0, double (&T)[8]) {
    uint32_t v[ND][8];
#pragma unroll
    for (int d = 0; d < ND; ++d) tmem_ld8(tmem_addr(tm, 32 * qd, d * IN + c0), v[d]);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const long long hi = ((long long)(int)v[0][j] << 16) + ((long long)(int)v[1][j] << 8) + (long long)(int)v[2][j];
        const long long lo = ((long long)(int)v[3][j] << 16) + ((long long)(int)v[4][j] << 8) + (long long)(int)v[5][j];
        T[j] = fma((double)hi, 16777216.0, (double)lo);
    }
}

// ---- rowquad ----------------------------------------------------------------------------------------------------------
struct RqArgs {
    RqSched sc;        // tile filled in by the CTA
    const float* sa;   // row exponents
    const double* sb;  // column scales 2^(eb-52)
    const double* ao;  // FP64 ao rows for the row-dot epilogue
    const double* sgn; // MO form (non-null): q[g] = sum_k sgn[k] ((ao_0 L)[g,k])^2, no row-dot
    double* q;
    long ao_cstride, q_cstride;
    int Npad, ncomp, wide;
    double f[4];
};

__global__ void __launch_bounds__(I8_THREADS, 1) rowquad_i8_kernel(const RqArgs a) {
    extern __shared__ unsigned char smem_raw[];
    uint32_t* tslot;
    const Pipe p = pipe_setup(smem_raw, &tslot);
    const int warp = warp_uniform_idx(), lane = threadIdx.x & 31;
    RqSched sc = a.sc;
    sc.tile = blockIdx.x / sc.P;  // the P CTAs of a row tile are neighbours: they run together and share its A planes in L2
    sc.p = blockIdx.x % sc.P;
    if (warp == 1) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tslot;

    if (warp == 0) {
        if (lane == 0) i8_produce(sc, p);
    } else if (warp == 1) {
        i8_issue(sc, p, tm);
    } else {
        // epilogue: 16 warps; thread = one grid row x CW of the tile's columns
        const int qd = warp & 3;         // TMEM lane quadrant of this warp
        const int cq = (warp - 2) >> 2;  // columns [CW cq, CW cq + CW) of every 64-column tile
        const int r = qd * 32 + lane;
        const long g = sc.tile * IM + r;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        const int nu = sc.nunits();
        for (int u = 0; u < nu; ++u) {
            const int col = sc.ct_of(u) * IN + CW * cq;
            const bool live = col < a.Npad;  // Npad is a multiple of 32 and col of 16; digits of columns >= Nc are zero
            // the FP64 ao_0 values of this thread's row are fetched before the wait, so that after the drain the row-dot is
            // arithmetic only and the warp is back in time for the next (possibly short) column tile
            double a0[CW / 4][4];
            if (live && !a.sgn) {
                const double* ap = a.ao + g * a.Npad + col;
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) ldg256(ap + 4 * j, a0[j], a.wide);
            }
            mbar_wait(p.tfull, u & 1);
            __syncwarp();
            tc_fence_after();
            double T[CW];
#pragma unroll
            for (int b = 0; b < CW / 8; ++b) {
                double t8[8];
                drain8(tm, qd, CW * cq + 8 * b, t8);
#pragma unroll
                for (int j = 0; j < 8; ++j) T[8 * b + j] = t8[j];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p.tempty);  // the accumulators are free: the next column tile's MMAs overlap the row-dot
            if (live && a.sgn) {  // MO form: signed sum of squares of the (ao_0 L) row
                double s0 = 0.0;
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    const double t = T[j] * a.sb[col + j];
                    s0 = fma(a.sgn[col + j] * t, t, s0);
                }
                acc[0] += s0;
            } else if (live) {
#pragma unroll
                for (int j = 0; j < CW; ++j) T[j] *= a.sb[col + j];
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                    s0 = fma(T[4 * j], a0[j][0], s0);
                    s1 = fma(T[4 * j + 1], a0[j][1], s1);
                    s0 = fma(T[4 * j + 2], a0[j][2], s0);
                    s1 = fma(T[4 * j + 3], a0[j][3], s1);
                }
                acc[0] += s0 + s1;
#pragma unroll
                for (int c = 1; c < 4; ++c)
                    if (c < a.ncomp) {
                        const double* ap = a.ao + (long)c * a.ao_cstride + g * a.Npad + col;
                        double v[CW / 4][4];
#pragma unroll
                        for (int j = 0; j < CW / 4; ++j) ldg256(ap + 4 * j, v[j], a.wide);
                        s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int j = 0; j < CW / 4; ++j) {
                            s0 = fma(T[4 * j], v[j][0], s0);
                            s1 = fma(T[4 * j + 1], v[j][1], s1);
                            s0 = fma(T[4 * j + 2], v[j][2], s0);
                            s1 = fma(T[4 * j + 3], v[j][3], s1);
                        }
                        acc[c] += s0 + s1;
                    }
            }
        }
        // the column groups of a row: fixed-order sum through shared memory (every tile has been consumed by now)
        double* part = reinterpret_cast<double*>(p.Bs);
        if (cq != 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) part[((cq - 1) * 4 + c) * IM + r] = acc[c];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * NEW) : "memory");
        if (cq == 0) {
            const double r1 = pow2((int)a.sa[g]), rs = a.sgn ? r1 * r1 : r1;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < a.ncomp) {
                    double t = acc[c];
#pragma unroll
                    for (int k = 0; k < NEW / 4 - 1; ++k) t += part[(k * 4 + c) * IM + r];
                    // P > 1: partial sums of this CTA's column tiles, [p][c][g], added up by rowquad_i8_sum_kernel
                    a.q[((long)sc.p * (sc.P > 1 ? 4 : 0) + c) * a.q_cstride + g] = a.f[c] * rs * t;
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_free(tm, 512);
}

// q[c][g] = sum_p part[p][c][g] (fixed order)
__global__ void __launch_bounds__(256) rowquad_i8_sum_kernel(const double* __restrict__ part, long pstride, int P, double* __restrict__ q,
                                                             long q_cstride, int ncomp, long n) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    for (int c = 0; c < ncomp; ++c) {
        double t = 0.0;
        for (int p = 0; p < P; ++p) t += part[((long)p * 4 + c) * pstride + g];
        q[(long)c * q_cstride + g] = t;
    }
}

// ---- wsyrk ------------------------------------------------------------------------------------------------------------
struct WsArgs {
    const signed char *A, *B;   // unweighted / weighted column-scaled planes
    long plane_stride, a_plane;
    int nkc;
    const int *eA, *eB;         // block exponents [nblk][NpadK]; eA == nullptr: the A digits are plain integers (row-scaled planes)
    double* part;               // [nsplit][ntile][128 x 64]
    int NpadK, njt, njtL, ngc, nblk, ntile, nsplit, sym;  // njt: column tiles that hold data (tile enumeration); njtL: pitch of the plane layout
};

__device__ __forceinline__ void ws_tile(const WsArgs& a, int t, int& it, int& jt) {
    if (!a.sym) {
        it = t / a.njt;
        jt = t % a.njt;
        return;
    }
    it = 0;
    while (t >= a.njt - 2 * it) {  // row tile it holds the column tiles jt >= 2 it (on or above the diagonal)
        t -= a.njt - 2 * it;
        ++it;
    }
    jt = 2 * it + t;
}

template <bool AROWS>
__global__ void __launch_bounds__(I8_THREADS, 1) wsyrk_i8_kernel(const WsArgs a) {
    extern __shared__ unsigned char smem_raw[];
    uint32_t* tslot;
    const Pipe p = pipe_setup(smem_raw, &tslot);
    const int warp = warp_uniform_idx(), lane = threadIdx.x & 31;
    const int t = blockIdx.x % a.ntile, sp = blockIdx.x / a.ntile;
    WsSched<AROWS> sc;
    sc.A = a.A, sc.B = a.B, sc.a_plane = a.a_plane, sc.b_plane = a.plane_stride, sc.nkc = a.nkc;
    ws_tile(a, t, sc.it, sc.jt);
    sc.njt = a.njtL, sc.ngc = a.ngc;
    sc.blk0 = (int)((long)a.nblk * sp / a.nsplit), sc.blk1 = (int)((long)a.nblk * (sp + 1) / a.nsplit);
    if (warp == 1) tmem_alloc(tslot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tslot;

    if (warp == 0) {
        if (lane == 0) i8_produce(sc, p);
    } else if (warp == 1) {
        i8_issue(sc, p, tm);
    } else {
        // epilogue: thread = one output row i x CW of the tile's 64 columns, FP64 accumulators across the exponent blocks
        const int qd = warp & 3, cq = (warp - 2) >> 2;
        const int r = qd * 32 + lane;
        const int i = sc.it * IM + r, j0 = sc.jt * IN + CW * cq;
        double acc[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) acc[j] = 0.0;
        for (int u = 0; u < sc.blk1 - sc.blk0; ++u) {
            const long eo = (long)(sc.blk0 + u) * a.NpadK;
            const double si = a.eA ? pow2(a.eA[eo + i] - 26) : 1.4901161193847656e-08;  // row-scaled integer A digits: 2^-26, the row's 2^ea is in B
            int eb[CW];  // fetched before the wait: the accumulators are held for as short a time as possible
            {
                const int4* ebp = reinterpret_cast<const int4*>(a.eB + eo + j0);
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                    const int4 v = ebp[j];
                    eb[4 * j] = v.x, eb[4 * j + 1] = v.y, eb[4 * j + 2] = v.z, eb[4 * j + 3] = v.w;
                }
            }
            mbar_wait(p.tfull, u & 1);
            __syncwarp();
            tc_fence_after();
#pragma unroll
            for (int b = 0; b < CW / 8; ++b) {
                double t8[8];
                drain8(tm, qd, CW * cq + 8 * b, t8);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[8 * b + j] = fma(t8[j] * si, pow2(eb[8 * b + j] - 26), acc[8 * b + j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p.tempty);
        }
        double2* out = reinterpret_cast<double2*>(a.part + ((long)sp * a.ntile + t) * (IM * IN) + r * IN + CW * cq);
#pragma unroll
        for (int j = 0; j < CW / 2; ++j) out[j] = make_double2(acc[2 * j], acc[2 * j + 1]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_free(tm, 512);
}

// out[i][j] = scale * (H[i][j] + (tadd ? H[j][i] : 0)),  H = sum over the grid splits of the partial tiles (fixed order)
__global__ void __launch_bounds__(256) wsyrk_i8_reduce_kernel(const double* __restrict__ part, double* __restrict__ out, int N, int njt,
                                                              int ntile, int nsplit, int sym, double scale, int tadd) {
    const int j = blockIdx.x * 32 + (threadIdx.x & 31), i = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (i >= N || j >= N) return;
    auto H = [&](int r, int c) {
        const int it = r / IM, jt = c / IN;
        const int t = sym ? it * njt - it * (it - 1) + (jt - 2 * it) : it * njt + jt;
        const double* P = part + (long)t * (IM * IN) + (r - it * IM) * IN + (c - jt * IN);
        double h = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) h += P[(long)sp * ntile * (IM * IN)];
        return h;
    };
    if (sym) {
        if (j < i) return;
        const double v = scale * (tadd ? 2.0 : 1.0) * H(i, j);
        out[(long)i * N + j] = v;
        out[(long)j * N + i] = v;
    } else {
        out[(long)i * N + j] = scale * (H(i, j) + (tadd ? H(j, i) : 0.0));
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
struct I8Ws {
    signed char *A = nullptr, *Bs = nullptr;  // rowquad: row-scaled ao planes, S planes
    float* sa = nullptr;
    double* sb = nullptr;
    double* qpart = nullptr;                  // rowquad: partial sums [4 column groups][4 components][GpadMax]
    signed char *T = nullptr, *W = nullptr;   // wsyrk: column-scaled planes of ao_0 (per geometry) and of the weighted operand (per call)
    float *cmax = nullptr, *cmaxW = nullptr;  // 128-row column maxima of ao_0 / of a general B operand
    int *eA = nullptr, *eB = nullptr;
    int nkc = 0, NpadK = 0, njt = 0, ngcMax = 0, nblkMax = 0;
    const double* cmaxW_of = nullptr;         // cmaxW currently holds the maxima of this matrix (set by launch_build_aow_i8)
    long plane_stride = 0;
};

size_t i8_smem() { return 2 * ND * BTILE + NA * ATILE + 32 * 8 + 1024; }

int i8_alloc(qexxc_ctx* c) {
    if (c->i8ws) return QEXXC_OK;
    I8Ws* w = new I8Ws();
    w->nkc = (c->Npad + KC - 1) / KC;
    w->NpadK = w->nkc * KC;
    w->njt = w->NpadK / IN;
    w->ngcMax = c->GpadMax / 128;
    w->nblkMax = (w->ngcMax + KDC - 1) / KDC;
    w->plane_stride = (long)w->ngcMax * w->njt * BTILE;
    const size_t planes = (size_t)ND * c->GpadMax * w->NpadK;
    const size_t sbytes = (size_t)ND * w->njt * w->nkc * BTILE;
    auto A = [&](void** p, size_t bytes) {
        if (cudaMalloc(p, bytes) != cudaSuccess) return false;
        c->allocs.push_back(*p);
        c->bytes += bytes;
        return true;
    };
    const bool want_T = getenv("QEXXC_I8_T") && atoi(getenv("QEXXC_I8_T")) != 0;  // A/B runs of the column-scaled A planes
    bool ok = A((void**)&w->A, planes) && (!want_T || A((void**)&w->T, planes)) && A((void**)&w->W, planes) && A((void**)&w->Bs, sbytes) &&
              A((void**)&w->sa, sizeof(float) * c->GpadMax) && A((void**)&w->qpart, sizeof(double) * 16 * (size_t)c->GpadMax) && A((void**)&w->sb, sizeof(double) * w->NpadK) &&
              A((void**)&w->cmax, sizeof(float) * (size_t)w->ngcMax * w->NpadK) &&
              A((void**)&w->eA, sizeof(int) * (size_t)w->nblkMax * w->NpadK) && A((void**)&w->eB, sizeof(int) * (size_t)w->nblkMax * w->NpadK);
    if (ok && c->C == 4) ok = A((void**)&w->cmaxW, sizeof(float) * (size_t)w->ngcMax * w->NpadK);
    c->i8ws = w;  // the buffers are owned by c->allocs; the small struct is released in qexxc_destroy via i8_release
    if (!ok) {
        (void)cudaGetLastError();
        set_error("INT8 contraction workspace: cudaMalloc failed (%.1f GB of digit planes); set QEXXC_I8=0 for the FP64 path",
                  2.0 * planes / 1e9);
        return QEXXC_ERR_CUDA;
    }
    QX_CUDA(cudaFuncSetAttribute(rowquad_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8_smem()));
    QX_CUDA(cudaFuncSetAttribute(wsyrk_i8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8_smem()));
    QX_CUDA(cudaFuncSetAttribute(wsyrk_i8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)i8_smem()));
    QX_CUDA(cudaFuncSetAttribute(slice_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ND * 16384));
    return QEXXC_OK;
}

// digit planes that depend on the geometry only: row-scaled (rowquad) and column-scaled (wsyrk) planes of ao_0
int i8_prepare(qexxc_ctx* c, cudaStream_t st) {
    QX_TRY(i8_alloc(c));
    if (c->i8_valid) return QEXXC_OK;
    I8Ws* w = (I8Ws*)c->i8ws;
    ProfScope prof(c, QEXXC_PROF_SLICE, st);
    const unsigned nb = (unsigned)(c->Gpad / 128);
    switch (w->nkc) {
#define QX_SR(K)                                                                                                          \
    case K:                                                                                                               \
        slice_rows_kernel<K><<<nb, 256, 0, st>>>(c->ao, c->Npad, w->nkc, c->Gpad, w->A, w->sa, w->NpadK, w->cmax);      \
        break
        QX_SR(1); QX_SR(2); QX_SR(3); QX_SR(4); QX_SR(5); QX_SR(6); QX_SR(7); QX_SR(8);
        QX_SR(9); QX_SR(10); QX_SR(11); QX_SR(12); QX_SR(13); QX_SR(14); QX_SR(15); QX_SR(16);
#undef QX_SR
        default:
            set_error("INT8 contractions support nao <= 2048 (got %d)", c->N);
            return QEXXC_ERR_UNSUPPORTED;
    }
    QX_LAUNCH_CHECK(c);
    if (w->T) {  // QEXXC_I8_T=1: separate column-scaled planes for the wsyrk A operand (A/B runs; the default reads A8 MN-major)
        const int ngc = c->Gpad / 128, nblk = (ngc + KDC - 1) / KDC;
        blk_exp_kernel<<<nblk, 256, 0, st>>>(w->cmax, nullptr, nullptr, ngc, w->NpadK, w->eA);
        QX_LAUNCH_CHECK(c);
        slice_cols_kernel<<<dim3(ngc, w->NpadK / 128), 256, ND * 16384, st>>>(c->ao, c->Npad, nullptr, nullptr, w->eA, w->NpadK, w->njt,
                                                                             w->plane_stride, w->T);
        QX_LAUNCH_CHECK(c);
    }
    c->i8_valid = true;
    return QEXXC_OK;
}

}  // namespace

int i8_prepare_geometry(qexxc_ctx* c, cudaStream_t st) { return i8_prepare(c, st); }

// GGA: the weighted AO tensor of stage 4 and its 128-row column maxima in one pass (the general wsyrk that follows
// slices aow without reading it a second time for the maxima)
int launch_build_aow_i8(qexxc_ctx* c, const double* wv, long wv_cstride, const double* fac4, cudaStream_t st) {
    QX_TRY(i8_prepare(c, st));
    I8Ws* w = (I8Ws*)c->i8ws;
    if (w->T || !w->cmaxW) return QEXXC_ERR_UNSUPPORTED;  // A/B mode with column-scaled A planes: caller uses the plain kernel
    ProfScope prof(c, QEXXC_PROF_SLICE, st);
    build_aow_cmax_kernel<<<c->Gpad / 128, 256, 0, st>>>(c->ao, (long)c->GpadMax * c->Npad, wv, wv_cstride, fac4[0], fac4[1], fac4[2],
                                                        fac4[3], c->Npad, w->NpadK, w->sa, c->aow, w->cmaxW);
    QX_LAUNCH_CHECK(c);
    w->cmaxW_of = c->aow;
    return QEXXC_OK;
}
int i8_reserve(qexxc_ctx* c) { return i8_alloc(c); }

void i8_release(qexxc_ctx* c) {
    delete (I8Ws*)c->i8ws;
    c->i8ws = nullptr;
}

// INT8 operations (2 per MAC) of one launch: every scheduled (tile, k-chunk) pair runs the 21 digit products of a
// 128 x 64 x 128 block
double i8_executed_ops(const qexxc_ctx* c, int which, bool sym) {
    const double per_pair = 2.0 * 21.0 * IM * IN * KC;
    const int njt = (c->Nc + IN - 1) / IN, nit = (c->Nc + IM - 1) / IM;
    if (which == 0) {
        double pairs = 0.0;
        for (int ct = 0; ct < njt; ++ct) {
            pairs += sym ? (std::min(c->Nc, (ct + 1) * IN) + KC - 1) / KC : (c->Nc + KC - 1) / KC;
            if (sym && !(ct & 1) && (ct >> 1) < (std::min(c->Nc, (ct + 1) * IN) + KC - 1) / KC) pairs -= 0.5;  // half k-chunk on the diagonal
        }
        return per_pair * pairs * (c->Gpad / IM);
    }
    int ntile = 0;
    for (int it = 0; it < nit; ++it) ntile += sym ? std::max(0, njt - 2 * it) : njt;
    return per_pair * ntile * (c->Gpad / KC);
}

namespace {
// the rate probe: the same unrolled issue pattern as the contractions, N = 256 only, operands resident in shared memory
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < (int)(ATILE + 4 * BTILE) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x01010101u * (uint32_t)(i % 3);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    const int warp = warp_uniform_idx();
    if (warp == 0) tmem_alloc(&tslot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (warp == 1) {
        constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
        const uint32_t alo = ((smem_u32(base) >> 4) & 0x3FFFu) | (1u << 16), blo = ((smem_u32(base + ATILE) >> 4) & 0x3FFFu) | (1u << 16);
        uint32_t phase = 0;
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mma_i8_lohi(tm, alo + k * 2u, blo + k * 2u, DESC_HI, idesc_i8(IM, 256), 1u);
                    mma_i8_lohi(tm + 256, alo + k * 2u, blo + k * 2u, DESC_HI, idesc_i8(IM, 256), 1u);
                }
                if ((it & 31) == 31 || it == iters - 1) mma_commit(&bar);
            }
            __syncwarp();
            if ((it & 31) == 31 || it == iters - 1) {  // bound the number of MMAs in flight
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}
}  // namespace

int i8_peak_probe(int device, double* ops_per_second) {
    QX_CUDA(cudaSetDevice(device));
    int nsm = 0;
    QX_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    const int smem = ATILE + 4 * BTILE + 1024, iters = 20000;
    QX_CUDA(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    QX_CUDA(cudaEventCreate(&e0));
    QX_CUDA(cudaEventCreate(&e1));
    i8_peak_kernel<<<nsm, 128, smem>>>(2000);
    QX_CUDA(cudaEventRecord(e0));
    i8_peak_kernel<<<nsm, 128, smem>>>(iters);
    QX_CUDA(cudaEventRecord(e1));
    QX_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    QX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_second = 2.0 * nsm * (double)iters * 8.0 * IM * 256.0 * 32.0 / (ms * 1e-3);
    return QEXXC_OK;
}

bool i8_enabled(const qexxc_ctx* c) {
    // one AO tensor per context: a single molecule, or nset density matrices over a SHARED AO tensor (looped over the sets)
    if ((c->B != 1 && !c->ao_shared) || c->Npad > 2048) return false;
    const char* e = getenv("QEXXC_I8");
    if (e) return atoi(e) != 0;
    return c->N >= 256;
}

// B operand = the [nrow x ncol] matrix Smat (ld ldS): the padded S of rowquad, or L = C sqrt|occ| of the MO form (sgn != nullptr)
static int rowquad_i8_common(qexxc_ctx* c, const double* Smat, int ldS, int nrow, int ncol, int ncomp, int tri, const double* fac4,
                             const double* sgn, int ldsgn, double* q, long q_cstride, cudaStream_t st) {
    QX_TRY(i8_prepare(c, st));
    I8Ws* w = (I8Ws*)c->i8ws;
    const int nct = (ncol + IN - 1) / IN;  // column tiles that hold data
    {
        ProfScope prof(c, QEXXC_PROF_SLICE, st);
        slice_s_kernel<<<nct * IN, 256, 0, st>>>(Smat, ldS, nrow, ncol, w->nkc, w->njt, w->Bs, w->sb);
        QX_LAUNCH_CHECK(c);
    }
    ProfScope prof(c, QEXXC_PROF_ROWQUAD, st);
    RqArgs a{};
    a.sc.A = w->A;
    a.sc.B = w->Bs;
    const long ntiles = c->Gpad / IM;
    a.sc.a_plane = ntiles * w->nkc * (long)ATILE;
    a.sc.b_plane = (long)w->njt * w->nkc * BTILE;
    a.sc.nkc = w->nkc;
    a.sc.nct = nct;
    a.sc.Nc = nrow;
    a.sc.tri = tri;
    a.sa = w->sa;
    a.sb = w->sb;
    a.ao = c->ao;
    a.sgn = sgn;
    // column groups per row tile: the A planes of a row tile (6 nkc x 16 KB) are re-read by every column tile; with one CTA
    // per row tile the planes of all resident CTAs (148 x 768 KB at nao = 1000) do not stay in L2
    int P = 1;
    while (P < 4 && (long)(c->num_sms / P) * ND * w->nkc * ATILE > (64L << 20) && a.sc.nct >= 4 * P) P *= 2;
    if (getenv("QEXXC_I8_P")) P = std::max(1, std::min(4, atoi(getenv("QEXXC_I8_P"))));
    a.sc.P = P;
    a.q = P > 1 ? w->qpart : q;
    a.ao_cstride = (long)c->GpadMax * c->Npad;
    a.q_cstride = P > 1 ? (long)c->GpadMax : q_cstride;
    a.Npad = sgn ? ldsgn : c->Npad;  // columns beyond this hold zero digits (and lie outside sgn / the ao rows)
    a.ncomp = ncomp;
    a.wide = getenv("QEXXC_I8_LD128") ? 0 : 1;
    a.f[0] = (tri ? 2.0 : 1.0) * fac4[0];
    a.f[1] = fac4[1];
    a.f[2] = fac4[2];
    a.f[3] = fac4[3];
    rowquad_i8_kernel<<<(unsigned)(ntiles * P), I8_THREADS, i8_smem(), st>>>(a);
    QX_LAUNCH_CHECK(c);
    if (P > 1) {
        rowquad_i8_sum_kernel<<<(unsigned)((c->Gpad + 255) / 256), 256, 0, st>>>(w->qpart, c->GpadMax, P, q, q_cstride, ncomp, c->Gpad);
        QX_LAUNCH_CHECK(c);
    }
    return QEXXC_OK;
}

int launch_rowquad_i8(qexxc_ctx* c, int ncomp, int tri, const double* fac4, double* q, long q_bstride, long q_cstride, cudaStream_t st) {
    for (int b = 0; b < c->B; ++b)
        QX_TRY(rowquad_i8_common(c, c->S + (long)b * c->Npad * c->Npad, c->Npad, c->Nc, c->Nc, ncomp, tri, fac4, nullptr, 0,
                                 q + (long)b * q_bstride, q_cstride, st));
    return QEXXC_OK;
}

// MO form of rho (pyscf eval_rho2, numint_legacy.py:527-545): q[g] = sum_k sgn[k] ((ao_0 L)[g,k])^2, L [Npad][ldL], nk columns
int launch_rowquad_mo_i8(qexxc_ctx* c, const double* L, int ldL, int nk, const double* sgn, double* q, long q_bstride, cudaStream_t st) {
    static const double one4[4] = {1.0, 0.0, 0.0, 0.0};
    for (int b = 0; b < c->B; ++b)
        QX_TRY(rowquad_i8_common(c, L + (long)b * c->Npad * ldL, ldL, c->Nc, nk > 0 ? nk : 1, 1, 0, one4, sgn + (long)b * ldL, ldL,
                                 q + (long)b * q_bstride, 0, st));
    return QEXXC_OK;
}

// the per-call operand of wsyrk (s .* ao_0, or a general B): block exponents, then digit planes
static int i8_slice_weighted(qexxc_ctx* c, I8Ws* w, const double* s, const double* Bsrc, int ngc, int nblk, cudaStream_t st) {
    ProfScope prof(c, QEXXC_PROF_SLICE, st);
    const dim3 grid(ngc, w->NpadK / 128);
    // default: the A operand is the row-scaled integer planes A8 (ao[g,i] = 2^(ea[g]-46) vA[g,i]), so the row's 2^ea[g] goes
    // into the B operand here
    const float* rexp = w->T ? nullptr : w->sa;
    if (Bsrc == nullptr) {
        blk_exp_kernel<<<nblk, 256, 0, st>>>(w->cmax, s, rexp, ngc, w->NpadK, w->eB);
        QX_LAUNCH_CHECK(c);
        slice_cols_kernel<<<grid, 256, ND * 16384, st>>>(c->ao, c->Npad, s, rexp, w->eB, w->NpadK, w->njt, w->plane_stride, w->W);
    } else {
        if (w->cmaxW_of != Bsrc || rexp == nullptr) {  // not already produced together with Bsrc (launch_build_aow_i8)
            colmax_kernel<<<ngc, 256, 0, st>>>(Bsrc, c->Npad, w->NpadK, rexp, w->cmaxW);
            QX_LAUNCH_CHECK(c);
        }
        w->cmaxW_of = nullptr;
        blk_exp_kernel<<<nblk, 256, 0, st>>>(w->cmaxW, nullptr, nullptr, ngc, w->NpadK, w->eB);
        QX_LAUNCH_CHECK(c);
        slice_cols_kernel<<<grid, 256, ND * 16384, st>>>(Bsrc, c->Npad, nullptr, rexp, w->eB, w->NpadK, w->njt, w->plane_stride, w->W);
    }
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

static int wsyrk_i8_one(qexxc_ctx* c, const double* s, const double* Bsrc, double scale, int tadd, double* out, cudaStream_t st);

int launch_wsyrk_i8(qexxc_ctx* c, const double* s, long s_bstride, const double* Bsrc, double scale, int tadd, double* out,
                    long out_bstride, cudaStream_t st) {
    for (int b = 0; b < c->B; ++b)
        QX_TRY(wsyrk_i8_one(c, s ? s + (long)b * s_bstride : nullptr, Bsrc ? Bsrc + (long)b * c->GpadMax * c->Npad : nullptr, scale, tadd,
                            out + (long)b * out_bstride, st));
    return QEXXC_OK;
}

static int wsyrk_i8_one(qexxc_ctx* c, const double* s, const double* Bsrc, double scale, int tadd, double* out, cudaStream_t st) {
    QX_TRY(i8_prepare(c, st));
    I8Ws* w = (I8Ws*)c->i8ws;
    const bool sym = (Bsrc == nullptr);
    const int ngc = c->Gpad / 128, nblk = (ngc + KDC - 1) / KDC;
    QX_TRY(i8_slice_weighted(c, w, s, Bsrc, ngc, nblk, st));
    ProfScope prof(c, QEXXC_PROF_WSYRK, st);
    const int nit = (c->Nc + IM - 1) / IM, njt = (c->Nc + IN - 1) / IN;  // tiles that hold data
    int ntile = 0;  // general: nit x njt; symmetric: the column tiles jt >= 2 it of each row tile
    for (int it = 0; it < nit; ++it) ntile += sym ? std::max(0, njt - 2 * it) : njt;
    // grid splits: as many CTAs as fill whole waves of the machine
    int nsplit = 1;
    double beff = 0.0;
    for (int ns = 1; ns <= nblk; ++ns) {
        const long ctas = (long)ntile * ns, waves = (ctas + c->num_sms - 1) / c->num_sms;
        const double eff = (double)ctas / (double)(waves * c->num_sms);
        if (eff > beff + 1e-9) beff = eff, nsplit = ns;
        if (ctas >= 2L * c->num_sms) break;
    }
    if ((size_t)ntile * nsplit * IM * IN > c->part_doubles) {
        set_error("internal: INT8 wsyrk partial tiles exceed the workspace");
        return QEXXC_ERR_STATE;
    }
    WsArgs a{};
    a.A = w->T ? w->T : w->A;
    a.B = w->W;
    a.plane_stride = w->plane_stride;
    a.a_plane = w->T ? w->plane_stride : (long)(c->Gpad / IM) * w->nkc * (long)ATILE;
    a.nkc = w->nkc;
    a.eA = w->T ? w->eA : nullptr;
    a.eB = w->eB;
    a.part = c->part;
    a.NpadK = w->NpadK;
    a.njt = njt;
    a.njtL = w->njt;
    a.ngc = ngc;
    a.nblk = nblk;
    a.ntile = ntile;
    a.nsplit = nsplit;
    a.sym = sym ? 1 : 0;
    if (w->T) wsyrk_i8_kernel<false><<<(unsigned)(ntile * nsplit), I8_THREADS, i8_smem(), st>>>(a);
    else wsyrk_i8_kernel<true><<<(unsigned)(ntile * nsplit), I8_THREADS, i8_smem(), st>>>(a);
    QX_LAUNCH_CHECK(c);
    wsyrk_i8_reduce_kernel<<<dim3((c->N + 31) / 32, (c->N + 7) / 8), 256, 0, st>>>(c->part, out, c->N, njt, ntile, nsplit, sym ? 1 : 0, scale, tadd);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

// Stage 3 for LocalMLP: the learned XC functional evaluated at every grid point, and its
// second-order reverse rule.
//
//   forward  : exc_g = MLP(theta, x_g),  vrho_g = d exc_g / d rho_g  (, vgamma_g = d exc_g / d sigma_g)
//              = exc_and_vrho_local  qedft/train/td/trainer_legacy_no_jit.py:56-63 with
//                build_local_mlp.apply_fn  qedft/models/classical/classical_models.py:168-172
//                (x/density_normalization_factor -> [Dense, act] x n_layers -> Dense(1)).
//              The input derivative is carried as forward-mode tangent rows next to the value
//              rows, so one pass gives (exc, vrho[, vgamma]) without a tape.
//   reverse  : cotangents (exc_bar, vrho_bar[, vgamma_bar]) -> (rho_bar, theta_bar): forward with
//              the tangent direction v = (vrho_bar, vgamma_bar), then reverse over both row
//              streams with seeds (exc_bar, 1)  (oracle/mlp_ref.py).
//
// Layout: a CTA (16 warps) owns R = 128 rows = 8-row blocks that alternate value/tangent streams
// of the same 8 grid points, so the activation epilogue finds z and z-dot of one point in the
// same thread.  Every Dense layer is a [128 x 64] x [64 x 64] product from shared memory:
// DMMA.8x8x4 for float64, register-blocked FFMA for float32, both with the same accumulator
// ownership (lane t: rows t>>2 + 8*mi, cols 2*(t&3)+{0,1} + 8*nj).  Weights stay resident in
// shared memory (pitch 68 => conflict-free fragment loads in all four transposition cases).
// CTAs are persistent over point blocks; the reverse-mode tape is a per-CTA slot (L2 resident).
#include "common.cuh"
#include "dmma.cuh"
#include "xc_act.cuh"

namespace qexxc {
namespace {

constexpr int HP = 64;        // padded hidden width
constexpr int LD = HP + 4;    // shared-memory pitch of activation / weight rows
constexpr int R = 128;        // rows per CTA iteration
constexpr int LDX = 4;        // pitch of the layer-0 feature rows (4 columns; 4 = 4 mod 16 is conflict-free)
constexpr int NW = 16;        // warps: 4 (m) x 4 (n), warp tile 32 x 16
constexpr int MLP_THREADS = NW * 32;
constexpr int MB = 4, NB = 2;
constexpr int MAXL = 3;       // hidden layers supported by the resident-weight kernel

struct MlpParams {
    int F, L, H, act, out_transform, xctype;
    double in_scale, out_scale;
    const double* rho;
    long rho_bstride, rho_cstride;
    const double* theta;
    double *exc, *vrho, *vgamma;
    long out_bstride;
    int Gpad, B, blocks_per_batch, nblocks, P;
    const double *exc_bar, *vrho_bar, *vgamma_bar;
    double* rho_bar;
    int accumulate;
    void* tape;
    double* theta_part;
    long n_theta;
};

// ---- warp-level [8*MBX x 8*NBX] (+)= A[.. x K] * B[K x ..] from shared memory -------------------
template <typename T, int MBX, int NBX, bool TA, bool TB>
struct WarpGemm;

template <int MBX, int NBX, bool TA, bool TB>
struct WarpGemm<double, MBX, NBX, TA, TB> {
    __device__ static __forceinline__ void run(double (&acc)[MBX][NBX][2], const double* A, int lda,
                                               const double* B, int ldb, int K, int g, int qd) {
#pragma unroll 4
        for (int k0 = 0; k0 < K; k0 += 4) {
            double a[MBX], bf[NBX];
#pragma unroll
            for (int mi = 0; mi < MBX; ++mi)
                a[mi] = TA ? A[(k0 + qd) * lda + mi * 8 + g] : A[(mi * 8 + g) * lda + k0 + qd];
#pragma unroll
            for (int nj = 0; nj < NBX; ++nj)
                bf[nj] = TB ? B[(nj * 8 + g) * ldb + k0 + qd] : B[(k0 + qd) * ldb + nj * 8 + g];
#pragma unroll
            for (int mi = 0; mi < MBX; ++mi)
#pragma unroll
                for (int nj = 0; nj < NBX; ++nj) dmma884(acc[mi][nj], a[mi], bf[nj]);
        }
    }
};

template <int MBX, int NBX, bool TA, bool TB>
struct WarpGemm<float, MBX, NBX, TA, TB> {
    __device__ static __forceinline__ void run(float (&acc)[MBX][NBX][2], const float* A, int lda,
                                               const float* B, int ldb, int K, int g, int qd) {
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float a[MBX], bf[NBX][2];
#pragma unroll
            for (int mi = 0; mi < MBX; ++mi) a[mi] = TA ? A[k * lda + mi * 8 + g] : A[(mi * 8 + g) * lda + k];
#pragma unroll
            for (int nj = 0; nj < NBX; ++nj) {
                bf[nj][0] = TB ? B[(nj * 8 + 2 * qd) * ldb + k] : B[k * ldb + nj * 8 + 2 * qd];
                bf[nj][1] = TB ? B[(nj * 8 + 2 * qd + 1) * ldb + k] : B[k * ldb + nj * 8 + 2 * qd + 1];
            }
#pragma unroll
            for (int mi = 0; mi < MBX; ++mi)
#pragma unroll
                for (int nj = 0; nj < NBX; ++nj) {
                    acc[mi][nj][0] = fmaf(a[mi], bf[nj][0], acc[mi][nj][0]);
                    acc[mi][nj][1] = fmaf(a[mi], bf[nj][1], acc[mi][nj][1]);
                }
        }
    }
};

template <typename T>
struct Smem {
    T* W1;     // [8][LD]      first Dense (rows >= F zero)
    T* Wh;     // [L-1][HP][LD] hidden Dense layers 2..L
    T* wl;     // [HP]         last Dense (out = 1)
    T* bias;   // [L+1][HP]
    T* X0;     // [R][LDX]
    T* buf0;   // [R][LD]
    T* buf1;   // [R][LD]
    T* u;      // [R]
    T* gacc;   // [L+4][HP]    reverse: bias, last-layer and first-layer weight gradient accumulators
    T* scratch;  // [8][HP]     reverse: partial column sums
};

template <typename T>
__host__ __device__ inline size_t smem_elems(int L, bool vjp) {
    size_t n = 8 * LD + (size_t)(L - 1) * HP * LD + HP + (size_t)(L + 1) * HP + (size_t)R * LDX +
               2 * (size_t)R * LD + R;
    if (vjp) n += (size_t)(L + 4) * HP + 8 * HP;
    return n;
}

template <typename T>
__device__ __forceinline__ Smem<T> carve(unsigned char* raw, int L, bool vjp) {
    Smem<T> s;
    T* p = reinterpret_cast<T*>(raw);
    s.W1 = p; p += 8 * LD;
    s.Wh = p; p += (size_t)(L - 1) * HP * LD;
    s.wl = p; p += HP;
    s.bias = p; p += (size_t)(L + 1) * HP;
    s.X0 = p; p += (size_t)R * LDX;
    s.buf0 = p; p += (size_t)R * LD;
    s.buf1 = p; p += (size_t)R * LD;
    s.u = p; p += R;
    s.gacc = vjp ? p : nullptr;
    s.scratch = vjp ? p + (size_t)(L + 4) * HP : nullptr;
    return s;
}

__device__ __forceinline__ long theta_off(int F, int H, int l) {
    // offset of Dense l (0-based) in flat theta: W [in][out] then b[out]
    long off = 0;
    for (int k = 0; k < l; ++k) {
        const int in = k == 0 ? F : H, out = H;
        off += (long)in * out + out;
    }
    return off;
}

template <typename T>
__device__ void load_weights(const MlpParams& p, const Smem<T>& s) {
    const int F = p.F, H = p.H, L = p.L;
    const double* th = p.theta;
    for (int i = threadIdx.x; i < 8 * LD; i += blockDim.x) {
        const int r = i / LD, c = i % LD;
        s.W1[i] = (r < F && c < H) ? (T)th[(long)r * H + c] : (T)0;
    }
    for (int l = 1; l < L; ++l) {
        const long off = theta_off(F, H, l);
        T* W = s.Wh + (size_t)(l - 1) * HP * LD;
        for (int i = threadIdx.x; i < HP * LD; i += blockDim.x) {
            const int r = i / LD, c = i % LD;
            W[i] = (r < H && c < H) ? (T)th[off + (long)r * H + c] : (T)0;
        }
    }
    const long offl = theta_off(F, H, L);
    for (int i = threadIdx.x; i < HP; i += blockDim.x) s.wl[i] = i < H ? (T)th[offl + i] : (T)0;
    for (int i = threadIdx.x; i < (L + 1) * HP; i += blockDim.x) {
        const int l = i / HP, c = i % HP;
        T v = (T)0;
        if (l < L) {
            if (c < H) v = (T)th[theta_off(F, H, l) + (long)(l == 0 ? F : H) * H + c];
        } else if (c == 0) {
            v = (T)th[offl + H];
        }
        s.bias[i] = v;
    }
}

// row of (point p in block, stream s) with NS streams interleaved in 8-row blocks
__device__ __forceinline__ int row_of(int p, int s, int NS) { return (p >> 3) * (8 * NS) + s * 8 + (p & 7); }

template <typename T, int NS>
__device__ __forceinline__ void init_acc_bias(T (&acc)[MB][NB][2], const T* bias, int col0, int qd) {
#pragma unroll
    for (int mi = 0; mi < MB; ++mi)
#pragma unroll
        for (int nj = 0; nj < NB; ++nj) {
            const bool val = (mi % NS) == 0;
            acc[mi][nj][0] = val ? bias[col0 + nj * 8 + 2 * qd] : (T)0;
            acc[mi][nj][1] = val ? bias[col0 + nj * 8 + 2 * qd + 1] : (T)0;
        }
}

// activation epilogue: value stream h = sigma(z); tangent streams hdot = sigma'(z) zdot
template <typename T, int NS>
__device__ __forceinline__ void act_store(const T (&acc)[MB][NB][2], T* dst, int row0, int col0, int g,
                                          int qd, int act) {
#pragma unroll
    for (int pg = 0; pg < MB / NS; ++pg)
#pragma unroll
        for (int nj = 0; nj < NB; ++nj) {
            T s0[2], s1[2], s2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) act_d012<T>(act, acc[pg * NS][nj][e], s0[e], s1[e], s2[e]);
            T* d = dst + (size_t)(row0 + (pg * NS) * 8 + g) * LD + col0 + nj * 8 + 2 * qd;
            d[0] = s0[0];
            d[1] = s0[1];
#pragma unroll
            for (int s = 1; s < NS; ++s) {
                T* dd = d + (size_t)s * 8 * LD;
                dd[0] = s1[0] * acc[pg * NS + s][nj][0];
                dd[1] = s1[1] * acc[pg * NS + s][nj][1];
            }
        }
}

// u[r] = sum_j buf[r][j] * wl[j]  (rotated start => conflict-free)
template <typename T>
__device__ __forceinline__ void last_layer_dots(const T* buf, const T* wl, T* u) {
    // 4 threads per row, 16 columns each (rotated start => conflict-free), fixed-order quad reduction
    const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
    T acc = (T)0;
#pragma unroll 8
    for (int jj = 0; jj < HP / 4; ++jj) {
        const int j = part * (HP / 4) + ((jj + r + 4 * part) & (HP / 4 - 1));
        acc = fma(buf[(size_t)r * LD + j], wl[j], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (part == 0) u[r] = acc;
}

// gout[j] += sum_{r < nrows} coef(r) * M[rowmap(r)][j] for j < HP, using all 512 threads: 8 row slices
// summed in parallel into `scratch` [8][HP], then added in fixed order.  Contains its own barriers.
template <typename T, typename RowMap, typename Coef>
__device__ __forceinline__ void col_sums(T* scratch, T* gout, const T* M, int nrows, RowMap rowmap, Coef coef) {
    const int j = threadIdx.x & (HP - 1), part = threadIdx.x >> 6;  // 512 threads = 8 parts x 64 columns
    const int per = nrows >> 3;
    T a = (T)0;
    for (int r = part * per; r < (part + 1) * per; ++r) a = fma(coef(r), M[(size_t)rowmap(r) * LD + j], a);
    __syncthreads();
    scratch[part * HP + j] = a;
    __syncthreads();
    if (threadIdx.x < HP) {
        T sacc = (T)0;
#pragma unroll
        for (int k = 0; k < 8; ++k) sacc += scratch[k * HP + j];
        gout[j] += sacc;
    }
}

__device__ __forceinline__ void block_to_bg(const MlpParams& p, int blk, int& b, long& g0) {
    b = blk / p.blocks_per_batch;
    g0 = (long)(blk - b * p.blocks_per_batch) * p.P;
}

// =================================================================================================
// forward: exc, vrho (, vgamma)
// =================================================================================================
// ACT >= 0: activation fixed at compile time (tanh, the reference default: keeps the 9-way switch and
// its exp/log code out of the instruction stream); ACT < 0: run-time p.act.
template <typename T, int NS, int ACT>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_fwd_kernel(const MlpParams p) {
    const int act = ACT >= 0 ? ACT : p.act;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem<T> s = carve<T>(smem_raw, p.L, false);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, qd = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int row0 = wm * (MB * 8), col0 = wn * (NB * 8);
    constexpr int P = R / NS;
    load_weights<T>(p, s);
    __syncthreads();

    for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
        int b;
        long g0;
        block_to_bg(p, blk, b, g0);
        // ---- features -> X0 ----
        if (threadIdx.x < P) {
            const int pt = threadIdx.x;
            const double* rb = p.rho + (long)b * p.rho_bstride + g0 + pt;
            double x[2];
            x[0] = rb[0];
            x[1] = 0.0;
            if (p.xctype == QEXXC_XC_GGA) {
                const double r1 = rb[p.rho_cstride], r2 = rb[2 * p.rho_cstride], r3 = rb[3 * p.rho_cstride];
                x[1] = r1 * r1 + r2 * r2 + r3 * r3;
            } else if (p.F == 2) {
                x[1] = rb[p.rho_cstride];
            }
#pragma unroll
            for (int st = 0; st < NS; ++st) {
                T* xr = s.X0 + (size_t)row_of(pt, st, NS) * LDX;
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    T v = (T)0;
                    if (st == 0) v = f < p.F ? (T)(p.in_scale * x[f & 1]) : (T)0;
                    else if (f == st - 1 && f < p.F) v = (T)p.in_scale;
                    xr[f] = v;
                }
            }
        }
        __syncthreads();
        // ---- layers ----
        T acc[MB][NB][2];
        T* cur = s.buf0;
        T* nxt = s.buf1;
        for (int l = 0; l < p.L; ++l) {
            init_acc_bias<T, NS>(acc, s.bias + l * HP, col0, qd);
            if (l == 0)
                WarpGemm<T, MB, NB, false, false>::run(acc, s.X0 + (size_t)row0 * LDX, LDX, s.W1 + col0, LD, 4, g, qd);
            else
                WarpGemm<T, MB, NB, false, false>::run(acc, cur + (size_t)row0 * LD, LD,
                                                       s.Wh + (size_t)(l - 1) * HP * LD + col0, LD, HP, g, qd);
            act_store<T, NS>(acc, nxt, row0, col0, g, qd, act);
            __syncthreads();
            T* t = cur;
            cur = nxt;
            nxt = t;
        }
        last_layer_dots<T>(cur, s.wl, s.u);
        __syncthreads();
        if (threadIdx.x < P) {
            const int pt = threadIdx.x;
            T u0 = s.u[row_of(pt, 0, NS)] + s.bias[p.L * HP];
            T y = u0, d1 = (T)1;
            if (p.out_transform == 1) {
                T s0, s1, s2;
                act_d012<T>(QEXXC_ACT_SWISH, u0, s0, s1, s2);
                y = (T)(-p.out_scale) * s0;
                d1 = (T)(-p.out_scale) * s1;
            }
            const long o = (long)b * p.out_bstride + g0 + pt;
            p.exc[o] = (double)y;
            p.vrho[o] = (double)(d1 * s.u[row_of(pt, 1, NS)]);
            if (NS > 2 && p.vgamma) p.vgamma[o] = (double)(d1 * s.u[row_of(pt, 2 % NS, NS)]);
        }
        __syncthreads();
    }
}

// =================================================================================================
// reverse: (exc_bar, vrho_bar[, vgamma_bar]) -> rho_bar, theta_bar partials
// =================================================================================================
template <typename T, int ACT>
__global__ void __launch_bounds__(MLP_THREADS, 1) mlp_vjp_kernel(const MlpParams p) {
    const int act = ACT >= 0 ? ACT : p.act;
    constexpr int NS = 2;
    constexpr int P = R / NS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Smem<T> s = carve<T>(smem_raw, p.L, true);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, qd = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    const int row0 = wm * (MB * 8), col0 = wn * (NB * 8);
    const int L = p.L;
    typedef typename Vec2<T>::type T2;
    T2* tape = reinterpret_cast<T2*>(p.tape) + (size_t)blockIdx.x * L * NW * MB * NB * 32;

    load_weights<T>(p, s);
    for (int i = threadIdx.x; i < (L + 4) * HP; i += blockDim.x) s.gacc[i] = (T)0;
    // weight-gradient accumulators: hidden layers 2..L as 16x16 warp tiles, layer 1 per thread
    T wacc[MAXL - 1][2][2][2];
#pragma unroll
    for (int l = 0; l < MAXL - 1; ++l)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) wacc[l][a][c][0] = wacc[l][a][c][1] = (T)0;
    __syncthreads();

    for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
        int b;
        long g0;
        block_to_bg(p, blk, b, g0);
        double rk[3] = {0.0, 0.0, 0.0};
        if (threadIdx.x < P) {
            const int pt = threadIdx.x;
            const double* rb = p.rho + (long)b * p.rho_bstride + g0 + pt;
            double x[2], v[2];
            x[0] = rb[0];
            x[1] = 0.0;
            const long o = (long)b * p.out_bstride + g0 + pt;
            v[0] = p.vrho_bar[o];
            v[1] = (p.F == 2 && p.vgamma_bar) ? p.vgamma_bar[o] : 0.0;
            if (p.xctype == QEXXC_XC_GGA) {
                rk[0] = rb[p.rho_cstride];
                rk[1] = rb[2 * p.rho_cstride];
                rk[2] = rb[3 * p.rho_cstride];
                x[1] = rk[0] * rk[0] + rk[1] * rk[1] + rk[2] * rk[2];
            } else if (p.F == 2) {
                x[1] = rb[p.rho_cstride];
            }
            T* x0 = s.X0 + (size_t)row_of(pt, 0, NS) * LDX;
            T* x1 = s.X0 + (size_t)row_of(pt, 1, NS) * LDX;
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                x0[f] = f < p.F ? (T)(p.in_scale * x[f & 1]) : (T)0;
                x1[f] = f < p.F ? (T)(p.in_scale * v[f & 1]) : (T)0;
            }
        }
        __syncthreads();
        // ---------------- forward with tape ----------------
        T acc[MB][NB][2];
        T* cur = s.buf0;
        T* nxt = s.buf1;
        for (int l = 0; l < L; ++l) {
            init_acc_bias<T, NS>(acc, s.bias + l * HP, col0, qd);
            if (l == 0)
                WarpGemm<T, MB, NB, false, false>::run(acc, s.X0 + (size_t)row0 * LDX, LDX, s.W1 + col0, LD, 4, g, qd);
            else
                WarpGemm<T, MB, NB, false, false>::run(acc, cur + (size_t)row0 * LD, LD,
                                                       s.Wh + (size_t)(l - 1) * HP * LD + col0, LD, HP, g, qd);
            // tape: tangent rows keep zdot; value rows keep z, or -- tanh build -- h = tanh(z) itself, from
            // which the reverse sweep gets sigma' = 1 - h^2 and sigma'' = -2 h sigma' without another tanh
            T2* tp = tape + ((size_t)l * NW + warp) * MB * NB * 32 + lane;
#pragma unroll
            for (int pg = 0; pg < MB / NS; ++pg)
#pragma unroll
                for (int nj = 0; nj < NB; ++nj) {
                    T s0[2], s1[2], s2[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) act_d012<T>(act, acc[pg * NS][nj][e], s0[e], s1[e], s2[e]);
                    T* d = nxt + (size_t)(row0 + (pg * NS) * 8 + g) * LD + col0 + nj * 8 + 2 * qd;
                    d[0] = s0[0];
                    d[1] = s0[1];
                    d[8 * LD] = s1[0] * acc[pg * NS + 1][nj][0];
                    d[8 * LD + 1] = s1[1] * acc[pg * NS + 1][nj][1];
                    tp[((pg * NS) * NB + nj) * 32] = (ACT == QEXXC_ACT_TANH) ? Vec2<T>::make(s0[0], s0[1])
                                                                            : Vec2<T>::make(acc[pg * NS][nj][0], acc[pg * NS][nj][1]);
                    tp[((pg * NS + 1) * NB + nj) * 32] = Vec2<T>::make(acc[pg * NS + 1][nj][0], acc[pg * NS + 1][nj][1]);
                }
            __syncthreads();
            T* t = cur;
            cur = nxt;
            nxt = t;
        }
        // ---------------- output layer and its adjoint ----------------
        last_layer_dots<T>(cur, s.wl, s.u);
        __syncthreads();
        if (threadIdx.x < P) {
            const int pt = threadIdx.x;
            const int r0 = row_of(pt, 0, NS), r1 = row_of(pt, 1, NS);
            const T u0 = s.u[r0] + s.bias[L * HP], ud = s.u[r1];
            const long o = (long)b * p.out_bstride + g0 + pt;
            const T yb = (T)(p.xctype == QEXXC_XC_NN_GLOBAL ? p.exc_bar[b] : p.exc_bar[o]);
            T ub = yb, udb = (T)1;  // seeds (exc_bar, 1) on (y, ydot)
            if (p.out_transform == 1) {
                T s0, s1, s2;
                act_d012<T>(QEXXC_ACT_SWISH, u0, s0, s1, s2);
                const T c = (T)(-p.out_scale);
                ub = yb * c * s1 + c * s2 * ud;
                udb = c * s1;
            }
            s.u[r0] = ub;
            s.u[r1] = udb;
        }
        __syncthreads();
        // d wl[j] = sum_r u[r] h_L[r][j] (all rows, both streams); d b_last = sum of the value-row seeds
        col_sums<T>(s.scratch, s.gacc + (L + 1) * HP, cur, R, [](int r) { return r; },
                    [&](int r) { return s.u[r]; });
        if (warp == 0) {
            T bsum = s.u[row_of(lane, 0, NS)] + s.u[row_of(lane + 32, 0, NS)];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
            if (lane == 0) s.gacc[L * HP] += bsum;
        }
        // adjoint of the last hidden activations, directly in accumulator layout
#pragma unroll
        for (int mi = 0; mi < MB; ++mi)
#pragma unroll
            for (int nj = 0; nj < NB; ++nj) {
                const T ur = s.u[row0 + mi * 8 + g];
                acc[mi][nj][0] = ur * s.wl[col0 + nj * 8 + 2 * qd];
                acc[mi][nj][1] = ur * s.wl[col0 + nj * 8 + 2 * qd + 1];
            }
        __syncthreads();  // everyone is done reading `cur` (h_L) and s.u
        T* Bz = s.buf0;
        T* Bh = s.buf1;
        // ---------------- reverse over hidden layers ----------------
#pragma unroll
        for (int li = MAXL - 1; li >= 0; --li) {
            if (li < L) {
                const int l = li;  // Dense l feeds activation l
                const T2* tp = tape + ((size_t)l * NW + warp) * MB * NB * 32 + lane;
#pragma unroll
                for (int pg = 0; pg < MB / NS; ++pg)
#pragma unroll
                    for (int nj = 0; nj < NB; ++nj) {
                        const T2 z = tp[((pg * NS) * NB + nj) * 32];
                        const T2 zd = tp[((pg * NS + 1) * NB + nj) * 32];
                        const T zz[2] = {z.x, z.y}, zzd[2] = {zd.x, zd.y};
                        T* d = Bz + (size_t)(row0 + (pg * NS) * 8 + g) * LD + col0 + nj * 8 + 2 * qd;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            T s0, s1, s2;
                            if (ACT == QEXXC_ACT_TANH) {  // the tape holds h = tanh(z)
                                s1 = (T)1 - zz[e] * zz[e];
                                s2 = (T)-2 * zz[e] * s1;
                            } else {
                                act_d012<T>(act, zz[e], s0, s1, s2);
                            }
                            const T hb = acc[pg * NS][nj][e], hdb = acc[pg * NS + 1][nj][e];
                            d[e] = hb * s1 + hdb * s2 * zzd[e];  // z_bar
                            d[8 * LD + e] = hdb * s1;            // zdot_bar
                        }
                    }
                if (l > 0) {  // inputs of Dense l: h_{l-1} = sigma(z_{l-1}) from the tape
                    const T2* tq = tape + ((size_t)(l - 1) * NW + warp) * MB * NB * 32 + lane;
#pragma unroll
                    for (int pg = 0; pg < MB / NS; ++pg)
#pragma unroll
                        for (int nj = 0; nj < NB; ++nj) {
                            const T2 a = tq[((pg * NS) * NB + nj) * 32];
                            const T2 ad = tq[((pg * NS + 1) * NB + nj) * 32];
                            const T av[2] = {a.x, a.y}, adv[2] = {ad.x, ad.y};
                            T* d = Bh + (size_t)(row0 + (pg * NS) * 8 + g) * LD + col0 + nj * 8 + 2 * qd;
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                T s0, s1, s2;
                                if (ACT == QEXXC_ACT_TANH) {
                                    s0 = av[e];
                                    s1 = (T)1 - s0 * s0;
                                } else {
                                    act_d012<T>(act, av[e], s0, s1, s2);
                                }
                                d[e] = s0;
                                d[8 * LD + e] = s1 * adv[e];
                            }
                        }
                }
                __syncthreads();
                // bias gradient: column sums of z_bar over the value rows
                col_sums<T>(s.scratch, s.gacc + l * HP, Bz, P, [](int r) { return row_of(r, 0, NS); },
                            [](int) { return (T)1; });
                if (l > 0) {
                    // dW_l (64x64) += [h; hdot]^T [z_bar; zdot_bar], warp tile 16 x 16
                    WarpGemm<T, 2, 2, true, false>::run(wacc[l - 1 < MAXL - 1 ? l - 1 : 0], Bh + wm * 16, LD,
                                                        Bz + wn * 16, LD, R, g, qd);
                    // adjoint of the previous activations: [z_bar; zdot_bar] W_l^T
#pragma unroll
                    for (int mi = 0; mi < MB; ++mi)
#pragma unroll
                        for (int nj = 0; nj < NB; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = (T)0;
                    WarpGemm<T, MB, NB, false, true>::run(acc, Bz + (size_t)row0 * LD, LD,
                                                          s.Wh + (size_t)(l - 1) * HP * LD + (size_t)col0 * LD, LD, HP, g, qd);
                } else {
                    // dW_1 [F][H] and the input cotangent
                    for (int f = 0; f < p.F; ++f)
                        col_sums<T>(s.scratch, s.gacc + (L + 2 + f) * HP, Bz, R, [](int r) { return r; },
                                    [&](int r) { return s.X0[(size_t)r * LDX + f]; });
                    if (threadIdx.x < 4 * P) {  // 4 threads per point, 16 columns each, quad reduction
                        const int pt = threadIdx.x >> 2, part = threadIdx.x & 3;
                        const int r = row_of(pt, 0, NS);
                        T xb[2] = {(T)0, (T)0};
                        for (int jj = 0; jj < HP / 4; ++jj) {
                            const int j = part * (HP / 4) + ((jj + pt + 4 * part) & (HP / 4 - 1));
                            const T zb = Bz[(size_t)r * LD + j];
                            xb[0] = fma(zb, s.W1[j], xb[0]);
                            xb[1] = fma(zb, s.W1[LD + j], xb[1]);
                        }
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            xb[e] += __shfl_xor_sync(0xffffffffu, xb[e], 1);
                            xb[e] += __shfl_xor_sync(0xffffffffu, xb[e], 2);
                        }
                        if (part == 0) {
                            s.u[pt] = xb[0] * (T)p.in_scale;
                            s.u[P + pt] = xb[1] * (T)p.in_scale;
                        }
                    }
                }
                __syncthreads();
            }
        }
        // ---------------- write rho_bar ----------------
        if (threadIdx.x < P) {
            const int pt = threadIdx.x;
            double* ob = p.rho_bar + (long)b * p.rho_bstride + g0 + pt;
            const double x0b = (double)s.u[pt], x1b = (double)s.u[P + pt];
            ob[0] = (p.accumulate ? ob[0] : 0.0) + x0b;
            if (p.xctype == QEXXC_XC_GGA) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double* q = ob + (long)(k + 1) * p.rho_cstride;
                    *q = (p.accumulate ? *q : 0.0) + 2.0 * x1b * rk[k];
                }
            } else if (p.F == 2) {
                double* q = ob + p.rho_cstride;
                *q = (p.accumulate ? *q : 0.0) + x1b;
            }
        }
        __syncthreads();
    }
    // ---------------- per-CTA theta_bar partial ----------------
    double* out = p.theta_part + (size_t)blockIdx.x * p.n_theta;
    const int H = p.H, F = p.F;
    __syncthreads();
    if (threadIdx.x < 2 * HP) {
        const int f = threadIdx.x / HP, j = threadIdx.x % HP;
        if (f < F && j < H) out[(long)f * H + j] = (double)s.gacc[(L + 2 + f) * HP + j];
    }
#pragma unroll
    for (int l = 1; l < MAXL; ++l) {
        if (l < L) {
            const long off = theta_off(F, H, l);
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int i = wm * 16 + a * 8 + g, j = wn * 16 + c * 8 + 2 * qd + e;
                        if (i < H && j < H) out[off + (long)i * H + j] = (double)wacc[l - 1][a][c][e];
                    }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (L + 2) * HP; i += blockDim.x) {
        const int l = i / HP, j = i % HP;
        if (l < L) {
            if (j < H) out[theta_off(F, H, l) + (long)(l == 0 ? F : H) * H + j] = (double)s.gacc[i];
        } else if (l == L) {
            if (j == 0) out[theta_off(F, H, L) + H] = (double)s.gacc[i];
        } else if (j < H) {
            out[theta_off(F, H, L) + j] = (double)s.gacc[i];
        }
    }
}

// theta_bar[k] = sum over CTA partials in fixed order
__global__ void theta_reduce_kernel(const double* __restrict__ part, int nparts, long n, double* __restrict__ out,
                                    int accumulate) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double a = accumulate ? out[k] : 0.0;
    for (int c = 0; c < nparts; ++c) a += part[(size_t)c * n + k];
    out[k] = a;
}

template <typename K>
int set_smem_attr(K kernel, size_t bytes) {
    QX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return QEXXC_OK;
}

int check_supported(const qexxc_net_desc& net) {
    if (net.width > HP || net.width < 1) {
        set_error("LocalMLP: n_neurons=%d not supported by the resident-weight kernel (max %d)", net.width, HP);
        return QEXXC_ERR_UNSUPPORTED;
    }
    if (net.n_hidden < 1 || net.n_hidden > MAXL) {
        set_error("LocalMLP: n_layers=%d not supported (1..%d)", net.n_hidden, MAXL);
        return QEXXC_ERR_UNSUPPORTED;
    }
    if (net.n_features < 1 || net.n_features > 2) {
        set_error("LocalMLP: n_features=%d not supported (1 or 2)", net.n_features);
        return QEXXC_ERR_UNSUPPORTED;
    }
    return QEXXC_OK;
}

MlpParams base_params(const qexxc_ctx* c, int xctype) {
    MlpParams p{};
    p.F = c->net.n_features;
    p.L = c->net.n_hidden;
    p.H = c->net.width;
    p.act = c->net.activation;
    p.out_transform = c->net.out_transform;
    p.xctype = xctype;
    p.in_scale = c->net.in_scale;
    p.out_scale = c->net.out_scale;
    p.n_theta = c->n_theta;
    return p;
}

}  // namespace

int mlp_local_grid(const qexxc_ctx* c) { return c->num_sms; }
size_t mlp_local_tape_bytes(const qexxc_ctx* c) {
    const size_t el = c->net.precision == QEXXC_PREC_F32 ? 4 : 8;
    const size_t resident = (size_t)mlp_local_grid(c) * c->net.n_hidden * NW * MB * NB * 32 * 2 * el;
    const size_t tc = mlp_tc_enabled(c) ? mlp_tc_tape_bytes(c) : 0;
    return tc > resident ? tc : resident;
}

// rho/rho_bar etc. are given as [B][C][ld]-style arrays through explicit strides; npts_per_batch
// must be a multiple of 64.
int launch_mlp_local_fwd(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                         const double* theta, double* exc, double* vrho, double* vgamma, long out_bstride,
                         int nbatch, long npts_per_batch, cudaStream_t st) {
    if (mlp_is_wide(c->net))
        return launch_mlp_wide_fwd(c, xctype, rho, rho_bstride, rho_cstride, theta, exc, vrho, vgamma, out_bstride, nbatch,
                                   npts_per_batch, st);
    QX_TRY(check_supported(c->net));
    if (mlp_tc_enabled(c))
        return launch_mlp_tc_fwd(c, xctype, rho, rho_bstride, rho_cstride, theta, exc, vrho, vgamma, out_bstride, nbatch,
                                 npts_per_batch, st);
    MlpParams p = base_params(c, xctype);
    p.rho = rho;
    p.rho_bstride = rho_bstride;
    p.rho_cstride = rho_cstride;
    p.theta = theta;
    p.exc = exc;
    p.vrho = vrho;
    p.vgamma = vgamma;
    p.out_bstride = out_bstride;
    p.B = nbatch;
    const bool f32 = c->net.precision == QEXXC_PREC_F32;
    const int NS = p.F == 1 ? 2 : 4;
    p.P = R / NS;
    p.blocks_per_batch = (int)((npts_per_batch + p.P - 1) / p.P);
    p.nblocks = p.blocks_per_batch * nbatch;
    const int grid = p.nblocks < c->num_sms ? p.nblocks : c->num_sms;
    if (grid <= 0) return QEXXC_OK;
    ProfScope prof(c, QEXXC_PROF_XC_FWD, st);
#define QX_FWD(T, NSV)                                                                      \
    do {                                                                                    \
        const size_t sm = smem_elems<T>(p.L, false) * sizeof(T);                            \
        if (p.act == QEXXC_ACT_TANH) {                                                      \
            QX_TRY(set_smem_attr(mlp_fwd_kernel<T, NSV, QEXXC_ACT_TANH>, sm));              \
            mlp_fwd_kernel<T, NSV, QEXXC_ACT_TANH><<<grid, MLP_THREADS, sm, st>>>(p);       \
        } else {                                                                            \
            QX_TRY(set_smem_attr(mlp_fwd_kernel<T, NSV, -1>, sm));                          \
            mlp_fwd_kernel<T, NSV, -1><<<grid, MLP_THREADS, sm, st>>>(p);                   \
        }                                                                                   \
    } while (0)
    if (f32) {
        if (NS == 2) QX_FWD(float, 2);
        else QX_FWD(float, 4);
    } else {
        if (NS == 2) QX_FWD(double, 2);
        else QX_FWD(double, 4);
    }
#undef QX_FWD
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_mlp_local_vjp(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                         const double* theta, const double* exc_bar, const double* vrho_bar,
                         const double* vgamma_bar, long in_bstride, double* rho_bar, int accumulate,
                         double* theta_bar, int accumulate_theta, int nbatch, long npts_per_batch,
                         cudaStream_t st) {
    if (mlp_is_wide(c->net))
        return launch_mlp_wide_vjp(c, xctype, rho, rho_bstride, rho_cstride, theta, exc_bar, vrho_bar, vgamma_bar, in_bstride,
                                   rho_bar, accumulate, theta_bar, accumulate_theta, nbatch, npts_per_batch, st);
    QX_TRY(check_supported(c->net));
    MlpParams p = base_params(c, xctype);
    p.rho = rho;
    p.rho_bstride = rho_bstride;
    p.rho_cstride = rho_cstride;
    p.theta = theta;
    p.exc_bar = exc_bar;
    p.vrho_bar = vrho_bar;
    p.vgamma_bar = vgamma_bar;
    p.out_bstride = in_bstride;
    p.rho_bar = rho_bar;
    p.accumulate = accumulate;
    p.B = nbatch;
    p.P = R / 2;
    p.blocks_per_batch = (int)((npts_per_batch + p.P - 1) / p.P);
    p.nblocks = p.blocks_per_batch * nbatch;
    p.tape = c->tape;
    p.theta_part = c->red;
    const int grid = p.nblocks < c->num_sms ? p.nblocks : c->num_sms;
    if (grid <= 0) return QEXXC_OK;
    const bool f32 = c->net.precision == QEXXC_PREC_F32;
    ProfScope prof(c, QEXXC_PROF_XC_VJP, st);
    if (mlp_tc_enabled(c)) {
        int tgrid = 0;
        QX_TRY(launch_mlp_tc_vjp(c, xctype, rho, rho_bstride, rho_cstride, theta, exc_bar, vrho_bar, vgamma_bar, in_bstride,
                                 rho_bar, accumulate, nbatch, npts_per_batch, &tgrid, st));
        if (tgrid > 0) {
            theta_reduce_kernel<<<(unsigned)((c->n_theta + 255) / 256), 256, 0, st>>>(c->red, tgrid, c->n_theta, theta_bar,
                                                                                   accumulate_theta);
            QX_LAUNCH_CHECK(c);
        }
        return QEXXC_OK;
    }
#define QX_VJP(T, A)                                                                        \
    do {                                                                                    \
        const size_t sm = smem_elems<T>(p.L, true) * sizeof(T);                             \
        QX_TRY(set_smem_attr(mlp_vjp_kernel<T, A>, sm));                                    \
        mlp_vjp_kernel<T, A><<<grid, MLP_THREADS, sm, st>>>(p);                             \
    } while (0)
    const bool tanh_act = p.act == QEXXC_ACT_TANH;
    if (f32) {
        if (tanh_act) QX_VJP(float, QEXXC_ACT_TANH);
        else QX_VJP(float, -1);
    } else {
        if (tanh_act) QX_VJP(double, QEXXC_ACT_TANH);
        else QX_VJP(double, -1);
    }
#undef QX_VJP
    QX_LAUNCH_CHECK(c);
    theta_reduce_kernel<<<(unsigned)((c->n_theta + 255) / 256), 256, 0, st>>>(c->red, grid, c->n_theta, theta_bar,
                                                                           accumulate_theta);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

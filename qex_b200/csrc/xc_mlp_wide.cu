// Stage 3 for LocalMLP of ANY width and depth (the reference accepts any n_neurons / n_layers:
// qedft/models/networks.py:103-108, classical_models.py:75-115; the 3D trainer's default network is a 512-wide
// flax MLP, trainer_legacy_no_jit.py:96-107,136-140).  xc_mlp.cu keeps the weights of a <= 64 x 3 network resident
// in shared memory; beyond that the activations of a point block no longer fit on chip, so this path runs the
// network layer by layer over chunks of grid points with the activations in HBM/L2:
//
//   * one FP64 tensor-core GEMM kernel (DMMA.8x8x4, 64 x 64 CTA tiles, register-prefetched K slabs) in its NN / NT /
//     TN forms does every product: Dense layers, back-propagation through W^T, weight gradients H^T Z_bar, and --
//     with a 64-column coefficient matrix -- the bias / first-layer / last-layer gradient reductions and the input
//     cotangent; accumulation over chunks is `beta = 1` on the same output, i.e. sequential and deterministic;
//   * value and tangent streams are stacked as row blocks (stream s = rows [s P, (s+1) P)), so one GEMM serves all
//     streams of a layer; the activation kernels pair the rows of one point.
//
// Same functional and reverse rule as xc_mlp.cu (exc_and_vrho_local, trainer_legacy_no_jit.py:56-63); always float64
// (a `precision = "f32"` request for a wide network is served at the higher precision).
#include "common.cuh"
#include "dmma.cuh"
#include "xc_act.cuh"

namespace qexxc {
namespace {

constexpr int WT = 64;  // tile / padding unit
constexpr int KS = 32;  // K slab
constexpr int PA = KS + 4, PB = WT + 4;

// C[M x N] = beta * C + op(A) op(B);  TA: A is stored [K x M];  TB: B is stored [N x K].  M, N % 64 == 0, K % 32 == 0.
template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm64_kernel(const double* __restrict__ A, long lda, const double* __restrict__ B,
                                                     long ldb, double* __restrict__ C, long ldc, int K, int beta,
                                                     int kchunk, long cz) {
    __shared__ __align__(16) double As[WT * PA > KS * PB ? WT * PA : KS * PB];
    __shared__ __align__(16) double Bs[WT * PA > KS * PB ? WT * PA : KS * PB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, qd = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const long m0 = (long)blockIdx.y * WT, n0 = (long)blockIdx.x * WT;
    double acc[2][4][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) acc[mi][nj][0] = acc[mi][nj][1] = 0.0;
    // slab loaders: 2048 doubles per operand per slab = 4 double2 per thread
    double2 ra[4], rb[4];
    auto load_a = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;  // double2 index
            if (!TA) {  // [64 m][32 k]: 16 double2 per row
                const int r = e >> 4, c = (e & 15) * 2;
                ra[i] = *reinterpret_cast<const double2*>(A + (m0 + r) * lda + k0 + c);
            } else {  // [32 k][64 m]: 32 double2 per row
                const int r = e >> 5, c = (e & 31) * 2;
                ra[i] = *reinterpret_cast<const double2*>(A + (long)(k0 + r) * lda + m0 + c);
            }
        }
    };
    auto load_b = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            if (!TB) {  // [32 k][64 n]
                const int r = e >> 5, c = (e & 31) * 2;
                rb[i] = *reinterpret_cast<const double2*>(B + (long)(k0 + r) * ldb + n0 + c);
            } else {  // [64 n][32 k]
                const int r = e >> 4, c = (e & 15) * 2;
                rb[i] = *reinterpret_cast<const double2*>(B + (n0 + r) * ldb + k0 + c);
            }
        }
    };
    auto store_ab = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            if (!TA) {
                const int r = e >> 4, c = (e & 15) * 2;
                *reinterpret_cast<double2*>(As + r * PA + c) = ra[i];
            } else {
                const int r = e >> 5, c = (e & 31) * 2;
                *reinterpret_cast<double2*>(As + r * PB + c) = ra[i];
            }
            if (!TB) {
                const int r = e >> 5, c = (e & 31) * 2;
                *reinterpret_cast<double2*>(Bs + r * PB + c) = rb[i];
            } else {
                const int r = e >> 4, c = (e & 15) * 2;
                *reinterpret_cast<double2*>(Bs + r * PA + c) = rb[i];
            }
        }
    };
    // split-K: block z multiplies K range [z kchunk, (z+1) kchunk) into its own partial output C + z cz
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    C += (long)blockIdx.z * cz;
    load_a(kbeg);
    load_b(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += KS) {
        __syncthreads();  // the previous slab has been consumed
        store_ab();
        __syncthreads();
        if (k0 + KS < kend) {  // prefetch the next slab into registers while this one is multiplied
            load_a(k0 + KS);
            load_b(k0 + KS);
        }
#pragma unroll
        for (int kk = 0; kk < KS; kk += 4) {
            double a[2], b[4];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
                a[mi] = TA ? As[(kk + qd) * PB + 16 * wm + 8 * mi + g] : As[(16 * wm + 8 * mi + g) * PA + kk + qd];
#pragma unroll
            for (int nj = 0; nj < 4; ++nj)
                b[nj] = TB ? Bs[(32 * wn + 8 * nj + g) * PA + kk + qd] : Bs[(kk + qd) * PB + 32 * wn + 8 * nj + g];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int nj = 0; nj < 4; ++nj) dmma884(acc[mi][nj], a[mi], b[nj]);
        }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
            double* c = C + (m0 + 16 * wm + 8 * mi + g) * ldc + n0 + 32 * wn + 8 * nj + 2 * qd;
            double2 v = make_double2(acc[mi][nj][0], acc[mi][nj][1]);
            if (beta) {
                const double2 o = *reinterpret_cast<const double2*>(c);
                v.x += o.x;
                v.y += o.y;
            }
            *reinterpret_cast<double2*>(c) = v;
        }
}

struct WideWs {
    int Hp, L, S, Pc;
    double *Wp, *Wg, *W1p, *wlp, *bias, *Z, *Hh, *D0, *D1, *U, *A64, *XB, *Gb, *Gwl, *G2, *feat, *part;
    long part_doubles;
};

struct WideNet {
    int F, L, H, Hp, act, out_transform, xctype;
    double in_scale, out_scale;
};

__device__ __forceinline__ long wth_off(int F, int H, int l) {
    long off = 0;
    for (int k = 0; k < l; ++k) off += (long)(k == 0 ? F : H) * H + H;
    return off;
}

// theta -> zero-padded operand matrices
__global__ void wide_pack_kernel(WideNet n, const double* __restrict__ th, WideWs w) {
    const long Hp = n.Hp, H = n.H;
    const long total = (long)(n.L - 1) * Hp * Hp + 64 * Hp + Hp * 64 + (long)(n.L + 1) * Hp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        long k = i;
        if (k < (long)(n.L - 1) * Hp * Hp) {
            const int l = (int)(k / (Hp * Hp)) + 1;
            const long r = (k % (Hp * Hp)) / Hp, c = k % Hp;
            w.Wp[k] = (r < H && c < H) ? th[wth_off(n.F, n.H, l) + r * H + c] : 0.0;
            continue;
        }
        k -= (long)(n.L - 1) * Hp * Hp;
        if (k < 64 * Hp) {  // W1p [64][Hp]: row f = first Dense weights of feature f
            const long f = k / Hp, c = k % Hp;
            w.W1p[k] = (f < n.F && c < H) ? th[f * H + c] : 0.0;
            continue;
        }
        k -= 64 * Hp;
        if (k < Hp * 64) {  // wlp [Hp][64]: column 0 = last Dense weights
            const long r = k / 64, c = k % 64;
            w.wlp[k] = (c == 0 && r < H) ? th[wth_off(n.F, n.H, n.L) + r] : 0.0;
            continue;
        }
        k -= Hp * 64;
        const int l = (int)(k / Hp);
        const long c = k % Hp;
        if (l < n.L) w.bias[k] = c < H ? th[wth_off(n.F, n.H, l) + (long)(l == 0 ? n.F : n.H) * H + c] : 0.0;
        else w.bias[k] = c == 0 ? th[wth_off(n.F, n.H, n.L) + H] : 0.0;  // row L: the output bias
    }
}

struct WideIo {
    const double* rho;
    long rho_bstride, rho_cstride;
    double *exc, *vrho, *vgamma;
    const double *exc_bar, *vrho_bar, *vgamma_bar;
    double* rho_bar;
    long io_bstride;  // batch stride of exc / vrho / the cotangents
    long npts;        // points per batch element
    long total;       // nbatch * npts
    long t0;          // first flattened point of this chunk
    int accumulate;
};

// feat[k][p]: k = 0,1 scaled features, 2,3 scaled tangent direction (reverse), 4 exc_bar seed
__global__ void wide_prep_kernel(WideNet n, WideIo io, WideWs w, int vjp) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= w.Pc) return;
    const long t = io.t0 + p;
    double x0 = 0, x1 = 0, d0 = 0, d1 = 0, yb = 0;
    if (t < io.total) {
        const long b = t / io.npts, g = t - b * io.npts;
        const double* rb = io.rho + b * io.rho_bstride + g;
        x0 = rb[0];
        if (n.xctype == QEXXC_XC_GGA) {
            const double r1 = rb[io.rho_cstride], r2 = rb[2 * io.rho_cstride], r3 = rb[3 * io.rho_cstride];
            x1 = r1 * r1 + r2 * r2 + r3 * r3;
        } else if (n.F == 2) {
            x1 = rb[io.rho_cstride];
        }
        if (vjp) {
            const long o = b * io.io_bstride + g;
            d0 = io.vrho_bar[o];
            if (n.F == 2 && io.vgamma_bar) d1 = io.vgamma_bar[o];
            yb = n.xctype == QEXXC_XC_NN_GLOBAL ? io.exc_bar[b] : io.exc_bar[o];
        }
    }
    w.feat[p] = n.in_scale * x0;
    w.feat[w.Pc + p] = n.in_scale * x1;
    w.feat[2 * (long)w.Pc + p] = n.in_scale * d0;
    w.feat[3 * (long)w.Pc + p] = n.in_scale * d1;
    w.feat[4 * (long)w.Pc + p] = yb;
}

// first Dense + activation.  Forward: tangent stream s = d/d feature (s-1); reverse: one tangent stream along feat[2,3].
__global__ void wide_first_kernel(WideNet n, WideWs w, int vjp) {
    const long Hp = n.Hp, Pc = w.Pc;
    const long total = Pc * Hp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long p = i / Hp, j = i % Hp;
        const double w0 = w.W1p[j], w1 = w.W1p[Hp + j];
        const double z = w.bias[j] + w.feat[p] * w0 + w.feat[Pc + p] * w1;
        double s0, s1, s2;
        act_d012<double>(n.act, z, s0, s1, s2);
        w.Z[i] = z;
        w.Hh[i] = s0;
        if (vjp) {
            const double zd = w.feat[2 * Pc + p] * w0 + w.feat[3 * Pc + p] * w1;
            w.Z[Pc * Hp + i] = zd;
            w.Hh[Pc * Hp + i] = s1 * zd;
        } else {
            for (int s = 1; s < w.S; ++s) {
                const double zd = n.in_scale * (s == 1 ? w0 : w1);
                w.Z[(long)s * Pc * Hp + i] = zd;
                w.Hh[(long)s * Pc * Hp + i] = s1 * zd;
            }
        }
    }
}

// hidden activation of layer l: Z (value rows, bias added here) -> Hh; tangent rows: Hh = sigma'(z) zdot
__global__ void wide_act_kernel(WideNet n, WideWs w, int l) {
    const long Hp = n.Hp, Pc = w.Pc, lay = (long)l * w.S * Pc * Hp;
    const long total = Pc * Hp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long j = i % Hp;
        const double z = w.Z[lay + i] + w.bias[(long)l * Hp + j];
        double s0, s1, s2;
        act_d012<double>(n.act, z, s0, s1, s2);
        w.Z[lay + i] = z;
        w.Hh[lay + i] = s0;
        for (int s = 1; s < w.S; ++s) w.Hh[lay + (long)s * Pc * Hp + i] = s1 * w.Z[lay + (long)s * Pc * Hp + i];
    }
}

__device__ __forceinline__ void out_head(const WideNet& n, double u0, double& y, double& d1, double& d2) {
    y = u0;
    d1 = 1.0;
    d2 = 0.0;
    if (n.out_transform == 1) {
        double s0, s1, s2;
        act_d012<double>(QEXXC_ACT_SWISH, u0, s0, s1, s2);
        y = -n.out_scale * s0;
        d1 = -n.out_scale * s1;
        d2 = -n.out_scale * s2;
    }
}

__global__ void wide_out_kernel(WideNet n, WideIo io, WideWs w) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= w.Pc) return;
    const long t = io.t0 + p;
    if (t >= io.total) return;
    const long b = t / io.npts, g = t - b * io.npts, o = b * io.io_bstride + g;
    double y, d1, d2;
    out_head(n, w.U[(long)p * 64] + w.bias[(long)n.L * n.Hp], y, d1, d2);
    io.exc[o] = y;
    io.vrho[o] = d1 * w.U[((long)w.Pc + p) * 64];
    if (w.S > 2 && io.vgamma) io.vgamma[o] = d1 * w.U[(2L * w.Pc + p) * 64];
}

// reverse seeds: A64 rows (value | tangent): col 0 = (u_bar | udot_bar), col 1 = (1 | 0), cols 2,3 = (x_f | xdot_f)
__global__ void wide_seed_kernel(WideNet n, WideIo io, WideWs w) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= w.Pc) return;
    const long Pc = w.Pc;
    const bool live = io.t0 + p < io.total;
    const double u0 = w.U[(long)p * 64] + w.bias[(long)n.L * n.Hp], ud = w.U[(Pc + p) * 64];
    double y, d1, d2;
    out_head(n, u0, y, d1, d2);
    const double yb = w.feat[4 * Pc + p], one = live ? 1.0 : 0.0;
    const double ub = yb * d1 + d2 * ud * one, udb = d1 * one;
    double* av = w.A64 + (long)p * 64;
    double* at = w.A64 + (Pc + p) * 64;
    av[0] = ub;
    av[1] = one;
    av[2] = w.feat[p];
    av[3] = w.feat[Pc + p];
    at[0] = udb;
    at[1] = 0.0;
    at[2] = w.feat[2 * Pc + p];
    at[3] = w.feat[3 * Pc + p];
}

// D[r][j] = A64[r][0] * wl[j]   (adjoint of the last hidden activations, both streams)
__global__ void wide_seed2_kernel(WideNet n, WideWs w, double* __restrict__ D) {
    const long Hp = n.Hp, total = 2L * w.Pc * Hp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / Hp, j = i % Hp;
        D[i] = w.A64[r * 64] * w.wlp[j * 64];
    }
}

// in place: (h_bar, hdot_bar) -> (z_bar, zdot_bar) of layer l
__global__ void wide_actbwd_kernel(WideNet n, WideWs w, int l, double* __restrict__ D) {
    const long Hp = n.Hp, Pc = w.Pc, lay = (long)l * 2 * Pc * Hp;
    const long total = Pc * Hp;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        double s0, s1, s2;
        act_d012<double>(n.act, w.Z[lay + i], s0, s1, s2);
        const double zd = w.Z[lay + Pc * Hp + i], hb = D[i], hdb = D[Pc * Hp + i];
        D[i] = hb * s1 + hdb * s2 * zd;
        D[Pc * Hp + i] = hdb * s1;
    }
}

__global__ void wide_xbar_kernel(WideNet n, WideIo io, WideWs w) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= w.Pc) return;
    const long t = io.t0 + p;
    if (t >= io.total) return;
    const long b = t / io.npts, g = t - b * io.npts;
    const double x0b = w.XB[(long)p * 64] * n.in_scale, x1b = w.XB[(long)p * 64 + 1] * n.in_scale;
    double* ob = io.rho_bar + b * io.rho_bstride + g;
    const double* rb = io.rho + b * io.rho_bstride + g;
    ob[0] = (io.accumulate ? ob[0] : 0.0) + x0b;
    if (n.xctype == QEXXC_XC_GGA) {
        for (int k = 1; k <= 3; ++k) {
            double* q = ob + k * io.rho_cstride;
            *q = (io.accumulate ? *q : 0.0) + 2.0 * x1b * rb[k * io.rho_cstride];
        }
    } else if (n.F == 2) {
        double* q = ob + io.rho_cstride;
        *q = (io.accumulate ? *q : 0.0) + x1b;
    }
}

// padded gradient accumulators -> flat theta_bar
__global__ void wide_unpack_kernel(WideNet n, WideWs w, double* __restrict__ tb, long n_theta, int accumulate) {
    const long H = n.H, Hp = n.Hp, F = n.F;
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n_theta; k += (long)gridDim.x * blockDim.x) {
        double v;
        long r = k;
        if (r < F * H) {
            v = w.Gb[(2 + r / H) * Hp + r % H];  // layer-0 reduction rows 2, 3
        } else if ((r -= F * H) < H) {
            v = w.Gb[Hp + r];  // row 1 of layer 0
        } else {
            r -= H;
            const long per = H * H + H;
            const long l = r / per + 1;
            if (l < n.L) {
                const long q = r % per;
                v = q < H * H ? w.Wg[(l - 1) * Hp * Hp + (q / H) * Hp + q % H] : w.Gb[l * 64 * Hp + Hp + (q - H * H)];
            } else {
                const long q = r - (long)(n.L - 1) * per;
                v = q < H ? w.Gwl[q] : w.G2[1];
            }
        }
        tb[k] = (accumulate ? tb[k] : 0.0) + v;
    }
}

inline unsigned blocks_for(long total, int num_sms) {
    long b = (total + 255) / 256;
    const long cap = (long)num_sms * 8;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

// C = beta * C + sum_z part[z]  (fixed order => deterministic)
__global__ void splitk_reduce_kernel(const double* __restrict__ part, int nsplit, long mn, long N, double* __restrict__ C,
                                     long ldc, int beta) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < mn; i += (long)gridDim.x * blockDim.x) {
        const long r = i / N, c = i - r * N;
        double a = beta ? C[r * ldc + c] : 0.0;
        for (int z = 0; z < nsplit; ++z) a += part[(long)z * mn + i];
        C[r * ldc + c] = a;
    }
}

template <bool TA, bool TB>
void gemm(cudaStream_t st, const double* A, long lda, const double* B, long ldb, double* C, long ldc, long M, long N, long K,
          int beta, double* part = nullptr, long part_doubles = 0, int num_sms = 148) {
    dim3 grid((unsigned)(N / WT), (unsigned)(M / WT), 1);
    const long tiles = (long)grid.x * grid.y;
    // few output tiles and a long K (the weight-gradient and reduction products): split K over blockIdx.z into partial
    // outputs, then add them in a fixed order
    int nsplit = 1;
    if (part && tiles < num_sms && K >= 1024) {
        nsplit = (int)((2L * num_sms + tiles - 1) / tiles);
        if (nsplit > K / 256) nsplit = (int)(K / 256);
        if ((long)nsplit * M * N > part_doubles) nsplit = (int)(part_doubles / (M * N));
    }
    if (nsplit <= 1) {
        gemm64_kernel<TA, TB><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, (int)K, beta, (int)K, 0);
        return;
    }
    long kchunk = ((K + nsplit - 1) / nsplit + KS - 1) / KS * KS;
    nsplit = (int)((K + kchunk - 1) / kchunk);
    grid.z = nsplit;
    gemm64_kernel<TA, TB><<<grid, 256, 0, st>>>(A, lda, B, ldb, part, N, (int)K, 0, (int)kchunk, M * N);
    splitk_reduce_kernel<<<blocks_for(M * N, num_sms), 256, 0, st>>>(part, nsplit, M * N, N, C, ldc, beta);
}

WideNet wide_net(const qexxc_ctx* c, int xctype) {
    WideNet n{};
    n.F = c->net.n_features;
    n.L = c->net.n_hidden;
    n.H = c->net.width;
    n.Hp = round_up(c->net.width, WT);
    n.act = c->net.activation;
    n.out_transform = c->net.out_transform;
    n.xctype = xctype;
    n.in_scale = c->net.in_scale;
    n.out_scale = c->net.out_scale;
    return n;
}

long wide_chunk(const qexxc_ctx* c) {
    const long Hp = round_up(c->net.width, WT), L = c->net.n_hidden;
    const long per_pt = 2 * L * 3 * Hp * 8;  // Z and Hh of every layer, three streams
    long pc = ((512L << 20) / per_pt) / WT * WT;
    const long cap = round_up((long)c->B * c->GpadMax, WT);
    if (pc > 16384) pc = 16384;
    if (pc > cap) pc = cap;
    if (pc < WT) pc = WT;
    return pc;
}

WideWs wide_carve(const qexxc_ctx* c, double* base, size_t* total) {
    WideWs w{};
    const long Hp = round_up(c->net.width, WT), L = c->net.n_hidden, Pc = wide_chunk(c);
    w.Hp = (int)Hp;
    w.L = (int)L;
    w.S = 3;
    w.Pc = (int)Pc;
    size_t off = 0;
    auto take = [&](size_t n) {
        double* p = base ? base + off : nullptr;
        off += (n + 1) & ~(size_t)1;
        return p;
    };
    w.Wp = take((size_t)(L - 1) * Hp * Hp);
    w.W1p = take((size_t)64 * Hp);   // pack_kernel writes Wp | W1p | wlp | bias through separate pointers
    w.wlp = take((size_t)Hp * 64);
    w.bias = take((size_t)(L + 1) * Hp);
    w.Wg = take((size_t)(L - 1) * Hp * Hp);
    w.Z = take((size_t)L * 3 * Pc * Hp);
    w.Hh = take((size_t)L * 3 * Pc * Hp);
    w.D0 = take((size_t)2 * Pc * Hp);
    w.D1 = take((size_t)2 * Pc * Hp);
    w.U = take((size_t)3 * Pc * 64);
    w.A64 = take((size_t)2 * Pc * 64);
    w.XB = take((size_t)Pc * 64);
    w.Gb = take((size_t)L * 64 * Hp);
    w.Gwl = take((size_t)64 * Hp);
    w.G2 = take((size_t)64 * 64);
    w.feat = take((size_t)8 * Pc);
    w.part_doubles = 8L * Hp * Hp > 64L * 64 * Hp ? 8L * Hp * Hp : 64L * 64 * Hp;  // split-K partial outputs
    w.part = take((size_t)w.part_doubles);
    *total = off;
    return w;
}

}  // namespace

bool mlp_is_wide(const qexxc_net_desc& net) {
    return net.kind == QEXXC_NET_LOCAL_MLP && (net.width > 64 || net.n_hidden > 3);
}
size_t mlp_wide_ws_doubles(const qexxc_ctx* c) {
    size_t total = 0;
    wide_carve(c, nullptr, &total);
    return total;
}

static int wide_forward(qexxc_ctx* c, const WideNet& n, const WideWs& w, WideIo& io, int vjp, cudaStream_t st) {
    const long Hp = n.Hp, Pc = w.Pc, S = w.S, L = n.L;
    wide_prep_kernel<<<(unsigned)((Pc + 255) / 256), 256, 0, st>>>(n, io, w, vjp);
    wide_first_kernel<<<blocks_for(Pc * Hp, c->num_sms), 256, 0, st>>>(n, w, vjp);
    for (int l = 1; l < L; ++l) {
        gemm<false, false>(st, w.Hh + (size_t)(l - 1) * S * Pc * Hp, Hp, w.Wp + (size_t)(l - 1) * Hp * Hp, Hp,
                           w.Z + (size_t)l * S * Pc * Hp, Hp, S * Pc, Hp, Hp, 0);
        wide_act_kernel<<<blocks_for(Pc * Hp, c->num_sms), 256, 0, st>>>(n, w, l);
    }
    gemm<false, false>(st, w.Hh + (size_t)(L - 1) * S * Pc * Hp, Hp, w.wlp, 64, w.U, 64, S * Pc, 64, Hp, 0);
    c->launches += 2 * L + 1;
    return QEXXC_OK;
}

int launch_mlp_wide_fwd(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                        const double* theta, double* exc, double* vrho, double* vgamma, long out_bstride, int nbatch,
                        long npts_per_batch, cudaStream_t st) {
    if (!c->wide_ws) {
        set_error("LocalMLP: the wide-network workspace was not allocated");
        return QEXXC_ERR_STATE;
    }
    ProfScope prof(c, QEXXC_PROF_XC_FWD, st);
    const WideNet n = wide_net(c, xctype);
    size_t tot;
    WideWs w = wide_carve(c, c->wide_ws, &tot);
    w.S = 1 + n.F;
    WideIo io{};
    io.rho = rho;
    io.rho_bstride = rho_bstride;
    io.rho_cstride = rho_cstride;
    io.exc = exc;
    io.vrho = vrho;
    io.vgamma = vgamma;
    io.io_bstride = out_bstride;
    io.npts = npts_per_batch;
    io.total = (long)nbatch * npts_per_batch;
    wide_pack_kernel<<<blocks_for((long)(n.L - 1) * n.Hp * n.Hp + 130L * n.Hp, c->num_sms), 256, 0, st>>>(n, theta, w);
    QX_LAUNCH_CHECK(c);
    for (long t0 = 0; t0 < io.total; t0 += w.Pc) {
        io.t0 = t0;
        QX_TRY(wide_forward(c, n, w, io, 0, st));
        wide_out_kernel<<<(unsigned)((w.Pc + 255) / 256), 256, 0, st>>>(n, io, w);
        QX_LAUNCH_CHECK(c);
    }
    return QEXXC_OK;
}

int launch_mlp_wide_vjp(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                        const double* theta, const double* exc_bar, const double* vrho_bar, const double* vgamma_bar,
                        long in_bstride, double* rho_bar, int accumulate, double* theta_bar, int accumulate_theta,
                        int nbatch, long npts_per_batch, cudaStream_t st) {
    if (!c->wide_ws) {
        set_error("LocalMLP: the wide-network workspace was not allocated");
        return QEXXC_ERR_STATE;
    }
    ProfScope prof(c, QEXXC_PROF_XC_VJP, st);
    const WideNet n = wide_net(c, xctype);
    size_t tot;
    WideWs w = wide_carve(c, c->wide_ws, &tot);
    w.S = 2;
    const long Hp = n.Hp, Pc = w.Pc, L = n.L;
    WideIo io{};
    io.rho = rho;
    io.rho_bstride = rho_bstride;
    io.rho_cstride = rho_cstride;
    io.exc_bar = exc_bar;
    io.vrho_bar = vrho_bar;
    io.vgamma_bar = vgamma_bar;
    io.rho_bar = rho_bar;
    io.accumulate = accumulate;
    io.io_bstride = in_bstride;
    io.npts = npts_per_batch;
    io.total = (long)nbatch * npts_per_batch;
    wide_pack_kernel<<<blocks_for((long)(L - 1) * Hp * Hp + 130L * Hp, c->num_sms), 256, 0, st>>>(n, theta, w);
    QX_LAUNCH_CHECK(c);
    int beta = 0;  // gradient accumulators: overwritten by the first chunk, added to by the rest (sequential => deterministic)
    for (long t0 = 0; t0 < io.total; t0 += Pc, beta = 1) {
        io.t0 = t0;
        QX_TRY(wide_forward(c, n, w, io, 1, st));
        wide_seed_kernel<<<(unsigned)((Pc + 255) / 256), 256, 0, st>>>(n, io, w);
        // b_last_bar = sum u_bar = (A64^T A64)[0][1];  wl_bar = (A64^T [H; Hdot])[0][:]
        gemm<true, false>(st, w.A64, 64, w.A64, 64, w.G2, 64, 64, 64, 2 * Pc, beta, w.part, w.part_doubles, c->num_sms);
        gemm<true, false>(st, w.A64, 64, w.Hh + (size_t)(L - 1) * 2 * Pc * Hp, Hp, w.Gwl, Hp, 64, Hp, 2 * Pc, beta, w.part,
                          w.part_doubles, c->num_sms);
        double *D = w.D0, *Dn = w.D1;
        wide_seed2_kernel<<<blocks_for(2 * Pc * Hp, c->num_sms), 256, 0, st>>>(n, w, D);
        for (int l = (int)L - 1; l >= 0; --l) {
            wide_actbwd_kernel<<<blocks_for(Pc * Hp, c->num_sms), 256, 0, st>>>(n, w, l, D);
            // rows of A64^T [z_bar; zdot_bar]: 1 = bias gradient, (l == 0) 2, 3 = first-Dense weight gradients
            gemm<true, false>(st, w.A64, 64, D, Hp, w.Gb + (size_t)l * 64 * Hp, Hp, 64, Hp, 2 * Pc, beta, w.part, w.part_doubles,
                              c->num_sms);
            if (l > 0) {
                gemm<true, false>(st, w.Hh + (size_t)(l - 1) * 2 * Pc * Hp, Hp, D, Hp, w.Wg + (size_t)(l - 1) * Hp * Hp, Hp, Hp,
                                  Hp, 2 * Pc, beta, w.part, w.part_doubles, c->num_sms);
                gemm<false, true>(st, D, Hp, w.Wp + (size_t)(l - 1) * Hp * Hp, Hp, Dn, Hp, 2 * Pc, Hp, Hp, 0);
                double* t = D;
                D = Dn;
                Dn = t;
            } else {
                gemm<false, true>(st, D, Hp, w.W1p, Hp, w.XB, 64, Pc, 64, Hp, 0);  // x_bar[p][f] = z_bar[p] . W1[f]
                wide_xbar_kernel<<<(unsigned)((Pc + 255) / 256), 256, 0, st>>>(n, io, w);
            }
        }
        c->launches += 5 + 4 * L;
        QX_LAUNCH_CHECK(c);
    }
    wide_unpack_kernel<<<blocks_for(c->n_theta, c->num_sms), 256, 0, st>>>(n, w, theta_bar, c->n_theta, accumulate_theta);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

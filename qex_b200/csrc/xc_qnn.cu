// Stage 3 for LocalQNN: a small statevector circuit per grid point, forward and reverse.
//
//   per point x:  |0..0> -RY(x) on every qubit-> hea(n, L): per layer RX,RY,RX on each qubit, then
//   the CNOT ring repeated n times -> sum_i <Z_i>
//     QNN.__call__            qedft/models/quantum/quantum_models.py:115-157
//     direct_gates            qedft/models/quantum/feature_maps.py:184-221
//     hea                     qedft/models/quantum/hardware_ansatz.py:88-147 (ring x n at :143-144)
//     total_magnetization_ops qedft/models/quantum/measurement.py:140-167
//   exc = <O>, vrho = d<O>/dx as in exc_and_vrho_local (trainer_legacy_no_jit.py:56-63).
//
// One warp per grid point.  The 2^n amplitudes live in registers: amplitude i sits in lane i&31,
// slot i>>5 (qubit q is bit n-1-q of i, horqrux axis order).  Gates on lane bits exchange
// partners with __shfl_xor_sync; the n CNOT rings of a layer are one precomputed basis
// permutation applied with indexed shuffles; <Z> sums are warp-shuffle reductions.
//   forward : psi and psi_dot = d psi/dx (the feature map is a product state, so both start in
//             closed form) -> exc, vrho.
//   reverse : adjoint-state method.  lam = exc_bar*O psi + vrho_bar*O psi_dot, lam_dot =
//             vrho_bar*O psi are pulled back gate by gate with U^dagger while psi, psi_dot are
//             un-computed; each parametrised gate contributes
//             Im(<lam|P|psi> + <lam_dot|P|psi_dot>); rho_bar = 2 Re(<lam_0|phi'> + <lam_dot_0|phi''>).
#include "common.cuh"
#include "xc_act.cuh"

namespace qexxc {
namespace {

constexpr int QT = 128;  // threads per CTA (4 warps = 4 points in flight)
constexpr int QMAXP = 3 * 8 * 8;

struct QnnParams {
    int nq, nl;
    const double* x;  // [npts]
    long npts;
    const double* theta;
    double *exc, *vrho;
    const double *exc_bar, *vrho_bar;
    double* x_bar;
    int accumulate;
    double* theta_part;  // [nwarps_total][n_theta]
    const unsigned char* perm;      // out[i] = in[perm[i]]
    const unsigned char* perm_inv;
};

template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <typename T>
__device__ __forceinline__ T shfl_idx_t(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }
template <typename T>
__device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void sincos_t(double a, double& s, double& c) { sincos(a, &s, &c); }
__device__ __forceinline__ void sincos_t(float a, float& s, float& c) { sincosf(a, &s, &c); }

template <int NQ>
struct QCfg {
    static constexpr int A = NQ > 5 ? (1 << (NQ - 5)) : 1;  // amplitudes per lane
    static constexpr int DIM = 1 << NQ;
};

template <typename T, int NQ>
struct State {
    T re[QCfg<NQ>::A], im[QCfg<NQ>::A];
};

// partner amplitudes for a gate on bit position bp
template <typename T, int NQ>
__device__ __forceinline__ void partner(const State<T, NQ>& s, int bp, State<T, NQ>& o) {
    constexpr int A = QCfg<NQ>::A;
    if (bp < 5) {
#pragma unroll
        for (int k = 0; k < A; ++k) {
            o.re[k] = shfl_xor_t(s.re[k], 1 << bp);
            o.im[k] = shfl_xor_t(s.im[k], 1 << bp);
        }
    } else {
#pragma unroll
        for (int k = 0; k < A; ++k) {
            o.re[k] = s.re[k];
            o.im[k] = s.im[k];
        }
#pragma unroll
        for (int sb = 0; sb < (NQ > 5 ? NQ - 5 : 0); ++sb)
            if (bp - 5 == sb) {
#pragma unroll
                for (int k = 0; k < A; ++k) {
                    o.re[k] = s.re[k ^ (1 << sb)];
                    o.im[k] = s.im[k ^ (1 << sb)];
                }
            }
    }
}

// bit value of amplitude (lane, slot k) at position bp
__device__ __forceinline__ int bit_of(int lane, int k, int bp) { return bp < 5 ? (lane >> bp) & 1 : (k >> (bp - 5)) & 1; }

// s <- R_axis(angle) s given partner amplitudes o.  axis 0 = X, 1 = Y.  (c, sn) = cos, sin(angle/2)
template <typename T, int NQ>
__device__ __forceinline__ void rotate(State<T, NQ>& s, const State<T, NQ>& o, int lane, int bp, int axis, T c, T sn) {
    constexpr int A = QCfg<NQ>::A;
#pragma unroll
    for (int k = 0; k < A; ++k) {
        if (axis == 0) {  // new = c m - i sn o
            const T r = c * s.re[k] + sn * o.im[k];
            const T i = c * s.im[k] - sn * o.re[k];
            s.re[k] = r;
            s.im[k] = i;
        } else {  // bit 0: c m - sn o ; bit 1: c m + sn o
            const T sg = bit_of(lane, k, bp) ? sn : -sn;
            s.re[k] = c * s.re[k] + sg * o.re[k];
            s.im[k] = c * s.im[k] + sg * o.im[k];
        }
    }
}

// Im <l|P|s> restricted to this lane's amplitudes; o = partner amplitudes of s
template <typename T, int NQ>
__device__ __forceinline__ T im_lps(const State<T, NQ>& l, const State<T, NQ>& o, int lane, int bp, int axis) {
    constexpr int A = QCfg<NQ>::A;
    T acc = (T)0;
#pragma unroll
    for (int k = 0; k < A; ++k) {
        if (axis == 0) acc += l.re[k] * o.im[k] - l.im[k] * o.re[k];
        else {
            const T v = l.re[k] * o.re[k] + l.im[k] * o.im[k];
            acc += bit_of(lane, k, bp) ? v : -v;
        }
    }
    return acc;
}

template <typename T, int NQ>
__device__ __forceinline__ void permute(State<T, NQ>& s, const unsigned char* perm, int lane) {
    constexpr int A = QCfg<NQ>::A;
    State<T, NQ> o;
#pragma unroll
    for (int k = 0; k < A; ++k) {
        const int src = perm[(lane + 32 * k) & (QCfg<NQ>::DIM - 1)];
        const int sl = src & 31, ss = src >> 5;
        T r = (T)0, i = (T)0;
#pragma unroll
        for (int c = 0; c < A; ++c) {
            const T rr = shfl_idx_t(s.re[c], sl), ii = shfl_idx_t(s.im[c], sl);
            if (ss == c) {
                r = rr;
                i = ii;
            }
        }
        const bool valid = lane + 32 * k < QCfg<NQ>::DIM;
        o.re[k] = valid ? r : (T)0;
        o.im[k] = valid ? i : (T)0;
    }
    s = o;
}

// product feature state and its first two x-derivatives (real): Taylor-series product over qubits
template <typename T, int NQ>
__device__ __forceinline__ void feature(T x, int lane, T (&p0)[QCfg<NQ>::A], T (&p1)[QCfg<NQ>::A], T (&p2)[QCfg<NQ>::A]) {
    constexpr int A = QCfg<NQ>::A;
    T sn, c;
    sincos_t((T)0.5 * x, sn, c);
#pragma unroll
    for (int k = 0; k < A; ++k) {
        const int idx = lane + 32 * k;
        T a0 = (T)1, a1 = (T)0, a2 = (T)0;  // series a0 + a1 t + a2 t^2
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int bit = (idx >> q) & 1;
            // factor f(x+t): bit 0 -> cos((x+t)/2), bit 1 -> sin((x+t)/2)
            const T f0 = bit ? sn : c;
            const T f1 = bit ? (T)0.5 * c : (T)-0.5 * sn;
            const T f2 = (T)-0.125 * f0;
            const T n2 = a0 * f2 + a1 * f1 + a2 * f0;
            const T n1 = a0 * f1 + a1 * f0;
            a0 = a0 * f0;
            a1 = n1;
            a2 = n2;
        }
        const bool valid = idx < QCfg<NQ>::DIM;
        p0[k] = valid ? a0 : (T)0;
        p1[k] = valid ? a1 : (T)0;
        p2[k] = valid ? (T)2 * a2 : (T)0;
    }
}

__device__ __forceinline__ int zsum_of(int idx, int nq) { return nq - 2 * __popc(idx); }

template <typename T, int NQ, bool VJP>
__global__ void __launch_bounds__(QT) qnn_kernel(const QnnParams p) {
    constexpr int A = QCfg<NQ>::A;
    __shared__ double s_theta_cs[2 * QMAXP];  // cos, sin of theta/2
    __shared__ double s_grad[QT / 32][QMAXP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nparam = 3 * NQ * p.nl;
    for (int k = threadIdx.x; k < nparam; k += blockDim.x) {
        double s, c;
        sincos(0.5 * p.theta[k], &s, &c);
        s_theta_cs[2 * k] = c;
        s_theta_cs[2 * k + 1] = s;
    }
    if (VJP)
        for (int k = lane; k < nparam; k += 32) s_grad[warp][k] = 0.0;
    __syncthreads();
    const long gw = (long)blockIdx.x * (QT / 32) + warp, nw = (long)gridDim.x * (QT / 32);

    for (long pt = gw; pt < p.npts; pt += nw) {
        const T x = (T)p.x[pt];
        State<T, NQ> psi, dps, o;
        T ph0[A], ph1[A], ph2[A];
        feature<T, NQ>(x, lane, ph0, ph1, ph2);
#pragma unroll
        for (int k = 0; k < A; ++k) {
            psi.re[k] = ph0[k];
            psi.im[k] = (T)0;
            dps.re[k] = ph1[k];
            dps.im[k] = (T)0;
        }
        // ---- forward through the ansatz ----
        int gk = 0;
        for (int l = 0; l < p.nl; ++l) {
            for (int q = 0; q < NQ; ++q) {
                const int bp = NQ - 1 - q;
#pragma unroll
                for (int r = 0; r < 3; ++r, ++gk) {
                    const T c = (T)s_theta_cs[2 * gk], sn = (T)s_theta_cs[2 * gk + 1];
                    const int axis = r == 1 ? 1 : 0;
                    partner<T, NQ>(psi, bp, o);
                    rotate<T, NQ>(psi, o, lane, bp, axis, c, sn);
                    partner<T, NQ>(dps, bp, o);
                    rotate<T, NQ>(dps, o, lane, bp, axis, c, sn);
                }
            }
            permute<T, NQ>(psi, p.perm, lane);
            permute<T, NQ>(dps, p.perm, lane);
        }
        // ---- observable ----
        T e = (T)0, e1 = (T)0;
#pragma unroll
        for (int k = 0; k < A; ++k) {
            const T oz = (T)zsum_of(lane + 32 * k, NQ);
            e += oz * (psi.re[k] * psi.re[k] + psi.im[k] * psi.im[k]);
            e1 += oz * (dps.re[k] * psi.re[k] + dps.im[k] * psi.im[k]);
        }
        e = warp_sum_t(e);
        e1 = (T)2 * warp_sum_t(e1);
        if (!VJP) {
            if (lane == 0) {
                p.exc[pt] = (double)e;
                if (p.vrho) p.vrho[pt] = (double)e1;
            }
            continue;
        }
        // ---- reverse ----
        const T eb = (T)p.exc_bar[pt], vb = p.vrho_bar ? (T)p.vrho_bar[pt] : (T)0;
        State<T, NQ> lam, lmd;
#pragma unroll
        for (int k = 0; k < A; ++k) {
            const T oz = (T)zsum_of(lane + 32 * k, NQ);
            lam.re[k] = oz * (eb * psi.re[k] + vb * dps.re[k]);
            lam.im[k] = oz * (eb * psi.im[k] + vb * dps.im[k]);
            lmd.re[k] = oz * vb * psi.re[k];
            lmd.im[k] = oz * vb * psi.im[k];
        }
        for (int l = p.nl - 1; l >= 0; --l) {
            permute<T, NQ>(psi, p.perm_inv, lane);
            permute<T, NQ>(dps, p.perm_inv, lane);
            permute<T, NQ>(lam, p.perm_inv, lane);
            permute<T, NQ>(lmd, p.perm_inv, lane);
            for (int q = NQ - 1; q >= 0; --q) {
                const int bp = NQ - 1 - q;
#pragma unroll
                for (int r = 2; r >= 0; --r) {
                    --gk;
                    const T c = (T)s_theta_cs[2 * gk], sn = (T)s_theta_cs[2 * gk + 1];
                    const int axis = r == 1 ? 1 : 0;
                    State<T, NQ> op, od;
                    partner<T, NQ>(psi, bp, op);
                    partner<T, NQ>(dps, bp, od);
                    T gsum = im_lps<T, NQ>(lam, op, lane, bp, axis) + im_lps<T, NQ>(lmd, od, lane, bp, axis);
                    gsum = warp_sum_t(gsum);
                    if (lane == 0) s_grad[warp][gk] += (double)gsum;
                    // un-compute with U^dagger = R(-angle)
                    rotate<T, NQ>(psi, op, lane, bp, axis, c, -sn);
                    rotate<T, NQ>(dps, od, lane, bp, axis, c, -sn);
                    partner<T, NQ>(lam, bp, o);
                    rotate<T, NQ>(lam, o, lane, bp, axis, c, -sn);
                    partner<T, NQ>(lmd, bp, o);
                    rotate<T, NQ>(lmd, o, lane, bp, axis, c, -sn);
                }
            }
        }
        T xb = (T)0;
#pragma unroll
        for (int k = 0; k < A; ++k) xb += lam.re[k] * ph1[k] + lmd.re[k] * ph2[k];
        xb = (T)2 * warp_sum_t(xb);
        if (lane == 0) p.x_bar[pt] = (p.accumulate ? p.x_bar[pt] : 0.0) + (double)xb;
    }
    if (VJP) {
        __syncwarp();
        double* out = p.theta_part + (size_t)gw * nparam;
        for (int k = lane; k < nparam; k += 32) out[k] = s_grad[warp][k];
    }
}

// hardware_ansatz.py:143-144: ring CNOT(control=i, target=(i+1)%n), i = 0..n-1, repeated n times,
// as one permutation: out[idx] = in[perm[idx]]
void build_ring_perm(int n, std::vector<unsigned char>& perm, std::vector<unsigned char>& inv) {
    const int dim = 1 << n;
    std::vector<int> acc(dim), tmp(dim);
    for (int i = 0; i < dim; ++i) acc[i] = i;
    for (int rep = 0; rep < n; ++rep)
        for (int i = 0; i < n; ++i) {
            const int cb = n - 1 - i, tb = n - 1 - ((i + 1) % n);
            for (int idx = 0; idx < dim; ++idx) {
                const int pg = ((idx >> cb) & 1) ? (idx ^ (1 << tb)) : idx;
                tmp[idx] = acc[pg];
            }
            acc = tmp;
        }
    perm.resize(256);
    inv.resize(256);
    for (int i = 0; i < 256; ++i) perm[i] = inv[i] = 0;
    for (int i = 0; i < dim; ++i) {
        perm[i] = (unsigned char)acc[i];
        inv[acc[i]] = (unsigned char)i;
    }
}

}  // namespace

int qnn_grid(const qexxc_ctx* c, long npts) {
    long blocks = (npts + (QT / 32) - 1) / (QT / 32);
    const long cap = (long)c->num_sms * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}
size_t qnn_red_doubles(const qexxc_ctx* c, long npts_max) { return (size_t)qnn_grid(c, npts_max) * (QT / 32) * c->n_theta; }

int qnn_upload_perm(qexxc_ctx* c, unsigned char* dev_tables /* 512 bytes */) {
    std::vector<unsigned char> perm, inv;
    build_ring_perm(c->net.width, perm, inv);
    QX_CUDA(cudaMemcpy(dev_tables, perm.data(), 256, cudaMemcpyHostToDevice));
    QX_CUDA(cudaMemcpy(dev_tables + 256, inv.data(), 256, cudaMemcpyHostToDevice));
    return QEXXC_OK;
}

__global__ void theta_reduce_kernel3(const double* __restrict__ part, int nparts, long n, double* __restrict__ out,
                                     int accumulate) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double a = accumulate ? out[k] : 0.0;
    for (int c = 0; c < nparts; ++c) a += part[(size_t)c * n + k];
    out[k] = a;
}

template <typename T, bool VJP>
static int dispatch_nq(int nq, int grid, const QnnParams& p, cudaStream_t st) {
    switch (nq) {
#define QX_CASE(N)                                           \
    case N:                                                  \
        qnn_kernel<T, N, VJP><<<grid, QT, 0, st>>>(p);      \
        break;
        QX_CASE(2)
        QX_CASE(3)
        QX_CASE(4)
        QX_CASE(5)
        QX_CASE(6)
        QX_CASE(7)
        QX_CASE(8)
#undef QX_CASE
        default:
            set_error("LocalQNN: n_qubits=%d not supported (2..8)", nq);
            return QEXXC_ERR_UNSUPPORTED;
    }
    return QEXXC_OK;
}

int launch_qnn(qexxc_ctx* c, bool vjp, const unsigned char* tables, const double* x, long npts, const double* theta,
               double* exc, double* vrho, const double* exc_bar, const double* vrho_bar, double* x_bar,
               int accumulate, double* theta_bar, int accumulate_theta, cudaStream_t st) {
    if (c->net.n_hidden < 1 || 3 * c->net.width * c->net.n_hidden > QMAXP) {
        set_error("LocalQNN: n_layers=%d unsupported (3*n_qubits*n_layers <= %d)", c->net.n_hidden, QMAXP);
        return QEXXC_ERR_UNSUPPORTED;
    }
    if (npts <= 0) return QEXXC_OK;
    ProfScope prof(c, vjp ? QEXXC_PROF_XC_VJP : QEXXC_PROF_XC_FWD, st);
    QnnParams p{};
    p.nq = c->net.width;
    p.nl = c->net.n_hidden;
    p.x = x;
    p.npts = npts;
    p.theta = theta;
    p.exc = exc;
    p.vrho = vrho;
    p.exc_bar = exc_bar;
    p.vrho_bar = vrho_bar;
    p.x_bar = x_bar;
    p.accumulate = accumulate;
    p.theta_part = c->red;
    p.perm = tables;
    p.perm_inv = tables + 256;
    const int grid = qnn_grid(c, npts);
    const bool f32 = c->net.precision == QEXXC_PREC_F32;
    int rc;
    if (vjp) rc = f32 ? dispatch_nq<float, true>(p.nq, grid, p, st) : dispatch_nq<double, true>(p.nq, grid, p, st);
    else rc = f32 ? dispatch_nq<float, false>(p.nq, grid, p, st) : dispatch_nq<double, false>(p.nq, grid, p, st);
    QX_TRY(rc);
    QX_LAUNCH_CHECK(c);
    if (vjp) {
        theta_reduce_kernel3<<<(unsigned)((c->n_theta + 255) / 256), 256, 0, st>>>(
            c->red, grid * (QT / 32), c->n_theta, theta_bar, accumulate_theta);
        QX_LAUNCH_CHECK(c);
    }
    return QEXXC_OK;
}

}  // namespace qexxc

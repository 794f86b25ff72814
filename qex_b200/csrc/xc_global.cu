// Stage 3 for GlobalMLP ("NN-AmplitudeEncoding"): the network sees the whole density vector.
//
//   forward : exc = sum(MLP(theta, rho / norm))  (scalar),  vrho = d exc / d rho  [G]
//             = exc_and_vrho_global  qedft/train/td/trainer_legacy_no_jit.py:46-53 with
//               build_global_mlp.apply_fn  qedft/models/classical/classical_models.py:216-222
//   reverse : (exc_bar scalar, vrho_bar [G]) -> rho_bar [G], theta_bar: tangent direction
//             v = vrho_bar, then reverse over (value, tangent) with seeds (exc_bar, 1).
//
// One CTA per batch element (molecule); the work is a chain of small matrix-vector products
// (G x H then H x H), float64 throughout.  All reductions have a fixed order.
#include "common.cuh"
#include "xc_act.cuh"

namespace qexxc {
namespace {

constexpr int GT = 512;       // threads
constexpr int GMAXH = 512;    // max hidden width
constexpr int GMAXL = QEXXC_MAX_LAYERS;

struct GlobalParams {
    int G, L, H, act, out_transform;
    double in_scale, out_scale;
    const double* rho;   // [B][ld]
    long ld;
    const double* theta;
    double* exc;         // [B] (stride exc_stride)
    long exc_stride;
    double* vrho;        // [B][ld]
    const double* exc_bar;   // [B] (stride exc_stride)
    const double* vrho_bar;  // [B][ld]
    double* rho_bar;     // [B][ld]
    double* theta_part;  // [B][n_theta]
    long n_theta;
};

__device__ __forceinline__ long goff(int G, int H, int l) {  // offset of Dense l
    long off = 0;
    for (int k = 0; k < l; ++k) off += (long)(k == 0 ? G : H) * H + H;
    return off;
}

// out[j] = bias[j] + sum_i x[i] W[i][j]; outd[j] = sum_i xd[i] W[i][j]   (W [nin][H] row-major)
// x given as scale * xs[i].  tmp: [2][nsl][H] scratch.
__device__ void matvec_fwd(const double* W, const double* bias, int nin, int H, const double* xs, double scale,
                           const double* xds, double* out, double* outd, double* tmp) {
    const int nsl = GT / H;
    const int t = threadIdx.x, sl = t / H, j = t % H;
    if (sl < nsl) {
        double a = 0.0, ad = 0.0;
        const int chunk = (nin + nsl - 1) / nsl;
        const int i0 = sl * chunk, i1 = min(nin, i0 + chunk);
        for (int i = i0; i < i1; ++i) {
            const double w = W[(long)i * H + j];
            a = fma(scale * xs[i], w, a);
            if (xds) ad = fma(scale * xds[i], w, ad);
        }
        tmp[sl * H + j] = a;
        tmp[(nsl + sl) * H + j] = ad;
    }
    __syncthreads();
    if (t < H) {
        double a = bias ? bias[t] : 0.0, ad = 0.0;
        for (int s = 0; s < nsl; ++s) {
            a += tmp[s * H + t];
            ad += tmp[(nsl + s) * H + t];
        }
        out[t] = a;
        if (outd) outd[t] = ad;
    }
    __syncthreads();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool VJP>
__global__ void __launch_bounds__(GT, 1) global_mlp_kernel(const GlobalParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);
    const int H = p.H, L = p.L, G = p.G;
    double* z = sm;                 // [L][H]
    double* zd = z + L * H;         // [L][H]
    double* h = zd + L * H;         // [H] current activations
    double* hd = h + H;             // [H]
    double* a = hd + H;             // [H] adjoint (value stream)
    double* ad = a + H;             // [H] adjoint (tangent stream)
    double* zb = ad + H;            // [H]
    double* zdb = zb + H;           // [H]
    double* tmp = zdb + H;          // [2][GT/H][H] <= 2*GT
    double* sc = tmp + 2 * GT;      // scalars
    const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const double* rho = p.rho + (long)b * p.ld;
    const double* v = VJP ? p.vrho_bar + (long)b * p.ld : nullptr;
    const double* th = p.theta;

    // ---------------- forward ----------------
    for (int l = 0; l < L; ++l) {
        const long off = goff(G, H, l);
        const int nin = l == 0 ? G : H;
        if (l == 0)
            matvec_fwd(th + off, th + off + (long)nin * H, nin, H, rho, p.in_scale, v, z, zd, tmp);
        else
            matvec_fwd(th + off, th + off + (long)nin * H, nin, H, h, 1.0, VJP ? hd : nullptr, z + l * H, zd + l * H, tmp);
        if (t < H) {
            double s0, s1, s2;
            act_d012<double>(p.act, z[l * H + t], s0, s1, s2);
            h[t] = s0;
            hd[t] = VJP ? s1 * zd[l * H + t] : 0.0;
        }
        __syncthreads();
    }
    const long offl = goff(G, H, L);
    {   // output Dense (n_out = 1)
        double pu = 0.0, pud = 0.0;
        for (int j = t; j < H; j += GT) {
            pu = fma(h[j], th[offl + j], pu);
            pud = fma(hd[j], th[offl + j], pud);
        }
        pu = warp_sum(pu);
        pud = warp_sum(pud);
        if (lane == 0) {
            tmp[warp] = pu;
            tmp[32 + warp] = pud;
        }
        __syncthreads();
        if (t == 0) {
            double u = th[offl + H], ud = 0.0;
            for (int w = 0; w < GT / 32; ++w) {
                u += tmp[w];
                ud += tmp[32 + w];
            }
            double y = u, d1 = 1.0, d2 = 0.0;
            if (p.out_transform == 1) {
                double s0, s1, s2;
                act_d012<double>(QEXXC_ACT_SWISH, u, s0, s1, s2);
                y = -p.out_scale * s0;
                d1 = -p.out_scale * s1;
                d2 = -p.out_scale * s2;
            }
            double ub, udb;
            if (VJP) {
                const double yb = p.exc_bar[(long)b * p.exc_stride];
                ub = yb * d1 + d2 * ud;  // seeds (exc_bar, 1)
                udb = d1;
            } else {
                p.exc[(long)b * p.exc_stride] = y;
                ub = d1;  // seed 1 on y
                udb = 0.0;
            }
            sc[0] = ub;
            sc[1] = udb;
        }
        __syncthreads();
    }
    const double ub = sc[0], udb = sc[1];
    double* tp = VJP ? p.theta_part + (long)b * p.n_theta : nullptr;
    if (VJP) {
        for (int j = t; j < H; j += GT) tp[offl + j] = ub * h[j] + udb * hd[j];
        if (t == 0) tp[offl + H] = ub;
    }
    for (int j = t; j < H; j += GT) {
        a[j] = ub * th[offl + j];
        ad[j] = udb * th[offl + j];
    }
    __syncthreads();
    // ---------------- reverse ----------------
    for (int l = L - 1; l >= 0; --l) {
        const long off = goff(G, H, l);
        const int nin = l == 0 ? G : H;
        if (t < H) {
            double s0, s1, s2;
            act_d012<double>(p.act, z[l * H + t], s0, s1, s2);
            zb[t] = a[t] * s1 + ad[t] * s2 * zd[l * H + t];
            zdb[t] = ad[t] * s1;
            if (VJP) tp[off + (long)nin * H + t] = zb[t];
        }
        __syncthreads();
        // inputs of Dense l (value, tangent) for the weight gradient
        if (l > 0 && t < H) {
            double s0, s1, s2;
            act_d012<double>(p.act, z[(l - 1) * H + t], s0, s1, s2);
            h[t] = s0;
            hd[t] = s1 * zd[(l - 1) * H + t];
        }
        __syncthreads();
        // warp per input row i: adjoint of the input and the outer-product weight gradient
        const double* W = th + off;
        for (int i = warp; i < nin; i += GT / 32) {
            double hi, hdi;
            if (l == 0) {
                hi = p.in_scale * rho[i];
                hdi = VJP ? p.in_scale * v[i] : 0.0;
            } else {
                hi = h[i];
                hdi = hd[i];
            }
            double s = 0.0, sd = 0.0;
            for (int j = lane; j < H; j += 32) {
                const double w = W[(long)i * H + j];
                s = fma(w, zb[j], s);
                sd = fma(w, zdb[j], sd);
                if (VJP) tp[off + (long)i * H + j] = hi * zb[j] + hdi * zdb[j];
            }
            s = warp_sum(s);
            sd = warp_sum(sd);
            if (lane == 0) {
                if (l == 0) {
                    if (VJP) p.rho_bar[(long)b * p.ld + i] = p.in_scale * s;
                    else p.vrho[(long)b * p.ld + i] = p.in_scale * s;
                } else {
                    tmp[i] = s;
                    tmp[GT + i] = sd;
                }
            }
        }
        __syncthreads();
        if (l > 0) {
            for (int j = t; j < H; j += GT) {
                a[j] = tmp[j];
                ad[j] = tmp[GT + j];
            }
            __syncthreads();
        }
    }
}

}  // namespace

size_t global_mlp_smem(int L, int H) { return (size_t)(2 * L * H + 6 * H + 2 * GT + 8) * 8; }

int launch_global_mlp(qexxc_ctx* c, bool vjp, const double* rho, long ld, int G, const double* theta, double* exc,
                      long exc_stride, double* vrho, const double* exc_bar, const double* vrho_bar,
                      double* rho_bar, double* theta_bar, int accumulate_theta, int nbatch, cudaStream_t st);

__global__ void theta_reduce_kernel2(const double* __restrict__ part, int nparts, long n, double* __restrict__ out,
                                     int accumulate) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double a = accumulate ? out[k] : 0.0;
    for (int c = 0; c < nparts; ++c) a += part[(size_t)c * n + k];
    out[k] = a;
}

int launch_global_mlp(qexxc_ctx* c, bool vjp, const double* rho, long ld, int G, const double* theta, double* exc,
                      long exc_stride, double* vrho, const double* exc_bar, const double* vrho_bar,
                      double* rho_bar, double* theta_bar, int accumulate_theta, int nbatch, cudaStream_t st) {
    const qexxc_net_desc& net = c->net;
    if (net.width < 1 || net.width > GMAXH || net.n_hidden < 1 || net.n_hidden > GMAXL) {
        set_error("GlobalMLP: width=%d (1..%d) / n_layers=%d (1..%d) unsupported", net.width, GMAXH, net.n_hidden, GMAXL);
        return QEXXC_ERR_UNSUPPORTED;
    }
    ProfScope prof(c, vjp ? QEXXC_PROF_XC_VJP : QEXXC_PROF_XC_FWD, st);
    GlobalParams p{};
    p.G = G;
    p.L = net.n_hidden;
    p.H = net.width;
    p.act = net.activation;
    p.out_transform = net.out_transform;
    p.in_scale = net.in_scale;
    p.out_scale = net.out_scale;
    p.rho = rho;
    p.ld = ld;
    p.theta = theta;
    p.exc = exc;
    p.exc_stride = exc_stride;
    p.vrho = vrho;
    p.exc_bar = exc_bar;
    p.vrho_bar = vrho_bar;
    p.rho_bar = rho_bar;
    p.theta_part = c->red;
    // the first Dense has G inputs: the parameter count follows the grid in use (<= the context's capacity)
    const long nth = qexxc_n_params(&net, G);
    p.n_theta = nth;
    const size_t sm = global_mlp_smem(p.L, p.H);
    if (vjp) {
        QX_CUDA(cudaFuncSetAttribute(global_mlp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        global_mlp_kernel<true><<<nbatch, GT, sm, st>>>(p);
        QX_LAUNCH_CHECK(c);
        theta_reduce_kernel2<<<(unsigned)((nth + 255) / 256), 256, 0, st>>>(c->red, nbatch, nth, theta_bar,
                                                                             accumulate_theta);
        QX_LAUNCH_CHECK(c);
    } else {
        QX_CUDA(cudaFuncSetAttribute(global_mlp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        global_mlp_kernel<false><<<nbatch, GT, sm, st>>>(p);
        QX_LAUNCH_CHECK(c);
    }
    return QEXXC_OK;
}

}  // namespace qexxc

// Streaming per-grid-point parts of stage 4 and its VJP, deterministic reductions, layout copies.
//
// Forward (numint_legacy.py:304-308 "NN", :328-332 "NN-AmplitudeEncoding", :189-194 "GGA" with
// _rks_gga_wv0 :485-491):
//   den = rho0*w; nelec += sum(den); excsum += dot(den, exc)  (global: excsum = exc scalar)
//   wv0 = 0.5*w*vrho;  wv_k = 2*w*vgamma*rho_k (GGA)
// Reverse: SURVEY.md a12.
#include "common.cuh"

namespace qexxc {
namespace {

constexpr int PW_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double* sh) {
    // fixed-order tree: warp shuffle, then warp partials summed serially by thread 0
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    return t;  // valid on thread 0
}

__global__ void __launch_bounds__(PW_THREADS)
stage4_fwd_kernel(int xctype, const double* __restrict__ rho, const double* __restrict__ exc,
                  const double* __restrict__ vrho, const double* __restrict__ vgamma,
                  const double* __restrict__ w, double* __restrict__ wv, double* __restrict__ part,
                  int Gpad, long ld, int C) {
    __shared__ double sh[PW_THREADS / 32];
    const int b = blockIdx.y;
    const double* rb = rho + (long)b * C * ld;
    double* wvb = wv + (long)b * C * ld;
    double s_n = 0.0, s_e = 0.0;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < Gpad; g += (long)gridDim.x * blockDim.x) {
        const double wg = w[(long)b * ld + g];
        const double r0 = rb[g];
        const double den = r0 * wg;
        s_n += den;
        if (xctype != QEXXC_XC_NN_GLOBAL) s_e = fma(den, exc[(long)b * ld + g], s_e);
        wvb[g] = 0.5 * wg * vrho[(long)b * ld + g];
        if (xctype == QEXXC_XC_GGA) {
            const double t = 2.0 * wg * vgamma[(long)b * ld + g];
#pragma unroll
            for (int k = 1; k < 4; ++k) wvb[k * ld + g] = t * rb[k * ld + g];
        }
    }
    const double tn = block_sum(s_n, sh);
    const double te = block_sum(s_e, sh);
    if (threadIdx.x == 0) {
        part[((long)b * gridDim.x + blockIdx.x) * 2 + 0] = te;
        part[((long)b * gridDim.x + blockIdx.x) * 2 + 1] = tn;
    }
}

// sums[b] = (excsum, nelec): one warp per molecule, lane-strided partial sums then a shuffle tree
// (a fixed order, so the result is bit-reproducible)
__global__ void stage4_final_kernel(int xctype, const double* __restrict__ part, int nblocks,
                                    const double* __restrict__ exc, long ld, double* __restrict__ sums,
                                    long sums_bstride) {
    const int b = blockIdx.x, lane = threadIdx.x;
    double te = 0.0, tn = 0.0;
    for (int k = lane; k < nblocks; k += 32) {
        te += part[((long)b * nblocks + k) * 2 + 0];
        tn += part[((long)b * nblocks + k) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        te += __shfl_xor_sync(0xffffffffu, te, o);
        tn += __shfl_xor_sync(0xffffffffu, tn, o);
    }
    if (lane != 0) return;
    if (xctype == QEXXC_XC_NN_GLOBAL) te = exc[(long)b * ld];
    sums[(long)b * sums_bstride + 0] = te;
    sums[(long)b * sums_bstride + 1] = tn;
}

__global__ void __launch_bounds__(PW_THREADS)
stage4_vjp_kernel(int xctype, const double* __restrict__ rho, const double* __restrict__ exc,
                  const double* __restrict__ vrho, const double* __restrict__ vgamma,
                  const double* __restrict__ w, const double* __restrict__ e_bar,
                  const double* __restrict__ wvb, double* __restrict__ rho_bar,
                  double* __restrict__ exc_bar, double* __restrict__ vrho_bar,
                  double* __restrict__ vgamma_bar, int Gpad, long ld, int C) {
    const int b = blockIdx.y;
    const double eb = e_bar[b];
    const double* rb = rho + (long)b * C * ld;
    const double* wb = wvb + (long)b * C * ld;
    double* ob = rho_bar + (long)b * C * ld;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < Gpad; g += (long)gridDim.x * blockDim.x) {
        const double wg = w[(long)b * ld + g];
        vrho_bar[(long)b * ld + g] = 0.5 * wg * wb[g];
        if (xctype == QEXXC_XC_NN_GLOBAL) {
            ob[g] = 0.0;
            if (g == 0) exc_bar[(long)b * ld] = eb;
        } else {
            exc_bar[(long)b * ld + g] = eb * wg * rb[g];
            ob[g] = eb * wg * exc[(long)b * ld + g];
        }
        if (xctype == QEXXC_XC_GGA) {
            const double vg = vgamma[(long)b * ld + g];
            double acc = 0.0;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                acc = fma(rb[k * ld + g], wb[k * ld + g], acc);
                ob[k * ld + g] = 2.0 * wg * vg * wb[k * ld + g];
            }
            vgamma_bar[(long)b * ld + g] = 2.0 * wg * acc;
        }
    }
}

// dst[b][c][i] = src[b][c][i], i < n, with independent strides; pads dst to n_pad with 0
__global__ void copy_rows_kernel(double* __restrict__ dst, long dst_bs, long dst_cs, const double* __restrict__ src,
                                 long src_bs, long src_cs, long n, long n_pad) {
    const long c = blockIdx.y, b = blockIdx.z;
    double* d = dst + b * dst_bs + c * dst_cs;
    const double* s = src + b * src_bs + c * src_cs;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (long)gridDim.x * blockDim.x)
        d[i] = i < n ? s[i] : 0.0;
}

// x[npts][F] (point-major, user) <-> feat[F][ld] (feature-major, internal)
__global__ void transpose_in_kernel(double* __restrict__ feat, long ld, const double* __restrict__ x,
                                    long npts, int F, long n_pad) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += (long)gridDim.x * blockDim.x)
        for (int f = 0; f < F; ++f) feat[f * ld + i] = i < npts ? x[i * F + f] : 0.0;
}
__global__ void transpose_out_kernel(double* __restrict__ x, const double* __restrict__ feat, long ld,
                                     long npts, int F) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += (long)gridDim.x * blockDim.x)
        for (int f = 0; f < F; ++f) x[i * F + f] = feat[f * ld + i];
}

unsigned nblk(long n, int num_sms) {
    long b = (n + PW_THREADS - 1) / PW_THREADS;
    const long cap = (long)num_sms * 8;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

int stage4_nblocks(const qexxc_ctx* c) { return (int)nblk(c->GpadMax, c->num_sms); }

int launch_stage4_pointwise(qexxc_ctx* c, int xctype, const double* rho, const double* exc,
                            const double* vrho, const double* vgamma, double* wv, double* sums,
                            long sums_bstride, cudaStream_t st) {
    const unsigned nb = nblk(c->Gpad, c->num_sms);
    dim3 grid(nb, c->B);
    ProfScope prof(c, QEXXC_PROF_STAGE4, st);
    stage4_fwd_kernel<<<grid, PW_THREADS, 0, st>>>(xctype, rho, exc, vrho, vgamma, c->weights, wv, c->red,
                                                  c->Gpad, c->GpadMax, c->C);
    QX_LAUNCH_CHECK(c);
    stage4_final_kernel<<<c->B, 32, 0, st>>>(xctype, c->red, (int)nb, exc, c->GpadMax, sums, sums_bstride);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_stage4_pointwise_vjp(qexxc_ctx* c, int xctype, const double* rho, const double* exc,
                                const double* vrho, const double* vgamma, const double* e_bar,
                                const double* wvb, double* rho_bar, double* exc_bar, double* vrho_bar,
                                double* vgamma_bar, cudaStream_t st) {
    dim3 grid(nblk(c->Gpad, c->num_sms), c->B);
    ProfScope prof(c, QEXXC_PROF_STAGE4, st);
    stage4_vjp_kernel<<<grid, PW_THREADS, 0, st>>>(xctype, rho, exc, vrho, vgamma, c->weights, e_bar, wvb,
                                                  rho_bar, exc_bar, vrho_bar, vgamma_bar, c->Gpad,
                                                  c->GpadMax, c->C);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_copy_rows(qexxc_ctx* c, double* dst, long dst_bstride, long dst_cstride, const double* src,
                     long src_bstride, long src_cstride, int nb, int nc, long n, long n_pad, cudaStream_t st) {
    if (nb <= 0 || nc <= 0 || n_pad <= 0) return QEXXC_OK;
    dim3 grid(nblk(n_pad, c->num_sms), (unsigned)nc, (unsigned)nb);
    copy_rows_kernel<<<grid, PW_THREADS, 0, st>>>(dst, dst_bstride, dst_cstride, src, src_bstride, src_cstride, n,
                                                 n_pad);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_transpose_in(qexxc_ctx* c, double* feat, long ld, const double* x, long npts, int F, long n_pad,
                        cudaStream_t st) {
    transpose_in_kernel<<<nblk(n_pad, c->num_sms), PW_THREADS, 0, st>>>(feat, ld, x, npts, F, n_pad);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}
int launch_transpose_out(qexxc_ctx* c, double* x, const double* feat, long ld, long npts, int F,
                         cudaStream_t st) {
    transpose_out_kernel<<<nblk(npts, c->num_sms), PW_THREADS, 0, st>>>(x, feat, ld, npts, F);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

// Stage 3 for LocalMLP, FP32 network path on the 5th-generation tensor cores (tcgen05 / TMEM).
//
// Same functional and reverse rule as xc_mlp.cu (exc_and_vrho_local, trainer_legacy_no_jit.py:56-63, with
// build_local_mlp.apply_fn, classical_models.py:168-172), for `precision = "f32"` networks of width <= 64 and
// up to 3 hidden layers (the reference default 64 x 3).  What changes is the execution model:
//
//   * a CTA owns 128 grid points per iteration: thread (quadrant q = warp & 3, column chunk cc = warp >> 2) owns
//     point 32 q + lane and neurons [16 cc, 16 cc + 16) of EVERY stream of that point (value, tangent(s)), so the
//     activation epilogue finds z and z-dot in the same thread;
//   * every hidden Dense layer is a tcgen05.mma (kind::f16 on bf16 operands, M = 128 points, N = 64, K = 64) per
//     stream, issued by one thread, operands in 128-byte-swizzled shared-memory planes (csrc/tc05.cuh), FP32
//     accumulators in tensor memory, read back with tcgen05.ld (32 lanes x 16 columns per warp);
//   * FP32 accuracy on the bf16 pipe comes from the exact three-way split x = x1 + x2 + x3 (each part a bf16) and
//     the six products x1 w1 + x1 w2 + x2 w1 + x2 w2 + x1 w3 + x3 w1, smallest first: measured 1.6e-7 relative on
//     K = 64 (scripts/probe/tc_probe16.cu), i.e. FP32-grade.  (3xTF32 costs the same tensor time but MN-major
//     TF32 operands need a different swizzle, so one plane could not serve two views; bf16 planes can);
//   * the first Dense (1 or 2 inputs) and the last (1 output) are CUDA-core work inside the epilogue.
//
// The reverse kernel keeps the same ownership; its products per hidden layer -- back-propagation [z_bar] W^T and the
// weight gradient H^T [z_bar] -- read the SAME planes once as a K-major and once as an MN-major operand, and the
// weight gradients accumulate in tensor memory over all the tiles of the persistent CTA (the tangent-stream half of
// the weight gradient runs on the tensor cores while the threads already work on the next layer's epilogue).
#include "common.cuh"
#include "tc05.cuh"
#include "xc_act.cuh"

namespace qexxc {
namespace {
using namespace tc05;

constexpr int TC_THREADS = 512;
constexpr int TP = 128;  // points per tile
constexpr int HP = 64;   // padded hidden width
constexpr int CW = 16;   // neurons per thread
constexpr uint32_t PL = TP * HP * 2;    // one [128 x 64] bf16 plane: 16 KB
constexpr uint32_t MAT = 3 * PL;       // a matrix = three planes (bf16x3 split): 48 KB
constexpr uint32_t WPL = HP * HP * 2;  // one [64 x 64] bf16 weight plane: 8 KB
constexpr uint32_t WMAT = 3 * WPL;
constexpr int MAXL_TC = 3;

struct TcParams {
    int F, L, H, act, out_transform, xctype;
    float in_scale, out_scale;
    const double* rho;
    long rho_bstride, rho_cstride;
    const double* theta;
    double *exc, *vrho, *vgamma;
    long out_bstride;
    long npts;  // valid points per batch element
    int B, blocks_per_batch, nblocks;
    const double *exc_bar, *vrho_bar, *vgamma_bar;
    double* rho_bar;
    int accumulate;
    float4* tape;
    double* theta_part;
    long n_theta;
};

__device__ __forceinline__ long th_off(int F, int H, int l) {
    long off = 0;
    for (int k = 0; k < l; ++k) off += (long)(k == 0 ? F : H) * H + H;
    return off;
}

// tanh with ~3e-7 relative error: odd Taylor polynomial below 0.25, 1 - 2/(e^{2|x|}+1) above (ex2.approx + rcp.approx)
__device__ __forceinline__ float tanh_f32(float x) {
    const float ax = fabsf(x), x2 = x * x;
    float ps = fmaf(x2, 62.0f / 2835.0f, -17.0f / 315.0f);
    ps = fmaf(ps, x2, 2.0f / 15.0f);
    ps = fmaf(ps, x2, -1.0f / 3.0f);
    ps = fmaf(ps * x2, x, x);
    const float e = exp2f(ax * 2.885390081777927f);  // e^{2|x|}
    const float t = copysignf(1.0f - __fdividef(2.0f, e + 1.0f), x);
    return ax < 0.25f ? ps : t;
}

template <int ACT>
__device__ __forceinline__ void act3(int act, float z, float& s0, float& s1, float& s2) {
    if (ACT == QEXXC_ACT_TANH) {
        s0 = tanh_f32(z);
        s1 = 1.0f - s0 * s0;
        s2 = -2.0f * s0 * s1;
    } else {
        act_d012<float>(act, z, s0, s1, s2);
    }
}

struct TcSmem {
    unsigned char* X;   // matrix region 0 (value stream)
    unsigned char* Y;   // matrix region 1 (tangent stream / previous-layer activations)
    unsigned char* Z;   // matrix region 2 (second tangent stream / zdot_bar)
    unsigned char* W;   // hidden Dense l = 1..L-1: three planes [64 in x 64 out]
    float* W1;          // [2][64]
    float* bias;        // [L][64]
    float* wl;          // [64]
    float* scr;         // [4][128][4] cross-chunk partial sums
    float* red;         // [16 warps][8][16] end-of-kernel reductions (reverse only)
    uint64_t* bar;
    uint32_t* tslot;
};
__host__ __device__ inline size_t tc_smem_bytes(int L, bool vjp) {
    size_t n = 3 * (size_t)MAT + (size_t)(L - 1) * WMAT;
    n += (2 * HP + (size_t)L * HP + HP + 4 * TP * 4) * 4;
    if (vjp) n += (size_t)16 * 8 * 16 * 4;
    n += 16;
    return n + 1024;  // alignment slack
}
__device__ __forceinline__ TcSmem tc_carve(unsigned char* raw, int L, bool vjp) {
    TcSmem s;
    unsigned char* b = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    s.X = b;
    s.Y = b + MAT;
    s.Z = b + 2 * MAT;
    s.W = b + 3 * MAT;
    float* f = (float*)(s.W + (size_t)(L - 1) * WMAT);
    s.W1 = f; f += 2 * HP;
    s.bias = f; f += (size_t)L * HP;
    s.wl = f; f += HP;
    s.scr = f; f += 4 * TP * 4;
    s.red = f;
    if (vjp) f += 16 * 8 * 16;
    s.bar = (uint64_t*)f;
    s.tslot = (uint32_t*)(s.bar + 1);
    return s;
}

__device__ void tc_load_weights(const TcParams& p, const TcSmem& s) {
    const int F = p.F, H = p.H, L = p.L;
    const double* th = p.theta;
    for (int i = threadIdx.x; i < 2 * HP; i += blockDim.x) {
        const int f = i / HP, c = i % HP;
        s.W1[i] = (f < F && c < H) ? (float)th[(long)f * H + c] : 0.0f;
    }
    for (int l = 1; l < L; ++l) {
        const long off = th_off(F, H, l);
        unsigned char* Wb = s.W + (size_t)(l - 1) * WMAT;
        for (int i = threadIdx.x; i < HP * HP; i += blockDim.x) {
            const int r = i / HP, c = i % HP;
            const float v = (r < H && c < H) ? (float)th[off + (long)r * H + c] : 0.0f;
            uint32_t b1, b2, b3;
            split_bf16x3(v, b1, b2, b3);
            const uint32_t o = plane16_off(r, c);
            *(unsigned short*)(Wb + o) = (unsigned short)(b1 >> 16);
            *(unsigned short*)(Wb + WPL + o) = (unsigned short)(b2 >> 16);
            *(unsigned short*)(Wb + 2 * WPL + o) = (unsigned short)(b3 >> 16);
        }
    }
    const long offl = th_off(F, H, L);
    for (int i = threadIdx.x; i < HP; i += blockDim.x) s.wl[i] = i < H ? (float)th[offl + i] : 0.0f;
    for (int i = threadIdx.x; i < L * HP; i += blockDim.x) {
        const int l = i / HP, c = i % HP;
        s.bias[i] = c < H ? (float)th[th_off(F, H, l) + (long)(l == 0 ? F : H) * H + c] : 0.0f;
    }
}

// thread (row r, columns [j0, j0+16)) writes its 16 values, split three ways, into the three planes of matrix `P`
__device__ __forceinline__ void store_planes(unsigned char* P, int r, int j0, const float (&v)[CW]) {
    const uint32_t rowb = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    const int ch0 = j0 >> 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        uint32_t w1[4], w2[4], w3[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uint32_t a1, a2, a3, c1, c2, c3;
            split_bf16x3(v[8 * i + 2 * e], a1, a2, a3);
            split_bf16x3(v[8 * i + 2 * e + 1], c1, c2, c3);
            w1[e] = pack_hi16(a1, c1);
            w2[e] = pack_hi16(a2, c2);
            w3[e] = pack_hi16(a3, c3);
        }
        const uint32_t o = rowb + (uint32_t)(((ch0 + i) ^ (r & 7)) << 4);
        *(uint4*)(P + o) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        *(uint4*)(P + PL + o) = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        *(uint4*)(P + 2 * PL + o) = make_uint4(w3[0], w3[1], w3[2], w3[3]);
    }
}

// the six products of the bf16x3 split, smallest first: (part of A, part of B)
__device__ __constant__ int kPa[6] = {2, 0, 1, 1, 0, 0};
__device__ __constant__ int kPb[6] = {0, 2, 1, 0, 1, 0};

// D[128 x 64] (tmem) = A (matrix at sA) x W_l (matrix at sW, stored [in][out] => MN-major B)
__device__ __forceinline__ void issue_fwd(uint32_t d, uint32_t sA, uint32_t sW) {
    constexpr uint32_t id = idesc_bf16(TP, HP, 0, 1);
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_bf16(d, desc16_k(sA + kPa[t] * PL, k), desc16_mn(sW + kPb[t] * WPL, HP, k), id, (t | k) != 0);
}
// D[128 x 64] = A x W_l^T  (B[n = in][k = out] = W[in][out] => K-major B)
__device__ __forceinline__ void issue_bwd(uint32_t d, uint32_t sA, uint32_t sW) {
    constexpr uint32_t id = idesc_bf16(TP, HP, 0, 0);
#pragma unroll
    for (int t = 0; t < 6; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k)
            mma_bf16(d, desc16_k(sA + kPa[t] * PL, k), desc16_k(sW + kPb[t] * WPL, k), id, (t | k) != 0);
}
// weight gradient Y^T X over the 128 points of the tile (both operands MN-major):
//   main accumulator (M = 128 spans planes 1|2 of Y): rows 0..63 += Y1^T (X1+X2+X3), rows 64..127 += Y2^T (X1+X2+X3)
//   aux accumulator (M = 64): += Y3^T X1            -- together (Y1+Y2+Y3)^T (X1+X2+X3) up to O(2^-24) terms
__device__ __forceinline__ void issue_wgrad(uint32_t dmain, uint32_t daux, uint32_t sY, uint32_t sX, uint32_t accumulate) {
    constexpr uint32_t id = idesc_bf16(TP, HP, 1, 1), id64 = idesc_bf16(64, HP, 1, 1);
#pragma unroll
    for (int pb = 2; pb >= 0; --pb)
#pragma unroll
        for (int k = 0; k < 8; ++k)
            mma_bf16(dmain, desc16_mn(sY, TP, k), desc16_mn(sX + pb * PL, TP, k), id, accumulate | (uint32_t)(pb != 2 || k != 0));
#pragma unroll
    for (int k = 0; k < 8; ++k)
        mma_bf16(daux, desc16_mn(sY + 2 * PL, TP, k), desc16_mn(sX, TP, k), id64, accumulate | (uint32_t)(k != 0));
}

__device__ __forceinline__ void tc_block_to_bg(const TcParams& p, int blk, int& b, long& g0) {
    b = blk / p.blocks_per_batch;
    g0 = (long)(blk - b * p.blocks_per_batch) * TP;
}

// features of point (b, g): x[0] = rho, x[1] = sigma (GGA) or the second feature row; rk = grad rho (GGA)
__device__ __forceinline__ void load_features(const TcParams& p, int b, long g, bool live, float (&x)[2], double (&rk)[3]) {
    x[0] = x[1] = 0.0f;
    rk[0] = rk[1] = rk[2] = 0.0;
    if (!live) return;
    const double* rb = p.rho + (long)b * p.rho_bstride + g;
    const double r0 = rb[0];
    double r1 = 0.0;
    if (p.xctype == QEXXC_XC_GGA) {
        rk[0] = rb[p.rho_cstride];
        rk[1] = rb[2 * p.rho_cstride];
        rk[2] = rb[3 * p.rho_cstride];
        r1 = rk[0] * rk[0] + rk[1] * rk[1] + rk[2] * rk[2];
    } else if (p.F == 2) {
        r1 = rb[p.rho_cstride];
    }
    x[0] = (float)((double)p.in_scale * r0);
    x[1] = (float)((double)p.in_scale * r1);
}

// =================================================================================================
// forward: exc, vrho (, vgamma)
// =================================================================================================
template <int NT, int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_fwd_kernel(const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const TcSmem s = tc_carve(smem_raw, p.L, false);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cc = warp >> 2, pt = 32 * q + lane, j0 = CW * cc;
    const int L = p.L, act = p.act;

    tc_load_weights(p, s);
    if (tid == 0) {
        mbar_init(s.bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(s.tslot, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *s.tslot;
    const uint32_t sX = smem_u32(s.X), sY = smem_u32(s.Y), sZ = smem_u32(s.Z), sW = smem_u32(s.W);
    uint32_t phase = 0;

    for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
        int b;
        long g0;
        tc_block_to_bg(p, blk, b, g0);
        const bool live = g0 + pt < p.npts;
        float x[2];
        double rk[3];
        load_features(p, b, g0 + pt, live, x, rk);
        // ---- first Dense + activation on the CUDA cores ----
        float h[CW], hd[NT][CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const float w0 = s.W1[j0 + j], w1 = s.W1[HP + j0 + j];
            const float z = fmaf(x[1], w1, fmaf(x[0], w0, s.bias[j0 + j]));
            float s0, s1, s2;
            act3<ACT>(act, z, s0, s1, s2);
            h[j] = s0;
            hd[0][j] = s1 * (p.in_scale * w0);
            if (NT == 2) hd[NT - 1][j] = s1 * (p.in_scale * w1);
        }
        // ---- hidden Dense layers on the tensor cores ----
        for (int l = 1; l < L; ++l) {
            const uint32_t sWl = sW + (uint32_t)(l - 1) * WMAT;
            store_planes(s.X, pt, j0, h);
            store_planes(s.Y, pt, j0, hd[0]);
            if (NT == 2) store_planes(s.Z, pt, j0, hd[NT - 1]);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_fwd(tm, sX, sWl);
                issue_fwd(tm + HP, sY, sWl);
                if (NT == 2) issue_fwd(tm + 2 * HP, sZ, sWl);
                mma_commit(s.bar);
            }
            mbar_wait(s.bar, phase);
            phase ^= 1;
            __syncwarp();
            tc_fence_after();
            float z[CW], zd[NT][CW];
            tmem_ld16(tmem_addr(tm, 32 * q, j0), z);
            tmem_ld16(tmem_addr(tm, 32 * q, HP + j0), zd[0]);
            if (NT == 2) tmem_ld16(tmem_addr(tm, 32 * q, 2 * HP + j0), zd[NT - 1]);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                float s0, s1, s2;
                act3<ACT>(act, z[j] + s.bias[l * HP + j0 + j], s0, s1, s2);
                h[j] = s0;
#pragma unroll
                for (int t = 0; t < NT; ++t) hd[t][j] = s1 * zd[t][j];
            }
        }
        // ---- last Dense (one output) ----
        float pd[1 + NT];
#pragma unroll
        for (int t = 0; t <= NT; ++t) pd[t] = 0.0f;
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const float w = s.wl[j0 + j];
            pd[0] = fmaf(h[j], w, pd[0]);
#pragma unroll
            for (int t = 0; t < NT; ++t) pd[1 + t] = fmaf(hd[t][j], w, pd[1 + t]);
        }
        {
            float4 v = make_float4(pd[0], pd[1], NT == 2 ? pd[NT] : 0.0f, 0.0f);
            *(float4*)(s.scr + ((size_t)cc * TP + pt) * 4) = v;
        }
        __syncthreads();
        if (cc == 0 && live) {
            float4 a = *(float4*)(s.scr + (size_t)pt * 4);
#pragma unroll
            for (int c2 = 1; c2 < 4; ++c2) {
                const float4 v = *(float4*)(s.scr + ((size_t)c2 * TP + pt) * 4);
                a.x += v.x;
                a.y += v.y;
                a.z += v.z;
            }
            const float u0 = a.x + (float)p.theta[th_off(p.F, p.H, L) + p.H];
            float y = u0, d1 = 1.0f;
            if (p.out_transform == 1) {
                float s0, s1, s2;
                act_d012<float>(QEXXC_ACT_SWISH, u0, s0, s1, s2);
                y = -p.out_scale * s0;
                d1 = -p.out_scale * s1;
            }
            const long o = (long)b * p.out_bstride + g0 + pt;
            p.exc[o] = (double)y;
            p.vrho[o] = (double)(d1 * a.y);
            if (NT == 2 && p.vgamma) p.vgamma[o] = (double)(d1 * a.z);
        }
        __syncthreads();  // scr is rewritten by the next tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 256);
}

// column sums over the 32 lanes of a warp: lane ends with S[(lane >> 1) & 15] = sum over lanes of v[that index]
// (fixed butterfly => bitwise deterministic)
__device__ __forceinline__ float colsum16(const float (&v)[CW], int lane) {
    float a[8], b[4], c[2];
    bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = up ? v[i + 8] : v[i], send = up ? v[i] : v[i + 8];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = up ? a[i + 4] : a[i], send = up ? a[i] : a[i + 4];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = up ? b[i + 2] : b[i], send = up ? b[i] : b[i + 2];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    up = lane & 2;
    const float keep = up ? c[1] : c[0], send = up ? c[0] : c[1];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// tape slot of (layer l, chunk cc, quad k, point pt): a warp writes 32 consecutive float4 (512 bytes)
__device__ __forceinline__ float4* tape_at(float4* tape, int l, int cc, int k, int pt) {
    return tape + (((size_t)l * 4 + cc) * 8 + k) * TP + pt;
}
__device__ __forceinline__ void tape_put(float4* tape, int l, int cc, int pt, const float (&a)[CW], const float (&bq)[CW]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        __stcg(tape_at(tape, l, cc, k, pt), make_float4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]));
        __stcg(tape_at(tape, l, cc, 4 + k, pt), make_float4(bq[4 * k], bq[4 * k + 1], bq[4 * k + 2], bq[4 * k + 3]));
    }
}
__device__ __forceinline__ void tape_get(float4* tape, int l, int cc, int pt, float (&a)[CW], float (&bq)[CW]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float4 u = __ldcg(tape_at(tape, l, cc, k, pt));
        const float4 w = __ldcg(tape_at(tape, l, cc, 4 + k, pt));
        a[4 * k] = u.x; a[4 * k + 1] = u.y; a[4 * k + 2] = u.z; a[4 * k + 3] = u.w;
        bq[4 * k] = w.x; bq[4 * k + 1] = w.y; bq[4 * k + 2] = w.z; bq[4 * k + 3] = w.w;
    }
}

// =================================================================================================
// reverse: (exc_bar, vrho_bar[, vgamma_bar]) -> rho_bar, theta_bar partials
// =================================================================================================
template <int ACT>
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_vjp_kernel(const TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const TcSmem s = tc_carve(smem_raw, p.L, true);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3, cc = warp >> 2, pt = 32 * q + lane, j0 = CW * cc;
    const int L = p.L, act = p.act, F = p.F, H = p.H;
    float4* tape = p.tape + (size_t)blockIdx.x * L * 4 * 8 * TP;

    tc_load_weights(p, s);
    if (tid == 0) {
        mbar_init(s.bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(s.tslot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *s.tslot;
    const uint32_t sX = smem_u32(s.X), sY = smem_u32(s.Y), sZ = smem_u32(s.Z), sW = smem_u32(s.W);
    const float bl = (float)p.theta[th_off(F, H, L) + H];
    uint32_t phase = 0;
    // per-thread gradient accumulators: lane pair (2i, 2i+1) carries neuron j0 + i
    float bacc[MAXL_TC], wlacc = 0.0f, w1acc[2] = {0.0f, 0.0f}, blacc = 0.0f;
#pragma unroll
    for (int l = 0; l < MAXL_TC; ++l) bacc[l] = 0.0f;
    uint32_t wg_started = 0;  // bit l: the tensor-memory accumulator of Dense l has been written once
    bool pending = false;     // a committed MMA batch (tangent half of a weight gradient) has not been waited for yet

    for (int blk = blockIdx.x; blk < p.nblocks; blk += gridDim.x) {
        int b;
        long g0;
        tc_block_to_bg(p, blk, b, g0);
        const bool live = g0 + pt < p.npts;
        const long o = (long)b * p.out_bstride + g0 + pt;
        float x[2], xd[2] = {0.0f, 0.0f}, yb = 0.0f;
        double rk[3];
        load_features(p, b, g0 + pt, live, x, rk);
        if (live) {
            xd[0] = (float)((double)p.in_scale * p.vrho_bar[o]);
            if (F == 2 && p.vgamma_bar) xd[1] = (float)((double)p.in_scale * p.vgamma_bar[o]);
            yb = (float)(p.xctype == QEXXC_XC_NN_GLOBAL ? p.exc_bar[b] : p.exc_bar[o]);
        }
        // ---------------- forward with tape ----------------
        float h[CW], hd[CW];
        {
            float ta[CW], tz[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const float w0 = s.W1[j0 + j], w1 = s.W1[HP + j0 + j];
                const float z = fmaf(x[1], w1, fmaf(x[0], w0, s.bias[j0 + j]));
                const float zd = fmaf(xd[1], w1, xd[0] * w0);
                float s0, s1, s2;
                act3<ACT>(act, z, s0, s1, s2);
                h[j] = s0;
                hd[j] = s1 * zd;
                ta[j] = ACT == QEXXC_ACT_TANH ? s0 : z;
                tz[j] = zd;
            }
            tape_put(tape, 0, cc, pt, ta, tz);
        }
        for (int l = 1; l < L; ++l) {
            const uint32_t sWl = sW + (uint32_t)(l - 1) * WMAT;
            if (pending) {  // the previous tile's last weight-gradient batch still reads Y and Z
                mbar_wait(s.bar, phase);
                phase ^= 1;
                pending = false;
            }
            store_planes(s.X, pt, j0, h);
            store_planes(s.Y, pt, j0, hd);
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                issue_fwd(tm, sX, sWl);
                issue_fwd(tm + HP, sY, sWl);
                mma_commit(s.bar);
            }
            mbar_wait(s.bar, phase);
            phase ^= 1;
            __syncwarp();
            tc_fence_after();
            float z[CW], zd[CW];
            tmem_ld16(tmem_addr(tm, 32 * q, j0), z);
            tmem_ld16(tmem_addr(tm, 32 * q, HP + j0), zd);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                z[j] += s.bias[l * HP + j0 + j];
                float s0, s1, s2;
                act3<ACT>(act, z[j], s0, s1, s2);
                h[j] = s0;
                hd[j] = s1 * zd[j];
                if (ACT == QEXXC_ACT_TANH) z[j] = s0;
            }
            tape_put(tape, l, cc, pt, z, zd);
        }
        // ---------------- output layer and its adjoint ----------------
        float hb[CW], hdb[CW];
        {
            float pd0 = 0.0f, pd1 = 0.0f;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const float w = s.wl[j0 + j];
                pd0 = fmaf(h[j], w, pd0);
                pd1 = fmaf(hd[j], w, pd1);
            }
            float2* slot = (float2*)(s.scr + ((size_t)cc * TP + pt) * 4);
            *slot = make_float2(pd0, pd1);
            __syncthreads();
            float u0 = bl, ud = 0.0f;
#pragma unroll
            for (int c2 = 0; c2 < 4; ++c2) {
                const float2 v = *(const float2*)(s.scr + ((size_t)c2 * TP + pt) * 4);
                u0 += v.x;
                ud += v.y;
            }
            float ub = yb, udb = live ? 1.0f : 0.0f;  // seeds (exc_bar, 1) on (y, ydot)
            if (p.out_transform == 1) {
                float s0, s1, s2;
                act_d012<float>(QEXXC_ACT_SWISH, u0, s0, s1, s2);
                const float c = -p.out_scale;
                ub = yb * c * s1 + c * s2 * ud * udb;
                udb = c * s1 * udb;
            }
            float g[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const float w = s.wl[j0 + j];
                g[j] = fmaf(udb, hd[j], ub * h[j]);
                hb[j] = ub * w;
                hdb[j] = udb * w;
            }
            wlacc += colsum16(g, lane);
            float bs = ub;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) bs += __shfl_xor_sync(0xffffffffu, bs, o2);
            blacc += bs;
        }
        // ---------------- reverse over the hidden layers ----------------
        for (int l = L - 1; l >= 0; --l) {
            float zb[CW], zdb[CW];
            {
                float ta[CW], tz[CW];
                tape_get(tape, l, cc, pt, ta, tz);
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    float s0, s1, s2;
                    if (ACT == QEXXC_ACT_TANH) {  // the tape holds h = tanh(z)
                        s1 = 1.0f - ta[j] * ta[j];
                        s2 = -2.0f * ta[j] * s1;
                    } else {
                        act_d012<float>(act, ta[j], s0, s1, s2);
                    }
                    zb[j] = fmaf(hdb[j] * s2, tz[j], hb[j] * s1);
                    zdb[j] = hdb[j] * s1;
                }
            }
            const float bsum = colsum16(zb, lane);
#pragma unroll
            for (int k = 0; k < MAXL_TC; ++k)
                if (k == l) bacc[k] += bsum;
            if (l > 0) {
                const uint32_t sWl = sW + (uint32_t)(l - 1) * WMAT;
                const uint32_t wg = tm + 2 * HP + (uint32_t)(l - 1) * 2 * HP;  // main | aux accumulators
                float a0[CW], a1[CW];
                {
                    float ta[CW], tz[CW];
                    tape_get(tape, l - 1, cc, pt, ta, tz);
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        float s0, s1, s2;
                        if (ACT == QEXXC_ACT_TANH) {
                            s0 = ta[j];
                            s1 = 1.0f - s0 * s0;
                        } else {
                            act_d012<float>(act, ta[j], s0, s1, s2);
                        }
                        a0[j] = s0;
                        a1[j] = s1 * tz[j];
                    }
                }
                // round a: both back-propagation products and the value half of the weight gradient
                if (pending) {
                    mbar_wait(s.bar, phase);
                    phase ^= 1;
                    pending = false;
                }
                store_planes(s.X, pt, j0, zb);
                store_planes(s.Z, pt, j0, zdb);
                store_planes(s.Y, pt, j0, a0);
                fence_async_smem();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    issue_bwd(tm, sX, sWl);
                    issue_bwd(tm + HP, sZ, sWl);
                    issue_wgrad(wg, wg + HP, sY, sX, (wg_started >> l) & 1u);
                    mma_commit(s.bar);
                }
                wg_started |= 1u << l;
                mbar_wait(s.bar, phase);
                phase ^= 1;
                // round b: the tangent half of the weight gradient; it runs while the threads go on
                store_planes(s.Y, pt, j0, a1);
                fence_async_smem();
                tc_fence_before();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after();
                    issue_wgrad(wg, wg + HP, sY, sZ, 1u);
                    mma_commit(s.bar);
                }
                pending = true;
                __syncwarp();
                tc_fence_after();
                tmem_ld16(tmem_addr(tm, 32 * q, j0), hb);
                tmem_ld16(tmem_addr(tm, 32 * q, HP + j0), hdb);
                tmem_ld_wait();
            } else {
                // dW_1 [F][H] and the input cotangent
                float g0v[CW], g1v[CW];
                float xb0 = 0.0f, xb1 = 0.0f;
#pragma unroll
                for (int j = 0; j < CW; ++j) {
                    g0v[j] = fmaf(xd[0], zdb[j], x[0] * zb[j]);
                    g1v[j] = fmaf(xd[1], zdb[j], x[1] * zb[j]);
                    xb0 = fmaf(zb[j], s.W1[j0 + j], xb0);
                    xb1 = fmaf(zb[j], s.W1[HP + j0 + j], xb1);
                }
                w1acc[0] += colsum16(g0v, lane);
                if (F == 2) w1acc[1] += colsum16(g1v, lane);
                float2* slot = (float2*)(s.scr + ((size_t)cc * TP + pt) * 4 + 2);
                *slot = make_float2(xb0, xb1);
                __syncthreads();
                if (cc == 0 && live) {
                    float sx0 = 0.0f, sx1 = 0.0f;
#pragma unroll
                    for (int c2 = 0; c2 < 4; ++c2) {
                        const float2 v = *(const float2*)(s.scr + ((size_t)c2 * TP + pt) * 4 + 2);
                        sx0 += v.x;
                        sx1 += v.y;
                    }
                    const double x0b = (double)(sx0 * p.in_scale), x1b = (double)(sx1 * p.in_scale);
                    double* ob = p.rho_bar + (long)b * p.rho_bstride + g0 + pt;
                    ob[0] = (p.accumulate ? ob[0] : 0.0) + x0b;
                    if (p.xctype == QEXXC_XC_GGA) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            double* qq = ob + (long)(k + 1) * p.rho_cstride;
                            *qq = (p.accumulate ? *qq : 0.0) + 2.0 * x1b * rk[k];
                        }
                    } else if (F == 2) {
                        double* qq = ob + p.rho_cstride;
                        *qq = (p.accumulate ? *qq : 0.0) + x1b;
                    }
                }
            }
        }
    }
    // ---------------- per-CTA theta_bar partial ----------------
    double* out = p.theta_part + (size_t)blockIdx.x * p.n_theta;
    if (pending) {
        mbar_wait(s.bar, phase);
        phase ^= 1;
        pending = false;
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if ((lane & 1) == 0) {
        float* r = s.red + (size_t)warp * 8 * 16 + (lane >> 1);
#pragma unroll
        for (int l = 0; l < MAXL_TC; ++l) r[l * 16] = bacc[l];
        r[3 * 16] = wlacc;
        r[4 * 16] = w1acc[0];
        r[5 * 16] = w1acc[1];
        if (lane == 0) r[6 * 16] = blacc;
    }
    // hidden weight gradients: main accumulator rows i (Y1 part) and 64 + i (Y2 part), aux accumulator (M = 64 MMA:
    // row i lives in lane 32 (i / 16) + i % 16) for the Y3 part
    float* stage = (float*)s.X;            // [128][64] floats (32 KB); the planes are free now
    float* stage2 = stage + TP * HP;       // [64][64] floats (16 KB)
    for (int l = 1; l < L; ++l) {
        float v[CW], u[CW];
        tmem_ld16(tmem_addr(tm, 32 * q, 2 * HP + (l - 1) * 2 * HP + j0), v);
        tmem_ld16(tmem_addr(tm, 32 * q, 2 * HP + (l - 1) * 2 * HP + HP + j0), u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CW; ++j) stage[(size_t)pt * HP + j0 + j] = v[j];
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < CW; ++j) stage2[(size_t)(16 * q + lane) * HP + j0 + j] = u[j];
        }
        __syncthreads();
        const long off = th_off(F, H, l);
        for (int i = tid; i < HP * HP; i += TC_THREADS) {
            const int r = i / HP, c = i % HP;
            if (r < H && c < H) out[off + (long)r * H + c] = (double)((stage[i] + stage[HP * HP + i]) + stage2[i]);
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < 6 * HP; i += TC_THREADS) {
        const int k = i / HP, j = i % HP, c2 = j >> 4, jj = j & 15;
        float a = 0.0f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) a += s.red[((size_t)(c2 * 4 + qq) * 8 + k) * 16 + jj];
        if (j >= H) continue;
        if (k < MAXL_TC) {
            if (k < L) out[th_off(F, H, k) + (long)(k == 0 ? F : H) * H + j] = (double)a;
        } else if (k == 3) {
            out[th_off(F, H, L) + j] = (double)a;
        } else if (k - 4 < F) {
            out[(long)(k - 4) * H + j] = (double)a;
        }
    }
    if (tid == 0) {
        float a = 0.0f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) a += s.red[((size_t)qq * 8 + 6) * 16];
        out[th_off(F, H, L) + H] = (double)a;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_free(tm, 512);
}

int tc_supported(const qexxc_net_desc& net) {
    return net.precision == QEXXC_PREC_F32 && net.width >= 1 && net.width <= HP && net.n_hidden >= 1 &&
           net.n_hidden <= MAXL_TC && net.n_features >= 1 && net.n_features <= 2;
}

TcParams tc_base(const qexxc_ctx* c, int xctype) {
    TcParams p{};
    p.F = c->net.n_features;
    p.L = c->net.n_hidden;
    p.H = c->net.width;
    p.act = c->net.activation;
    p.out_transform = c->net.out_transform;
    p.xctype = xctype;
    p.in_scale = (float)c->net.in_scale;
    p.out_scale = (float)c->net.out_scale;
    p.n_theta = c->n_theta;
    return p;
}

template <typename K>
int tc_set_smem(K kernel, size_t bytes) {
    QX_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return QEXXC_OK;
}

}  // namespace

bool mlp_tc_enabled(const qexxc_ctx* c) {
    static const int off = getenv("QEXXC_MLP_TC") ? atoi(getenv("QEXXC_MLP_TC")) == 0 : 0;
    return !off && tc_supported(c->net);
}

int launch_mlp_tc_fwd(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                      const double* theta, double* exc, double* vrho, double* vgamma, long out_bstride, int nbatch,
                      long npts_per_batch, cudaStream_t st) {
    TcParams p = tc_base(c, xctype);
    p.rho = rho;
    p.rho_bstride = rho_bstride;
    p.rho_cstride = rho_cstride;
    p.theta = theta;
    p.exc = exc;
    p.vrho = vrho;
    p.vgamma = vgamma;
    p.out_bstride = out_bstride;
    p.B = nbatch;
    p.npts = npts_per_batch;
    p.blocks_per_batch = (int)((npts_per_batch + TP - 1) / TP);
    p.nblocks = p.blocks_per_batch * nbatch;
    const int grid = p.nblocks < c->num_sms ? p.nblocks : c->num_sms;
    if (grid <= 0) return QEXXC_OK;
    ProfScope prof(c, QEXXC_PROF_XC_FWD, st);
    const size_t sm = tc_smem_bytes(p.L, false);
#define QX_TCF(NTV, A)                                                         \
    do {                                                                       \
        QX_TRY(tc_set_smem(mlp_tc_fwd_kernel<NTV, A>, sm));                    \
        mlp_tc_fwd_kernel<NTV, A><<<grid, TC_THREADS, sm, st>>>(p);            \
    } while (0)
    const bool th = p.act == QEXXC_ACT_TANH;
    if (p.F == 1) {
        if (th) QX_TCF(1, QEXXC_ACT_TANH);
        else QX_TCF(1, -1);
    } else {
        if (th) QX_TCF(2, QEXXC_ACT_TANH);
        else QX_TCF(2, -1);
    }
#undef QX_TCF
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

size_t mlp_tc_tape_bytes(const qexxc_ctx* c) { return (size_t)c->num_sms * c->net.n_hidden * 4 * 8 * TP * sizeof(float4); }

int launch_mlp_tc_vjp(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                      const double* theta, const double* exc_bar, const double* vrho_bar, const double* vgamma_bar,
                      long in_bstride, double* rho_bar, int accumulate, int nbatch, long npts_per_batch, int* grid_out,
                      cudaStream_t st) {
    TcParams p = tc_base(c, xctype);
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
    p.npts = npts_per_batch;
    p.blocks_per_batch = (int)((npts_per_batch + TP - 1) / TP);
    p.nblocks = p.blocks_per_batch * nbatch;
    p.tape = reinterpret_cast<float4*>(c->tape);
    p.theta_part = c->red;
    const int grid = p.nblocks < c->num_sms ? p.nblocks : c->num_sms;
    *grid_out = grid;
    if (grid <= 0) return QEXXC_OK;
    const size_t sm = tc_smem_bytes(p.L, true);
    if (p.act == QEXXC_ACT_TANH) {
        QX_TRY(tc_set_smem(mlp_tc_vjp_kernel<QEXXC_ACT_TANH>, sm));
        mlp_tc_vjp_kernel<QEXXC_ACT_TANH><<<grid, TC_THREADS, sm, st>>>(p);
    } else {
        QX_TRY(tc_set_smem(mlp_tc_vjp_kernel<-1>, sm));
        mlp_tc_vjp_kernel<-1><<<grid, TC_THREADS, sm, st>>>(p);
    }
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

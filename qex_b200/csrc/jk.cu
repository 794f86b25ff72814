// Incore Coulomb / exchange build on the dense N^4 ERI tensor and its reverse mode
// (SURVEY.md 8f row N2).  Replaces the two einsums of the reference's jitted `_dot_eri_dm_s1`
// (qedft/train/td/hf_legacy.py:275-286, called from `get_jk` :452-470 and `get_veff`
// rks_legacy.py:91-117):
//     vj[x,k,l] = sum_ij eri[i,j,k,l] dm[x,j,i]        vk[x,i,l] = sum_jk eri[i,j,k,l] dm[x,j,k]
//
// The tensor is read ONCE for J and K together: 8 N^4 bytes against 4 N^4 flops, so the kernel
// is HBM-bound and everything here is about streaming it with enough loads in flight.
//   - grid = (row chunks of one [k][l] slab) x (contiguous ranges of the slab index p = i*N + j),
//     sized to one CTA per SM (persistent-style, no wave tail);
//   - a thread owns column l and up to JR rows k = k_lo + kg + KG*n of the slab (a warp reads
//     32 consecutive doubles of one row: coalesced; JR independent 8-byte loads in flight per
//     thread) and keeps its J accumulators in registers across the whole p range;
//   - the K multiplier dm[j,k] is warp-uniform in this layout (one broadcast L1 load), and the sum
//     over k happens inside the thread, so K costs ONE accumulator per thread; it is flushed
//     through shared memory whenever i changes and written as per-(range, segment, chunk) rows;
//   - small deterministic reduce kernels finish J and K (no atomics: results are bit-reproducible).
// The reverse mode (cotangents vj_bar, vk_bar -> dm_bar) is the transposed contraction
//     dm_bar[x,j,i] += sum_kl eri[i,j,k,l] vj_bar[x,k,l]    dm_bar[x,j,k] += sum_il eri[i,j,k,l] vk_bar[x,i,l]
// streamed the same way with the slab loop ordered (j, i) so that the K part accumulates in
// registers over i; no symmetry of the ERI tensor is assumed in either direction.
#include <atomic>

#include "common.cuh"

namespace {

using namespace qexxc;

constexpr int JT = 512;  // threads per CTA
constexpr int JR = 16;   // slab rows per thread
constexpr int JW = JT / 32;
constexpr int JQ = 256;  // row-dot results buffered in shared memory between drains (vjp)

std::atomic<long> g_jk_launches{0};

struct JkPlan {
    int TL, KG, rows, nchunk, npr, maxseg;
    long jpart, kpart;  // doubles
};

__host__ __device__ inline long long jk_split(long long total, int parts, int k) { return (total * k) / parts; }

// TL lanes cover the columns l (TL = N rounded up to whole warps), KG = JT / TL thread groups
// interleave the rows; a chunk is KG*JR rows at most, balanced over the chunks.
JkPlan jk_plan(int N, int num_sms) {
    JkPlan p;
    const long NN = (long)N * N;
    p.TL = round_up(N, 32);
    p.KG = JT / p.TL;
    const int max_rows = p.KG * JR;
    p.nchunk = (N + max_rows - 1) / max_rows;
    p.rows = round_up((N + p.nchunk - 1) / p.nchunk, p.KG);
    long npr = num_sms / p.nchunk;
    if (npr < 1) npr = 1;
    if (npr > NN) npr = NN;
    p.npr = (int)npr;
    const long maxlen = (NN + p.npr - 1) / p.npr + 1;
    p.maxseg = (int)((maxlen + N - 1) / N + 1);
    p.jpart = (long)((p.npr > p.nchunk) ? p.npr : p.nchunk) * NN;
    p.kpart = (long)p.npr * p.maxseg * p.nchunk * N;
    return p;
}

// Loads the (up to JR) owned rows, row n at byte offset n*sb.  The loop carries no predicate and no zero
// fill: rows n >= nv re-read row nv-1 (an L1/L2 hit, no DRAM traffic) and are ignored by the caller, so all
// JR loads of a thread are in flight before the first use whatever nv is.
template <bool STREAM>
__device__ __forceinline__ void jk_load_rows(double (&v)[JR], const double* __restrict__ row, unsigned sb, int nvm1) {
    const char* r = reinterpret_cast<const char*>(row);
#pragma unroll
    for (int n = 0; n < JR; ++n) {
        const unsigned idx = (unsigned)(n < nvm1 ? n : nvm1);
        const double* a = reinterpret_cast<const double*>(r + idx * sb);
        v[n] = STREAM ? __ldcs(a) : __ldg(a);
    }
}

// Which rows of the slab this thread owns (blockDim.x = KG * TL threads).
struct JkOwn {
    int kg, l, nv, nvm1, k_lo, k_hi;
    bool lane_ok;
    unsigned sb;  // byte stride between owned rows
    long base, stride;
    __device__ JkOwn(int N, int TL, int KG, int rows, int ce) {
        const int tid = threadIdx.x;
        kg = tid / TL;
        const int lt = tid - kg * TL;
        lane_ok = lt < N;
        l = lt < N ? lt : N - 1;  // idle lanes re-read the last column (same sectors) and never write
        k_lo = ce * rows;
        k_hi = (k_lo + rows < N) ? k_lo + rows : N;
        nv = (k_lo + kg < k_hi) ? (k_hi - k_lo - kg + KG - 1) / KG : 0;  // warp-uniform
        if (nv > JR) nv = JR;
        nvm1 = nv > 0 ? nv - 1 : 0;
        if (nv == 0) kg = 0;  // a row group beyond the ragged last chunk: read valid memory, contribute nothing
        stride = (long)KG * N;
        sb = (unsigned)(KG * N) * 8u;
        base = (long)(k_lo + kg) * N + l;
    }
};

// BATCHED = false is the single-tensor instantiation: no per-molecule base pointers are kept live (at the
// 128-register cap two more 64-bit values are enough to make ptxas split the 16 streaming loads of a slab
// into two batches: 0.27 -> 0.35 ms for the J+K reverse kernel at nao = 120).
template <bool WJ, bool WK, bool BATCHED>
__global__ void __launch_bounds__(JT, 1)
jk_fwd_kernel(const double* __restrict__ eri, const double* __restrict__ dm, int N, int TL, int rows, int npr,
              int maxseg, double* __restrict__ Jpart, double* __restrict__ Kpart, long jstride, long kstride) {
    __shared__ double Sk[JT];
    extern __shared__ double Sd[];  // WK: this chunk's dm multipliers, [N][KG][JR], zero beyond the chunk
    const int ce = blockIdx.x, pr = blockIdx.y, nchunk = gridDim.x, tid = threadIdx.x, KG = blockDim.x / TL;
    const int kg_true = tid / TL;
    const JkOwn o(N, TL, KG, rows, ce);
    const long long NN = (long long)N * N;
    const long long p0 = jk_split(NN, npr, pr), p1 = jk_split(NN, npr, pr + 1);
    // blockIdx.z: molecule of a batch (its own tensor, density matrix and partial buffers)
    if (BATCHED) {
        eri += (long long)blockIdx.z * NN * NN;
        dm += (long long)blockIdx.z * NN;
        Jpart += (long long)blockIdx.z * jstride;
        Kpart += (long long)blockIdx.z * kstride;
    }
    if (WK) {
        // dm[j][k] for the rows k of this chunk, stored so that a thread's JR multipliers are contiguous:
        // the K loop reads them with broadcast LDS at use time instead of holding JR more loads in registers
        const int RP = KG * JR;
        for (int idx = tid; idx < N * RP; idx += blockDim.x) {
            const int jj = idx / RP, r = idx - jj * RP, g = r / JR, n = r - g * JR;
            const int k = o.k_lo + g + KG * n;
            Sd[idx] = (k < o.k_hi) ? dm[(long)jj * N + k] : 0.0;
        }
        __syncthreads();
    }
    double aj[JR], ak[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int n = 0; n < JR; ++n) aj[n] = 0.0;
    int i = (int)(p0 / N), j = (int)(p0 - (long long)i * N), seg = 0;
    const double* __restrict__ row = eri + p0 * NN + o.base;
    for (long long p = p0; p < p1; ++p, row += NN) {
        const double dji = WJ ? __ldg(dm + (long)j * N + i) : 0.0;
        const double2* __restrict__ sd = reinterpret_cast<const double2*>(Sd + ((long)j * KG + kg_true) * JR);
        double v[JR];
        jk_load_rows<true>(v, row, o.sb, o.nvm1);
#pragma unroll
        for (int n = 0; n < JR; n += 2) {
            if (WJ) {
                aj[n] = fma(dji, v[n], aj[n]);
                aj[n + 1] = fma(dji, v[n + 1], aj[n + 1]);
            }
            if (WK) {
                const double2 d = sd[n / 2];
                ak[n & 3] = fma(d.x, v[n], ak[n & 3]);
                ak[(n + 1) & 3] = fma(d.y, v[n + 1], ak[(n + 1) & 3]);
            }
        }
        const bool row_done = (++j == N);
        if (row_done) { j = 0; ++i; }
        if (WK && (row_done || p == p1 - 1)) {
            // sum the KG row groups of every column in a fixed order -> one partial row of K
            Sk[tid] = o.lane_ok ? (ak[0] + ak[1]) + (ak[2] + ak[3]) : 0.0;
            ak[0] = ak[1] = ak[2] = ak[3] = 0.0;
            __syncthreads();
            if (tid < N) {
                double s = 0.0;
                for (int g = 0; g < KG; ++g) s += Sk[g * TL + tid];
                Kpart[(((long)pr * maxseg + seg) * nchunk + ce) * N + tid] = s;
            }
            __syncthreads();
            ++seg;
        }
    }
    if (WJ && o.lane_ok) {
#pragma unroll
        for (int n = 0; n < JR; ++n)
            if (n < o.nv) Jpart[(long long)pr * NN + o.base + n * o.stride] = aj[n];
    }
}

// Index of the range that holds slab p (inverse of jk_split).
__device__ __forceinline__ int jk_range_of(long long p, long long total, int parts) {
    int r = (int)((p * parts) / total);
    while (r + 1 < parts && jk_split(total, parts, r + 1) <= p) ++r;
    while (r > 0 && jk_split(total, parts, r) > p) --r;
    return r;
}

// One launch finishes both outputs, every sum in a fixed order:
//   blocks [0, nblk_flat):  out_flat[e] = sum_r flat[r][e]                                  (J of the forward pass)
//   then one block per row a: out_rows[a][t] = sum over the ranges r that touch row a of rows_part[r][a - a0(r)][c][t]
//       over all chunks c (owner_rows == 0) or from the chunk that owns column t (owner_rows > 0),
//       plus sum_c add[c][a*N + t] when `add` is given (reverse pass: the J part of dm_bar).
__global__ void jk_finish_kernel(const double* __restrict__ flat, int nflat, long NN, double* __restrict__ out_flat,
                                 int nblk_flat, const double* __restrict__ rows_part, int N, int nranges, int maxseg,
                                 int nchunk, int owner_rows, const double* __restrict__ add, int nadd,
                                 double* __restrict__ out_rows, long jstride, long kstride) {
    {  // blockIdx.y: molecule of a batch
        const long long m = blockIdx.y;
        if (flat) flat += m * jstride;
        if (out_flat) out_flat += m * NN;
        if (rows_part) rows_part += m * kstride;
        if (add) add += m * jstride;
        if (out_rows) out_rows += m * NN;
    }
    if ((int)blockIdx.x < nblk_flat) {
        const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (e >= NN) return;
        double s = 0.0;
        for (int r = 0; r < nflat; ++r) s += flat[(long)r * NN + e];
        out_flat[e] = s;
        return;
    }
    const int a = blockIdx.x - nblk_flat;
    int r_lo = 0, r_hi = -1;
    if (rows_part != nullptr) {
        r_lo = jk_range_of((long long)a * N, NN, nranges);
        r_hi = jk_range_of((long long)a * N + N - 1, NN, nranges);
    }
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        double s = 0.0;
        for (int r = r_lo; r <= r_hi; ++r) {
            const long long p0 = jk_split(NN, nranges, r);
            if (jk_split(NN, nranges, r + 1) <= p0) continue;  // empty range
            const int a0 = (int)(p0 / N);
            const double* src = rows_part + (((long)r * maxseg + (a - a0)) * nchunk) * N + t;
            if (owner_rows > 0) {
                s += src[(long)(t / owner_rows) * N];
            } else {
                for (int c = 0; c < nchunk; ++c) s += src[(long)c * N];
            }
        }
        for (int c = 0; c < nadd; ++c) s += add[(long)c * NN + (long)a * N + t];
        out_rows[(long)a * N + t] = s;
    }
}

template <bool WJ, bool WK, bool BATCHED>
__global__ void __launch_bounds__(JT, 1)
jk_vjp_kernel(const double* __restrict__ eri, const double* __restrict__ vjb, const double* __restrict__ vkb, int N,
              int TL, int rows, int nqr, int maxseg, double* __restrict__ DJpart, double* __restrict__ DKpart,
              long jstride, long kstride) {
    __shared__ double Sk[JR * JW];
    __shared__ double Sred[JQ * JW];
    const int ce = blockIdx.x, qr = blockIdx.y, nchunk = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int KG = blockDim.x / TL, WPG = TL / 32, NW = blockDim.x / 32;  // WPG: warps per row group
    const JkOwn o(N, TL, KG, rows, ce);
    const long long NN = (long long)N * N;
    const long long q0 = jk_split(NN, nqr, qr), q1 = jk_split(NN, nqr, qr + 1);
    if (BATCHED) {
        eri += (long long)blockIdx.z * NN * NN;
        if (WJ) vjb += (long long)blockIdx.z * NN;
        if (WK) vkb += (long long)blockIdx.z * NN;
        DJpart += (long long)blockIdx.z * jstride;
        DKpart += (long long)blockIdx.z * kstride;
    }
    double jb[JR], ak[JR];
#pragma unroll
    for (int n = 0; n < JR; ++n) {
        jb[n] = (WJ && o.lane_ok && n < o.nv) ? vjb[o.base + n * o.stride] : 0.0;
        ak[n] = 0.0;
    }
    int j = (int)(q0 / N), i = (int)(q0 - (long long)j * N), seg = 0;
    long long qbase = q0;
    const long long istep = (long long)N * NN;
    const double* __restrict__ row = eri + ((long long)i * N + j) * NN + o.base;
    for (long long q = q0; q < q1; ++q) {
        const double kb = (WK && o.lane_ok) ? __ldg(vkb + (long)i * N + o.l) : 0.0;
        double v[JR];
        jk_load_rows<true>(v, row, o.sb, o.nvm1);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int n = 0; n < JR; ++n) {
            if (WJ) { if (n & 1) s1 = fma(v[n], jb[n], s1); else s0 = fma(v[n], jb[n], s0); }
            if (WK) ak[n] = fma(kb, v[n], ak[n]);
        }
        if (WJ) {
            double s = s0 + s1;
#pragma unroll
            for (int x = 16; x > 0; x >>= 1) s += __shfl_xor_sync(0xffffffffu, s, x);
            if (lane == 0) Sred[(int)(q - qbase) * JW + warp] = s;
            if (q - qbase == JQ - 1 || q == q1 - 1) {
                __syncthreads();
                const int cnt = (int)(q - qbase) + 1;
                if (tid < cnt) {
                    double t = 0.0;
                    for (int w = 0; w < NW; ++w) t += Sred[tid * JW + w];
                    DJpart[(long long)ce * NN + qbase + tid] = t;
                }
                __syncthreads();
                qbase = q + 1;
            }
        }
        const bool row_done = (++i == N);
        if (row_done) { i = 0; ++j; row = eri + (long long)j * NN + o.base; } else { row += istep; }
        if (WK && (row_done || q == q1 - 1)) {
            // dm_bar[j][k] partial: sum over the columns l (lanes, then the WPG warps of the row group)
#pragma unroll
            for (int n = 0; n < JR; ++n) {
                double s = ak[n];
                ak[n] = 0.0;
#pragma unroll
                for (int x = 16; x > 0; x >>= 1) s += __shfl_xor_sync(0xffffffffu, s, x);
                if (lane == 0) Sk[n * JW + warp] = s;
            }
            __syncthreads();
            const int k_hi = (o.k_lo + rows < N) ? o.k_lo + rows : N;
            if (tid < k_hi - o.k_lo) {
                const int g = tid % KG, n = tid / KG;
                double s = 0.0;
                for (int w = 0; w < WPG; ++w) s += Sk[n * JW + g * WPG + w];
                DKpart[(((long)qr * maxseg + seg) * nchunk + ce) * N + o.k_lo + tid] = s;
            }
            __syncthreads();
            ++seg;
        }
    }
}

int jk_device(int device, int* num_sms) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available: libqexxc has no CPU fallback");
        return QEXXC_ERR_NODEVICE;
    }
    QX_ARG(device >= 0 && device < ndev, "device index out of range");
    QX_CUDA(cudaSetDevice(device));
    QX_CUDA(cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, device));
    return QEXXC_OK;
}

#define JK_LAUNCH_CHECK()                                                                              \
    do {                                                                                               \
        g_jk_launches++;                                                                               \
        cudaError_t _e = cudaGetLastError();                                                           \
        if (_e != cudaSuccess) {                                                                       \
            set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return QEXXC_ERR_CUDA;                                                                     \
        }                                                                                              \
    } while (0)

int jk_check_nao(int nao) {
    if (nao > JT) {
        set_error("incore J/K: nao = %d > %d is not supported (the dense tensor would need %.0f GB)", nao, JT,
                  8e-9 * (double)nao * nao * nao * nao);
        return QEXXC_ERR_UNSUPPORTED;
    }
    return QEXXC_OK;
}

}  // namespace

extern "C" {

long qexxc_jk_launch_count(void) { return g_jk_launches.load(); }

int qexxc_jk_workspace_doubles(int device, int nao, long* out) {
    QX_ARG(out != nullptr && nao >= 1, "nao must be >= 1 and out non-null");
    QX_TRY(jk_check_nao(nao));
    int sms = 0;
    QX_TRY(jk_device(device, &sms));
    const JkPlan p = jk_plan(nao, sms);
    *out = p.jpart + p.kpart;
    return QEXXC_OK;
}

static int jk_forward(int device, const double* eri_dev, const double* dm_dev, int nmol, int nset, int nao,
                      int with_j, int with_k, double* vj_dev, double* vk_dev, double* work_dev, long work_doubles,
                      void* stream) {
    QX_ARG(eri_dev && dm_dev && work_dev, "null device pointer");
    QX_ARG(nset >= 1 && nao >= 1 && nmol >= 1, "nmol, nset and nao must be >= 1");
    QX_TRY(jk_check_nao(nao));
    QX_ARG((!with_j || vj_dev) && (!with_k || vk_dev), "output pointer is null for a requested matrix");
    if (!with_j && !with_k) return QEXXC_OK;
    int sms = 0;
    QX_TRY(jk_device(device, &sms));
    const JkPlan p = jk_plan(nao, sms);
    QX_ARG(work_doubles >= (p.jpart + p.kpart) * nmol, "workspace smaller than nmol * qexxc_jk_workspace_doubles");
    cudaStream_t st = (cudaStream_t)stream;
    const long NN = (long)nao * nao;
    double* Jpart = work_dev;
    double* Kpart = work_dev + p.jpart * nmol;
    const dim3 grid(p.nchunk, p.npr, nmol);
    const size_t smem = with_k ? sizeof(double) * nao * p.KG * JR : 0;  // <= 64 KB
    if (with_k) {
        QX_CUDA(cudaFuncSetAttribute(jk_fwd_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        QX_CUDA(cudaFuncSetAttribute(jk_fwd_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        QX_CUDA(cudaFuncSetAttribute(jk_fwd_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        QX_CUDA(cudaFuncSetAttribute(jk_fwd_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    }
#define JK_FWD(WJV, WKV, SM)                                                                                         \
    do {                                                                                                             \
        if (nmol > 1)                                                                                                \
            jk_fwd_kernel<WJV, WKV, true><<<grid, p.KG * p.TL, SM, st>>>(eri_dev, dm, nao, p.TL, p.rows, p.npr,       \
                                                                         p.maxseg, Jpart, Kpart, p.jpart, p.kpart);  \
        else                                                                                                         \
            jk_fwd_kernel<WJV, WKV, false><<<grid, p.KG * p.TL, SM, st>>>(eri_dev, dm, nao, p.TL, p.rows, p.npr,      \
                                                                          p.maxseg, Jpart, Kpart, p.jpart, p.kpart); \
    } while (0)
    for (int x = 0; x < nset; ++x) {  // nset > 1 only with nmol == 1 (several density matrices, one tensor)
        const double* dm = dm_dev + x * NN;
        if (with_j && with_k) JK_FWD(true, true, smem);
        else if (with_j) JK_FWD(true, false, 0);
        else JK_FWD(false, true, smem);
        JK_LAUNCH_CHECK();
        const int nblk_flat = with_j ? (int)((NN + 255) / 256) : 0;
        jk_finish_kernel<<<dim3(nblk_flat + (with_k ? nao : 0), nmol), 256, 0, st>>>(
            Jpart, p.npr, NN, with_j ? vj_dev + x * NN : nullptr, nblk_flat, Kpart, nao, p.npr, p.maxseg, p.nchunk, 0,
            nullptr, 0, with_k ? vk_dev + x * NN : nullptr, p.jpart, p.kpart);
        JK_LAUNCH_CHECK();
    }
    return QEXXC_OK;
}

static int jk_reverse(int device, const double* eri_dev, const double* vj_bar_dev, const double* vk_bar_dev, int nmol,
                      int nset, int nao, double* dm_bar_dev, double* work_dev, long work_doubles, void* stream) {
    QX_ARG(eri_dev && dm_bar_dev && work_dev, "null device pointer");
    QX_ARG(nset >= 1 && nao >= 1 && nmol >= 1, "nmol, nset and nao must be >= 1");
    QX_TRY(jk_check_nao(nao));
    int sms = 0;
    QX_TRY(jk_device(device, &sms));
    const JkPlan p = jk_plan(nao, sms);
    QX_ARG(work_doubles >= (p.jpart + p.kpart) * nmol, "workspace smaller than nmol * qexxc_jk_workspace_doubles");
    cudaStream_t st = (cudaStream_t)stream;
    const long NN = (long)nao * nao;
    const bool wj = vj_bar_dev != nullptr, wk = vk_bar_dev != nullptr;
    if (!wj && !wk) {
        QX_CUDA(cudaMemsetAsync(dm_bar_dev, 0, sizeof(double) * nset * nmol * NN, st));
        return QEXXC_OK;
    }
    double* DJpart = work_dev;
    double* DKpart = work_dev + p.jpart * nmol;
    const dim3 grid(p.nchunk, p.npr, nmol);
    for (int x = 0; x < nset; ++x) {
        const double* vjb = wj ? vj_bar_dev + x * NN : nullptr;
        const double* vkb = wk ? vk_bar_dev + x * NN : nullptr;
#define JK_VJP(WJV, WKV)                                                                                             \
    do {                                                                                                             \
        if (nmol > 1)                                                                                                \
            jk_vjp_kernel<WJV, WKV, true><<<grid, p.KG * p.TL, 0, st>>>(eri_dev, vjb, vkb, nao, p.TL, p.rows, p.npr,  \
                                                                        p.maxseg, DJpart, DKpart, p.jpart, p.kpart); \
        else                                                                                                         \
            jk_vjp_kernel<WJV, WKV, false><<<grid, p.KG * p.TL, 0, st>>>(eri_dev, vjb, vkb, nao, p.TL, p.rows, p.npr, \
                                                                         p.maxseg, DJpart, DKpart, p.jpart, p.kpart);\
    } while (0)
        if (wj && wk) JK_VJP(true, true);
        else if (wj) JK_VJP(true, false);
        else JK_VJP(false, true);
#undef JK_VJP
        JK_LAUNCH_CHECK();
        double* out = dm_bar_dev + x * NN;
        jk_finish_kernel<<<dim3(nao, nmol), 256, 0, st>>>(nullptr, 0, NN, nullptr, 0, wk ? DKpart : nullptr, nao, p.npr,
                                                         p.maxseg, p.nchunk, p.rows, wj ? DJpart : nullptr,
                                                         wj ? p.nchunk : 0, out, p.jpart, p.kpart);
        JK_LAUNCH_CHECK();
    }
    return QEXXC_OK;
}

int qexxc_dot_eri_dm(int device, const double* eri_dev, const double* dm_dev, int nset, int nao, int with_j,
                     int with_k, double* vj_dev, double* vk_dev, double* work_dev, long work_doubles, void* stream) {
    return jk_forward(device, eri_dev, dm_dev, 1, nset, nao, with_j, with_k, vj_dev, vk_dev, work_dev, work_doubles, stream);
}

int qexxc_dot_eri_dm_vjp(int device, const double* eri_dev, const double* vj_bar_dev, const double* vk_bar_dev,
                         int nset, int nao, double* dm_bar_dev, double* work_dev, long work_doubles, void* stream) {
    return jk_reverse(device, eri_dev, vj_bar_dev, vk_bar_dev, 1, nset, nao, dm_bar_dev, work_dev, work_doubles, stream);
}

int qexxc_dot_eri_dm_batched(int device, const double* eri_dev, const double* dm_dev, int nmol, int nao, int with_j,
                             int with_k, double* vj_dev, double* vk_dev, double* work_dev, long work_doubles,
                             void* stream) {
    return jk_forward(device, eri_dev, dm_dev, nmol, 1, nao, with_j, with_k, vj_dev, vk_dev, work_dev, work_doubles, stream);
}

int qexxc_dot_eri_dm_vjp_batched(int device, const double* eri_dev, const double* vj_bar_dev, const double* vk_bar_dev,
                                 int nmol, int nao, double* dm_bar_dev, double* work_dev, long work_doubles,
                                 void* stream) {
    return jk_reverse(device, eri_dev, vj_bar_dev, vk_bar_dev, nmol, 1, nao, dm_bar_dev, work_dev, work_doubles, stream);
}

}  // extern "C"

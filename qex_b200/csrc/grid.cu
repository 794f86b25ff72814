// Becke fuzzy-cell partition of an atom-centred integration grid (grid generation).
//
// The reference builds its grids with pyscf: `Grids(mol); level = 0; becke_scheme = stratmann; build()`
// (qedft/train/td/trainer_legacy_no_jit.py:248-251, :316-317; data_io/td/dataset_generation.py:139-142).  The
// radial / angular tables are a few hundred numbers per element and stay on the host (qex_b200/gen_grid.py);
// the partition is the O(G natm^2) part -- every point needs the cell function of every atom:
//     P_i(r) = prod_{j != i} 1/2 (1 - s(nu_ij)),   mu_ij = (|r - R_i| - |r - R_j|) / |R_i - R_j|,
//     nu_ij = mu_ij + a_ij (1 - mu_ij^2)           (Treutler's atomic-size adjustment, a_ji = -a_ij),
//     w(r)  = vol(r) P_owner(r) / sum_i P_i(r),
// with s = Becke's three-fold iterated polynomial or the Stratmann-Scuseria-Frisch switch (both odd, so the
// factor pyscf multiplies into the partner atom, 1/2 (1 + s(nu_ij)), is the 1/2 (1 - s(nu_ji)) of this form).
//
// One thread per grid point.  The distances to all atoms are computed once and kept in a shared-memory
// column private to the thread ([natm][128] doubles, conflict-free); atom centres and, when they fit, the
// inverse inter-atomic distances sit in shared memory too.  Point data are read once, coalesced; the kernel is
// FP64-FMA bound (natm^2 ~ 25-flop pair terms per point), not HBM bound: 32 B in / 8 B out per point.
#include "common.cuh"

namespace {

using namespace qexxc;

constexpr int kTpb = 128;

__device__ __forceinline__ double switch_becke(double g) {
#pragma unroll
    for (int k = 0; k < 3; ++k) g = (3.0 - g * g) * g * 0.5;
    return g;
}

__device__ __forceinline__ double switch_stratmann(double g) {
    constexpr double a = 0.64;
    const double ma = g / a, ma2 = ma * ma;
    const double p = (1.0 / 16.0) * (ma * (35.0 + ma2 * (-35.0 + ma2 * (21.0 - 5.0 * ma2))));
    return g <= -a ? -1.0 : (g >= a ? 1.0 : p);
}

// INVR_SMEM: the natm x natm table of 1/|R_i - R_j| lives in shared memory, else it is read from `inv_dist`
// in global memory (every lane of a warp reads the same word: one broadcast transaction, L1-resident).
template <int SCHEME, bool INVR_SMEM>
__global__ void __launch_bounds__(kTpb)
becke_partition_kernel(const double* __restrict__ coords, long ngrids, const int* __restrict__ owner,
                       const double* __restrict__ vol, const double* __restrict__ atom_coords,
                       const double* __restrict__ inv_dist, const double* __restrict__ adjust, int natm,
                       double* __restrict__ weights) {
    extern __shared__ double smem[];
    double* ac = smem;                                            // [natm][3]
    double* invr = ac + 3 * natm;                                 // [natm][natm] when INVR_SMEM
    double* dcol = invr + (INVR_SMEM ? natm * natm : 0);          // [natm][kTpb]
    for (int k = threadIdx.x; k < 3 * natm; k += kTpb) ac[k] = atom_coords[k];
    if (INVR_SMEM)
        for (int k = threadIdx.x; k < natm * natm; k += kTpb) invr[k] = inv_dist[k];
    __syncthreads();
    const double* ir = INVR_SMEM ? invr : inv_dist;
    const long stride = (long)gridDim.x * kTpb;
    for (long g = (long)blockIdx.x * kTpb + threadIdx.x; g < ngrids; g += stride) {
        const double x = coords[3 * g], y = coords[3 * g + 1], z = coords[3 * g + 2];
        for (int j = 0; j < natm; ++j) {
            const double dx = x - ac[3 * j], dy = y - ac[3 * j + 1], dz = z - ac[3 * j + 2];
            dcol[j * kTpb + threadIdx.x] = sqrt(dx * dx + dy * dy + dz * dz);
        }
        const int own = owner[g];
        if (SCHEME == 1 && adjust == nullptr) {
            // Stratmann's screening (CPL 257, 213, eq. 15): within 1/2 (1 - a) of the nearest neighbour every mu_own,j
            // is below -a, so P_own = 1 and all other cells vanish -- exactly what the loops below would compute
            double inv_nn = 0.0;
            for (int j = 0; j < natm; ++j) inv_nn = fmax(inv_nn, ir[own * natm + j]);
            if (dcol[own * kTpb + threadIdx.x] * inv_nn < 0.5 * (1.0 - 0.64)) {
                weights[g] = vol[g];
                continue;
            }
        }
        auto cell = [&](int i, int j, double di) -> double {
            double mu = (di - dcol[j * kTpb + threadIdx.x]) * ir[i * natm + j];
            if (adjust != nullptr) mu += adjust[i * natm + j] * (1.0 - mu * mu);
            const double s = SCHEME == 0 ? switch_becke(mu) : switch_stratmann(mu);
            return 0.5 * (1.0 - s);
        };
        double psum = 0.0, pown = 0.0;
        for (int i = 0; i < natm; ++i) {
            const double di = dcol[i * kTpb + threadIdx.x];
            // the factor against the owner first: with Stratmann's switch it is exactly zero for most atoms, which
            // ends their product after one pair term; two interleaved partial products shorten the FP64 chain
            double p0 = i == own ? 1.0 : cell(i, own, di), p1 = 1.0;
            if (p0 != 0.0) {
                for (int j = 0; j < natm; ++j) {
                    if (j == i || j == own) continue;
                    const double c = cell(i, j, di);
                    if (j & 1) p1 *= c;
                    else p0 *= c;
                    if (c == 0.0) break;
                }
            }
            const double p = p0 * p1;
            psum += p;
            if (i == own) pown = p;
        }
        weights[g] = vol[g] * pown / psum;
    }
}

__global__ void inv_dist_kernel(const double* __restrict__ atom_coords, int natm, double* __restrict__ inv_dist) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= natm * natm) return;
    const int i = k / natm, j = k % natm;
    const double dx = atom_coords[3 * i] - atom_coords[3 * j], dy = atom_coords[3 * i + 1] - atom_coords[3 * j + 1],
                 dz = atom_coords[3 * i + 2] - atom_coords[3 * j + 2];
    inv_dist[k] = i == j ? 0.0 : 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
}

long g_grid_launches = 0;

}  // namespace

extern "C" long qexxc_grid_launch_count(void) { return g_grid_launches; }

extern "C" int qexxc_becke_partition(int device, const double* coords_dev, long ngrids, const int* owner_dev,
                                     const double* vol_dev, const double* atom_coords_dev, const double* adjust_dev,
                                     int natm, int scheme, double* inv_dist_work_dev, double* weights_dev,
                                     void* stream) {
    QX_ARG(ngrids >= 0 && natm >= 1, "ngrids must be >= 0 and natm >= 1");
    QX_ARG(scheme == QEXXC_BECKE_ORIGINAL || scheme == QEXXC_BECKE_STRATMANN, "unknown switching function");
    if (ngrids == 0) return QEXXC_OK;
    QX_ARG(coords_dev && owner_dev && vol_dev && atom_coords_dev && inv_dist_work_dev && weights_dev,
           "null device pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        qexxc::set_error("no CUDA device available: libqexxc has no CPU fallback");
        return QEXXC_ERR_NODEVICE;
    }
    QX_ARG(device >= 0 && device < ndev, "device index out of range");
    QX_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    int max_smem = 0, num_sms = 0;
    QX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    QX_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    const size_t base = sizeof(double) * ((size_t)3 * natm + (size_t)natm * kTpb);
    const size_t with_invr = base + sizeof(double) * (size_t)natm * natm;
    const bool invr_smem = with_invr <= (size_t)max_smem;
    const size_t smem = invr_smem ? with_invr : base;
    if (smem > (size_t)max_smem) {
        qexxc::set_error("qexxc_becke_partition keeps one distance per atom and thread in shared memory: natm = %d needs "
                         "%zu B, the device offers %d B", natm, smem, max_smem);
        return QEXXC_ERR_UNSUPPORTED;
    }
    inv_dist_kernel<<<(natm * natm + 255) / 256, 256, 0, st>>>(atom_coords_dev, natm, inv_dist_work_dev);
    ++g_grid_launches;
    long blocks = (ngrids + kTpb - 1) / kTpb;
    const long cap = (long)num_sms * 8;  // grid-stride beyond a few waves
    if (blocks > cap) blocks = cap;
#define QX_BECKE_LAUNCH(S, I)                                                                                       \
    do {                                                                                                            \
        QX_CUDA(cudaFuncSetAttribute(becke_partition_kernel<S, I>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                     (int)smem));                                                                   \
        becke_partition_kernel<S, I><<<(unsigned)blocks, kTpb, smem, st>>>(coords_dev, ngrids, owner_dev, vol_dev,  \
                                                                           atom_coords_dev, inv_dist_work_dev,      \
                                                                           adjust_dev, natm, weights_dev);          \
    } while (0)
    if (scheme == QEXXC_BECKE_ORIGINAL) {
        if (invr_smem) QX_BECKE_LAUNCH(0, true);
        else QX_BECKE_LAUNCH(0, false);
    } else {
        if (invr_smem) QX_BECKE_LAUNCH(1, true);
        else QX_BECKE_LAUNCH(1, false);
    }
#undef QX_BECKE_LAUNCH
    ++g_grid_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        qexxc::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e));
        return QEXXC_ERR_CUDA;
    }
    return QEXXC_OK;
}

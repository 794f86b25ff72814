// Batched generalised symmetric eigensolver for SMALL matrices:
//     A v = w B v,  B SPD,  n <= 16,  thousands of independent problems, one thread each.
// Follows the reference's `generalized_eigh` (qedft/train/td/generalized_eigensolver.py:264-330) step
// by step -- symmetrise A and B, shift B by (eps - lambda_min(B)) if that is positive, Cholesky
// B = L L^T, C = L^-1 A L^-T by two triangular solves, symmetrise, diagonalise, back-transform
// V = L^-T U, eigenvalues ascending -- with the dense LAPACK calls replaced by in-register cyclic Jacobi
// sweeps.  For the 4 x 4 matrices of an H2 / 6-31G dissociation curve a library call per SCF cycle is pure
// launch and host-synchronisation latency (cuSOLVER's batched path checks `info` on the host); this kernel
// needs neither, so the whole SCF cycle stays asynchronous.  Larger matrices stay with cuSOLVER.
#include "common.cuh"

namespace {

using namespace qexxc;

// cyclic Jacobi: a (symmetric, NM x NM storage, n used) -> eigenvalues on its diagonal, u = eigenvectors
template <int NM>
__device__ __forceinline__ void jacobi(double (&a)[NM * NM], double (&u)[NM * NM], int n, bool want_vectors) {
    if (want_vectors) {
#pragma unroll
        for (int i = 0; i < NM; ++i)
#pragma unroll
            for (int j = 0; j < NM; ++j) u[i * NM + j] = (i == j) ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < NM; ++i)
#pragma unroll
            for (int j = 0; j < NM; ++j)
                if (i < n && j < n) {
                    if (i == j) diag += a[i * NM + j] * a[i * NM + j];
                    else off += a[i * NM + j] * a[i * NM + j];
                }
        if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
        for (int p = 0; p < NM - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < NM; ++q) {
                if (q >= n) continue;
                const double apq = a[p * NM + q];
                if (apq == 0.0) continue;
                const double app = a[p * NM + p], aqq = a[q * NM + q];
                const double tau = (aqq - app) / (2.0 * apq);
                const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
#pragma unroll
                for (int k = 0; k < NM; ++k) {  // columns p, q
                    if (k >= n) continue;
                    const double akp = a[k * NM + p], akq = a[k * NM + q];
                    a[k * NM + p] = c * akp - s * akq;
                    a[k * NM + q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < NM; ++k) {  // rows p, q
                    if (k >= n) continue;
                    const double apk = a[p * NM + k], aqk = a[q * NM + k];
                    a[p * NM + k] = c * apk - s * aqk;
                    a[q * NM + k] = s * apk + c * aqk;
                }
                if (want_vectors) {
#pragma unroll
                    for (int k = 0; k < NM; ++k) {
                        if (k >= n) continue;
                        const double ukp = u[k * NM + p], ukq = u[k * NM + q];
                        u[k * NM + p] = c * ukp - s * ukq;
                        u[k * NM + q] = s * ukp + c * ukq;
                    }
                }
            }
    }
}

template <int NM>
__global__ void __launch_bounds__(64)
geigh_kernel(const double* __restrict__ Ain, const double* __restrict__ Bin, int nb, int n, double eps,
             double* __restrict__ wout, double* __restrict__ Vout) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nb) return;
    const double* A = Ain + (long)m * n * n;
    const double* B = Bin + (long)m * n * n;
    double a[NM * NM], b[NM * NM], u[NM * NM];
#pragma unroll
    for (int i = 0; i < NM; ++i)
#pragma unroll
        for (int j = 0; j < NM; ++j) {
            const bool in = i < n && j < n;
            a[i * NM + j] = in ? 0.5 * (A[i * n + j] + A[j * n + i]) : 0.0;
            b[i * NM + j] = in ? 0.5 * (B[i * n + j] + B[j * n + i]) : 0.0;
        }
    // SPD guard: lambda_min(B) by Jacobi on a copy (u doubles as the scratch copy)
    {
#pragma unroll
        for (int i = 0; i < NM * NM; ++i) u[i] = b[i];
        double dummy[NM * NM];
        jacobi<NM>(u, dummy, n, false);
        double lam_min = u[0];
#pragma unroll
        for (int i = 1; i < NM; ++i)
            if (i < n) lam_min = fmin(lam_min, u[i * NM + i]);
        const double shift = lam_min < eps ? eps - lam_min : 0.0;
#pragma unroll
        for (int i = 0; i < NM; ++i)
            if (i < n) b[i * NM + i] += shift;
    }
    // Cholesky B = L L^T (L overwrites the lower triangle of b)
#pragma unroll
    for (int j = 0; j < NM; ++j) {
        if (j >= n) continue;
        double d = b[j * NM + j];
#pragma unroll
        for (int k = 0; k < NM; ++k)
            if (k < j) d -= b[j * NM + k] * b[j * NM + k];
        d = sqrt(d);
        b[j * NM + j] = d;
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            if (i <= j || i >= n) continue;
            double s = b[i * NM + j];
#pragma unroll
            for (int k = 0; k < NM; ++k)
                if (k < j) s -= b[i * NM + k] * b[j * NM + k];
            b[i * NM + j] = s / d;
        }
    }
    // Y = L^-1 A (in place, column by column), then C = L^-1 Y^T, symmetrised
#pragma unroll
    for (int col = 0; col < NM; ++col) {
        if (col >= n) continue;
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            if (i >= n) continue;
            double s = a[i * NM + col];
#pragma unroll
            for (int k = 0; k < NM; ++k)
                if (k < i) s -= b[i * NM + k] * a[k * NM + col];
            a[i * NM + col] = s / b[i * NM + i];
        }
    }
    // now a = Y; C^T = L^-1 Y^T  <=>  solve on the rows of Y
#pragma unroll
    for (int row = 0; row < NM; ++row) {
        if (row >= n) continue;
#pragma unroll
        for (int i = 0; i < NM; ++i) {
            if (i >= n) continue;
            double s = a[row * NM + i];
#pragma unroll
            for (int k = 0; k < NM; ++k)
                if (k < i) s -= b[i * NM + k] * a[row * NM + k];
            a[row * NM + i] = s / b[i * NM + i];
        }
    }
#pragma unroll
    for (int i = 0; i < NM; ++i)
#pragma unroll
        for (int j = i + 1; j < NM; ++j) {
            const double v = 0.5 * (a[i * NM + j] + a[j * NM + i]);
            a[i * NM + j] = v;
            a[j * NM + i] = v;
        }
    jacobi<NM>(a, u, n, true);
    // ascending order (selection sort on the eigenpairs)
#pragma unroll
    for (int i = 0; i < NM - 1; ++i) {
        if (i >= n - 1) continue;
        int best = i;
#pragma unroll
        for (int j = 0; j < NM; ++j)
            if (j > i && j < n && a[j * NM + j] < a[best * NM + best]) best = j;
        if (best != i) {
            const double t = a[i * NM + i];
            a[i * NM + i] = a[best * NM + best];
            a[best * NM + best] = t;
#pragma unroll
            for (int k = 0; k < NM; ++k) {
                const double x = u[k * NM + i];
                u[k * NM + i] = u[k * NM + best];
                u[k * NM + best] = x;
            }
        }
    }
    // V = L^-T U (back substitution, column by column)
    double* w = wout + (long)m * n;
    double* V = Vout + (long)m * n * n;
#pragma unroll
    for (int col = 0; col < NM; ++col) {
        if (col >= n) continue;
        w[col] = a[col * NM + col];
#pragma unroll
        for (int ii = 0; ii < NM; ++ii) {
            const int i = NM - 1 - ii;
            if (i >= n) continue;
            double s = u[i * NM + col];
#pragma unroll
            for (int k = 0; k < NM; ++k)
                if (k > i && k < n) s -= b[k * NM + i] * u[k * NM + col];
            u[i * NM + col] = s / b[i * NM + i];
        }
#pragma unroll
        for (int i = 0; i < NM; ++i)
            if (i < n) V[i * n + col] = u[i * NM + col];
    }
}

}  // namespace

extern "C" int qexxc_generalized_eigh_batched(int device, const double* a_dev, const double* b_dev, int nbatch, int n,
                                              double eps, double* w_dev, double* v_dev, void* stream) {
    QX_ARG(a_dev && b_dev && w_dev && v_dev, "null device pointer");
    QX_ARG(nbatch >= 1 && n >= 1, "nbatch and n must be >= 1");
    if (n > 16) {
        qexxc::set_error("qexxc_generalized_eigh_batched handles n <= 16 (one thread per matrix); n = %d belongs to cuSOLVER", n);
        return QEXXC_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        qexxc::set_error("no CUDA device available: libqexxc has no CPU fallback");
        return QEXXC_ERR_NODEVICE;
    }
    QX_ARG(device >= 0 && device < ndev, "device index out of range");
    QX_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (nbatch + 63) / 64;
    if (n <= 4) geigh_kernel<4><<<blocks, 64, 0, st>>>(a_dev, b_dev, nbatch, n, eps, w_dev, v_dev);
    else if (n <= 8) geigh_kernel<8><<<blocks, 64, 0, st>>>(a_dev, b_dev, nbatch, n, eps, w_dev, v_dev);
    else geigh_kernel<16><<<blocks, 64, 0, st>>>(a_dev, b_dev, nbatch, n, eps, w_dev, v_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        qexxc::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e));
        return QEXXC_ERR_CUDA;
    }
    return QEXXC_OK;
}

// Analytic LDA exchange on the grid (libxc LDA_X, spin-unpolarised), the functional behind pyscf's `xc = "lda"`.
//
// The reference's `nr_rks` sends every xc_code without "NN" to libxc through pyscfad (`ni.eval_xc(xc_code, rho, ...)`,
// qedft/train/td/numint_legacy.py:175-198 and the LDA branch above it); its data generator runs exactly that functional
// (`mf.xc = "lda"`, data_io/td/dataset_generation.py:385-389), and the energy that run prints is the reference-held number
// this repo pins its whole path against (tests/test_zz_pyscf_pin.py).  libxc is third-party C, absent from
// /root/reference; the closed form is Dirac/Slater exchange:
//     exc(rho)  = -3/4 (3/pi)^(1/3) rho^(1/3)            energy per particle
//     vrho(rho) = d(rho exc)/d rho = 4/3 exc             (the libxc convention for vrho, not the NN glue's d exc/d rho)
// Pointwise and HBM-bound: 8 B read, 16 B written per grid point; one thread per point, coalesced.
#include "common.cuh"

namespace {

__global__ void lda_exchange_kernel(const double* __restrict__ rho, long n, double* __restrict__ exc,
                                    double* __restrict__ vrho) {
    const double cx = -0.75 * cbrt(3.0 / 3.14159265358979323846);
    const long stride = (long)gridDim.x * blockDim.x;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        const double r = fmax(rho[g], 0.0);  // quadrature noise can leave -1e-18 in the tails
        const double e = cx * cbrt(r);
        exc[g] = e;
        vrho[g] = (4.0 / 3.0) * e;
    }
}

long g_lda_launches = 0;

}  // namespace

extern "C" long qexxc_lda_launch_count(void) { return g_lda_launches; }

extern "C" int qexxc_lda_exchange(int device, const double* rho_dev, long npts, double* exc_dev, double* vrho_dev,
                                  void* stream) {
    QX_ARG(npts >= 0, "npts must be >= 0");
    if (npts == 0) return QEXXC_OK;
    QX_ARG(rho_dev && exc_dev && vrho_dev, "null device pointer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        qexxc::set_error("no CUDA device available: libqexxc has no CPU fallback");
        return QEXXC_ERR_NODEVICE;
    }
    QX_ARG(device >= 0 && device < ndev, "device index out of range");
    QX_CUDA(cudaSetDevice(device));
    int num_sms = 0;
    QX_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    long blocks = (npts + 255) / 256;
    if (blocks > (long)num_sms * 16) blocks = (long)num_sms * 16;
    lda_exchange_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rho_dev, npts, exc_dev, vrho_dev);
    ++g_lda_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        qexxc::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e));
        return QEXXC_ERR_CUDA;
    }
    return QEXXC_OK;
}

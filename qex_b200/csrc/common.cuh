// Internal declarations shared by the libqexxc translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qexxc.h"

namespace qexxc {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define QX_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            qexxc::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                             cudaGetErrorString(_e));                                         \
            return QEXXC_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)
#define QX_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            qexxc::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,         \
                             cudaGetErrorString(_e));                                         \
            return QEXXC_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)
#define QX_TRY(expr)                                                                          \
    do {                                                                                      \
        int _r = (expr);                                                                      \
        if (_r != QEXXC_OK) return _r;                                                        \
    } while (0)
#define QX_ARG(cond, msg)                                                                     \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            qexxc::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__);        \
            return QEXXC_ERR_ARG;                                                             \
        }                                                                                     \
    } while (0)

constexpr int kGTile = 128;  // grid rows are padded to a multiple of this
constexpr int kNTile = 32;   // AO row pitch (storage) is a multiple of this; pad columns hold real zeros
constexpr int kNBlock = 8;   // compute extents are rounded to one DMMA block only
inline int round_up(long x, int m) { return (int)(((x + m - 1) / m) * m); }

// One contracted shell of the basis, flattened for the AO kernel.
struct ShellDev {
    int atom_coord;  // offset of the centre (x,y,z) in env
    int l;
    int nprim;
    int nctr;
    int ptr_exp;
    int ptr_coef;
    int ao_off;  // first AO index of this shell
    int rad_off; // first radial-function index (one per contraction) of this shell
};
// Per-AO lookup for the tiled AO kernel: which radial function, which atom, which harmonic.
struct AoMeta {
    int rad;     // radial-function index
    short atom;  // atom index
    short lm;    // l*l + m, m = 0..2l
};

}  // namespace qexxc

struct qexxc_ctx {
    int device = 0;
    int B = 1, C = 1, Gmax = 0, GpadMax = 0, N = 0, Npad = 0, Nc = 0;  // Npad: storage pitch, Nc: compute extent
    int G = 0, Gpad = 0;  // current grid
    qexxc_net_desc net{};
    long n_theta = 0;
    int num_sms = 148;
    long launches = 0;
    size_t bytes = 0;

    // stage 1
    double* coords = nullptr;   // [B][GpadMax][3]
    double* weights = nullptr;  // [B][GpadMax] (zero beyond G)
    double* ao = nullptr;       // [B or 1][C][GpadMax][Npad], zero padded
    int ao_ncomp = 0;           // components currently valid in `ao`
    bool have_grid = false, have_basis = false;
    bool ao_shared = false;     // QEXXC_FLAG_SHARED_AO: one AO tensor / grid / geometry for the whole batch (nset density matrices)
    qexxc::ShellDev* shells = nullptr;
    int nshell = 0;
    double* env = nullptr;  // [B][nenv]
    int nenv = 0;
    qexxc::AoMeta* ao_meta = nullptr;  // [Npad]
    int* atom_coord = nullptr;         // [natm] offset of (x,y,z) in env
    int* shell_atom = nullptr;         // [nshell]
    int natm = 0, nrad = 0, lmax = 0;

    // stage 2/4 workspaces (all [B][...][GpadMax] rows are zero beyond G)
    double* S = nullptr;        // [B][Npad][Npad] padded symmetric operand (dm or V_bar + V_bar^T); also MO factor L
    double* mosgn = nullptr;    // [B][Npad] occupation signs of the MO form
    double* rho = nullptr;      // [B][C][GpadMax]
    double* exc = nullptr;      // [B][GpadMax]
    double* vrho = nullptr;     // [B][GpadMax]
    double* vgamma = nullptr;   // [B][GpadMax]
    double* wv = nullptr;       // [B][C][GpadMax]  per-point scale factors of the V_xc contraction
    double* wvb = nullptr;      // [B][C][GpadMax]  cotangent of wv
    double* rbar = nullptr;     // [B][C][GpadMax]
    double* excb = nullptr;     // [B][GpadMax]
    double* vrhob = nullptr;    // [B][GpadMax]
    double* vgammab = nullptr;  // [B][GpadMax]
    double* aow = nullptr;      // [B][GpadMax][Npad] (C == 4 only)
    double* rq_part = nullptr;  // rowquad split-tail partials [NT][4][num_sms][128]
    void* i8ws = nullptr;       // INT8 (Ozaki) contraction workspace: digit planes of ao_0 / S / the weighted operand (contract_i8.cu)
    bool i8_valid = false;      // the geometry-dependent digit planes match the current ao
    double* rq_pair = nullptr;  // rowquad pair-mode partials [2][4][GpadMax] (single-molecule contexts with wide N)
    double* part = nullptr;     // wsyrk partial tiles, one compact [BN][BN] slot per (batch, tile, grid chunk)
    size_t part_doubles = 0;
    void* ws_items[2] = {nullptr, nullptr};  // wsyrk static schedules (general, symmetric)
    int* ws_start[2] = {nullptr, nullptr};
    long ws_key[2] = {-1, -1};
    size_t ws_items_bytes = 0, ws_start_cap = 0;
    double* red = nullptr;      // per-CTA partial sums (excsum, nelec, theta_bar ...)
    size_t red_doubles = 0;
    double* wide_ws = nullptr;  // LocalMLP wider than 64 or deeper than 3: layer-by-layer workspace (xc_mlp_wide.cu)
    void* tape = nullptr;       // MLP reverse-mode tape (per-CTA slots)
    size_t tape_bytes = 0;
    unsigned char* qperm = nullptr;  // QNN ring permutation tables (2 x 256 bytes)
    std::vector<void*> allocs;
    // optional per-kernel-class CUDA-event timing (bench.py roofline): class -> (start, stop) pairs
    bool prof = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev[QEXXC_PROF_NCLASS];
};

namespace qexxc {
// RAII: records an event pair around the launches made in its scope when profiling is on
struct ProfScope {
    qexxc_ctx* c;
    int cls;
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    ProfScope(qexxc_ctx* ctx, int k, cudaStream_t s) : c(ctx), cls(k), st(s) {
        if (c->prof) {
            cudaEvent_t e0;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, st);
            c->prof_ev[cls].push_back({e0, e1});
        }
    }
    ~ProfScope() {
        if (e1) cudaEventRecord(e1, st);
    }
};
}  // namespace qexxc

namespace qexxc {

// ---- launchers implemented in the .cu files --------------------------------------------------
// contract.cu
void wsyrk_workspace(int num_sms, int Nc, int GpadMax, int B, bool general, size_t* part_doubles,
                     size_t* item_bytes, size_t* start_ints);
double rowquad_executed_flops(const qexxc_ctx* c, int tri);
double wsyrk_executed_flops(const qexxc_ctx* c, bool sym);
// mode 0: (a+a^T)/2, 1: a, 2: a+a^T; tri: keep the upper triangle only (diagonal halved)
int launch_pad_sym(qexxc_ctx* c, const double* src, int mode, int tri, cudaStream_t st);
// q[b][k][g] = fac[k] * sum_ij ao[b][k][g][i] S[b][i][j] ao[b][0][g][j], k < ncomp
// tri: S was built with tri=1 (only valid for ncomp == 1, where the form is symmetric in i,j)
int launch_rowquad(qexxc_ctx* c, int ncomp, int tri, const double* fac4, double* q, long q_bstride,
                   long q_cstride, cudaStream_t st);
// H[b] = ao0^T diag(s) ao0 (Bsrc == nullptr, symmetric) or ao0^T Bsrc (general);
// out[b][i][j] = scale * (H[i][j] + (tadd ? H[j][i] : 0)), i,j < N, row stride N
int launch_wsyrk(qexxc_ctx* c, const double* s, long s_bstride, const double* Bsrc, double scale, int tadd,
                 double* out, long out_bstride, cudaStream_t st);
// MO form of rho (pyscf eval_rho2, reached by numint_legacy.py:527-545)
int launch_pack_mo(qexxc_ctx* c, const double* C, const double* occ, int nmo, double* L, double* sgn, int ldL,
                   cudaStream_t st);
int launch_rowquad_mo(qexxc_ctx* c, const double* L, int ldL, int nk, const double* sgn, double* q, long q_bstride,
                      cudaStream_t st);
// aow[b][g][n] = sum_c f[c] wv[b][c][g] ao[b][c][g][n]
int launch_build_aow(qexxc_ctx* c, const double* wv, long wv_bstride, long wv_cstride, const double* fac4,
                     cudaStream_t st);
// contract_i8.cu: exact INT8 (Ozaki) form of rowquad / wsyrk on tcgen05 (QEXXC_I8=0/1; default: nao >= 256, one AO tensor per context)
bool i8_enabled(const qexxc_ctx* c);
void i8_release(qexxc_ctx* c);
int launch_rowquad_i8(qexxc_ctx* c, int ncomp, int tri, const double* fac4, double* q, long q_bstride, long q_cstride, cudaStream_t st);
int launch_wsyrk_i8(qexxc_ctx* c, const double* s, long s_bstride, const double* Bsrc, double scale, int tadd, double* out,
                    long out_bstride, cudaStream_t st);
int launch_rowquad_mo_i8(qexxc_ctx* c, const double* L, int ldL, int nk, const double* sgn, double* q, long q_bstride, cudaStream_t st);
double i8_executed_ops(const qexxc_ctx* c, int which, bool sym);
int i8_prepare_geometry(qexxc_ctx* c, cudaStream_t st);
int launch_build_aow_i8(qexxc_ctx* c, const double* wv, long wv_cstride, const double* fac4, cudaStream_t st);
int i8_reserve(qexxc_ctx* c);  // allocate the digit-plane workspace now (qexxc_create), not on first use
int i8_peak_probe(int device, double* ops_per_second);
// ao.cu
int launch_eval_ao(qexxc_ctx* c, int deriv, cudaStream_t st);
int launch_pack_ao(qexxc_ctx* c, const double* src, int ncomp, int G, cudaStream_t st);
int launch_unpack_ao(qexxc_ctx* c, double* dst, int ncomp, cudaStream_t st);
int launch_set_grid(qexxc_ctx* c, const double* coords, const double* weights, int G, cudaStream_t st);
// pointwise.cu
int stage4_nblocks(const qexxc_ctx* c);
int launch_stage4_pointwise(qexxc_ctx* c, int xctype, const double* rho, const double* exc, const double* vrho,
                            const double* vgamma, double* wv, double* sums, long sums_bstride, cudaStream_t st);
int launch_stage4_pointwise_vjp(qexxc_ctx* c, int xctype, const double* rho, const double* exc,
                                const double* vrho, const double* vgamma, const double* e_bar, const double* wvb,
                                double* rho_bar, double* exc_bar, double* vrho_bar, double* vgamma_bar,
                                cudaStream_t st);
// dst[b][c][i] = i < n ? src[b][c][i] : 0 for i < n_pad
int launch_copy_rows(qexxc_ctx* c, double* dst, long dst_bstride, long dst_cstride, const double* src,
                     long src_bstride, long src_cstride, int nb, int nc, long n, long n_pad, cudaStream_t st);
int launch_transpose_in(qexxc_ctx* c, double* feat, long ld, const double* x, long npts, int F, long n_pad,
                        cudaStream_t st);
int launch_transpose_out(qexxc_ctx* c, double* x, const double* feat, long ld, long npts, int F, cudaStream_t st);
// xc_mlp.cu
int mlp_local_grid(const qexxc_ctx* c);
size_t mlp_local_tape_bytes(const qexxc_ctx* c);
int launch_mlp_local_fwd(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                         const double* theta, double* exc, double* vrho, double* vgamma, long out_bstride,
                         int nbatch, long npts_per_batch, cudaStream_t st);
int launch_mlp_local_vjp(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                         const double* theta, const double* exc_bar, const double* vrho_bar,
                         const double* vgamma_bar, long in_bstride, double* rho_bar, int accumulate,
                         double* theta_bar, int accumulate_theta, int nbatch, long npts_per_batch,
                         cudaStream_t st);
// xc_mlp_wide.cu: any width / depth, layer by layer (float64)
bool mlp_is_wide(const qexxc_net_desc& net);
size_t mlp_wide_ws_doubles(const qexxc_ctx* c);
int launch_mlp_wide_fwd(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                        const double* theta, double* exc, double* vrho, double* vgamma, long out_bstride, int nbatch,
                        long npts_per_batch, cudaStream_t st);
int launch_mlp_wide_vjp(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                        const double* theta, const double* exc_bar, const double* vrho_bar, const double* vgamma_bar,
                        long in_bstride, double* rho_bar, int accumulate, double* theta_bar, int accumulate_theta,
                        int nbatch, long npts_per_batch, cudaStream_t st);
// xc_mlp_tc.cu: FP32 network path on tcgen05 / TMEM (width <= 64, <= 3 hidden layers)
bool mlp_tc_enabled(const qexxc_ctx* c);
int launch_mlp_tc_fwd(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                      const double* theta, double* exc, double* vrho, double* vgamma, long out_bstride, int nbatch,
                      long npts_per_batch, cudaStream_t st);
size_t mlp_tc_tape_bytes(const qexxc_ctx* c);
int launch_mlp_tc_vjp(qexxc_ctx* c, int xctype, const double* rho, long rho_bstride, long rho_cstride,
                      const double* theta, const double* exc_bar, const double* vrho_bar, const double* vgamma_bar,
                      long in_bstride, double* rho_bar, int accumulate, int nbatch, long npts_per_batch, int* grid_out,
                      cudaStream_t st);
// xc_global.cu
size_t global_mlp_smem(int L, int H);
int launch_global_mlp(qexxc_ctx* c, bool vjp, const double* rho, long ld, int G, const double* theta, double* exc,
                      long exc_stride, double* vrho, const double* exc_bar, const double* vrho_bar,
                      double* rho_bar, double* theta_bar, int accumulate_theta, int nbatch, cudaStream_t st);
// xc_qnn.cu
int qnn_grid(const qexxc_ctx* c, long npts);
size_t qnn_red_doubles(const qexxc_ctx* c, long npts_max);
int qnn_upload_perm(qexxc_ctx* c, unsigned char* dev_tables);
int launch_qnn(qexxc_ctx* c, bool vjp, const unsigned char* tables, const double* x, long npts, const double* theta,
               double* exc, double* vrho, const double* exc_bar, const double* vrho_bar, double* x_bar,
               int accumulate, double* theta_bar, int accumulate_theta, cudaStream_t st);

}  // namespace qexxc

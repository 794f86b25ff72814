/*
 * qexxc.h -- C ABI of libqexxc.so: the B200 (sm_100a) implementation of QEX's 3D Kohn-Sham
 * exchange-correlation (XC) grid-integration hot path, forward and reverse mode.
 *
 * The reference (pasqal-io/qex, Python package `qedft`) has no native/FFI interface: its boundary
 * for this path is four Python call signatures.  Each entry point below
 * cites the reference function it stands in for (paths relative to the reference tree).  The
 * Python host side (qex_b200/) binds these with ctypes; qex_b200/jax_ffi_shim.py shows the
 * jax.ffi / jax.custom_vjp registration a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative QEXXC_ERR_* code; qexxc_last_error()
 *     returns a thread-local human-readable message.  Nothing throws, nothing aborts.
 *   - `*_dev` pointers are device pointers on the context's device; all arrays are float64,
 *     C-order, unpadded (the library keeps its own padded copies).  `stream` is a cudaStream_t
 *     passed as void*; all work is enqueued on it and no call synchronises the device unless
 *     stated.  No hot call allocates: all workspaces are sized in qexxc_create().
 *   - B = batch of independent molecules with identical (nao, ngrids_max, basis structure);
 *     arrays carry a leading [B] dimension.  theta (network parameters) is shared by the batch
 *     and theta_bar is summed over it.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef QEXXC_H
#define QEXXC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QEXXC_VERSION 210

#define QEXXC_OK 0
#define QEXXC_ERR_CUDA (-1)        /* a CUDA runtime call or kernel launch failed */
#define QEXXC_ERR_ARG (-2)         /* bad argument (null pointer, size out of range) */
#define QEXXC_ERR_STATE (-3)       /* call order violated (e.g. no AO set before nr_rks) */
#define QEXXC_ERR_UNSUPPORTED (-4) /* mirrors the reference's NotImplementedError / ValueError */
#define QEXXC_ERR_NODEVICE (-5)    /* no CUDA device: there is no CPU fallback */

/* xctype: the branch of nr_rks (qedft/train/td/numint_legacy.py:134-140) */
#define QEXXC_XC_NN 0        /* "NN": local functional, LDA features (numint_legacy.py:290-310) */
#define QEXXC_XC_NN_GLOBAL 1 /* "NN-AmplitudeEncoding": global functional (numint_legacy.py:311-334) */
#define QEXXC_XC_GGA 2       /* "GGA" assembly (numint_legacy.py:175-198) with features (rho, sigma) */

/* network kinds (qedft/models/networks.py) */
#define QEXXC_NET_NONE 0       /* exc/vrho supplied by the caller (host eval_xc callback) */
#define QEXXC_NET_LOCAL_MLP 1  /* LocalMLP networks.py:83-109, classical_models.py:123-174 */
#define QEXXC_NET_GLOBAL_MLP 2 /* GlobalMLP networks.py:112-138, classical_models.py:177-224 */
#define QEXXC_NET_LOCAL_QNN 3  /* LocalQNN networks.py:180-228, quantum_models.py:115-157 */

/* activations (classical_models.py:39-49 ACTIVATION_MAP; swish for the flax MLP trainer :96-107) */
#define QEXXC_ACT_TANH 0
#define QEXXC_ACT_RELU 1
#define QEXXC_ACT_SOFTPLUS 2
#define QEXXC_ACT_SIGMOID 3
#define QEXXC_ACT_ELU 4
#define QEXXC_ACT_LEAKY_RELU 5
#define QEXXC_ACT_SELU 6
#define QEXXC_ACT_GELU 7
#define QEXXC_ACT_SWISH 8

/* switching function of the Becke partition (pyscf.dft.gen_grid.original_becke / .stratmann) */
#define QEXXC_BECKE_ORIGINAL 0
#define QEXXC_BECKE_STRATMANN 1

#define QEXXC_PREC_F64 0
#define QEXXC_PREC_F32 1

#define QEXXC_MAX_LAYERS 8

typedef struct qexxc_net_desc {
    int kind;          /* QEXXC_NET_* */
    int n_features;    /* MLP: inputs per point (1 = rho, 2 = rho,sigma); global MLP: ignored (= ngrids) */
    int n_hidden;      /* MLP: number of hidden Dense+activation pairs; QNN: ansatz layers */
    int width;         /* MLP: neurons per hidden layer; QNN: number of qubits */
    int activation;    /* QEXXC_ACT_* */
    int out_transform; /* 0 none; 1 = -out_scale*swish(.) (flax MLP, trainer_legacy_no_jit.py:107) */
    int precision;     /* QEXXC_PREC_F64 (reference runs jax_enable_x64) or QEXXC_PREC_F32 */
    int reserved;
    double in_scale;   /* 1/density_normalization_factor (classical_models.py:169); QNN: 1.0 */
    double out_scale;  /* 1e-2 for the flax MLP */
} qexxc_net_desc;

typedef struct qexxc_ctx qexxc_ctx;

/* ---- lifetime ---------------------------------------------------------------------------- */
int qexxc_version(void);
const char* qexxc_last_error(void);
/* number of network parameters (flat theta length) for a descriptor; global MLP needs ngrids */
long qexxc_n_params(const qexxc_net_desc* net, int ngrids);
/* Allocates every workspace for problems up to (nbatch, ncomp, ngrids_max, nao).  ncomp = 1
 * (AO values, "LDA"/"NN") or 4 (values + gradient, "GGA").  Replaces nothing in the reference
 * (JAX allocates implicitly); it is the price of "no allocation on the hot call". */
int qexxc_create(qexxc_ctx** ctx, int device, int nbatch, int ncomp, int ngrids_max, int nao,
                 const qexxc_net_desc* net);
/* As qexxc_create, with flags.  QEXXC_FLAG_SHARED_AO: the nbatch elements are `nset` density matrices of ONE
 * molecule / grid (the `for idm in range(nset)` loop of nr_rks, numint_legacy.py:141-156,292-310, as one launch
 * per stage): a single AO tensor [C][G][N], grid and geometry serve the whole batch -- set_grid takes coords
 * [G][3], weights [G]; set_basis env [nenv]; set_ao / get_ao [C][G][N]; dm, out, resid keep their [B] dimension. */
#define QEXXC_FLAG_SHARED_AO 1u
int qexxc_create_ex(qexxc_ctx** ctx, int device, int nbatch, int ncomp, int ngrids_max, int nao,
                    const qexxc_net_desc* net, unsigned flags);
int qexxc_destroy(qexxc_ctx* ctx);
/* bytes of device memory held by the context */
size_t qexxc_workspace_bytes(const qexxc_ctx* ctx);

/* ---- stage 1: grid + AO values ------------------------------------------------------------
 * qexxc_set_grid: grids.coords [B][G][3], grids.weights [B][G] (device), as yielded by pyscf
 *   block_loop (numint_legacy.py:292).  G <= ngrids_max.
 * qexxc_set_basis: libcint tables mol._atm [natm][6], mol._bas [nbas][8] (HOST int32) and
 *   mol._env [B][nenv] (HOST float64; per-batch geometry) -- what pyscf's eval_gto reads
 *   (qedft/train/td/eval_gto.py:48-70).  Spherical shells, l <= 3.
 * qexxc_eval_ao: K1 -- replaces numint.eval_ao / block_loop's eval_ao (numint_legacy.py:292,313;
 *   trainer_legacy_no_jit.py:273): fills the context's AO tensor from basis + grid.
 * qexxc_set_ao: alternative to K1: upload pre-evaluated AO [B][C][G][N] (e.g. from pyscf).
 * qexxc_get_ao: copy the AO tensor out as [B][C][G][N] (eval_ao's return value). */
int qexxc_set_grid(qexxc_ctx* ctx, const double* coords_dev, const double* weights_dev, int ngrids,
                   void* stream);
int qexxc_set_basis(qexxc_ctx* ctx, const int* atm, int natm, const int* bas, int nbas,
                    const double* env, int nenv);
int qexxc_eval_ao(qexxc_ctx* ctx, int deriv, void* stream);
int qexxc_set_ao(qexxc_ctx* ctx, const double* ao_dev, int ncomp, int ngrids, void* stream);
int qexxc_get_ao(qexxc_ctx* ctx, double* ao_dev, int ncomp, void* stream);

/* ---- stage 2: density on the grid ----------------------------------------------------------
 * eval_rho(mol, ao, dm, non0tab, xctype, hermi) numint_legacy.py:351-397.  dm [B][N][N];
 * rho out [B][C][G] with C = 1 ("LDA") or 4 ("GGA": rho, 2<c0,d_x ao>, ...).  hermi=0
 * symmetrises dm first (:362-365).  The VJP maps rho_bar [B][C][G] -> dm_bar [B][N][N]. */
int qexxc_eval_rho(qexxc_ctx* ctx, const double* dm_dev, int ncomp, int hermi, double* rho_dev,
                   void* stream);
int qexxc_eval_rho_vjp(qexxc_ctx* ctx, const double* rho_bar_dev, int ncomp, int hermi,
                       double* dm_bar_dev, void* stream);

/* ---- stage 3: the learned XC functional -----------------------------------------------------
 * qexxc_xc_fwd: eval_xc(..., rho, ..., params) trainer_legacy_no_jit.py:76-93 ->
 *   exc_and_vrho_local :56-63 (xctype NN: exc [B][G], vrho [B][G]),
 *   exc_and_vrho_global :46-53 (xctype NN_GLOBAL: exc [B] scalars, vrho [B][G]),
 *   GGA extension (xctype GGA: rho [B][4][G] in; exc [B][G], vrho [B][G], vgamma [B][G]).
 * qexxc_xc_vjp: reverse rule of the above w.r.t. (rho, theta): cotangents exc_bar, vrho_bar
 *   (, vgamma_bar) -> rho_bar [B][C][G], theta_bar [n_params] (summed over batch and grid).
 * qexxc_apply_fn_fwd/_vjp: network apply_fn(params, inputs) (networks.py:43-75;
 *   classical_models.py:168-172; quantum_models.py:759-774): x [npts][n_features] -> y [npts]
 *   (global MLP: x [ngrids] -> y [1]); VJP: y_bar -> x_bar, theta_bar.
 * Every entry that takes `theta_dev` also takes its length `n_theta` and fails with QEXXC_ERR_ARG unless it
 * equals qexxc_n_params(net, G) for the grid in use (the reference raises a shape error when the parameter
 * tree does not fit the network; a bare pointer could not be checked). */
int qexxc_xc_fwd(qexxc_ctx* ctx, int xctype, const double* rho_dev, const double* theta_dev, long n_theta,
                 double* exc_dev, double* vrho_dev, double* vgamma_dev, void* stream);
int qexxc_xc_vjp(qexxc_ctx* ctx, int xctype, const double* rho_dev, const double* theta_dev, long n_theta,
                 const double* exc_bar_dev, const double* vrho_bar_dev,
                 const double* vgamma_bar_dev, double* rho_bar_dev, double* theta_bar_dev,
                 void* stream);
int qexxc_apply_fn_fwd(qexxc_ctx* ctx, const double* x_dev, long npts, const double* theta_dev, long n_theta,
                       double* y_dev, void* stream);
int qexxc_apply_fn_vjp(qexxc_ctx* ctx, const double* x_dev, long npts, const double* theta_dev, long n_theta,
                       const double* y_bar_dev, double* x_bar_dev, double* theta_bar_dev,
                       void* stream);

/* ---- stage 4: E_xc, nelec, V_xc from caller-supplied (exc, vxc) ------------------------------
 * The body of the nr_rks block loop after eval_xc (numint_legacy.py:304-309 / :328-333 /
 * :189-197) plus vmat + vmat.T (:336-337), for functionals evaluated by a host callback.
 * out [B][N*N + 2] = vmat (N*N, C-order) | excsum | nelec.  exc: [B][G] (NN, GGA) or [B] (global).
 * The VJP returns the cotangents of (rho, exc, vrho, vgamma) given (e_bar [B], v_bar [B][N][N]). */
int qexxc_vxc_assemble(qexxc_ctx* ctx, int xctype, const double* rho_dev, const double* exc_dev,
                       const double* vrho_dev, const double* vgamma_dev, double* out_dev,
                       void* stream);
int qexxc_vxc_assemble_vjp(qexxc_ctx* ctx, int xctype, const double* rho_dev,
                           const double* exc_dev, const double* vrho_dev,
                           const double* vgamma_dev, const double* e_bar_dev,
                           const double* v_bar_dev, double* rho_bar_dev, double* exc_bar_dev,
                           double* vrho_bar_dev, double* vgamma_bar_dev, void* stream);

/* ---- the fused hot path --------------------------------------------------------------------
 * qexxc_nr_rks_fwd: NumInt.nr_rks(mol, grids, xc_code, dms, hermi, params=params)
 *   numint_legacy.py:122-348 for one dm per batch element (nset == 1), with the context's
 *   network as ni.eval_xc.  out [B][N*N + 2] = vmat | excsum | nelec (one packed buffer so a
 *   multi-GPU caller all-reduces it in a single collective).  resid (nullable) receives the
 *   residuals the reverse pass needs: qexxc_resid_doubles(ctx) float64.
 * qexxc_nr_rks_vjp: reverse rule JAX derives for the above under jax.value_and_grad
 *   (trainer_legacy_no_jit.py:284): (e_bar [B], v_bar [B][N][N]) -> bar [B*N*N + n_params] =
 *   dm_bar (B*N*N) | theta_bar (n_params, summed over batch); nelec is stop-gradient (:305). */
size_t qexxc_resid_doubles(const qexxc_ctx* ctx);
/* MO form of the density: NumInt._gen_rho_evaluator's `dms.mo_coeff` branch (numint_legacy.py:527-545 ->
 * pyscf eval_rho2): rho = sum_k occ_k (ao C_k)^2 over orbitals with |occ_k| > 1e-12 (negative occupations
 * subtract).  mo_coeff [B][N][nmo], mo_occ [B][nmo], nmo <= N; 2*G*N*nmo instead of 2*G*N^2 FLOPs.
 * qexxc_nr_rks_fwd_mo = qexxc_nr_rks_fwd with that stage 2 (LDA-type branches); outputs, residuals and
 * the reverse pass (qexxc_nr_rks_vjp, cotangent w.r.t. dm = C occ C^T) are unchanged. */
int qexxc_eval_rho_mo(qexxc_ctx* ctx, const double* mo_coeff_dev, const double* mo_occ_dev, int nmo,
                      double* rho_dev, void* stream);
int qexxc_nr_rks_fwd_mo(qexxc_ctx* ctx, int xctype, const double* mo_coeff_dev, const double* mo_occ_dev, int nmo,
                        const double* theta_dev, long n_theta, double* out_dev, double* resid_dev, void* stream);
int qexxc_nr_rks_fwd(qexxc_ctx* ctx, int xctype, int hermi, const double* dm_dev,
                     const double* theta_dev, long n_theta, double* out_dev, double* resid_dev, void* stream);
int qexxc_nr_rks_vjp(qexxc_ctx* ctx, int xctype, int hermi, const double* theta_dev, long n_theta,
                     const double* resid_dev, const double* e_bar_dev, const double* v_bar_dev,
                     double* bar_dev, void* stream);

/* ---- multi-GPU: the one exchange step of the path -------------------------------
 * The grid shards by point ranges, one process per GPU; dm / theta / basis are replicated.  The reference has no
 * multi-device code; these calls are what a grid-sharded caller adds around nr_rks:
 *   qexxc_allreduce: in-place sum over ranks of a packed output buffer -- `out` [B][N*N+2] of qexxc_nr_rks_fwd
 *     (vmat | excsum | nelec) or `bar` of qexxc_nr_rks_vjp (dm_bar | theta_bar) -- one collective per direction;
 *   qexxc_bcast: replicate (dm | theta | cotangents) from the rank that received them from the host.
 * A qexxc_comm wraps an ncclComm_t: either created here (rank 0 calls qexxc_comm_unique_id, ships the 128 bytes to
 * the other ranks by any side channel, every rank calls qexxc_comm_create) or borrowed from the host framework
 * (qexxc_comm_wrap; not destroyed by qexxc_comm_destroy).  NCCL is bound with dlopen at first use; without it the
 * calls fail with QEXXC_ERR_STATE.  Collectives are enqueued on `stream` and never synchronise the host. */
typedef struct qexxc_comm qexxc_comm;
int qexxc_comm_nccl_version(int* version);
int qexxc_comm_unique_id(unsigned char id[128]);
int qexxc_comm_create(qexxc_comm** comm, int device, int world, int rank, const unsigned char id[128]);
int qexxc_comm_wrap(qexxc_comm** comm, void* nccl_comm, int device, int world, int rank);
int qexxc_comm_destroy(qexxc_comm* comm);
int qexxc_comm_rank(const qexxc_comm* comm);
int qexxc_comm_world(const qexxc_comm* comm);
long qexxc_comm_calls(const qexxc_comm* comm); /* collectives enqueued since creation */
int qexxc_allreduce(qexxc_comm* comm, double* buf_dev, long count, void* stream);
int qexxc_bcast(qexxc_comm* comm, double* buf_dev, long count, int root, void* stream);

/* ---- introspection for the benchmark -------------------------------------------------------
 * Number of kernels this library launched on the context since creation (gpu_launches). */
long qexxc_launch_count(const qexxc_ctx* ctx);
/* Per-kernel-class device timing with CUDA events recorded on the caller's stream around the
 * launches of one class (bench.py's roofline line).  qexxc_profile_read synchronises on the
 * recorded events, returns the summed milliseconds and launch count and clears the class. */
#define QEXXC_PROF_ROWQUAD 0
#define QEXXC_PROF_WSYRK 1
#define QEXXC_PROF_XC_FWD 2
#define QEXXC_PROF_XC_VJP 3
#define QEXXC_PROF_EVAL_AO 4
#define QEXXC_PROF_STAGE4 5 /* the streaming stage-4 kernels: wv / E_xc / nelec forward, and their adjoint */
#define QEXXC_PROF_SLICE 6  /* INT8 path only: the digit-slicing passes (per geometry and per call) */
#define QEXXC_PROF_NCLASS 7
int qexxc_profile_enable(qexxc_ctx* ctx, int on);
int qexxc_profile_read(qexxc_ctx* ctx, int cls, double* ms_total, long* count);
/* DMMA FLOPs the contraction kernels actually execute for the current problem shape (zero-padded and
 * structurally-zero blocks excluded): which = 0 rowquad, 1 wsyrk; symmetric = triangular variant. */
int qexxc_contraction_flops(qexxc_ctx* ctx, int which, int symmetric, double* executed_flops);
/* Which tensor pipe the two contractions (numint_legacy.py:351-410 eval_rho, :432-456 V_xc) run on for this context:
 * 0 = FP64 DMMA (contract.cu), 1 = exact INT8 digit split on tcgen05 (contract_i8.cu; env QEXXC_I8=0/1 overrides the
 * default "nao >= 256, single molecule").  Results agree to ~1e-12 of the largest element either way. */
int qexxc_contraction_mode(const qexxc_ctx* ctx);
/* Builds now, on `stream`, whatever the contractions derive from the AO tensor alone (mode 1: the digit planes of ao_0), so
 * that it overlaps the upload of the density matrix instead of running inside the first contraction.  Optional: the
 * contractions build it on first use.  No-op in mode 0.  Needs qexxc_eval_ao / qexxc_set_ao first. */
int qexxc_prepare_contractions(qexxc_ctx* ctx, void* stream);
/* INT8 multiply-add operations (2 per MAC) one launch of rowquad (0) / wsyrk (1) executes in mode 1. */
int qexxc_contraction_i8_ops(qexxc_ctx* ctx, int which, int symmetric, double* executed_ops);
/* Measures the INT8 tensor-core rate of `device` with the library's own issue loop (kind::i8, M = 128, N = 256, K = 32,
 * operands resident in shared memory, one CTA per SM): the roofline denominator of mode 1, in ops/s (2 per MAC). */
int qexxc_i8_peak(int device, double* ops_per_second);
/* Runs only the dominant contraction kernel once on the current AO/S buffers (roofline timing):
 * which = 0 rowquad (rho-type), 1 wsyrk (vmat-type). */
int qexxc_debug_run_contraction(qexxc_ctx* ctx, int which, void* stream);

/* ---- "next" row N2: incore Coulomb / exchange build on the dense s1 ERI tensor --------
 * Replaces the two einsums of `_dot_eri_dm_s1` (qedft/train/td/hf_legacy.py:275-286; called from
 * `SCF.get_jk` :452-470 -> `get_veff` rks_legacy.py:91-117):
 *     vj[x,k,l] = sum_ij eri[i,j,k,l] dm[x,j,i]      vk[x,i,l] = sum_jk eri[i,j,k,l] dm[x,j,k]
 * All pointers are DEVICE pointers (the 8*nao^4-byte tensor stays resident in HBM across SCF cycles);
 * eri is [nao]^4 row-major, dm / vj / vk are [nset][nao][nao].  One pass over the tensor produces J and K.
 * `work` needs qexxc_jk_workspace_doubles doubles.  These calls are context-free (no grid involved). */
int qexxc_jk_workspace_doubles(int device, int nao, long* out);
int qexxc_dot_eri_dm(int device, const double* eri_dev, const double* dm_dev, int nset, int nao, int with_j,
                     int with_k, double* vj_dev, double* vk_dev, double* work_dev, long work_doubles, void* stream);
/* Reverse mode of the call above w.r.t. dm (what jax.vjp of the einsums gives; the ERI tensor is not a
 * parameter): dm_bar[x,j,i] = sum_kl eri[i,j,k,l] vj_bar[x,k,l] + sum_{i',l} eri[i',j,i,l] vk_bar[x,i',l].
 * Either cotangent may be NULL (= zero). No permutational symmetry of eri is assumed. */
int qexxc_dot_eri_dm_vjp(int device, const double* eri_dev, const double* vj_bar_dev, const double* vk_bar_dev,
                         int nset, int nao, double* dm_bar_dev, double* work_dev, long work_doubles, void* stream);
/* Batched over `nmol` molecules of equal nao, each with its own tensor and ONE density matrix (the c4 pattern:
 * a dissociation curve of small molecules in one launch): eri [nmol][nao^4], dm / vj / vk / cotangents
 * [nmol][nao][nao]; `work` needs nmol * qexxc_jk_workspace_doubles doubles. */
int qexxc_dot_eri_dm_batched(int device, const double* eri_dev, const double* dm_dev, int nmol, int nao, int with_j,
                             int with_k, double* vj_dev, double* vk_dev, double* work_dev, long work_doubles,
                             void* stream);
int qexxc_dot_eri_dm_vjp_batched(int device, const double* eri_dev, const double* vj_bar_dev, const double* vk_bar_dev,
                                 int nmol, int nao, double* dm_bar_dev, double* work_dev, long work_doubles,
                                 void* stream);
/* ---- "next" row N3 for SMALL matrices: batched generalised symmetric eigensolver A v = w B v, n <= 16 ---------
 * `generalized_eigh` of qedft/train/td/generalized_eigensolver.py:264-330 (symmetrise, SPD shift of B by
 * eps - lambda_min(B) when positive, Cholesky, two triangular solves, eigh, back-transform; eigenvalues
 * ascending, V^T B V = I) for nbatch independent problems, one thread each, no host synchronisation.
 * a, b, v: [nbatch][n][n]; w: [nbatch][n].  n > 16 returns QEXXC_ERR_UNSUPPORTED (use cuSOLVER). */
int qexxc_generalized_eigh_batched(int device, const double* a_dev, const double* b_dev, int nbatch, int n,
                                   double eps, double* w_dev, double* v_dev, void* stream);
/* kernels launched by the three J/K calls since the library was loaded */
long qexxc_jk_launch_count(void);

/* ---- "next" row N4: grid generation --------------------------------------------------------------------------
 * qexxc_becke_partition: the O(G natm^2) half of `pyscf.dft.gen_grid.Grids.build()` as the reference calls it
 * (qedft/train/td/trainer_legacy_no_jit.py:248-251, :316-317; data_io/td/dataset_generation.py:139-142 --
 * `get_partition` / `gen_grid_partition` in pyscf): Becke fuzzy-cell weights of every grid point,
 *     weights[g] = vol[g] * P_owner[g](r_g) / sum_i P_i(r_g).
 * coords [G][3] (Bohr), owner [G] (index of the atom the point was generated around), vol [G] (radial x angular
 * quadrature weight), atom_coords [natm][3], adjust [natm][natm] = Treutler's a_ij (NULL: no atomic-size
 * adjustment), scheme = QEXXC_BECKE_ORIGINAL | QEXXC_BECKE_STRATMANN, inv_dist_work: natm*natm doubles of scratch.
 * All pointers are device pointers.  natm is limited by shared memory (one distance per atom and thread:
 * ~220 atoms); beyond that QEXXC_ERR_UNSUPPORTED. */
int qexxc_becke_partition(int device, const double* coords_dev, long ngrids, const int* owner_dev,
                          const double* vol_dev, const double* atom_coords_dev, const double* adjust_dev, int natm,
                          int scheme, double* inv_dist_work_dev, double* weights_dev, void* stream);
/* kernels launched by qexxc_becke_partition since the library was loaded */
long qexxc_grid_launch_count(void);

/* ---- analytic LDA exchange (libxc LDA_X, unpolarised): the functional of pyscf's `xc = "lda"` -------------------
 * What `ni.eval_xc("lda", rho, ...)` returns inside the reference's nr_rks (numint_legacy.py LDA branch; run by
 * dataset_generation.py:385-389): exc[g] = -3/4 (3/pi)^(1/3) rho^(1/3) (energy per particle),
 * vrho[g] = d(rho exc)/d rho = 4/3 exc.  rho, exc, vrho: npts doubles on the device (any batch layout, flat).
 * Feed the result to qexxc_vxc_assemble with xctype QEXXC_XC_NN (the same assembly as the LDA branch). */
int qexxc_lda_exchange(int device, const double* rho_dev, long npts, double* exc_dev, double* vrho_dev, void* stream);
long qexxc_lda_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* QEXXC_H */

// extern "C" entry points of libqexxc.so (see include/qexxc.h for the contract and the reference
// functions each one stands in for).  Host-side only: argument checks, workspace bookkeeping and
// the order in which the stage kernels are enqueued on the caller's stream.
#include <cstdarg>

#include "common.cuh"

namespace qexxc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static const double kFacGGA[4] = {1.0, 2.0, 2.0, 2.0};  // _rks_gga_assemble_rho numint_legacy.py:401-410
static const double kFacOne[4] = {1.0, 1.0, 1.0, 1.0};

template <typename T>
static int dev_alloc(qexxc_ctx* c, T** p, size_t count, bool zero = true) {
    if (count == 0) count = 1;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) {
        set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
        return QEXXC_ERR_CUDA;
    }
    if (zero) {
        e = cudaMemset(q, 0, count * sizeof(T));
        if (e != cudaSuccess) {
            set_error("cudaMemset failed: %s", cudaGetErrorString(e));
            return QEXXC_ERR_CUDA;
        }
    }
    c->allocs.push_back(q);
    c->bytes += count * sizeof(T);
    *p = static_cast<T*>(q);
    return QEXXC_OK;
}

// frees a buffer obtained from dev_alloc (count = the element count it was allocated with)
template <typename T>
static void dev_free(qexxc_ctx* c, T** p, size_t count) {
    if (!*p) return;
    if (count == 0) count = 1;
    for (size_t k = 0; k < c->allocs.size(); ++k)
        if (c->allocs[k] == (void*)*p) {
            c->allocs.erase(c->allocs.begin() + k);
            break;
        }
    cudaFree(*p);
    c->bytes -= count * sizeof(T);
    *p = nullptr;
}

static int ncomp_of(int xctype) { return xctype == QEXXC_XC_GGA ? 4 : 1; }

static int check_xctype(const qexxc_ctx* c, int xctype, bool need_net) {
    if (xctype < QEXXC_XC_NN || xctype > QEXXC_XC_GGA) {
        set_error("unknown xctype %d", xctype);
        return QEXXC_ERR_ARG;
    }
    if (ncomp_of(xctype) > c->C) {
        set_error("xctype GGA needs a context created with ncomp = 4");
        return QEXXC_ERR_STATE;
    }
    if (!need_net) return QEXXC_OK;
    const int k = c->net.kind;
    if (k == QEXXC_NET_NONE) {
        set_error("this context has no network (kind NONE): use qexxc_vxc_assemble with caller-supplied exc/vxc");
        return QEXXC_ERR_STATE;
    }
    if (xctype == QEXXC_XC_NN_GLOBAL && k != QEXXC_NET_GLOBAL_MLP) {
        set_error("xctype NN-AmplitudeEncoding needs a global network");
        return QEXXC_ERR_UNSUPPORTED;
    }
    if (xctype != QEXXC_XC_NN_GLOBAL && k == QEXXC_NET_GLOBAL_MLP) {
        set_error("a global network can only be used with xctype NN-AmplitudeEncoding");
        return QEXXC_ERR_UNSUPPORTED;
    }
    if (xctype == QEXXC_XC_GGA && (k != QEXXC_NET_LOCAL_MLP || c->net.n_features != 2)) {
        set_error("xctype GGA needs a LocalMLP with n_features = 2 (rho, sigma)");
        return QEXXC_ERR_UNSUPPORTED;
    }
    if (xctype == QEXXC_XC_NN && k == QEXXC_NET_LOCAL_MLP && c->net.n_features != 1) {
        set_error("xctype NN needs a LocalMLP with n_features = 1");
        return QEXXC_ERR_UNSUPPORTED;
    }
    return QEXXC_OK;
}

// ---- the network on the context's internal buffers ---------------------------------------------
static int net_fwd(qexxc_ctx* c, int xctype, const double* rho, const double* theta, double* exc, double* vrho,
                   double* vgamma, cudaStream_t st) {
    const long ld = c->GpadMax;
    switch (c->net.kind) {
        case QEXXC_NET_LOCAL_MLP:
            return launch_mlp_local_fwd(c, xctype, rho, (long)c->C * ld, ld, theta, exc, vrho, vgamma, ld, c->B,
                                        c->Gpad, st);
        case QEXXC_NET_GLOBAL_MLP:
            return launch_global_mlp(c, false, rho, (long)c->C * ld, c->G, theta, exc, ld, vrho, nullptr, nullptr,
                                     nullptr, nullptr, 0, c->B, st);
        case QEXXC_NET_LOCAL_QNN:
            for (int b = 0; b < c->B; ++b)
                QX_TRY(launch_qnn(c, false, c->qperm, rho + (long)b * c->C * ld, c->Gpad, theta, exc + b * ld,
                                  vrho + b * ld, nullptr, nullptr, nullptr, 0, nullptr, 0, st));
            return QEXXC_OK;
        default:
            set_error("no network in this context");
            return QEXXC_ERR_STATE;
    }
}

static int net_vjp(qexxc_ctx* c, int xctype, const double* rho, const double* theta, const double* excb,
                   const double* vrhob, const double* vgammab, double* rbar, int accumulate, double* theta_bar,
                   cudaStream_t st) {
    const long ld = c->GpadMax;
    switch (c->net.kind) {
        case QEXXC_NET_LOCAL_MLP:
            return launch_mlp_local_vjp(c, xctype, rho, (long)c->C * ld, ld, theta, excb, vrhob, vgammab, ld, rbar,
                                        accumulate, theta_bar, 0, c->B, c->Gpad, st);
        case QEXXC_NET_GLOBAL_MLP:
            // the direct part of rho_bar is zero for the global branch (excsum is not weighted by rho,
            // numint_legacy.py:331), so the network's rho_bar simply overwrites rows [0, G)
            return launch_global_mlp(c, true, rho, (long)c->C * ld, c->G, theta, nullptr, ld, nullptr, excb, vrhob,
                                     rbar, theta_bar, 0, c->B, st);
        case QEXXC_NET_LOCAL_QNN:
            for (int b = 0; b < c->B; ++b)
                QX_TRY(launch_qnn(c, true, c->qperm, rho + (long)b * c->C * ld, c->Gpad, theta, nullptr, nullptr,
                                  excb + b * ld, vrhob + b * ld, rbar + (long)b * c->C * ld, accumulate, theta_bar,
                                  b > 0 ? 1 : 0, st));
            return QEXXC_OK;
        default:
            set_error("no network in this context");
            return QEXXC_ERR_STATE;
    }
}

static int assemble_vmat(qexxc_ctx* c, int xctype, const double* wv, double* out, cudaStream_t st) {
    const long ld = c->GpadMax;
    const long ob = (long)c->N * c->N + 2;
    if (xctype == QEXXC_XC_GGA) {
        QX_TRY(launch_build_aow(c, wv, (long)c->C * ld, ld, kFacOne, st));
        return launch_wsyrk(c, nullptr, 0, c->aow, 1.0, 1, out, ob, st);
    }
    return launch_wsyrk(c, wv, (long)c->C * ld, nullptr, 1.0, 1, out, ob, st);
}

// stages 3 and 4 of nr_rks on the context's rho buffer, plus the residual copy-out
static int nr_rks_after_rho(qexxc_ctx* c, int xctype, const double* theta_dev, double* out_dev, double* resid_dev,
                            cudaStream_t st) {
    const long ld = c->GpadMax;
    // stage 3: exc, vrho (, vgamma)                    numint_legacy.py:295-303
    QX_TRY(net_fwd(c, xctype, c->rho, theta_dev, c->exc, c->vrho, c->vgamma, st));
    // stage 4: nelec, excsum, wv, vmat + vmat.T        numint_legacy.py:304-309, 336-337
    const long ob = (long)c->N * c->N + 2;
    QX_TRY(launch_stage4_pointwise(c, xctype, c->rho, c->exc, c->vrho, c->vgamma, c->wv,
                                   out_dev + (long)c->N * c->N, ob, st));
    QX_TRY(assemble_vmat(c, xctype, c->wv, out_dev, st));
    if (resid_dev) {
        const size_t n = (size_t)c->B * ld;
        QX_CUDA(cudaMemcpyAsync(resid_dev, c->rho, sizeof(double) * n * c->C, cudaMemcpyDeviceToDevice, st));
        QX_CUDA(cudaMemcpyAsync(resid_dev + n * c->C, c->exc, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        QX_CUDA(cudaMemcpyAsync(resid_dev + n * (c->C + 1), c->vrho, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
        QX_CUDA(cudaMemcpyAsync(resid_dev + n * (c->C + 2), c->vgamma, sizeof(double) * n, cudaMemcpyDeviceToDevice,
                                st));
    }
    return QEXXC_OK;
}

// theta length check (the reference raises a shape error when the parameter tree does not fit the network)
static int check_theta(const qexxc_ctx* c, long n_theta, long G) {
    const long want = qexxc_n_params(&c->net, (int)G);
    if (n_theta != want) {
        set_error("theta has %ld parameters but the context's network needs %ld (kind %d, n_features %d, n_hidden %d, "
                  "width %d, grid %ld)", n_theta, want, c->net.kind, c->net.n_features, c->net.n_hidden, c->net.width, G);
        return QEXXC_ERR_ARG;
    }
    return QEXXC_OK;
}

static int need_ao(const qexxc_ctx* c, int ncomp) {
    if (!c->have_grid) {
        set_error("qexxc_set_grid has not been called");
        return QEXXC_ERR_STATE;
    }
    if (c->ao_ncomp < ncomp) {
        set_error("AO tensor holds %d component(s), %d needed: call qexxc_eval_ao / qexxc_set_ao first", c->ao_ncomp,
                  ncomp);
        return QEXXC_ERR_STATE;
    }
    return QEXXC_OK;
}

}  // namespace qexxc

using namespace qexxc;

extern "C" {

int qexxc_version(void) { return QEXXC_VERSION; }
const char* qexxc_last_error(void) { return g_err; }

long qexxc_n_params(const qexxc_net_desc* net, int ngrids) {
    if (!net) return -1;
    switch (net->kind) {
        case QEXXC_NET_NONE: return 0;
        case QEXXC_NET_LOCAL_MLP: {
            const long F = net->n_features, H = net->width, L = net->n_hidden;
            return F * H + H + (L - 1) * (H * H + H) + H + 1;
        }
        case QEXXC_NET_GLOBAL_MLP: {
            const long F = ngrids, H = net->width, L = net->n_hidden;
            return F * H + H + (L - 1) * (H * H + H) + H + 1;
        }
        case QEXXC_NET_LOCAL_QNN: return 3L * net->width * net->n_hidden;
        default: return -1;
    }
}

int qexxc_create(qexxc_ctx** out, int device, int nbatch, int ncomp, int ngrids_max, int nao,
                 const qexxc_net_desc* net) {
    return qexxc_create_ex(out, device, nbatch, ncomp, ngrids_max, nao, net, 0u);
}

int qexxc_create_ex(qexxc_ctx** out, int device, int nbatch, int ncomp, int ngrids_max, int nao,
                    const qexxc_net_desc* net, unsigned flags) {
    QX_ARG(out != nullptr, "ctx out pointer is null");
    QX_ARG((flags & ~(unsigned)QEXXC_FLAG_SHARED_AO) == 0, "unknown flag bits");
    *out = nullptr;
    QX_ARG(nbatch >= 1 && ngrids_max >= 1 && nao >= 1, "nbatch, ngrids_max, nao must be >= 1");
    QX_ARG(ncomp == 1 || ncomp == 4, "ncomp must be 1 or 4");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available: libqexxc has no CPU fallback");
        (void)cudaGetLastError();
        return QEXXC_ERR_NODEVICE;
    }
    QX_ARG(device >= 0 && device < ndev, "device index out of range");
    QX_CUDA(cudaSetDevice(device));
    qexxc_ctx* c = new qexxc_ctx();
    c->device = device;
    c->ao_shared = (flags & QEXXC_FLAG_SHARED_AO) != 0;
    c->B = nbatch;
    c->C = ncomp;
    c->Gmax = ngrids_max;
    c->GpadMax = round_up(ngrids_max, kGTile);
    c->N = nao;
    c->Npad = round_up(nao, kNTile);
    c->Nc = round_up(nao, kNBlock);
    if (net) c->net = *net;
    else c->net.kind = QEXXC_NET_NONE;
    c->n_theta = qexxc_n_params(&c->net, ngrids_max);
    if (c->n_theta < 0) {
        set_error("unknown network kind %d", c->net.kind);
        delete c;
        return QEXXC_ERR_ARG;
    }
    cudaDeviceProp prop;
    QX_CUDA(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    if (prop.major < 10) {
        set_error("libqexxc is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        delete c;
        return QEXXC_ERR_NODEVICE;
    }

    const size_t B = c->B, C = c->C, Gp = c->GpadMax, Np = c->Npad;
    int rc = QEXXC_OK;
#define QX_A(ptr, count)                                  \
    if (rc == QEXXC_OK) rc = dev_alloc(c, &(ptr), (count))
    QX_A(c->coords, B * Gp * 3);
    QX_A(c->weights, B * Gp);
    QX_A(c->ao, (c->ao_shared ? 1 : B) * C * Gp * Np);
    QX_A(c->S, B * Np * Np);
    QX_A(c->mosgn, B * Np);
    QX_A(c->rho, B * C * Gp);
    QX_A(c->exc, B * Gp);
    QX_A(c->vrho, B * Gp);
    QX_A(c->vgamma, B * Gp);
    QX_A(c->wv, B * C * Gp);
    QX_A(c->wvb, B * C * Gp);
    QX_A(c->rbar, B * C * Gp);
    QX_A(c->excb, B * Gp);
    QX_A(c->vrhob, B * Gp);
    QX_A(c->vgammab, B * Gp);
    if (C == 4) QX_A(c->aow, B * Gp * Np);
    QX_A(c->rq_part, (size_t)((c->Nc + 31) / 32) * 4 * c->num_sms * 128);
    if (B == 1 && (size_t)c->Npad * 128 * 8 * c->num_sms >= ((size_t)96 << 20)) QX_A(c->rq_pair, (size_t)8 * Gp);
    {
        size_t part_doubles = 0;
        wsyrk_workspace(c->num_sms, c->Nc, c->GpadMax, c->B, C == 4, &part_doubles, &c->ws_items_bytes,
                        &c->ws_start_cap);
        QX_A(c->part, part_doubles);
        c->part_doubles = part_doubles;
        for (int k = 0; k < 2; ++k) {
            unsigned char* p = nullptr;
            if (rc == QEXXC_OK) rc = dev_alloc(c, &p, c->ws_items_bytes);
            c->ws_items[k] = p;
            QX_A(c->ws_start[k], c->ws_start_cap);
        }
    }
    size_t red = 2 * B * (size_t)stage4_nblocks(c) + 16;
    if (c->net.kind == QEXXC_NET_LOCAL_MLP && mlp_is_wide(c->net)) {
        if (c->net.n_features < 1 || c->net.n_features > 2 || c->net.width < 1 || c->net.n_hidden < 1) {
            set_error("LocalMLP: n_features must be 1 or 2, n_neurons and n_layers positive");
            rc = QEXXC_ERR_UNSUPPORTED;
        }
        QX_A(c->wide_ws, mlp_wide_ws_doubles(c));
    } else if (c->net.kind == QEXXC_NET_LOCAL_MLP) {
        const size_t r = (size_t)mlp_local_grid(c) * c->n_theta;
        if (r > red) red = r;
        c->tape_bytes = mlp_local_tape_bytes(c);
        unsigned char* tp = nullptr;
        if (rc == QEXXC_OK) rc = dev_alloc(c, &tp, c->tape_bytes);
        c->tape = tp;
    } else if (c->net.kind == QEXXC_NET_GLOBAL_MLP) {
        const size_t r = B * (size_t)c->n_theta;
        if (r > red) red = r;
    } else if (c->net.kind == QEXXC_NET_LOCAL_QNN) {
        const size_t r = qnn_red_doubles(c, (long)B * Gp);
        if (r > red) red = r;
        QX_A(c->qperm, 512);
        if (rc == QEXXC_OK) {
            if (c->net.width < 2 || c->net.width > 8) {
                set_error("LocalQNN: n_qubits=%d not supported (2..8)", c->net.width);
                rc = QEXXC_ERR_UNSUPPORTED;
            } else {
                rc = qnn_upload_perm(c, c->qperm);
            }
        }
    }
    c->red_doubles = red;
    QX_A(c->red, red);
#undef QX_A
    // INT8 contraction workspace (digit planes): reserved here when the policy selects that path for this shape, so that no
    // call on the hot path allocates (CUDA-graph capture); a later QEXXC_I8=1 still allocates on first use
    if (rc == QEXXC_OK && i8_enabled(c)) rc = i8_reserve(c);
    if (rc != QEXXC_OK) {
        qexxc_destroy(c);
        return rc;
    }
    *out = c;
    return QEXXC_OK;
}

int qexxc_destroy(qexxc_ctx* c) {
    if (!c) return QEXXC_OK;
    cudaSetDevice(c->device);
    for (void* p : c->allocs) cudaFree(p);
    i8_release(c);
    delete c;
    return QEXXC_OK;
}

size_t qexxc_workspace_bytes(const qexxc_ctx* c) { return c ? c->bytes : 0; }
long qexxc_launch_count(const qexxc_ctx* c) { return c ? c->launches : 0; }

// ---- stage 1 --------------------------------------------------------------------------------
int qexxc_set_grid(qexxc_ctx* c, const double* coords_dev, const double* weights_dev, int ngrids, void* stream) {
    QX_ARG(c != nullptr, "ctx is null");
    QX_ARG(weights_dev != nullptr || ngrids == 0, "weights pointer is null");
    QX_ARG(ngrids >= 0 && ngrids <= c->Gmax, "ngrids exceeds the ngrids_max the context was created with");
    QX_CUDA(cudaSetDevice(c->device));
    c->G = ngrids;
    c->Gpad = round_up(ngrids > 0 ? ngrids : 1, kGTile);
    c->ao_ncomp = 0;
    c->i8_valid = false;
    QX_TRY(launch_set_grid(c, coords_dev, weights_dev, ngrids, (cudaStream_t)stream));
    c->have_grid = true;
    return QEXXC_OK;
}

int qexxc_set_basis(qexxc_ctx* c, const int* atm, int natm, const int* bas, int nbas, const double* env, int nenv) {
    QX_ARG(c != nullptr, "ctx is null");
    QX_ARG(atm && bas && env && natm > 0 && nbas > 0 && nenv > 0, "null/empty basis tables");
    QX_CUDA(cudaSetDevice(c->device));
    const int nb = c->ao_shared ? 1 : c->B;  // geometries held: one per batch element unless the AO tensor is shared
    std::vector<ShellDev> sh(nbas);
    std::vector<AoMeta> meta(c->Npad, AoMeta{0, 0, -1});
    std::vector<int> acoord(natm), shatom(nbas);
    for (int ia = 0; ia < natm; ++ia) acoord[ia] = atm[6 * ia + 1];
    int off = 0, rad = 0, lmax = 0;
    for (int ib = 0; ib < nbas; ++ib) {
        const int* b = bas + 8 * ib;  // ATOM_OF, ANG_OF, NPRIM_OF, NCTR_OF, KAPPA_OF, PTR_EXP, PTR_COEFF
        const int ia = b[0], l = b[1], nprim = b[2], nctr = b[3];
        QX_ARG(ia >= 0 && ia < natm, "bas: atom index out of range");
        if (l < 0 || l > 3) {
            set_error("shell %d has angular momentum %d; only l <= 3 is supported", ib, l);
            return QEXXC_ERR_UNSUPPORTED;
        }
        const int pc = atm[6 * ia + 1];  // PTR_COORD
        QX_ARG(pc >= 0 && pc + 3 <= nenv && b[5] >= 0 && b[5] + nprim <= nenv && b[6] >= 0 &&
                   b[6] + nprim * nctr <= nenv,
               "bas/atm pointers exceed env");
        sh[ib] = ShellDev{pc, l, nprim, nctr, b[5], b[6], off, rad};
        shatom[ib] = ia;
        if (l > lmax) lmax = l;
        for (int ic = 0; ic < nctr; ++ic)
            for (int m = 0; m < 2 * l + 1; ++m) {
                const int n = off + ic * (2 * l + 1) + m;
                if (n < c->Npad) meta[n] = AoMeta{rad + ic, (short)ia, (short)(l * l + m)};
            }
        off += (2 * l + 1) * nctr;
        rad += nctr;
    }
    QX_ARG(natm < 32768, "too many atoms");
    if (off != c->N) {
        set_error("basis has %d spherical AOs but the context was created with nao = %d", off, c->N);
        return QEXXC_ERR_ARG;
    }
    // (re)allocate the basis tables; a replaced buffer is freed and dropped from the context's bookkeeping
    if (c->nshell != nbas || !c->shells || !c->shell_atom) {
        dev_free(c, &c->shells, (size_t)c->nshell);
        dev_free(c, &c->shell_atom, (size_t)c->nshell);
        QX_TRY(dev_alloc(c, &c->shells, (size_t)nbas, false));
        QX_TRY(dev_alloc(c, &c->shell_atom, (size_t)nbas, false));
        c->nshell = nbas;
    }
    if (c->nenv != nenv || !c->env) {
        dev_free(c, &c->env, (size_t)nb * c->nenv);
        QX_TRY(dev_alloc(c, &c->env, (size_t)nb * nenv, false));
        c->nenv = nenv;
    }
    if (!c->ao_meta) QX_TRY(dev_alloc(c, &c->ao_meta, (size_t)c->Npad, false));
    if (c->natm != natm || !c->atom_coord) {
        dev_free(c, &c->atom_coord, (size_t)c->natm);
        QX_TRY(dev_alloc(c, &c->atom_coord, (size_t)natm, false));
    }
    c->natm = natm;
    c->nrad = rad;
    c->lmax = lmax;
    QX_CUDA(cudaMemcpy(c->ao_meta, meta.data(), sizeof(AoMeta) * c->Npad, cudaMemcpyHostToDevice));
    QX_CUDA(cudaMemcpy(c->atom_coord, acoord.data(), sizeof(int) * natm, cudaMemcpyHostToDevice));
    QX_CUDA(cudaMemcpy(c->shell_atom, shatom.data(), sizeof(int) * nbas, cudaMemcpyHostToDevice));
    QX_CUDA(cudaMemcpy(c->shells, sh.data(), sizeof(ShellDev) * nbas, cudaMemcpyHostToDevice));
    QX_CUDA(cudaMemcpy(c->env, env, sizeof(double) * (size_t)nb * nenv, cudaMemcpyHostToDevice));
    c->have_basis = true;
    c->ao_ncomp = 0;
    c->i8_valid = false;
    return QEXXC_OK;
}

int qexxc_eval_ao(qexxc_ctx* c, int deriv, void* stream) {
    QX_ARG(c != nullptr, "ctx is null");
    if (deriv != 0 && deriv != 1) {
        set_error("eval_ao: deriv must be 0 or 1");
        return QEXXC_ERR_UNSUPPORTED;
    }
    if (!c->have_grid || !c->have_basis) {
        set_error("eval_ao needs qexxc_set_grid (with coords) and qexxc_set_basis first");
        return QEXXC_ERR_STATE;
    }
    if (deriv == 1 && c->C < 4) {
        set_error("eval_ao deriv=1 needs a context created with ncomp = 4");
        return QEXXC_ERR_STATE;
    }
    QX_CUDA(cudaSetDevice(c->device));
    QX_TRY(launch_eval_ao(c, deriv, (cudaStream_t)stream));
    c->ao_ncomp = deriv ? 4 : 1;
    c->i8_valid = false;
    return QEXXC_OK;
}

int qexxc_set_ao(qexxc_ctx* c, const double* ao_dev, int ncomp, int ngrids, void* stream) {
    QX_ARG(c != nullptr && (ao_dev != nullptr || ngrids == 0), "null pointer");
    QX_ARG(ncomp == 1 || ncomp == 4, "ncomp must be 1 or 4");
    QX_ARG(ncomp <= c->C, "ncomp exceeds the context's ncomp");
    if (!c->have_grid || ngrids != c->G) {
        set_error("set_ao: call qexxc_set_grid with the same ngrids first");
        return QEXXC_ERR_STATE;
    }
    QX_CUDA(cudaSetDevice(c->device));
    QX_TRY(launch_pack_ao(c, ao_dev, ncomp, ngrids, (cudaStream_t)stream));
    c->ao_ncomp = ncomp;
    c->i8_valid = false;
    return QEXXC_OK;
}

int qexxc_get_ao(qexxc_ctx* c, double* ao_dev, int ncomp, void* stream) {
    QX_ARG(c != nullptr && ao_dev != nullptr, "null pointer");
    QX_TRY(need_ao(c, ncomp));
    QX_CUDA(cudaSetDevice(c->device));
    return launch_unpack_ao(c, ao_dev, ncomp, (cudaStream_t)stream);
}

// ---- stage 2 --------------------------------------------------------------------------------
int qexxc_eval_rho(qexxc_ctx* c, const double* dm_dev, int ncomp, int hermi, double* rho_dev, void* stream) {
    QX_ARG(c != nullptr && dm_dev != nullptr && rho_dev != nullptr, "null pointer");
    QX_ARG(ncomp == 1 || ncomp == 4, "ncomp must be 1 (LDA) or 4 (GGA)");
    QX_TRY(need_ao(c, ncomp));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    // LDA: rho is a quadratic form, which only sees the symmetric part of dm, so symmetrising is
    // exact for hermi = 0 and 1 alike and lets rowquad use the upper triangle only
    QX_TRY(launch_pad_sym(c, dm_dev, (ncomp == 1 || !hermi) ? 0 : 1, ncomp == 1, st));
    QX_TRY(launch_rowquad(c, ncomp, ncomp == 1, kFacGGA, c->rho, (long)c->C * ld, ld, st));
    return launch_copy_rows(c, rho_dev, (long)ncomp * c->G, c->G, c->rho, (long)c->C * ld, ld, c->B, ncomp, c->G,
                            c->G, st);
}

int qexxc_eval_rho_vjp(qexxc_ctx* c, const double* rho_bar_dev, int ncomp, int hermi, double* dm_bar_dev,
                       void* stream) {
    QX_ARG(c != nullptr && rho_bar_dev != nullptr && dm_bar_dev != nullptr, "null pointer");
    QX_ARG(ncomp == 1 || ncomp == 4, "ncomp must be 1 (LDA) or 4 (GGA)");
    QX_TRY(need_ao(c, ncomp));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    QX_TRY(launch_copy_rows(c, c->rbar, (long)c->C * ld, ld, rho_bar_dev, (long)ncomp * c->G, c->G, c->B, ncomp,
                            c->G, c->Gpad, st));
    const long nn = (long)c->N * c->N;
    if (ncomp == 1) return launch_wsyrk(c, c->rbar, (long)c->C * ld, nullptr, 0.5, 1, dm_bar_dev, nn, st);
    QX_TRY(launch_build_aow(c, c->rbar, (long)c->C * ld, ld, kFacGGA, st));
    return launch_wsyrk(c, nullptr, 0, c->aow, hermi ? 1.0 : 0.5, hermi ? 0 : 1, dm_bar_dev, nn, st);
}

// ---- stage 3 --------------------------------------------------------------------------------
int qexxc_xc_fwd(qexxc_ctx* c, int xctype, const double* rho_dev, const double* theta_dev, long n_theta,
                 double* exc_dev, double* vrho_dev, double* vgamma_dev, void* stream) {
    QX_ARG(c != nullptr && rho_dev && theta_dev && exc_dev && vrho_dev, "null pointer");
    QX_TRY(check_xctype(c, xctype, true));
    if (!c->have_grid) {
        set_error("qexxc_set_grid has not been called");
        return QEXXC_ERR_STATE;
    }
    QX_TRY(check_theta(c, n_theta, c->G));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    const int nc = ncomp_of(xctype);
    QX_TRY(launch_copy_rows(c, c->rho, (long)c->C * ld, ld, rho_dev, (long)nc * c->G, c->G, c->B, nc, c->G, c->Gpad,
                            st));
    QX_TRY(net_fwd(c, xctype, c->rho, theta_dev, c->exc, c->vrho, c->vgamma, st));
    if (xctype == QEXXC_XC_NN_GLOBAL)
        QX_TRY(launch_copy_rows(c, exc_dev, 1, 1, c->exc, ld, ld, c->B, 1, 1, 1, st));
    else
        QX_TRY(launch_copy_rows(c, exc_dev, c->G, c->G, c->exc, ld, ld, c->B, 1, c->G, c->G, st));
    QX_TRY(launch_copy_rows(c, vrho_dev, c->G, c->G, c->vrho, ld, ld, c->B, 1, c->G, c->G, st));
    if (xctype == QEXXC_XC_GGA && vgamma_dev)
        QX_TRY(launch_copy_rows(c, vgamma_dev, c->G, c->G, c->vgamma, ld, ld, c->B, 1, c->G, c->G, st));
    return QEXXC_OK;
}

int qexxc_xc_vjp(qexxc_ctx* c, int xctype, const double* rho_dev, const double* theta_dev, long n_theta,
                 const double* exc_bar_dev, const double* vrho_bar_dev, const double* vgamma_bar_dev,
                 double* rho_bar_dev, double* theta_bar_dev, void* stream) {
    QX_ARG(c != nullptr && rho_dev && theta_dev && exc_bar_dev && vrho_bar_dev && rho_bar_dev && theta_bar_dev,
           "null pointer");
    QX_TRY(check_xctype(c, xctype, true));
    if (!c->have_grid) {
        set_error("qexxc_set_grid has not been called");
        return QEXXC_ERR_STATE;
    }
    QX_ARG(xctype != QEXXC_XC_GGA || vgamma_bar_dev != nullptr, "vgamma_bar is null for xctype GGA");
    QX_TRY(check_theta(c, n_theta, c->G));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    const int nc = ncomp_of(xctype);
    QX_TRY(launch_copy_rows(c, c->rho, (long)c->C * ld, ld, rho_dev, (long)nc * c->G, c->G, c->B, nc, c->G, c->Gpad,
                            st));
    if (xctype == QEXXC_XC_NN_GLOBAL)
        QX_TRY(launch_copy_rows(c, c->excb, ld, ld, exc_bar_dev, 1, 1, c->B, 1, 1, 1, st));
    else
        QX_TRY(launch_copy_rows(c, c->excb, ld, ld, exc_bar_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    QX_TRY(launch_copy_rows(c, c->vrhob, ld, ld, vrho_bar_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    if (xctype == QEXXC_XC_GGA)
        QX_TRY(launch_copy_rows(c, c->vgammab, ld, ld, vgamma_bar_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    QX_TRY(net_vjp(c, xctype, c->rho, theta_dev, c->excb, c->vrhob, c->vgammab, c->rbar, 0, theta_bar_dev, st));
    return launch_copy_rows(c, rho_bar_dev, (long)nc * c->G, c->G, c->rbar, (long)c->C * ld, ld, c->B, nc, c->G, c->G,
                            st);
}

static int apply_fn_common(qexxc_ctx* c, bool vjp, const double* x_dev, long npts, const double* theta_dev,
                           double* y_dev, const double* y_bar_dev, double* x_bar_dev, double* theta_bar_dev,
                           cudaStream_t st) {
    const long cap = (long)c->B * c->GpadMax;
    const int kind = c->net.kind;
    if (kind == QEXXC_NET_NONE) {
        set_error("this context has no network");
        return QEXXC_ERR_STATE;
    }
    if (kind == QEXXC_NET_GLOBAL_MLP) {
        // x [ngrids] -> y [1]; the parameter count fixes ngrids
        QX_ARG(npts >= 1 && npts <= c->Gmax, "global apply_fn: npts must be <= ngrids_max");
        const long npad = round_up(npts, kGTile);
        QX_TRY(launch_copy_rows(c, c->rho, 0, 0, x_dev, 0, 0, 1, 1, npts, npad, st));
        if (!vjp) {
            QX_TRY(launch_global_mlp(c, false, c->rho, 0, (int)npts, theta_dev, c->exc, 0, c->vrho, nullptr, nullptr,
                                     nullptr, nullptr, 0, 1, st));
            return launch_copy_rows(c, y_dev, 0, 0, c->exc, 0, 0, 1, 1, 1, 1, st);
        }
        QX_TRY(launch_copy_rows(c, c->excb, 0, 0, y_bar_dev, 0, 0, 1, 1, 1, 1, st));
        QX_CUDA(cudaMemsetAsync(c->vrhob, 0, sizeof(double) * npad, st));
        QX_TRY(launch_global_mlp(c, true, c->rho, 0, (int)npts, theta_dev, nullptr, 0, nullptr, c->excb, c->vrhob,
                                 c->rbar, theta_bar_dev, 0, 1, st));
        return launch_copy_rows(c, x_bar_dev, 0, 0, c->rbar, 0, 0, 1, 1, npts, npts, st);
    }
    const int F = kind == QEXXC_NET_LOCAL_MLP ? c->net.n_features : 1;
    const long npad = round_up(npts, kGTile);
    if (npts < 1 || npad > cap || (long)F * npad > (long)c->B * c->C * c->GpadMax) {
        set_error("apply_fn: npts=%ld exceeds the context capacity (%ld points)", npts, cap);
        return QEXXC_ERR_ARG;
    }
    QX_TRY(launch_transpose_in(c, c->rho, npad, x_dev, npts, F, npad, st));
    if (!vjp) {
        if (kind == QEXXC_NET_LOCAL_MLP)
            QX_TRY(launch_mlp_local_fwd(c, QEXXC_XC_NN, c->rho, 0, npad, theta_dev, c->exc, c->vrho, c->vgamma, 0, 1,
                                        npad, st));
        else
            QX_TRY(launch_qnn(c, false, c->qperm, c->rho, npts, theta_dev, c->exc, nullptr, nullptr, nullptr, nullptr,
                              0, nullptr, 0, st));
        return launch_copy_rows(c, y_dev, 0, 0, c->exc, 0, 0, 1, 1, npts, npts, st);
    }
    QX_TRY(launch_copy_rows(c, c->excb, 0, 0, y_bar_dev, 0, 0, 1, 1, npts, npad, st));
    QX_CUDA(cudaMemsetAsync(c->vrhob, 0, sizeof(double) * npad, st));
    QX_CUDA(cudaMemsetAsync(c->vgammab, 0, sizeof(double) * npad, st));
    if (kind == QEXXC_NET_LOCAL_MLP)
        QX_TRY(launch_mlp_local_vjp(c, QEXXC_XC_NN, c->rho, 0, npad, theta_dev, c->excb, c->vrhob, c->vgammab, 0,
                                    c->rbar, 0, theta_bar_dev, 0, 1, npad, st));
    else
        QX_TRY(launch_qnn(c, true, c->qperm, c->rho, npts, theta_dev, nullptr, nullptr, c->excb, c->vrhob, c->rbar, 0,
                          theta_bar_dev, 0, st));
    return launch_transpose_out(c, x_bar_dev, c->rbar, npad, npts, F, st);
}

int qexxc_apply_fn_fwd(qexxc_ctx* c, const double* x_dev, long npts, const double* theta_dev, long n_theta,
                       double* y_dev, void* stream) {
    QX_ARG(c != nullptr && x_dev && theta_dev && y_dev, "null pointer");
    QX_TRY(check_theta(c, n_theta, npts));
    QX_CUDA(cudaSetDevice(c->device));
    return apply_fn_common(c, false, x_dev, npts, theta_dev, y_dev, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int qexxc_apply_fn_vjp(qexxc_ctx* c, const double* x_dev, long npts, const double* theta_dev, long n_theta,
                       const double* y_bar_dev, double* x_bar_dev, double* theta_bar_dev, void* stream) {
    QX_ARG(c != nullptr && x_dev && theta_dev && y_bar_dev && x_bar_dev && theta_bar_dev, "null pointer");
    QX_TRY(check_theta(c, n_theta, npts));
    QX_CUDA(cudaSetDevice(c->device));
    return apply_fn_common(c, true, x_dev, npts, theta_dev, nullptr, y_bar_dev, x_bar_dev, theta_bar_dev,
                           (cudaStream_t)stream);
}

// ---- stage 4 --------------------------------------------------------------------------------
int qexxc_vxc_assemble(qexxc_ctx* c, int xctype, const double* rho_dev, const double* exc_dev,
                       const double* vrho_dev, const double* vgamma_dev, double* out_dev, void* stream) {
    QX_ARG(c != nullptr && rho_dev && exc_dev && vrho_dev && out_dev, "null pointer");
    QX_TRY(check_xctype(c, xctype, false));
    QX_ARG(xctype != QEXXC_XC_GGA || vgamma_dev != nullptr, "vgamma is null for xctype GGA");
    const int nc = ncomp_of(xctype);
    QX_TRY(need_ao(c, nc));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    QX_TRY(launch_copy_rows(c, c->rho, (long)c->C * ld, ld, rho_dev, (long)nc * c->G, c->G, c->B, nc, c->G, c->Gpad,
                            st));
    if (xctype == QEXXC_XC_NN_GLOBAL)
        QX_TRY(launch_copy_rows(c, c->exc, ld, ld, exc_dev, 1, 1, c->B, 1, 1, 1, st));
    else
        QX_TRY(launch_copy_rows(c, c->exc, ld, ld, exc_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    QX_TRY(launch_copy_rows(c, c->vrho, ld, ld, vrho_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    if (xctype == QEXXC_XC_GGA)
        QX_TRY(launch_copy_rows(c, c->vgamma, ld, ld, vgamma_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    const long ob = (long)c->N * c->N + 2;
    QX_TRY(launch_stage4_pointwise(c, xctype, c->rho, c->exc, c->vrho, c->vgamma, c->wv,
                                   out_dev + (long)c->N * c->N, ob, st));
    return assemble_vmat(c, xctype, c->wv, out_dev, st);
}

int qexxc_vxc_assemble_vjp(qexxc_ctx* c, int xctype, const double* rho_dev, const double* exc_dev,
                           const double* vrho_dev, const double* vgamma_dev, const double* e_bar_dev,
                           const double* v_bar_dev, double* rho_bar_dev, double* exc_bar_dev,
                           double* vrho_bar_dev, double* vgamma_bar_dev, void* stream) {
    QX_ARG(c != nullptr && rho_dev && exc_dev && vrho_dev && e_bar_dev && v_bar_dev && rho_bar_dev && exc_bar_dev &&
               vrho_bar_dev,
           "null pointer");
    QX_TRY(check_xctype(c, xctype, false));
    QX_ARG(xctype != QEXXC_XC_GGA || (vgamma_dev != nullptr && vgamma_bar_dev != nullptr), "vgamma pointers are null");
    const int nc = ncomp_of(xctype);
    QX_TRY(need_ao(c, nc));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    QX_TRY(launch_copy_rows(c, c->rho, (long)c->C * ld, ld, rho_dev, (long)nc * c->G, c->G, c->B, nc, c->G, c->Gpad,
                            st));
    if (xctype == QEXXC_XC_NN_GLOBAL)
        QX_TRY(launch_copy_rows(c, c->exc, ld, ld, exc_dev, 1, 1, c->B, 1, 1, 1, st));
    else
        QX_TRY(launch_copy_rows(c, c->exc, ld, ld, exc_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    QX_TRY(launch_copy_rows(c, c->vrho, ld, ld, vrho_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    if (xctype == QEXXC_XC_GGA)
        QX_TRY(launch_copy_rows(c, c->vgamma, ld, ld, vgamma_dev, c->G, c->G, c->B, 1, c->G, c->Gpad, st));
    QX_TRY(launch_pad_sym(c, v_bar_dev, 2, nc == 1, st));
    QX_TRY(launch_rowquad(c, nc, nc == 1, kFacOne, c->wvb, (long)c->C * ld, ld, st));
    QX_TRY(launch_stage4_pointwise_vjp(c, xctype, c->rho, c->exc, c->vrho, c->vgamma, e_bar_dev, c->wvb, c->rbar,
                                       c->excb, c->vrhob, c->vgammab, st));
    QX_TRY(launch_copy_rows(c, rho_bar_dev, (long)nc * c->G, c->G, c->rbar, (long)c->C * ld, ld, c->B, nc, c->G, c->G,
                            st));
    if (xctype == QEXXC_XC_NN_GLOBAL)
        QX_TRY(launch_copy_rows(c, exc_bar_dev, 1, 1, c->excb, ld, ld, c->B, 1, 1, 1, st));
    else
        QX_TRY(launch_copy_rows(c, exc_bar_dev, c->G, c->G, c->excb, ld, ld, c->B, 1, c->G, c->G, st));
    QX_TRY(launch_copy_rows(c, vrho_bar_dev, c->G, c->G, c->vrhob, ld, ld, c->B, 1, c->G, c->G, st));
    if (xctype == QEXXC_XC_GGA)
        QX_TRY(launch_copy_rows(c, vgamma_bar_dev, c->G, c->G, c->vgammab, ld, ld, c->B, 1, c->G, c->G, st));
    return QEXXC_OK;
}

// ---- fused hot path -------------------------------------------------------------------------
size_t qexxc_resid_doubles(const qexxc_ctx* c) { return c ? (size_t)c->B * c->GpadMax * (c->C + 3) : 0; }

int qexxc_nr_rks_fwd(qexxc_ctx* c, int xctype, int hermi, const double* dm_dev, const double* theta_dev,
                     long n_theta, double* out_dev, double* resid_dev, void* stream) {
    QX_ARG(c != nullptr && dm_dev && theta_dev && out_dev, "null pointer");
    QX_TRY(check_xctype(c, xctype, true));
    const int nc = ncomp_of(xctype);
    QX_TRY(need_ao(c, nc));
    QX_TRY(check_theta(c, n_theta, c->G));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    // stage 2: rho = rowdot(ao, ao sym(dm))            numint_legacy.py:294 -> :351-397
    QX_TRY(launch_pad_sym(c, dm_dev, (nc == 1 || !hermi) ? 0 : 1, nc == 1, st));
    QX_TRY(launch_rowquad(c, nc, nc == 1, kFacGGA, c->rho, (long)c->C * ld, ld, st));
    return nr_rks_after_rho(c, xctype, theta_dev, out_dev, resid_dev, st);
}

int qexxc_eval_rho_mo(qexxc_ctx* c, const double* mo_coeff_dev, const double* mo_occ_dev, int nmo, double* rho_dev,
                      void* stream) {
    QX_ARG(c != nullptr && mo_coeff_dev && mo_occ_dev && rho_dev, "null pointer");
    QX_ARG(nmo >= 1 && nmo <= c->N, "nmo must be in [1, nao]");
    QX_TRY(need_ao(c, 1));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    const int ldL = round_up(nmo, kNTile);
    QX_TRY(launch_pack_mo(c, mo_coeff_dev, mo_occ_dev, nmo, c->S, c->mosgn, ldL, st));
    QX_TRY(launch_rowquad_mo(c, c->S, ldL, nmo, c->mosgn, c->rho, (long)c->C * ld, st));
    return launch_copy_rows(c, rho_dev, c->G, c->G, c->rho, (long)c->C * ld, ld, c->B, 1, c->G, c->G, st);
}

int qexxc_nr_rks_fwd_mo(qexxc_ctx* c, int xctype, const double* mo_coeff_dev, const double* mo_occ_dev, int nmo,
                        const double* theta_dev, long n_theta, double* out_dev, double* resid_dev, void* stream) {
    QX_ARG(c != nullptr && mo_coeff_dev && mo_occ_dev && theta_dev && out_dev, "null pointer");
    QX_ARG(nmo >= 1 && nmo <= c->N, "nmo must be in [1, nao]");
    QX_TRY(check_xctype(c, xctype, true));
    if (xctype == QEXXC_XC_GGA) {
        set_error("the MO form of rho is implemented for the LDA-type branches (NN, NN-AmplitudeEncoding) only");
        return QEXXC_ERR_UNSUPPORTED;
    }
    QX_TRY(need_ao(c, 1));
    QX_TRY(check_theta(c, n_theta, c->G));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    const int ldL = round_up(nmo, kNTile);
    // stage 2, MO form: rho = sum_k occ_k (ao C_k)^2   numint_legacy.py:527-545 -> pyscf eval_rho2
    QX_TRY(launch_pack_mo(c, mo_coeff_dev, mo_occ_dev, nmo, c->S, c->mosgn, ldL, st));
    QX_TRY(launch_rowquad_mo(c, c->S, ldL, nmo, c->mosgn, c->rho, (long)c->C * ld, st));
    return nr_rks_after_rho(c, xctype, theta_dev, out_dev, resid_dev, st);
}

int qexxc_nr_rks_vjp(qexxc_ctx* c, int xctype, int hermi, const double* theta_dev, long n_theta,
                     const double* resid_dev, const double* e_bar_dev, const double* v_bar_dev, double* bar_dev,
                     void* stream) {
    QX_ARG(c != nullptr && theta_dev && resid_dev && e_bar_dev && v_bar_dev && bar_dev, "null pointer");
    QX_TRY(check_xctype(c, xctype, true));
    const int nc = ncomp_of(xctype);
    QX_TRY(need_ao(c, nc));
    QX_TRY(check_theta(c, n_theta, c->G));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    const size_t n = (size_t)c->B * ld;
    const double* rho = resid_dev;
    const double* exc = resid_dev + n * c->C;
    const double* vrho = resid_dev + n * (c->C + 1);
    const double* vgamma = resid_dev + n * (c->C + 2);
    const long nn = (long)c->N * c->N;
    // adjoint of vmat = H + H^T, H = ao^T aow: wv_bar_c = rowdot(ao_c, ao_0 (V_bar + V_bar^T))
    QX_TRY(launch_pad_sym(c, v_bar_dev, 2, nc == 1, st));
    QX_TRY(launch_rowquad(c, nc, nc == 1, kFacOne, c->wvb, (long)c->C * ld, ld, st));
    // adjoint of the per-point stage-4 algebra
    QX_TRY(launch_stage4_pointwise_vjp(c, xctype, rho, exc, vrho, vgamma, e_bar_dev, c->wvb, c->rbar, c->excb,
                                       c->vrhob, c->vgammab, st));
    // second-order reverse of the network: rbar += ..., theta_bar
    QX_TRY(net_vjp(c, xctype, rho, theta_dev, c->excb, c->vrhob, c->vgammab, c->rbar, 1, bar_dev + (long)c->B * nn,
                   st));
    // adjoint of eval_rho w.r.t. dm
    if (nc == 1) return launch_wsyrk(c, c->rbar, (long)c->C * ld, nullptr, 0.5, 1, bar_dev, nn, st);
    QX_TRY(launch_build_aow(c, c->rbar, (long)c->C * ld, ld, kFacGGA, st));
    return launch_wsyrk(c, nullptr, 0, c->aow, hermi ? 1.0 : 0.5, hermi ? 0 : 1, bar_dev, nn, st);
}

int qexxc_profile_enable(qexxc_ctx* c, int on) {
    QX_ARG(c != nullptr, "ctx is null");
    c->prof = on != 0;
    return QEXXC_OK;
}

int qexxc_profile_read(qexxc_ctx* c, int cls, double* ms_total, long* count) {
    QX_ARG(c != nullptr && ms_total && count, "null pointer");
    QX_ARG(cls >= 0 && cls < QEXXC_PROF_NCLASS, "profile class out of range");
    double tot = 0.0;
    long n = 0;
    for (auto& pr : c->prof_ev[cls]) {
        QX_CUDA(cudaEventSynchronize(pr.second));
        float ms = 0.f;
        QX_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
        ++n;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    c->prof_ev[cls].clear();
    *ms_total = tot;
    *count = n;
    return QEXXC_OK;
}

int qexxc_contraction_flops(qexxc_ctx* c, int which, int symmetric, double* executed) {
    QX_ARG(c != nullptr && executed != nullptr, "null pointer");
    *executed = which == 0 ? rowquad_executed_flops(c, symmetric) : wsyrk_executed_flops(c, symmetric != 0);
    return QEXXC_OK;
}

int qexxc_prepare_contractions(qexxc_ctx* c, void* stream) {
    QX_ARG(c != nullptr, "ctx is null");
    if (!i8_enabled(c)) return QEXXC_OK;
    QX_TRY(need_ao(c, 1));
    QX_CUDA(cudaSetDevice(c->device));
    return i8_prepare_geometry(c, (cudaStream_t)stream);
}

int qexxc_contraction_mode(const qexxc_ctx* c) { return (c && i8_enabled(c)) ? 1 : 0; }

int qexxc_contraction_i8_ops(qexxc_ctx* c, int which, int symmetric, double* executed) {
    QX_ARG(c != nullptr && executed != nullptr, "null pointer");
    *executed = i8_executed_ops(c, which, symmetric != 0);
    return QEXXC_OK;
}

int qexxc_i8_peak(int device, double* ops_per_second) {
    QX_ARG(ops_per_second != nullptr, "null pointer");
    return i8_peak_probe(device, ops_per_second);
}

int qexxc_debug_run_contraction(qexxc_ctx* c, int which, void* stream) {
    QX_ARG(c != nullptr, "ctx is null");
    QX_TRY(need_ao(c, 1));
    QX_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long ld = c->GpadMax;
    if (which == 0) return launch_rowquad(c, 1, 1, kFacGGA, c->wvb, (long)c->C * ld, ld, st);
    if (which == 2) return launch_rowquad(c, 1, 0, kFacGGA, c->wvb, (long)c->C * ld, ld, st);  // dense S (timing only)
    return launch_wsyrk(c, c->wv, (long)c->C * ld, nullptr, 1.0, 1, c->S, (long)c->N * c->N, st);
}

}  // extern "C"

// Stage 1 (K1): atomic-orbital values and gradients on the grid, plus layout helpers.
//
// Replaces the pyscf C evaluator the reference calls on every nr_rks
// (pyscf.gto.eval_gto "GTOval_sph_deriv0/1" via qedft/train/td/eval_gto.py:48-70, reached from
// block_loop at numint_legacy.py:292,313 and numint.eval_ao at trainer_legacy_no_jit.py:273).
// phi_lm(r) = S_lm(r - A) * sum_p c_p exp(-a_p |r-A|^2), S_lm real solid harmonics in pyscf order,
// coefficients as stored in mol._env (already normalised).  Output goes straight into the
// context's padded AO tensor ao[b][c][g][n] (n fastest, like pyscf's [comp, grid, ao]).
#include "common.cuh"

namespace qexxc {
namespace {

struct V4 {
    double v, x, y, z;
};
__device__ __forceinline__ V4 mk(double v, double x, double y, double z) { return V4{v, x, y, z}; }

// real solid harmonic m of order l (value + gradient) at (x,y,z)
template <int L>
__device__ __forceinline__ V4 solid(int m, double x, double y, double z);
template <>
__device__ __forceinline__ V4 solid<0>(int, double, double, double) {
    return mk(0.28209479177387814, 0, 0, 0);
}
template <>
__device__ __forceinline__ V4 solid<1>(int m, double x, double y, double z) {
    const double c = 0.4886025119029199;
    return m == 0 ? mk(c * x, c, 0, 0) : (m == 1 ? mk(c * y, 0, c, 0) : mk(c * z, 0, 0, c));
}
template <>
__device__ __forceinline__ V4 solid<2>(int m, double x, double y, double z) {
    const double c = 1.0925484305920792, a = 0.6307831305050401, h = 0.31539156525252005,
                 e = 0.5462742152960396;
    switch (m) {
        case 0: return mk(c * x * y, c * y, c * x, 0);
        case 1: return mk(c * y * z, 0, c * z, c * y);
        case 2: return mk(a * z * z - h * (x * x + y * y), -2 * h * x, -2 * h * y, 2 * a * z);
        case 3: return mk(c * x * z, c * z, 0, c * x);
        default: return mk(e * (x * x - y * y), 2 * e * x, -2 * e * y, 0);
    }
}
template <>
__device__ __forceinline__ V4 solid<3>(int m, double x, double y, double z) {
    const double A = 0.5900435899266435, B = 2.8906114426405543, C = 0.4570457994644657,
                 D = 0.3731763325901154, E = 1.4453057213202771;
    const double xx = x * x, yy = y * y, zz = z * z;
    switch (m) {
        case 0: return mk(A * (3 * xx * y - yy * y), A * 6 * x * y, A * (3 * xx - 3 * yy), 0);
        case 1: return mk(B * x * y * z, B * y * z, B * x * z, B * x * y);
        case 2: return mk(C * y * (4 * zz - xx - yy), -2 * C * x * y, C * (4 * zz - xx - 3 * yy), 8 * C * y * z);
        case 3:
            return mk(D * z * (2 * zz - 3 * xx - 3 * yy), -6 * D * x * z, -6 * D * y * z,
                      D * (6 * zz - 3 * xx - 3 * yy));
        case 4: return mk(C * x * (4 * zz - xx - yy), C * (4 * zz - 3 * xx - yy), -2 * C * x * y, 8 * C * x * z);
        case 5: return mk(E * z * (xx - yy), 2 * E * x * z, -2 * E * y * z, E * (xx - yy));
        default: return mk(A * (xx * x - 3 * x * yy), A * (3 * xx - 3 * yy), -6 * A * x * y, 0);
    }
}

template <int L>
__device__ __forceinline__ void emit_shell(double* __restrict__ row0, long cstride, int deriv, int off,
                                           double x, double y, double z, double R0, double R1) {
#pragma unroll
    for (int m = 0; m < 2 * L + 1; ++m) {
        const V4 s = solid<L>(m, x, y, z);
        row0[off + m] = s.v * R0;
        if (deriv) {
            row0[cstride + off + m] = s.x * R0 + s.v * x * R1;
            row0[2 * cstride + off + m] = s.y * R0 + s.v * y * R1;
            row0[3 * cstride + off + m] = s.z * R0 + s.v * z * R1;
        }
    }
}

// one warp per grid point; lanes stride over shells
__global__ void __launch_bounds__(256)
eval_ao_kernel(const double* __restrict__ coords, const ShellDev* __restrict__ shells, int nshell,
               const double* __restrict__ env, int nenv, double* __restrict__ ao, int G, int Gpad,
               int GpadMax, int N, int Npad, int C, int deriv) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const long wid = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const long cstride = (long)GpadMax * Npad;
    const double* envb = env + (long)b * nenv;
    const int ncomp = deriv ? 4 : 1;
    for (long gi = wid; gi < Gpad; gi += nwarps) {
        double* row0 = ao + ((long)b * C * GpadMax + gi) * Npad;
        if (gi >= G) {  // padding rows stay exactly zero
            for (int c = 0; c < ncomp; ++c)
                for (int n = lane; n < N; n += 32) row0[c * cstride + n] = 0.0;
            continue;
        }
        const double* r = coords + ((long)b * GpadMax + gi) * 3;
        const double rx = r[0], ry = r[1], rz = r[2];
        for (int s = lane; s < nshell; s += 32) {
            const ShellDev sh = shells[s];
            const double x = rx - envb[sh.atom_coord], y = ry - envb[sh.atom_coord + 1],
                         z = rz - envb[sh.atom_coord + 2];
            const double rr = x * x + y * y + z * z;
            const double* ex = envb + sh.ptr_exp;
            const int nf = 2 * sh.l + 1;
            for (int ic = 0; ic < sh.nctr; ++ic) {
                const double* cf = envb + sh.ptr_coef + ic * sh.nprim;
                double R0 = 0.0, R1 = 0.0;
                for (int p = 0; p < sh.nprim; ++p) {
                    const double a = ex[p];
                    const double e = cf[p] * exp(-a * rr);
                    R0 += e;
                    R1 -= 2.0 * a * e;
                }
                const int off = sh.ao_off + ic * nf;
                switch (sh.l) {
                    case 0: emit_shell<0>(row0, cstride, deriv, off, x, y, z, R0, R1); break;
                    case 1: emit_shell<1>(row0, cstride, deriv, off, x, y, z, R0, R1); break;
                    case 2: emit_shell<2>(row0, cstride, deriv, off, x, y, z, R0, R1); break;
                    default: emit_shell<3>(row0, cstride, deriv, off, x, y, z, R0, R1); break;
                }
            }
        }
    }
}


// ---- tiled evaluator -------------------------------------------------------------------------------
// CTA = P grid points.  Phase A: every (point, atom) gets its 16 real solid harmonics (and their
// gradients), every (point, radial function) its contracted radial sum -- each exp is evaluated
// exactly once and staged in shared memory.  Phase B: warps sweep (point, 32-AO chunk) pairs;
// lane n multiplies harmonic x radial for AO n and the warp stores 256 contiguous bytes of the AO
// row, so the 8*N bytes per point stream to HBM fully coalesced.
template <int L>
__device__ __forceinline__ void fill_harm(double* __restrict__ H, int stride, bool deriv, double x, double y, double z) {
#pragma unroll
    for (int m = 0; m < 2 * L + 1; ++m) {
        const V4 s = solid<L>(m, x, y, z);
        H[(L * L + m) * stride] = s.v;
        if (deriv) {  // rows 17.., 33.., 49.. hold d/dx, d/dy, d/dz
            H[(17 + L * L + m) * stride] = s.x;
            H[(33 + L * L + m) * stride] = s.y;
            H[(49 + L * L + m) * stride] = s.z;
        }
    }
}

template <bool DERIV>
__global__ void __launch_bounds__(512, 2)
eval_ao_tiled_kernel(const double* __restrict__ coords, const ShellDev* __restrict__ shells,
                     const int* __restrict__ shell_atom, const int* __restrict__ atom_coord,
                     const AoMeta* __restrict__ meta, int nshell, int natm, int nrad, int lmax,
                     const double* __restrict__ env, int nenv, double* __restrict__ ao, int G, int Gpad,
                     int GpadMax, int N, int Npad, int C, int P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Hs[row][col], col = p*natm + atom, odd column stride (conflict-free both ways).
    // rows: 0..15 harmonics | 16 r^2 | (DERIV) 17..64 d/dx,d/dy,d/dz of the harmonics | 65..67 x,y,z
    constexpr int NROW = DERIV ? 68 : 17;
    constexpr int RS = DERIV ? 2 : 1;  // per (point, radial): R0 (+ R1)
    const int hstr = (P * natm) | 1;
    double* Hs = reinterpret_cast<double*>(smem_raw);
    double* Rs = Hs + (size_t)NROW * hstr;  // [P][nrad][RS]
    const int b = blockIdx.y;
    const long g0 = (long)blockIdx.x * P;
    const double* envb = env + (long)b * nenv;
    const long cstride = (long)GpadMax * Npad;
    const int tid = threadIdx.x;

    // ---- phase A1: harmonics and r^2 per (point, atom) ----
    for (int it = tid; it < P * natm; it += blockDim.x) {
        const int p = it / natm, ia = it - p * natm;
        const long gi = g0 + p;
        double x = 0, y = 0, z = 0;
        if (gi < G) {
            const double* r = coords + ((long)b * GpadMax + gi) * 3;
            const int ac = atom_coord[ia];
            x = r[0] - envb[ac];
            y = r[1] - envb[ac + 1];
            z = r[2] - envb[ac + 2];
        }
        double* H = Hs + it;
        fill_harm<0>(H, hstr, DERIV, x, y, z);
        fill_harm<1>(H, hstr, DERIV, x, y, z);
        if (lmax >= 2) fill_harm<2>(H, hstr, DERIV, x, y, z);
        if (lmax >= 3) fill_harm<3>(H, hstr, DERIV, x, y, z);
        H[16 * hstr] = x * x + y * y + z * z;
        if (DERIV) {
            H[65 * hstr] = x;
            H[66 * hstr] = y;
            H[67 * hstr] = z;
        }
    }
    __syncthreads();
    // ---- phase A2: radial sums; an item = one shell x two points (shell data loaded once) ----
    const int npair = (P + 1) >> 1;
    for (int it = tid; it < nshell * npair; it += blockDim.x) {
        const int pp = it / nshell, s = it - pp * nshell;
        const ShellDev sh = shells[s];
        const int ia = shell_atom[s];
        const double* ex = envb + sh.ptr_exp;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = 2 * pp + h;
            if (p >= P) break;
            const bool live = g0 + p < G;
            const double rr = Hs[16 * hstr + p * natm + ia];
            for (int ic = 0; ic < sh.nctr; ++ic) {
                const double* cf = envb + sh.ptr_coef + ic * sh.nprim;
                double R0 = 0.0, R1 = 0.0;
                for (int q = 0; q < sh.nprim; ++q) {
                    const double a = ex[q];
                    const double e = cf[q] * exp(-a * rr);
                    R0 += e;
                    if (DERIV) R1 -= 2.0 * a * e;
                }
                double* R = Rs + ((size_t)p * nrad + sh.rad_off + ic) * RS;
                R[0] = live ? R0 : 0.0;  // padding rows come out exactly zero
                if (DERIV) R[1] = live ? R1 : 0.0;
            }
        }
    }
    __syncthreads();
    // ---- phase B: a warp owns a 32-AO chunk for all P points; 256-byte coalesced row stores ----
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int nchunk = (Npad + 31) >> 5;
    int pmax = P;
    if (g0 + pmax > Gpad) pmax = (int)(Gpad - g0);
    for (int ch = warp; ch < nchunk; ch += nwarps) {
        const int n = (ch << 5) + lane;
        if (n >= Npad) continue;
        const AoMeta m = meta[n];
        const bool on = m.lm >= 0;
        const double* H = Hs + (on ? m.lm * hstr + m.atom : 0);
        const double* R = Rs + (on ? m.rad * RS : 0);
        double* row = ao + ((long)b * C * GpadMax + g0) * Npad + n;
        for (int p = 0; p < pmax; ++p, row += Npad) {
            double v0 = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
            if (on) {
                const double* Hp = H + p * natm;
                const double h = Hp[0], R0 = R[(size_t)p * nrad * RS];
                v0 = h * R0;
                if (DERIV) {
                    const double R1 = R[(size_t)p * nrad * RS + 1];
                    const double* Xp = Hs + p * natm + m.atom;
                    vx = Hp[17 * hstr] * R0 + h * Xp[65 * hstr] * R1;
                    vy = Hp[33 * hstr] * R0 + h * Xp[66 * hstr] * R1;
                    vz = Hp[49 * hstr] * R0 + h * Xp[67 * hstr] * R1;
                }
            }
            row[0] = v0;
            if (DERIV) {
                row[cstride] = vx;
                row[2 * cstride] = vy;
                row[3 * cstride] = vz;
            }
        }
    }
}

// user AO [B][ncomp][G][N] -> internal padded tensor (zero fill) and back
__global__ void pack_ao_kernel(const double* __restrict__ src, double* __restrict__ ao, int ncomp, int G,
                               int Gpad, int GpadMax, int N, int Npad, int C) {
    const int b = blockIdx.z, c = blockIdx.y;
    const long total = (long)Gpad * Npad;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const long gi = idx / Npad;
        const int n = (int)(idx - gi * Npad);
        double v = 0.0;
        if (gi < G && n < N) v = src[(((long)b * ncomp + c) * G + gi) * N + n];
        ao[(((long)b * C + c) * GpadMax + gi) * Npad + n] = v;
    }
}
__global__ void unpack_ao_kernel(double* __restrict__ dst, const double* __restrict__ ao, int ncomp, int G,
                                 int GpadMax, int N, int Npad, int C) {
    const int b = blockIdx.z, c = blockIdx.y;
    const long total = (long)G * N;
    for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long)gridDim.x * blockDim.x) {
        const long gi = idx / N;
        const int n = (int)(idx - gi * N);
        dst[(((long)b * ncomp + c) * G + gi) * N + n] = ao[(((long)b * C + c) * GpadMax + gi) * Npad + n];
    }
}

// shared: one source grid [G] for every batch element (weights are replicated, coordinates kept once)
__global__ void set_grid_kernel(const double* __restrict__ coords, const double* __restrict__ weights,
                                double* __restrict__ cdst, double* __restrict__ wdst, int G, int GpadMax, int shared) {
    const int b = blockIdx.y, bs = shared ? 0 : b;
    for (long gi = (long)blockIdx.x * blockDim.x + threadIdx.x; gi < GpadMax;
         gi += (long)gridDim.x * blockDim.x) {
        const bool in = gi < G;
        wdst[(long)b * GpadMax + gi] = in ? weights[(long)bs * G + gi] : 0.0;
        if (coords && (!shared || b == 0)) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                cdst[((long)b * GpadMax + gi) * 3 + k] = in ? coords[((long)b * G + gi) * 3 + k] : 0.0;
        }
    }
}

inline unsigned grid_for(long total, int threads, int num_sms) {
    long blocks = (total + threads - 1) / threads;
    const long cap = (long)num_sms * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace

int launch_set_grid(qexxc_ctx* c, const double* coords, const double* weights, int G, cudaStream_t st) {
    dim3 grid(grid_for(c->GpadMax, 256, c->num_sms), c->B);
    set_grid_kernel<<<grid, 256, 0, st>>>(coords, weights, c->coords, c->weights, G, c->GpadMax, c->ao_shared ? 1 : 0);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_eval_ao(qexxc_ctx* c, int deriv, cudaStream_t st) {
    ProfScope prof(c, QEXXC_PROF_EVAL_AO, st);
    // tiled kernel when the per-point staging fits in shared memory (it does up to ~1500 atoms)
    const size_t per_pt = ((size_t)c->natm * (deriv ? 68 : 17) + (size_t)c->nrad * (deriv ? 2 : 1)) * 8;
    int P = (int)((deriv ? 140 * 1024 : 100 * 1024) / (per_pt ? per_pt : 1));
    if (P > 16) P = 16;
    if (P >= 1) {
        const size_t smem = ((size_t)(deriv ? 68 : 17) * (((size_t)P * c->natm) | 1) +
                             (size_t)P * c->nrad * (deriv ? 2 : 1)) * 8;
        dim3 tgrid((unsigned)((c->Gpad + P - 1) / P), c->ao_shared ? 1 : c->B);
        // two resident CTAs share the SM's shared memory whatever the block size, so wide blocks double the
        // resident warps (the kernel is latency-bound, not pipe-bound); small molecules keep 256 threads
        int nthr = ((long)P * c->natm >= 256 || c->Npad >= 512) ? 512 : 256;
        if (getenv("QEXXC_AO_THREADS")) nthr = atoi(getenv("QEXXC_AO_THREADS"));
        if (deriv) {
            QX_CUDA(cudaFuncSetAttribute(eval_ao_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            eval_ao_tiled_kernel<true><<<tgrid, nthr, smem, st>>>(c->coords, c->shells, c->shell_atom, c->atom_coord,
                c->ao_meta, c->nshell, c->natm, c->nrad, c->lmax, c->env, c->nenv, c->ao, c->G, c->Gpad, c->GpadMax,
                c->N, c->Npad, c->C, P);
        } else {
            QX_CUDA(cudaFuncSetAttribute(eval_ao_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            eval_ao_tiled_kernel<false><<<tgrid, nthr, smem, st>>>(c->coords, c->shells, c->shell_atom, c->atom_coord,
                c->ao_meta, c->nshell, c->natm, c->nrad, c->lmax, c->env, c->nenv, c->ao, c->G, c->Gpad, c->GpadMax,
                c->N, c->Npad, c->C, P);
        }
        QX_LAUNCH_CHECK(c);
        return QEXXC_OK;
    }
    dim3 grid(grid_for((long)c->Gpad * 32, 256, c->num_sms), c->ao_shared ? 1 : c->B);
    eval_ao_kernel<<<grid, 256, 0, st>>>(c->coords, c->shells, c->nshell, c->env, c->nenv, c->ao, c->G,
                                         c->Gpad, c->GpadMax, c->N, c->Npad, c->C, deriv);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_pack_ao(qexxc_ctx* c, const double* src, int ncomp, int G, cudaStream_t st) {
    dim3 grid(grid_for((long)c->Gpad * c->Npad, 256, c->num_sms), ncomp, c->ao_shared ? 1 : c->B);
    pack_ao_kernel<<<grid, 256, 0, st>>>(src, c->ao, ncomp, G, c->Gpad, c->GpadMax, c->N, c->Npad, c->C);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

int launch_unpack_ao(qexxc_ctx* c, double* dst, int ncomp, cudaStream_t st) {
    dim3 grid(grid_for((long)c->G * c->N, 256, c->num_sms), ncomp, c->ao_shared ? 1 : c->B);
    unpack_ao_kernel<<<grid, 256, 0, st>>>(dst, c->ao, ncomp, c->G, c->GpadMax, c->N, c->Npad, c->C);
    QX_LAUNCH_CHECK(c);
    return QEXXC_OK;
}

}  // namespace qexxc

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
// CTA = P grid points.  Phase A: every (point, atom) gets its (lmax+1)^2 real solid harmonics (and their
// gradients), every (point, radial function) its contracted radial sum -- each exp is evaluated
// exactly once and staged in shared memory.  Phase B: warps sweep (point, 32-AO chunk) pairs;
// lane n multiplies harmonic x radial for AO n and the warp stores 256 contiguous bytes of the AO
// row, so the 8*N bytes per point stream to HBM fully coalesced.
//
// The kernel is co-limited by the FP64 pipe (one exp per primitive per point: 800 per point at c5) and the
// HBM write stream (8 KB per point), so the exponential is a table-driven one: exp(x) = 2^(k/64) * e^r with
// k = round(64 x / ln 2), |r| <= ln2/128, a degree-5 Taylor tail (remainder r^6/720 < 4e-17) and a
// 64-entry table of correctly rounded 2^(j/64) -- 9 FP64 instructions instead of libm's ~17, relative error
// <= 2.5e-16 -- and the radial sums of four points are interleaved so four exp chains are in flight per thread.
__device__ const double kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

// exp(x) for x <= 0; arguments below -700 are clamped (returns ~1e-304 where the true value is smaller still)
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ T) {
    const double L = 92.33248261689366, C1 = 0x1.62e42fe000000p-7, C2 = 0x1.f473de6af278fp-36;
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: the low word of x*L + MAGIC is round(x*L)
    const double xs = x < -700.0 ? -700.0 : x;
    double kd = fma(xs, L, MAGIC);
    const int k = __double2loint(kd);
    kd -= MAGIC;
    double r = fma(kd, -C1, xs);
    r = fma(kd, -C2, r);
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double y = T[k & 63] * p;
    return __hiloint2double(__double2hiint(y) + ((k >> 6) << 20), __double2loint(y));
}

template <int L>
__device__ __forceinline__ void fill_harm(double* __restrict__ H, int stride, int NH, bool deriv, double x, double y, double z) {
#pragma unroll
    for (int m = 0; m < 2 * L + 1; ++m) {
        const V4 s = solid<L>(m, x, y, z);
        H[(L * L + m) * stride] = s.v;
        if (deriv) {  // row blocks NH+1.., 2NH+1.., 3NH+1.. hold d/dx, d/dy, d/dz
            H[(NH + 1 + L * L + m) * stride] = s.x;
            H[(2 * NH + 1 + L * L + m) * stride] = s.y;
            H[(3 * NH + 1 + L * L + m) * stride] = s.z;
        }
    }
}

// rows of the harmonic table for NH = (lmax+1)^2 harmonics
__host__ __device__ inline int ao_tab_rows(int NH, bool deriv) { return deriv ? 4 * NH + 4 : NH + 1; }

// radial sums of one shell for every point of the CTA: the shell's primitives (exponent, coefficient) sit in
// registers (fast path: one contraction, <= 4 primitives -- every standard split-valence / polarisation shell), the
// points go four at a time so four exp chains are in flight
template <bool DERIV>
__device__ __forceinline__ void radial_phase(const ShellDev* __restrict__ shells, const int* __restrict__ shell_atom,
                                             const double* __restrict__ envb, const double* __restrict__ Hs,
                                             double* __restrict__ Rs, const double* __restrict__ T, int hstr, int NH,
                                             int nshell, int natm, int nrad, int P, long g0, int G) {
    constexpr int RS = DERIV ? 2 : 1;
    constexpr int Q = 4;
    const double* rrow = Hs + NH * hstr;
    // an item = one shell x `ppi` points: all P points when there are enough shells to occupy the CTA (the shell's
    // primitives are loaded once), one quad of points otherwise (small molecules: more items than threads matter more)
    const int ppi = (2 * nshell >= (int)blockDim.x) ? P : Q;
    const int ngrp = (P + ppi - 1) / ppi;
    for (int it = threadIdx.x; it < nshell * ngrp; it += blockDim.x) {
        const int grp = it / nshell, s = it - grp * nshell;
        const int pbeg = grp * ppi, pend = min(P, pbeg + ppi);
        const ShellDev sh = shells[s];
        const int ia = shell_atom[s];
        const double* ex = envb + sh.ptr_exp;
        if (sh.nctr == 1 && sh.nprim <= 4) {
            const double* cf = envb + sh.ptr_coef;
            double a[4], c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool in = q < sh.nprim;
                a[q] = in ? ex[q] : 0.0;
                c[q] = in ? cf[q] : 0.0;
            }
            double* R = Rs + (size_t)sh.rad_off * RS;
            for (int p0 = pbeg; p0 < pend; p0 += Q) {
                double rr[Q], R0[Q], R1[Q];
#pragma unroll
                for (int h = 0; h < Q; ++h) {
                    rr[h] = rrow[min(p0 + h, P - 1) * natm + ia];
                    R0[h] = R1[h] = 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < sh.nprim) {
#pragma unroll
                        for (int h = 0; h < Q; ++h) {
                            const double e = c[q] * exp_neg(-a[q] * rr[h], T);
                            R0[h] += e;
                            if (DERIV) R1[h] = fma(-2.0 * a[q], e, R1[h]);
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < Q; ++h) {
                    const int p = p0 + h;
                    if (p < P) {
                        const bool live = g0 + p < G;  // padding rows come out exactly zero
                        R[(size_t)p * nrad * RS] = live ? R0[h] : 0.0;
                        if (DERIV) R[(size_t)p * nrad * RS + 1] = live ? R1[h] : 0.0;
                    }
                }
            }
            continue;
        }
        // general shells: several contractions and/or many primitives
        for (int p0 = pbeg; p0 < pend; p0 += Q) {
            double rr[Q];
#pragma unroll
            for (int h = 0; h < Q; ++h) rr[h] = rrow[min(p0 + h, P - 1) * natm + ia];
            for (int ic = 0; ic < sh.nctr; ++ic) {
                const double* cf = envb + sh.ptr_coef + ic * sh.nprim;
                double R0[Q], R1[Q];
#pragma unroll
                for (int h = 0; h < Q; ++h) R0[h] = R1[h] = 0.0;
                for (int q = 0; q < sh.nprim; ++q) {
                    const double aq = ex[q], cq = cf[q];
#pragma unroll
                    for (int h = 0; h < Q; ++h) {
                        const double e = cq * exp_neg(-aq * rr[h], T);
                        R0[h] += e;
                        if (DERIV) R1[h] -= 2.0 * aq * e;
                    }
                }
#pragma unroll
                for (int h = 0; h < Q; ++h) {
                    const int p = p0 + h;
                    if (p < P) {
                        const bool live = g0 + p < G;
                        double* R = Rs + ((size_t)p * nrad + sh.rad_off + ic) * RS;
                        R[0] = live ? R0[h] : 0.0;
                        if (DERIV) R[1] = live ? R1[h] : 0.0;
                    }
                }
            }
        }
    }
}

template <bool DERIV, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 2)
eval_ao_tiled_kernel(const double* __restrict__ coords, const ShellDev* __restrict__ shells,
                     const int* __restrict__ shell_atom, const int* __restrict__ atom_coord,
                     const AoMeta* __restrict__ meta, int nshell, int natm, int nrad, int lmax,
                     const double* __restrict__ env, int nenv, double* __restrict__ ao, int G, int Gpad,
                     int GpadMax, int N, int Npad, int C, int P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Hs[row][col], col = p*natm + atom, odd column stride (conflict-free both ways).
    // rows: 0..NH-1 harmonics | NH r^2 | (DERIV) d/dx, d/dy, d/dz blocks of NH rows | x, y, z
    const int NH = (lmax + 1) * (lmax + 1);
    const int NROW = ao_tab_rows(NH, DERIV);
    constexpr int RS = DERIV ? 2 : 1;  // per (point, radial): R0 (+ R1)
    const int hstr = (P * natm) | 1;
    double* Hs = reinterpret_cast<double*>(smem_raw);
    double* Rs = Hs + (size_t)NROW * hstr;  // [P][nrad][RS]
    double* T = Rs + (size_t)P * nrad * RS;  // [64] 2^(j/64)
    const int b = blockIdx.y;
    const long g0 = (long)blockIdx.x * P;
    const double* envb = env + (long)b * nenv;
    const long cstride = (long)GpadMax * Npad;
    const int tid = threadIdx.x;
    if (tid < 64) T[tid] = kExp2Tab[tid];

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
        fill_harm<0>(H, hstr, NH, DERIV, x, y, z);
        if (lmax >= 1) fill_harm<1>(H, hstr, NH, DERIV, x, y, z);
        if (lmax >= 2) fill_harm<2>(H, hstr, NH, DERIV, x, y, z);
        if (lmax >= 3) fill_harm<3>(H, hstr, NH, DERIV, x, y, z);
        H[NH * hstr] = x * x + y * y + z * z;
        if (DERIV) {
            H[(4 * NH + 1) * hstr] = x;
            H[(4 * NH + 2) * hstr] = y;
            H[(4 * NH + 3) * hstr] = z;
        }
    }
    __syncthreads();
    // ---- phase A2: radial sums; a thread = one shell, all P points, four exp chains in flight ----
    radial_phase<DERIV>(shells, shell_atom, envb, Hs, Rs, T, hstr, NH, nshell, natm, nrad, P, g0, G);
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
        if (!DERIV) {
            // 32-bit table indices, one 64-bit row pointer: ~10 instructions per stored value
            int hi = 0, ri = 0, p = 0;
            for (; p + 4 <= pmax; p += 4) {  // four independent load pairs in flight
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    v[u] = H[hi] * R[ri];
                    hi += natm;
                    ri += nrad;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    __stcs(row, on ? v[u] : 0.0);
                    row += Npad;
                }
            }
            for (; p < pmax; ++p, row += Npad, hi += natm, ri += nrad) __stcs(row, on ? H[hi] * R[ri] : 0.0);
            continue;
        }
        for (int p = 0; p < pmax; ++p, row += Npad) {
            double v0 = 0.0, vx = 0.0, vy = 0.0, vz = 0.0;
            if (on) {
                const double* Hp = H + p * natm;
                const double h = Hp[0], R0 = R[(size_t)p * nrad * RS];
                v0 = h * R0;
                const double R1 = R[(size_t)p * nrad * RS + 1];
                const double* Xp = Hs + p * natm + m.atom;
                vx = Hp[(NH + 1) * hstr] * R0 + h * Xp[(4 * NH + 1) * hstr] * R1;
                vy = Hp[(2 * NH + 1) * hstr] * R0 + h * Xp[(4 * NH + 2) * hstr] * R1;
                vz = Hp[(3 * NH + 1) * hstr] * R0 + h * Xp[(4 * NH + 3) * hstr] * R1;
            }
            __stcs(row, v0);
            __stcs(row + cstride, vx);
            __stcs(row + 2 * cstride, vy);
            __stcs(row + 3 * cstride, vz);
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
    const int NH = (c->lmax + 1) * (c->lmax + 1);
    const int nrow = ao_tab_rows(NH, deriv != 0);
    const size_t per_pt = ((size_t)c->natm * nrow + (size_t)c->nrad * (deriv ? 2 : 1)) * 8;
    // two 512-thread CTAs per SM (the register file allows no more): up to ~110 KB of tables each
    int P = (int)((110 * 1024) / (per_pt ? per_pt : 1));
    if (P > 16) P = 16;
    if (P >= 4) P &= ~3;  // radial sums run four points at a time
    const bool reg85 = getenv("QEXXC_AO_REG85") && atoi(getenv("QEXXC_AO_REG85")) != 0;
    if (reg85) {  // three CTAs per SM: up to ~72 KB of tables each
        P = (int)((72 * 1024) / (per_pt ? per_pt : 1));
        if (P > 16) P = 16;
        if (P >= 4) P &= ~3;
    }
    if (getenv("QEXXC_AO_P")) P = atoi(getenv("QEXXC_AO_P"));
    if (P >= 1) {
        const size_t smem = ((size_t)nrow * (((size_t)P * c->natm) | 1) + (size_t)P * c->nrad * (deriv ? 2 : 1) + 64) * 8;
        dim3 tgrid((unsigned)((c->Gpad + P - 1) / P), c->ao_shared ? 1 : c->B);
        // two resident CTAs share the SM's shared memory whatever the block size, so wide blocks double the
        // resident warps (the kernel is latency-bound, not pipe-bound); small molecules keep 256 threads
        int nthr = ((long)P * c->natm >= 256 || c->Npad >= 512) ? 512 : 256;
        if (getenv("QEXXC_AO_THREADS")) nthr = atoi(getenv("QEXXC_AO_THREADS"));
        if (reg85) nthr = 256;
#define QX_AO(D, T)                                                                                              \
    do {                                                                                                         \
        QX_CUDA(cudaFuncSetAttribute(eval_ao_tiled_kernel<D, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        eval_ao_tiled_kernel<D, T><<<tgrid, nthr, smem, st>>>(c->coords, c->shells, c->shell_atom, c->atom_coord, c->ao_meta, \
            c->nshell, c->natm, c->nrad, c->lmax, c->env, c->nenv, c->ao, c->G, c->Gpad, c->GpadMax, c->N, c->Npad, c->C, P); \
    } while (0)
        if (reg85) {  // 256-thread CTAs, three per SM, up to 85 registers per thread
            if (deriv) QX_AO(true, 256);
            else QX_AO(false, 256);
        } else {
            if (deriv) QX_AO(true, 512);
            else QX_AO(false, 512);
        }
#undef QX_AO
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

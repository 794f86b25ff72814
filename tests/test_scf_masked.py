"""Padded / masked batched SCF (second half): molecules of DIFFERENT nao and grid size in one
fixed-shape batch.  Mirrors the reference's own checks (scf_functions_masked.py:596-821 `test_*_masked`,
`compare_padded_vs_non_padded` :970-1113; generalized_eigensolver_masked.py:91-170): the masked pieces on padded
inputs must reproduce the unpadded pieces, whatever sits in the padding."""
import numpy as np
import pytest

from oracle import gto_ref, ints_ref, mlp_ref, scf_ref
from qex_b200 import gen_grid, gto


def _problem(R, basis, n_rad):
    m = gto.h2(R, basis)
    I = ints_ref.integrals(m._atm, m._bas, m._env)
    g = gen_grid.Grids(m, n_rad=n_rad, n_theta=5, n_phi=4).build()
    return m, I, g


def _pad(a, n, fill=0.0):
    out = np.full((n,) * a.ndim, fill)
    out[tuple(slice(0, k) for k in a.shape)] = a
    return out


def test_oracle_masked_pieces_reproduce_the_unpadded_ones():
    rng = np.random.default_rng(0)
    n, npad = 3, 5
    f = rng.standard_normal((n, n)); f = f @ f.T
    s = rng.standard_normal((n, n)); s = s @ s.T + 5.0 * np.eye(n)
    mask = np.arange(npad) < n
    w_ref, v_ref = scf_ref.generalized_eigh(f, s)
    for fill in (0.0, 100.0):  # the reference's test pads with zeros and with 100s
        fp, sp = _pad(f, npad, fill), _pad(s, npad, fill)
        fp[:n, :n], sp[:n, :n] = f, s
        w, v = scf_ref.masked_generalized_eigh(fp, sp, mask)
        assert np.abs(w[:n] - w_ref).max() < 1e-12 and np.all(w[n:] == 0.0)
        assert np.abs(np.abs(v[:n, :n]) - np.abs(v_ref)).max() < 1e-10 and np.all(v[n:] == 0) and np.all(v[:, n:] == 0)
    occ = scf_ref.get_occ_masked(2, np.array([0.3, -0.5, 1.0, -9.0, -9.0]), mask)
    assert np.array_equal(occ, [0.0, 2.0, 0.0, 0.0, 0.0])  # padded "eigenvalues" are never occupied
    c = rng.standard_normal((npad, npad))
    dm = scf_ref.make_rdm1_masked(c, np.array([2.0, 0, 0, 2.0, 2.0]), mask)
    assert np.abs(dm[:n, :n] - scf_ref.make_rdm1(c[:n, :n], np.array([2.0, 0, 0]))).max() < 1e-14 and np.all(dm[n:] == 0)


def test_host_masked_pieces_match_the_oracle_on_cpu():
    import torch

    from qex_b200 import scf

    rng = np.random.default_rng(1)
    sizes, npad = [3, 5, 2], 5
    fs, ss, masks = [], [], []
    for n in sizes:
        f = rng.standard_normal((n, n)); f = f @ f.T
        s = rng.standard_normal((n, n)); s = s @ s.T + 5.0 * np.eye(n)
        fp, sp = _pad(f, npad, 7.0), _pad(s, npad, 7.0)
        fp[:n, :n], sp[:n, :n] = f, s
        fs.append(fp); ss.append(sp); masks.append(np.arange(npad) < n)
    F, S, M = (torch.as_tensor(np.stack(x)) for x in (fs, ss, masks))
    w, v = scf.masked_generalized_eigh(F, S, M)
    ne = torch.as_tensor([2, 4, 2])
    occ = scf.get_occ_masked(ne, w, M)
    dm = scf.make_rdm1_masked(v, occ, M)
    for b, n in enumerate(sizes):
        w_ref, v_ref = scf_ref.masked_generalized_eigh(fs[b], ss[b], masks[b])
        assert np.abs(w[b].numpy() - w_ref).max() < 1e-11
        occ_ref = scf_ref.get_occ_masked(int(ne[b]), w_ref, masks[b])
        assert np.array_equal(occ[b].numpy(), occ_ref)
        dm_ref = scf_ref.make_rdm1_masked(v_ref, occ_ref, masks[b])
        assert np.abs(dm[b].numpy() - dm_ref).max() < 1e-10
    stack, mask2 = scf.pad_stack([torch.ones(2, 2), torch.ones(4, 4)])
    assert stack.shape == (2, 4, 4) and mask2.tolist() == [[True, True, False, False], [True] * 4]
    assert float(stack[0].sum()) == 4.0


def test_oracle_padded_loop_equals_unpadded_loop():
    """`compare_padded_vs_non_padded`.  With exact-zero padding the padded loop is the same arithmetic as the plain
    one.  The reference's "stable" get_veff (the definition in force, scf_functions_masked.py:546-588) pads with
    eps = 1e-12 and adds eps to rho at EVERY grid point; the level-0 grid reaches out to ~100 Bohr, so sum(w) ~ 1e6 and
    E_xc moves by ~1e-6 Ha -- a property of the reference's prototype, reproduced by the oracle (eps = 1e-12), not by
    the product path (which must equal the unpadded loop)."""
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    fxc = lambda rho: mlp_ref.exc_and_vrho_local(spec, theta, rho)  # noqa: E731
    m, I, g = _problem(0.74, "sto-3g", 15)
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
    n, npad = 2, 4
    mask = np.arange(npad) < n
    dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
    kw = dict(max_cycle=6, diis_start_cycle=10**6)
    e_ref, dm_ref, hist_ref = scf_ref.scf_loop(dm0, I["eri"], ao, g.weights, I["s1e"], I["h1e"], I["enuc"], 2, fxc, **kw)
    aop = np.zeros((g.size, npad)); aop[:, :n] = ao
    e, dm, hist = scf_ref.scf_loop_padded(_pad(dm0, npad), _pad(I["eri"], npad), aop, g.weights, _pad(I["s1e"], npad),
                                          _pad(I["h1e"], npad), I["enuc"], 2, mask, fxc, **kw)
    assert np.abs(hist - hist_ref).max() < 1e-5 and np.abs(dm[:n, :n] - dm_ref).max() < 1e-5
    e0, _, hist0 = scf_ref.scf_loop_padded(_pad(dm0, npad), _pad(I["eri"], npad), aop, g.weights, _pad(I["s1e"], npad),
                                           _pad(I["h1e"], npad), I["enuc"], 2, mask, fxc, eps=0.0, **kw)
    assert np.abs(hist0 - hist_ref).max() < 1e-12  # with exact-zero padding the two loops are the same arithmetic


@pytest.mark.gpu
def test_cuda_padded_batch_of_mixed_sizes_matches_unpadded_loops():
    """Three H2 molecules with nao = 4 / 2 / 4 and three different grid sizes in ONE padded batch (batched XC launches,
    batched J kernel, masked batched eigensolver) against three unpadded single-molecule GPU loops (1e-10 Ha) and the
    oracle's padded loop."""
    import torch

    from qex_b200 import _lib, scf
    from qex_b200.engine import NetSpec, XCContext

    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 3))
    fxc = lambda rho: mlp_ref.exc_and_vrho_local(spec, theta, rho)  # noqa: E731
    probs = [_problem(0.74, "6-31g", 31), _problem(0.9, "sto-3g", 20), _problem(1.5, "6-31g", 25)]
    B, Nmax = len(probs), 4
    Gmax = max(g.size for _, _, g in probs)
    assert len({g.size for _, _, g in probs}) == 3
    net = NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=1, n_hidden=3, width=64)
    th = torch.as_tensor(theta).cuda()
    kw = dict(max_cycle=8, diis_start_cycle=10**6)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()  # noqa: E731
    # unpadded single-molecule loops; their AO values (from the AO kernel) also fill the padded batch
    singles, aos = [], []
    for m, I, g in probs:
        n = m.nao_nr()
        x1 = XCContext(nao=n, ngrids_max=g.size, ncomp=1, net=net)
        x1.set_grid(g.coords, g.weights).set_basis(m._atm, m._bas, m._env).eval_ao(0)
        dm0 = scf_ref.core_guess(I["h1e"], I["s1e"], 2)
        _, dm1, hist1 = scf.scf_loop(x1, th, dev(dm0), dev(I["eri"]), dev(I["s1e"]), dev(I["h1e"]), I["enuc"], 2, **kw)
        singles.append((dm0, dm1.cpu().numpy(), hist1.cpu().numpy()))
        aos.append(x1.get_ao(1).cpu().numpy()[0, 0])
        x1.close()
    ao_pad = np.zeros((B, 1, Gmax, Nmax))
    w_pad = np.zeros((B, Gmax))
    for b, ((m, I, g), ao) in enumerate(zip(probs, aos)):
        ao_pad[b, 0, : g.size, : ao.shape[1]] = ao
        w_pad[b, : g.size] = g.weights
    xb = XCContext(nao=Nmax, ngrids_max=Gmax, ncomp=1, nbatch=B, net=net)
    xb.set_grid(None, w_pad).set_ao(ao_pad, 1)
    stack = lambda k: scf.pad_stack([dev(I[k]) for _, I, _ in probs], Nmax)[0]  # noqa: E731
    dm0s, mask = scf.pad_stack([dev(s[0]) for s in singles], Nmax)
    assert mask.sum(-1).tolist() == [4, 2, 4]
    enuc = dev(np.array([I["enuc"] for _, I, _ in probs]))
    e_b, dm_b, hist_b = scf.scf_loop_padded(xb, th, dm0s, stack("eri"), stack("s1e"), stack("h1e"), enuc, 2, mask, **kw)
    for b, ((m, I, g), (dm0, dm1, hist1)) in enumerate(zip(probs, singles)):
        n = m.nao_nr()
        assert np.abs(hist_b[:, b].cpu().numpy() - hist1).max() < 1e-10          # the VERDICT's bar: 1e-10 Ha
        got = dm_b[b].cpu().numpy()
        assert np.abs(got[:n, :n] - dm1).max() < 1e-9 and np.all(got[n:] == 0) and np.all(got[:, n:] == 0)
        aop = np.zeros((g.size, Nmax)); aop[:, :n] = gto_ref.eval_ao(m._atm, m._bas, m._env, g.coords, 0)
        mk = np.arange(Nmax) < n
        _, _, hist_o = scf_ref.scf_loop_padded(_pad(dm0, Nmax), _pad(I["eri"], Nmax), aop, g.weights, _pad(I["s1e"], Nmax),
                                               _pad(I["h1e"], Nmax), I["enuc"], 2, mk, fxc, eps=0.0, **kw)
        assert np.abs(hist_b[:, b].cpu().numpy() - hist_o).max() < 1e-9        # oracle padded loop, exact-zero padding
        _, _, hist_e = scf_ref.scf_loop_padded(_pad(dm0, Nmax), _pad(I["eri"], Nmax), aop, g.weights, _pad(I["s1e"], Nmax),
                                               _pad(I["h1e"], Nmax), I["enuc"], 2, mk, fxc, **kw)
        # the reference's eps = 1e-12 variant: rho + eps over a grid whose weights sum to ~1e6-1e7 Bohr^3 (see the CPU test)
        assert np.abs(hist_b[:, b].cpu().numpy() - hist_e).max() < 1e-3
    # reference defaults (DIIS on): finite, right shape; gradient w.r.t. theta flows through the padded loop
    th2 = th.clone().requires_grad_(True)
    e_d, _, hist_d = scf.scf_loop_padded(xb, th2, dm0s, stack("eri"), stack("s1e"), stack("h1e"), enuc, 2, mask, max_cycle=4)
    assert torch.isfinite(hist_d).all() and hist_d.shape == (4, B)
    (g_th,) = torch.autograd.grad(e_d.sum(), th2)
    assert torch.isfinite(g_th).all() and float(g_th.abs().max()) > 0
    xb.close()

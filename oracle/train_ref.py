"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the training step that calls the hot path.

Restates ``TDKSDFTTrainer`` of qedft/train/td/trainer_legacy_no_jit.py for two-electron molecules:

* ``make_dataset``     -- ``prepare_dataset`` :169-235 + ``DataGenerator.generate_data`` (dataset_generation.py:111-160,
  :360-395): target energy and AO density matrix from CCSD (exact for two electrons -> a full CI of the oracle
  integrals), the level-0 Stratmann grid, target density ``eval_rho(eval_ao(coords), dm_ao)``;
* ``batch_loss``       -- ``batch_loss_fn`` :240-283: per molecule a KS-SCF with the learned functional, then
  ``energy_weight (E - E_goal)^2`` and ``density_weight mean((rho - rho_true)^2)``; batch means are added;
* ``adam_init/adam_update`` -- ``optax.adam(learning_rate)`` :427-428 (b1 0.9, b2 0.999, eps 1e-8, bias-corrected).

The SCF inside is the reference's fixed-cycle form (``_scf_test_non_padded``, scf_functions_masked.py:856-904, =
``scf_ref.scf_loop``) started from the core guess; the trainer's ``mf.kernel`` is pyscfad's driver (minao guess,
convergence test) whose converged result is the same fixed point.  Pin status: the target energies reproduce the
notebook's E(CCSD) values and the grid/AO/rho chain its LDA energy (tests/test_scf.py, tests/test_zz_pyscf_pin.py);
the loss itself has no reference-held number ("parity unpinned": JAX's PRNG stream initialises the network).
"""
import numpy as np

from . import grid_ref, gto_ref, ints_ref, mlp_ref, scf_ref


def full_ci_two_electron(I):
    """-> (e_tot, dm_ao) of the singlet ground state of a two-electron molecule (CCSD == FCI there)."""
    _, C = scf_ref.generalized_eigh(I["h1e"], I["s1e"])
    h = C.T @ I["h1e"] @ C
    e = np.einsum("pi,qj,rk,sl,pqrs->ijkl", C, C, C, C, I["eri"], optimize=True)
    n = h.shape[0]
    eye = np.eye(n)
    H = (np.einsum("ik,jl->ijkl", h, eye) + np.einsum("ik,jl->ijkl", eye, h) + e.transpose(0, 2, 1, 3)).reshape(n * n, n * n)
    ev, vec = np.linalg.eigh(H)
    c = vec[:, 0].reshape(n, n)
    return float(ev[0] + I["enuc"]), C @ (2.0 * c @ c.T) @ C.T


def make_dataset(mols, level=0):
    """[(E_goal, density_goal [G,4] = (x, y, z, rho), mol, extras)] ; extras: integrals, grid weights, AO values."""
    out = []
    for m in mols:
        I = ints_ref.integrals(m._atm, m._bas, m._env)
        e, dm_ao = full_ci_two_electron(I)
        coords, weights = grid_ref.build(m.atom_charges(), m.atom_coords(), level=level, becke_scheme=grid_ref.stratmann)
        ao = gto_ref.eval_ao(m._atm, m._bas, m._env, coords, 0)
        rho = np.einsum("gi,ij,gj->g", ao, dm_ao, ao)
        out.append((e, np.concatenate([coords, rho[:, None]], axis=1), m, dict(I=I, weights=weights, ao=ao, dm_ao=dm_ao)))
    return out


def _veff(dm, eri, ao, w, exc_vrho, is_global):
    J = np.einsum("ijkl,kl->ij", eri, dm)
    rho = np.einsum("gi,ij,gj->g", ao, dm, ao)
    exc, vrho = exc_vrho(rho)
    Vxc = np.einsum("gi,g,gj->ij", ao, w * vrho, ao)
    # "NN-AmplitudeEncoding" adds the network's scalar, unweighted (numint_legacy.py:331); "NN": sum w rho exc (:306)
    exc_e = float(exc) if is_global else float(np.sum(exc * rho * w))
    return J + Vxc, exc_e, J


def mlp_functional(spec, is_global=False):
    """-> functional(theta, rho) = (exc, vrho) of trainer_legacy_no_jit.py:46-63 for an MLP spec."""
    if is_global:
        return lambda theta, rho: mlp_ref.exc_and_vrho_global(spec, theta, rho)
    return lambda theta, rho: mlp_ref.exc_and_vrho_local(spec, theta, rho)


def ks_scf(theta, functional, I, ao, w, nelectron=2, is_global=False, max_cycle=15, diis=True):
    """-> (e_tot, dm) of the fixed-cycle KS loop; ``functional(theta, rho) -> (exc, vrho)``."""
    def exc_vrho(rho):
        return functional(theta, rho)

    dm = scf_ref.core_guess(I["h1e"], I["s1e"], nelectron)
    vhf, exc_e, J = _veff(dm, I["eri"], ao, w, exc_vrho, is_global)
    st = scf_ref.initialize_diis(15)
    e_tot = None
    for cycle in range(max_cycle):
        fock = I["h1e"] + vhf
        if diis and cycle >= 1:
            fock, st = scf_ref.apply_diis(st, fock, dm, I["s1e"], 15, 2, 0.0)
        mo_e, mo_c = scf_ref.generalized_eigh(fock, I["s1e"])
        dm = scf_ref.make_rdm1(mo_c, scf_ref.get_occ(nelectron, mo_e))
        vhf, exc_e, J = _veff(dm, I["eri"], ao, w, exc_vrho, is_global)
        e_tot = scf_ref.energy_tot(dm, I["h1e"], J, exc_e, I["enuc"])
    return e_tot, dm


def batch_loss(theta, functional, batch, energy_weight=1.0, density_weight=1.0, is_global=False, max_cycle=15, diis=True):
    """trainer_legacy_no_jit.py:240-283 -> scalar loss of one batch of ``make_dataset`` entries.
    ``functional``: ``mlp_functional(spec, is_global)`` or any ``(theta, rho) -> (exc, vrho)``; an ``MLPSpec`` is accepted."""
    if isinstance(functional, mlp_ref.MLPSpec):
        functional = mlp_functional(functional, is_global)
    le, ln = [], []
    for e_goal, density_goal, _mol, x in batch:
        e, dm = ks_scf(theta, functional, x["I"], x["ao"], x["weights"], 2, is_global, max_cycle, diis)
        le.append(energy_weight * (e - e_goal) ** 2)
        rho = np.einsum("gi,ij,gj->g", x["ao"], dm, x["ao"])
        ln.append(density_weight * np.mean((rho - density_goal[:, 3]) ** 2))
    return float(np.mean(le) + np.mean(ln))


def adam_init(theta):
    return dict(count=0, mu=np.zeros_like(theta), nu=np.zeros_like(theta))


def adam_update(grads, state, theta, learning_rate=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam + optax.apply_updates."""
    count = state["count"] + 1
    mu = b1 * state["mu"] + (1 - b1) * grads
    nu = b2 * state["nu"] + (1 - b2) * grads**2
    mu_hat = mu / (1 - b1**count)
    nu_hat = nu / (1 - b2**count)
    return theta - learning_rate * mu_hat / (np.sqrt(nu_hat) + eps), dict(count=count, mu=mu, nu=nu)

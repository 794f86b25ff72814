"""The training step around the hot path: device-resident mirror of ``TDKSDFTTrainer``
(qedft/train/td/trainer_legacy_no_jit.py:110-560), the caller `north_star` names ("KSDFTTrainer still trains end to end").

Per loss evaluation the reference builds, for every molecule of the batch, a pyscf grid, an RKS object with the
learned functional, runs ``mf.kernel(params)`` and adds ``energy_weight (E - E_goal)^2`` and
``density_weight mean((rho - rho_true)^2)`` (:240-283), then differentiates the batch loss with
``jax.value_and_grad`` and steps ``optax.adam`` (:427-428, :470-500).  Here the whole batch is ONE device-resident
problem: a batched ``XCContext`` (grid, AO values evaluated once per geometry instead of once per ``nr_rks`` call), the
batched fixed-cycle KS loop ``scf.scf_loop_batched`` (XC kernels + batched J kernel + small-matrix eigensolver), the
density loss through the differentiable ``eval_rho``, reverse mode through the kernels' own VJPs, and an Adam update
on the flat parameter vector.  torch is the tape and the N x N plumbing; the grid-sized arithmetic is libqexxc's.

Kept from the reference: config keys (``train_bond_lengths``, ``val_bond_lengths``, ``basis``, ``method``,
``grid_density``, ``n_iterations``, ``batch_size``, ``is_global_xc``, ``learning_rate``, ``energy_weight``,
``density_weight``, ``max_cycle``, ``validation_interval``; extra: ``cuda_graph`` replays each iteration as one CUDA graph),
the dataset entries ``(energy, density_true [G,4], mol)``,
``_compute_loss_and_grad(params, batch, ew, dw) -> (loss, grads)``, ``_compute_validation_loss``, ``train() ->
(params, opt_state, train_losses, val_losses)``.  Different on purpose: the SCF is the reference's fixed-cycle form
(``_scf_test_non_padded``) from the core guess, not pyscfad's driver; data generation covers two-electron molecules
with s-type bases (targets by full CI == CCSD there), because the integrals come from ``qex_b200.ints`` instead of
pyscf -- a pyscf-generated dataset can be passed to ``train(training_data=...)`` unchanged if the integrals are supplied.
No checkpoints, plots or logging (out of scope).
"""
from __future__ import annotations

import numpy as np
import torch

from . import autograd as _ag
from . import dist as _dist
from . import gen_grid, gto, ints, scf
from .engine import XCContext
from .networks import GlobalMLP, LocalMLP
from .xc import _native_apply, _theta


def full_ci_two_electron(I):
    """CCSD target of dataset_generation.py:360-375 for two electrons (where CCSD is exact): -> (e_tot, dm_ao)."""
    h1e, s1e, eri = (torch.as_tensor(I[k], dtype=torch.float64) for k in ("h1e", "s1e", "eri"))
    L = torch.linalg.cholesky(s1e)
    X = torch.linalg.inv(L).T  # X^T S X = 1
    h = X.T @ h1e @ X
    e = torch.einsum("pi,qj,rk,sl,pqrs->ijkl", X, X, X, X, eri)
    n = h.shape[0]
    eye = torch.eye(n, dtype=torch.float64)
    H = (torch.einsum("ik,jl->ijkl", h, eye) + torch.einsum("ik,jl->ijkl", eye, h) + e.permute(0, 2, 1, 3)).reshape(n * n, n * n)
    ev, vec = torch.linalg.eigh(H)
    c = vec[:, 0].reshape(n, n)
    return float(ev[0]) + float(I["enuc"]), (X @ (2.0 * c @ c.T) @ X.T).numpy()


def adam_init(theta: torch.Tensor) -> dict:
    return dict(count=0, mu=torch.zeros_like(theta), nu=torch.zeros_like(theta))


def adam_update(grads, state, theta, learning_rate=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """optax.adam(learning_rate) followed by optax.apply_updates."""
    count = state["count"] + 1
    mu = b1 * state["mu"] + (1 - b1) * grads
    nu = b2 * state["nu"] + (1 - b2) * grads * grads
    step = learning_rate * (mu / (1 - b1**count)) / (torch.sqrt(nu / (1 - b2**count)) + eps)
    return theta - step, dict(count=count, mu=mu, nu=nu)


class _BatchProblem:
    """Everything of one batch that does not depend on theta, resident on the device."""

    def __init__(self, entries, spec, device, grid_density):
        mols = [e[2] for e in entries]
        self.nmol = len(mols)
        nao = {m.nao_nr() for m in mols}
        if len(nao) != 1:
            raise NotImplementedError("one batch holds molecules of equal nao (pad or split the batch)")
        self.nao = nao.pop()
        self.nelectron = {int(sum(m.atom_charges())) for m in mols}
        if len(self.nelectron) != 1:
            raise NotImplementedError("one batch holds molecules with the same electron count")
        self.nelectron = self.nelectron.pop()
        for m in mols[1:]:  # one basis table serves the whole batch (set_basis below): only the geometry (env) may differ
            if not (np.array_equal(m._atm, mols[0]._atm) and np.array_equal(m._bas, mols[0]._bas)):
                raise NotImplementedError("one batch holds molecules with identical _atm / _bas tables (same atoms, same basis)")
        grids = []
        for e in entries:
            g = gen_grid.Grids(e[2])
            g.level = grid_density
            g.becke_scheme = gen_grid.stratmann
            grids.append(g.build(device=device))
        G = {g.size for g in grids}
        if len(G) != 1:
            raise NotImplementedError("one batch holds molecules with equal grid size")
        self.ngrids = G.pop()
        self.xc = XCContext(nao=self.nao, ngrids_max=self.ngrids, ncomp=1, nbatch=self.nmol, net=spec, device=device)
        self.xc.set_grid(np.stack([g.coords for g in grids]), np.stack([g.weights for g in grids]))
        self.xc.set_basis(mols[0]._atm, mols[0]._bas, np.stack([m._env for m in mols])).eval_ao(0)
        dev = self.xc.tdev
        I = [e[3]["I"] if len(e) > 3 and "I" in e[3] else ints.integrals(e[2]._atm, e[2]._bas, e[2]._env) for e in entries]
        st = lambda k: torch.as_tensor(np.stack([np.asarray(x[k]) for x in I]), dtype=torch.float64, device=dev)  # noqa: E731
        self.eri, self.s1e, self.h1e = st("eri"), st("s1e"), st("h1e")
        self.enuc = torch.as_tensor(np.array([float(x["enuc"]) for x in I]), dtype=torch.float64, device=dev)
        self.e_goal = torch.as_tensor(np.array([float(e[0]) for e in entries]), dtype=torch.float64, device=dev)
        for e, g in zip(entries, grids):  # the density loss reuses the context's AO values: same points required
            if e[1].shape[0] != g.size or np.abs(np.asarray(e[1])[:, :3] - g.coords).max() > 1e-10:
                raise NotImplementedError("density targets must live on the molecule's own level-%d grid" % grid_density)
        self.rho_goal = torch.as_tensor(np.stack([np.asarray(e[1])[:, 3] for e in entries]), dtype=torch.float64, device=dev)
        with torch.no_grad():
            w, c = scf.generalized_eigh_batched(self.h1e, self.s1e)
            self.dm0 = scf.make_rdm1(c, scf.get_occ_batched(self.nelectron, w))


class _GraphedIteration:
    """One training iteration of one batch -- forward SCF, losses, reverse pass, Adam -- captured as ONE CUDA graph.
    Nothing inside synchronises with the host (batched eigensolver kernel, `solve_ex` without `info`, Adam's step
    counter on the device), so a replay is a single launch: 10.5 ms against 37.6 ms of stream launches for the README
    example (profiles/r01/train_c1_local.json).  All iterations share the optimiser buffers in `shared`."""

    def __init__(self, trainer, batch_data, ew, dw, lr, shared, b1=0.9, b2=0.999, eps=1e-8):
        self.shared = shared
        self.loss = torch.zeros((), dtype=torch.float64, device=shared["theta"].device)
        n = len(batch_data)

        def iteration():
            th = shared["theta"].detach().requires_grad_(True)
            loss = trainer._loss_sum(th, batch_data, ew, dw) / n
            (g,) = torch.autograd.grad(loss, th)
            shared["count"].add_(1.0)
            shared["mu"].mul_(b1).add_(g, alpha=1 - b1)
            shared["nu"].mul_(b2).addcmul_(g, g, value=1 - b2)
            c = shared["count"]
            mu_hat = shared["mu"] / (1.0 - torch.pow(torch.full_like(c, b1), c))
            nu_hat = shared["nu"] / (1.0 - torch.pow(torch.full_like(c, b2), c))
            shared["theta"].sub_(lr * mu_hat / (torch.sqrt(nu_hat) + eps))
            self.loss.copy_(loss.detach())

        keep = {k: v.clone() for k, v in shared.items()}  # warm-up iterations must not move the optimiser
        side = torch.cuda.Stream(device=shared["theta"].device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                iteration()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            iteration()
        for k, v in keep.items():
            shared[k].copy_(v)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss.clone()


class TDKSDFTTrainer:
    def __init__(self, config_dict: dict, network=None, seed: int = 0, device: int | None = None):
        self.config = dict(config_dict)
        self.seed = seed
        self.device = torch.cuda.current_device() if device is None and torch.cuda.is_available() else device
        self.is_global_xc = bool(self.config.get("is_global_xc", True))
        self.network = network  # (init_fn, apply_fn); None: built for the first molecule's grid in train()
        self._problems = {}

    # ---- data (prepare_dataset :169-235) ----------------------------------------------------------------
    def _generate(self, bond_length):
        basis = {"631g": "6-31g"}.get(str(self.config.get("basis", "631g")).lower(), self.config.get("basis"))
        method = str(self.config.get("method", "CCSD")).lower()
        if method not in ("ccsd", "fci"):
            raise ValueError(f"  Method {method} for data generation is not implemented (e.g., ccsd, fci).")
        mol = gto.h2(float(bond_length), basis)
        I = ints.integrals(mol._atm, mol._bas, mol._env)
        energy, dm_ao = full_ci_two_electron(I)
        g = gen_grid.Grids(mol)
        g.level = int(self.config.get("grid_density", 0))
        g.becke_scheme = gen_grid.stratmann
        g.build(device=self.device)
        from . import numint

        ao = numint.eval_ao(mol, g.coords, deriv=0)
        rho = numint.eval_rho(mol, ao, dm_ao, xctype="LDA")
        return (energy, np.concatenate([g.coords, rho[:, None]], axis=1), mol, dict(I=I, dm_ao=dm_ao))

    def prepare_dataset(self):
        train = [self._generate(b) for b in self.config.get("train_bond_lengths", [0.74, 0.5, 1.5])]
        val = [self._generate(b) for b in self.config.get("val_bond_lengths", [0.6, 0.9, 1.2])]
        return train, val

    # ---- loss (:237-285) -----------------------------------------------------------------------------------
    def _spec(self):
        return _native_apply(self.network).qex_spec

    def _problem(self, batch_data) -> _BatchProblem:
        # keyed on CONTENT (geometry / basis tables and the targets), not on object identity: the same Mole objects reused
        # with other targets, or ids recycled after garbage collection, must not hit a stale entry; bounded (LRU)
        import hashlib

        h = hashlib.sha1()
        for e in batch_data:
            for part in (e[0], e[1], e[2]._atm, e[2]._bas, e[2]._env):
                h.update(np.ascontiguousarray(np.asarray(part, dtype=np.float64)).tobytes())
        key = h.hexdigest()
        if key in self._problems:
            self._problems[key] = self._problems.pop(key)  # most recently used last
        else:
            while len(self._problems) >= 32:
                self._problems.pop(next(iter(self._problems)))
            self._problems[key] = _BatchProblem(batch_data, self._spec(), self.device, int(self.config.get("grid_density", 0)))
        return self._problems[key]

    def _loss_sum(self, theta, batch_data, energy_weight, density_weight):
        """Sum over the given molecules of the two per-molecule loss terms (a differentiable 0-d tensor)."""
        p = self._problem(batch_data)
        xctype = "NN-AmplitudeEncoding" if self.is_global_xc else "NN"
        e, dm, _ = scf.scf_loop_batched(p.xc, theta, p.dm0, p.eri, p.s1e, p.h1e, p.enuc, p.nelectron, xctype=xctype,
                                        max_cycle=int(self.config.get("max_cycle", 20)),
                                        diis_start_cycle=int(self.config.get("diis_start_cycle", 1)))
        loss_e = energy_weight * (e - p.e_goal) ** 2
        rho = _ag.eval_rho(p.xc, dm, 1, 1)[:, 0, :]  # [B, G]
        loss_n = density_weight * ((rho - p.rho_goal) ** 2).mean(dim=1)
        return loss_e.sum() + loss_n.sum()

    def _loss_and_grad_flat(self, theta, batch_data, energy_weight, density_weight, want_grad=True):
        """mean_E + mean_n of the batch (:277-281) and its theta gradient.  Under torch.distributed the molecules of
        the batch are dealt round-robin to the ranks (`dist.shard_batch`: replicas, no data-path collective) and the
        packed `[grad | loss]` is all-reduced once, so every rank steps the same replicated theta."""
        rank, world = _dist.rank_world()
        mine = [batch_data[i] for i in _dist.shard_batch(len(batch_data), rank, world)]
        theta = theta.detach().requires_grad_(want_grad)
        packed = torch.zeros(theta.numel() + 1, dtype=torch.float64, device=theta.device)
        if mine:
            with torch.set_grad_enabled(want_grad):
                loss = self._loss_sum(theta, mine, energy_weight, density_weight) / len(batch_data)
            if want_grad:
                (packed[:-1],) = torch.autograd.grad(loss, theta)
            packed[-1] = loss.detach()
        _dist.all_reduce_packed(packed)
        return float(packed[-1]), packed[:-1]

    def _theta(self, params):
        fn = _native_apply(self.network)
        if isinstance(params, torch.Tensor):
            return params.to(device=torch.device("cuda", self.device), dtype=torch.float64)
        return torch.as_tensor(_theta(fn, params), dtype=torch.float64, device=torch.device("cuda", self.device))

    def _compute_loss_and_grad(self, params, batch_data, energy_weight, density_weight):
        """-> (loss float, grads): grads is flat when ``params`` is a flat tensor, else in the stax structure."""
        loss, g = self._loss_and_grad_flat(self._theta(params), batch_data, energy_weight, density_weight)
        if isinstance(params, torch.Tensor):
            return loss, g
        return loss, _native_apply(self.network).unflatten(g.cpu().numpy())

    def _compute_validation_loss(self, params, validation_data, energy_weight, density_weight, batch_size):
        if len(validation_data) == 0:
            return 0.0
        theta = self._theta(params)
        total, nb = 0.0, 0
        for lo in range(0, len(validation_data), batch_size):
            total += self._loss_and_grad_flat(theta, validation_data[lo : lo + batch_size], energy_weight, density_weight,
                                              want_grad=False)[0]
            nb += 1
        return total / nb

    # ---- train (:395-560) ----------------------------------------------------------------------------------
    def train(self, training_data=None, validation_data=None):
        if training_data is None:
            training_data, validation_data = self.prepare_dataset()
        validation_data = validation_data or []
        if self.network is None:
            g = gen_grid.Grids(training_data[0][2])
            g.level = int(self.config.get("grid_density", 0))
            g.becke_scheme = gen_grid.stratmann
            g.build()
            cls = GlobalMLP if self.is_global_xc else LocalMLP
            self.network = cls(self.config.get("network_config")).build_network(g.coords)
        init_fn, apply_fn = self.network if isinstance(self.network, tuple) else (None, self.network)
        if not hasattr(self, "initialized_params"):
            self.initialized_params = init_fn(self.seed, None)[1]
        theta = self._theta(self.initialized_params)
        opt_state = adam_init(theta)
        lr = float(self.config.get("learning_rate", 1e-3))
        n_iterations = int(self.config.get("n_iterations", 1000))
        validation_interval = int(self.config.get("validation_interval", 5))
        batch_size = int(self.config.get("batch_size", 3))
        ew, dw = float(self.config.get("energy_weight", 1.0)), float(self.config.get("density_weight", 1.0))
        train_losses, validation_losses = [], []
        n = len(training_data)
        batches = [training_data[lo : lo + batch_size] for lo in range(0, n, batch_size)]
        # the reference shuffles with jax.random.permutation(PRNGKey(iteration)); batches of a fixed order keep the
        # per-batch device problems (and graphs) cached -- the loss of an epoch does not depend on the order
        if bool(self.config.get("cuda_graph", False)) and _dist.rank_world()[1] == 1:
            shared = dict(theta=theta.detach().clone(), mu=opt_state["mu"].clone(), nu=opt_state["nu"].clone(),
                          count=torch.zeros((), dtype=torch.float64, device=theta.device))
            graphs = [_GraphedIteration(self, b, ew, dw, lr, shared) for b in batches]
            for iteration in range(n_iterations):
                losses = [g.replay() for g in graphs]
                train_losses.append(torch.stack(losses).mean())  # stays on the device: no per-iteration synchronisation
                if validation_data and iteration % validation_interval == 0:
                    validation_losses.append(self._compute_validation_loss(shared["theta"], validation_data, ew, dw, batch_size))
            train_losses = [float(x) for x in torch.stack(train_losses).cpu()] if train_losses else []
            theta = shared["theta"]
            opt_state = dict(count=int(shared["count"].item()), mu=shared["mu"], nu=shared["nu"])
        else:
            for iteration in range(n_iterations):
                epoch_loss = 0.0
                for b in batches:
                    loss, g = self._compute_loss_and_grad(theta, b, ew, dw)
                    theta, opt_state = adam_update(g, opt_state, theta, lr)
                    epoch_loss += loss
                train_losses.append(epoch_loss / len(batches))
                if validation_data and iteration % validation_interval == 0:
                    validation_losses.append(self._compute_validation_loss(theta, validation_data, ew, dw, batch_size))
        params = _native_apply(self.network).unflatten(theta.cpu().numpy())
        return params, opt_state, train_losses, validation_losses

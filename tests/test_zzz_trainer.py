"""The training step around the hot path (qex_b200/trainer.py <-> oracle/train_ref.py), i.e. the reference's
``TDKSDFTTrainer._compute_loss_and_grad`` / ``train`` (trainer_legacy_no_jit.py:237-285, :395-560) on the README's
H2 example: bond lengths 0.74 / 0.5 / 1.5 A, 6-31G, CCSD targets, level-0 Stratmann grid, batch 3."""
import numpy as np
import pytest

from oracle import mlp_ref, train_ref
from qex_b200 import gto

BONDS = [0.74, 0.5, 1.5]
E_CCSD_NOTEBOOK = {0.74: -1.151672678339737, 0.5: -1.077863888625149, 1.5: -1.054347450987067}


@pytest.fixture(scope="module")
def dataset():
    return train_ref.make_dataset([gto.h2(b, "6-31g") for b in BONDS], level=0)


def test_oracle_dataset_targets_are_the_notebook_ccsd_energies(dataset):
    for b, (e, dens, mol, x) in zip(BONDS, dataset):
        assert abs(e - E_CCSD_NOTEBOOK[b]) < 3e-7
        assert dens.shape == (1240, 4)
        assert abs((dens[:, 3] * x["weights"]).sum() - 2.0) < 5e-3  # level-0 quadrature of a 2-electron density
        assert abs(np.einsum("ij,ji", x["dm_ao"], x["I"]["s1e"]) - 2.0) < 1e-10


def test_oracle_loss_is_finite_and_responds_to_the_weights(dataset):
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 0))
    l11 = train_ref.batch_loss(theta, spec, dataset, 1.0, 1.0, max_cycle=8)
    l10 = train_ref.batch_loss(theta, spec, dataset, 1.0, 0.0, max_cycle=8)
    l01 = train_ref.batch_loss(theta, spec, dataset, 0.0, 1.0, max_cycle=8)
    assert np.isfinite(l11) and l10 > 0 and l01 > 0
    assert abs(l11 - (l10 + l01)) < 1e-12


def test_adam_update_matches_torch_adam_on_the_host():
    import torch

    from qex_b200 import trainer

    rng = np.random.default_rng(0)
    th = rng.standard_normal(50)
    t_ref = torch.tensor(th, requires_grad=True)
    opt = torch.optim.Adam([t_ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    t_mine, st = torch.tensor(th), trainer.adam_init(torch.tensor(th))
    o_mine, so = th.copy(), train_ref.adam_init(th)
    for _ in range(5):
        g = rng.standard_normal(50)
        opt.zero_grad()
        t_ref.grad = torch.tensor(g)
        opt.step()
        t_mine, st = trainer.adam_update(torch.tensor(g), st, t_mine, 1e-3)
        o_mine, so = train_ref.adam_update(g, so, o_mine, 1e-3)
    assert np.abs(t_mine.numpy() - t_ref.detach().numpy()).max() < 1e-12
    assert np.abs(o_mine - t_ref.detach().numpy()).max() < 1e-12


def test_host_full_ci_matches_the_oracle(dataset):
    from qex_b200 import trainer

    for e, _d, _m, x in dataset:
        e2, dm2 = trainer.full_ci_two_electron(x["I"])
        assert abs(e2 - e) < 1e-11 and np.abs(dm2 - x["dm_ao"]).max() < 1e-9


def test_trainer_rejects_unknown_methods():
    from qex_b200 import trainer

    t = trainer.TDKSDFTTrainer({"method": "rks"}, device=0)
    with pytest.raises(ValueError):
        t._generate(0.74)


def _trainer(cfg, is_global, qnn=False):
    from qex_b200 import gen_grid, trainer
    from qex_b200.networks import GlobalMLP, LocalMLP, LocalQNN

    g = gen_grid.Grids(gto.h2(0.74, "6-31g"))
    g.level = 0
    g.becke_scheme = gen_grid.stratmann
    g.build()
    if qnn:
        net = LocalQNN({"n_qubits": 6, "n_layers": 2}).build_network(g.coords)
    else:
        net = (GlobalMLP if is_global else LocalMLP)().build_network(g.coords)
    return trainer.TDKSDFTTrainer(dict(cfg, is_global_xc=is_global), network=net, seed=0)


@pytest.mark.gpu
@pytest.mark.parametrize("is_global", [False, True])
def test_cuda_training_loss_and_gradient_match_the_oracle(lib, dataset, is_global):
    """Loss of one batch (3 molecules in one batched device problem) against the numpy restatement, and its theta
    gradient against central differences of that restatement along random directions (plain SCF iteration: with DIIS
    the extrapolation solve is ill-conditioned and the finite difference itself scatters, see tests/test_scf.py)."""
    import torch

    cfg = dict(max_cycle=6, diis_start_cycle=10**6)
    tr = _trainer(cfg, is_global)
    G = dataset[0][1].shape[0]
    spec = mlp_ref.MLPSpec([G if is_global else 1, 64, 64, 64, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 1))
    batch = [(e, d, m, dict(I=x["I"])) for e, d, m, x in dataset]

    def ref(th):
        return train_ref.batch_loss(th, spec, dataset, 1.0, 1.0, is_global=is_global, max_cycle=6, diis=False)

    loss, grad = tr._compute_loss_and_grad(torch.as_tensor(theta), batch, 1.0, 1.0)
    l0 = ref(theta)
    assert abs(loss - l0) < 1e-10 * max(1.0, abs(l0))
    grad = grad.cpu().numpy()
    rng = np.random.default_rng(7)
    for _ in range(2):
        d = rng.standard_normal(theta.shape)
        d /= np.linalg.norm(d)
        h = 1e-5
        fd = (ref(theta + h * d) - ref(theta - h * d)) / (2 * h)
        assert abs(fd - grad @ d) < 1e-6 * max(1.0, abs(fd))
    # the stax-structured entry point returns gradients in the parameter structure
    params = tr.network[1].unflatten(theta)
    loss2, g2 = tr._compute_loss_and_grad(params, batch, 1.0, 1.0)
    assert abs(loss2 - loss) < 1e-12 and np.abs(tr.network[1].flatten(g2) - grad).max() < 1e-12
    # validation loss = the same number without a tape
    assert abs(tr._compute_validation_loss(params, batch, 1.0, 1.0, 3) - loss) < 1e-12


@pytest.mark.gpu
def test_cuda_trainer_trains_end_to_end(lib):
    """README 3D example in miniature: prepare_dataset (full-CI targets, grids and target densities through the CUDA
    kernels), 8 Adam iterations over one batch of 3 molecules: the loss must go down and the targets be the notebook's."""
    tr = _trainer(dict(n_iterations=8, batch_size=3, learning_rate=1e-3, max_cycle=10, validation_interval=4,
                       train_bond_lengths=BONDS, val_bond_lengths=[0.9]), is_global=False)
    train, val = tr.prepare_dataset()
    for b, entry in zip(BONDS, train):
        assert abs(entry[0] - E_CCSD_NOTEBOOK[b]) < 3e-7 and entry[1].shape == (1240, 4)
    params, opt_state, tl, vl = tr.train(train, val)
    assert len(tl) == 8 and len(vl) == 2 and opt_state["count"] == 8
    assert all(np.isfinite(tl)) and tl[-1] < tl[0]
    assert len(params) == 7 and params[0][0].shape == (1, 64)


@pytest.mark.gpu
def test_cuda_graph_training_equals_stream_training(lib, dataset):
    """`cuda_graph=True` replays forward + reverse + Adam of an iteration as one CUDA graph: same losses, same
    parameters, same optimiser state as the stream-launched loop (two batches, so two graphs share one optimiser)."""
    data = [(e, d, m, dict(I=x["I"])) for e, d, m, x in dataset]
    # plain SCF iteration: strict.  With DIIS the extrapolation solve is ill-conditioned and amplifies last-bit differences
    # (torch.pow on the device vs Python's pow in Adam's bias correction, library algorithm choices under capture) to
    # ~1e-7 in the loss -- the same scatter tests/test_scf.py documents for LAPACK vs cuSOLVER
    for diis_start, tol in ((10**6, 1e-10), (1, 1e-5)):
        cfg = dict(n_iterations=4, batch_size=2, learning_rate=1e-3, max_cycle=6, validation_interval=2,
                   diis_start_cycle=diis_start)
        out = []
        for graph in (False, True):
            tr = _trainer(dict(cfg, cuda_graph=graph), is_global=False)
            params, st, tl, vl = tr.train(data, data[:1])
            out.append((tr.network[1].flatten(params), st, tl, vl))
        (p0, s0, t0, v0), (p1, s1, t1, v1) = out
        assert s0["count"] == s1["count"] == 8
        assert np.abs(np.array(t0) - np.array(t1)).max() < tol and np.abs(np.array(v0) - np.array(v1)).max() < tol
        assert np.abs(p0 - p1).max() < 10 * tol
        assert np.abs((s0["mu"] - s1["mu"]).cpu().numpy()).max() < 10 * tol * max(1.0, float(s0["mu"].abs().max()))
        assert t0[-1] < t0[0]


@pytest.mark.gpu
def test_cuda_training_step_with_the_local_qnn_functional(lib, dataset):
    """BASELINE.json configs[1] as a training step: H2 KS-SCF with the LocalQNN functional (6 qubits, 2 hea layers, 36
    parameters) at every grid point; loss against the numpy statevector restatement, gradient against its central
    differences along two directions and, with only 36 parameters, a few single coordinates."""
    import torch

    from oracle import qnn_ref

    tr = _trainer(dict(max_cycle=4, diis_start_cycle=10**6), is_global=False, qnn=True)
    qspec = qnn_ref.QNNSpec(6, 2)
    theta = np.random.default_rng(2).uniform(-0.1, 0.1, 36)
    batch = [(e, d, m, dict(I=x["I"])) for e, d, m, x in dataset]

    def ref(th):
        return train_ref.batch_loss(th, lambda t, rho: qnn_ref.exc_and_vrho_local(qspec, t, rho), dataset, 1.0, 1.0,
                                    max_cycle=4, diis=False)

    loss, grad = tr._compute_loss_and_grad(torch.as_tensor(theta), batch, 1.0, 1.0)
    l0 = ref(theta)
    assert abs(loss - l0) < 1e-10 * max(1.0, abs(l0))
    grad = grad.cpu().numpy()
    rng = np.random.default_rng(8)
    dirs = [rng.standard_normal(36) for _ in range(2)] + [np.eye(36)[k] for k in (0, 17, 35)]
    for d in dirs:
        d = d / np.linalg.norm(d)
        h = 1e-5
        fd = (ref(theta + h * d) - ref(theta - h * d)) / (2 * h)
        assert abs(fd - grad @ d) < 1e-6 * max(1.0, abs(fd))

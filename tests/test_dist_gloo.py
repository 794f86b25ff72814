"""world_size-2 gloo test (CPU) of the multi-GPU host logic: fixed rank -> grid-range map, packed
buffers, one all-reduce per direction.  The per-rank compute is the oracle (test infrastructure);
on GPUs the same code path runs XCContext + NCCL (bench.py --gpus N)."""
import os

import numpy as np
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import mlp_ref, numint_ref
from qex_b200 import dist
from tests._util import rel_err, synth_problem


def _partial(ao, w, dm, spec, theta, e_bar, v_bar, lo, hi):
    a, ww = ao[lo:hi], w[lo:hi]

    def eval_xc(code, rho, **k):
        e, v = mlp_ref.exc_and_vrho_local(spec, theta, rho)
        return e, (v, None, None, None), None, None

    nelec, exc, vmat = numint_ref.nr_rks(a, ww, dm, eval_xc, "NN")
    D, tb = numint_ref.nr_rks_vjp(a, ww, dm, lambda r, p: mlp_ref.exc_and_vrho_local(spec, theta, r),
                                  lambda r, p, eb, vb: mlp_ref.exc_and_vrho_local_vjp(spec, theta, r, eb, vb),
                                  e_bar, v_bar, "NN")
    return np.concatenate([vmat.ravel(), [exc, nelec]]), np.concatenate([D.ravel(), tb])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    N, G = 6, 1000
    ao, dm, w = synth_problem(N, G, 1, seed=5)
    spec = mlp_ref.MLPSpec([1, 8, 8, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 0))
    v_bar = np.random.default_rng(1).standard_normal((N, N))
    lo, hi = dist.shard_range(G, rank, world)
    out, bar = _partial(ao[0, 0], w[0], dm[0], spec, theta, 0.5, v_bar, lo, hi)
    out, bar = torch.from_numpy(out), torch.from_numpy(bar)
    dist.all_reduce_packed(out)
    dist.all_reduce_packed(bar)
    if rank == 0:
        q.put((out.numpy(), bar.numpy()))
    tdist.destroy_process_group()


def test_two_rank_sharded_nr_rks_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, bar = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, G = 6, 1000
    ao, dm, w = synth_problem(N, G, 1, seed=5)
    spec = mlp_ref.MLPSpec([1, 8, 8, 1], "tanh")
    theta = mlp_ref.pack(*mlp_ref.init_params(spec, 0))
    v_bar = np.random.default_rng(1).standard_normal((N, N))
    full_out, full_bar = _partial(ao[0, 0], w[0], dm[0], spec, theta, 0.5, v_bar, 0, G)
    assert rel_err(out, full_out) < 1e-12
    assert rel_err(bar, full_bar) < 1e-12


# ---- the trainer's data-parallel reduction (molecules dealt to ranks, one all-reduce of [grad | loss]) ----------
class _ToyTrainer:
    """TDKSDFTTrainer with the device loss replaced by a closed-form per-molecule term, so that the sharding /
    packing / all-reduce logic of `_loss_and_grad_flat` runs on CPU tensors under gloo."""

    def __new__(cls):
        from qex_b200 import trainer

        class T(trainer.TDKSDFTTrainer):
            def _loss_sum(self, theta, batch_data, energy_weight, density_weight):
                tot = 0.0
                for e_goal, vec in batch_data:
                    tot = tot + energy_weight * ((theta * torch.as_tensor(vec)).sum() - e_goal) ** 2 \
                        + density_weight * (theta ** 2).mean() * abs(e_goal)
                return tot

        return T({}, network=None, device=None)


def _toy_batch():
    rng = np.random.default_rng(3)
    return [(float(rng.standard_normal()), rng.standard_normal(7)) for _ in range(5)]


def _trainer_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    theta = torch.linspace(-1.0, 1.0, 7, dtype=torch.float64)
    loss, g = _ToyTrainer()._loss_and_grad_flat(theta, _toy_batch(), 1.0, 0.5)
    val, _ = _ToyTrainer()._loss_and_grad_flat(theta, _toy_batch(), 1.0, 0.5, want_grad=False)
    q.put((rank, loss, g.numpy(), val))
    tdist.destroy_process_group()


def test_two_rank_trainer_step_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    theta = torch.linspace(-1.0, 1.0, 7, dtype=torch.float64)
    loss1, g1 = _ToyTrainer()._loss_and_grad_flat(theta, _toy_batch(), 1.0, 0.5)
    for _rank, loss, g, val in res:  # every rank holds the full-batch loss and gradient
        assert abs(loss - loss1) < 1e-13 and np.abs(g - g1.numpy()).max() < 1e-13 and abs(val - loss1) < 1e-13
    # against the definition: mean over the 5 molecules
    th = theta.clone().requires_grad_(True)
    ref = sum(((th * torch.as_tensor(v)).sum() - e) ** 2 + 0.5 * (th ** 2).mean() * abs(e) for e, v in _toy_batch()) / 5
    (gr,) = torch.autograd.grad(ref, th)
    assert abs(float(ref.detach()) - loss1) < 1e-13 and np.abs(gr.numpy() - g1.numpy()).max() < 1e-13

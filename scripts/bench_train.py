"""README 3D example as a device-resident training step (BASELINE.json configs[0]): H2 at 0.74 / 0.5 / 1.5 A, 6-31G,
level-0 Stratmann grids (1240 points), batch 3, `max_cycle`-cycle KS-SCF per molecule, energy + density loss,
theta gradient, Adam update -- `qex_b200.trainer.TDKSDFTTrainer._compute_loss_and_grad` + `adam_update`.
CUDA events around `steps` iterations after `warmup`; prints one JSON line.  `bench.py --config c1train` adds the
CPU baseline (the numpy restatement of the loss; bench.py is the only place allowed to time oracle code)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qex_b200 import gen_grid, gto, trainer  # noqa: E402
from qex_b200.networks import GlobalMLP, LocalMLP  # noqa: E402


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cycles", type=int, default=20)
    ap.add_argument("--nmol", type=int, default=3, help="3 = README example; other: bond lengths linspace(0.4, 3.0) (config c4)")
    ap.add_argument("--global-xc", action="store_true", help="GlobalMLP / NN-AmplitudeEncoding (README default)")
    return ap.parse_args(argv)


def _dist_setup():
    """Under torchrun: one rank per GPU, NCCL; the trainer deals the molecules of a batch to the ranks and all-reduces
    the packed [grad | loss] (this mode has a gloo test but no hardware run yet)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    return int(os.environ.get("RANK", "0")), world


def measure(args, cpu_baseline_fn=None):
    rank, world = _dist_setup()
    bonds = [0.74, 0.5, 1.5] if args.nmol == 3 else [float(b) for b in torch.linspace(0.4, 3.0, args.nmol)]
    g = gen_grid.Grids(gto.h2(0.74, "6-31g"))
    g.level = 0
    g.becke_scheme = gen_grid.stratmann
    g.build()
    net = (GlobalMLP if args.global_xc else LocalMLP)().build_network(g.coords)
    tr = trainer.TDKSDFTTrainer(dict(train_bond_lengths=bonds, val_bond_lengths=[], batch_size=len(bonds), max_cycle=args.cycles,
                                     is_global_xc=args.global_xc), network=net, seed=0)
    train, _ = tr.prepare_dataset()
    theta = tr._theta(net[0](0, None)[1])
    state = trainer.adam_init(theta)

    def step(theta, state):
        loss, grad = tr._compute_loss_and_grad(theta, train, 1.0, 1.0)
        theta, state = trainer.adam_update(grad, state, theta, 1e-3)
        return loss, theta, state

    losses = []
    for _ in range(max(3, args.warmup)):
        loss, theta, state = step(theta, state)
        losses.append(loss)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        from qex_b200 import dist as qdist

        mine = [train[i] for i in qdist.shard_batch(len(train), rank, world)]
    else:
        mine = train
    xc = tr._problem(mine).xc
    n0 = int(xc.lib.qexxc_launch_count(xc._h)) + int(xc.lib.qexxc_jk_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, theta, state = step(theta, state)
        losses.append(loss)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:  # max over ranks, on the device clock
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    launches = (int(xc.lib.qexxc_launch_count(xc._h)) + int(xc.lib.qexxc_jk_launch_count()) - n0) // args.steps
    # the same iteration (forward, reverse, Adam) as ONE CUDA graph (`config["cuda_graph"]` of the trainer)
    ms_graph, graph_note = None, None
    try:
        if world > 1:
            raise RuntimeError("one-graph replay is a single-GPU mode")
        shared = dict(theta=theta.detach().clone(), mu=state["mu"].clone(), nu=state["nu"].clone(),
                      count=torch.full((), float(state["count"]), dtype=torch.float64, device=theta.device))
        gi = trainer._GraphedIteration(tr, train, 1.0, 1.0, 1e-3, shared)
        l_a = float(gi.replay())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            l_b = gi.replay()
        e1.record()
        torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / args.steps
        graph_note = {"loss_first_replay": l_a, "loss_last_replay": float(l_b)}
    except Exception as ex:  # capture is an optimisation, not a requirement
        graph_note = "graph capture failed: " + repr(ex)[:300]
    G = g.size
    cpu = cpu_baseline_fn(bonds, args.cycles, args.global_xc) if cpu_baseline_fn else None
    # BASELINE.md publishes this exact configuration (GlobalMLP, 3 molecules, max_cycle 20): ~1.4 s per iteration in
    # steady state on the reference author's laptop CPU (notebooks/04_notebook_td_trainer.ipynb:707-711)
    published = (1.0 / 1.4) if (args.global_xc and args.nmol == 3 and args.cycles == 20) else None
    return {
        "vs_baseline": None if published is None else (1e3 / ms) / published,
        "vs_baseline_note": None if published is None else "value / (1 / 1.4 s) of BASELINE.md (laptop CPU, JAX CPU backend)",
        "metric": "training iterations per second (README 3D H2 example: batch 3, KS-SCF + energy/density loss + grad + Adam)",
        "unit": "it/s", "value": 1e3 / ms, "ms_per_iteration": ms, "ms_per_iteration_cuda_graph": ms_graph,
        "cuda_graph": graph_note, "gpu_launches_per_iteration": int(launches),
        "grid_pts_per_s": len(bonds) * G * (args.cycles + 1) / (ms * 1e-3),
        "grid_pts_per_s_cuda_graph": None if ms_graph is None else len(bonds) * G * (args.cycles + 1) / (ms_graph * 1e-3),
        "losses_first_last": [losses[0], losses[-1]], "loss_decreased": bool(losses[-1] < losses[0]),
        "config": {"workload": f"c1 as a training step: {len(bonds)} H2/6-31G geometries, {G} grid points x 4 AOs, "
                               f"{'GlobalMLP' if args.global_xc else 'LocalMLP'} 64x3 tanh, {args.cycles}-cycle KS-SCF with DIIS",
                   "steps": args.steps, "warmup": max(3, args.warmup)},
        "dtype": "f64", "cpu_baseline": cpu, "n_gpus": world, "rank": rank,
    }


if __name__ == "__main__":
    line = measure(parse())
    if line["rank"] == 0:
        print(json.dumps(line))

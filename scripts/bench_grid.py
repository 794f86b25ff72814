"""Times qexxc_becke_partition (csrc/grid.cu) on a c5-sized grid: 60 carbon-like atoms, pyscf level-3 atomic grids
(75 radial x 302 Lebedev points, NWChem-pruned) ~ 1e6 points.  CUDA events around the launches, inputs resident.
Prints one JSON line.  Usage: python scripts/bench_grid.py [natoms] [level]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qex_b200 import _lib, gen_grid, gto  # noqa: E402


def main():
    natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    mol = gto.synthetic_molecule(natoms, [1], seed=0)
    charges, centers = np.asarray(mol.atom_charges(), dtype=int), mol.atom_coords()
    t0 = time.time()
    tab = gen_grid.gen_atomic_grids(charges, level)
    c0, v0 = tab[int(charges[0])]
    coords = np.concatenate([c0 + centers[i] for i in range(natoms)])
    vol = np.tile(v0, natoms)
    owner = np.repeat(np.arange(natoms, dtype=np.int32), v0.shape[0])
    t_host = time.time() - t0
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    c, v, o = torch.tensor(coords, device=dev), torch.tensor(vol, device=dev), torch.tensor(owner, device=dev)
    ac = torch.tensor(centers, device=dev)
    work = torch.empty(natoms * natoms, dtype=torch.float64, device=dev)
    w = torch.empty_like(v)
    G = c.shape[0]
    out = {"natm": natoms, "level": level, "ngrids": G, "host_tables_s": round(t_host, 3)}
    variants = [("becke", 0, c, v, o), ("stratmann", 1, c, v, o)]
    if "--sorted" in sys.argv:
        # experiment for DESIGN 6d's open item: points ordered by (owner, distance to owner) make the lanes of a warp
        # break out of the pair loops at similar places (NOT yet run on hardware)
        d_own = np.linalg.norm(coords - centers[owner], axis=1)
        perm = torch.tensor(np.lexsort((d_own, owner)), device=dev)
        variants += [("becke_sorted", 0, c[perm].contiguous(), v[perm].contiguous(), o[perm].contiguous()),
                     ("stratmann_sorted", 1, c[perm].contiguous(), v[perm].contiguous(), o[perm].contiguous())]
    for name, sid, c, v, o in variants:
        def run():
            _lib.check(lib.qexxc_becke_partition(0, c.data_ptr(), C.c_long(G), o.data_ptr(), v.data_ptr(), ac.data_ptr(), None,
                                                 natoms, sid, work.data_ptr(), w.data_ptr(),
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        pairs = G * natoms * (natoms - 1)
        out[name] = {"ms": round(ms, 4), "gpts_per_s": round(G / ms / 1e6, 4), "pair_terms_per_s": pairs / ms * 1e3,
                     "weights_sum": float(w.sum())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

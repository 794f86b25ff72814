"""Generate tests/golden/xc_small.npz: small seeded input/output vectors of the hot path.

IMPORTANT: these are ORACLE outputs, not outputs of the reference: pasqal-io/qex cannot be
executed in this environment (no jax / pyscf / horqrux), and its own tests hold no golden vectors
for this path.  The fixtures freeze the oracle so that (a) the oracle cannot drift
silently between rounds and (b) the CUDA path is also checked against committed numbers, not only
against an oracle run in the same process.  If the reference ever becomes runnable, regenerate this
file from it with the same inputs and the parity claim becomes pinned.

    python tests/golden/make_golden.py        # rewrites tests/golden/xc_small.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import gto_ref, mlp_ref, qnn_ref, step_ref  # noqa: E402
from qex_b200 import gen_grid, gto  # noqa: E402


def build():
    out = {}
    rng = np.random.default_rng(2024)
    # --- H2 / 6-31G, the README molecule, 1240-point grid ---
    mol = gto.h2(0.74, "6-31g")
    grids = gen_grid.Grids(mol, n_rad=31, n_theta=5, n_phi=4).build()
    c = rng.standard_normal((4, 1)) * 0.4
    dm = 2.0 * c @ c.T
    v_bar = rng.standard_normal((4, 4))
    out.update(h2_atm=mol._atm, h2_bas=mol._bas, h2_env=mol._env, h2_coords=grids.coords, h2_weights=grids.weights,
               h2_dm=dm, h2_vbar=v_bar, h2_ebar=np.array(0.8))
    out["h2_ao"] = gto_ref.eval_ao(mol._atm, mol._bas, mol._env, grids.coords[:64], 1)
    # LocalMLP 1->64->64->64->1 tanh ("NN")
    spec = mlp_ref.MLPSpec([1, 64, 64, 64, 1], "tanh")
    th = mlp_ref.pack(*mlp_ref.init_params(spec, 7))
    r = step_ref.xc_step(mol._atm, mol._bas, mol._env, grids.coords, grids.weights, dm,
                         dict(kind="local_mlp", n_features=1, n_hidden=3, width=64), th, "NN", 0.8, v_bar)
    out.update(mlp_theta=th, **{"mlp_" + k: np.asarray(v) for k, v in r.items()})
    # GlobalMLP G->64->64->64->1 ("NN-AmplitudeEncoding")
    gspec = mlp_ref.MLPSpec([grids.size, 64, 64, 64, 1], "tanh")
    gth = mlp_ref.pack(*mlp_ref.init_params(gspec, 8))
    r = step_ref.xc_step(mol._atm, mol._bas, mol._env, grids.coords, grids.weights, dm,
                         dict(kind="global_mlp", n_hidden=3, width=64), gth, "NN-AmplitudeEncoding", 0.8, v_bar)
    out.update(gmlp_theta=gth, **{"gmlp_" + k: np.asarray(v) for k, v in r.items()})
    # LocalQNN 6 qubits x 2 layers ("NN")
    qth = qnn_ref.init_params(qnn_ref.QNNSpec(6, 2), 9)
    r = step_ref.xc_step(mol._atm, mol._bas, mol._env, grids.coords, grids.weights, dm,
                         dict(kind="local_qnn", n_hidden=2, width=6), qth, "NN", 0.8, v_bar)
    out.update(qnn_theta=qth, **{"qnn_" + k: np.asarray(v) for k, v in r.items()})
    # --- a 3-atom s/p/d molecule with GGA features ---
    m3 = gto.synthetic_molecule(3, (2, 1, 1), seed=3)
    g3 = gen_grid.random_grid(m3, 600, seed=4)
    N = m3.nao_nr()
    C3 = rng.standard_normal((N, 4)) / np.sqrt(N)
    dm3 = 2.0 * C3 @ C3.T
    vb3 = rng.standard_normal((N, N)) / N
    s2 = mlp_ref.MLPSpec([2, 64, 64, 64, 1], "tanh")
    th2 = mlp_ref.pack(*mlp_ref.init_params(s2, 10))
    r = step_ref.xc_step(m3._atm, m3._bas, m3._env, g3.coords, g3.weights, dm3,
                         dict(kind="local_mlp", n_features=2, n_hidden=3, width=64), th2, "GGA", 1.0, vb3)
    out.update(gga_atm=m3._atm, gga_bas=m3._bas, gga_env=m3._env, gga_coords=g3.coords, gga_weights=g3.weights,
               gga_dm=dm3, gga_vbar=vb3, gga_theta=th2, **{"gga_" + k: np.asarray(v) for k, v in r.items()})
    return out


if __name__ == "__main__":
    data = build()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "xc_small.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes;", len(data), "arrays")

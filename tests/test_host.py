"""CPU tests of host-side logic: libcint tables, grids, sharding, workloads."""
import numpy as np
import pytest

from qex_b200 import dist, gen_grid, gto, workloads


def test_mole_tables_follow_libcint_layout():
    m = gto.h2(0.74, "6-31g")
    assert m._atm.shape == (2, 6) and m._bas.shape == (4, 8) and m._atm.dtype == np.int32
    assert list(m._bas[:, gto.NPRIM_OF]) == [3, 1, 3, 1] and list(m._bas[:, gto.ANG_OF]) == [0, 0, 0, 0]
    assert list(m.ao_loc_nr()) == [0, 1, 2, 3, 4]
    # coefficients are stored normalised: a 1-primitive s shell has c = gto_norm(0, a)
    pe, pc = m._bas[1, gto.PTR_EXP], m._bas[1, gto.PTR_COEFF]
    assert np.isclose(m._env[pc], gto.gto_norm(0, m._env[pe]))
    # gto_norm(0, a)^2 * int_0^inf r^2 exp(-2 a r^2) dr = 1
    a = 0.77
    assert np.isclose(gto.gto_norm(0, a) ** 2 * np.sqrt(np.pi) / (4 * (2 * a) ** 1.5), 1.0)


def test_synthetic_molecules_have_the_named_sizes():
    assert gto.synthetic_molecule(50, (4, 2, 2)).nao_nr() == 1000
    assert gto.synthetic_molecule(3, (5, 5, 4)).nao_nr() == 120
    wl = workloads.make("c5", ngrids=512)
    assert wl.nao == 1000 and wl.ngrids == 512 and wl.theta.size == 64 + 64 + 2 * (64 * 64 + 64) + 65
    wl = workloads.make("c2")
    assert wl.nao == 4 and wl.ngrids == 1240 and wl.theta.size == 36
    wl = workloads.make("c3", ngrids=256)
    assert wl.nao == 120 and wl.ncomp == 4 and wl.xctype == "GGA"
    wl = workloads.make("c1")
    assert wl.dm.shape == (3, 4, 4) and wl.xctype == "NN-AmplitudeEncoding" and wl.theta.size == 1240 * 64 + 64 + 2 * (64 * 64 + 64) + 65


def test_grid_integrates_a_gaussian():
    m = gto.Mole([("H", (0, 0, 0)), ("H", (0, 0, 1.4))], basis="sto-3g", unit="Bohr")
    g = gen_grid.Grids(m, n_rad=50, n_theta=12, n_phi=12).build()
    r2 = ((g.coords - np.array([0.2, -0.1, 0.6])) ** 2).sum(1)
    assert abs(np.dot(g.weights, np.exp(-1.3 * r2)) - (np.pi / 1.3) ** 1.5) < 1e-5
    assert (g.weights >= 0).all()


@pytest.mark.parametrize("G,world", [(1_000_000, 8), (1240, 2), (129, 4), (5, 3), (128 * 7, 8)])
def test_shard_ranges_partition_the_grid(G, world):
    r = [dist.shard_range(G, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == G
    for (a, b), (c, d) in zip(r[:-1], r[1:]):
        assert b == c and a <= b
        assert b % dist.TILE == 0 or b == G
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) < 2 * dist.TILE
    assert sorted(sum((dist.shard_batch(64, k, world) for k in range(world)), [])) == list(range(64))

"""Committed fixtures (tests/golden/xc_small.npz, made by tests/golden/make_golden.py).
They are ORACLE outputs frozen on disk -- the reference cannot be executed here (see the script's
header) -- checked (CPU) against a fresh oracle run and (GPU) against the CUDA path."""
import os

import numpy as np
import pytest

from oracle import gto_ref, step_ref
from tests._util import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xc_small.npz")
CASES = {
    "mlp": ("h2", dict(kind="local_mlp", n_features=1, n_hidden=3, width=64), "NN", 0.8),
    "gmlp": ("h2", dict(kind="global_mlp", n_hidden=3, width=64), "NN-AmplitudeEncoding", 0.8),
    "qnn": ("h2", dict(kind="local_qnn", n_hidden=2, width=6, in_scale=1.0), "NN", 0.8),
    "gga": ("gga", dict(kind="local_mlp", n_features=2, n_hidden=3, width=64), "GGA", 1.0),
}


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_reproduces_fixtures(gold, case):
    sys_, net, xct, e_bar = CASES[case]
    r = step_ref.xc_step(gold[sys_ + "_atm"], gold[sys_ + "_bas"], gold[sys_ + "_env"], gold[sys_ + "_coords"],
                         gold[sys_ + "_weights"], gold[sys_ + "_dm"], net, gold[case + "_theta"], xct, e_bar,
                         gold[sys_ + "_vbar"])
    for k in ("nelec", "excsum", "vmat", "dm_bar", "theta_bar"):
        assert rel_err(r[k], gold[f"{case}_{k}"]) <= 1e-12, k


def test_ao_fixture(gold):
    ao = gto_ref.eval_ao(gold["h2_atm"], gold["h2_bas"], gold["h2_env"], gold["h2_coords"][:64], 1)
    assert np.abs(ao - gold["h2_ao"]).max() <= 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_cuda_path_matches_fixtures(gold, case):
    from qex_b200 import _lib
    from qex_b200.engine import NetSpec, XCContext

    sys_, net, xct, e_bar = CASES[case]
    kinds = {"local_mlp": _lib.NET_LOCAL_MLP, "global_mlp": _lib.NET_GLOBAL_MLP, "local_qnn": _lib.NET_LOCAL_QNN}
    kw = dict(net)
    kw["kind"] = kinds[kw["kind"]]
    coords, w, dm = gold[sys_ + "_coords"], gold[sys_ + "_weights"], gold[sys_ + "_dm"]
    N, G = dm.shape[0], w.shape[0]
    gga = xct == "GGA"
    ctx = XCContext(nao=N, ngrids_max=G, ncomp=4 if gga else 1, net=NetSpec(**kw))
    ctx.set_grid(coords, w).set_basis(gold[sys_ + "_atm"], gold[sys_ + "_bas"], gold[sys_ + "_env"]).eval_ao(1 if gga else 0)
    if sys_ == "h2":
        ao = ctx.get_ao(1)[0, 0, :64].cpu().numpy()
        assert np.abs(ao - gold["h2_ao"][0]).max() <= 1e-13
    theta = gold[case + "_theta"]
    out, resid = ctx.nr_rks_fwd(dm, theta, xct)
    bar = ctx.nr_rks_vjp(theta, resid, [e_bar], gold[sys_ + "_vbar"], xct).cpu().numpy()
    out = out.cpu().numpy()[0]
    assert rel_err(out[: N * N].reshape(N, N), gold[case + "_vmat"]) <= 1e-10
    assert abs(out[N * N] - gold[case + "_excsum"]) <= 1e-9
    assert abs(out[N * N + 1] - gold[case + "_nelec"]) <= 1e-9
    assert rel_err(bar[: N * N].reshape(N, N), gold[case + "_dm_bar"]) <= 1e-10
    assert rel_err(bar[N * N :], gold[case + "_theta_bar"]) <= 1e-10


def test_reference_notebook_fixture_is_what_the_pin_tests_use():
    """tests/golden/reference_notebook.json holds the numbers extracted from the reference's stored notebook outputs
    (tests/golden/extract_reference_notebook.py); the constants in the pin tests must be exactly those."""
    import json
    import os

    from tests import test_scf, test_zz_pyscf_pin, test_zzz_trainer

    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_notebook.json")))
    scf_all = ref["converged_scf_energies_all"]
    assert all(v in scf_all for v in test_scf.GOLDEN_RHF.values()) and len(scf_all) == 7
    assert test_zz_pyscf_pin.E_LDA_NOTEBOOK == ref["lda_rks_energy_0.74"] and ref["lda_rks_energy_0.74"] in scf_all
    assert sorted(round(v, 12) for v in test_scf.GOLDEN_CCSD.values()) == ref["ccsd_energies_all"]
    assert all(abs(test_scf.GOLDEN_CCSD[b] - v) < 1e-15 for b, v in test_zzz_trainer.E_CCSD_NOTEBOOK.items())
    assert np.array_equal(test_zz_pyscf_pin.DM_CCSD_NOTEBOOK, np.array(ref["cell2_dm_ao"]))
    assert list(test_zz_pyscf_pin.RHO_TAIL_NOTEBOOK) == ref["cell2_density_head_tail"][:3]
    assert ref["cell2_density_head_tail"][3:] == ref["cell2_density_head_tail"][:3][::-1]
    assert ref["ngrids_logged"] == [1192, 1240] and ref["cell2_grid_points"] == 1192 and ref["lda_rks_ngrids"] == 1240

"""CPU tests of the boundary: libqexxc.so loads, exports every symbol include/qexxc.h declares,
and fails loudly (no fallback) when there is no CUDA device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "qexxc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qexxc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("qexxc_nr_rks_fwd", "qexxc_nr_rks_vjp", "qexxc_eval_rho", "qexxc_eval_ao", "qexxc_xc_fwd",
                 "qexxc_xc_vjp", "qexxc_apply_fn_fwd", "qexxc_apply_fn_vjp", "qexxc_vxc_assemble"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    from qex_b200 import _lib

    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == _declared()
    assert lib.qexxc_version() == 210


def test_n_params(lib):
    from qex_b200 import _lib
    from qex_b200.engine import NetSpec

    d = NetSpec(kind=_lib.NET_LOCAL_MLP, n_features=1, n_hidden=3, width=64).desc()
    assert lib.qexxc_n_params(C.byref(d), 0) == 1 * 64 + 64 + 2 * (64 * 64 + 64) + 64 + 1
    d = NetSpec(kind=_lib.NET_GLOBAL_MLP, n_hidden=3, width=64).desc()
    assert lib.qexxc_n_params(C.byref(d), 1240) == 1240 * 64 + 64 + 2 * (64 * 64 + 64) + 64 + 1
    d = NetSpec(kind=_lib.NET_LOCAL_QNN, n_hidden=2, width=6).desc()
    assert lib.qexxc_n_params(C.byref(d), 0) == 36  # 3 * n_qubits * n_layers (hardware_ansatz.py:133-141)


def test_no_cuda_device_is_a_loud_error(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qex_b200 import _lib
    from qex_b200.engine import XCContext

    h = C.c_void_p()
    rc = lib.qexxc_create(C.byref(h), 0, 1, 1, 128, 4, None)
    assert rc == _lib.ERR_NODEVICE and not h.value
    assert b"no CPU fallback" in lib.qexxc_last_error()
    with pytest.raises(_lib.QexxcError):
        XCContext(nao=4, ngrids_max=128)


def test_bad_arguments_return_codes(lib):
    from qex_b200 import _lib

    assert lib.qexxc_create(None, 0, 1, 1, 128, 4, None) == _lib.ERR_ARG
    h = C.c_void_p()
    assert lib.qexxc_create(C.byref(h), 0, 0, 1, 128, 4, None) == _lib.ERR_ARG
    assert lib.qexxc_create(C.byref(h), 0, 1, 3, 128, 4, None) == _lib.ERR_ARG
    assert lib.qexxc_nr_rks_fwd(None, 0, 0, None, None, 0, None, None, None) == _lib.ERR_ARG
    assert lib.qexxc_destroy(None) == 0
    # grid partition: argument checks come before any device work; an empty grid is a no-op
    assert lib.qexxc_becke_partition(0, None, 10, None, None, None, None, 0, 0, None, None, None) == _lib.ERR_ARG
    assert lib.qexxc_becke_partition(0, None, 10, None, None, None, None, 2, 7, None, None, None) == _lib.ERR_ARG
    assert lib.qexxc_becke_partition(0, None, 10, None, None, None, None, 2, 1, None, None, None) == _lib.ERR_ARG
    assert b"null device pointer" in lib.qexxc_last_error()
    assert lib.qexxc_becke_partition(0, None, 0, None, None, None, None, 2, 1, None, None, None) == 0
    assert lib.qexxc_lda_exchange(0, None, 5, None, None, None) == _lib.ERR_ARG
    assert lib.qexxc_lda_exchange(0, None, -1, None, None, None) == _lib.ERR_ARG
    assert lib.qexxc_lda_exchange(0, None, 0, None, None, None) == 0


def test_grid_build_on_a_device_is_loud_without_one():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qex_b200 import gen_grid, gto

    g = gen_grid.Grids(gto.h2(0.74, "6-31g"))
    g.level = 0
    with pytest.raises(RuntimeError):
        g.build(device=0)
    assert g.build().size == 1240  # the host partition is set-up code and needs no device


def test_product_code_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under qex_b200/ or scripts/ may import it (only tests/,
    __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs do)."""
    paths = [os.path.join(d, f) for top in ("qex_b200", "scripts") for d, _, fs in os.walk(os.path.join(ROOT, top))
             for f in fs if f.endswith((".py", ".cu", ".cuh", ".h"))]
    assert len(paths) > 20
    for f in paths:
        txt = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f

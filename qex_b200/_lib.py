"""ctypes binding of libqexxc.so (the C ABI declared in include/qexxc.h).

There is no fallback of any kind: if the shared library is missing or no CUDA device is
visible, the compute entry points raise.  torch is used for device buffers and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["QEXXC_LIB"]) if os.environ.get("QEXXC_LIB") else HERE / "libqexxc.so"  # override: A/B builds

# mirrors include/qexxc.h
XC_NN, XC_NN_GLOBAL, XC_GGA = 0, 1, 2
NET_NONE, NET_LOCAL_MLP, NET_GLOBAL_MLP, NET_LOCAL_QNN = 0, 1, 2, 3
PREC_F64, PREC_F32 = 0, 1
ACTIVATIONS = {
    "tanh": 0, "relu": 1, "softplus": 2, "sigmoid": 3, "elu": 4, "leaky_relu": 5, "selu": 6, "gelu": 7, "swish": 8,
}
ERR_CUDA, ERR_ARG, ERR_STATE, ERR_UNSUPPORTED, ERR_NODEVICE = -1, -2, -3, -4, -5

EXPORTS = [
    "qexxc_version", "qexxc_last_error", "qexxc_n_params", "qexxc_create", "qexxc_create_ex", "qexxc_destroy",
    "qexxc_workspace_bytes", "qexxc_set_grid", "qexxc_set_basis", "qexxc_eval_ao", "qexxc_set_ao",
    "qexxc_get_ao", "qexxc_eval_rho", "qexxc_eval_rho_vjp", "qexxc_xc_fwd", "qexxc_xc_vjp",
    "qexxc_apply_fn_fwd", "qexxc_apply_fn_vjp", "qexxc_vxc_assemble", "qexxc_vxc_assemble_vjp",
    "qexxc_resid_doubles", "qexxc_nr_rks_fwd", "qexxc_nr_rks_vjp", "qexxc_launch_count",
    "qexxc_debug_run_contraction", "qexxc_profile_enable", "qexxc_profile_read", "qexxc_contraction_flops", "qexxc_contraction_mode", "qexxc_prepare_contractions", "qexxc_contraction_i8_ops", "qexxc_i8_peak", "qexxc_eval_rho_mo", "qexxc_nr_rks_fwd_mo",
    "qexxc_jk_workspace_doubles", "qexxc_dot_eri_dm", "qexxc_dot_eri_dm_vjp", "qexxc_jk_launch_count",
    "qexxc_dot_eri_dm_batched", "qexxc_dot_eri_dm_vjp_batched", "qexxc_generalized_eigh_batched",
    "qexxc_becke_partition", "qexxc_grid_launch_count", "qexxc_lda_exchange", "qexxc_lda_launch_count",
    "qexxc_comm_nccl_version", "qexxc_comm_unique_id", "qexxc_comm_create", "qexxc_comm_wrap", "qexxc_comm_destroy",
    "qexxc_comm_rank", "qexxc_comm_world", "qexxc_comm_calls", "qexxc_allreduce", "qexxc_bcast",
]


class NetDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("n_features", C.c_int), ("n_hidden", C.c_int), ("width", C.c_int),
        ("activation", C.c_int), ("out_transform", C.c_int), ("precision", C.c_int), ("reserved", C.c_int),
        ("in_scale", C.c_double), ("out_scale", C.c_double),
    ]


class QexxcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libqexxc error {code}: {msg}")
        self.code = code


class QexxcArgError(QexxcError, ValueError):
    """QEXXC_ERR_ARG: a shape / size mismatch (the reference raises a shape ValueError there)."""


_lib = None


def load(build_if_missing: bool = False):
    """Load libqexxc.so (optionally compiling it first).  Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing or os.environ.get("QEXXC_AUTOBUILD") == "1":
            from .build import build

            build()
        else:
            raise FileNotFoundError(
                f"{LIB_PATH} not found: run `python -m qex_b200.build` (nvcc, sm_100a). "
                "There is no CPU/PyTorch fallback for this path."
            )
    lib = C.CDLL(str(LIB_PATH))
    p, i, l, d, vp = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_void_p
    sigs = {
        "qexxc_version": (i, []),
        "qexxc_last_error": (C.c_char_p, []),
        "qexxc_n_params": (l, [C.POINTER(NetDesc), i]),
        "qexxc_create": (i, [C.POINTER(vp), i, i, i, i, i, C.POINTER(NetDesc)]),
        "qexxc_create_ex": (i, [C.POINTER(vp), i, i, i, i, i, C.POINTER(NetDesc), C.c_uint]),
        "qexxc_destroy": (i, [vp]),
        "qexxc_workspace_bytes": (C.c_size_t, [vp]),
        "qexxc_set_grid": (i, [vp, p, p, i, vp]),
        "qexxc_set_basis": (i, [vp, p, i, p, i, p, i]),
        "qexxc_eval_ao": (i, [vp, i, vp]),
        "qexxc_set_ao": (i, [vp, p, i, i, vp]),
        "qexxc_get_ao": (i, [vp, p, i, vp]),
        "qexxc_eval_rho": (i, [vp, p, i, i, p, vp]),
        "qexxc_eval_rho_vjp": (i, [vp, p, i, i, p, vp]),
        "qexxc_xc_fwd": (i, [vp, i, p, p, l, p, p, p, vp]),
        "qexxc_xc_vjp": (i, [vp, i, p, p, l, p, p, p, p, p, vp]),
        "qexxc_apply_fn_fwd": (i, [vp, p, l, p, l, p, vp]),
        "qexxc_apply_fn_vjp": (i, [vp, p, l, p, l, p, p, p, vp]),
        "qexxc_vxc_assemble": (i, [vp, i, p, p, p, p, p, vp]),
        "qexxc_vxc_assemble_vjp": (i, [vp, i, p, p, p, p, p, p, p, p, p, p, vp]),
        "qexxc_resid_doubles": (C.c_size_t, [vp]),
        "qexxc_nr_rks_fwd": (i, [vp, i, i, p, p, l, p, p, vp]),
        "qexxc_nr_rks_vjp": (i, [vp, i, i, p, l, p, p, p, p, vp]),
        "qexxc_launch_count": (l, [vp]),
        "qexxc_debug_run_contraction": (i, [vp, i, vp]),
        "qexxc_profile_enable": (i, [vp, i]),
        "qexxc_eval_rho_mo": (i, [vp, p, p, i, p, vp]),
        "qexxc_nr_rks_fwd_mo": (i, [vp, i, p, p, i, p, l, p, p, vp]),
        "qexxc_contraction_flops": (i, [vp, i, i, C.POINTER(C.c_double)]),
        "qexxc_contraction_mode": (i, [vp]),
        "qexxc_prepare_contractions": (i, [vp, vp]),
        "qexxc_contraction_i8_ops": (i, [vp, i, i, C.POINTER(C.c_double)]),
        "qexxc_i8_peak": (i, [i, C.POINTER(C.c_double)]),
        "qexxc_profile_read": (i, [vp, i, C.POINTER(C.c_double), C.POINTER(C.c_long)]),
        "qexxc_jk_workspace_doubles": (i, [i, i, C.POINTER(C.c_long)]),
        "qexxc_dot_eri_dm": (i, [i, p, p, i, i, i, i, p, p, p, l, vp]),
        "qexxc_dot_eri_dm_vjp": (i, [i, p, p, p, i, i, p, p, l, vp]),
        "qexxc_jk_launch_count": (l, []),
        "qexxc_dot_eri_dm_batched": (i, [i, p, p, i, i, i, i, p, p, p, l, vp]),
        "qexxc_dot_eri_dm_vjp_batched": (i, [i, p, p, p, i, i, p, p, l, vp]),
        "qexxc_generalized_eigh_batched": (i, [i, p, p, i, i, d, p, p, vp]),
        "qexxc_becke_partition": (i, [i, p, l, p, p, p, p, i, i, p, p, vp]),
        "qexxc_grid_launch_count": (l, []),
        "qexxc_lda_exchange": (i, [i, p, l, p, p, vp]),
        "qexxc_lda_launch_count": (l, []),
        "qexxc_comm_nccl_version": (i, [C.POINTER(i)]),
        "qexxc_comm_unique_id": (i, [p]),
        "qexxc_comm_create": (i, [C.POINTER(vp), i, i, i, p]),
        "qexxc_comm_wrap": (i, [C.POINTER(vp), vp, i, i, i]),
        "qexxc_comm_destroy": (i, [vp]),
        "qexxc_comm_rank": (i, [vp]),
        "qexxc_comm_world": (i, [vp]),
        "qexxc_comm_calls": (l, [vp]),
        "qexxc_allreduce": (i, [vp, p, l, vp]),
        "qexxc_bcast": (i, [vp, p, l, i, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().qexxc_last_error().decode()


def check(rc: int):
    if rc != 0:
        msg = last_error()
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError(msg)  # the reference raises NotImplementedError / ValueError here
        if rc == ERR_ARG:
            raise QexxcArgError(rc, msg)
        raise QexxcError(rc, msg)

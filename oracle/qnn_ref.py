"""Oracle: per-grid-point statevector QNN (stage 3, LocalQNN).  TEST INFRASTRUCTURE ONLY.

Restates ``QNN.__call__`` qedft/models/quantum/quantum_models.py:115-157 as used by
``LocalQNN`` (:384-494, via ``DirectQNN`` :312-330 and ``build_qnn.apply_fn`` :759-774):

    per point x:  |0...0>  --RY(x) on every qubit (``direct_gates`` feature_maps.py:184-221)-->
                  hea(n, L) (hardware_ansatz.py:88-147): per layer, for q = 0..n-1
                  RX(t), RY(t), RX(t) on q; then the CNOT ring
                  NOT(target=(i+1)%n, control=i), i = 0..n-1, repeated n times (:143-144)
                  --> sum_i <Z_i>  (``total_magnetization_ops`` measurement.py:140-167).

theta index = layer*3n + 3q + k in gate-list order (quantum_models.py:136).  Init
U(-0.1, 0.1) (:752-757).  No input normalisation.

horqrux ^0.9.2 is a third-party dependency absent from /root/reference; its published gate
definitions are restated: R_P(t) = cos(t/2) I - i sin(t/2) P; the state is a rank-n tensor of
shape (2,)*n with qubit i on axis i (so qubit 0 is the most significant bit of the flat
index); Z = diag(+1, -1).  These conventions are pinned by the reference's known-answer tests
(tests/test_measurements.py:31-61: |00> -> [1, 1]; after X(0) -> [-1, 1];
tests/test_quantum_measurement.py:46-59).  Circuit outputs: parity unpinned.

The derivative code here is deliberately *forward-mode* (one derivative circuit per parameter)
so that it is algorithmically independent from the adjoint-state method the CUDA kernel uses.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_I = np.eye(2, dtype=np.complex128)
_PAULI = {"X": _X, "Y": _Y, "Z": _Z}


@dataclass
class QNNSpec:
    n_qubits: int = 6
    n_layers: int = 2

    def n_params(self):
        return 3 * self.n_qubits * self.n_layers

    def dim(self):
        return 1 << self.n_qubits


def init_params(spec: QNNSpec, seed=0):
    return np.random.default_rng(seed).uniform(-0.1, 0.1, spec.n_params())


def rot(axis, t):
    """R_P(t) = cos(t/2) I - i sin(t/2) P; t scalar or [G] -> [2,2] or [G,2,2]."""
    t = np.asarray(t, dtype=np.float64)
    c = np.cos(t / 2)[..., None, None]
    s = np.sin(t / 2)[..., None, None]
    return c * _I - 1j * s * _PAULI[axis]


def apply_1q(state, U, q, n):
    """state [..., 2**n]; U [2,2] or [G,2,2] acting on qubit q (axis q of the (2,)*n tensor)."""
    lead = state.shape[:-1]
    s = state.reshape(lead + (1 << q, 2, 1 << (n - 1 - q)))
    if U.ndim == 2:
        out = np.einsum("ab,...ibj->...iaj", U, s)
    else:  # per-point gate, state [G, ..., dim] with G leading
        out = np.einsum("gab,g...ibj->g...iaj", U, s)
    return out.reshape(state.shape)


def cnot_perm(control, target, n):
    """index map: new_state[idx] = old_state[perm[idx]] for NOT(target, control)."""
    dim = 1 << n
    idx = np.arange(dim)
    cbit = (idx >> (n - 1 - control)) & 1
    return np.where(cbit == 1, idx ^ (1 << (n - 1 - target)), idx)


def ring_perm(n):
    """hardware_ansatz.py:143-144: the ring (i -> (i+1)%n), i=0..n-1, repeated n times, as one
    permutation: out[idx] = in[perm[idx]]."""
    perm = np.arange(1 << n)
    for _ in range(n):
        for i in range(n):
            p = cnot_perm(i, (i + 1) % n, n)
            # applying gate: new[idx] = old[p[idx]]; compose with accumulated perm
            perm = perm[p]
    return perm


def z_sum_diag(n):
    idx = np.arange(1 << n)
    pop = np.zeros_like(idx)
    for b in range(n):
        pop += (idx >> b) & 1
    return (n - 2 * pop).astype(np.float64)


def ansatz_gates(spec: QNNSpec):
    """list of ("RX"/"RY", qubit, theta_index) and ("RING",) in circuit order."""
    n, out, k = spec.n_qubits, [], 0
    for _ in range(spec.n_layers):
        for q in range(n):
            for ax in ("X", "Y", "X"):
                out.append((ax, q, k))
                k += 1
        out.append(("RING",))
    return out


def run_ansatz(spec, theta, state, dgate=None):
    """Apply hea to ``state`` [..., dim]; if dgate=k, gate k is replaced by dU/dtheta_k."""
    n = spec.n_qubits
    ring = ring_perm(n)
    for g in ansatz_gates(spec):
        if g[0] == "RING":
            state = state[..., ring]
        else:
            ax, q, k = g
            U = rot(ax, theta[k])
            if dgate == k:
                U = (-0.5j * _PAULI[ax]) @ U
            state = apply_1q(state, U, q, n)
    return state


def feature_states(spec, x):
    """phi(x) = prod_q RY_q(x)|0..0> and its first two x-derivatives, each [G, dim]."""
    n = spec.n_qubits
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    G = x.shape[0]
    phi = np.zeros((G, 1 << n), dtype=np.complex128)
    phi[:, 0] = 1.0
    U = rot("Y", x)
    for q in range(n):
        phi = apply_1q(phi, U, q, n)

    def D(s):  # sum_q (-i/2) Y_q s
        out = np.zeros_like(s)
        for q in range(n):
            out += apply_1q(s, -0.5j * _Y, q, n)
        return out

    d1 = D(phi)
    return phi, d1, D(d1)


def apply(spec, theta, x):
    """LocalQNN apply_fn: x [G] (or [G,1]) -> [G]."""
    phi, _, _ = feature_states(spec, x)
    psi = run_ansatz(spec, theta, phi)
    return np.einsum("gi,i,gi->g", psi.conj(), z_sum_diag(spec.n_qubits), psi).real


def exc_and_vrho_local(spec, theta, rho):
    """trainer_legacy_no_jit.py:56-63 with the QNN as network: exc [G], vrho = d exc/d rho [G]."""
    O = z_sum_diag(spec.n_qubits)
    phi, d1, _ = feature_states(spec, rho)
    st = run_ansatz(spec, theta, np.stack([phi, d1], axis=1))  # [G,2,dim]
    psi, dpsi = st[:, 0], st[:, 1]
    exc = np.einsum("gi,i,gi->g", psi.conj(), O, psi).real
    vrho = 2.0 * np.einsum("gi,i,gi->g", dpsi.conj(), O, psi).real
    return exc, vrho


def exc_and_vrho_local_vjp(spec, theta, rho, exc_bar, vrho_bar):
    """(exc_bar [G], vrho_bar [G]) -> (rho_bar [G], theta_bar [n_params]); forward mode."""
    O = z_sum_diag(spec.n_qubits)
    phi, d1, d2 = feature_states(spec, rho)
    init = np.stack([phi, d1, d2], axis=1)
    st = run_ansatz(spec, theta, init)
    psi, p1, p2 = st[:, 0], st[:, 1], st[:, 2]

    def ev(a, b):
        return np.einsum("gi,i,gi->g", a.conj(), O, b).real

    e1 = 2.0 * ev(p1, psi)
    e2 = 2.0 * ev(p2, psi) + 2.0 * ev(p1, p1)
    rho_bar = exc_bar * e1 + vrho_bar * e2
    tb = np.zeros(spec.n_params())
    for k in range(spec.n_params()):
        dk = run_ansatz(spec, theta, init[:, :2], dgate=k)
        dpsi, dp1 = dk[:, 0], dk[:, 1]
        de = 2.0 * ev(dpsi, psi)
        de1 = 2.0 * (ev(dp1, psi) + ev(p1, dpsi))
        tb[k] = np.dot(exc_bar, de) + np.dot(vrho_bar, de1)
    return rho_bar, tb


def apply_vjp(spec, theta, x, ybar):
    rb, tb = exc_and_vrho_local_vjp(spec, theta, x, np.asarray(ybar, dtype=np.float64), np.zeros(len(np.ravel(x))))
    return rb, tb


def per_qubit_z(state, n):
    """<Z_i> for i = 0..n-1 of one state vector (known-answer helper)."""
    idx = np.arange(1 << n)
    p = np.abs(state) ** 2
    return np.array([np.sum(p * (1 - 2 * ((idx >> (n - 1 - i)) & 1))) for i in range(n)])

"""Ozaki-style split of the FP64 contractions into exact INT8 products: how many 7-bit slices do
the c5 operands need for the north-star tolerance (1e-10 of the largest element)?

Emulates on the CPU, exactly (int64 arithmetic), what INT8 tensor-core GEMMs with INT32 accumulation would compute for
the stage-2 product T = ao . S and the row-dot rho = rowsum(T * ao):
    a[g,:] = 2^ea[g] * sum_s A_s[g,:] 2^(-7(s+1)),   S[:,j] = 2^eb[j] * sum_t B_t[:,j] 2^(-7(t+1)),  A_s, B_t in [-64, 63]
    T ~= 2^(ea+eb) * sum_{s+t<ns} A_s B_t 2^(-7(s+t+2))      (every A_s B_t exact: 12 bits + log2(N=1000) = 22 < 31)
Prints, per slice count, the number of INT8 GEMMs and the max-norm error of T and rho.
    python tests/studies/ozaki_study.py [ngrids]   (lives under tests/ because it uses the oracle's AO values)"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import gto_ref  # noqa: E402
from qex_b200 import workloads  # noqa: E402


def slices(x, axis, ns):
    """x = 2^e * sum_s X_s 2^(-7(s+1)) along `axis` scaling; X_s integer in [-64, 64]."""
    mx = np.abs(x).max(axis=axis, keepdims=True)
    e = np.where(mx > 0, np.ceil(np.log2(np.maximum(mx, 1e-300))) + 1, 0.0)
    r = x / np.exp2(e)            # |r| < 0.5
    out = []
    for s in range(ns):
        r = r * 128.0
        q = np.rint(r)
        out.append(q.astype(np.int64))
        r = r - q
    return e, out


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    wl = workloads.make("c5", ngrids=G)
    m = wl.mol
    ao = gto_ref.eval_ao(m._atm, m._bas, m._env, wl.coords, 0)
    S = 0.5 * (wl.dm + wl.dm.T)
    T = ao @ S
    rho = (T * ao).sum(1)
    rows = []
    for ns in range(3, 8):
        ea, A = slices(ao, 1, ns)
        eb, B = slices(S, 0, ns)
        acc = np.zeros_like(T)
        ngemm = 0
        for d in range(ns):  # diagonals s + t = d share one INT32 accumulator
            part = np.zeros(T.shape, dtype=np.int64)
            for s in range(d + 1):
                part += A[s] @ B[d - s]
                ngemm += 1
            assert np.abs(part).max() < 2**31
            acc += part.astype(np.float64) * 2.0 ** (-7 * (d + 2))
        Tz = acc * np.exp2(ea) * np.exp2(eb)
        rz = (Tz * ao).sum(1)
        rows.append(dict(slices=ns, int8_gemms=ngemm, T_rel_err=float(np.abs(Tz - T).max() / np.abs(T).max()),
                         rho_rel_err=float(np.abs(rz - rho).max() / np.abs(rho).max())))
        print(rows[-1])
    return dict(workload=f"c5 operands: ao [{G} x {wl.nao}] (oracle AO values), S = sym(dm) [{wl.nao} x {wl.nao}]", rows=rows)


if __name__ == "__main__":
    res = main()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "profiles", "r02", "ozaki_slices.json")
    json.dump(res, open(out, "w"), indent=1)

"""Incore Coulomb / exchange build (reference hf_legacy.py:275-299).
CPU: the oracle restatement against explicit loops and its own adjoint identity; the host mirror's
argument handling.  GPU: csrc/jk.cu through the C ABI against the oracle, bit-reproducibility,
and size-independent properties at the c3 size (120 AOs, 1.66 GB tensor)."""
import numpy as np
import pytest

from oracle import jk_ref
from tests._util import rel_err


def test_oracle_einsums_match_explicit_loops():
    n = 3
    rng = np.random.default_rng(0)
    eri = rng.standard_normal((n,) * 4)
    dm = rng.standard_normal((2, n, n))
    vj, vk = jk_ref.dot_eri_dm(eri, dm)
    ej, ek = np.zeros_like(dm), np.zeros_like(dm)
    for x in range(2):
        for i in range(n):
            for j in range(n):
                for k in range(n):
                    for l in range(n):
                        ej[x, k, l] += eri[i, j, k, l] * dm[x, j, i]
                        ek[x, i, l] += eri[i, j, k, l] * dm[x, j, k]
    assert np.allclose(vj, ej, atol=1e-13) and np.allclose(vk, ek, atol=1e-13)
    # flat / 2-D inputs keep the shape of dm, missing outputs are None (reference :279-286)
    vj1, vk1 = jk_ref.dot_eri_dm(eri.ravel(), dm[0], with_k=False)
    assert vj1.shape == (n, n) and vk1 is None and np.allclose(vj1, ej[0])
    with pytest.raises(NotImplementedError):
        jk_ref.dot_eri_dm(eri.ravel()[:20], dm[0])


def test_oracle_vjp_is_the_adjoint():
    n = 5
    rng = np.random.default_rng(1)
    eri = jk_ref.synthetic_eri(n, 2, symmetric=False)
    dm, a, b = (rng.standard_normal((2, n, n)) for _ in range(3))
    vj, vk = jk_ref.dot_eri_dm(eri, dm)
    bar = jk_ref.dot_eri_dm_s1_vjp(eri, a, b)
    assert abs(np.sum(vj * a) + np.sum(vk * b) - np.sum(bar * dm)) < 1e-11
    # physical symmetry: for a symmetric tensor the J adjoint is J itself
    es = jk_ref.synthetic_eri(n, 3)
    assert np.allclose(es, es.transpose(1, 0, 2, 3)) and np.allclose(es, es.transpose(2, 3, 0, 1))
    s = a[0] + a[0].T
    assert np.allclose(jk_ref.dot_eri_dm_s1_vjp(es, s, None), jk_ref.dot_eri_dm(es, s)[0], atol=1e-12)
    assert np.isclose(jk_ref.energy_coulomb(s, jk_ref.dot_eri_dm(es, s)[0]), 0.5 * np.einsum("ijkl,ji,lk", es, s, s))


def test_host_mirror_rejects_what_the_reference_does_not_accelerate():
    from qex_b200 import hf

    with pytest.raises(NotImplementedError):
        hf.dot_eri_dm(np.zeros(55), np.zeros((4, 4)))  # s8-packed tensor -> pyscfad _vhf.incore in the reference
    with pytest.raises(NotImplementedError):
        hf.dot_eri_dm(np.zeros(16, dtype=complex), np.zeros((2, 2)))
    c = np.array([[1.0, 0.5], [0.2, -1.0]])
    dm = hf.make_rdm1(c, np.array([2.0, 0.0])).numpy()
    assert np.allclose(dm, 2.0 * np.outer(c[:, 0], c[:, 0]))


# ------------------------------------------------------------------ GPU parity (through the C ABI)
@pytest.mark.gpu
@pytest.mark.parametrize("nao,nset", [(1, 1), (4, 1), (7, 2), (30, 1), (65, 1), (92, 3)])
@pytest.mark.parametrize("which", ["jk", "j", "k"])
def test_cuda_jk_matches_oracle(nao, nset, which):
    from qex_b200 import hf

    rng = np.random.default_rng(nao)
    eri = jk_ref.synthetic_eri(nao, nao + 1, symmetric=False)  # no symmetry: catches index transpositions
    dm = rng.standard_normal((nset, nao, nao)) if nset > 1 else rng.standard_normal((nao, nao))
    wj, wk = "j" in which, "k" in which
    rj, rk = jk_ref.dot_eri_dm(eri, dm, with_j=wj, with_k=wk)
    vj, vk = hf.dot_eri_dm(eri, dm, with_j=wj, with_k=wk)
    assert (vj is None) == (not wj) and (vk is None) == (not wk)
    if wj:
        assert vj.shape == dm.shape and rel_err(vj.cpu().numpy(), rj) <= 1e-10
    if wk:
        assert vk.shape == dm.shape and rel_err(vk.cpu().numpy(), rk) <= 1e-10
    a = rng.standard_normal(dm.shape) if wj else None
    b = rng.standard_normal(dm.shape) if wk else None
    bar = hf._dot_eri_dm_s1_vjp(eri, nao, a, b)
    assert rel_err(bar.cpu().numpy(), jk_ref.dot_eri_dm_s1_vjp(eri, a, b)) <= 1e-10


@pytest.mark.gpu
def test_cuda_jk_autograd_and_reproducibility():
    import torch

    from qex_b200 import hf

    nao = 40
    eri = torch.as_tensor(jk_ref.synthetic_eri(nao, 5)).cuda()
    dm = torch.randn(nao, nao, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    dm = (dm + dm.T).requires_grad_(True)
    vj, vk = hf.dot_eri_dm_autograd(eri, dm)
    # E = 0.5 sum dm (J - 0.5 K): for a symmetric tensor and symmetric dm, dE/ddm = J - 0.5 K
    e = hf.energy_coulomb(dm, vj) - 0.25 * torch.sum(dm * vk.T)
    (g,) = torch.autograd.grad(e, dm)
    want = (vj - 0.5 * vk).detach()
    assert rel_err(g.cpu().numpy(), want.T.cpu().numpy()) <= 1e-10
    n0 = hf.jk_launch_count()
    vj2, vk2 = hf.dot_eri_dm(eri, dm.detach())
    assert hf.jk_launch_count() - n0 == 2  # streaming kernel + one finishing reduction
    assert torch.equal(vj2, vj.detach()) and torch.equal(vk2, vk.detach())  # no atomics: bit-reproducible


@pytest.mark.gpu
def test_cuda_jk_c3_size_properties():
    """120 AOs (BASELINE c3): the oracle would need the 1.66 GB tensor on the host, so use properties:
    a Gram-form tensor eri = L^T L gives J = L^T (L . dm^T) and K by a GEMM chain in torch fp64."""
    import torch

    from qex_b200 import hf

    nao, naux = 120, 96
    g = torch.Generator("cuda").manual_seed(7)
    L = torch.randn(naux, nao, nao, dtype=torch.float64, device="cuda", generator=g) / naux**0.5
    eri = (L.reshape(naux, -1).T @ L.reshape(naux, -1)).reshape((nao,) * 4)
    dm = torch.randn(nao, nao, dtype=torch.float64, device="cuda", generator=g)
    vj, vk = hf.dot_eri_dm(eri, dm)
    wj = torch.einsum("Pkl,P->kl", L, torch.einsum("Pij,ji->P", L, dm))
    wk = torch.einsum("Pij,Pjl->il", L, torch.einsum("Pkl,jk->Pjl", L, dm))
    assert rel_err(vj.cpu().numpy(), wj.cpu().numpy()) <= 1e-10
    assert rel_err(vk.cpu().numpy(), wk.cpu().numpy()) <= 1e-10
    # linearity in dm and the adjoint identity <vj,a> + <vk,b> = <dm, vjp(a,b)>
    a = torch.randn(nao, nao, dtype=torch.float64, device="cuda", generator=g)
    b = torch.randn(nao, nao, dtype=torch.float64, device="cuda", generator=g)
    bar = hf._dot_eri_dm_s1_vjp(eri, nao, a, b)
    lhs = (vj * a).sum() + (vk * b).sum()
    rhs = (bar * dm).sum()
    assert abs(lhs - rhs).item() <= 1e-10 * abs(lhs).item()
    vj2, _ = hf.dot_eri_dm(eri, 2.0 * dm + a, with_k=False)
    vja, _ = hf.dot_eri_dm(eri, a, with_k=False)
    assert rel_err(vj2.cpu().numpy(), (2.0 * vj + vja).cpu().numpy()) <= 1e-12

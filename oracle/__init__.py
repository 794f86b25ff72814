"""CPU oracle for the QEX 3D XC grid-integration hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``qex_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do.  It is a NumPy float64 restatement of the reference's
pure-Python/JAX algorithm (pasqal-io/qex, package ``qedft``), each function citing
the reference file:line it follows.

PARITY PIN STATUS
-----------------
The reference cannot be executed in this environment (jax, pyscf, pyscfad, horqrux,
flax are not installable: no network, not in the wheelhouse), and its own tests hold
no golden vectors for this path.  Its notebooks, however, freeze
outputs of the path, and those ARE reproduced.  What is pinned:

* the whole grid -> AO -> rho -> E_xc / V_xc -> SCF chain (``grid_ref``, ``gto_ref`` s shells,
  ``numint_ref`` conventions, ``scf_ref``): PINNED by the reference notebook's LDA-exchange RKS
  energy of H2/6-31G on the level-0 Stratmann grid the trainers use
  (notebooks/04_notebook_td_trainer.ipynb cell 1: "converged SCF energy = -1.03718794786902",
  1240 grid points) -- the oracle gives -1.0371879478690555 (|diff| < 1e-13 Ha), the CUDA
  kernels through ``NumInt.nr_rks`` < 1e-9 Ha -- and by the three tail densities the same
  notebook prints for its CCSD density matrix (1e-8 relative, and in pyscf's point order with
  ``grid_ref.build(sort_grids=True)``); tests/test_zz_pyscf_pin.py; the extracted numbers live in
  tests/golden/reference_notebook.json; ``train_ref`` restates the trainer's loss on top of this chain;

* contraction / assembly conventions (``numint_ref``): pinned by the closed-form toy
  functional of ``tests/test_numint.py:96-103`` (exc = 0.01 rho^2, vrho = 0.02 rho), by
  the einsum restatement ``scf_functions_masked.py:143-159`` and by finite differences;
* MLP / second-order VJP (``mlp_ref``): pinned by torch float64 double-autograd and
  central finite differences (semantics of stax ``Dense``/``Tanh``/``Gelu`` are standard);
* statevector circuit (``qnn_ref``): conventions pinned by the reference's known-answer
  tests (``tests/test_measurements.py:31-61``, ``tests/test_quantum_measurement.py:46-59``);
  gate matrices are the published horqrux 0.9.2 definitions -> circuit outputs
  "parity unpinned";
* AO evaluation (``gto_ref``): restates pyscf 2.9 ``GTOval_sph_deriv0/1`` (third-party C,
  absent from /root/reference); pinned by orthonormality (numerical overlap integrals =
  identity for normalised shells) and, for the H2/6-31G s shells, by agreement of the grid
  overlap with the closed-form overlap of ``ints_ref`` (itself pinned below) -> values of
  general shells "parity unpinned";
* J/K contraction, generalised eigensolver, DIIS, SCF loop, 6-31G tables and s-type integrals
  (``jk_ref``, ``scf_ref``, ``ints_ref``, ``qex_b200/gto.py``): PINNED TOGETHER by six converged RHF
  energies of H2/6-31G that the reference's own notebook prints
  (notebooks/04_notebook_td_trainer.ipynb cells 1 and 5: 0.74/0.5/1.5/0.6/0.9/1.2 Angstrom),
  reproduced to < 1e-12 Ha by the oracle and to < 1e-10 Ha by the CUDA path (tests/test_scf.py).
"""

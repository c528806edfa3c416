"""CPU oracle for the truncated-EM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in vectorised NumPy float64, what the reference's
`select_Hprimes / E_step / M_step` compute for each sparse-coding model
(prosper/em/camodels/*_et.py).  Every function cites the reference file:line it
follows.  The formulation is deliberately the *direct* one (explicit W-bar
reconstructions, explicit squared errors) -- not the Gram/score-GEMM algebra the
CUDA path uses -- so that it is an independent check.

Who may import it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, always as the checker or the reported CPU
baseline.  Nothing under `prosper_b200/` imports it; the product path raises if
its CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, run in the build container through `oracle/ref_harness.py`; the minted
vectors and the script that made them live in `tests/golden/`.
"""

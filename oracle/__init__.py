"""CPU oracle for the dominant-eigenpair hot path.  TEST INFRASTRUCTURE ONLY.

This package is a from-scratch CPU restatement (numpy for the integer/bit work,
torch-CPU fp64 for the vector loops) of the algorithm the reference
buwantaiji/DominantSparseEigenAD runs on the path SURVEY.md section 8 names.
Every function cites the reference file:line it restates.

It exists to CHECK the CUDA product, never to be it.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  Nothing under `dominantsparseeigenad_b200/` imports it, and
the product raises if its CUDA library is missing instead of falling back here.

Parity status: PINNED.  `oracle/gen_golden.py` (run in the build container where
`/root/reference` is mounted) imports the unmodified reference with three
compatibility shims and (1) asserts this restatement reproduces it bit-for-bit
on the integer tables and to <=1e-12 on identical injected start vectors, and
(2) writes the reference's outputs to `tests/golden/*.npz`.  `tests/test_oracle.py`
re-checks the oracle against those committed vectors and against the upstream
result files `examples/TFIM/datas/*.npz` (copied as data into the golden set).
"""
from .dsea_oracle import *  # noqa: F401,F403

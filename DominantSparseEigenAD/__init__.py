"""Import alias: lets scripts written against buwantaiji/DominantSparseEigenAD run unedited on this library.

    import DominantSparseEigenAD.symeig as symeig             # examples/TFIM/E0.py:59, chiF.py:45
    from DominantSparseEigenAD.symeig import DominantSymeig    # examples/TFIM/E0.py:44, schrodinger1D.py:56
    from DominantSparseEigenAD.Lanczos import symeigLanczos    # tests/test_Lanczos.py
    from DominantSparseEigenAD.CG import CG_torch, CGSubspace  # tests/test_CG.py
    from DominantSparseEigenAD.eig import DominantEig          # examples/TFIM_vumps/general.py:47

Each name resolves to the SAME module object as `dominantsparseeigenad_b200.<name>`, so the reference's
module-global protocol (`symeig.setDominantSparseSymeig(...)` rebinding `symeig.DominantSparseSymeig`,
symeig.py:66,87) keeps working through either spelling.  The compute path is libdsea.so (CUDA, sm_100a);
there is no CPU fallback behind this alias either.
"""
import sys as _sys

import dominantsparseeigenad_b200 as _impl
from dominantsparseeigenad_b200 import CG, Lanczos, eig, symeig  # noqa: F401

for _name in ("symeig", "CG", "Lanczos", "eig"):
    _sys.modules[__name__ + "." + _name] = getattr(_impl, _name)
del _name

"""dominantsparseeigenad_b200 — B200-native (sm_100a) dominant-eigenpair solver with reverse-mode AD.

Drop-in for the hot path of buwantaiji/DominantSparseEigenAD: the modules `symeig`, `CG`, `Lanczos`
(and `eig`) keep the reference's names and signatures; the arithmetic runs in hand-written CUDA behind
the C ABI of include/dsea.h (libdsea.so).  There is no CPU fallback.
"""
from . import _lib, runtime                                        # noqa: F401
from . import CG, Lanczos, eig, symeig                             # noqa: F401
from .operators import (CallbackOperator, DenseOperator, SparseMatrixOperator, TFIM, dot,  # noqa: F401
                        project, scale)

__version__ = "0.1.0"

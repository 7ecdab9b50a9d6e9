"""admmsolver_b200 -- B200-native ADMM engine behind the API of SpM-lab/admmsolver.

Sub-modules mirror the reference package (``/root/reference/src/admmsolver``):
``matrix``, ``objectivefunc``, ``optimizer``, ``util``; ``batch`` exposes the two fused CUDA
engines directly for large batches; ``problems`` holds the synthetic BASELINE configurations.
Importing the package loads ``libadmm_b200.so`` (built by ``__graft_entry__.build()``); there is
no CPU fallback.
"""
__license__ = "MIT"
__version__ = "0.7.6+b200.1"

from . import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)

"""Drop-in shim: put ``<repo>/compat`` (and ``<repo>``) on ``PYTHONPATH`` and existing code that does
``from admmsolver.optimizer import SimpleOptimizer, Problem`` etc. runs on the B200 engine unchanged."""
from admmsolver_b200 import __license__, __version__  # noqa: F401

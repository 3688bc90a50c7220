"""PYTHONPATH shim: `import operators` (reference src/operators.py) -> ppnn_b200.operators."""
from ppnn_b200.operators import *  # noqa: F401,F403
from ppnn_b200.operators import __all__  # noqa: F401

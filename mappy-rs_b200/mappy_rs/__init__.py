"""mappy_rs -- drop-in mirror of the reference's Python module (host side above the C ABI).

Placeholder import surface; the full `Aligner` mirror lives in aligner.py.
"""
from . import _mmg  # noqa: F401
try:
    from .aligner import Aligner, Mapping  # noqa: F401
except ImportError:  # aligner.py arrives with the API milestone
    pass

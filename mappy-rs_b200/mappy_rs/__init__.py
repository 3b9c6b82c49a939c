"""mappy_rs -- drop-in mirror of the reference's Python module (Adoni5/mappy-rs src/lib.rs:995-999
exports `Aligner`), hosted above the C ABI of the B200 mapping library (include/mmg.h, libmmg.so)."""
from . import _mmg  # noqa: F401
from .aligner import Aligner, Mapping, AlignmentBatchResultIter, BatchMappings  # noqa: F401

__all__ = ["Aligner", "Mapping"]

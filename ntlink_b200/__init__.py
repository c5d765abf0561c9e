"""ntlink_b200 -- B200-native (sm_100a) minimizer sketching + minimizer mapping for ntLink.

The CUDA kernels live in ntlink_b200/csrc and are reached through the C ABI of include/ntlink_b200.h
(libntlink_b200.so, built in-tree by ntlink_b200/build.py). There is no CPU fallback."""
from .api import Context, MapResult, SeqBatch, Sketch, name_ranks, read_sequences  # noqa: F401
from ._lib import NtlError  # noqa: F401

__version__ = "0.1.0"

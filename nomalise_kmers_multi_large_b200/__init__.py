"""B200-native k-mer coverage normalisation: drop-in for the hot path of normalise_kmers_multi_large.

The product is the C-ABI shared library ``csrc/libnk_b200.so`` (sm_100a kernels + C host pipeline) and the
command-line program ``csrc/normalise_kmers_multi_large_b200``; this package only binds them (ctypes).
"""
from . import capi
from .capi import Engine, NkError, load_library
from .pipeline import Pipeline, count_chunk_lines, plan_ranges, run_cli

__all__ = ["capi", "Engine", "Pipeline", "NkError", "load_library", "plan_ranges", "count_chunk_lines", "run_cli"]

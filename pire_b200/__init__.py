"""pire_b200 -- B200-native implementation of Pire's inner DFA scan path.

Only what the path needs lives here:
    csrc/         sm_100a CUDA kernels + the extern "C" boundary (include/pire_b200.h)
    _native.py    ctypes binding of that boundary (fails loudly if the .so is missing)
    scanner.py    Python mirror of Pire's Scanner / Runner / Matches for batches
    workloads.py  the BASELINE.json pattern sets and synthetic corpora
    dist.py       shard-by-string + the one bitmap all-reduce
"""
from ._native import PireGpuError, RUN_BEGIN, RUN_END, VARIANT_AUTO, VARIANT_PLAIN, VARIANT_PRED, VARIANT_PRIV, VARIANT_LOOK  # noqa: F401
from .scanner import (Batch, BeginMark, EndMark, HalfFinalCount, HalfFinalResult, LongestPrefix, LongestSuffix, Matches, RunHelper, Runner, Scanner,  # noqa: F401
                      ShortestPrefix, ShortestSuffix)

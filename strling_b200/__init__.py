"""strling_b200: B200 (sm_100a) implementation of STRling's data-parallel hot path -- the per-read repeat-unit
scan (`strling extract`, utils.nim:236) and STR-read clustering (`strling call` / `merge`, cluster.nim:364) --
behind a C ABI (include/strgpu.h, strling_b200/libstrgpu.so).  This package is the thin Python mirror of that
ABI used by the tests and bench.py; there is no CPU fallback: without the built CUDA library it raises."""
from .binding import (  # noqa: F401
    BOUNDS_DTYPE,
    Masks,
    REPEAT_DTYPE,
    TREAD_DTYPE,
    SEGMENT_DTYPE,
    StrGpu,
    StrGpuError,
    load_library,
    pack_reads,
)

__version__ = "0.1.0"

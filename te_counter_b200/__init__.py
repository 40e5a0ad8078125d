"""
te_counter_b200 -- B200-native counting hot path of te_counter (read <-> gene/TE annotation overlap
and tally, bulk SE/PE and 10x-style single cell) behind the reference's own `measureTE` API.

    from te_counter_b200 import measureTE        # same surface as te_count.measureTE

The compute path is hand-written CUDA for sm_100a in te_counter_b200/csrc, loaded through the
C ABI of include/tecount.h.  Nothing here falls back to the CPU.
"""
from .te_count import measureTE, ScResult      # noqa: F401
from . import index, reads                       # noqa: F401

__all__ = ["measureTE", "ScResult", "index", "reads"]

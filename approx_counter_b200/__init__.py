"""approx_counter_b200 — B200 (sm_100a) implementation of approx_counter's
approximate k-mer counting path.  CUDA kernels and the C ABI live in csrc/
(libapc.so); this package is the ctypes binding used by tests and bench.py.
"""
from ._lib import ApcError, LIB_PATH, load  # noqa: F401
from .api import ApproxCounter, device_count, plan_queries  # noqa: F401
from .sharded import ShardedApproxCounter, allreduce_counts, shard_bounds  # noqa: F401
from . import host  # noqa: F401

__all__ = ["ApproxCounter", "ShardedApproxCounter", "ApcError", "device_count", "load", "LIB_PATH",
           "allreduce_counts", "shard_bounds", "host", "plan_queries"]

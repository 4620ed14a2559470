"""plumed2_b200 -- B200-native COORDINATION + neighbour-list engine behind PLUMED's action API.

  capi           ctypes binding of the C ABI (include/b200coord.h, lib/libb200coord.so)
  coordination   host-side mirror of the reference COORDINATION action (keywords, prepare/calculate)
  csrc/          CUDA kernels (sm_100a), the C ABI and the PLUMED plugin action

There is no CPU implementation in this package; the oracle lives in /oracle and is test infrastructure.
"""
from . import capi  # noqa: F401
from .coordination import Coordination, PlumedInputError, comm_unique_id, parse_atom_list, shard_range  # noqa: F401

__all__ = ["capi", "Coordination", "PlumedInputError", "comm_unique_id", "parse_atom_list", "shard_range"]

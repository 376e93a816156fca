"""wignerd.jl_b200 -- B200-native (sm_100a) Wigner-d/D engine behind the WignerD.jl API.

    from wignerd.jl_b200 import wignerd, wignerD, wignerdjmn, wignerDjmn, Equator, NorthPole, SouthPole

``csrc/`` holds the CUDA kernels and the C ABI (``include/wignerd_b200.h``); ``api.py`` is the host-side
mirror of the reference's Julia functions over that ABI; ``shard.py`` splits angle batches over ranks.
"""
from .api import (ABI_SYMBOLS, ArgumentError, CudaError, Engine, Equator, InexactError, NorthPole, SouthPole,  # noqa: F401
                  SpecialCoLatitudes, WignerMatrix, _offsetmatrix, _sincos, default_engine, load_library,
                  twoj_of, wignerD, wignerD_, wignerd, wignerd_, wignerdjmn, wignerDjmn,
                  GENERIC, NORTH, SOUTH, EQUATOR, MEM_HOST, MEM_DEVICE, LIB_PATH)
from .shard import shard_bounds, shard_bounds_native, symmetric_shard, gather_slabs  # noqa: F401

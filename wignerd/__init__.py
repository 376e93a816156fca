"""Namespace package: the engine lives in ``wignerd.jl_b200`` (also reachable as ./wignerd.jl_b200/)."""

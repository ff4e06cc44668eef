"""exadem_b200 — B200-native (sm_100a) DEM contact hot path behind exaDEM's operator interface.

Only what the hot path needs lives here: csrc/ (CUDA kernels + C ABI), host/ (C++ operators with the reference's
names/slots), capi.py (ctypes view of the ABI), workloads.py (seeded synthetic inputs), parallel.py (brick
decomposition + NCCL bootstrap).  The CPU oracle is NOT part of this package (see oracle/).
"""
from .capi import (Config, Counters, EdbError, Engine, Library, ParticleArrays, make_config, make_shape,  # noqa: F401
                   MODE_SPHERE, MODE_POLYHEDRON, LAW_NONE, LAW_COHESIVE, LAW_DMT, DRIVER_CYLINDER, DRIVER_SURFACE,
                   DRIVER_BALL, DRIVER_RSHAPE, NTYPES, INTERACTION_DTYPE)

__version__ = "0.1.0"

"""Test-only loader of the CPU oracle (oracle/liboracle_dem.so) and of the compiled reference math
(oracle/_ref/libexadem_refmath.so).  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs ONLY — never from the exadem_b200 package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

from exadem_b200 import capi

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"


def build_oracle(force=False):
    so = ORACLE_DIR / "liboracle_dem.so"
    if force or not so.exists() or so.stat().st_mtime < max((ORACLE_DIR / f).stat().st_mtime
                                                              for f in ("dem_oracle.cpp", "dem_math.hpp")):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "liboracle_dem.so"], check=True, capture_output=True)
    return so


class OracleLibrary:
    prefix = "orc_"

    def __init__(self):
        self.path = build_oracle()
        self.dll = C.CDLL(str(self.path))

    def fn(self, name):
        f = getattr(self.dll, self.prefix + name)
        f.restype = C.c_int
        return f


class OracleEngine(capi.Engine):
    """Same scripting surface as the product Engine, running the CPU restatement."""

    def set_threads(self, n):
        self._call("set_threads", C.c_int(n))

    def set_virtual_bricks(self, pgrid):
        """Multi-rank emulation: treat cells of another brick like the ghost cells they would be on the owning rank."""
        self._call("set_virtual_bricks", (C.c_int32 * 3)(*pgrid))

    def set_virtual_cuts(self, cell_cuts):
        """Virtual bricks bounded by cell-index cuts per axis (after set_virtual_bricks): emulates Engine.set_brick_cuts."""
        arr = [None if c is None else (C.c_int32 * len(c))(*[int(x) for x in c]) for c in cell_cuts]
        self._call("set_virtual_cuts", *arr)

    def set_particle_clusters(self, ids, cluster):
        """field::cluster per particle id (oracle-only input of stick_polyhedra)."""
        import numpy as np
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
        cl = np.ascontiguousarray(cluster, dtype=np.float64)
        self._call("set_particle_clusters", C.c_uint64(len(ids)), ids.ctypes.data_as(C.c_void_p), cl.ctypes.data_as(C.c_void_p))

    def stick_polyhedra(self, threshold):
        self._call("stick_polyhedra", C.c_double(threshold))

    def stick_polyhedra_ids(self, threshold, ids, cluster):
        """same call shape as the product's edb_stick_polyhedra"""
        self.set_particle_clusters(ids, cluster)
        self.stick_polyhedra(threshold)

    def sync(self):
        pass

    def set_timing(self, on):
        pass

    def flush_l2(self):
        pass


def ref_math():
    """ctypes handle on the reference's own header-only math (None when oracle/_ref was not built)."""
    so = ORACLE_DIR / "_ref" / "libexadem_refmath.so"
    if not so.exists():
        return None
    return C.CDLL(str(so))

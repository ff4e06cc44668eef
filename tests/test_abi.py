"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/exadem_b200.h
declares, struct layouts agree with the ctypes mirror, and the product path fails loudly without a CUDA device."""
import ctypes as C
import re
from pathlib import Path

import pytest

from exadem_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / "include" / "exadem_b200.h").read_text()
    return sorted(set(re.findall(r"\b(edb_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_the_operator_surface():
    syms = declared_symbols()
    for need in ("edb_update_particle_neighbors", "edb_contact_sphere", "edb_contact_polyhedron", "edb_combined_compute_prolog",
                 "edb_combined_compute_epilog", "edb_ghost_update_rq", "edb_compute_vertices", "edb_run_verlet",
                 "edb_export_interactions", "edb_comm_init", "edb_set_driver_motions", "edb_push_f_v_r_driver", "edb_push_f_v_driver",
                 "edb_push_av_to_quat_driver", "edb_reset_force_driver", "edb_force_to_accel_driver", "edb_driver_vertices",
                 "edb_grid_rshape_driver", "edb_backup_drivers", "edb_driver_displ_over"):
        assert need in syms


def test_library_exports_every_declared_symbol(product_lib):
    missing = [s for s in declared_symbols() if not hasattr(product_lib.dll, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_struct_layouts():
    assert C.sizeof(capi.Interaction) == 104            # Interaction = 56 B pair + 48 B payload
    assert capi.Interaction.id_j.offset == 24 and capi.Interaction.type.offset == 48
    assert capi.Interaction.friction.offset == 56
    assert C.sizeof(capi.ContactParams) == 64
    assert C.sizeof(capi.Particles) == 8 + 31 * 8
    assert C.sizeof(capi.Config) == 13 * 8 + 12 * 4
    # edb_driver / edb_driver_motion (static_assert'ed to the same numbers in csrc/edb_lib.cu)
    assert C.sizeof(capi.Driver) == 184 and C.sizeof(capi.DriverMotion) == 400
    assert capi.DriverMotion.tab_time.offset == 80 and capi.DriverMotion.mass.offset == 224
    assert capi.INNER_BOND_DTYPE.itemsize == 128 and C.sizeof(capi.InnerBondParams) == 48   # edb_inner_bond, edb_inner_bond_params


def test_no_cpu_fallback(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = capi.make_config((0, 0, 0), (10, 10, 10), 2.5)
    with pytest.raises(capi.EdbError) as e:
        capi.Engine(product_lib, cfg)
    assert "no CUDA device" in str(e.value) or "rc=-2" in str(e.value)


def test_product_package_does_not_reference_the_oracle():
    for p in (ROOT / "exadem_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"):
            txt = p.read_text()
            for needle in ("liboracle", "oracle_binding", "dem_oracle", "orc_", "oracle/_ref", "refmath"):
                assert needle not in txt, (p, needle)

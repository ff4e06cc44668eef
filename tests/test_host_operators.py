"""The host-side mirror of the reference's operator interface (exadem_b200/host/operators.cpp -> libexadem_host_ops.so): the
operators are registered under the reference's YAML names with the reference's slot names, and a Verlet step
driven by NAME through the registry — what the .msp interpreter does with `numerical_scheme_verlet_macro`
(data/config/config_numerical_schemes.msp:6-30) — equals the direct C-ABI path bit for bit."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from exadem_b200 import Engine, capi, workloads

ROOT = Path(__file__).resolve().parent.parent
HOSTLIB = ROOT / "exadem_b200" / "libexadem_host_ops.so"

# YAML operator names of SURVEY.md 2b that a reference deck may call on the hot path
EXPECTED = """chunk_neighbors_contact update_traversals nbh_sphere nbh_polyhedron compute_vertices update_interaction_ghost
classify_interactions unclassify_interactions compress_interaction contact_sphere contact_singlemat_sphere
contact_singlemat_sphere_hooke contact_singlemat_sphere_hooke_none contact_singlemat_sphere_hooke_cohesive
contact_singlemat_sphere_hooke_dmt contact_multimat_sphere_hooke_none contact_multimat_sphere_hooke_cohesive
contact_multimat_sphere_hooke_dmt contact_polyhedron contact_singlemat_polyhedron contact_singlemat_polyhedron_hooke_none
contact_singlemat_polyhedron_hooke_cohesive contact_singlemat_polyhedron_hooke_dmt contact_multimat_polyhedron_hooke_none
contact_multimat_polyhedron_hooke_cohesive contact_multimat_polyhedron_hooke_dmt gravity_force reset_force_moment force_to_accel
combined_compute_prolog combined_compute_epilog push_to_quaternion push_to_angular_velocity push_to_angular_acceleration
freeze_particles pas_libere_pas_delivre ghost_update_rq vertex_displ_over particle_displ_over backup_dem
migrate_cell_particles_interaction move_particles_interaction compute_driver_force_moment stress_tensor simulation_state
push_f_v_r_driver push_f_v_driver push_av_to_quat_driver reset_force_driver force_to_accel_driver driver_vertices grid_rshape_driver
backup_drivers driver_displ_over stick_polyhedra inner_bond_polyhedron update_interfaces apply_interface_fracture_criterion
broken_interface_friction_updater""".split()


def host_lib(product_lib):
    assert HOSTLIB.exists(), f"{HOSTLIB} missing: run __graft_entry__.build()"
    h = C.CDLL(str(HOSTLIB), mode=C.RTLD_GLOBAL)
    h.edbh_graph_create.restype = C.c_void_p
    h.edbh_graph_create.argtypes = [C.c_void_p]
    h.edbh_last_error.restype = C.c_char_p
    h.edbh_last_error.argtypes = [C.c_void_p]
    h.edbh_get_doubles.restype = C.c_longlong
    for f in ("edbh_graph_destroy", "edbh_run", "edbh_set_double", "edbh_set_bool", "edbh_set_vec3", "edbh_set_contact_params",
              "edbh_get_bool", "edbh_get_doubles"):
        getattr(h, f).argtypes = None
    return h


def test_registry_has_the_reference_operator_names(product_lib):
    h = host_lib(product_lib)
    buf = C.create_string_buffer(1 << 16)
    n = h.edbh_list_operators(buf, len(buf))
    assert n > 0
    names = set(buf.value.decode().split())
    missing = [e for e in EXPECTED if e not in names]
    assert not missing, missing
    doc = C.create_string_buffer(4096)
    assert h.edbh_documentation(b"contact_sphere", doc, len(doc)) > 0 and b"contact law" in doc.value


@pytest.mark.gpu
def test_verlet_step_by_operator_name_equals_direct_path(product_lib):
    h = host_lib(product_lib)
    wl = workloads.rotating_drum(5000, jitter=0.02)
    direct = wl.setup(Engine(product_lib, wl.config()))
    named = wl.setup(Engine(product_lib, wl.config()))
    g = C.c_void_p(h.edbh_graph_create(named.h))

    def run(op):
        rc = h.edbh_run(g, op.encode())
        assert rc == 0, (op, h.edbh_last_error(g))

    h.edbh_set_double(g, b"dt", C.c_double(wl.dt))
    h.edbh_set_double(g, b"rcut_inc", C.c_double(wl.rcut_inc))
    h.edbh_set_double(g, b"threshold", C.c_double(0.5 * wl.rcut_inc))
    h.edbh_set_bool(g, b"symetric", 1)
    h.edbh_set_vec3(g, b"gravity", (C.c_double * 3)(*wl.gravity))
    cp = capi.ContactParams(**wl.params)
    cpd = capi.ContactParams(**wl.params_driver)
    h.edbh_set_contact_params(g, b"config", C.byref(cp))
    h.edbh_set_contact_params(g, b"config_driver", C.byref(cpd))
    # update_particles_full (config_move_particles.msp:141-154), then numerical_scheme_verlet_macro, by YAML name
    for op in ("chunk_neighbors_contact", "nbh_sphere", "update_interaction_ghost", "classify_interactions"):
        run(op)
    direct.update_particle_neighbors(wl.rcut_inc, True)
    for it in range(6):
        run("combined_compute_prolog")
        run("particle_displ_over")
        over = C.c_int(0)
        assert h.edbh_get_bool(g, b"result", C.byref(over)) == 0
        if over.value:
            for op in ("unclassify_interactions", "compress_interaction", "chunk_neighbors_contact", "nbh_sphere", "classify_interactions"):
                run(op)
        else:
            run("ghost_update_rq")
        for op in ("reset_force_moment", "gravity_force", "contact_sphere", "force_to_accel", "combined_compute_epilog"):
            run(op)
        direct.combined_compute_prolog(wl.dt)
        if direct.displ_over(0.5 * wl.rcut_inc) > 0:
            direct.update_particle_neighbors(wl.rcut_inc, True)
        else:
            direct.ghost_update_rq()
        direct.reset_force_moment(); direct.gravity_force(wl.gravity); direct.contact_sphere(wl.dt); direct.force_to_accel()
        direct.combined_compute_epilog(wl.dt)
    pa, pb = named.download_particles(), direct.download_particles()
    for f in capi.F64_FIELDS:
        assert np.array_equal(getattr(pa, f), getattr(pb, f)), f
    # consumers of the contact state, by name
    run("stress_tensor")
    n = h.edbh_get_doubles(g, b"stress", None, C.c_longlong(0))
    assert n == 9 * pa.n
    out = np.zeros(n)
    h.edbh_get_doubles(g, b"stress", out.ctypes.data_as(C.POINTER(C.c_double)), C.c_longlong(n))
    assert np.array_equal(out.reshape(-1, 3, 3), direct.stress_tensor())
    h.edbh_graph_destroy(g)

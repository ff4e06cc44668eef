// operators.cpp -> libexadem_host_ops.so — C++ host operators with exaDEM's YAML operator NAMES, SLOTS and PARAMETERS
//  whose execute() calls the C ABI of libexadem_b200 instead of the reference's OpenMP / Onika loops.
// Each class cites the reference operator it stands in for (paths relative to /root/reference).
//
// Slot types: `grid` carries the device context (the particle grid lives in HBM between operators, replacing the
// reference's CUDA managed memory); `ges`/`ic` (GridCellParticleInteraction / Classifier) are held inside the same
// context, so classify/unclassify/compress become bookkeeping no-ops that are still registered under their names so
// that the reference's .msp decks run unchanged.
#include <cstring>
#include <sstream>

#include "../../include/exadem_b200.h"
#include "operator_node.hpp"

using namespace onika_lite;

namespace exaDEM {

struct Vec3d { double x, y, z; };
struct GridB200 { edb_context* ctx = nullptr; };           // `grid` slot (+ ges + ic + cvf, all device-resident)
struct ContactParams : edb_contact_params {};               // `config`, `config_driver`
struct MultiMatParams { int n_groups = 0, n_drivers = 0; std::vector<edb_contact_params> pp, pd; };  // `multimat_cp`

static void ck(GridB200& g, int rc, const char* op) {
  // color_log::error(op, msg) + std::exit in the reference (include/exaDEM/color_log.hpp:78-82); an exception here so that
  // the embedding framework decides
  if (rc != EDB_OK) throw std::runtime_error(std::string(op) + ": " + edb_last_error(g.ctx));
}

// -------------------------------------------------------------------------------------------------------------------
// neighbour / interaction rebuild  (data/config/config_spheres.msp:7-12, config_polyhedra.msp:11-18)
// -------------------------------------------------------------------------------------------------------------------
// src/grid/chunk_neighbors_contact.cpp:60-115 — the Verlet candidate build is fused into nbh_*: this operator only
// validates its slots (the candidate filter runs inside edb_update_particle_neighbors).
class ChunkNeighborsContact : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, rcut_inc, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "chunk_neighbors_contact: candidate filter 0<d2<(ra+rb+rcut_inc)^2, fused into nbh_sphere/nbh_polyhedron on the device."; }
  void execute() final { if (*rcut_inc < 0) throw std::runtime_error("chunk_neighbors_contact: rcut_inc < 0"); }
};

// src/sphere/nbh_sphere.cu:43-310
class UpdateContactInteractionSphere : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, rcut_inc, INPUT, REQUIRED, DocString{"value added to the search distance to update neighbor list less frequently. in physical space"});
  ADD_SLOT(bool, symetric, INPUT, REQUIRED, DocString{"Activate the use of symetric feature (contact law)"});
 public:
  std::string documentation() const final { return "This operator build the interactions lists.\n\nYAML example:\n\n  - nbh_sphere:\n     symetric: true"; }
  void execute() final { ck(*grid, edb_update_particle_neighbors(grid->ctx, *rcut_inc, *symetric ? 1 : 0), "nbh_sphere"); }
};
// src/polyhedron/nbh_polyhedron.cpp:40-385
class UpdateGridCellInteractionPolyhedron : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, rcut_inc, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "nbh_polyhedron: OBB cull + vertex-vertex / vertex-edge / vertex-face / edge-edge filters, both directions."; }
  void execute() final { ck(*grid, edb_update_particle_neighbors(grid->ctx, *rcut_inc, 1), "nbh_polyhedron"); }
};
// classify_interactions.cpp:29-137, unclassify_interactions.cpp:29-195, compress_interaction.cpp:30-57,
// update_interaction_ghost.cpp:38-285, traversals.cpp:29-96: bookkeeping that the device layout makes implicit.
class DeviceImplicit : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "no-op on the B200 path: interactions are produced classified (per-type, owner-sorted) and stay on the device."; }
  void execute() final {}
};

// -------------------------------------------------------------------------------------------------------------------
// contact operators  (src/sphere/contact_force_sphere.cu.in:47-333, src/polyhedron/contact_force_polyhedron.cu.in:46-319;
// registered names: src/sphere/CMakeLists.txt:62-107, src/polyhedron/CMakeLists.txt:46-105)
// -------------------------------------------------------------------------------------------------------------------
template <bool POLY, bool MULTIMAT, int LAW>
class ComputeContactClassifier : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT_OUTPUT, REQUIRED);
  ADD_SLOT(ContactParams, config, INPUT, OPTIONAL, DocString{"Contact parameters for sphere interactions"});
  ADD_SLOT(ContactParams, config_driver, INPUT, OPTIONAL, DocString{"Contact parameters for drivers, optional"});
  ADD_SLOT(MultiMatParams, multimat_cp, INPUT, OPTIONAL, DocString{"List of contact parameters for simulations with multiple materials"});
  ADD_SLOT(double, dt, INPUT, REQUIRED);
  ADD_SLOT(bool, symetric, INPUT_OUTPUT, REQUIRED, DocString{"Activate the use of symetric feature (contact law)"});
 public:
  std::string documentation() const final { return "This operator computes forces between particles and particles/drivers using the contact law."; }
  void execute() final {
    GridB200& g = *grid;
    if (!MULTIMAT) {
      if (!config.has_value()) throw std::runtime_error("contact: slot 'config' is required in single-material mode");  // check_slots, .cu.in:202-284
      ck(g, edb_set_contact_params(g.ctx, &*config, config_driver.has_value() ? &*config_driver : nullptr), "contact");
    } else {
      if (!multimat_cp.has_value()) throw std::runtime_error("contact: slot 'multimat_cp' is required in multi-material mode");
      ck(g, edb_set_contact_params_multimat(g.ctx, multimat_cp->n_groups, multimat_cp->pp.data(), multimat_cp->n_drivers, multimat_cp->pd.data()), "contact");
    }
    ck(g, edb_set_law(g.ctx, LAW), "contact");
    ck(g, POLY ? edb_contact_polyhedron(g.ctx, *dt) : edb_contact_sphere(g.ctx, *dt), "contact");
  }
};

// -------------------------------------------------------------------------------------------------------------------
// per-particle operators
// -------------------------------------------------------------------------------------------------------------------
// driver operators of the time loop (config_numerical_schemes.msp:13-24, config_move_particles.msp:17-24,61,92,152): the `drivers`
// slot is device-resident (edb_set_drivers / edb_set_driver_motions), the operators keep the reference's names and slots
// -------------------------------------------------------------------------------------------------------------------
// src/driver/push_f_v_r_driver.cpp:40-64
class PushAccVelocityToPositionDriver : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, physical_time, INPUT, REQUIRED);
  ADD_SLOT(double, dt, INPUT, REQUIRED, DocString{"dt is the time increment of the timeloop"});
  ADD_SLOT(double, dt_scale, INPUT, 1.0);
 public:
  std::string documentation() const final { return "This operator updates driver centers using their velocities. Not that accelerations are not used."; }
  void execute() final { ck(*grid, edb_push_f_v_r_driver(grid->ctx, *physical_time, (*dt) * (*dt_scale)), "push_f_v_r_driver"); }
};
// src/driver/push_f_v_driver.cpp:42-70
class PushAccelToVelocityDriver : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, dt, INPUT, REQUIRED, DocString{"dt is the time increment of the timeloop"});
  ADD_SLOT(double, dt_scale, INPUT, 1.0);
 public:
  std::string documentation() const final { return "This operator updates driver velocities.\n\nParameter:\n  - dt_scale [double]: apply a scale factor.\n\nYAML example:\n\n  - push_f_v_driver"; }
  void execute() final { ck(*grid, edb_push_f_v_driver(grid->ctx, (*dt) * (*dt_scale)), "push_f_v_driver"); }
};
// src/driver/push_av_to_quat_driver.cpp:43-70
class PushAngularVelocityToQuaternionDriver : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, physical_time, INPUT, REQUIRED);
  ADD_SLOT(double, dt, INPUT, REQUIRED, DocString{"dt is the time increment of the timeloop"});
  ADD_SLOT(double, dt_scale, INPUT, 1.0);
 public:
  std::string documentation() const final { return "This operator compute the new orientation using the angular velocity.\n\nYAML example [no option]:\n\n  - push_av_to_quat_driver"; }
  void execute() final { ck(*grid, edb_push_av_to_quat_driver(grid->ctx, *physical_time, (*dt) * (*dt_scale)), "push_av_to_quat_driver"); }
};
// src/driver/reset_force_driver.cpp:27-43
class ResetForceDriverFunctor : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "This operator reset the forces of drivers."; }
  void execute() final { ck(*grid, edb_reset_force_driver(grid->ctx), "reset_force_driver"); }
};
// src/driver/force_to_accel_driver.cpp:100-146
class ForceToAccelDriverFunctor : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, dt, INPUT, REQUIRED, DocString{"dt is the time increment of the timeloop"});
  ADD_SLOT(double, system_mass, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "force_to_accel_driver: sums the reaction of the contacts on the drivers that need it (all ranks), then converts forces to accelerations."; }
  void execute() final { ck(*grid, edb_force_to_accel_driver(grid->ctx, *dt, *system_mass), "force_to_accel_driver"); }
};
// src/driver/driver_vertices.cu:42-117 (force_host is meaningless here: there is no host path)
class DriverVertices : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(bool, force_host, INPUT, false, DocString{"Force computations on the host"});
 public:
  std::string documentation() const final { return "This operator calculates new vertex positions.\nIf rshape mesh velocity and angular velocity are equal to [0,0,0], vertices are not calculated."; }
  void execute() final { ck(*grid, edb_driver_vertices(grid->ctx), "driver_vertices"); }
};
// src/driver/grid_stl_mesh.cpp:35-269
class UpdateGridSTLMeshOperator : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, rcut_inc, INPUT, 0.0);
  ADD_SLOT(bool, force_reset, INPUT, false);
 public:
  std::string documentation() const final { return "grid_rshape_driver: update the list of mesh vertices / edges / faces of every cell (device build)."; }
  void execute() final { ck(*grid, edb_grid_rshape_driver(grid->ctx), "grid_rshape_driver"); }
};
// -------------------------------------------------------------------------------------------------------------------
// inner bonds (fragmentation decks, data/config/config_fragmentation.msp): per-step operators
// -------------------------------------------------------------------------------------------------------------------
// src/polyhedron/stick_polyhedra.cpp:45-365.  field::cluster is not part of the device-resident field set: the binding hands it over
// per particle id (two vectors copied from the grid's id / cluster fields).
class StickPolyhedraOperator : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, sticking_threshold, INPUT, REQUIRED);
  ADD_SLOT(std::vector<uint64_t>, cluster_particle_ids, INPUT, REQUIRED);
  ADD_SLOT(std::vector<double>, cluster, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "stick_polyhedra: glue the matching faces of neighbouring polyhedra of the same cluster with inner bonds."; }
  void execute() final {
    if (cluster_particle_ids->size() != cluster->size()) throw std::runtime_error("stick_polyhedra: ids / cluster size mismatch");
    ck(*grid, edb_stick_polyhedra(grid->ctx, *sticking_threshold, cluster->size(), cluster_particle_ids->data(), cluster->data()), "stick_polyhedra");
  }
};
// src/polyhedron/inner_bond_force.cu:44-139
class ComputeInnerBondForce : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, dt, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "This operator computes forces between inner_bonded particles using the contact law.\n\nYAML example:\n\n  - inner_bond_polyhedron"; }
  void execute() final { ck(*grid, edb_inner_bond_polyhedron(grid->ctx, *dt), "inner_bond_polyhedron"); }
};
// src/interaction/update_interfaces.cpp:28-95
class UpdateInterfaces : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "update_interfaces: rebuilds the interfaces (runs of inner bonds between the same two particles) and re-resolves the bonds by particle id."; }
  void execute() final { ck(*grid, edb_update_interfaces(grid->ctx, nullptr), "update_interfaces"); }
};
// src/interaction/apply_interface_fracture_criterion.cpp + include/exaDEM/interface/apply_interface_fracture_criterion.hpp:40-73
class ApplyInterfaceFractureCriterion : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(bool, display, INPUT, false);
  ADD_SLOT(bool, result, OUTPUT);
 public:
  std::string documentation() const final { return "apply_interface_fracture_criterion: breaks the interfaces whose energies exceed their criterion; result = true if at least one interface broke."; }
  void execute() final {
    uint64_t nb = 0;
    ck(*grid, edb_apply_interface_fracture_criterion(grid->ctx, &nb), "apply_interface_fracture_criterion");
    *result = nb > 0;
  }
};

// src/interaction/broken_interface_friction_updater.cpp:30-100
class BrokenInterfaceFrictionUpdater : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "This operator modify the cumulated friction [stored in interfactions] if an interface is broken: frition = max(friction, mu * fn)."; }
  void execute() final { ck(*grid, edb_broken_interface_friction_updater(grid->ctx), "broken_interface_friction_updater"); }
};

// src/driver/backup_drivers.cu
class BackupDrivers : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
 public:
  std::string documentation() const final { return "backup_drivers: copy the drivers into backup_drvs for driver_displ_over."; }
  void execute() final { ck(*grid, edb_backup_drivers(grid->ctx), "backup_drivers"); }
};
// src/driver/driver_displ_over.cu:196-243
class DriverDisplacementOver : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, threshold, INPUT, 0.0);
  ADD_SLOT(bool, result, OUTPUT);
 public:
  std::string documentation() const final { return "Compute the distance between each driver and its backup; sets result to true if a driver (a vertex of a mesh) has moved further than threshold."; }
  void execute() final {
    uint64_t over = 0;
    ck(*grid, edb_driver_displ_over(grid->ctx, *threshold, &over), "driver_displ_over");
    *result = over > 0;
  }
};

// -------------------------------------------------------------------------------------------------------------------
#define EDB_SIMPLE_OP(Class, call, doc)                                   \
  class Class : public OperatorNode {                                     \
    ADD_SLOT(GridB200, grid, INPUT_OUTPUT, REQUIRED);                     \
   public:                                                                \
    std::string documentation() const final { return doc; }               \
    void execute() final { ck(*grid, call(grid->ctx), #Class); }          \
  }
#define EDB_DT_OP(Class, call, doc)                                       \
  class Class : public OperatorNode {                                     \
    ADD_SLOT(GridB200, grid, INPUT_OUTPUT, REQUIRED);                     \
    ADD_SLOT(double, dt, INPUT, REQUIRED);                                \
   public:                                                                \
    std::string documentation() const final { return doc; }               \
    void execute() final { ck(*grid, call(grid->ctx, *dt), #Class); }     \
  }
// reset_force_moment / freeze_particles with the reference's optional `particle_regions` + `region` slots (reset_force_moment.cu:36-57,
// freeze_particles.cu:47-66).  ParticleRegions = the region table, ParticleRegionCSG = the expression over it, here already compiled to
// reverse Polish notation (what exaNBody's build_from_expression_string produces from `region: A and not B`).
using ParticleRegions = std::vector<edb_region>;
using ParticleRegionCSG = std::vector<int32_t>;
#define EDB_REGION_OP(Class, call, call_region, doc)                                                        \
  class Class : public OperatorNode {                                                                       \
    ADD_SLOT(GridB200, grid, INPUT_OUTPUT, REQUIRED);                                                       \
    ADD_SLOT(ParticleRegions, particle_regions, INPUT, OPTIONAL);                                           \
    ADD_SLOT(ParticleRegionCSG, region, INPUT, OPTIONAL);                                                   \
   public:                                                                                                  \
    std::string documentation() const final { return doc; }                                                 \
    void execute() final {                                                                                  \
      if (!region.has_value()) { ck(*grid, call(grid->ctx), #Class); return; }                              \
      if (!particle_regions.has_value())                                                                    \
        throw std::runtime_error(#Class ": Region is defined, but particle_regions has no value");          \
      ck(*grid, edb_set_particle_regions(grid->ctx, (int32_t)particle_regions->size(), particle_regions->data()), #Class); \
      ck(*grid, call_region(grid->ctx, (int32_t)region->size(), region->data()), #Class);                   \
    }                                                                                                       \
  }
EDB_REGION_OP(ResetForceMomentNode, edb_reset_force_moment, edb_reset_force_moment_region,
              "reset_force_moment: f = 0, mom = 0, optionally inside a region (src/numericalscheme/reset_force_moment.cu:33-73)");
EDB_SIMPLE_OP(ForceToAccel, edb_force_to_accel, "force_to_accel: f /= m (src/numericalscheme/force_to_accel.cu:32-53)");
EDB_SIMPLE_OP(PushToAngularAcceleration, edb_push_to_angular_acceleration, "push_to_angular_acceleration (src/numericalscheme/push_to_angular_acceleration.cu)");
EDB_REGION_OP(FreezeParticles, edb_freeze_particles, edb_freeze_particles_region,
              "freeze_particles / pas_libere_pas_delivre, optionally inside a region (src/numericalscheme/freeze_particles.cu:33-88)");
EDB_SIMPLE_OP(PolyhedraComputeVertices, edb_compute_vertices, "compute_vertices (src/polyhedron/compute_vertices.cu:34-109)");
EDB_SIMPLE_OP(UpdateGhostsRQ, edb_ghost_update_rq, "ghost_update_rq (src/grid/update_ghosts.cu:46-63)");
EDB_DT_OP(CombinedComputeProlog, edb_combined_compute_prolog, "combined_compute_prolog (src/numericalscheme/combined_compute_prolog.cu:81-117)");
EDB_DT_OP(CombinedComputeEpilog, edb_combined_compute_epilog, "combined_compute_epilog (src/numericalscheme/combined_compute_epilog.cu:81-114)");
EDB_DT_OP(PushToQuaternion, edb_push_to_quaternion, "push_to_quaternion (src/numericalscheme/push_to_quaternion.cu)");
EDB_DT_OP(PushToAngularVelocity, edb_push_to_angular_velocity, "push_to_angular_velocity (src/numericalscheme/push_to_angular_velocity.cu)");

// src/forcefield/gravity_force.cu:39-87
class GravityForce : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT_OUTPUT, REQUIRED);
  ADD_SLOT(Vec3d, gravity, INPUT, Vec3d{0.0, 0.0, -9.807}, DocString{"define the gravity constant in function of the gravity axis, default value are x axis = 0, y axis = 0 and z axis = -9.807"});
 public:
  std::string documentation() const final { return "This operator computes forces related to the gravity."; }
  void execute() final {
    const double g[3] = {gravity->x, gravity->y, gravity->z};
    ck(*grid, edb_gravity_force(grid->ctx, g), "gravity_force");
  }
};
// src/polyhedron/vertex_displ_over.cu:44-143 (and exaNBody particle_displ_over for spheres)
class DisplOver : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(double, threshold, INPUT, 0.0);
  ADD_SLOT(bool, result, OUTPUT);
 public:
  std::string documentation() const final { return "compute the distance between each particle (vertex) in grid input and its backup position; sets result to true if at least one moved further than threshold."; }
  void execute() final {
    uint64_t n = 0;
    ck(*grid, edb_displ_over(grid->ctx, *threshold, &n), "displ_over");
    result.set(n > 0);
  }
};

// src/interaction/stress_tensor.cu:142-178 — per-particle virial from the resident contact state; result in slot `stress`
class StressTensor : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT_OUTPUT, REQUIRED);
  ADD_SLOT(std::vector<double>, stress, OUTPUT, DocString{"9 doubles per particle, row-major, upload order"});
 public:
  std::string documentation() const final { return "This operator computes the total stress tensor and the stress tensor for each particles."; }
  void execute() final {
    uint64_t nr = 0;
    ck(*grid, edb_num_particles(grid->ctx, &nr, nullptr), "stress_tensor");
    std::vector<double> out(9 * nr);
    ck(*grid, edb_stress_tensor(grid->ctx, out.data()), "stress_tensor");
    stress.set(std::move(out));
  }
};
// src/interaction/dem_cost_model.cpp:37-92 — per-cell cost estimate for the load balancer; result in slot `cell_costs` (x fastest
// over the real cells) with the grid dimensions in `cell_costs_dims`
class DEMCostModel : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(std::vector<double>, cell_costs, OUTPUT);
  ADD_SLOT(std::vector<int>, cell_costs_dims, OUTPUT);
 public:
  std::string documentation() const final { return "Estimates the cost of every inner cell (3 N + 1 + weighted interaction counts) for the load balancing."; }
  void execute() final {
    int32_t dims[3] = {0, 0, 0};
    ck(*grid, edb_dem_cost_model(grid->ctx, dims, nullptr, 0), "dem_cost_model");
    std::vector<double> out((size_t)dims[0] * dims[1] * dims[2]);
    ck(*grid, edb_dem_cost_model(grid->ctx, dims, out.data(), out.size()), "dem_cost_model");
    cell_costs.set(std::move(out));
    cell_costs_dims.set(std::vector<int>{dims[0], dims[1], dims[2]});
  }
};
// src/io/simulation_state.cu:80-175 — the reductions of the log line (this rank's share); result in slot `simulation_state` as
// {kinetic xyz, rotation xyz, mass, min_dn, n_particles, n_active, n_total}
class SimulationStateNode : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT, REQUIRED);
  ADD_SLOT(bool, symetric, INPUT, REQUIRED, DocString{"Use of symetric feature (contact law)"});
  ADD_SLOT(std::vector<double>, simulation_state, OUTPUT);
  ADD_SLOT(double, system_mass, OUTPUT);
 public:
  std::string documentation() const final { return "Computes the kinetic / rotation energies, the mass, the particle and interaction counts and the largest overlap."; }
  void execute() final {
    edb_sim_state st;
    ck(*grid, edb_simulation_state(grid->ctx, *symetric ? 1 : 0, &st), "simulation_state");
    simulation_state.set(std::vector<double>{st.kinetic_energy[0], st.kinetic_energy[1], st.kinetic_energy[2], st.rotation_energy[0],
                                             st.rotation_energy[1], st.rotation_energy[2], st.mass, st.min_dn, (double)st.n_particles,
                                             (double)st.n_active_interactions, (double)st.n_total_interactions});
    system_mass.set(st.mass);
  }
};
// src/interaction/compute_driver_force_moment.cu:79-146 — results in slots `driver_forces` / `driver_moments` (3 doubles per driver)
class ComputeDriverForceMoment : public OperatorNode {
  ADD_SLOT(GridB200, grid, INPUT_OUTPUT, REQUIRED);
  ADD_SLOT(double, n_drivers, INPUT, REQUIRED);
  ADD_SLOT(std::vector<double>, driver_forces, OUTPUT);
  ADD_SLOT(std::vector<double>, driver_moments, OUTPUT);
 public:
  std::string documentation() const final { return "Reduces -(fn+ft) and the torque about the driver centre over the driver waves of the classifier."; }
  void execute() final {
    const size_t nd = (size_t)*n_drivers;
    std::vector<double> f(3 * nd), m(3 * nd);
    ck(*grid, edb_compute_driver_force_moment(grid->ctx, f.data(), m.data()), "compute_driver_force_moment");
    driver_forces.set(std::move(f));
    driver_moments.set(std::move(m));
  }
};

}  // namespace exaDEM

// === register factories (same names as the reference's ONIKA_AUTORUN_INIT blocks) ===
ONIKA_AUTORUN_INIT(exadem_b200_hot_path) {
  using namespace exaDEM;
  auto* f = OperatorNodeFactory::instance();
  f->register_factory("chunk_neighbors_contact", make_simple_operator<ChunkNeighborsContact>);
  f->register_factory("nbh_sphere", make_simple_operator<UpdateContactInteractionSphere>);
  f->register_factory("nbh_polyhedron", make_simple_operator<UpdateGridCellInteractionPolyhedron>);
  f->register_factory("stress_tensor", make_simple_operator<StressTensor>);
  f->register_factory("dem_cost_model", make_simple_operator<DEMCostModel>);
  f->register_factory("simulation_state", make_simple_operator<SimulationStateNode>);
  f->register_factory("compute_driver_force_moment", make_simple_operator<ComputeDriverForceMoment>);
  // bookkeeping operators of update_particles_full (config_move_particles.msp:141-154) that the device layout makes implicit
  for (const char* n : {"classify_interactions", "unclassify_interactions", "compress_interaction", "update_interaction_ghost", "update_traversals",
                        "backup_dem", "backup_r", "migrate_cell_particles_interaction", "move_particles_interaction"})
    f->register_factory(n, make_simple_operator<DeviceImplicit>);
  // src/sphere/CMakeLists.txt:62-107
  f->register_factory("contact_sphere", make_simple_operator<ComputeContactClassifier<false, false, 0>>);
  f->register_factory("contact_singlemat_sphere", make_simple_operator<ComputeContactClassifier<false, false, 0>>);
  f->register_factory("contact_singlemat_sphere_hooke", make_simple_operator<ComputeContactClassifier<false, false, 0>>);
  f->register_factory("contact_singlemat_sphere_hooke_none", make_simple_operator<ComputeContactClassifier<false, false, 0>>);
  f->register_factory("contact_singlemat_sphere_hooke_cohesive", make_simple_operator<ComputeContactClassifier<false, false, 1>>);
  f->register_factory("contact_singlemat_sphere_hooke_dmt", make_simple_operator<ComputeContactClassifier<false, false, 2>>);
  f->register_factory("contact_multimat_sphere_hooke_none", make_simple_operator<ComputeContactClassifier<false, true, 0>>);
  f->register_factory("contact_multimat_sphere_hooke_cohesive", make_simple_operator<ComputeContactClassifier<false, true, 1>>);
  f->register_factory("contact_multimat_sphere_hooke_dmt", make_simple_operator<ComputeContactClassifier<false, true, 2>>);
  // src/polyhedron/CMakeLists.txt:46-105
  f->register_factory("contact_polyhedron", make_simple_operator<ComputeContactClassifier<true, false, 0>>);
  f->register_factory("contact_singlemat_polyhedron", make_simple_operator<ComputeContactClassifier<true, false, 0>>);
  f->register_factory("contact_singlemat_polyhedron_hooke_none", make_simple_operator<ComputeContactClassifier<true, false, 0>>);
  f->register_factory("contact_singlemat_polyhedron_hooke_cohesive", make_simple_operator<ComputeContactClassifier<true, false, 1>>);
  f->register_factory("contact_singlemat_polyhedron_hooke_dmt", make_simple_operator<ComputeContactClassifier<true, false, 2>>);
  f->register_factory("contact_multimat_polyhedron_hooke_none", make_simple_operator<ComputeContactClassifier<true, true, 0>>);
  f->register_factory("contact_multimat_polyhedron_hooke_cohesive", make_simple_operator<ComputeContactClassifier<true, true, 1>>);
  f->register_factory("contact_multimat_polyhedron_hooke_dmt", make_simple_operator<ComputeContactClassifier<true, true, 2>>);
  f->register_factory("gravity_force", make_simple_operator<GravityForce>);
  f->register_factory("reset_force_moment", make_simple_operator<ResetForceMomentNode>);
  f->register_factory("force_to_accel", make_simple_operator<ForceToAccel>);
  f->register_factory("combined_compute_prolog", make_simple_operator<CombinedComputeProlog>);
  f->register_factory("combined_compute_epilog", make_simple_operator<CombinedComputeEpilog>);
  f->register_factory("push_to_quaternion", make_simple_operator<PushToQuaternion>);
  f->register_factory("push_to_angular_velocity", make_simple_operator<PushToAngularVelocity>);
  f->register_factory("push_to_angular_acceleration", make_simple_operator<PushToAngularAcceleration>);
  f->register_factory("freeze_particles", make_simple_operator<FreezeParticles>);
  f->register_factory("pas_libere_pas_delivre", make_simple_operator<FreezeParticles>);
  f->register_factory("compute_vertices", make_simple_operator<PolyhedraComputeVertices>);
  f->register_factory("ghost_update_rq", make_simple_operator<UpdateGhostsRQ>);
  f->register_factory("vertex_displ_over", make_simple_operator<DisplOver>);
  f->register_factory("particle_displ_over", make_simple_operator<DisplOver>);
  f->register_factory("push_f_v_r_driver", make_simple_operator<PushAccVelocityToPositionDriver>);
  f->register_factory("push_f_v_driver", make_simple_operator<PushAccelToVelocityDriver>);
  f->register_factory("push_av_to_quat_driver", make_simple_operator<PushAngularVelocityToQuaternionDriver>);
  f->register_factory("reset_force_driver", make_simple_operator<ResetForceDriverFunctor>);
  f->register_factory("force_to_accel_driver", make_simple_operator<ForceToAccelDriverFunctor>);
  f->register_factory("driver_vertices", make_simple_operator<DriverVertices>);
  f->register_factory("grid_rshape_driver", make_simple_operator<UpdateGridSTLMeshOperator>);
  f->register_factory("backup_drivers", make_simple_operator<BackupDrivers>);
  f->register_factory("driver_displ_over", make_simple_operator<DriverDisplacementOver>);
  f->register_factory("stick_polyhedra", make_simple_operator<StickPolyhedraOperator>);
  f->register_factory("inner_bond_polyhedron", make_simple_operator<ComputeInnerBondForce>);
  f->register_factory("update_interfaces", make_simple_operator<UpdateInterfaces>);
  f->register_factory("apply_interface_fracture_criterion", make_simple_operator<ApplyInterfaceFractureCriterion>);
  f->register_factory("broken_interface_friction_updater", make_simple_operator<BrokenInterfaceFrictionUpdater>);
}

// -------------------------------------------------------------------------------------------------------------------
// tiny C surface so that tests can drive the operator registry by YAML name (what the .msp interpreter does)
// -------------------------------------------------------------------------------------------------------------------
struct edbh_graph {
  std::map<std::string, std::shared_ptr<std::any>> pool;
  std::string err;
  template <class T>
  void put(const char* name, T v) {
    auto& p = pool[name];
    if (!p) p = std::make_shared<std::any>();
    *p = std::move(v);
  }
};

extern "C" {
edbh_graph* edbh_graph_create(edb_context* ctx) {
  auto* g = new edbh_graph();
  g->put("grid", exaDEM::GridB200{ctx});
  return g;
}
void edbh_graph_destroy(edbh_graph* g) { delete g; }
const char* edbh_last_error(edbh_graph* g) { return g->err.c_str(); }
void edbh_set_double(edbh_graph* g, const char* name, double v) { g->put(name, v); }
void edbh_set_bool(edbh_graph* g, const char* name, int v) { g->put(name, (bool)(v != 0)); }
void edbh_set_vec3(edbh_graph* g, const char* name, const double* v) { g->put(name, exaDEM::Vec3d{v[0], v[1], v[2]}); }
void edbh_set_contact_params(edbh_graph* g, const char* name, const edb_contact_params* p) {
  exaDEM::ContactParams c;
  static_cast<edb_contact_params&>(c) = *p;
  g->put(name, c);
}
void edbh_set_multimat(edbh_graph* g, int n_groups, const edb_contact_params* pp, int n_drivers, const edb_contact_params* pd) {
  exaDEM::MultiMatParams m;
  m.n_groups = n_groups; m.n_drivers = n_drivers;
  m.pp.assign(pp, pp + n_groups * n_groups);
  m.pd.assign(pd, pd + n_groups * n_drivers);
  g->put("multimat_cp", m);
}
// copies a std::vector<double> slot (e.g. `stress`, `driver_forces`) out; returns its length, or -1 when absent
long long edbh_get_doubles(edbh_graph* g, const char* name, double* out, long long cap) {
  auto it = g->pool.find(name);
  if (it == g->pool.end() || !it->second->has_value()) return -1;
  const auto* v = std::any_cast<std::vector<double>>(it->second.get());
  if (!v) return -1;
  if (out) memcpy(out, v->data(), sizeof(double) * (size_t)std::min<long long>(cap, (long long)v->size()));
  return (long long)v->size();
}
int edbh_get_bool(edbh_graph* g, const char* name, int* out) {
  auto it = g->pool.find(name);
  if (it == g->pool.end() || !it->second->has_value()) return -1;
  *out = std::any_cast<bool>(*it->second) ? 1 : 0;
  return 0;
}
int edbh_run(edbh_graph* g, const char* op_name) {
  try {
    auto op = OperatorNodeFactory::instance()->make_operator(op_name);
    op->bind(g->pool);
    op->check_required();
    op->execute();
    return 0;
  } catch (const std::exception& e) {
    g->err = e.what();
    return -1;
  }
}
int edbh_list_operators(char* buf, int cap) {
  std::ostringstream o;
  for (auto& n : OperatorNodeFactory::instance()->names()) o << n << "\n";
  const std::string s = o.str();
  if ((int)s.size() + 1 > cap) return -(int)s.size() - 1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}
int edbh_documentation(const char* op_name, char* buf, int cap) {
  try {
    const std::string s = OperatorNodeFactory::instance()->make_operator(op_name)->documentation();
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
  } catch (...) { return -1; }
}
}

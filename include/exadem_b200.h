/* exadem_b200.h — C ABI of libexadem_b200.so: the B200-native (sm_100a) DEM contact hot path.
 *
 * Drop-in boundary for the hot-path operators of exaDEM v1.2.2.  The reference has NO C ABI:
 * its operators are Onika `OperatorNode` plugins (C++), so each entry point below names the reference
 * operator (YAML name, file:line under /root/reference) whose `execute()` it replaces.  The C++ host operators
 * in exadem_b200/host/ keep the reference's operator names, slots and parameters and call these functions;
 * INTEGRATION.md shows the Onika-side binding.
 *
 * Conventions
 *  - plain pointers + sizes only; every function returns 0 on success, <0 on error (edb_last_error()).
 *  - all particle pointers are BORROWED host pointers (pageable or pinned) unless a function says "device".
 *  - one context per GPU / rank; a context is not re-entrant (reference: operators run sequentially from one
 *    host thread, SURVEY.md 8b "Threading").  Launches are asynchronous on the context stream; edb_sync() waits.
 *  - no CPU fallback: every compute entry point fails with EDB_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef EXADEM_B200_H
#define EXADEM_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define EDB_OK 0
#define EDB_ERR_INVALID (-1)
#define EDB_ERR_NO_DEVICE (-2)
#define EDB_ERR_CUDA (-3)
#define EDB_ERR_NCCL (-4)
#define EDB_ERR_STATE (-5)
#define EDB_ERR_CAPACITY (-6)

/* interaction type ids: src/interaction/include/exaDEM/interaction/interaction_enum.hpp:24-57 */
#define EDB_NTYPES 14
enum { EDB_MODE_SPHERE = 0, EDB_MODE_POLYHEDRON = 1 };
/* cohesive-law variants of the contact operators: src/forcefield/include/exaDEM/forcefield/contact_force.hpp:32-52 */
enum { EDB_LAW_HOOKE_NONE = 0, EDB_LAW_HOOKE_COHESIVE = 1, EDB_LAW_HOOKE_DMT = 2 };
/* DRIVER_TYPE: src/driver/include/exaDEM/driver_base.hpp */
enum { EDB_DRIVER_CYLINDER = 0, EDB_DRIVER_SURFACE = 1, EDB_DRIVER_BALL = 2, EDB_DRIVER_RSHAPE = 3, EDB_DRIVER_UNDEFINED = 4 };

typedef struct edb_context edb_context;

/* `domain` + `grid` slots (exaNBody Domain / Grid; contract in SURVEY.md 2f).  The brick [lo,hi) is THIS rank's
 * sub-domain; glo/ghi the global periodic box.  cell_size >= 2*r_max + rcut_inc (check_rcut.cpp:72,88-116). */
typedef struct edb_config {
  double lo[3], hi[3];   /* local brick bounds */
  double glo[3], ghi[3]; /* global domain bounds (== lo/hi on one GPU) */
  double cell_size;
  int32_t periodic[3];
  int32_t mode;          /* EDB_MODE_* */
  int32_t device;        /* CUDA ordinal */
  int32_t rank, nranks;  /* position in the process grid */
  int32_t pgrid[3];      /* process grid (bricks per axis), product == nranks */
  int32_t emit_contact_state; /* 1: keep dn/cp/fn/ft per interaction like classifier.hpp:238-249 (default 1) */
  int32_t reserved;
} edb_config;

/* DEMFieldSet (include/exaDEM/fields.h:34-54) as borrowed SoA pointers; NULL = field not transferred. */
typedef struct edb_particles {
  uint64_t n;
  double *rx, *ry, *rz, *vx, *vy, *vz, *fx, *fy, *fz;
  double *mass, *homothety, *radius;
  double *qw, *qx, *qy, *qz;          /* orient (Quaternion w,x,y,z) */
  double *momx, *momy, *momz;
  double *vrotx, *vroty, *vrotz;
  double *arotx, *aroty, *arotz;
  double *inertiax, *inertiay, *inertiaz;
  uint64_t* id;
  uint32_t* type;
  uint32_t* group;
} edb_particles;

/* ContactParams (src/forcefield/include/exaDEM/forcefield/contact_parameters.hpp:32-41) */
typedef struct edb_contact_params {
  double kn, kt, kr, mu, damp_rate, fc, dncut, gamma;
} edb_contact_params;

/* One driver (src/driver/include/exaDEM/{cylinder,surface,ball}.hpp fields).  Cylinder `center` is given as in
 * the YAML deck; the library applies the reference's component-wise `center*axis` pre-multiplication
 * (cylinder.hpp:104, quirk Q11). */
typedef struct edb_driver {
  int32_t kind;        /* EDB_DRIVER_* */
  int32_t need_forces; /* motion_type.hpp:133-141: accumulate -f on the driver */
  double radius;       /* cylinder, ball */
  double axis[3];      /* cylinder */
  double center[3];    /* cylinder, ball, surface reference centre */
  double offset;       /* surface */
  double normal[3];    /* surface */
  double vel[3], vrot[3];
  int32_t shape;       /* RSHAPE: mesh shape index (edb_set_driver_shapes table); center = position of the mesh */
  int32_t reserved;
  double quat[4];      /* RSHAPE: orientation (w,x,y,z), rshape.hpp:370-384 */
} edb_driver;

/* MotionType: src/driver/include/exaDEM/motion_type.hpp:23-35 (same numbering) */
enum { EDB_MOTION_STATIONARY = 0, EDB_MOTION_LINEAR_MOTION = 1, EDB_MOTION_COMPRESSIVE_FORCE = 2, EDB_MOTION_LINEAR_FORCE_MOTION = 3,
       EDB_MOTION_PARTICLE = 4, EDB_MOTION_LINEAR_COMPRESSIVE_MOTION = 5, EDB_MOTION_TABULATED = 6, EDB_MOTION_SHAKER = 7,
       EDB_MOTION_PENDULUM_MOTION = 8, EDB_MOTION_EXPRESSION = 9, EDB_MOTION_UNKNOWN = 10 };
#define EDB_MAX_TAB 64 /* entries of a TABULATED trajectory kept on the device per driver */

/* Motion of one driver = `Driver_params` (src/driver/include/exaDEM/driver_params.hpp:33-66) + the integrator state the
 * reference keeps in the driver's own fields (SurfaceFields surface.hpp:30-46, BallFields ball.hpp, RShapeDriverFields
 * rshape.hpp:37-52).  Entry i belongs to driver i of edb_set_drivers.  The values are taken AFTER the reference's
 * `initialize()` checks (normalised motion_vector / shaker_dir, surface motion_vector aligned with the normal). */
typedef struct edb_driver_motion {
  int32_t motion_type;             /* EDB_MOTION_* */
  int32_t n_tab;                   /* TABULATED: number of (time, position) entries, <= EDB_MAX_TAB */
  double motion_vector[3];
  double motion_start_threshold, motion_end_threshold;
  double const_vel, const_force;   /* LINEAR_MOTION / LINEAR_FORCE_MOTION */
  double sigma, damprate;          /* compressive motions */
  const double* tab_time;          /* n_tab, sorted, tab_time[0] == 0 (driver_params.hpp:141-166) */
  const double* tab_pos;           /* n_tab * 3 */
  double omega, amplitude;         /* SHAKER, PENDULUM_MOTION */
  double shaker_dir[3];
  double pendulum_anchor_point[3], pendulum_initial_position[3], pendulum_swing_dir[3];
  /* EXPRESSION (expr.hpp: tinyexpr strings): the strings stay on the host; the binding evaluates them at
   * `time - motion_start_threshold` and hands the values over with edb_set_driver_expr before push_f_v_r_driver. */
  int32_t expr_use_v, expr_use_vrot, expr_use_mom;
  int32_t drive_by_mom;            /* RShapeDriverFields::drive_by_mom_ */
  /* integrator state */
  double mass;                     /* fields_.mass_ (surface, ball, rshape) */
  double surface;                  /* fields_.surface_ (surface, rshape) */
  double acc[3];                   /* fields_.acc_ (surface: scalar in acc[0]) */
  double rv, ra;                   /* ball: radius velocity / acceleration */
  double inertia[3], applied_mom[3], mom_axis[3], mom[3]; /* rshape */
  double forces[3];                /* fields_.forces_: reaction summed by force_to_accel_driver (output of edb_get_drivers) */
} edb_driver_motion;

/* InnerBondParams (src/forcefield/include/exaDEM/forcefield/inner_bond_parameters.hpp:33-45), the `multimat_ibp` slot entries */
enum { EDB_RUPTURE_ENERGY_MIXED = 0, EDB_RUPTURE_ENERGY_SEPARATE = 1, EDB_RUPTURE_STRESS_ENERGY_SEPARATE = 2, EDB_RUPTURE_NONE = 3 };
typedef struct edb_inner_bond_params {
  double kn, kt, damp_rate;
  int32_t mode;     /* EDB_RUPTURE_* (rupture_criterion.hpp:26) */
  int32_t pad;
  double crit1, crit2;
} edb_inner_bond_params;
/* One inner bond = the fields of InnerBondInteraction (src/interaction/include/exaDEM/interaction/inner_bond_interaction.hpp:33-41)
 * with the particles named by id (the reference's (cell, p) addressing is rebuilt by update_persistent_interactions at every move) */
typedef struct edb_inner_bond {
  uint64_t id_i, id_j;
  uint32_t sub_i, sub_j;        /* glued vertices */
  double friction[3], en, tds[3], et, dn0, weight;
  double criteria_1, criteria_2; /* RuptureCriteria */
  int32_t mode;
  uint8_t unbroken, ghost;
  uint16_t pad;
} edb_inner_bond;

/* `shape` (src/shape/include/exaDEM/shape.hpp:50-69) flattened: faces as the [n0, v.., n1, v..] stream. */
typedef struct edb_shape {
  int32_t nv, ne, nf, faces_len;
  const double* vertices; /* nv*3 */
  const int32_t* edges;   /* ne*2 */
  const int32_t* faces;   /* faces_len ints: per face {n, v0..v(n-1)} */
  double obb[15];         /* center, e1, e2, e3, extent (extent includes the Minkowski radius, shape_prepro.hpp:95) */
  double radius;          /* Minkowski radius */
} edb_shape;

/* One exported interaction = `Interaction` (interaction.hpp:44-48 + interaction_pair.hpp:30-70), 104 B, same
 * member order/offsets as the reference record so dumps/restarts interoperate. */
typedef struct edb_interaction {
  uint64_t id_i; uint32_t cell_i; uint16_t p_i; uint16_t pad_i; uint32_t sub_i; uint32_t pad_i2;
  uint64_t id_j; uint32_t cell_j; uint16_t p_j; uint16_t pad_j; uint32_t sub_j; uint32_t pad_j2;
  uint16_t type; uint8_t swap; uint8_t ghost; uint32_t pad_t;
  double friction[3];
  double moment[3];
} edb_interaction;

/* ---- lifetime --------------------------------------------------------------------------------------- */
int edb_create(const edb_config* cfg, edb_context** out);
int edb_destroy(edb_context* ctx);
const char* edb_last_error(const edb_context* ctx); /* ctx may be NULL: last error of a failed edb_create */
int edb_sync(edb_context* ctx);                     /* ONIKA_CU_DEVICE_SYNCHRONIZE equivalent */
int edb_version(void);

/* ---- slots: parameters -------------------------------------------------------------------------------- */
/* `config` / `config_driver` slots of contact_sphere / contact_polyhedron (contact_force_sphere.cu.in:52-73) */
int edb_set_contact_params(edb_context*, const edb_contact_params* particle, const edb_contact_params* driver);
/* `multimat_cp` / drivers_contact_params tables (multimat_parameters.hpp:149-200): pp[gA*n_groups+gB], pd[g*n_drivers+d] */
int edb_set_contact_params_multimat(edb_context*, int32_t n_groups, const edb_contact_params* pp, int32_t n_drivers,
                                    const edb_contact_params* pd);
int edb_set_law(edb_context*, int32_t law); /* operator variant: ..._hooke_none | _cohesive | _dmt */
int edb_set_drivers(edb_context*, int32_t n, const edb_driver* drivers);      /* `drivers` slot */
int edb_set_shapes(edb_context*, int32_t n, const edb_shape* shapes);         /* `shapes_collection` slot */
/* shapes of the RShape (STL mesh) drivers: RShapeDriver::shp_, src/driver/include/exaDEM/rshape.hpp.  Call BEFORE edb_set_drivers.
 * The mesh shape arrays are an input (quirk Q12: the reference's STL->shape numbering is thread-order dependent). */
int edb_set_driver_shapes(edb_context*, int32_t n, const edb_shape* shapes);

/* motions of the drivers of the last edb_set_drivers call (n must match).  Without this call every driver is STATIONARY. */
int edb_set_driver_motions(edb_context*, int32_t n, const edb_driver_motion* motions);
/* EXPRESSION motion: values of the v / vrot / mom expressions for the coming push (Driver_params::driver_expr_*, driver_params.hpp:373-389) */
int edb_set_driver_expr(edb_context*, int32_t driver, const double v[3], const double vrot[3], const double mom[3]);
/* read the drivers back (state of the device-resident `drivers` slot): out / motions may be NULL; tab pointers are not filled */
int edb_get_drivers(edb_context*, int32_t n, edb_driver* out, edb_driver_motion* motions);

/* ---- slots: grid ------------------------------------------------------------------------------------- */
int edb_upload_particles(edb_context*, const edb_particles* host);   /* full DEMFieldSet, any order */
int edb_download_particles(edb_context*, edb_particles* host);       /* same order as uploaded/last rebuilt; see edb_get_order */
int edb_num_particles(edb_context*, uint64_t* n_real, uint64_t* n_ghost);
/* dynamic fields only (r, v, vrot, orient in; + f, mom, arot out) — the per-step host<->device traffic of a host-owned grid */
int edb_upload_dynamic(edb_context*, const edb_particles* host);
int edb_download_dynamic(edb_context*, edb_particles* host);
/* Asynchronous forms for host arrays in PINNED memory: the calls return once the copies are enqueued (own copy streams, ordered against
   the kernels by events).  An upload whose host arrays are the target of the previous asynchronous download waits for that download
   field by field, so on a full-duplex link the download of step n, the upload of step n+1 and the kernels of step n+1 overlap
   (the exaNBody grid a host operator hands over between two steps is exactly that: the arrays the previous step filled).
   field_mask of the download: bit 0 r, 1 v, 2 vrot, 3 orient, 4 f (acceleration), 5 mom, 6 arot; EDB_DYN_ALL = everything
   edb_download_dynamic returns.  edb_sync() (or any blocking entry point) before the host reads or writes the arrays. */
#define EDB_DYN_ALL 0x7fu
#define EDB_DYN_STATE 0x0fu /* r, v, vrot, orient: what the next step needs back */
int edb_upload_dynamic_async(edb_context*, const edb_particles* host_pinned);
int edb_download_dynamic_async(edb_context*, edb_particles* host_pinned, uint32_t field_mask);

/* ---- operators (reference YAML name in brackets) ----------------------------------------------------- */
/* [chunk_neighbors_contact + nbh_sphere | nbh_polyhedron + update_interaction_ghost + classify_interactions]
 * = chunk_neighbors_impl, data/config/config_spheres.msp:7-12, config_polyhedra.msp:11-18.  Re-bins particles,
 * rebuilds ghosts, enumerates candidates (src/sphere/nbh_sphere.cu:64-309, src/polyhedron/nbh_polyhedron.cpp:75-384),
 * carries friction/moment history over (interaction_manager.hpp:205-217) and classifies into waves. */
int edb_update_particle_neighbors(edb_context*, double rcut_inc, int32_t symetric);
int edb_ghost_update_rq(edb_context*);                       /* [ghost_update_rq] src/grid/update_ghosts.cu:46-63 */
int edb_compute_vertices(edb_context*);                      /* [compute_vertices] src/polyhedron/compute_vertices.cu:68-108 */
int edb_reset_force_moment(edb_context*);                    /* [reset_force_moment] src/numericalscheme/reset_force_moment.cu:55-72 */
int edb_gravity_force(edb_context*, const double g[3]);      /* [gravity_force] src/forcefield/gravity_force.cu:67-86 */
int edb_contact_sphere(edb_context*, double dt);             /* [contact_sphere*] src/sphere/contact_force_sphere.cu.in:132-188 */
int edb_contact_polyhedron(edb_context*, double dt);         /* [contact_polyhedron*] src/polyhedron/contact_force_polyhedron.cu.in:148-245 */
int edb_force_to_accel(edb_context*);                        /* [force_to_accel] src/numericalscheme/force_to_accel.cu:48-52 */
int edb_combined_compute_prolog(edb_context*, double dt);    /* [combined_compute_prolog] src/numericalscheme/combined_compute_prolog.cu:94-116 */
int edb_combined_compute_epilog(edb_context*, double dt);    /* [combined_compute_epilog] src/numericalscheme/combined_compute_epilog.cu:94-113 */
int edb_push_to_quaternion(edb_context*, double dt);         /* [push_to_quaternion] src/numericalscheme/push_to_quaternion.cu */
int edb_push_to_angular_velocity(edb_context*, double dt);   /* [push_to_angular_velocity] */
int edb_push_to_angular_acceleration(edb_context*);          /* [push_to_angular_acceleration] */
int edb_freeze_particles(edb_context*);                      /* [freeze_particles] src/numericalscheme/freeze_particles.cu:33-88 (no region) */
/* The `region` slot of reset_force_moment / freeze_particles (reset_force_moment.cu:36-57, freeze_particles.cu:47-66): exaNBody's
   ParticleRegions + ParticleRegionCSG.  A region is the conjunction of an axis-aligned box (lo <= r < hi), an id range
   [id_start, id_end) and a quadric  (x y z 1) Q (x y z 1)^T <= 0  (row-major 4x4), each optional; a CSG expression over the regions is
   given in reverse Polish notation: a token k >= 0 pushes "the particle is inside region k", EDB_CSG_AND / EDB_CSG_OR / EDB_CSG_NOT
   combine the top of the stack (the form exaNBody's build_from_expression_string compiles `region: A and not B` into).
   Parity unpinned: exaNBody's sources are not in the reference tree; restated from the operators' call sites and documentation. */
typedef struct edb_region {
  double lo[3], hi[3];
  double quadric[16];
  uint64_t id_start, id_end;
  int32_t use_bounds, use_ids, use_quadric, pad;
} edb_region;
#define EDB_CSG_AND (-1)
#define EDB_CSG_OR (-2)
#define EDB_CSG_NOT (-3)
#define EDB_CSG_MAX 32
int edb_set_particle_regions(edb_context*, int32_t n, const edb_region* regions);
int edb_reset_force_moment_region(edb_context*, int32_t n_expr, const int32_t* expr);   /* f = mom = 0 for the particles inside */
int edb_freeze_particles_region(edb_context*, int32_t n_expr, const int32_t* expr);     /* f = mom = v = vrot = 0 for the particles inside */
/* [particle_displ_over | vertex_displ_over] rebuild trigger: *over = #particles (vertices) moved >= threshold since
 * the last edb_update_particle_neighbors (which takes the backup, backup_dem.cpp:40-80); all-reduced over ranks. */
int edb_displ_over(edb_context*, double threshold, uint64_t* over);
/* ---- driver operators of the time loop (config_numerical_schemes.msp:13-24), SURVEY 8(f) rank 3.  The `drivers` slot lives on
 * the device: these run as one small kernel each on the context stream, no host round trip. ---------------------------- */
/* [push_f_v_r_driver] src/driver/push_f_v_r_driver.cpp:28-37 (only when is_motion_triggered(time + dt)); delta_t = dt * dt_scale */
int edb_push_f_v_r_driver(edb_context*, double physical_time, double delta_t);
int edb_push_f_v_driver(edb_context*, double delta_t);                          /* [push_f_v_driver] src/driver/push_f_v_driver.cpp:28-35 */
int edb_push_av_to_quat_driver(edb_context*, double physical_time, double delta_t); /* [push_av_to_quat_driver] src/driver/push_av_to_quat_driver.cpp:27-36 */
int edb_reset_force_driver(edb_context*);                                       /* [reset_force_driver] src/driver/reset_force_driver.cpp:38-42 */
/* [force_to_accel_driver] src/driver/force_to_accel_driver.cpp:30-146: sums the reaction -f (and moment) of the driver waves of the
 * last contact operator into the drivers that need them (deterministic block-ordered sum, all-reduced over the ranks), then
 * Ball::f_ra / force_to_accel of every driver.  system_mass = the `system_mass` slot. */
int edb_force_to_accel_driver(edb_context*, double dt, double system_mass);
/* [driver_vertices] src/driver/driver_vertices.cu:28-35: world vertices of the non-stationary RShape drivers */
int edb_driver_vertices(edb_context*);
/* [grid_rshape_driver] src/driver/grid_stl_mesh.cpp:89-258: per-cell vertex / edge / face lists of the RShape drivers, built on the
 * device (count -> scan -> fill -> per-list sort) from the current world vertices.  Also run by edb_update_particle_neighbors. */
int edb_grid_rshape_driver(edb_context*);
/* `physical_time` slot used by edb_run_verlet for the driver operators; advanced by dt per step (next_time_step) */
int edb_set_physical_time(edb_context*, double t);
int edb_get_physical_time(edb_context*, double* t);
/* `system_mass` slot (output of simulation_state, simulation_state.cu:199) used by edb_run_verlet for force_to_accel_driver;
 * when never set, the loop takes the mass sum of edb_simulation_state at its first step */
int edb_set_system_mass(edb_context*, double m);
int edb_backup_drivers(edb_context*);                                           /* [backup_drivers] src/driver/backup_drivers.cu */
/* [driver_displ_over] src/driver/driver_displ_over.cu:60-243: *result = number of drivers (mesh: vertices) that moved >= threshold
 * since the last backup (taken by edb_update_particle_neighbors / edb_backup_drivers) */
int edb_driver_displ_over(edb_context*, double threshold, uint64_t* result);

/* [compute_driver_force_moment] src/interaction/compute_driver_force_moment.cu:79-146: forces[n_drivers*3], moments[n_drivers*3] */
int edb_compute_driver_force_moment(edb_context*, double* forces, double* moments);
/* [stress_tensor] src/interaction/stress_tensor.cu:142-178: per-particle 3x3 (row-major, upload order), scaled by 1/volume outside */
int edb_stress_tensor(edb_context*, double* stress9);
/* [simulation_state] src/io/simulation_state.cu:80-175 (+ simulation_state.hpp:48-57, itools/counter.hpp:125-152): the reductions
 * behind the log line, on the device — per-axis kinetic 0.5 m v^2 and rotation 0.5 m vrot^2 energies, mass and particle count over
 * the real particles; interaction totals over the waves (PartnerGhost records skipped, particle-particle waves counted twice when
 * `symetric`), active = dn < 0 or non-zero friction/moment history, min_dn over the active ones.  Values of THIS rank (the
 * reference's operator then MPI_Allreduce's them). */
/* Cost-balanced decomposition (the reference's `load_balance_rcb` step of load_balancing_if_triggered, config_move_particles.msp:114-139,
 * fed by dem_cost_model): the bricks of the process grid are bounded by the per-axis cut positions c{x,y,z}[0..pgrid[d]] (first = glo,
 * last = ghi, NULL = uniform along that axis) instead of equal parts; this rank's lo/hi must be the cuts of its own brick.  Call on every
 * rank with the same cuts, before edb_comm_init and before the first upload. */
int edb_set_brick_cuts(edb_context*, const double* cx, const double* cy, const double* cz);

/* [dem_cost_model] src/interaction/dem_cost_model.cpp:63-92: per real cell of this rank's grid (x fastest, the ghost shell excluded)
 * cost = 3 N + 1 + sum over the interaction records owned by the cell's particles of the type weight {1,3,5,4, 1,1,1, 1,3,5,4,3,5, 1}
 * (types 0..13); with no interaction list built yet cost = 3 N + 1.  grid_dims[3] receives the real-cell grid dimensions; costs may
 * be NULL to query them.  The values feed the load balancer (exaNBody load_balance_rcb; here exadem_b200.parallel.balanced_cuts). */
int edb_dem_cost_model(edb_context*, int32_t grid_dims[3], double* costs, uint64_t cap);

typedef struct edb_sim_state {
  double kinetic_energy[3], rotation_energy[3], mass, min_dn;
  uint64_t n_particles, n_active_interactions, n_total_interactions;
} edb_sim_state;
int edb_simulation_state(edb_context*, int32_t symetric, edb_sim_state* out);

/* ---- inner bonds (fragmentation, wave 13), SURVEY 8(f) rank 4 ---------------------------------------------------------------
 * Bonds are created by edb_stick_polyhedra or enter through edb_import_inner_bonds (the fragmentation dump reader stays on the host
 * side of the binding).  While a bond holds, the polyhedron contact enumeration ignores the pair it glues
 * (InteractionManager::update_ignore_interaction).  With bonds present edb_run_verlet runs the fragmentation macro (fracture
 * criterion on the previous step's energies, friction cap + neighbour rebuild on a break, inner_bond_polyhedron inside
 * compute_force) operator by operator, together with the driver operators when drivers move.  NOT built yet: bonds across periodic /
 * brick boundaries (OwnerGhost bonds, several ranks). */
int edb_set_inner_bond_params(edb_context*, int32_t n_groups, const edb_inner_bond_params* ibp /* n_groups^2 */);
int edb_import_inner_bonds(edb_context*, const edb_inner_bond* in, uint64_t n);
/* [stick_polyhedra] src/polyhedron/stick_polyhedra.cpp:88-360 on the device: glues the matching faces of neighbouring polyhedra of the
 * same cluster (id_i < id_j) — one bond per vertex pair closer than the threshold, weight 1/#bonds, rupture criterion from the face
 * area and the bond parameters.  Call after edb_update_particle_neighbors (it walks the same neighbour cells); `cluster` is
 * field::cluster given per particle id (ids[n], cluster[n]).  Replaces the bond list.  Partners in ghost cells are skipped. */
int edb_stick_polyhedra(edb_context*, double sticking_threshold, uint64_t n, const uint64_t* ids, const double* cluster);
int edb_export_inner_bonds(edb_context*, edb_inner_bond* out, uint64_t cap, uint64_t* n); /* owner-sorted; out may be NULL to query n */
/* [update_interfaces] src/interaction/update_interfaces.cpp:44-84: interfaces = runs of bonds between the same two particles; also
 * re-resolves the particle slots of the bonds by id (to be called after every edb_update_particle_neighbors, like the deck does) */
int edb_update_interfaces(edb_context*, uint64_t* n_interfaces);
/* [inner_bond_polyhedron] src/polyhedron/inner_bond_force.cu + include/exaDEM/polyhedron/inner_bond.hpp:106-185: the bond law
 * (inner_bond_force.hpp:26-70) on every bond, forces and torques ADDED to the particles in a fixed order (no atomics) */
int edb_inner_bond_polyhedron(edb_context*, double dt);
/* [apply_interface_fracture_criterion] apply_interface_fracture_criterion.hpp:40-73: breaks the interfaces whose energies exceed
 * their criterion (all their bonds become unbroken = 0); *n_broken = interfaces broken by this call */
int edb_apply_interface_fracture_criterion(edb_context*, uint64_t* n_broken);
/* [broken_interface_friction_updater] src/interaction/broken_interface_friction_updater.cpp:58-98: caps the cumulated friction of the
 * bonds of the broken interfaces at mu * |fn| (CONTACT parameters) and zeroes it where the vertices are apart.  At the next
 * edb_update_particle_neighbors the broken bonds leave wave 13 and those with a non-zero friction become the history of the
 * vertex-vertex contact between the same two vertices (unclassify_interactions.cpp:135-145). */
int edb_broken_interface_friction_updater(edb_context*);
/* ContactState of wave 13 (dn, cp, fn, ft), one entry per bond in export order */
int edb_get_inner_bond_state(edb_context*, double* dn, double* cp, double* fn, double* ft);

/* Fused fast path = numerical_scheme_verlet (data/config/config_numerical_schemes.msp:6-30), `nsteps` iterations:
 * prolog -> (rebuild if displacement >= rcut_inc/2 | ghost_update_rq) -> [compute_vertices] -> reset+gravity+contact
 * +force_to_accel -> epilog.  rebuild_every > 0 forces a fixed rebuild schedule instead of the displacement trigger
 * (parity runs make the schedule an explicit input); 0 = trigger; <0 = never rebuild.
 * When a driver has a motion (edb_set_driver_motions) the loop also runs the driver operators of the macro in the reference's order —
 * push_f_v_r_driver, push_f_v_driver(dt/2), push_av_to_quat_driver after the prolog; driver_displ_over in the trigger; backup_drivers,
 * driver_vertices, grid_rshape_driver in the rebuild; reset_force_driver, gravity on PARTICLE meshes, force_to_accel_driver (system mass:
 * edb_set_system_mass, default = sum of the particle masses), push_f_v_driver(dt/2) before the epilog — fused into one small kernel per
 * step boundary, still without a host round trip per step.  `physical_time` (edb_set_physical_time) advances by dt per step. */
int edb_run_verlet(edb_context*, double dt, const double g[3], double rcut_inc, int32_t symetric, int32_t nsteps,
                   int32_t rebuild_every, uint64_t* n_rebuilds);

/* ---- classifier / ges views --------------------------------------------------------------------------- */
int edb_interaction_counts(edb_context*, uint64_t counts[EDB_NTYPES], uint64_t active[EDB_NTYPES]);
/* [unclassify_interactions + compress_interaction] view: all records (type-major, owner-sorted); cap = capacity of out */
int edb_export_interactions(edb_context*, edb_interaction* out, uint64_t cap, uint64_t* n);
int edb_import_interactions(edb_context*, const edb_interaction* in, uint64_t n); /* restart: history source for the next rebuild */
/* ContactState buffers (classifier/contact_state.hpp:22-51) of one wave; each array has counts[type] entries (x3 for vectors) */
int edb_get_contact_state(edb_context*, int32_t type, double* dn, double* cp, double* fn, double* ft);

/* ---- multi-GPU (replaces MPI in ghost_update_rq / migrate_cell_particles_interaction) ------------------- */
int edb_comm_unique_id(uint8_t id_out[128]);                       /* rank 0: ncclGetUniqueId */
int edb_comm_init(edb_context*, const uint8_t id[128]);            /* all ranks: ncclCommInitRank(nranks, rank) */

/* ---- measurement hooks -------------------------------------------------------------------------------- */
typedef struct edb_counters {
  uint64_t kernel_launches;     /* kernels of this library launched since creation */
  uint64_t interaction_updates; /* sum over waves of records processed by the contact kernels */
  uint64_t particle_steps;      /* real particles x steps */
  uint64_t rebuilds;
  double   contact_ms;          /* CUDA-event time of all contact wave kernels (accumulated, when timing enabled) */
  double   gather_ms, prolog_ms, epilog_ms, rebuild_ms, ghost_ms;
  uint64_t contact_launches;
  uint64_t bytes_contact_algorithmic; /* accumulated algorithmic bytes of the contact kernels */
  double   contact_pp_ms;       /* wave 0 (particle-particle) kernel alone: the dominant kernel of the step */
  uint64_t contact_pp_launches;
  uint64_t contact_pp_records;  /* records processed by those launches */
  uint64_t gather_launches;
} edb_counters;
int edb_set_timing(edb_context*, int32_t enabled);
int edb_get_counters(edb_context*, edb_counters* out);
int edb_reset_counters(edb_context*);
int edb_flush_l2(edb_context*); /* writes a >L2-sized scratch buffer (bench hygiene) */
/* CUDA events on the context stream (the stream every kernel of this library is launched on): slot in [0,8) */
int edb_event_record(edb_context*, int32_t slot);
int edb_event_elapsed_ms(edb_context*, int32_t slot_begin, int32_t slot_end, double* ms); /* synchronises slot_end */

#ifdef __cplusplus
}
#endif
#endif /* EXADEM_B200_H */

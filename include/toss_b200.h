/* include/toss_b200.h -- C ABI of libtoss_b200.so: the B200 (sm_100a) population-fitness path of TOSS.
 *
 * This is the drop-in boundary.  The reference (rasmusmarak/TOSS) is pure
 * Python; what its hot path binds today are three third-party packages, one call per RK stage:
 *     polyhedral_gravity.GravityEvaluable.__call__      toss/trajectory/equations_of_motion.py:58
 *     desolver.OdeSystem(...).integrate(events=...)     toss/trajectory/compute_trajectory.py:243-259
 *     pyquaternion.Quaternion(axis, angle).rotate       toss/trajectory/equations_of_motion.py:112-115
 * and numpy for resampling / fitness / ray casting.  Each entry point below names the reference
 * interface it replaces.  A maintainer binds this library with ctypes (INTEGRATION.md shows the stub);
 * toss_b200/_native.py is that binding.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  Pointers named d_* are DEVICE pointers owned by
 *     the caller (torch CUDA tensors in the Python host); h_* are HOST pointers.
 *   - every function returns 0 on success, non-zero on failure; toss_last_error() gives the message
 *     (thread-local).  No exceptions cross the ABI.
 *   - `stream` is a cudaStream_t passed as uintptr_t (0 = legacy default stream).  Calls are
 *     asynchronous w.r.t. the host unless stated; a context is used from one host thread at a time.
 *   - FP64 throughout; integer outputs (voxel counts, collision flags) are exact.
 */
#ifndef TOSS_B200_H
#define TOSS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct toss_ctx toss_ctx;

/* The scalar fields of the reference's `args` DotMap that the hot path reads
 * (toss/resources/default_cfg.toml + derived fields of toss/optimization/setup_parameters.py:49-56). */
typedef struct {
    double density;                 /* args.body.density */
    double spin_axis[3];            /* args.body.spin_axis (normalised on use, as pyquaternion does) */
    double spin_velocity;           /* args.body.spin_velocity */
    int32_t activate_rotation;      /* args.problem.activate_rotation */
    double rtol, atol;              /* args.integrator.rtol / atol */
    double start_time, final_time;  /* args.problem.start_time / final_time (int32-truncated on use) */
    double initial_time_step;       /* args.problem.initial_time_step */
    int32_t activate_event;         /* args.problem.activate_event */
    int32_t number_of_spacecrafts;  /* args.problem.number_of_spacecrafts (coverage chunking quirk) */
    double radius_inner;            /* args.problem.radius_inner_bounding_sphere */
    double radius_outer;            /* args.problem.radius_outer_bounding_sphere */
    double measurement_period;      /* args.problem.measurement_period */
    double penalty_scaling_factor;  /* args.problem.penalty_scaling_factor */
    double target_squared_altitude; /* args.problem.target_squared_altitude */
    double maximal_measurement_sphere_volume;
    double total_measurable_volume;
    int32_t max_steps;              /* per-segment cap on attempted steps (safety; status 1 when hit) */
    int32_t stop_on_collision;      /* 0 = reference behaviour: integrate to final_time, judge afterwards */
} toss_params;

/* per-trajectory accounting (throughput metrics: gravity evaluations = n_rhs) */
typedef struct {
    int64_t n_rhs;
    int32_t n_accepted;
    int32_t n_rejected;
    int32_t n_event_points;   /* accepted states with |r| <= radius_inner (tested by the ray caster) */
    int32_t status;           /* 0 ok, 1 step cap, 2 non-finite state, 3 knot capacity exhausted */
} toss_counters;

/* byte offsets of the arrays inside a propagation workspace (see toss_workspace_layout) */
typedef struct {
    size_t knots_t;     /* double [pop][knot_cap]                                              */
    size_t knots_y;     /* double [pop][knot_cap][6]   accepted states (desolver OdeSystem.y)   */
    size_t knots_f;     /* double [pop][knot_cap][6]   RHS at the knots (Hermite dense output)  */
    size_t seg_start;   /* int32  [pop][n_man+2]       first knot of each segment; [n_seg] = n_knots */
    size_t intervals;   /* double [pop][n_man+2]       integration_intervals                    */
    size_t n_seg;       /* int32  [pop]                                                        */
    size_t counters;    /* toss_counters [pop]                                                 */
    size_t collision;   /* int32  [pop]                collision_detected                      */
    size_t queue;       /* int32  [4 + 2 pop]          work-queue cursor | order | predicted cost (internal) */
    size_t total_bytes;
} toss_ws_layout;

const char *toss_version(void);
const char *toss_last_error(void);

/* One context per process/GPU.  `device` is the CUDA ordinal. */
int toss_ctx_create(int device, toss_ctx **out);
int toss_ctx_destroy(toss_ctx *ctx);

/* Replaces polyhedral_gravity.GravityEvaluable((vertices, faces), density)
 * (toss/optimization/setup_parameters.py:63) and the vertices/faces the ray caster reads
 * (toss/mesh/mesh_utility.py:70-89).  HOST arrays: verts V x 3 (metres), faces F x 3 (CCW outward).
 * Matches the faces into pairs that share an edge (the unit of the gravity sum: 4 vertex distances and 5 edge
 * terms per pair instead of 6 and 6; host, integer work) and precomputes the per-pair constants (normals, edge
 * normals, edge lengths, 2*area) in a kernel -- what GravityEvaluable precomputes at construction
 * (setup_parameters.py:63).  Fails (error 2, toss_last_error says why) for a surface that is not closed and
 * consistently oriented (an edge not shared by exactly two faces in opposite directions), for inward-pointing
 * winding (negative volume) and for zero-area faces. */
int toss_mesh_upload(toss_ctx *ctx, const double *h_verts, int V, const int32_t *h_faces, int F, double density);
int toss_mesh_num_faces(const toss_ctx *ctx);
/* The pairing: n = toss_mesh_num_pairs records (fA, qA, fB) ascending in fA; A's edge slot qA (v_qA -> v_qA+1)
 * is the shared edge; fB = -1 for a face left without a partner.  out = int32[3 * n].
 * toss_pair_faces is the matching alone (host integer work, needs no device or context): out = int32[3 * F]
 * capacity, returns the number of records or -1. */
int toss_mesh_num_pairs(const toss_ctx *ctx);
int toss_mesh_pairs(const toss_ctx *ctx, int32_t *out);
int toss_pair_faces(const int32_t *h_faces, int F, int V, int32_t *out);
/* The per-pair constants as the device built them (pairs.cu: pair_tables_kernel): out = double[n_pairs][44], the
 * 43 fields of a pair record (v0 v1 v2 v3 | N_A N_B | 3 + 3 in-plane edge normals | 5 edge lengths | 2A_A 2A_B) + one
 * pad word.  For inspection and tests: the constants are plain IEEE + - * / sqrt in a fixed order, so a host
 * re-computation reproduces them bit for bit. */
int toss_mesh_pair_records(toss_ctx *ctx, double *h_out);

/* Replaces args.problem.tensor_grid_r/theta/phi + bool_tensor
 * (toss/fitness/fitness_function_utils.py:231-279).  HOST arrays; bool_tensor is nr*nth*nph bytes. */
int toss_grid_upload(toss_ctx *ctx, const double *h_r, int nr, const double *h_theta, int nth,
                     const double *h_phi, int nph, const uint8_t *h_bool_tensor);

/* Replaces args.mesh.evaluable(x, parallel) (equations_of_motion.py:58) for n points at once.
 * d_pts n x 3 -> d_acc n x 3 in the PACKAGE sign convention (TOSS negates it, :59);
 * d_pot (n) may be NULL. */
int toss_gravity_eval(toss_ctx *ctx, const double *d_pts, int64_t n, double *d_acc, double *d_pot, uintptr_t stream);

/* Replaces compute_motion(t, y, args) (equations_of_motion.py:62-95) for n states: d_t (n), d_y n x 6
 * -> d_out n x 6 = [v, -gravity(rotate(-w t) r)]. */
int toss_compute_motion(toss_ctx *ctx, const toss_params *p, const double *d_t, const double *d_y, int64_t n,
                        double *d_out, uintptr_t stream);

/* Replaces rotate_point(t, x, spin_axis, spin_velocity) (equations_of_motion.py:98-117, pyquaternion
 * Quaternion(axis, angle = -w t).rotate) for n points: d_t (n), d_x n x 3 -> d_out n x 3. */
int toss_rotate_points(toss_ctx *ctx, const double *h_axis3, double spin_velocity, const double *d_t,
                       const double *d_x, int64_t n, double *d_out, uintptr_t stream);

/* workspace for toss_propagate: the caller allocates total_bytes of device memory */
int toss_workspace_layout(int pop, int n_man, int knot_cap, toss_ws_layout *out);

/* Replaces compute_trajectory(x, args, compute_motion) (compute_trajectory.py:13-162) for a population.
 * d_x: pop x dim decision vectors (row-major); h_fixed: the n_fixed leading values the UDP prepends
 * (udp_initial_condition.py:117-120; 0, 3 or 7).  n_fixed + dim must equal 7 + 5 n_man.
 * Fills the workspace (knots, segments, counters, collision flags: every accepted state inside the risk sphere is
 * ray-cast against the mesh by the stepper itself, compute_trajectory.py:141-159). */
int toss_propagate(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim,
                   const double *h_fixed, int n_fixed, int n_man, void *d_workspace, int knot_cap,
                   uintptr_t stream);

/* The two halves of toss_propagate, separately launchable (bench.py times the first alone):
 * toss_integrate = the stepper (desolver OdeSystem.integrate for every segment, compute_trajectory.py:130-148);
 * toss_collision_verdict = is_outside over the recorded event points (compute_trajectory.py:154-159). */
int toss_integrate(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                   int n_fixed, int n_man, void *d_workspace, int knot_cap, uintptr_t stream);
int toss_collision_verdict(toss_ctx *ctx, const toss_params *p, void *d_workspace, int pop, int n_man, int knot_cap,
                           uintptr_t stream);

/* Replaces get_trajectory_fixed_step(args, list_of_ode_objects) (trajectory_tools.py:28-76) for the
 * trajectories in the workspace: d_pos / d_vel are pop x 3 x N (N = toss_num_samples), d_ts is N. */
int64_t toss_num_samples(const toss_params *p);
int toss_resample(toss_ctx *ctx, const toss_params *p, const void *d_workspace, int pop, int n_man, int knot_cap,
                  double *d_pos, double *d_vel, double *d_ts, uintptr_t stream);

/* Replaces ode_object._OdeSystem__sol(t) (trajectory_tools.py:67): dense output of ONE segment given
 * its knots (n_knots of t, y[6], f[6]) at nt times -> d_out nt x 6. */
int toss_dense_eval(toss_ctx *ctx, const double *d_kt, const double *d_ky, const double *d_kf, int n_knots,
                    const double *d_t, int64_t nt, double *d_out, uintptr_t stream);

/* Replaces get_fitness(enum, args, positions, velocities, timesteps) (fitness_functions.py:7-45) for a
 * batch: d_pos is batch x 3 x N, d_ts is N; fn = FitnessFunctions value 1..9.
 * d_fitness (batch); d_n_visited (batch, may be NULL) = number of True voxels after the OR (exact). */
int toss_fitness_eval(toss_ctx *ctx, int fn, const toss_params *p, const double *d_pos, int batch, int64_t N,
                      const double *d_ts, double *d_fitness, int64_t *d_n_visited, uintptr_t stream);

/* Replaces update_spherical_tensor_grid(...) (fitness_function_utils.py:165-228): ORs the voxels visited
 * by ONE trajectory (d_pos 3 x N) into the context's bool_tensor; h_bool_tensor_out (may be NULL)
 * receives the updated tensor (synchronises the stream). */
int toss_update_tensor(toss_ctx *ctx, const toss_params *p, const double *d_pos, int64_t N, const double *d_ts,
                       uint8_t *h_bool_tensor_out, uintptr_t stream);

/* Replaces is_outside(points, vertices, faces) (mesh_utility.py:70-166): d_pts n x 3 -> d_out n bytes,
 * 1 = outside (even number of +z ray hits). */
int toss_is_outside(toss_ctx *ctx, const double *d_pts, int64_t n, uint8_t *d_out, uintptr_t stream);

/* Replaces udp_initial_condition.fitness (udp_initial_condition.py:106-137) for a whole population
 * (the new batch_fitness): propagate -> collision verdict -> resample -> fitness enum 9, with the
 * resample+fitness fused (positions never materialised).  Device buffers:
 * d_fitness (pop), d_collision (pop int32, may be NULL), d_counters (pop, may be NULL),
 * d_n_visited (pop int64, may be NULL; -1 for collided trajectories). */
int toss_population_fitness(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim,
                            const double *h_fixed, int n_fixed, int n_man, void *d_workspace, int knot_cap,
                            double *d_fitness, int32_t *d_collision, toss_counters *d_counters,
                            int64_t *d_n_visited, uintptr_t stream);

/* Same with HOST buffers (what a pygmo batch_fitness call holds): copies h_x in, runs, copies
 * h_fitness (and the optional arrays) out and synchronises.  Uses an internal workspace that grows
 * on demand. */
int toss_population_fitness_host(toss_ctx *ctx, const toss_params *p, const double *h_x, int pop, int dim,
                                 const double *h_fixed, int n_fixed, int n_man, int knot_cap,
                                 double *h_fitness, int32_t *h_collision, toss_counters *h_counters,
                                 int64_t *h_n_visited);

/* ---- one population over several GPUs --------------------------------------------------------
 * Replaces the process pool of pg.mp_bfe (toss.py:110; every worker evaluates a share of the generation,
 * scripts/scaling_performance_benchmark.py:126-137).  One process per GPU; every rank holds the whole pop x dim
 * chromosome matrix (a pygmo batch_fitness call hands it to all of them) and evaluates the rows DEALT to it:
 *   toss_predict_cost   cost key of each row of d_x (a short proxy integration through a mascon model of the body;
 *                       all zero when the mesh is so small that the proxy would not pay) -> d_cost int32[pop].
 *                       Each rank predicts a contiguous slice; the host all-gathers the slices (NCCL).
 *   toss_cost_order     d_order int32[n] = the rows in longest-first order, ties by index: a pure function of d_cost,
 *                       identical on every rank.
 *   toss_deal_rows      rank g's shard: d_x_local[j] = d_x[d_order[g + j*world]], j < ceil((n - g) / world) -- equal
 *                       predicted work for every rank, each shard already longest-first.
 *   toss_set_queue_order(ctx, 1) tells toss_population_fitness / toss_propagate that the rows come pre-ordered (no
 *                       second proxy run); 0 restores the default (the launcher orders its queue itself).
 *   toss_pack_result    d_send double[cap + 2] = the shard's fitness padded with NaN to cap | #trajectories with
 *                       status 3 | #with status 1 or 2 -- the all-gather payload.
 *   toss_unpack_result  d_recv double[world][cap + 2] (the gathered payloads) -> d_fitness double[n] in population
 *                       order, d_flags double[2] = the two counts summed over the ranks. */
int toss_set_queue_order(toss_ctx *ctx, int mode);
int toss_predict_cost(toss_ctx *ctx, const toss_params *p, const double *d_x, int pop, int dim, const double *h_fixed,
                      int n_fixed, int n_man, int32_t *d_cost, uintptr_t stream);
int toss_cost_order(toss_ctx *ctx, const int32_t *d_cost, int n, int32_t *d_order, uintptr_t stream);
int toss_deal_rows(toss_ctx *ctx, const double *d_x, int n, int dim, const int32_t *d_order, int rank, int world,
                   double *d_x_local, uintptr_t stream);
int toss_pack_result(toss_ctx *ctx, const double *d_fitness_local, const toss_counters *d_counters_local, int n_local,
                     int cap, double *d_send, uintptr_t stream);
int toss_unpack_result(toss_ctx *ctx, const double *d_recv, const int32_t *d_order, int n, int world, int cap,
                       double *d_fitness, double *d_flags, uintptr_t stream);

/* ---- the optimiser's candidates, generated where they are consumed ----------------------------
 * Replaces the sampling step of pygmo's gaco (pg.gaco(...).evolve, toss.py:51-66 with default_cfg.toml:72-82): ants
 * first_ant .. first_ant + n - 1 of generation `generation` draw a kernel k from the cumulative rank weights
 * h_cum_weights[ker] and every gene h from N(h_archive_x[k][h], h_sigma[h]), re-drawn while outside
 * [h_lb[h], h_ub[h]] (clamped after 64 draws), into d_x_out[n][dim] -- the chromosome matrix toss_population_fitness
 * reads, so a generation's decision vectors never cross PCIe.  Counter-based random numbers: ant j's genes depend on
 * (seed, generation, j) only (toss_b200/optimization/gaco.py has the numpy mirror). */
int toss_gaco_sample(toss_ctx *ctx, const double *h_archive_x, int ker, int dim, const double *h_cum_weights,
                     const double *h_sigma, const double *h_lb, const double *h_ub, uint64_t seed, uint64_t generation,
                     int first_ant, int n, double *d_x_out, uintptr_t stream);

/* Measures the FP64 FMA peak of the device (register-resident DFMA chains on every SM) in TFLOP/s;
 * the roofline denominator of bench.py (MEASURED_PEAKS.json has no FP64 entry). */
int toss_measure_fp64_peak(toss_ctx *ctx, double *tflops_out);

/* Test hooks of the arithmetic specification (toss_b200/csrc/tmath.cuh): evaluates elementary function `fn`
 * (0 two_atanh, 1 log, 2 atan2(a,b), 3 sincos -> (out,out2), 4 x^(1/8), 5 x^(1/4), 6 8-term atan series,
 * 7 a/b, 8 sqrt, 9 2 atanh minimax polynomial, 10 2 atan minimax polynomial, 11 / 12 far-tier 2 atanh / 2 atan,
 * 13 middle-tier 2 atanh, 14 wide 2 atanh) over n device values; and counts, over `samples` pseudo-random operands, how often the
 * guard-free division / square root differ from the IEEE-rounded intrinsics (must be 0). */
int toss_math_probe(toss_ctx *ctx, int fn, const double *d_a, const double *d_b, int64_t n, double *d_out,
                    double *d_out2, uintptr_t stream);
int toss_selftest_div_sqrt(toss_ctx *ctx, int64_t samples, int64_t *bad_div, int64_t *bad_sqrt);

/* number of kernel launches issued through this context so far (bench.py's gpu_launches) */
int64_t toss_launch_count(const toss_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif

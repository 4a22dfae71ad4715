/* oracle/toss_oracle.h -- CPU restatement of the TOSS population-fitness path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under toss_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker / CPU baseline.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * rasmusmarak/TOSS).  Third-party arithmetic that is NOT in the reference tree
 * (polyhedral-gravity-model, desolver, pyquaternion; all un-pinned in
 * environment.yml:5-19) is restated from the published algorithms:
 *   - Tsoulis (2012) line-integral polyhedral gravity, with the LN / AN
 *     arguments in their cancellation-free closed forms (2 atanh(|G|/(l1+l2)); signed solid
 *     angle by one atan2) -- same values, ~100x less rounding noise than the literal form,
 *     which is kept as orc_gravity_direct and compared in tests/test_oracle_gravity.py,
 *   - Prince & Dormand (1981) RK8(7)-13M with desolver's step controller
 *     (safety 0.8, exponent 1/8, max-norm) and desolver's
 *     cubic-Hermite dense output,
 *   - axis-angle rotation (Euler-Rodrigues == quaternion sandwich).
 * Parity pin: tests/test_oracle_golden.py checks this file against every golden
 * vector in the reference's tests/ (integration, fixed-step, manoeuvres, fitness,
 * covered-space, rotation).  Collision flags / event points and gravity-alone
 * values have no golden in the reference: "parity unpinned" for those two
 * (gravity is instead checked against a 50-digit mpmath evaluation).
 */
#ifndef TOSS_ORACLE_H
#define TOSS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_mesh orc_mesh;

typedef struct {
    /* body */
    double density;
    double spin_axis[3];
    double spin_velocity;
    int32_t activate_rotation;
    /* integrator */
    double rtol, atol;
    /* problem */
    double start_time, final_time, initial_time_step;
    int32_t activate_event;
    int32_t number_of_spacecrafts;      /* chunking quirk, fitness_function_utils.py:96 */
    double radius_inner, radius_outer;
    double measurement_period;
    double penalty_scaling_factor;
    double target_squared_altitude;
    double maximal_measurement_sphere_volume;
    double total_measurable_volume;
    int32_t max_steps;                  /* safety cap per segment */
    int32_t stop_on_collision;          /* 0 = reference behaviour (integrate to the end) */
} orc_params;

typedef struct {
    int64_t n_rhs;        /* gravity evaluations */
    int32_t n_accepted;
    int32_t n_rejected;
    int32_t n_event_points;
    int32_t status;       /* 0 ok, 1 step cap hit, 2 non-finite state */
} orc_counters;

orc_mesh *orc_mesh_create(const double *verts, int V, const int32_t *faces, int F, double density);
void orc_mesh_free(orc_mesh *m);
int orc_mesh_faces(const orc_mesh *m);
/* the face pairing the gravity sum runs over: out = int32[n][3] (fA, qA, fB; fB = -1: no partner), or NULL;
 * returns n.  At most orc_mesh_faces() records. */
int orc_mesh_pairs(const orc_mesh *m, int32_t *out);

/* polyhedral_gravity.GravityEvaluable.__call__ (call site equations_of_motion.py:58):
 * acc is the PACKAGE sign convention (TOSS negates it, equations_of_motion.py:59). */
void orc_gravity(const orc_mesh *m, const double *pts, int64_t n, double *acc, double *pot, int threads);
/* the same sum with LN/AN transcribed literally from Tsoulis (2012); ~1e-12 relative rounding
 * noise on the fine meshes (tests/test_oracle_gravity.py) -- cross-check only. */
void orc_gravity_direct(const orc_mesh *m, const double *pts, int64_t n, double *acc, double *pot);
/* census of the tiers / fall-backs orc_gravity takes for these points: far rows, middle rows, per-term rows,
 * ln-based edge terms, full-range atan2 terms */
void orc_gravity_tiers(const orc_mesh *m, const double *pts, int64_t n, int64_t out[5]);

/* equations_of_motion.py:98-117 */
void orc_rotate_point(double t, const double x[3], const double axis[3], double w, double out[3]);
/* equations_of_motion.py:62-95 */
void orc_compute_motion(const orc_mesh *m, const orc_params *p, double t, const double y[6], double out[6]);

/* compute_trajectory.py:13-162.  x = [r(3), v_mag, v_dir(3), (t_m, dv_mag, dv_dir(3)) x M].
 * Knots of all segments are concatenated; seg_start[k] is the first knot of segment k,
 * seg_start[n_seg] = n_knots.  intervals has n_seg+1 entries.
 * Returns 0 on success, <0 if capacity is too small. */
int orc_compute_trajectory(const orc_mesh *m, const orc_params *p, const double *x, int n_man,
                           double *knots_t, double *knots_y, double *knots_f, int cap,
                           int *seg_start, double *intervals, int *n_seg, int *n_knots,
                           int *collision_detected, double *event_pts, int event_cap,
                           orc_counters *cnt);

/* trajectory_tools.py:28-76 (index logic) + desolver cubic-Hermite dense output. */
int64_t orc_num_samples(const orc_params *p);
void orc_fixed_step(const orc_params *p, const double *knots_t, const double *knots_y,
                    const double *knots_f, const int *seg_start, int n_seg,
                    double *pos /*3xN*/, double *vel /*3xN*/, double *ts /*N*/);
/* evaluates the dense output of ONE segment at arbitrary times (the _OdeSystem__sol callable) */
void orc_dense_eval(const double *kt, const double *ky, const double *kf, int n,
                    const double *t, int64_t nt, double *out /*nt x 6*/);

/* mesh_utility.py:70-89,132-166 */
void orc_is_outside(const orc_mesh *m, const double *pts, int64_t n, uint8_t *out);

/* fitness_functions.py:7-45 (enum ids as fitness_function_enums.py:4-48) */
double orc_get_fitness(int fn, const orc_params *p, const double *pos /*3xN row-major*/, int64_t N,
                       const double *ts /*N*/, const double *gr, int nr, const double *gth, int nth,
                       const double *gph, int nph, const uint8_t *bool_tensor,
                       int64_t *n_visited /*optional: total true voxels after OR*/);
/* fitness_function_utils.py:54-162; single = 1 selects the positions.ndim==1 branch (:84-93). */
double orc_space_coverage(int n_spacecrafts, const double axis[3], double w, const double *pos,
                          int64_t N, const double *ts, double rmin, double rmax,
                          const double *gr, int nr, const double *gth, int nth, const double *gph, int nph,
                          const uint8_t *bool_tensor, int single, int64_t *n_visited,
                          uint8_t *new_tensor /*optional out*/);
/* fitness_function_utils.py:165-228 (in-place OR into bool_tensor) */
void orc_update_tensor(int n_spacecrafts, const double axis[3], double w, const double *pos,
                       int64_t N, const double *ts, double rmin, double rmax,
                       const double *gr, int nr, const double *gth, int nth, const double *gph, int nph,
                       uint8_t *bool_tensor);

/* udp_initial_condition.py:106-137 for a whole population (OpenMP over chromosomes).
 * xs: pop x dim decision vectors, fixed: n_fixed leading values (0, 3 or 7). */
void orc_population_fitness(const orc_mesh *m, const orc_params *p,
                            const double *gr, int nr, const double *gth, int nth, const double *gph, int nph,
                            const uint8_t *bool_tensor,
                            const double *xs, int pop, int dim, const double *fixed, int n_fixed, int n_man,
                            double *fitness, int32_t *collision, orc_counters *cnt, int64_t *n_visited,
                            int threads);

/* orc_math.h functions by id (tests only): 0 two_atanh 1 log 2 atan2(a,b) 3 sincos 4 root8 5 root4
 * 6 atan series (8 terms) 7 a/b 8 sqrt */
void orc_math_probe(int fn, const double *a, const double *b, int64_t n, double *out, double *out2);

#ifdef __cplusplus
}
#endif
#endif

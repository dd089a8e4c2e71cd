/*
 * flock_oracle.h -- CPU oracle for the feriphys flocking step.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * algorithm (jalberse/feriphys, src/simulation/flocking/ *.rs, state.rs, ...),
 * one separately rounded binary32 operation per source operation, evaluated in
 * the source's order.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (feriphys_b200/, include/feriphys_cuda.h) never links, imports or calls it.
 *
 * PARITY STATUS: the flocking arithmetic is "parity unpinned" -- the reference
 * has no tests, fixtures or golden vectors for flocking/obstacle/attractor/
 * bounding-box code, and no Rust toolchain exists here to run it.  The oracle
 * is pinned only (a) against the two State golden tests the reference does
 * hold (state.rs:166-185 Euler, state.rs:218-280 RK4) and (b) against
 * hand-derived known answers (tests/test_oracle_kat.py).
 *
 * Semantics of un-vendored dependencies (cgmath 0.18.0, approx 0.4.0,
 * num-traits 0.2.15, Rust std) are declared in flock_oracle.c next to the
 * helper that restates each one.
 */
#ifndef FLOCK_ORACLE_H
#define FLOCK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flocking::Config, flocking.rs:15-34 (defaults :36-51).  Duration is
 * carried as whole seconds + nanoseconds, like std::time::Duration. */
typedef struct {
    float dt;
    float avoidance_factor;
    float centering_factor;
    float velocity_matching_factor;
    float distance_weight_threshold;
    float distance_weight_threshold_falloff;
    float max_sight_angle;
    float max_sight_angle_to_lead_boid;
    uint64_t time_to_start_steering_secs;
    uint32_t time_to_start_steering_nanos;
    int32_t steering_overrides;
} orc_config;

/* Everything Simulation owns besides the boids (flocking.rs:53-60).
 * NULL / 0 == None.  Layouts: lead = pos3 vel3 weight (7 floats);
 * attractor = pos3 mass (4); obstacle = pos3 radius (4);
 * bbox = x.start x.end y.start y.end z.start z.end (6). */
typedef struct {
    const float *leads;
    uint32_t n_leads;
    const float *attractors;
    uint32_t n_attractors;
    const float *obstacles;
    uint32_t n_obstacles;
    const float *bbox;
} orc_scene;

/* Per-boid status bits raised where the reference would panic. */
#define ORC_FLAG_STEER_NEGATIVE 1u /* Duration::from_secs_f32(x<0)          */
#define ORC_FLAG_STEER_NAN_OVF 2u  /* Duration::from_secs_f32(NaN|overflow) */

void orc_config_default(orc_config *cfg);

/* boid.rs:139-166.  self6 = pos3 vel3.  out3 = acceleration on self. */
void orc_pair_accel(const float *self6, const float *other_pos3, const float *other_vel3,
                    float other_weight, float f_a, float f_c, float f_v, float thr, float fall,
                    float max_sight_angle, float *out3);

/* boid.rs:101-107 and :94-96 */
float orc_sight_angle(const float *self6, const float *other_pos3);
float orc_distance(const float *self6, const float *other_pos3);

/* point_attractor.rs:16-19, bounding_box.rs:13-26, flocking.rs:182-209 */
void orc_attractor_accel(const float *attr4, const float *pos3, float mass, float *out3);
void orc_bbox_accel(const float *bbox6, const float *pos3, float *out3);
void orc_steering_accel(const orc_config *cfg, const orc_scene *scene, const float *boid6,
                        float *out3, uint32_t *flags);

/* Accelerations of rows [i0,i1) of an N-boid flock (flocking.rs:101-114,
 * :133-209).  total3 is required; comp15 (boids, lead, attractors, bbox,
 * steering -- 3 floats each) and flags may be NULL.  threads<=1 is the
 * reference's single-threaded loop; threads>1 uses OpenMP over i (each row's
 * j-sum stays sequential, so results are identical). */
void orc_accel_rows(const orc_config *cfg, const orc_scene *scene, uint64_t n,
                    const float *state6, uint64_t i0, uint64_t i1, float *total3,
                    float *comp15, uint32_t *flags, int threads);

/* One Simulation::step (flocking.rs:97-122; leads are stepped by the caller,
 * see orc_lead_step).  state_out6 must not alias state_in6. */
void orc_step(const orc_config *cfg, const orc_scene *scene, uint64_t n, const float *state_in6,
              float *state_out6, uint32_t *flags, int threads);

/* Neighbour sets N(i) = { j : !(boid_j == boid_i) && !culled(i,j) &&
 * (dist <= thr || dist < thr+fall) } for rows [i0,i1): count and an
 * order-independent 64-bit hash (sum of mix64(j)).  list (if not NULL)
 * receives up to list_cap ascending indices per row at list + r*list_cap. */
void orc_neighbors_rows(const orc_config *cfg, uint64_t n, const float *state6, uint64_t i0,
                        uint64_t i1, uint32_t *count, uint64_t *hash, uint32_t *list,
                        uint32_t list_cap, int threads);
uint64_t orc_mix64(uint64_t j);

/* Pair-outcome census for the flop model: over rows [i0,i1) x all j,
 * out3 = { rejected by distance, rejected by FOV (in range), contributing }. */
void orc_pair_census(const orc_config *cfg, uint64_t n, const float *state6, uint64_t i0,
                     uint64_t i1, uint64_t *out3, int threads);

/* Grid-accelerated oracle: identical results to the literal loops (candidates
 * from the 27 surrounding cells, evaluated with the same literal pair function
 * in ascending j).  Proven identical in tests/test_oracle_grid.py. */
typedef struct orc_grid orc_grid;
orc_grid *orc_grid_build(const orc_config *cfg, uint64_t n, const float *state6);
void orc_grid_free(orc_grid *g);
void orc_grid_accel_rows(const orc_grid *g, const orc_config *cfg, const orc_scene *scene,
                         uint64_t n, const float *state6, uint64_t i0, uint64_t i1,
                         float *total3, float *comp15, uint32_t *flags, int threads);
void orc_grid_neighbors_rows(const orc_grid *g, const orc_config *cfg, uint64_t n,
                             const float *state6, uint64_t i0, uint64_t i1, uint32_t *count,
                             uint64_t *hash, int threads);
void orc_grid_step(const orc_config *cfg, const orc_scene *scene, uint64_t n,
                   const float *state_in6, float *state_out6, uint32_t *flags, int threads);

/* std::time::Duration::from_secs_f32 / as_secs_f32 [ext: Rust std].
 * Returns 0, or ORC_FLAG_* where Rust panics. */
uint32_t orc_duration_from_secs_f32(float secs, uint64_t *out_secs, uint32_t *out_nanos);
float orc_duration_as_secs_f32(uint64_t secs, uint32_t nanos);

/* LeadBoid::step + Parametric::step (boid.rs:46-53, parametric.rs:17-21).
 * lead7 = pos3 vel3 weight, updated in place; *curr_time updated in place.
 * path is the reference's fn(t)->Vector3.  Returns 0 or ORC_FLAG_*. */
typedef void (*orc_path_fn)(float t, float *out3, void *ctx);
uint32_t orc_lead_step(float *lead7, float *curr_time, float dt, orc_path_fn path, void *ctx);
/* The three closures of demos/flocking.rs:105-107,139-148 (kind 0,1,2). */
void orc_demo_path(int kind, float t, float *out3);

/* state.rs:75-106.  deriv maps a flat state vector to its derivative
 * (State::derivative over T::num_state_elements chunks is the caller's job). */
typedef void (*orc_deriv_fn)(const float *s, float *ds, size_t n, void *ctx);
void orc_state_euler(const float *s, size_t n, float h, orc_deriv_fn deriv, void *ctx, float *out);
void orc_state_rk4(const float *s, size_t n, float h, orc_deriv_fn deriv, void *ctx, float *out);
/* the two Stateful impls of the reference's own tests, state.rs:125-164, :193-216 */
void orc_deriv_test_point(const float *s, float *ds, size_t n, void *ctx);
void orc_deriv_test_examplefn(const float *s, float *ds, size_t n, void *ctx);

/* next_oracle.c -- SURVEY 8(f) rows.  Stateful::derivative of the reference's element types:
 * springy Point (10 floats, springy_mesh.rs:199-257), rigid-body State (29 floats,
 * rigidbody.rs:53-190), and a boid with frozen acceleration (9 floats). */
void orc_deriv_springy_point(const float *s, float *ds, size_t n, void *ctx);
void orc_deriv_rigidbody(const float *s, float *ds, size_t n, void *ctx);
void orc_deriv_boid(const float *s, float *ds, size_t n, void *ctx);
/* sph/mod.rs:89-121: k nearest within kernal_max_distance (ascending distance, ties by id) and the
 * Monaghan density; out_index n x k (0xffffffff beyond out_count). */
void orc_sph_neighbors(uint64_t n, const float *pos3, uint32_t k, float s, float mass, uint32_t *out_index,
                       uint32_t *out_count, float *out_density);

/* Largest c in [-1,1] with acosf(c) > theta (this libm); -2.0f if none.
 * Used by tests to check the GPU library's threshold form of the FOV test. */
float orc_acos_threshold(float theta);
/* Counts monotonicity violations of acosf over floats [lo_bits, hi_bits]
 * walking upward in value (both of the same sign). */
uint64_t orc_acos_monotone_violations(uint32_t lo_bits, uint32_t hi_bits);

/* Simulation::get_boid_instances (flocking.rs:230-245) with cgmath 0.18
 * Quaternion::from_arc: out8 = pos3, quat (s, x, y, z), scale. */
void orc_instances(uint64_t n, const float *state6, float *out8);

#ifdef __cplusplus
}
#endif
#endif

/*
 * flock_oracle.c -- CPU oracle for the feriphys flocking step (see header).
 *
 * TEST INFRASTRUCTURE ONLY; never on the product path.  PARITY UNPINNED for
 * the flocking arithmetic (the reference holds no golden vectors for it);
 * pinned for State::euler_step / rk4_step (state.rs:166-280).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -fno-fast-math -fopenmp (Makefile).
 * Every float operation below is one IEEE binary32 operation, written in the
 * order the Rust source evaluates it (Rust/LLVM neither reassociates nor
 * contracts float arithmetic).  x86-64 SSE has no excess precision.
 *
 * All file:line citations are relative to the reference repository root.
 */
#include "flock_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* cgmath 0.18.0 primitives [ext: Cargo.lock:209-210]                         */
/* ------------------------------------------------------------------------- */
typedef struct {
    float x, y, z;
} v3;

static inline v3 v3_new(float x, float y, float z) {
    v3 r = {x, y, z};
    return r;
}
static inline v3 v3_load(const float *p) { return v3_new(p[0], p[1], p[2]); }
static inline void v3_store(float *p, v3 a) {
    p[0] = a.x;
    p[1] = a.y;
    p[2] = a.z;
}
static inline v3 v3_zero(void) { return v3_new(0.0f, 0.0f, 0.0f); }
static inline v3 v3_add(v3 a, v3 b) { return v3_new(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_new(a.x - b.x, a.y - b.y, a.z - b.z); }
/* Vector3 * f32 and f32 * Vector3 are both component-wise products */
static inline v3 v3_scale(v3 a, float s) { return v3_new(a.x * s, a.y * s, a.z * s); }
static inline v3 s_times_v3(float s, v3 a) { return v3_new(s * a.x, s * a.y, s * a.z); }
/* Vector3 / f32 is three true divisions (used by boid.rs:51) */
static inline v3 v3_div(v3 a, float s) { return v3_new(a.x / s, a.y / s, a.z / s); }
/* InnerSpace::dot = mul_element_wise().sum() = (x + y) + z */
static inline float v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
/* InnerSpace::magnitude = sqrt(magnitude2) */
static inline float v3_magnitude(v3 a) { return sqrtf(v3_dot(a, a)); }
/* InnerSpace::normalize = normalize_to(1) = self * (1 / magnitude): ONE
 * division then three products, not three divisions */
static inline v3 v3_normalize(v3 a) { return v3_scale(a, 1.0f / v3_magnitude(a)); }
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_new(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* approx 0.4.0 abs_diff_eq! on Vector3<f32> [ext: Cargo.lock:63-64]:
 * every component |a-b| <= f32::EPSILON; NaN anywhere => false */
#define F32_EPSILON 1.1920929e-07f
static inline int f32_abs_diff_eq(float a, float b) {
    /* approx: (if a > b {a - b} else {b - a}) <= epsilon */
    float d = (a > b) ? (a - b) : (b - a);
    return d <= F32_EPSILON;
}
static inline int v3_abs_diff_eq(v3 a, v3 b) {
    return f32_abs_diff_eq(a.x, b.x) && f32_abs_diff_eq(a.y, b.y) && f32_abs_diff_eq(a.z, b.z);
}

/* Rust std [ext]: f32::powi(2) lowers to x*x; f32::powf(2.0) is folded by
 * LLVM (pow(x, 2.0) -> x*x) in optimised builds.  DECLARED CHOICE: x*x.
 * Isolated here so it can be flipped to powf(x, 2.0f). */
static inline float f32_square(float x) { return x * x; }

/* ------------------------------------------------------------------------- */
/* std::time::Duration [ext: Rust std, toolchain unpinned]                    */
/* ------------------------------------------------------------------------- */
/* Duration::from_secs_f32: panics on negative, NaN and >= 2^64 s; otherwise
 * the exact value of the float times 1e9, rounded to the nearest nanosecond,
 * ties to even (Rust >= 1.63; older toolchains truncated -- DECLARED CHOICE:
 * round-to-nearest-even).  x * 1e9 is exact in binary64 (24-bit significand
 * times 2^9 * 5^9 < 2^30 needs <= 54 bits only above 2^23 s, where the value
 * is an integer and the product has <= 45 significant bits), so rint() of the
 * double product is the exact answer. */
uint32_t orc_duration_from_secs_f32(float secs, uint64_t *out_secs, uint32_t *out_nanos) {
    *out_secs = 0;
    *out_nanos = 0;
    if (secs < 0.0f) return ORC_FLAG_STEER_NEGATIVE; /* -0.0 passes, as in Rust */
    if (!(secs < 18446744073709551616.0f)) return ORC_FLAG_STEER_NAN_OVF; /* NaN or >= 2^64 */
    if (secs >= 8388608.0f) { /* integer-valued: no fractional part */
        *out_secs = (uint64_t)secs;
        return 0;
    }
    double total_ns = rint((double)secs * 1e9); /* < 8.4e15 < 2^53: exact integer */
    uint64_t ns = (uint64_t)total_ns;
    *out_secs = ns / 1000000000ull;
    *out_nanos = (uint32_t)(ns % 1000000000ull);
    return 0;
}

/* Duration::as_secs_f32 = (secs as f32) + (nanos as f32) / 1e9f32 */
float orc_duration_as_secs_f32(uint64_t secs, uint32_t nanos) {
    return (float)secs + (float)nanos / 1000000000.0f;
}

typedef struct {
    uint64_t secs;
    uint32_t nanos;
} dur;
static const dur DUR_MAX = {UINT64_MAX, 999999999u};
static inline int dur_cmp(dur a, dur b) {
    if (a.secs != b.secs) return a.secs < b.secs ? -1 : 1;
    if (a.nanos != b.nanos) return a.nanos < b.nanos ? -1 : 1;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* flocking::Config::default, flocking.rs:36-51                               */
/* ------------------------------------------------------------------------- */
void orc_config_default(orc_config *cfg) {
    /* Duration::from_millis(1).as_secs_f32() = 0f32 + (1_000_000 as f32)/1e9 */
    cfg->dt = orc_duration_as_secs_f32(0, 1000000u);
    cfg->avoidance_factor = 1.0f;
    cfg->centering_factor = 0.1f;
    cfg->velocity_matching_factor = 0.5f;
    cfg->distance_weight_threshold = 15.0f;
    cfg->distance_weight_threshold_falloff = 1.0f;
    cfg->max_sight_angle = 3.14159274101257324f / 2.0f;  /* std::f32::consts::PI / 2.0 */
    cfg->max_sight_angle_to_lead_boid = 3.14159274101257324f;
    cfg->time_to_start_steering_secs = 4;
    cfg->time_to_start_steering_nanos = 0;
    cfg->steering_overrides = 0;
}

/* ------------------------------------------------------------------------- */
/* boid.rs: the pair function                                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    v3 pos, vel;
} boid;

static inline boid boid_load(const float *s6) {
    boid b = {v3_load(s6), v3_load(s6 + 3)};
    return b;
}

/* boid.rs:94-96 */
static inline float boid_distance(const boid *self, v3 other_pos) {
    return v3_magnitude(v3_sub(other_pos, self->pos));
}

/* boid.rs:101-107 */
static inline float boid_sight_angle(const boid *self, v3 other_pos) {
    return acosf(v3_dot(v3_normalize(self->vel), v3_normalize(v3_sub(other_pos, self->pos))));
}

/* boid.rs:110-117: -1.0 * factor / dist.powf(2.0) * normalize(d) * weight */
static inline v3 boid_avoidance(const boid *self, v3 other_pos, float other_weight, float factor) {
    if (v3_abs_diff_eq(other_pos, self->pos)) return v3_zero();
    float s = (-1.0f * factor) / f32_square(boid_distance(self, other_pos));
    return v3_scale(s_times_v3(s, v3_normalize(v3_sub(other_pos, self->pos))), other_weight);
}

/* boid.rs:120-128: factor * dist * normalize(d) * weight */
static inline v3 boid_centering(const boid *self, v3 other_pos, float other_weight, float factor) {
    if (v3_abs_diff_eq(other_pos, self->pos)) return v3_zero();
    float s = factor * boid_distance(self, other_pos);
    return v3_scale(s_times_v3(s, v3_normalize(v3_sub(other_pos, self->pos))), other_weight);
}

/* boid.rs:131-136: factor * (v_o - v_s) * weight */
static inline v3 boid_velocity_matching(const boid *self, v3 other_vel, float other_weight,
                                        float factor) {
    if (v3_abs_diff_eq(other_vel, self->vel)) return v3_zero();
    return v3_scale(s_times_v3(factor, v3_sub(other_vel, self->vel)), other_weight);
}

/* boid.rs:139-166 */
static inline v3 boid_get_acceleration(const boid *self, v3 other_pos, v3 other_vel,
                                       float other_weight, float f_a, float f_c, float f_v,
                                       float thr, float fall, float max_sight_angle) {
    if (boid_sight_angle(self, other_pos) > max_sight_angle) return v3_zero();
    float distance_weight;
    if (boid_distance(self, other_pos) <= thr) {
        distance_weight = 1.0f;
    } else if (boid_distance(self, other_pos) >= thr + fall) {
        distance_weight = 0.0f;
    } else {
        /* rises 0 -> 1 across the falloff band (doc comment says 1 -> 0) */
        distance_weight = (boid_distance(self, other_pos) - thr) / fall;
    }
    v3 sum = v3_add(v3_add(boid_avoidance(self, other_pos, other_weight, f_a),
                           boid_centering(self, other_pos, other_weight, f_c)),
                    boid_velocity_matching(self, other_vel, other_weight, f_v));
    return s_times_v3(distance_weight, sum);
}

/* #[derive(PartialEq)] on FlockingBoid (boid.rs:56-64): position, velocity,
 * weight, mass all ==; weight and mass are always 1.0 (boid.rs:81-88), so
 * this is IEEE == on the six state floats (NaN != NaN, -0 == +0).  SURVEY F8. */
static inline int boid_eq(const boid *a, const boid *b) {
    return a->pos.x == b->pos.x && a->pos.y == b->pos.y && a->pos.z == b->pos.z &&
           a->vel.x == b->vel.x && a->vel.y == b->vel.y && a->vel.z == b->vel.z;
}

void orc_pair_accel(const float *self6, const float *other_pos3, const float *other_vel3,
                    float other_weight, float f_a, float f_c, float f_v, float thr, float fall,
                    float max_sight_angle, float *out3) {
    boid s = boid_load(self6);
    v3_store(out3, boid_get_acceleration(&s, v3_load(other_pos3), v3_load(other_vel3),
                                         other_weight, f_a, f_c, f_v, thr, fall, max_sight_angle));
}

float orc_sight_angle(const float *self6, const float *other_pos3) {
    boid s = boid_load(self6);
    return boid_sight_angle(&s, v3_load(other_pos3));
}

float orc_distance(const float *self6, const float *other_pos3) {
    boid s = boid_load(self6);
    return boid_distance(&s, v3_load(other_pos3));
}

/* ------------------------------------------------------------------------- */
/* point_attractor.rs:16-19, consts.rs:1                                      */
/* ------------------------------------------------------------------------- */
#define GRAVITY 9.8f
static inline v3 attractor_accel(const float *attr4, v3 position, float mass) {
    v3 r = v3_sub(position, v3_load(attr4));
    /* -GRAVITY * (self.mass + mass) / |r|.powi(2) * normalize(r) */
    float s = (-GRAVITY * (attr4[3] + mass)) / f32_square(v3_magnitude(r));
    return s_times_v3(s, v3_normalize(r));
}
void orc_attractor_accel(const float *attr4, const float *pos3, float mass, float *out3) {
    v3_store(out3, attractor_accel(attr4, v3_load(pos3), mass));
}

/* ------------------------------------------------------------------------- */
/* bounding_box.rs:13-26                                                      */
/* ------------------------------------------------------------------------- */
static inline v3 bbox_accel(const float *b, v3 p) {
    float force_x_top = -1.0f / f32_square(b[1] - p.x);
    float force_x_bottom = 1.0f / f32_square(b[0] - p.x);
    float force_y_right = -1.0f / f32_square(b[3] - p.y);
    float force_y_left = 1.0f / f32_square(b[2] - p.y);
    float force_z_front = -1.0f / f32_square(b[5] - p.z);
    float force_z_back = 1.0f / f32_square(b[4] - p.z);
    /* note z: back + front */
    return v3_new(force_x_top + force_x_bottom, force_y_right + force_y_left,
                  force_z_back + force_z_front);
}
void orc_bbox_accel(const float *bbox6, const float *pos3, float *out3) {
    v3_store(out3, bbox_accel(bbox6, v3_load(pos3)));
}

/* ------------------------------------------------------------------------- */
/* obstacle.rs + flocking.rs:182-209                                          */
/* ------------------------------------------------------------------------- */
/* obstacle.rs:63-73.  num-traits 0.2.15 Signed::is_positive for f32 is
 * is_sign_positive: sign bit clear (+0.0 and +NaN count) [ext]. */
static inline int obstacle_will_collide_with_plane(const float *obs4, const boid *b) {
    v3 op = v3_load(obs4);
    v3 normal = v3_normalize(v3_sub(b->pos, op));
    float denom = v3_dot(normal, b->vel);
    if (fabsf(denom) > F32_EPSILON) {
        float t = v3_dot(v3_sub(op, b->pos), normal) / denom;
        if (!signbit(t)) return 1;
    }
    return 0;
}

/* obstacle.rs:76-81 */
static inline void obstacle_velocity_components(const float *obs4, const boid *b, v3 *vi, v3 *vt) {
    v3 dir = v3_normalize(v3_sub(v3_load(obs4), b->pos));
    *vi = s_times_v3(v3_dot(dir, b->vel), dir);
    *vt = v3_sub(b->vel, *vi);
}

/* obstacle.rs:20-28.  returns 1 = Some(*out), 0 = None; raises *flags where
 * Duration::from_secs_f32 panics (then treats the obstacle as None). */
static inline int obstacle_time_to_plane(const float *obs4, const boid *b, dur *out,
                                         uint32_t *flags) {
    if (!obstacle_will_collide_with_plane(obs4, b)) return 0;
    v3 vi, vt;
    obstacle_velocity_components(obs4, b, &vi, &vt);
    float t = (v3_magnitude(v3_sub(v3_load(obs4), b->pos)) - obs4[3]) / v3_magnitude(vi);
    uint32_t f = orc_duration_from_secs_f32(t, &out->secs, &out->nanos);
    if (f) {
        *flags |= f;
        return 0;
    }
    return 1;
}

/* obstacle.rs:31-46 */
static inline v3 obstacle_accel_to_avoid(const float *obs4, const boid *b, uint32_t *flags) {
    v3 vi, vt;
    obstacle_velocity_components(obs4, b, &vi, &vt);
    dur T;
    if (!obstacle_time_to_plane(obs4, b, &T, flags)) return v3_zero();
    float t = orc_duration_as_secs_f32(T.secs, T.nanos);
    if (t * v3_magnitude(vt) > obs4[3]) return v3_zero();
    float s = (2.0f * (obs4[3] - t * v3_magnitude(vt))) / f32_square(t);
    return s_times_v3(s, v3_normalize(vt));
}

/* flocking.rs:182-209.  Iterator::min_by keeps the FIRST minimum.  DECLARED
 * BEHAVIOUR where the reference panics (any obstacle whose plane is hit with a
 * negative / NaN / overflowing time): raise the flag and return
 * zero steering for this boid. */
static v3 steering_accel(const orc_config *cfg, const orc_scene *sc, const boid *b,
                         uint32_t *flags) {
    if (sc == NULL || sc->obstacles == NULL || sc->n_obstacles == 0) return v3_zero();
    uint32_t local = 0;
    uint32_t best = 0;
    dur best_t = DUR_MAX;
    int best_some = 0;
    for (uint32_t k = 0; k < sc->n_obstacles; ++k) {
        dur t = DUR_MAX;
        int some = obstacle_time_to_plane(sc->obstacles + 4 * k, b, &t, &local);
        if (!some) t = DUR_MAX;
        if (k == 0 || dur_cmp(t, best_t) < 0) { /* strictly less: first minimum wins */
            best = k;
            best_t = t;
            best_some = some;
        }
    }
    if (local) {
        *flags |= local;
        return v3_zero();
    }
    if (best_some) {
        dur start = {cfg->time_to_start_steering_secs, cfg->time_to_start_steering_nanos};
        if (dur_cmp(best_t, start) < 0)
            return obstacle_accel_to_avoid(sc->obstacles + 4 * best, b, flags);
    }
    return v3_zero();
}

void orc_steering_accel(const orc_config *cfg, const orc_scene *scene, const float *boid6,
                        float *out3, uint32_t *flags) {
    boid b = boid_load(boid6);
    uint32_t f = 0;
    v3_store(out3, steering_accel(cfg, scene, &b, &f));
    if (flags) *flags = f;
}

/* ------------------------------------------------------------------------- */
/* flocking.rs:101-114, :133-180: per-boid acceleration                       */
/* ------------------------------------------------------------------------- */
typedef struct {
    v3 boids, lead, attr, bbox, steer, total;
} accel_parts;

/* flocking.rs:153-170 */
static v3 accel_from_leads(const orc_config *cfg, const orc_scene *sc, const boid *b) {
    v3 total = v3_zero();
    if (sc && sc->leads) {
        for (uint32_t k = 0; k < sc->n_leads; ++k) {
            const float *l = sc->leads + 7 * k;
            total = v3_add(total, boid_get_acceleration(
                                      b, v3_load(l), v3_load(l + 3), l[6], cfg->avoidance_factor,
                                      cfg->centering_factor, cfg->velocity_matching_factor,
                                      cfg->distance_weight_threshold,
                                      cfg->distance_weight_threshold_falloff,
                                      cfg->max_sight_angle_to_lead_boid));
        }
    }
    return total;
}

/* flocking.rs:172-180; FlockingBoid mass is 1.0 (boid.rs:86) */
static v3 accel_from_attractors(const orc_scene *sc, const boid *b) {
    v3 total = v3_zero();
    if (sc && sc->attractors) {
        for (uint32_t k = 0; k < sc->n_attractors; ++k)
            total = v3_add(total, attractor_accel(sc->attractors + 4 * k, b->pos, 1.0f));
    }
    return total;
}

/* everything except the boid-boid sum; fills parts->{lead,attr,bbox,steer} */
static void accel_extras(const orc_config *cfg, const orc_scene *sc, const boid *b,
                         accel_parts *parts, uint32_t *flags) {
    parts->lead = accel_from_leads(cfg, sc, b);
    parts->attr = accel_from_attractors(sc, b);
    parts->bbox = (sc && sc->bbox) ? bbox_accel(sc->bbox, b->pos) : v3_zero();
    parts->steer = steering_accel(cfg, sc, b, flags);
}

/* flocking.rs:102-114 */
static void accel_combine(const orc_config *cfg, accel_parts *p) {
    if (cfg->steering_overrides)
        p->total = p->steer;
    else
        p->total = v3_add(v3_add(v3_add(v3_add(p->boids, p->lead), p->attr), p->bbox), p->steer);
}

/* flocking.rs:133-151: literal O(N) row */
static v3 accel_from_boids_literal(const orc_config *cfg, uint64_t n, const float *state6,
                                   const boid *b) {
    v3 total = v3_zero();
    for (uint64_t j = 0; j < n; ++j) {
        boid o = boid_load(state6 + 6 * j);
        if (boid_eq(&o, b)) continue;
        total = v3_add(total, boid_get_acceleration(
                                  b, o.pos, o.vel, 1.0f, cfg->avoidance_factor,
                                  cfg->centering_factor, cfg->velocity_matching_factor,
                                  cfg->distance_weight_threshold,
                                  cfg->distance_weight_threshold_falloff, cfg->max_sight_angle));
    }
    return total;
}

static void parts_store(const accel_parts *p, uint64_t r, float *total3, float *comp15) {
    v3_store(total3 + 3 * r, p->total);
    if (comp15) {
        float *c = comp15 + 15 * r;
        v3_store(c, p->boids);
        v3_store(c + 3, p->lead);
        v3_store(c + 6, p->attr);
        v3_store(c + 9, p->bbox);
        v3_store(c + 12, p->steer);
    }
}

void orc_accel_rows(const orc_config *cfg, const orc_scene *scene, uint64_t n,
                    const float *state6, uint64_t i0, uint64_t i1, float *total3, float *comp15,
                    uint32_t *flags, int threads) {
    int64_t rows = (int64_t)(i1 - i0);
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 1 ? threads : 1)
    for (int64_t r = 0; r < rows; ++r) {
        boid b = boid_load(state6 + 6 * (i0 + (uint64_t)r));
        accel_parts p;
        uint32_t f = 0;
        /* the reference evaluates the boid-boid sum even when steering overrides?  No:
         * flocking.rs:102-103 short-circuits; the debug components are still useful. */
        p.boids = accel_from_boids_literal(cfg, n, state6, &b);
        accel_extras(cfg, scene, &b, &p, &f);
        accel_combine(cfg, &p);
        parts_store(&p, (uint64_t)r, total3, comp15);
        if (flags) flags[r] = f;
    }
}

/* flocking.rs:97-122 (Jacobi update; explicit Euler, :116-117) */
static inline void euler_store(const orc_config *cfg, const boid *b, v3 a, float *out6) {
    v3_store(out6, v3_add(b->pos, s_times_v3(cfg->dt, b->vel)));
    v3_store(out6 + 3, v3_add(b->vel, s_times_v3(cfg->dt, a)));
}

void orc_step(const orc_config *cfg, const orc_scene *scene, uint64_t n, const float *state_in6,
              float *state_out6, uint32_t *flags, int threads) {
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 1 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        boid b = boid_load(state_in6 + 6 * i);
        accel_parts p;
        uint32_t f = 0;
        if (cfg->steering_overrides) {
            p.boids = p.lead = p.attr = p.bbox = v3_zero();
            p.steer = steering_accel(cfg, scene, &b, &f);
        } else {
            p.boids = accel_from_boids_literal(cfg, n, state_in6, &b);
            accel_extras(cfg, scene, &b, &p, &f);
        }
        accel_combine(cfg, &p);
        euler_store(cfg, &b, p.total, state_out6 + 6 * i);
        if (flags) flags[i] = f;
    }
}

/* ------------------------------------------------------------------------- */
/* neighbour sets and pair census                                             */
/* ------------------------------------------------------------------------- */
uint64_t orc_mix64(uint64_t j) { /* splitmix64 finaliser */
    uint64_t z = j + 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* 0 = rejected by distance, 1 = in range but FOV-culled, 2 = contributes
 * (member of N(i)), 3 = skipped as equal */
static inline int pair_outcome(const orc_config *cfg, const boid *b, const boid *o) {
    if (boid_eq(o, b)) return 3;
    float dist = boid_distance(b, o->pos);
    float thr = cfg->distance_weight_threshold;
    int in_range = (dist <= thr) || !(dist >= thr + cfg->distance_weight_threshold_falloff);
    if (!in_range) return 0;
    if (boid_sight_angle(b, o->pos) > cfg->max_sight_angle) return 1;
    return 2;
}

void orc_neighbors_rows(const orc_config *cfg, uint64_t n, const float *state6, uint64_t i0,
                        uint64_t i1, uint32_t *count, uint64_t *hash, uint32_t *list,
                        uint32_t list_cap, int threads) {
    int64_t rows = (int64_t)(i1 - i0);
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads > 1 ? threads : 1)
    for (int64_t r = 0; r < rows; ++r) {
        boid b = boid_load(state6 + 6 * (i0 + (uint64_t)r));
        uint32_t c = 0;
        uint64_t h = 0;
        for (uint64_t j = 0; j < n; ++j) {
            boid o = boid_load(state6 + 6 * j);
            if (pair_outcome(cfg, &b, &o) != 2) continue;
            if (list && c < list_cap) list[(uint64_t)r * list_cap + c] = (uint32_t)j;
            ++c;
            h += orc_mix64(j);
        }
        count[r] = c;
        if (hash) hash[r] = h;
    }
}

void orc_pair_census(const orc_config *cfg, uint64_t n, const float *state6, uint64_t i0,
                     uint64_t i1, uint64_t *out3, int threads) {
    uint64_t c0 = 0, c1 = 0, c2 = 0;
    int64_t rows = (int64_t)(i1 - i0);
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : c0, c1, c2) \
    num_threads(threads > 1 ? threads : 1)
    for (int64_t r = 0; r < rows; ++r) {
        boid b = boid_load(state6 + 6 * (i0 + (uint64_t)r));
        for (uint64_t j = 0; j < n; ++j) {
            boid o = boid_load(state6 + 6 * j);
            int k = pair_outcome(cfg, &b, &o);
            c0 += (k == 0);
            c1 += (k == 1);
            c2 += (k == 2);
        }
    }
    out3[0] = c0;
    out3[1] = c1;
    out3[2] = c2;
}

/* ------------------------------------------------------------------------- */
/* grid-accelerated mode (NOT in the reference): same results      */
/* ------------------------------------------------------------------------- */
struct orc_grid {
    double origin[3];
    double inv_cell;
    int64_t dim[3];
    uint64_t n;
    uint64_t *cell_start; /* ncells + 1 */
    uint32_t *cell_items; /* n, ascending j inside each cell */
};

static inline int64_t grid_coord(const orc_grid *g, int axis, float x) {
    double u = floor(((double)x - g->origin[axis]) * g->inv_cell);
    if (!(u >= 0.0)) u = 0.0; /* also NaN */
    if (u > (double)(g->dim[axis] - 1)) u = (double)(g->dim[axis] - 1);
    return (int64_t)u;
}

orc_grid *orc_grid_build(const orc_config *cfg, uint64_t n, const float *state6) {
    orc_grid *g = (orc_grid *)calloc(1, sizeof(orc_grid));
    float thr = cfg->distance_weight_threshold;
    float r = thr + cfg->distance_weight_threshold_falloff;
    double reach = (double)(r > thr ? r : thr);
    if (!(reach > 1e-6)) reach = 1e-6;
    double cell = reach * 1.001; /* > reach, so in-range pairs are <= 1 cell apart */
    double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint64_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            double x = state6[6 * i + a];
            if (isfinite(x)) {
                if (x < lo[a]) lo[a] = x;
                if (x > hi[a]) hi[a] = x;
            }
        }
    g->n = n;
    int64_t ncells = 1;
    for (int a = 0; a < 3; ++a) {
        if (!(lo[a] <= hi[a])) lo[a] = hi[a] = 0.0;
        /* cap the table at ~2^24 cells by coarsening (still >= reach) */
        g->origin[a] = lo[a];
    }
    for (;;) {
        ncells = 1;
        for (int a = 0; a < 3; ++a) {
            g->dim[a] = (int64_t)floor((hi[a] - lo[a]) / cell) + 1;
            ncells *= g->dim[a];
        }
        if (ncells <= (1ll << 24)) break;
        cell *= 1.26;
    }
    g->inv_cell = 1.0 / cell;
    g->cell_start = (uint64_t *)calloc((size_t)ncells + 1, sizeof(uint64_t));
    g->cell_items = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    uint32_t *cell_of = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < n; ++i) {
        int64_t cx = grid_coord(g, 0, state6[6 * i]), cy = grid_coord(g, 1, state6[6 * i + 1]),
                cz = grid_coord(g, 2, state6[6 * i + 2]);
        uint32_t c = (uint32_t)((cz * g->dim[1] + cy) * g->dim[0] + cx);
        cell_of[i] = c;
        g->cell_start[c + 1]++;
    }
    for (int64_t c = 0; c < ncells; ++c) g->cell_start[c + 1] += g->cell_start[c];
    uint64_t *fill = (uint64_t *)malloc(((size_t)ncells + 1) * sizeof(uint64_t));
    memcpy(fill, g->cell_start, ((size_t)ncells + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) g->cell_items[fill[cell_of[i]]++] = (uint32_t)i;
    free(fill);
    free(cell_of);
    return g;
}

void orc_grid_free(orc_grid *g) {
    if (!g) return;
    free(g->cell_start);
    free(g->cell_items);
    free(g);
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/* candidates of boid b from the 27 surrounding cells, ascending j */
static uint32_t grid_candidates(const orc_grid *g, const boid *b, uint32_t **buf, uint32_t *cap) {
    int64_t cx = grid_coord(g, 0, b->pos.x), cy = grid_coord(g, 1, b->pos.y),
            cz = grid_coord(g, 2, b->pos.z);
    uint32_t m = 0;
    for (int64_t z = cz - 1; z <= cz + 1; ++z) {
        if (z < 0 || z >= g->dim[2]) continue;
        for (int64_t y = cy - 1; y <= cy + 1; ++y) {
            if (y < 0 || y >= g->dim[1]) continue;
            for (int64_t x = cx - 1; x <= cx + 1; ++x) {
                if (x < 0 || x >= g->dim[0]) continue;
                uint64_t c = (uint64_t)((z * g->dim[1] + y) * g->dim[0] + x);
                for (uint64_t k = g->cell_start[c]; k < g->cell_start[c + 1]; ++k) {
                    if (m == *cap) {
                        *cap = *cap ? *cap * 2 : 1024;
                        *buf = (uint32_t *)realloc(*buf, *cap * sizeof(uint32_t));
                    }
                    (*buf)[m++] = g->cell_items[k];
                }
            }
        }
    }
    qsort(*buf, m, sizeof(uint32_t), cmp_u32);
    return m;
}

static v3 accel_from_boids_grid(const orc_grid *g, const orc_config *cfg, const float *state6,
                                const boid *b, uint32_t **buf, uint32_t *cap) {
    uint32_t m = grid_candidates(g, b, buf, cap);
    v3 total = v3_zero();
    for (uint32_t k = 0; k < m; ++k) {
        boid o = boid_load(state6 + 6 * (uint64_t)(*buf)[k]);
        if (boid_eq(&o, b)) continue;
        total = v3_add(total, boid_get_acceleration(
                                  b, o.pos, o.vel, 1.0f, cfg->avoidance_factor,
                                  cfg->centering_factor, cfg->velocity_matching_factor,
                                  cfg->distance_weight_threshold,
                                  cfg->distance_weight_threshold_falloff, cfg->max_sight_angle));
    }
    return total;
}

void orc_grid_accel_rows(const orc_grid *g, const orc_config *cfg, const orc_scene *scene,
                         uint64_t n, const float *state6, uint64_t i0, uint64_t i1,
                         float *total3, float *comp15, uint32_t *flags, int threads) {
    (void)n;
    int64_t rows = (int64_t)(i1 - i0);
#pragma omp parallel num_threads(threads > 1 ? threads : 1)
    {
        uint32_t *buf = NULL, cap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < rows; ++r) {
            boid b = boid_load(state6 + 6 * (i0 + (uint64_t)r));
            accel_parts p;
            uint32_t f = 0;
            p.boids = accel_from_boids_grid(g, cfg, state6, &b, &buf, &cap);
            accel_extras(cfg, scene, &b, &p, &f);
            accel_combine(cfg, &p);
            parts_store(&p, (uint64_t)r, total3, comp15);
            if (flags) flags[r] = f;
        }
        free(buf);
    }
}

void orc_grid_neighbors_rows(const orc_grid *g, const orc_config *cfg, uint64_t n,
                             const float *state6, uint64_t i0, uint64_t i1, uint32_t *count,
                             uint64_t *hash, int threads) {
    (void)n;
    int64_t rows = (int64_t)(i1 - i0);
#pragma omp parallel num_threads(threads > 1 ? threads : 1)
    {
        uint32_t *buf = NULL, cap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < rows; ++r) {
            boid b = boid_load(state6 + 6 * (i0 + (uint64_t)r));
            uint32_t m = grid_candidates(g, &b, &buf, &cap);
            uint32_t c = 0;
            uint64_t h = 0;
            for (uint32_t k = 0; k < m; ++k) {
                boid o = boid_load(state6 + 6 * (uint64_t)buf[k]);
                if (pair_outcome(cfg, &b, &o) != 2) continue;
                ++c;
                h += orc_mix64(buf[k]);
            }
            count[r] = c;
            if (hash) hash[r] = h;
        }
        free(buf);
    }
}

void orc_grid_step(const orc_config *cfg, const orc_scene *scene, uint64_t n,
                   const float *state_in6, float *state_out6, uint32_t *flags, int threads) {
    orc_grid *g = cfg->steering_overrides ? NULL : orc_grid_build(cfg, n, state_in6);
#pragma omp parallel num_threads(threads > 1 ? threads : 1)
    {
        uint32_t *buf = NULL, cap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < (int64_t)n; ++i) {
            boid b = boid_load(state_in6 + 6 * i);
            accel_parts p;
            uint32_t f = 0;
            if (cfg->steering_overrides) {
                p.boids = p.lead = p.attr = p.bbox = v3_zero();
                p.steer = steering_accel(cfg, scene, &b, &f);
            } else {
                p.boids = accel_from_boids_grid(g, cfg, state_in6, &b, &buf, &cap);
                accel_extras(cfg, scene, &b, &p, &f);
            }
            accel_combine(cfg, &p);
            euler_store(cfg, &b, p.total, state_out6 + 6 * i);
            if (flags) flags[i] = f;
        }
        free(buf);
    }
    orc_grid_free(g);
}

/* ------------------------------------------------------------------------- */
/* lead boids: boid.rs:46-53, parametric.rs:17-21                             */
/* ------------------------------------------------------------------------- */
uint32_t orc_lead_step(float *lead7, float *curr_time, float dt, orc_path_fn path, void *ctx) {
    /* flocking.rs:126: lead_boid.step(Duration::from_secs_f32(self.config.dt)) */
    dur D;
    uint32_t f = orc_duration_from_secs_f32(dt, &D.secs, &D.nanos);
    if (f) return f;
    if (D.secs == 0 && D.nanos == 0) return 0;          /* dt.is_zero() */
    float dts = orc_duration_as_secs_f32(D.secs, D.nanos); /* dt.as_secs_f32() */
    float np[3];
    path(*curr_time, np, ctx);     /* parametric.rs:18 */
    *curr_time = *curr_time + dts; /* parametric.rs:19 */
    v3 newp = v3_load(np);
    v3 vel = v3_div(v3_sub(newp, v3_load(lead7)), dts); /* boid.rs:51 */
    v3_store(lead7 + 3, vel);
    v3_store(lead7, newp);
    return 0;
}

/* demos/flocking.rs:105-107 (kind 0), :139-145 (kind 1), :146-148 (kind 2) */
void orc_demo_path(int kind, float t, float *out3) {
    switch (kind) {
    case 0:
        out3[0] = 25.0f * cosf(t / 12.0f);
        out3[1] = 0.5f;
        out3[2] = 0.0f;
        break;
    case 1:
        out3[0] = 15.0f * cosf(t / 12.0f);
        out3[1] = 6.0f + 5.0f * cosf(t / 12.0f);
        out3[2] = 15.0f * sinf(t / 12.0f);
        break;
    default:
        out3[0] = 25.0f * cosf(t / 10.0f);
        out3[1] = 1.0f;
        out3[2] = 10.0f * sinf(t / 9.0f);
        break;
    }
}

/* ------------------------------------------------------------------------- */
/* state.rs:75-106, utils.rs:5-21                                             */
/* ------------------------------------------------------------------------- */
void orc_state_euler(const float *s, size_t n, float h, orc_deriv_fn deriv, void *ctx,
                     float *out) {
    float *d = (float *)malloc((n ? n : 1) * sizeof(float));
    deriv(s, d, n, ctx);
    for (size_t i = 0; i < n; ++i) out[i] = s[i] + d[i] * h; /* state.rs:79,81 */
    free(d);
}

void orc_state_rk4(const float *s, size_t n, float h, orc_deriv_fn deriv, void *ctx, float *out) {
    size_t m = n ? n : 1;
    float *k1 = (float *)malloc(m * sizeof(float)), *k2 = (float *)malloc(m * sizeof(float)),
          *k3 = (float *)malloc(m * sizeof(float)), *k4 = (float *)malloc(m * sizeof(float)),
          *tmp = (float *)malloc(m * sizeof(float));
    deriv(s, k1, n, ctx);                                       /* :87 */
    for (size_t i = 0; i < n; ++i) tmp[i] = s[i] + k1[i] * (h * 0.5f); /* :88-89 */
    deriv(tmp, k2, n, ctx);
    for (size_t i = 0; i < n; ++i) tmp[i] = s[i] + k2[i] * (h * 0.5f); /* :91-92 */
    deriv(tmp, k3, n, ctx);
    for (size_t i = 0; i < n; ++i) tmp[i] = s[i] + k3[i] * h; /* :94-95 */
    deriv(tmp, k4, n, ctx);
    for (size_t i = 0; i < n; ++i) { /* :97-105 */
        float delta = ((h / 6.0f * k1[i] + h / 3.0f * k2[i]) + h / 3.0f * k3[i]) + h / 6.0f * k4[i];
        out[i] = s[i] + delta;
    }
    free(k1);
    free(k2);
    free(k3);
    free(k4);
    free(tmp);
}

/* state.rs:139-152: Point { position, velocity }, constant accel (1,-1,0) */
void orc_deriv_test_point(const float *s, float *ds, size_t n, void *ctx) {
    (void)ctx;
    for (size_t e = 0; e + 6 <= n; e += 6) {
        ds[e] = s[e + 3];
        ds[e + 1] = s[e + 4];
        ds[e + 2] = s[e + 5];
        ds[e + 3] = 1.0f;
        ds[e + 4] = -1.0f;
        ds[e + 5] = 0.0f;
    }
}

/* state.rs:209-211: ExampleFn { y, t, timestep }: y' = y - t^2 + 1, t' = 1 */
void orc_deriv_test_examplefn(const float *s, float *ds, size_t n, void *ctx) {
    (void)ctx;
    for (size_t e = 0; e + 3 <= n; e += 3) {
        ds[e] = s[e] - f32_square(s[e + 1]) + 1.0f;
        ds[e + 1] = 1.0f;
        ds[e + 2] = 0.0f;
    }
}

/* ------------------------------------------------------------------------- */
/* acosf threshold form of the FOV predicate                */
/* ------------------------------------------------------------------------- */
static inline uint32_t f32_bits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float bits_f32(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
/* order-preserving map float -> int64 (no NaN) */
static inline int64_t f32_ordinal(float f) {
    uint32_t u = f32_bits(f);
    return (u & 0x80000000u) ? -(int64_t)(u & 0x7fffffffu) : (int64_t)u;
}
static inline float ordinal_f32(int64_t o) {
    return o < 0 ? bits_f32(0x80000000u | (uint32_t)(-o)) : bits_f32((uint32_t)o);
}

float orc_acos_threshold(float theta) {
    if (!(acosf(-1.0f) > theta)) return -2.0f; /* nothing is ever culled (also NaN theta) */
    if (acosf(1.0f) > theta) return 1.0f;
    int64_t lo = f32_ordinal(-1.0f), hi = f32_ordinal(1.0f); /* acos(lo) > theta, !(acos(hi) > theta) */
    while (hi - lo > 1) {
        int64_t mid = lo + (hi - lo) / 2;
        if (acosf(ordinal_f32(mid)) > theta)
            lo = mid;
        else
            hi = mid;
    }
    return ordinal_f32(lo);
}

uint64_t orc_acos_monotone_violations(uint32_t lo_bits, uint32_t hi_bits) {
    /* walk the closed bit range; within one sign the bit pattern orders |x| */
    uint64_t bad = 0;
    int neg = (lo_bits & 0x80000000u) != 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t u = (int64_t)lo_bits; u < (int64_t)hi_bits; ++u) {
        float a = acosf(bits_f32((uint32_t)u)), b = acosf(bits_f32((uint32_t)u + 1));
        /* positive: value rises with u => acos must not rise; negative: value falls with u */
        if (neg ? (b < a) : (b > a)) ++bad;
    }
    return bad;
}

/* ------------------------------------------------------------------------- */
/* flocking.rs:230-245 instance export; cgmath 0.18 Quaternion::from_arc [ext] */
/* ------------------------------------------------------------------------- */
/* approx 0.4.0 ulps_eq! defaults: epsilon = f32::EPSILON, max_ulps = 4 */
static int f32_ulps_eq(float a, float b) {
    if (f32_abs_diff_eq(a, b)) return 1;
    if (signbit(a) != signbit(b)) return 0; /* a.signum() != b.signum(); NaN handled below */
    if (isnan(a) || isnan(b)) return 0;
    int32_t ia = (int32_t)f32_bits(a), ib = (int32_t)f32_bits(b);
    int64_t d = (int64_t)ia - (int64_t)ib;
    if (d < 0) d = -d;
    return d <= 4;
}

void orc_instances(uint64_t n, const float *state6, float *out8) {
    const v3 src = {0.0f, 0.0f, 1.0f}; /* Vector3::unit_z() */
    for (uint64_t i = 0; i < n; ++i) {
        v3 dst = v3_normalize(v3_load(state6 + 6 * i + 3));
        float qs, qx, qy, qz;
        /* from_arc: mag_avg = sqrt(|src|^2 * |dst|^2); dot = src . dst */
        float mag_avg = sqrtf(v3_dot(src, src) * v3_dot(dst, dst));
        float dot = v3_dot(src, dst);
        if (f32_ulps_eq(dot, mag_avg)) {
            qs = 1.0f;
            qx = qy = qz = 0.0f;
        } else if (f32_ulps_eq(dot, -mag_avg)) {
            /* fallback None: axis = unit_x x src, or unit_y x src if that is ~0; normalised;
             * from_axis_angle(axis, Rad::turn_div_2()) = (cos(pi/2), axis * sin(pi/2)) */
            v3 ax = v3_cross(v3_new(1.0f, 0.0f, 0.0f), src);
            if (f32_ulps_eq(ax.x, 0.0f) && f32_ulps_eq(ax.y, 0.0f) && f32_ulps_eq(ax.z, 0.0f))
                ax = v3_cross(v3_new(0.0f, 1.0f, 0.0f), src);
            ax = v3_normalize(ax);
            float half = 3.14159274101257324f * 0.5f;
            float sn = sinf(half), cs = cosf(half);
            qs = cs;
            qx = ax.x * sn;
            qy = ax.y * sn;
            qz = ax.z * sn;
        } else {
            /* Quaternion::from_sv(mag_avg + dot, src x dst).normalize() */
            float s = mag_avg + dot;
            v3 v = v3_cross(src, dst);
            /* Quaternion magnitude2 = s*s + v.dot(v); normalize = q * (1/magnitude) */
            float inv = 1.0f / sqrtf(s * s + v3_dot(v, v));
            qs = s * inv;
            qx = v.x * inv;
            qy = v.y * inv;
            qz = v.z * inv;
        }
        float *o = out8 + 8 * i;
        o[0] = state6[6 * i];
        o[1] = state6[6 * i + 1];
        o[2] = state6[6 * i + 2];
        o[3] = qs;
        o[4] = qx;
        o[5] = qy;
        o[6] = qz;
        o[7] = 0.1f;
    }
}

/*
 * next_oracle.c -- CPU restatements for the rows SURVEY.md 8(f) marks "next": the reference's
 * Stateful element types (springy Point, rigid-body State) for the generic State<T> integrator,
 * and the SPH neighbour pass.
 *
 * TEST INFRASTRUCTURE ONLY (see flock_oracle.h).  Plain C, one separately rounded binary32
 * operation per source operation, in the source's order (-ffp-contract=off).
 *
 * PARITY STATUS: unpinned.  The reference holds no test, fixture or golden vector for
 * Point::derivative, rigidbody::State::derivative or sph::Simulation::step, and cannot be built
 * here (no Rust toolchain).  What pins the integrator itself are the two State tests
 * (state.rs:166-185, :218-280), checked in tests/test_oracle_state.py and tests/test_gpu_state.py.
 * Semantics of un-vendored dependencies are declared where they are restated:
 *   cgmath 0.18.0  Matrix3::from(Quaternion), Matrix3 * Matrix3, Matrix3 * Vector3,
 *                  f32 * Quaternion, Quaternion * Quaternion, InnerSpace::dot / magnitude
 *   kiddo 0.2.4    KdTree::nearest (exact k nearest, ascending distance),
 *                  distance::squared_euclidean
 *   Rust std       f32::powi(2|3) -> repeated multiplication
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "flock_oracle.h"

/* springy_mesh.rs:223-240: [mass, p3, v3, accumulated_force3] */
void orc_deriv_springy_point(const float *s, float *d, size_t n, void *ctx) {
    (void)ctx;
    for (size_t e = 0; e + 10 <= n; e += 10) {
        const float *p = s + e;
        float *o = d + e;
        o[0] = 0.0f;
        o[1] = p[4]; o[2] = p[5]; o[3] = p[6];
        o[4] = p[7] / p[0]; o[5] = p[8] / p[0]; o[6] = p[9] / p[0];
        o[7] = o[8] = o[9] = 0.0f;
    }
}

/* [p3 v3 a3]: the Point pattern for a FlockingBoid with its acceleration frozen */
void orc_deriv_boid(const float *s, float *d, size_t n, void *ctx) {
    (void)ctx;
    for (size_t e = 0; e + 9 <= n; e += 9) {
        for (int i = 0; i < 6; ++i) d[e + i] = s[e + 3 + i];
        d[e + 6] = d[e + 7] = d[e + 8] = 0.0f;
    }
}

/* cgmath dot [ext]: (x x' + y y') + z z' */
static float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return (ax * bx + ay * by) + az * bz;
}
/* cgmath Matrix3 * Matrix3 [ext]: element (row r, column c) = lhs.row(r).dot(rhs[c]); column-major */
static void mat3_mul(const float *a, const float *b, float *o) {
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r)
            o[3 * c + r] = dot3(a[r], a[3 + r], a[6 + r], b[3 * c], b[3 * c + 1], b[3 * c + 2]);
}

/* rigidbody.rs:53-140.  Layout (as_state, :66-101): p3, rotation v3 then s, linear momentum 3,
 * angular momentum 3, mass, initial inverted inertia (3 columns), force 3, torque 3. */
void orc_deriv_rigidbody(const float *sv, float *dv, size_t n, void *ctx) {
    (void)ctx;
    for (size_t e = 0; e + 29 <= n; e += 29) {
        const float *s = sv + e;
        float *d = dv + e;
        const float qx = s[3], qy = s[4], qz = s[5], qs = s[6], mass = s[13];
        /* velocity() = linear_momentum / mass (:45-47) */
        d[0] = s[7] / mass; d[1] = s[8] / mass; d[2] = s[9] / mass;
        /* Matrix3::from(Quaternion) [ext] */
        const float x2 = qx + qx, y2 = qy + qy, z2 = qz + qz;
        const float xx2 = x2 * qx, xy2 = x2 * qy, xz2 = x2 * qz;
        const float yy2 = y2 * qy, yz2 = y2 * qz, zz2 = z2 * qz;
        const float sy2 = y2 * qs, sz2 = z2 * qs, sx2 = x2 * qs;
        float R[9], Rt[9], RI[9], Iinv[9];
        R[0] = 1.0f - yy2 - zz2; R[1] = xy2 + sz2; R[2] = xz2 - sy2;
        R[3] = xy2 - sz2; R[4] = 1.0f - xx2 - zz2; R[5] = yz2 + sx2;
        R[6] = xz2 + sy2; R[7] = yz2 - sx2; R[8] = 1.0f - xx2 - yy2;
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) Rt[3 * c + r] = R[3 * r + c];
        /* get_moment_of_inertia_inverted (:35-38): R * I0inv * R^T, left to right */
        mat3_mul(R, s + 14, RI);
        mat3_mul(RI, Rt, Iinv);
        /* angular_velocity() (:49-51): Matrix3 * Vector3 [ext] = col0 * x + col1 * y + col2 * z */
        float w[3];
        for (int r = 0; r < 3; ++r) w[r] = (Iinv[r] * s[10] + Iinv[3 + r] * s[11]) + Iinv[6 + r] * s[12];
        /* 0.5 * Quaternion::from_sv(0.0, w) * rotation (:105-106); f32 * Quaternion scales s and v [ext] */
        const float as = 0.5f * 0.0f, ax = 0.5f * w[0], ay = 0.5f * w[1], az = 0.5f * w[2];
        d[6] = as * qs - ax * qx - ay * qy - az * qz;
        d[3] = as * qx + ax * qs + ay * qz - az * qy;
        d[4] = as * qy + ay * qs + az * qx - ax * qz;
        d[5] = as * qz + az * qs + ax * qy - ay * qx;
        d[7] = s[23]; d[8] = s[24]; d[9] = s[25];
        d[10] = s[26]; d[11] = s[27]; d[12] = s[28];
        for (int i = 13; i < 29; ++i) d[i] = 0.0f;
    }
}

/* kernals.rs:6-16 */
static float monaghan(float r, float s) {
    const float q = r / s;
    float num;
    if (q >= 0.0f && q <= 1.0f) num = 1.0f - 1.5f * (q * q) + 0.75f * ((q * q) * q);
    else if (q >= 1.0f && q <= 2.0f) num = 0.25f * (((2.0f - q) * (2.0f - q)) * (2.0f - q));
    else num = 0.0f;
    return num / (3.14159274101257324f * ((s * s) * s));
}

/* sph/mod.rs:89-121: for every particle the k nearest (itself included) by squared_euclidean,
 * ascending, those with d2 < s^2 kept; density = sum of mass * monaghan(r, s) in that order.
 * Equal distances are ordered by particle id (DECLARED: kiddo's order among ties is unspecified).
 * out_index: n x k, 0xffffffff beyond the count.  Brute force, O(n^2). */
void orc_sph_neighbors(uint64_t n, const float *pos3, uint32_t k, float s, float mass, uint32_t *out_index,
                       uint32_t *out_count, float *out_density) {
    const float s2 = s * s;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        float bd[32];
        uint32_t bi[32], cnt = 0;
        const float *p = pos3 + 3 * i;
        for (uint64_t j = 0; j < n; ++j) {
            const float *o = pos3 + 3 * j;
            const float dx = p[0] - o[0], dy = p[1] - o[1], dz = p[2] - o[2];
            const float d2 = ((0.0f + dx * dx) + dy * dy) + dz * dz; /* squared_euclidean [ext] */
            if (!(d2 < s2)) continue;
            if (cnt == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && (uint32_t)j < bi[k - 1]))) continue;
            uint32_t t = cnt < k ? cnt : k - 1;
            while (t > 0 && (d2 < bd[t - 1] || (d2 == bd[t - 1] && (uint32_t)j < bi[t - 1]))) {
                bd[t] = bd[t - 1];
                bi[t] = bi[t - 1];
                --t;
            }
            bd[t] = d2;
            bi[t] = (uint32_t)j;
            if (cnt < k) ++cnt;
        }
        float density = 0.0f;
        for (uint32_t t = 0; t < k; ++t) {
            if (t < cnt) {
                const float r = bd[t] == 0.0f ? 0.0f : sqrtf(bd[t]); /* r_ij.is_zero() ? 0 : magnitude */
                density = density + mass * monaghan(r, s);
                out_index[(size_t)i * k + t] = bi[t];
            } else {
                out_index[(size_t)i * k + t] = 0xffffffffu;
            }
        }
        out_count[i] = cnt;
        out_density[i] = density;
    }
}

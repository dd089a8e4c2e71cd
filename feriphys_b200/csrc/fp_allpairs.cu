// fp_allpairs.cu -- K1: shared-memory tiled all-pairs influence pass fused
// with the per-boid extras and the Euler update.
//
// Replaces Simulation::step's O(N^2) loop (flocking.rs:97-151).  One thread
// owns one boid i and walks j = 0 .. N-1 in ascending index order through
// shared-memory tiles, so each boid's f32 sum is accumulated in exactly the
// order the reference accumulates it: with the exact pair function of
// fp_device.cuh the result is bit-identical to the Rust loop.
//
// FP32-pipe bound: per ordered pair 8 flops when rejected by
// distance, 18 when FOV-culled, 54 when it contributes.  State is 32 B/boid
// and is re-read from shared memory N times, so HBM traffic is negligible.
#include <stdlib.h>

#include "fp_internal.h"

namespace fp {

constexpr int AP_BLOCK = 128;  // threads per CTA == boids per shared-memory tile

template <int TAP>
__global__ void __launch_bounds__(AP_BLOCK)
allpairs_kernel(const DevParams P, const float4 *__restrict__ pos_all,
                const float4 *__restrict__ vel_all, uint32_t n_all, uint32_t row0, uint32_t nrows,
                float4 *__restrict__ pos_out, float4 *__restrict__ vel_out,
                unsigned *__restrict__ status, TapOut tap) {
    __shared__ float4 sp[AP_BLOCK];
    __shared__ float4 sv[AP_BLOCK];

    const uint32_t r = blockIdx.x * AP_BLOCK + threadIdx.x;
    const bool active = r < nrows;
    const uint32_t i = row0 + (active ? r : 0);
    const float4 pi4 = __ldg(pos_all + i);
    const float4 vi4 = __ldg(vel_all + i);
    const Self self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));

    V3 acc = v3zero();
    uint32_t n_count = 0;
    unsigned long long n_hash = 0;
    unsigned long long c_far = 0, c_cull = 0, c_in = 0;

    const bool need_pairs = (TAP != TAP_STEP) || !P.steering_overrides;
    if (need_pairs) {
        for (uint32_t j0 = 0; j0 < n_all; j0 += AP_BLOCK) {
            const uint32_t jl = j0 + threadIdx.x;
            if (jl < n_all) {
                sp[threadIdx.x] = __ldg(pos_all + jl);
                sv[threadIdx.x] = __ldg(vel_all + jl);
            }
            __syncthreads();
            const int cnt = (int)min((uint32_t)AP_BLOCK, n_all - j0);
            if (active) {
#pragma unroll 4
                for (int jj = 0; jj < cnt; ++jj) {
                    const float4 pj = sp[jj];
                    if (TAP == TAP_STEP || TAP == TAP_ACCEL) {
                        V3 d;
                        const float m2 = pair_m2(self, v3(pj.x, pj.y, pj.z), d);
                        if (m2 >= P.m2_cut) continue;
                        const float4 vj = sv[jj];
                        V3 contrib;
                        if (pair_flock(P, self, d, m2, v3(vj.x, vj.y, vj.z), contrib))
                            acc = vadd(acc, contrib);
                    } else {
                        const float4 vj = sv[jj];
                        bool equal;
                        const int o = pair_outcome(P, self, v3(pj.x, pj.y, pj.z),
                                                   v3(vj.x, vj.y, vj.z), equal);
                        if (equal) continue;
                        if (TAP == TAP_NEIGHBORS) {
                            if (o == PAIR_CONTRIB) {
                                ++n_count;
                                n_hash += mix64((unsigned long long)__float_as_uint(pj.w));
                            }
                        } else {
                            c_far += (o == PAIR_FAR);
                            c_cull += (o == PAIR_CULLED);
                            c_in += (o == PAIR_CONTRIB);
                        }
                    }
                }
            }
            __syncthreads();
        }
    }

    if (TAP == TAP_CENSUS) {
        // warp reduce, one atomic per warp per counter
        for (int off = 16; off > 0; off >>= 1) {
            c_far += __shfl_down_sync(0xffffffffu, c_far, off);
            c_cull += __shfl_down_sync(0xffffffffu, c_cull, off);
            c_in += __shfl_down_sync(0xffffffffu, c_in, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(tap.census + 0, c_far);
            atomicAdd(tap.census + 1, c_cull);
            atomicAdd(tap.census + 2, c_in);
            atomicAdd(tap.census + 3, c_far + c_cull + c_in);
        }
        return;
    }
    if (!active) return;

    const uint32_t idx = __float_as_uint(pi4.w);  // caller index
    if (TAP == TAP_NEIGHBORS) {
        tap.nbr_count[idx] = n_count;
        tap.nbr_hash[idx] = n_hash;
        return;
    }

    Extras e;
    unsigned flags = 0;
    const V3 a = accel_total(P, self, acc, e, flags, TAP == TAP_ACCEL);
    if (TAP == TAP_ACCEL) {
        float *o = tap.accel3 + 3ull * idx;
        o[0] = a.x; o[1] = a.y; o[2] = a.z;
        if (tap.comp15) {
            float *c = tap.comp15 + 15ull * idx;
            c[0] = acc.x; c[1] = acc.y; c[2] = acc.z;
            c[3] = e.lead.x; c[4] = e.lead.y; c[5] = e.lead.z;
            c[6] = e.attr.x; c[7] = e.attr.y; c[8] = e.attr.z;
            c[9] = e.bbox.x; c[10] = e.bbox.y; c[11] = e.bbox.z;
            c[12] = e.steer.x; c[13] = e.steer.y; c[14] = e.steer.z;
        }
        if (flags) atomicOr(status, flags);
        return;
    }
    V3 np, nv;
    euler(P, self.p, self.v, a, np, nv);
    pos_out[r] = make_float4(np.x, np.y, np.z, pi4.w);
    vel_out[r] = make_float4(nv.x, nv.y, nv.z, 0.0f);
    if (flags) atomicOr(status, flags);
}

// ---- production form for TAP_STEP / TAP_ACCEL: staged tiles, packed pre-gate, list, drain ----
// Same structure as the staged grid walk (fp_walk.cu): a tile of AP2_TJ candidates is staged in
// shared memory (positions SoA, velocities float4); every thread runs the packed, fused
// PRE-gate over the tile -- squared distance against m2_cut (1 + 1e-6) and the conservative
// field-of-view filter, two candidates per FP32 instruction, a superset of the contributing
// pairs -- and appends the survivors' tile offsets to its list; the list is then drained with
// the EXACT distance test and the exact pair function, in ascending j.  A batch of four
// candidates none of which is in range (the common case in a sparse flock) costs 6 shared
// loads, 12 packed operations and one branch.  Bit-identical to allpairs_kernel.
constexpr int AP2_TJ = 128;   // candidates per tile; also the list capacity (all may survive)

struct Ap2Smem {
    alignas(16) float tx[AP2_TJ], ty[AP2_TJ], tz[AP2_TJ];
    float4 tv[AP2_TJ];
    uint16_t list[AP2_TJ][AP_BLOCK];  // [entry][thread]: bank == lane
};

template <int TAP>
__global__ void __launch_bounds__(AP_BLOCK, 5)
allpairs2_kernel(const DevParams P, const float4 *__restrict__ pos_all,
                 const float4 *__restrict__ vel_all, uint32_t n_all, uint32_t row0, uint32_t nrows,
                 float4 *__restrict__ pos_out, float4 *__restrict__ vel_out,
                 unsigned *__restrict__ status, TapOut tap) {
    extern __shared__ __align__(16) unsigned char ap2_raw[];
    Ap2Smem &S = *reinterpret_cast<Ap2Smem *>(ap2_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t r = blockIdx.x * AP_BLOCK + tid;
    const bool active = r < nrows;
    const uint32_t i = row0 + (active ? r : 0);
    const float4 pi4 = __ldg(pos_all + i);
    const float4 vi4 = __ldg(vel_all + i);
    const Self self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
    V3 acc = v3zero();
    const bool need_pairs = (TAP != TAP_STEP) || !P.steering_overrides;
    if (need_pairs) {
        const float2 nsx = make_float2(-self.p.x, -self.p.x), nsy = make_float2(-self.p.y, -self.p.y),
                     nsz = make_float2(-self.p.z, -self.p.z);
        const float2 vhx = make_float2(self.vhat.x, self.vhat.x), vhy = make_float2(self.vhat.y, self.vhat.y),
                     vhz = make_float2(self.vhat.z, self.vhat.z);
        const float2 kh2 = make_float2(P.fov_kh, P.fov_kh), kl2 = make_float2(P.fov_kl, P.fov_kl);
        uint16_t *const lst = &S.list[0][tid];
        for (uint32_t j0 = 0; j0 < n_all; j0 += AP2_TJ) {
            const uint32_t jl = j0 + tid;
            float4 pj = make_float4(0, 0, 0, 0), vj = pj;
            if (jl < n_all) {
                pj = __ldg(pos_all + jl);
                vj = __ldg(vel_all + jl);
            }
            __syncthreads();  // the previous tile has been drained by everyone
            S.tx[tid] = pj.x;
            S.ty[tid] = pj.y;
            S.tz[tid] = pj.z;
            S.tv[tid] = vj;
            __syncthreads();
            const uint32_t cnt = min((uint32_t)AP2_TJ, n_all - j0);
            const uint32_t t_self = (i >= j0 && i < j0 + cnt) ? i - j0 : 0xffffu;
            uint32_t w = 0;  // list cursor, in entries * AP_BLOCK
            if (active) {
#pragma unroll 2
                for (uint32_t T = 0; T < cnt; T += 4) {
                    const float2 x01 = *reinterpret_cast<const float2 *>(&S.tx[T]);
                    const float2 x23 = *reinterpret_cast<const float2 *>(&S.tx[T + 2]);
                    const float2 y01 = *reinterpret_cast<const float2 *>(&S.ty[T]);
                    const float2 y23 = *reinterpret_cast<const float2 *>(&S.ty[T + 2]);
                    const float2 z01 = *reinterpret_cast<const float2 *>(&S.tz[T]);
                    const float2 z23 = *reinterpret_cast<const float2 *>(&S.tz[T + 2]);
                    const float2 dx01 = __fadd2_rn(x01, nsx), dx23 = __fadd2_rn(x23, nsx);
                    const float2 dy01 = __fadd2_rn(y01, nsy), dy23 = __fadd2_rn(y23, nsy);
                    const float2 dz01 = __fadd2_rn(z01, nsz), dz23 = __fadd2_rn(z23, nsz);
                    const float2 m01 = __ffma2_rn(dz01, dz01, __ffma2_rn(dy01, dy01, __fmul2_rn(dx01, dx01)));
                    const float2 m23 = __ffma2_rn(dz23, dz23, __ffma2_rn(dy23, dy23, __fmul2_rn(dx23, dx23)));
                    const float mm[4] = {m01.x, m01.y, m23.x, m23.y};
                    bool near[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) near[u] = T + u < cnt && !(mm[u] >= P.m2_cut_hi);
                    if (near[0] | near[1] | near[2] | near[3]) {
                        const float2 q01 = __ffma2_rn(vhz, dz01, __ffma2_rn(vhy, dy01, __fmul2_rn(vhx, dx01)));
                        const float2 q23 = __ffma2_rn(vhz, dz23, __ffma2_rn(vhy, dy23, __fmul2_rn(vhx, dx23)));
                        const float2 s01 = __fmul2_rn(q01, make_float2(fabsf(q01.x), fabsf(q01.y)));
                        const float2 s23 = __fmul2_rn(q23, make_float2(fabsf(q23.x), fabsf(q23.y)));
                        const float2 h01 = __fmul2_rn(kh2, m01), h23 = __fmul2_rn(kh2, m23);
                        const float2 l01 = __fmul2_rn(kl2, m01), l23 = __fmul2_rn(kl2, m23);
                        const float ss[4] = {s01.x, s01.y, s23.x, s23.y};
                        const float hh[4] = {h01.x, h01.y, h23.x, h23.y};
                        const float ll[4] = {l01.x, l01.y, l23.x, l23.y};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (near[u] && !(ss[u] < hh[u] && ss[u] > ll[u])) {
                                lst[w] = (uint16_t)(T + u);
                                w += AP_BLOCK;
                            }
                    }
                }
            }
            // drain: exact distance, exact pair function, ascending j; two entries per trip
            const int nb = (int)(w / AP_BLOCK);
            if (__any_sync(0xffffffffu, nb > 0)) {
                for (int k = 0; k < nb; k += 2) {
                    const bool hasb = k + 1 < nb;
                    const uint32_t ia = lst[k * AP_BLOCK], ib = hasb ? lst[(k + 1) * AP_BLOCK] : ia;
                    const float4 va = S.tv[ia], vb = S.tv[ib];
                    V3 da, db;
                    const float ma = pair_m2(self, v3(S.tx[ia], S.ty[ia], S.tz[ia]), da);
                    const float mb = pair_m2(self, v3(S.tx[ib], S.ty[ib], S.tz[ib]), db);
                    const bool oka = ia != t_self && !(ma >= P.m2_cut);
                    const bool okb = hasb && ib != t_self && !(mb >= P.m2_cut);
                    const bool fast = P.fast_ok && (!oka || (ma >= FAST_M2_LO && ma <= FAST_M2_HI)) &&
                                      (!okb || (mb >= FAST_M2_LO && mb <= FAST_M2_HI));
                    if (fast) {  // (a lane that does not count may hold inf / NaN; it is never used)
                        bool visa, visb;
                        const V3 fa = pair_force_fast(P, self, da, ma, v3(va.x, va.y, va.z), visa);
                        const V3 fb = pair_force_fast(P, self, db, mb, v3(vb.x, vb.y, vb.z), visb);
                        if (oka && visa) acc = vadd(acc, fa);
                        if (okb && visb) acc = vadd(acc, fb);
                    } else {
                        V3 contrib;
                        if (oka && pair_inrange<false>(P, self, da, ma, v3(va.x, va.y, va.z), 1.0f, P.cstar, contrib))
                            acc = vadd(acc, contrib);
                        if (okb && pair_inrange<false>(P, self, db, mb, v3(vb.x, vb.y, vb.z), 1.0f, P.cstar, contrib))
                            acc = vadd(acc, contrib);
                    }
                }
            }
        }
    }
    if (!active) return;
    const uint32_t idx = __float_as_uint(pi4.w);  // caller index
    Extras e;
    unsigned flags = 0;
    const V3 a = accel_total(P, self, acc, e, flags, TAP == TAP_ACCEL);
    if (TAP == TAP_ACCEL) {
        float *o = tap.accel3 + 3ull * idx;
        o[0] = a.x; o[1] = a.y; o[2] = a.z;
        if (tap.comp15) {
            float *c = tap.comp15 + 15ull * idx;
            c[0] = acc.x; c[1] = acc.y; c[2] = acc.z;
            c[3] = e.lead.x; c[4] = e.lead.y; c[5] = e.lead.z;
            c[6] = e.attr.x; c[7] = e.attr.y; c[8] = e.attr.z;
            c[9] = e.bbox.x; c[10] = e.bbox.y; c[11] = e.bbox.z;
            c[12] = e.steer.x; c[13] = e.steer.y; c[14] = e.steer.z;
        }
        if (flags) atomicOr(status, flags);
        return;
    }
    V3 np, nv;
    euler(P, self.p, self.v, a, np, nv);
    pos_out[r] = make_float4(np.x, np.y, np.z, pi4.w);
    vel_out[r] = make_float4(nv.x, nv.y, nv.z, 0.0f);
    if (flags) atomicOr(status, flags);
}

// ---- FAST numerics: j split across lanes, algebraic pre-gate, fused forces -------------------------
// A thread owns TWO boids (A, B) and one of JS interleaved slices of the candidates; the JS partial
// sums of a boid are combined with warp shuffles at the end (JS lanes of one warp), so a 100k-boid
// flock still fills the machine (C2: 64 rows per CTA, 1563 CTAs).  Per batch of four candidates:
//   pre-gate  -- |p_j - p_i|^2 = |p_j|^2 - 2 p_i . p_j + |p_i|^2 with |p_j|^2 staged per candidate and
//                |p_i|^2 folded into the threshold: three FMAs per pair (two pairs per FFMA2) instead
//                of three subtractions and three multiply-adds.  Coordinates are taken relative to
//                the flock's centre and the cancellation error (<= 1.2e-6 M^2, M the largest centred
//                norm) is added to the cut, so the survivors are a superset of the pairs in range;
//   survivors -- the reference's own squared distance (separately rounded) decides "in range" and
//                "weight 1"; the sight-angle decision and the forces are those of the fast grid walk
//                (fp_walk_nl.cu): fused cosine outside a 1e-5 guard band, exact sequence inside it.
// Neighbour sets are bit-exact; accelerations differ from the reference by rounding and summation
// order (~1e-6 relative).
constexpr int APF_TJ = 256;  // candidates per tile

struct ApfSmem {
    alignas(16) float xc[APF_TJ], yc[APF_TJ], zc[APF_TJ], w[APF_TJ];  // -2 x centred position, |centred|^2
    float4 tp[APF_TJ], tv[APF_TJ];                                    // the records themselves
};

struct ApfBoid {
    Self self;
    float2 ax2, ay2, az2;  // p - centre, both halves
    float thr;             // cut + margin - |p - centre|^2
    float ax, ay, az;      // partial acceleration
};

// FOV = false: the configuration never culls (max_sight_angle >= pi, the C2 workload): the cosine is
// not evaluated at all.  Branch-free but for the rare exact path: in a dense flock nearly every batch
// reaches this code, and a branch per pair would only add its own overhead to the same issue slots.
template <bool FOV>
__device__ __forceinline__ void apf_pair(const DevParams &P, ApfBoid &b, const float4 pj, const float4 *tvj,
                                         bool live) {
    // (the fast grid walk's pair, fp_walk_nl.cu: fast_gate + fast_force)
    const Self &self = b.self;
    const float dx = fsub(pj.x, self.p.x), dy = fsub(pj.y, self.p.y), dz = fsub(pj.z, self.p.z);
    const float m2 = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
    const float r = rsqrt_seed(m2);
    const bool in = live && !(m2 >= P.m2_cut);
    bool clear = m2 >= 1e-12f, vis = true;
    if (FOV) {
        const float q = fmaf(self.vhat.z, dz, fmaf(self.vhat.y, dy, self.vhat.x * dx));
        const float c = q * r;
        const float gc = (c - P.fz_a) * (c - P.fz_b);
        clear = clear && fabsf(gc) > P.fz_gc_tol;
        vis = gc > 0.0f;
    }
    const bool pass = in && clear && vis;
    const float4 vj = *tvj;
    const float g0 = m2 * r, h = 0.5f * r;
    const float mag = fmaf(fmaf(-g0, g0, m2), h, g0);
    const float coef = fmaf(P.f_c, mag, P.neg_f_a * (r * r)) * r;
    const float w = m2 <= P.m2_one ? 1.0f : (mag - P.thr) * P.fz_rinv_fall;
    const float dvx = vj.x - self.v.x, dvy = vj.y - self.v.y, dvz = vj.z - self.v.z;
    const bool vsm = fmaxf(fmaxf(fabsf(dvx), fabsf(dvy)), fabsf(dvz)) <= FP_F32_EPSILON;
    const float cw = pass ? coef * w : 0.0f, fw = (pass && !vsm) ? P.f_v * w : 0.0f;  // (selected: see fp_walk_nl.cu)
    b.ax = fmaf(cw, dx, fmaf(fw, dvx, b.ax));
    b.ay = fmaf(cw, dy, fmaf(fw, dvy, b.ay));
    b.az = fmaf(cw, dz, fmaf(fw, dvz, b.az));
    if (in && !clear) {
        // guard band / coincident / NaN (rare; the boid's own record lands here once): exact sequence.
        // The reference skips records equal to the boid (flocking.rs:137-139); their exact
        // contribution is +0 for finite states, so evaluating them changes nothing.
        V3 contrib;
        if (pair_inrange<false>(P, self, v3(dx, dy, dz), m2, v3(vj.x, vj.y, vj.z), 1.0f, P.cstar, contrib)) {
            b.ax += contrib.x;
            b.ay += contrib.y;
            b.az += contrib.z;
        }
    }
}

template <int TAP, bool FOV>
__global__ void __launch_bounds__(AP_BLOCK, 4)
allpairs_fast_kernel(const DevParams P, const float4 *__restrict__ pos_all, const float4 *__restrict__ vel_all,
                     uint32_t n_all, uint32_t row0, uint32_t nrows, int js_log2, const float *__restrict__ bounds8,
                     float4 *__restrict__ pos_out, float4 *__restrict__ vel_out, unsigned *__restrict__ status,
                     TapOut tap) {
    __shared__ ApfSmem S;
    const uint32_t tid = threadIdx.x;
    const uint32_t JS = 1u << js_log2, slice = tid & (JS - 1u), rowslot = tid >> js_log2;
    const uint32_t rows_per_cta = 2u * (AP_BLOCK >> js_log2);
    // centre of the flock and the cancellation margin of the pre-gate (bounds_kernel: finite positions)
    float cx = 0.0f, cy = 0.0f, cz = 0.0f, margin = INFINITY;
    {
        const float lx = __ldg(bounds8 + 0), ly = __ldg(bounds8 + 1), lz = __ldg(bounds8 + 2);
        const float hx = __ldg(bounds8 + 3), hy = __ldg(bounds8 + 4), hz = __ldg(bounds8 + 5);
        if (lx <= hx && ly <= hy && lz <= hz) {
            cx = 0.5f * (lx + hx); cy = 0.5f * (ly + hy); cz = 0.5f * (lz + hz);
            const float ex = fmaxf(hx - cx, cx - lx), ey = fmaxf(hy - cy, cy - ly), ez = fmaxf(hz - cz, cz - lz);
            const float M2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex)) * 1.0001f + 1e-30f;
            margin = 2.5e-6f * M2;  // twice the bound on |t - (|p_j|^2 - 2 p_i . p_j)| + the threshold's rounding
        }
        if (!(cx == cx && cy == cy && cz == cz) || fabsf(cx) == INFINITY || fabsf(cy) == INFINITY || fabsf(cz) == INFINITY) {
            cx = cy = cz = 0.0f;
            margin = INFINITY;  // (everything survives the pre-gate: still correct)
        }
    }
    ApfBoid b[2];
    float4 pi4[2], vi4[2];
    bool act[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t r = blockIdx.x * rows_per_cta + 2u * rowslot + k;
        act[k] = r < nrows;
        const uint32_t i = row0 + (act[k] ? r : 0u);
        pi4[k] = __ldg(pos_all + i);
        vi4[k] = __ldg(vel_all + i);
        b[k].self = make_self(v3(pi4[k].x, pi4[k].y, pi4[k].z), v3(vi4[k].x, vi4[k].y, vi4[k].z));
        const float px = pi4[k].x - cx, py = pi4[k].y - cy, pz = pi4[k].z - cz;
        b[k].ax2 = make_float2(px, px);
        b[k].ay2 = make_float2(py, py);
        b[k].az2 = make_float2(pz, pz);
        b[k].thr = (P.m2_cut + margin) - fmaf(pz, pz, fmaf(py, py, px * px));
        b[k].ax = b[k].ay = b[k].az = 0.0f;
    }
    const bool need_pairs = (TAP != TAP_STEP) || !P.steering_overrides;
    bool dense = false;
    if (need_pairs) {
        for (uint32_t j0 = 0; j0 < n_all; j0 += APF_TJ) {
            __syncthreads();  // everyone is done with the previous tile
#pragma unroll
            for (int h = 0; h < APF_TJ / AP_BLOCK; ++h) {
                const uint32_t t = tid + h * AP_BLOCK, jl = j0 + t;
                float4 pj = make_float4(1e18f, 1e18f, 1e18f, 0.0f), vj = make_float4(0, 0, 0, 0);  // padding: far, finite
                if (jl < n_all) {
                    pj = __ldg(pos_all + jl);
                    vj = __ldg(vel_all + jl);
                }
                const float x = pj.x - cx, y = pj.y - cy, z = pj.z - cz;
                S.xc[t] = -2.0f * x; S.yc[t] = -2.0f * y; S.zc[t] = -2.0f * z;  // (the -2 of -2 p_i . p_j, exact)
                S.w[t] = fmaf(z, z, fmaf(y, y, x * x));
                S.tp[t] = pj;
                S.tv[t] = vj;
            }
            __syncthreads();
            const uint32_t cnt = min((uint32_t)APF_TJ, n_all - j0);
            uint32_t tried = 0, hits = 0;
            // one compare per boid and batch of four candidates.  (fminf drops a NaN operand: a record
            // with a non-finite position can go unseen here -- FAST numerics are defined on finite states.)
            auto pregate = [&](uint32_t q) -> bool {
                const float4 X = *reinterpret_cast<const float4 *>(&S.xc[4 * q]);
                const float4 Y = *reinterpret_cast<const float4 *>(&S.yc[4 * q]);
                const float4 Z = *reinterpret_cast<const float4 *>(&S.zc[4 * q]);
                const float4 W = *reinterpret_cast<const float4 *>(&S.w[4 * q]);
                const float2 X01 = make_float2(X.x, X.y), X23 = make_float2(X.z, X.w);
                const float2 Y01 = make_float2(Y.x, Y.y), Y23 = make_float2(Y.z, Y.w);
                const float2 Z01 = make_float2(Z.x, Z.y), Z23 = make_float2(Z.z, Z.w);
                const float2 W01 = make_float2(W.x, W.y), W23 = make_float2(W.z, W.w);
                bool hit = false;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float2 t01 = __ffma2_rn(b[k].ax2, X01, __ffma2_rn(b[k].ay2, Y01, __ffma2_rn(b[k].az2, Z01, W01)));
                    const float2 t23 = __ffma2_rn(b[k].ax2, X23, __ffma2_rn(b[k].ay2, Y23, __ffma2_rn(b[k].az2, Z23, W23)));
                    hit |= !(fminf(fminf(t01.x, t01.y), fminf(t23.x, t23.y)) >= b[k].thr);
                }
                return hit;
            };
            auto pairs = [&](uint32_t q) {  // someone may be in range: the whole batch takes the exact distance test
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t t = 4 * q + u;
                    const float4 pj = S.tp[t];
                    apf_pair<FOV>(P, b[0], pj, &S.tv[t], t < cnt);
                    apf_pair<FOV>(P, b[1], pj, &S.tv[t], t < cnt);
                }
            };
            if (dense) {
                for (uint32_t q = slice; 4u * q < cnt; q += JS) pairs(q);
            } else {
                // two batches per trip: their pre-gates are independent chains
                for (uint32_t q = slice; 4u * q < cnt; q += 2 * JS) {
                    const uint32_t q2 = q + JS;
                    const bool two = 4u * q2 < cnt;
                    const bool h1 = pregate(q), h2 = two && pregate(q2);
                    tried += two ? 2u : 1u;
                    hits += (h1 ? 1u : 0u) + (h2 ? 1u : 0u);
                    if (h1) pairs(q);
                    if (h2) pairs(q2);
                }
            }
            // a dense flock (C2: half of all pairs are in range) gains nothing from the pre-gate: once
            // most batches of a tile hit, the warp stops asking (decided once, warp-uniform)
            if (!dense && j0 >= 4u * APF_TJ)
                dense = __reduce_add_sync(0xffffffffu, hits) * 4u > __reduce_add_sync(0xffffffffu, tried) * 3u;
        }
    }
    // the JS partial sums of each boid -> its slice-0 lane
    for (uint32_t off = JS >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            b[k].ax += __shfl_xor_sync(0xffffffffu, b[k].ax, off);
            b[k].ay += __shfl_xor_sync(0xffffffffu, b[k].ay, off);
            b[k].az += __shfl_xor_sync(0xffffffffu, b[k].az, off);
        }
    }
    if (slice != 0) return;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (!act[k]) continue;
        const uint32_t r = blockIdx.x * rows_per_cta + 2u * rowslot + k;
        const uint32_t idx = __float_as_uint(pi4[k].w);  // caller index
        const V3 acc = v3(b[k].ax, b[k].ay, b[k].az);
        Extras e;
        unsigned flags = 0;
        const V3 a = accel_total_fast(P, b[k].self, acc, e, flags, TAP == TAP_ACCEL);
        if (TAP == TAP_ACCEL) {
            float *o = tap.accel3 + 3ull * idx;
            o[0] = a.x; o[1] = a.y; o[2] = a.z;
            if (tap.comp15) {
                float *c = tap.comp15 + 15ull * idx;
                c[0] = acc.x; c[1] = acc.y; c[2] = acc.z;
                c[3] = e.lead.x; c[4] = e.lead.y; c[5] = e.lead.z;
                c[6] = e.attr.x; c[7] = e.attr.y; c[8] = e.attr.z;
                c[9] = e.bbox.x; c[10] = e.bbox.y; c[11] = e.bbox.z;
                c[12] = e.steer.x; c[13] = e.steer.y; c[14] = e.steer.z;
            }
        } else {
            V3 np, nv;
            euler(P, b[k].self.p, b[k].self.v, a, np, nv);
            pos_out[r] = make_float4(np.x, np.y, np.z, pi4[k].w);
            vel_out[r] = make_float4(nv.x, nv.y, nv.z, 0.0f);
        }
        if (flags) atomicOr(status, flags);
    }
}

// rows a CTA of the fast kernel covers, and the j-split that fills the machine
static int apf_js_log2(uint32_t nrows) {
    int js = 0;  // 256 rows per CTA when the flock is large
    while (js < 5 && (uint64_t)nrows * (1u << js) < 256ull * 148 * 8) ++js;
    return js;
}

template <int TAP>
static int launch_ap2(cudaStream_t st, const dim3 grid, const DevParams &P, const float4 *pos_all,
                      const float4 *vel_all, uint32_t n_all, uint32_t row0, uint32_t nrows, float4 *pos_out,
                      float4 *vel_out, unsigned *status, const TapOut &tap_out) {
    auto kern = allpairs2_kernel<TAP>;
    const int smem = (int)sizeof(Ap2Smem);
    FP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, AP_BLOCK, smem, st>>>(P, pos_all, vel_all, n_all, row0, nrows, pos_out, vel_out, status, tap_out);
    return FP_OK;
}

int launch_allpairs(cudaStream_t st, const DevParams &P, int tap, const float4 *pos_all,
                    const float4 *vel_all, uint32_t n_all, uint32_t row0, uint32_t nrows,
                    float4 *pos_out, float4 *vel_out, unsigned *status, const TapOut &tap_out, int variant,
                    const float *bounds8) {
    if (nrows == 0) return FP_OK;
    if (P.numerics_fast && bounds8 && (tap == TAP_STEP || tap == TAP_ACCEL)) {
        const int js = apf_js_log2(nrows);
        const uint32_t rows_per_cta = 2u * (AP_BLOCK >> js);
        const dim3 gridf((nrows + rows_per_cta - 1) / rows_per_cta);
        const bool fov = P.fz_a > -2.5f;  // (-3: the configuration never culls)
#define FP_APF(T, F)                                                                                        \
    allpairs_fast_kernel<T, F><<<gridf, AP_BLOCK, 0, st>>>(P, pos_all, vel_all, n_all, row0, nrows, js, bounds8, \
                                                           pos_out, vel_out, status, tap_out)
        if (tap == TAP_STEP) {
            if (fov) FP_APF(TAP_STEP, true); else FP_APF(TAP_STEP, false);
        } else {
            if (fov) FP_APF(TAP_ACCEL, true); else FP_APF(TAP_ACCEL, false);
        }
#undef FP_APF
        count_launch();
        FP_CUDA(cudaGetLastError());
        return FP_OK;
    }
    const dim3 grid((nrows + AP_BLOCK - 1) / AP_BLOCK), block(AP_BLOCK);
    // variant 0: staged (pre-gate + lists; wins when most pairs are out of range), 1: one-phase
    // (wins in dense flocks where most candidates contribute).  Same bits either way; the
    // handle times both and keeps the faster.  FP_ALLPAIRS_VARIANT=0|1 pins one (debug).
    static const int pinned = [] {
        const char *e = getenv("FP_ALLPAIRS_VARIANT");
        return e && *e ? atoi(e) : -1;
    }();
    const bool staged = (pinned >= 0 ? pinned : variant) == 0;
    switch (tap) {
        case TAP_STEP:
            if (staged) {
                int rc = launch_ap2<TAP_STEP>(st, grid, P, pos_all, vel_all, n_all, row0, nrows, pos_out, vel_out,
                                              status, tap_out);
                if (rc) return rc;
                break;
            }
            allpairs_kernel<TAP_STEP><<<grid, block, 0, st>>>(P, pos_all, vel_all, n_all, row0, nrows,
                                                              pos_out, vel_out, status, tap_out);
            break;
        case TAP_ACCEL:
            if (staged) {
                int rc = launch_ap2<TAP_ACCEL>(st, grid, P, pos_all, vel_all, n_all, row0, nrows, pos_out, vel_out,
                                               status, tap_out);
                if (rc) return rc;
                break;
            }
            allpairs_kernel<TAP_ACCEL><<<grid, block, 0, st>>>(P, pos_all, vel_all, n_all, row0, nrows,
                                                               pos_out, vel_out, status, tap_out);
            break;
        case TAP_NEIGHBORS:
            allpairs_kernel<TAP_NEIGHBORS><<<grid, block, 0, st>>>(P, pos_all, vel_all, n_all, row0,
                                                                   nrows, pos_out, vel_out, status,
                                                                   tap_out);
            break;
        default:
            allpairs_kernel<TAP_CENSUS><<<grid, block, 0, st>>>(P, pos_all, vel_all, n_all, row0, nrows,
                                                                pos_out, vel_out, status, tap_out);
            break;
    }
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

}  // namespace fp

// fp_misc.cu -- layout conversion, instance export, State integrators, scan.
#include "fp_internal.h"

namespace fp {

constexpr int MB = 256;
static inline unsigned blocks_for(size_t n, int per = MB) { return (unsigned)((n + per - 1) / per); }

// ---- AoS-6 (caller layout, flocking.rs FlockingBoid position+velocity) <-> SoA float4 ----
__global__ void aos6_to_soa_kernel(const float *__restrict__ aos6, float4 *__restrict__ pos,
                                   float4 *__restrict__ vel, uint32_t n, uint32_t first_index) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *s = aos6 + 6ull * i;
    pos[i] = make_float4(s[0], s[1], s[2], __uint_as_float(first_index + i));
    vel[i] = make_float4(s[3], s[4], s[5], 0.0f);
}

__global__ void soa_to_aos6_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                   float *__restrict__ aos6, uint32_t n, uint32_t first_index,
                                   int by_index) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i], v = vel[i];
    const uint32_t dst = by_index ? (__float_as_uint(p.w) - first_index) : i;
    float *s = aos6 + 6ull * dst;
    s[0] = p.x; s[1] = p.y; s[2] = p.z;
    s[3] = v.x; s[4] = v.y; s[5] = v.z;
}

__global__ void unpermute_kernel(const float4 *__restrict__ pos_in, const float4 *__restrict__ vel_in,
                                 float4 *__restrict__ pos_out, float4 *__restrict__ vel_out,
                                 uint32_t n, uint32_t first_index) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos_in[i];
    const uint32_t dst = __float_as_uint(p.w) - first_index;
    pos_out[dst] = p;
    vel_out[dst] = vel_in[i];
}

int launch_aos6_to_soa(cudaStream_t st, const float *aos6, float4 *pos, float4 *vel, uint32_t n,
                       uint32_t first_index) {
    if (!n) return FP_OK;
    aos6_to_soa_kernel<<<blocks_for(n), MB, 0, st>>>(aos6, pos, vel, n, first_index);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}
int launch_soa_to_aos6(cudaStream_t st, const float4 *pos, const float4 *vel, float *aos6, uint32_t n,
                       uint32_t first_index, int by_index) {
    if (!n) return FP_OK;
    soa_to_aos6_kernel<<<blocks_for(n), MB, 0, st>>>(pos, vel, aos6, n, first_index, by_index);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}
int launch_unpermute(cudaStream_t st, const float4 *pos_in, const float4 *vel_in, float4 *pos_out,
                     float4 *vel_out, uint32_t n, uint32_t first_index) {
    if (!n) return FP_OK;
    unpermute_kernel<<<blocks_for(n), MB, 0, st>>>(pos_in, vel_in, pos_out, vel_out, n, first_index);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// ---- get_boid_instances (flocking.rs:230-245) --------------------------------
// approx 0.4 ulps_eq! defaults (epsilon = f32::EPSILON, max_ulps = 4)
__device__ __forceinline__ bool ulps_eq(float a, float b) {
    const float d = (a > b) ? fsub(a, b) : fsub(b, a);
    if (d <= FP_F32_EPSILON) return true;
    if (isnan(a) || isnan(b)) return false;
    if ((__float_as_uint(a) >> 31) != (__float_as_uint(b) >> 31)) return false;
    long long ia = (int)__float_as_uint(a), ib = (int)__float_as_uint(b);
    long long dd = ia - ib;
    if (dd < 0) dd = -dd;
    return dd <= 4;
}

// cgmath 0.18 Quaternion::from_arc(unit_z, normalize(v), None)
__device__ __forceinline__ float4 quat_from_arc_z(V3 vel) {
    const V3 dst = vnormalize(vel);
    // src = (0,0,1): |src|^2 = (0*0 + 0*0) + 1*1 = 1; dot = (0*dx + 0*dy) + 1*dz
    const float mag_avg = fsqrt(fmul(1.0f, vdot(dst, dst)));
    const float dot = fadd(fadd(fmul(0.0f, dst.x), fmul(0.0f, dst.y)), fmul(1.0f, dst.z));
    if (ulps_eq(dot, mag_avg)) return make_float4(1.0f, 0.0f, 0.0f, 0.0f);
    if (ulps_eq(dot, -mag_avg)) {
        // unit_x.cross(unit_z) = (0*1-0*0, 0*0-1*1, 1*0-0*0) = (0,-1,0): not ~0, normalised (0,-1,0);
        // from_axis_angle(axis, pi): (cos(pi/2), axis*sin(pi/2)) with f32 pi/2
        const float cs = -4.37113883e-08f, sn = 1.0f;  // cosf / sinf of 1.57079637f
        return make_float4(cs, fmul(0.0f, sn), fmul(-1.0f, sn), fmul(0.0f, sn));
    }
    // Quaternion::from_sv(mag_avg + dot, src x dst).normalize()
    const float s = fadd(mag_avg, dot);
    // (0,0,1) x d = (0*dz - 1*dy, 1*dx - 0*dz, 0*dy - 0*dx)
    const V3 v = v3(fsub(fmul(0.0f, dst.z), fmul(1.0f, dst.y)), fsub(fmul(1.0f, dst.x), fmul(0.0f, dst.z)),
                    fsub(fmul(0.0f, dst.y), fmul(0.0f, dst.x)));
    const float inv = fdiv(1.0f, fsqrt(fadd(fmul(s, s), vdot(v, v))));
    return make_float4(fmul(s, inv), fmul(v.x, inv), fmul(v.y, inv), fmul(v.z, inv));
}

__global__ void instances_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                 float *__restrict__ out, uint32_t n, uint32_t first_index, int raw) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i], v = vel[i];
    const uint32_t dst = __float_as_uint(p.w) - first_index;
    const float4 q = quat_from_arc_z(v3(v.x, v.y, v.z));  // (s, x, y, z)
    const float scale = 0.1f;
    if (!raw) {
        float *o = out + 8ull * dst;
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
        o[3] = q.x; o[4] = q.y; o[5] = q.z; o[6] = q.w;
        o[7] = scale;
        return;
    }
    // Instance::to_raw (instance.rs:14-22): model = T * R * S, normal = Matrix3::from(q).
    // cgmath Matrix3::from(Quaternion): x2 = x + x, ... (column-major)
    const float qs = q.x, qx = q.y, qy = q.z, qz = q.w;
    const float x2 = fadd(qx, qx), y2 = fadd(qy, qy), z2 = fadd(qz, qz);
    const float xx2 = fmul(x2, qx), xy2 = fmul(x2, qy), xz2 = fmul(x2, qz);
    const float yy2 = fmul(y2, qy), yz2 = fmul(y2, qz), zz2 = fmul(z2, qz);
    const float sy2 = fmul(y2, qs), sz2 = fmul(z2, qs), sx2 = fmul(x2, qs);
    float r[9];
    r[0] = fsub(fsub(1.0f, yy2), zz2); r[1] = fadd(xy2, sz2); r[2] = fsub(xz2, sy2);
    r[3] = fsub(xy2, sz2); r[4] = fsub(fsub(1.0f, xx2), zz2); r[5] = fadd(yz2, sx2);
    r[6] = fadd(xz2, sy2); r[7] = fsub(yz2, sx2); r[8] = fsub(fsub(1.0f, xx2), yy2);
    float *o = out + 25ull * dst;
    // T*R leaves R's columns (x + 0 terms), then *S scales columns 0..2; column 3 = (p, 1)
    for (int c = 0; c < 3; ++c) {
        o[4 * c + 0] = fmul(r[3 * c + 0], scale);
        o[4 * c + 1] = fmul(r[3 * c + 1], scale);
        o[4 * c + 2] = fmul(r[3 * c + 2], scale);
        o[4 * c + 3] = 0.0f;
    }
    o[12] = p.x; o[13] = p.y; o[14] = p.z; o[15] = 1.0f;
    for (int k = 0; k < 9; ++k) o[16 + k] = r[k];
}

int launch_instances(cudaStream_t st, const float4 *pos, const float4 *vel, float *out, uint32_t n,
                     uint32_t first_index, int raw) {
    if (!n) return FP_OK;
    instances_kernel<<<blocks_for(n), MB, 0, st>>>(pos, vel, out, n, first_index, raw);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// ---- State<T>::euler_step / rk4_step (state.rs:75-106, utils.rs:5-21) --------
__global__ void state_euler_kernel(const float *__restrict__ s, const float *__restrict__ ds, float h,
                                   float *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fadd(s[i], fmul(ds[i], h));  // state.rs:79, :81
}

__global__ void state_rk4_kernel(const float *__restrict__ s, const float *__restrict__ k1,
                                 const float *__restrict__ k2, const float *__restrict__ k3,
                                 const float *__restrict__ k4, float h, float *__restrict__ out,
                                 size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float h6 = fdiv(h, 6.0f), h3 = fdiv(h, 3.0f);
    // timestep/6*k1 + timestep/3*k2 + timestep/3*k3 + timestep/6*k4, left-assoc (state.rs:97-104)
    const float delta = fadd(fadd(fadd(fmul(h6, k1[i]), fmul(h3, k2[i])), fmul(h3, k3[i])), fmul(h6, k4[i]));
    out[i] = fadd(s[i], delta);
}

int launch_state_combine_euler(cudaStream_t st, const float *s, const float *ds, float h, float *out,
                               size_t n) {
    if (!n) return FP_OK;
    state_euler_kernel<<<blocks_for(n), MB, 0, st>>>(s, ds, h, out, n);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}
int launch_state_combine_rk4(cudaStream_t st, const float *s, const float *k1, const float *k2,
                             const float *k3, const float *k4, float h, float *out, size_t n) {
    if (!n) return FP_OK;
    state_rk4_kernel<<<blocks_for(n), MB, 0, st>>>(s, k1, k2, k3, k4, h, out, n);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// The flock as State<boid>: 6 elements [p, v], derivative [v, a] with the
// acceleration accumulated beforehand and frozen (springy_mesh.rs:223-257).
// Euler: p + v*h, v + a*h.  RK4 with frozen a:
//   k1 = (v, a); k2 = (v + a*(h*0.5), a); k3 = k2; k4 = (v + a*h, a)
__global__ void flock_state_step_kernel(float4 *__restrict__ pos, float4 *__restrict__ vel,
                                        const float *__restrict__ accel3, uint32_t n,
                                        uint32_t first_index, float h, int rk4) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i], v = vel[i];
    const float *a = accel3 + 3ull * (__float_as_uint(p.w) - first_index);
    const float pv[3] = {p.x, p.y, p.z}, vv[3] = {v.x, v.y, v.z};
    float np[3], nv[3];
    if (!rk4) {
        for (int c = 0; c < 3; ++c) {
            np[c] = fadd(pv[c], fmul(vv[c], h));
            nv[c] = fadd(vv[c], fmul(a[c], h));
        }
    } else {
        const float hh = fmul(h, 0.5f), h6 = fdiv(h, 6.0f), h3 = fdiv(h, 3.0f);
        for (int c = 0; c < 3; ++c) {
            const float k1p = vv[c];
            const float k2p = fadd(vv[c], fmul(a[c], hh));  // velocity slot of S + k1*(h/2)
            const float k3p = k2p;                          // k2's velocity slot is again a
            const float k4p = fadd(vv[c], fmul(a[c], h));
            const float dp = fadd(fadd(fadd(fmul(h6, k1p), fmul(h3, k2p)), fmul(h3, k3p)), fmul(h6, k4p));
            const float dv = fadd(fadd(fadd(fmul(h6, a[c]), fmul(h3, a[c])), fmul(h3, a[c])), fmul(h6, a[c]));
            np[c] = fadd(pv[c], dp);
            nv[c] = fadd(vv[c], dv);
        }
    }
    pos[i] = make_float4(np[0], np[1], np[2], p.w);
    vel[i] = make_float4(nv[0], nv[1], nv[2], 0.0f);
}

int launch_flock_state_step(cudaStream_t st, float4 *pos, float4 *vel, const float *accel3, uint32_t n,
                            uint32_t first_index, float h, int rk4) {
    if (!n) return FP_OK;
    flock_state_step_kernel<<<blocks_for(n), MB, 0, st>>>(pos, vel, accel3, n, first_index, h, rk4);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// ---- bounds of finite positions: out8 = min xyz, max xyz, max |v|^2, unused ---------------
__device__ __forceinline__ void atomic_min_f(float *a, float v) {
    // monotone int mapping works for mixed signs with two atomics
    if (v >= 0.0f) atomicMin((int *)a, __float_as_int(v));
    else atomicMax((unsigned *)a, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float *a, float v) {
    if (v >= 0.0f) atomicMax((int *)a, __float_as_int(v));
    else atomicMin((unsigned *)a, __float_as_uint(v));
}
__global__ void bounds_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n,
                              float *out8) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    float v2 = 0.0f;  // max |v|^2 over finite velocities (sizes the re-binning skin)
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pos[i];
        const float c[3] = {p.x, p.y, p.z};
        for (int a = 0; a < 3; ++a)
            if (isfinite(c[a])) {
                lo[a] = fminf(lo[a], c[a]);
                hi[a] = fmaxf(hi[a], c[a]);
            }
        if (vel) {
            const float4 v = vel[i];
            const float m = fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x));
            if (isfinite(m)) v2 = fmaxf(v2, m);
        }
    }
    for (int off = 16; off > 0; off >>= 1) v2 = fmaxf(v2, __shfl_down_sync(0xffffffffu, v2, off));
    if ((threadIdx.x & 31) == 0 && v2 > 0.0f) atomicMax((unsigned *)(out8 + 6), __float_as_uint(v2));
    for (int a = 0; a < 3; ++a) {
        for (int off = 16; off > 0; off >>= 1) {
            lo[a] = fminf(lo[a], __shfl_down_sync(0xffffffffu, lo[a], off));
            hi[a] = fmaxf(hi[a], __shfl_down_sync(0xffffffffu, hi[a], off));
        }
        if ((threadIdx.x & 31) == 0) {
            if (lo[a] != INFINITY) atomic_min_f(out8 + a, lo[a]);
            if (hi[a] != -INFINITY) atomic_max_f(out8 + 3 + a, hi[a]);
        }
    }
}
int launch_bounds(cudaStream_t st, const float4 *pos, const float4 *vel, uint32_t n, float *out8) {
    const float init[8] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, 0.0f, 0.0f};
    FP_CUDA(cudaMemcpyAsync(out8, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (!n) return FP_OK;
    bounds_kernel<<<min(blocks_for(n), 148u * 8u), MB, 0, st>>>(pos, vel, n, out8);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// ---- lazy re-binning control (one thread; see SkinCtl in fp_internal.h) -------------------
// Displacement bound of one Euler move p' = fl(p + fl(dt * v)): at most dt * |v| (1 + 2^-23)
// plus half an ulp of the result per component; ulp(p') <= 2^-23 * 2 |p| covers a binade crossing.
__device__ __host__ inline float skin_step_bound(float v2max, float pmax, float dt) {
    return sqrtf(v2max) * fabsf(dt) * 1.000001f + pmax * 2.1e-7f + 1e-30f;
}
__global__ void skin_gate_kernel(SkinCtl *c, uint32_t ordinal, int rebin, float dt, float budget, int mode,
                                 const Mail *mail, int world, unsigned *status) {
    if (c->stale) return;  // sticky until the host settles: nothing after the first void step counts
    const int lane = threadIdx.x;  // one warp
    uint32_t v2 = 0, pm = 0;
    if (!rebin) {
        if (mode == GATE_MAILBOX) {
            bool ok = true;
            if (lane < world) {
                const volatile Mail *m = mail + (ordinal & 1u) * FP_MAX_WORLD + lane;
                const long long t0 = clock64();
                while (m->tag != ordinal) {
                    if (clock64() - t0 > 100000000000ll) {  // ~1 min: a peer died or the ranks disagree
                        ok = false;
                        break;
                    }
                    __nanosleep(64);
                }
                __threadfence_system();
                v2 = m->v2max;
                pm = m->pmax;
            }
            if (!__all_sync(0xffffffffu, ok)) {
                if (lane == 0) {
                    c->stale = 1u;
                    c->first_stale = ordinal;
                    c->fault = 1u;
                    if (status) atomicOr(status, 32u);
                }
                return;
            }
            v2 = __reduce_max_sync(0xffffffffu, v2);
            pm = __reduce_max_sync(0xffffffffu, pm);
        } else if (mode == GATE_REDUCED) {
            v2 = c->g_v2max;
            pm = c->g_pmax;
        } else {
            v2 = c->v2max;
            pm = c->pmax;
        }
    }
    if (lane != 0) return;
    if (rebin) {
        c->D = 0.0f;  // positions are about to be binned where they are
    } else {
        c->g_v2max = v2;
        c->g_pmax = pm;
        const float D = c->D + skin_step_bound(__uint_as_float(v2), __uint_as_float(pm), dt);
        c->D = D;
        if (!(D <= budget)) {  // also catches NaN
            c->stale = 1u;
            c->first_stale = ordinal;
            return;
        }
    }
    c->v2max = 0u;
    c->pmax = 0u;
}
int launch_skin_gate(cudaStream_t st, SkinCtl *ctl, uint32_t ordinal, int rebin, float dt, float budget, int mode,
                     const Mail *mail, int world, unsigned *status) {
    skin_gate_kernel<<<1, 32, 0, st>>>(ctl, ordinal, rebin, dt, budget, mode, mail, world, status);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// one thread per destination rank: payload, system fence, then the tag
__global__ void mail_post_kernel(const SkinCtl *c, MailPeers peers, int world, int rank, uint32_t tag) {
    if (c->stale) return;
    const int q = threadIdx.x;
    if (q >= world) return;
    Mail *dst = peers.box[q] + (tag & 1u) * FP_MAX_WORLD + rank;
    dst->v2max = c->v2max;
    dst->pmax = c->pmax;
    __threadfence_system();
    *(volatile uint32_t *)&dst->tag = tag;
}
int launch_mail_post(cudaStream_t st, const SkinCtl *ctl, const MailPeers &peers, int world, int rank,
                     uint32_t tag) {
    mail_post_kernel<<<1, 32, 0, st>>>(ctl, peers, world, rank, tag);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// ---- self-test of the branch-free exact sqrt / division (fp_device.cuh) ------------------
__global__ void fastmath_check_kernel(uint64_t n, uint64_t seed, unsigned long long *out) {
    unsigned long long bad_sqrt = 0, bad_div = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long r0 = mix64(seed + 3 * i), r1 = mix64(seed + 3 * i + 1),
                                 r2 = mix64(seed + 3 * i + 2);
        // random mantissas, exponents drawn from the admitted ranges
        auto make = [](unsigned long long r, int elo, int ehi, bool neg) {
            const unsigned mant = (unsigned)(r & 0x7fffffu);
            const int e = elo + (int)((r >> 23) % (unsigned)(ehi - elo + 1));
            const unsigned sign = neg ? (unsigned)((r >> 60) & 1u) << 31 : 0u;
            return __uint_as_float(sign | ((unsigned)(e + 127) << 23) | mant);
        };
        const float x = make(r0, -60, 60, false);  // m2
        if (__float_as_uint(sqrt_rn_fast(x)) != __float_as_uint(__fsqrt_rn(x))) ++bad_sqrt;
        const float mag = __fsqrt_rn(x);           // in 2^+-30
        const float a = make(r1, -40, 40, true);   // neg_f_a-like numerators
        const float b = make(r2, -40, 40, false);  // fall-like divisors
        const float m = fmul(mag, mag);
        const float num = make(r1 >> 7, -53, 41, true);
        if (__float_as_uint(div_rn_fast(1.0f, mag)) != __float_as_uint(__fdiv_rn(1.0f, mag))) ++bad_div;
        if (__float_as_uint(div_rn_fast(a, m)) != __float_as_uint(__fdiv_rn(a, m))) ++bad_div;
        if (__float_as_uint(div_rn_fast(num, b)) != __float_as_uint(__fdiv_rn(num, b))) ++bad_div;
        // the everyday range: distances of 1e-3 .. 1e3
        const float xe = make(r2 >> 9, -20, 20, false);
        const float me = __fsqrt_rn(xe);
        if (__float_as_uint(sqrt_rn_fast(xe)) != __float_as_uint(me)) ++bad_sqrt;
        if (__float_as_uint(div_rn_fast(1.0f, me)) != __float_as_uint(__fdiv_rn(1.0f, me))) ++bad_div;
        if (__float_as_uint(div_rn_fast(-1.0f, fmul(me, me))) != __float_as_uint(__fdiv_rn(-1.0f, fmul(me, me))))
            ++bad_div;
    }
    if (bad_sqrt) atomicAdd(out, bad_sqrt);
    if (bad_div) atomicAdd(out + 1, bad_div);
}

int launch_fastmath_check(uint64_t n, uint64_t seed, uint64_t out_mismatch[2]) {
    unsigned long long *d = nullptr;
    FP_CUDA(cudaMalloc((void **)&d, 2 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemset(d, 0, 2 * sizeof(unsigned long long));
    if (e == cudaSuccess) {
        fastmath_check_kernel<<<148 * 8, 256>>>(n, seed, d);
        count_launch();
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out_mismatch, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "fastmath check", __FILE__, __LINE__);
    return FP_OK;
}

// ---- exclusive scan (uint32, in place): reduce / scan partials / downsweep ----
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        for (int off = 1; off < SCAN_THREADS / 32; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const uint32_t base = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t *__restrict__ data,
                                                                   size_t n, uint32_t *partial) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const size_t i = base + (size_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += data[i];
    }
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_partials_kernel(uint32_t *partial, size_t m) {
    uint32_t carry = 0;
    for (size_t base = 0; base < m; base += SCAN_THREADS) {
        const size_t i = base + threadIdx.x;
        const uint32_t v = i < m ? partial[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        if (i < m) partial[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_down_kernel(uint32_t *data, size_t n,
                                                                 const uint32_t *__restrict__ partial) {
    // thread t owns SCAN_ITEMS consecutive elements (blocked) so the in-tile order is plain
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        s += v[k];
    }
    uint32_t run = block_exclusive_scan(s, nullptr) + partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

int launch_exclusive_scan(cudaStream_t st, uint32_t *data, size_t n, uint32_t *tmp) {
    if (!n) return FP_OK;
    const unsigned nb = (unsigned)((n + SCAN_TILE - 1) / SCAN_TILE);
    scan_reduce_kernel<<<nb, SCAN_THREADS, 0, st>>>(data, n, tmp);
    scan_partials_kernel<<<1, SCAN_THREADS, 0, st>>>(tmp, nb);
    scan_down_kernel<<<nb, SCAN_THREADS, 0, st>>>(data, n, tmp);
    count_launch(3);
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

}  // namespace fp

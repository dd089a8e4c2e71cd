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

int launch_allpairs(cudaStream_t st, const DevParams &P, int tap, const float4 *pos_all,
                    const float4 *vel_all, uint32_t n_all, uint32_t row0, uint32_t nrows,
                    float4 *pos_out, float4 *vel_out, unsigned *status, const TapOut &tap_out) {
    if (nrows == 0) return FP_OK;
    const dim3 grid((nrows + AP_BLOCK - 1) / AP_BLOCK), block(AP_BLOCK);
    switch (tap) {
        case TAP_STEP:
            allpairs_kernel<TAP_STEP><<<grid, block, 0, st>>>(P, pos_all, vel_all, n_all, row0, nrows,
                                                              pos_out, vel_out, status, tap_out);
            break;
        case TAP_ACCEL:
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

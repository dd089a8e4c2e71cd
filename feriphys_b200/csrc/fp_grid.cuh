// fp_grid.cuh -- device helpers shared by the uniform-grid kernels (fp_grid.cu, fp_walk.cu).
#pragma once

#include "fp_internal.h"

namespace fp {

__device__ __forceinline__ int cell_coord(float x, float origin, float inv_cell, int dim) {
    // monotone in x: fl(x - o) , fl(. * inv), floor, clamp are all monotone
    const int c = __float2int_rd(fmul(fsub(x, origin), inv_cell));  // NaN -> 0, saturating
    return min(max(c, 0), dim - 1);
}

// x coordinate: the GLOBAL cell layer (identical on every rank), shifted into this rank's
// slab-local frame (layer 0 = left halo) and clamped to it.  Single GPU: xoff = 0,
// gdimx = dim[0], i.e. plain cell_coord.
__device__ __forceinline__ int cell_coord_x(const GridDesc &g, float x) {
    const int c = cell_coord(x, g.origin[0], g.inv_cell, g.gdimx) - g.xoff;
    return min(max(c, 0), g.dim[0] - 1);
}

// Shared epilogue of the walk kernels: per-boid extras (flocking.rs:105-113), Euler
// update (flocking.rs:116-117) and the debug taps.
template <int TAP>
__device__ __forceinline__ void walk_finish(const DevParams &P, uint32_t s, float4 pi4, float4 vi4,
                                            const Self &self, V3 acc, uint32_t n_count,
                                            unsigned long long n_hash, float4 *__restrict__ pos_out,
                                            float4 *__restrict__ vel_out, unsigned *__restrict__ status,
                                            const TapOut &tap) {
    const bool ghost = __float_as_uint(vi4.w) != 0u;
    if (ghost) {
        if (TAP == TAP_STEP) {  // keep the slot well-defined; dropped by the next exchange
            pos_out[s] = pi4;
            vel_out[s] = vi4;
        }
        return;
    }
    const uint32_t idx = __float_as_uint(pi4.w);
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
    pos_out[s] = make_float4(np.x, np.y, np.z, pi4.w);
    vel_out[s] = make_float4(nv.x, nv.y, nv.z, 0.0f);
    if (flags) atomicOr(status, flags);
}

// FOV half of pair_inrange: true when the pair is culled (boid.rs:149)
__device__ __forceinline__ bool pair_fov_culled(const Self &s, V3 d, float m2, float cstar) {
    const float c = vdot(s.vhat, vscale(d, fdiv(1.0f, fsqrt(m2))));
    return c >= -1.0f && c <= cstar;
}

}  // namespace fp

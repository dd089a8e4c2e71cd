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

// Cell key: x slowest, z fastest.  A "row" of the walk is three consecutive z cells, and
// every x layer is one contiguous slot range of the sorted arrays (what the slab sharding
// exchanges as halos).
__device__ __forceinline__ uint32_t cell_key(const GridDesc &g, int cx, int cy, int cz) {
    return (uint32_t)((cx * g.dim[1] + cy) * g.dim[2] + cz);
}
__device__ __forceinline__ uint32_t cell_key_of(const GridDesc &g, float4 p) {
    return cell_key(g, cell_coord_x(g, p.x), cell_coord(p.y, g.origin[1], g.inv_cell, g.dim[1]),
                    cell_coord(p.z, g.origin[2], g.inv_cell_z, g.dim[2]));
}
// The cell a slot was BINNED under (its position may since have drifted by up to skin / 2).
__device__ __forceinline__ void home_cell(const GridDesc &g, uint32_t key, int &cx, int &cy, int &cz) {
    const uint32_t t = key / (uint32_t)g.dim[2];
    cz = (int)(key - t * (uint32_t)g.dim[2]);
    cx = (int)(t / (uint32_t)g.dim[1]);
    cy = (int)(t - (uint32_t)cx * (uint32_t)g.dim[1]);
}
// slot range of cells (x, y, z0 .. z1) -- one contiguous interval
__device__ __forceinline__ uint32_t row_base(const GridDesc &g, int x, int y) {
    return (uint32_t)((x * g.dim[1] + y) * g.dim[2]);
}

// Speed / extent tracking for the lazy re-binning: max |v|^2 and max |coordinate| over the
// INPUT records of a walk (the Euler move of this step is dt * v of the input velocity).
// Bit patterns of non-negative floats order like the floats; a NaN/inf poisons the bound,
// which only makes the flock re-bin.  Called by full warps.
__device__ __forceinline__ void track_motion(SkinCtl *ctl, bool counts, float4 p, float4 v) {
    uint32_t v2 = 0, pm = 0;
    if (counts) {
        v2 = __float_as_uint(fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x)));
        pm = __float_as_uint(fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)));
        if (v2 > 0x7f800000u) v2 = 0x7f800000u;  // NaN -> inf
        if (pm > 0x7f800000u) pm = 0x7f800000u;
    }
    v2 = __reduce_max_sync(0xffffffffu, v2);
    pm = __reduce_max_sync(0xffffffffu, pm);
    if ((threadIdx.x & 31) == 0) {
        // plain look first: after the first few warps almost nobody needs the atomic
        if (v2 > *(volatile uint32_t *)&ctl->v2max) atomicMax(&ctl->v2max, v2);
        if (pm > *(volatile uint32_t *)&ctl->pmax) atomicMax(&ctl->pmax, pm);
    }
}

// Shared epilogue of the walk kernels: per-boid extras (flocking.rs:105-113), Euler
// update (flocking.rs:116-117) and the debug taps.
template <int TAP, bool FAST = false>
__device__ __forceinline__ void walk_finish(const DevParams &P, uint32_t s, float4 pi4, float4 vi4,
                                            const Self &self, V3 acc, uint32_t n_count,
                                            unsigned long long n_hash, const WalkIO &io,
                                            unsigned *__restrict__ status, const TapOut &tap) {
    const bool ghost = __float_as_uint(vi4.w) != 0u;
    if (ghost) {
        if (TAP == TAP_STEP) {  // keep the slot well-defined; dropped by the next exchange
            io.pos_out[s] = pi4;
            io.vel_out[s] = vi4;
            io.soa_out[0][s] = pi4.x;
            io.soa_out[1][s] = pi4.y;
            io.soa_out[2][s] = pi4.z;
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
    const V3 a = FAST ? accel_total_fast(P, self, acc, e, flags, TAP == TAP_ACCEL)
                      : accel_total(P, self, acc, e, flags, TAP == TAP_ACCEL);
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
    io.pos_out[s] = make_float4(np.x, np.y, np.z, pi4.w);
    io.vel_out[s] = make_float4(nv.x, nv.y, nv.z, 0.0f);
    io.soa_out[0][s] = np.x;  // SoA copy: what the next walk stages through TMA
    io.soa_out[1][s] = np.y;
    io.soa_out[2][s] = np.z;
#pragma unroll
    for (int fc = 0; fc < 2; ++fc) {  // sharded: this boid is a ghost of the neighbour's next step
        const PeerFace &pf = io.push[fc];
        if (s >= pf.begin && s < pf.end) {
            const uint32_t k = s - pf.begin;
            pf.pos[k] = make_float4(np.x, np.y, np.z, pi4.w);
            pf.vel[k] = make_float4(nv.x, nv.y, nv.z, 0.0f);
            pf.sx[k] = np.x;
            pf.sy[k] = np.y;
            pf.sz[k] = np.z;
        }
    }
    if (flags) atomicOr(status, flags);
}

// FOV half of pair_inrange: true when the pair is culled (boid.rs:149)
__device__ __forceinline__ bool pair_fov_culled(const Self &s, V3 d, float m2, float cstar) {
    const float c = vdot(s.vhat, vscale(d, fdiv(1.0f, fsqrt(m2))));
    return c >= -1.0f && c <= cstar;
}

}  // namespace fp

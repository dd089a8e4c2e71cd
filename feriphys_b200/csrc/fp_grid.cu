// fp_grid.cu -- K2b/K3: uniform-grid variant of the influence pass.
//
// The reference is all-pairs only (flocking.rs:133-151).  Because a
// pair contributes exactly zero once dist >= thr + falloff (boid.rs:154-157),
// restricting each boid to the 27 cells around it is EXACT provided the cell
// edge exceeds that reach: the same f32 predicates decide the same pairs.
//
// A BINNING: cell keys + per-cell counts -> exclusive scan (cell_start) -> stable
// radix sort of (key, slot) -> gather into sorted float4 + SoA arrays.  A STEP:
// the walk over the 27 cells (9 rows of z slices) around each boid's HOME cell,
// fused with lead/attractor/bbox/steering terms and the Euler update.  One
// binning serves many steps (lazy re-binning, fp_api.cu: grid_steps / settle):
// the cell edge carries a skin, and a boid's home cell is the key its slot was
// binned under.  The state stays in sorted order (pos.w carries the caller
// index), so every global access is coalesced.  This file holds the binning
// kernels and the one-phase walk of the neighbour-set / census taps; steps run
// on standing candidate lists (fp_walk_nl.cu) or the staged walk (fp_walk.cu).
#include <stdlib.h>

#include "fp_grid.cuh"

namespace fp {

constexpr int GB = 256;

__global__ void __launch_bounds__(GB)
grid_keys_kernel(const GridDesc g, const float4 *__restrict__ pos, uint32_t n,
                 uint32_t *__restrict__ keys, uint32_t *__restrict__ cell_count) {
    const uint32_t i = blockIdx.x * GB + threadIdx.x;
    const bool valid = i < n;
    uint32_t key = 0xffffffffu;
    if (valid) {
        key = cell_key_of(g, pos[i]);
        keys[i] = key;
    }
    // neighbouring slots usually share a cell: one atomic per distinct key per warp
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (valid && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(cell_count + key, __popc(peers));
}

int launch_grid_keys(cudaStream_t st, const GridDesc &g, const float4 *pos, uint32_t n, GridWork &w) {
    FP_CUDA(cudaMemsetAsync(w.cell_start, 0, ((size_t)g.ncells + 1) * sizeof(uint32_t), st));
    if (n) {
        grid_keys_kernel<<<(n + GB - 1) / GB, GB, 0, st>>>(g, pos, n, w.keys[0], w.cell_start);
        count_launch();
        FP_CUDA(cudaGetLastError());
    }
    return launch_exclusive_scan(st, w.cell_start, (size_t)g.ncells + 1, w.scan_tmp);
}

__global__ void __launch_bounds__(GB)
grid_reorder_kernel(const uint32_t *__restrict__ vals, const float4 *__restrict__ pos_in,
                    const float4 *__restrict__ vel_in, float4 *__restrict__ pos_out,
                    float4 *__restrict__ vel_out, float *__restrict__ sx, float *__restrict__ sy,
                    float *__restrict__ sz, uint32_t n, const SkinCtl *__restrict__ ctl) {
    const uint32_t i = blockIdx.x * GB + threadIdx.x;
    if (i >= n) return;
    if (ctl && ctl->stale) return;  // the steps before this re-binning did not happen: keep the state
    const uint32_t src = vals[i];
    const float4 p = pos_in[src];
    pos_out[i] = p;
    vel_out[i] = vel_in[src];
    sx[i] = p.x;  // SoA copy of the positions: what the walk stages through TMA
    sy[i] = p.y;
    sz[i] = p.z;
}

int launch_grid_reorder(cudaStream_t st, const uint32_t *vals, const float4 *pos_in,
                        const float4 *vel_in, float4 *pos_out, float4 *vel_out, float *const *soa,
                        uint32_t n, const SkinCtl *ctl) {
    if (!n) return FP_OK;
    grid_reorder_kernel<<<(n + GB - 1) / GB, GB, 0, st>>>(vals, pos_in, vel_in, pos_out, vel_out, soa[0],
                                                          soa[1], soa[2], n, ctl);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

// K3: one thread per boid of the sorted state.  For each of the 9 (dx, dy) rows
// the cells cz-1 .. cz+1 are one contiguous slot range of the sorted arrays.
// Accumulation order: rows (dx, dy) ascending, then slot ascending -- fixed.
// The home cell comes from the key the slot was binned under, not from the current
// position (lazy re-binning: the position may have drifted by up to skin / 2).
constexpr int WALK_BLOCK = 128;

template <int TAP>
__global__ void __launch_bounds__(WALK_BLOCK)
grid_walk_kernel(const DevParams P, const GridDesc g, const WalkIO io, unsigned *__restrict__ status,
                 TapOut tap) {
    if (TAP == TAP_STEP && io.ctl && io.ctl->stale) return;
    const float4 *__restrict__ pos_s = io.pos_s;
    const float4 *__restrict__ vel_s = io.vel_s;
    const uint32_t *__restrict__ cell_start = io.cell_start;
    const uint32_t s = io.first + blockIdx.x * WALK_BLOCK + threadIdx.x;
    const bool active = s < io.last;
    unsigned long long c_far = 0, c_cull = 0, c_in = 0;
    V3 acc = v3zero();
    uint32_t n_count = 0;
    unsigned long long n_hash = 0;
    float4 pi4 = make_float4(0, 0, 0, 0), vi4 = make_float4(0, 0, 0, 0);
    Self self;
    if (active) {
        pi4 = pos_s[s];
        vi4 = vel_s[s];
    }
    if (TAP == TAP_STEP && io.ctl) track_motion(io.ctl, active, pi4, vi4);
    if (active) {
        self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
        const bool ghost = __float_as_uint(vi4.w) != 0u;  // halo copy owned by another rank
        const bool need_pairs = !ghost && ((TAP != TAP_STEP) || !P.steering_overrides);
        if (need_pairs) {
            int cx, cy, cz;
            home_cell(g, __ldg(io.home + (s - io.first)), cx, cy, cz);
            const int z0 = max(cz - g.zspan, 0), z1 = min(cz + g.zspan, g.dim[2] - 1);
            for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x) {
                for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
                    const uint32_t rowbase = row_base(g, x, y);
                    const uint32_t jb = __ldg(cell_start + rowbase + z0);
                    const uint32_t je = __ldg(cell_start + rowbase + z1 + 1);
                    for (uint32_t j = jb; j < je; ++j) {
                        const float4 pj = __ldg(pos_s + j);
                        if (TAP == TAP_STEP || TAP == TAP_ACCEL) {
                            V3 d;
                            const float m2 = pair_m2(self, v3(pj.x, pj.y, pj.z), d);
                            if (m2 >= P.m2_cut) continue;
                            const float4 vj = __ldg(vel_s + j);
                            V3 contrib;
                            if (pair_flock(P, self, d, m2, v3(vj.x, vj.y, vj.z), contrib))
                                acc = vadd(acc, contrib);
                        } else {
                            const float4 vj = __ldg(vel_s + j);
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
            }
        }
    }

    if (TAP == TAP_CENSUS) {
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
    walk_finish<TAP>(P, s, pi4, vi4, self, acc, n_count, n_hash, io, status, tap);
}

int launch_grid_walk(cudaStream_t st, const DevParams &P, const GridDesc &g, int tap, const WalkIO &io,
                     unsigned *status, const TapOut &tap_out) {
    if (io.last <= io.first) return FP_OK;
    // steps and the acceleration tap: the staged walk (fp_walk.cu); neighbour sets and the census:
    // the one-phase kernel above, which evaluates every predicate for every candidate
    if (tap == TAP_STEP || tap == TAP_ACCEL) return launch_grid_walk3(st, P, g, tap, io, status, tap_out);
    const uint32_t rows = io.last - io.first;
    const dim3 grid((rows + WALK_BLOCK - 1) / WALK_BLOCK), block(WALK_BLOCK);
    if (tap == TAP_NEIGHBORS)
        grid_walk_kernel<TAP_NEIGHBORS><<<grid, block, 0, st>>>(P, g, io, status, tap_out);
    else
        grid_walk_kernel<TAP_CENSUS><<<grid, block, 0, st>>>(P, g, io, status, tap_out);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

}  // namespace fp

// fp_small.cu -- K4: demo-sized flocks (config C1: 110 boids, 1000 steps).
//
// At N ~ 100 a step is launch- and latency-bound, so one CTA keeps the whole
// flock in shared memory and advances `nsteps` steps in a single launch
// (lead-boid rows pre-tabulated by the host).  One thread owns one
// boid and walks j = 0..N-1 in ascending order -- the reference's summation
// order (flocking.rs:133-151), so results are bit-identical to the Rust loop.
// Latency is hidden inside the thread: four pairs per trip are evaluated with
// the branch-free exact pair function (independent dependency chains the
// scheduler interleaves) and then added to the accumulator in index order.
#include <stdlib.h>

#include "fp_internal.h"

namespace fp {

constexpr int SM_THREADS = 256;
constexpr uint32_t SM_MAX = SM_THREADS;
constexpr int SM_BATCH = 4;

uint32_t small_max_boids() { return SM_MAX; }

// A boid of the single-CTA kernels comes from the device arrays or, when fp_flock_write_state left
// the caller's rows in mapped host memory, straight from there (the same conversion as
// aos6_to_soa_kernel: caller order is internal order for these flocks); it goes back to the device
// arrays and, for fp_flock_read_state, to mapped host memory as a caller-order row.
__device__ __forceinline__ void small_load(const float4 *gpos, const float4 *gvel, const float *aos_in,
                                           uint32_t first_index, uint32_t i, float4 &p, float4 &v) {
    if (aos_in) {
        const float2 *r = reinterpret_cast<const float2 *>(aos_in + 6ull * i);  // (rows are 8-byte aligned)
        const float2 a = r[0], b = r[1], c = r[2];
        p = make_float4(a.x, a.y, b.x, __uint_as_float(first_index + i));
        v = make_float4(b.y, c.x, c.y, 0.0f);
    } else {
        p = gpos[i];
        v = gvel[i];
    }
}
__device__ __forceinline__ void small_store(float4 *gpos, float4 *gvel, float *aos_out, uint32_t i, float4 p,
                                            float4 v) {
    gpos[i] = p;
    gvel[i] = v;
    if (aos_out) {
        float2 *r = reinterpret_cast<float2 *>(aos_out + 6ull * i);
        r[0] = make_float2(p.x, p.y);
        r[1] = make_float2(p.z, v.x);
        r[2] = make_float2(v.y, v.z);
    }
}

__global__ void __launch_bounds__(SM_THREADS, 1)
small_kernel(const DevParams P, float4 *__restrict__ gpos, float4 *__restrict__ gvel, uint32_t n,
             uint32_t nsteps, const float *__restrict__ lead_table, uint32_t lead_rows,
             unsigned *__restrict__ status, const float *__restrict__ aos_in, float *__restrict__ aos_out,
             uint32_t first_index) {
    __shared__ float4 spos[2][SM_MAX + SM_BATCH];
    __shared__ float4 svel[2][SM_MAX + SM_BATCH];
    const uint32_t i = threadIdx.x;
    const bool mine = i < n;
    for (int b = 0; b < 2; ++b)  // padding read by the masked tail of the last batch
        if (i < SM_BATCH) spos[b][SM_MAX + i] = svel[b][SM_MAX + i] = make_float4(0, 0, 0, 0);
    if (mine) {
        small_load(gpos, gvel, aos_in, first_index, i, spos[0][i], svel[0][i]);
    } else {
        spos[0][i] = svel[0][i] = spos[1][i] = svel[1][i] = make_float4(0, 0, 0, 0);
    }
    int cur = 0;
    unsigned flags = 0;
    const int lead_stride = P.n_leads * 8;
    for (uint32_t step = 0; step < nsteps; ++step) {
        __syncthreads();
        // (the by-value params stay read-only: writing P would move all of it to local memory)
        const float *leads = lead_table ? lead_table + (size_t)min(step, lead_rows - 1) * lead_stride : P.leads;
        if (mine) {
            const float4 pi4 = spos[cur][i], vi4 = svel[cur][i];
            const Self self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
            V3 acc = v3zero();
            if (!P.steering_overrides) {
                for (uint32_t j0 = 0; j0 < n; j0 += SM_BATCH) {
                    float4 pj[SM_BATCH], vj[SM_BATCH];
                    V3 d[SM_BATCH];
                    float m2[SM_BATCH];
                    bool fast = P.fast_ok != 0;
#pragma unroll
                    for (int u = 0; u < SM_BATCH; ++u) {
                        pj[u] = spos[cur][j0 + u];
                        vj[u] = svel[cur][j0 + u];
                    }
#pragma unroll
                    for (int u = 0; u < SM_BATCH; ++u) {
                        m2[u] = pair_m2(self, v3(pj[u].x, pj[u].y, pj[u].z), d[u]);
                        // pairs beyond the gate never reach the force code, whatever their m2
                        fast = fast && ((m2[u] >= FAST_M2_LO && m2[u] <= FAST_M2_HI) || m2[u] >= P.m2_cut);
                    }
                    if (fast) {
                        V3 f[SM_BATCH];
                        bool vis[SM_BATCH];
#pragma unroll
                        for (int u = 0; u < SM_BATCH; ++u)
                            f[u] = pair_force_fast(P, self, d[u], m2[u], v3(vj[u].x, vj[u].y, vj[u].z), vis[u]);
#pragma unroll
                        for (int u = 0; u < SM_BATCH; ++u)
                            if (j0 + u < n && !(m2[u] >= P.m2_cut) && vis[u]) acc = vadd(acc, f[u]);
                    } else {  // coincident boids (the self pair, m2 = 0), extreme distances, odd configs
#pragma unroll
                        for (int u = 0; u < SM_BATCH; ++u) {
                            V3 c;
                            if (j0 + u < n && !(m2[u] >= P.m2_cut) &&
                                pair_flock(P, self, d[u], m2[u], v3(vj[u].x, vj[u].y, vj[u].z), c))
                                acc = vadd(acc, c);
                        }
                    }
                }
            }
            Extras e;
            const V3 a = accel_total(P, self, acc, e, flags, false, leads);
            V3 np, nv;
            euler(P, self.p, self.v, a, np, nv);
            spos[cur ^ 1][i] = make_float4(np.x, np.y, np.z, pi4.w);
            svel[cur ^ 1][i] = make_float4(nv.x, nv.y, nv.z, 0.0f);
        }
        cur ^= 1;
    }
    __syncthreads();
    if (mine) small_store(gpos, gvel, aos_out, i, spos[cur][i], svel[cur][i]);
    if (flags) atomicOr(status, flags);
}

// ---- K4, lane-parallel form for N <= 128 (the demo's 110 boids) ---------------------------------
// The one-thread-per-boid kernel above leaves 7/8 of the SM idle and spends a step walking 110
// dependent pair evaluations per thread.  Here the N^2 exact pair terms of a step are evaluated by
// all 1024 threads of the CTA at once -- S2_LANES lanes per boid, lane s takes j = s, s + 8, ... --
// and written to shared memory as c[i][j]; then one thread per boid adds its row in ascending j,
// which is the reference's loop order (flocking.rs:136): same terms, same order, same bits.  The
// row sum is a chain of N dependent f32 adds (~0.25 us); the pair terms, the expensive part, run
// eight-wide.  Demo scene: 24 -> ~4 us per step.
constexpr int S2_THREADS = 1024;
constexpr int S2_MAX = 128;
constexpr int S2_STRIDE = S2_MAX + 1;  // odd row stride: the row sums of 32 boids hit 32 banks
constexpr int S2_LANES = 8;

uint32_t small2_max_boids() { return S2_MAX; }

struct Small2Smem {
    float4 pos[2][S2_MAX], vel[2][S2_MAX];
    float cx[S2_MAX * S2_STRIDE], cy[S2_MAX * S2_STRIDE], cz[S2_MAX * S2_STRIDE];
    unsigned char valid[S2_MAX * S2_MAX];  // the pair contributes (in range and visible)
};

__global__ void __launch_bounds__(S2_THREADS, 1)
small2_kernel(const DevParams P, float4 *__restrict__ gpos, float4 *__restrict__ gvel, uint32_t n,
              uint32_t nsteps, const float *__restrict__ lead_table, uint32_t lead_rows,
              unsigned *__restrict__ status, const float *__restrict__ aos_in, float *__restrict__ aos_out,
              uint32_t first_index) {
    extern __shared__ __align__(16) unsigned char s2_raw[];
    Small2Smem &S = *reinterpret_cast<Small2Smem *>(s2_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t i = tid / S2_LANES, lane = tid % S2_LANES;
    if (tid < n) small_load(gpos, gvel, aos_in, first_index, tid, S.pos[0][tid], S.vel[0][tid]);
    int cur = 0;
    unsigned flags = 0;
    const int lead_stride = P.n_leads * 8;
    for (uint32_t step = 0; step < nsteps; ++step) {
        __syncthreads();
        const float *leads = lead_table ? lead_table + (size_t)min(step, lead_rows - 1) * lead_stride : P.leads;
        if (i < n && !P.steering_overrides) {
            const float4 pi4 = S.pos[cur][i], vi4 = S.vel[cur][i];
            const Self self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
            for (uint32_t j = lane; j < n; j += S2_LANES) {
                const float4 pj = S.pos[cur][j];
                V3 d, c = v3zero();
                const float m2 = pair_m2(self, v3(pj.x, pj.y, pj.z), d);
                bool ok = false;
                if (!(m2 >= P.m2_cut)) {
                    const float4 vj = S.vel[cur][j];
                    ok = pair_flock(P, self, d, m2, v3(vj.x, vj.y, vj.z), c);
                }
                S.cx[i * S2_STRIDE + j] = c.x;
                S.cy[i * S2_STRIDE + j] = c.y;
                S.cz[i * S2_STRIDE + j] = c.z;
                S.valid[i * S2_MAX + j] = ok ? 1 : 0;
            }
        }
        __syncthreads();
        if (tid < n) {
            const float4 pi4 = S.pos[cur][tid], vi4 = S.vel[cur][tid];
            const Self self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
            V3 acc = v3zero();
            if (!P.steering_overrides) {
                const float *rx = S.cx + tid * S2_STRIDE, *ry = S.cy + tid * S2_STRIDE, *rz = S.cz + tid * S2_STRIDE;
                const unsigned char *rv = S.valid + tid * S2_MAX;
                for (uint32_t j = 0; j < n; ++j)
                    if (rv[j]) acc = vadd(acc, v3(rx[j], ry[j], rz[j]));  // ascending j: the reference's order
            }
            Extras e;
            const V3 a = accel_total(P, self, acc, e, flags, false, leads);
            V3 np, nv;
            euler(P, self.p, self.v, a, np, nv);
            S.pos[cur ^ 1][tid] = make_float4(np.x, np.y, np.z, pi4.w);
            S.vel[cur ^ 1][tid] = make_float4(nv.x, nv.y, nv.z, 0.0f);
        }
        cur ^= 1;
    }
    __syncthreads();
    if (tid < n) small_store(gpos, gvel, aos_out, tid, S.pos[cur][tid], S.vel[cur][tid]);
    if (flags) atomicOr(status, flags);
}

int launch_small(cudaStream_t st, const DevParams &P, float4 *pos, float4 *vel, uint32_t n,
                 uint32_t nsteps, const float *lead_table, uint32_t lead_rows, unsigned *status,
                 const float *aos_in, float *aos_out, uint32_t first_index) {
    if (!n || !nsteps) return FP_OK;
    if (n > SM_MAX) {
        set_error("flock too large for the single-CTA kernel");
        return FP_ERR_INVALID;
    }
    static const bool lane_parallel = [] {  // FP_SMALL_VARIANT=1: the one-thread-per-boid kernel (cross-check)
        const char *e = getenv("FP_SMALL_VARIANT");
        return !(e && *e == '1');
    }();
    if (n <= S2_MAX && lane_parallel) {
        const int smem = (int)sizeof(Small2Smem);
        FP_CUDA(cudaFuncSetAttribute(small2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        small2_kernel<<<1, S2_THREADS, smem, st>>>(P, pos, vel, n, nsteps, lead_rows ? lead_table : nullptr, lead_rows,
                                                  status, aos_in, aos_out, first_index);
    } else {
        small_kernel<<<1, SM_THREADS, 0, st>>>(P, pos, vel, n, nsteps, lead_rows ? lead_table : nullptr, lead_rows,
                                              status, aos_in, aos_out, first_index);
    }
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

}  // namespace fp

// fp_small.cu -- K4: demo-sized flocks (config C1: 110 boids, 1000 steps).
//
// At N ~ 100 a step is launch- and latency-bound, so one CTA keeps the whole
// flock in shared memory and advances `nsteps` steps in a single launch
// (lead-boid rows pre-tabulated by the host).  A warp owns a row i:
// its lanes evaluate the pair function against j = lane, lane+32, ... and
// park each contribution in shared memory; three lanes then add the x, y, z
// planes sequentially in ascending j -- the reference's summation order
// (flocking.rs:133-151) -- so results are bit-identical to the Rust loop.
#include "fp_internal.h"

namespace fp {

constexpr int SM_THREADS = 512;
constexpr int SM_WARPS = SM_THREADS / 32;
constexpr uint32_t SM_MAX = 256;

uint32_t small_max_boids() { return SM_MAX; }

struct SmallSmem {
    float4 pos[2][SM_MAX];
    float4 vel[2][SM_MAX];
    float vhx[SM_MAX], vhy[SM_MAX], vhz[SM_MAX];
    float contrib[SM_WARPS][3][SM_MAX];
};

__global__ void __launch_bounds__(SM_THREADS, 1)
small_kernel(DevParams P, float4 *__restrict__ gpos, float4 *__restrict__ gvel, uint32_t n,
             uint32_t nsteps, const float *__restrict__ lead_table, uint32_t lead_rows,
             unsigned *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SmallSmem &S = *reinterpret_cast<SmallSmem *>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < n; i += SM_THREADS) {
        S.pos[0][i] = gpos[i];
        S.vel[0][i] = gvel[i];
    }
    int cur = 0;
    unsigned flags = 0;
    const int lead_stride = P.n_leads * 8;
    for (uint32_t step = 0; step < nsteps; ++step) {
        __syncthreads();
        if (lead_table) P.leads = lead_table + (size_t)min(step, lead_rows - 1) * lead_stride;
        // normalize(v_i) once per boid per step (boid.rs:103)
        for (uint32_t i = threadIdx.x; i < n; i += SM_THREADS) {
            const float4 v = S.vel[cur][i];
            const V3 h = vnormalize(v3(v.x, v.y, v.z));
            S.vhx[i] = h.x; S.vhy[i] = h.y; S.vhz[i] = h.z;
        }
        __syncthreads();
        for (uint32_t i = wid; i < n; i += SM_WARPS) {
            const float4 pi4 = S.pos[cur][i], vi4 = S.vel[cur][i];
            Self self;
            self.p = v3(pi4.x, pi4.y, pi4.z);
            self.v = v3(vi4.x, vi4.y, vi4.z);
            self.vhat = v3(S.vhx[i], S.vhy[i], S.vhz[i]);
            if (!P.steering_overrides) {
                for (uint32_t j = lane; j < n; j += 32) {
                    const float4 pj = S.pos[cur][j];
                    V3 d, c = v3zero();
                    const float m2 = pair_m2(self, v3(pj.x, pj.y, pj.z), d);
                    if (!(m2 >= P.m2_cut)) {
                        const float4 vj = S.vel[cur][j];
                        V3 t;
                        if (pair_flock(P, self, d, m2, v3(vj.x, vj.y, vj.z), t)) c = t;
                    }
                    S.contrib[wid][0][j] = c.x;
                    S.contrib[wid][1][j] = c.y;
                    S.contrib[wid][2][j] = c.z;
                }
            }
            __syncwarp();
            // lanes 0..2 add one component plane each, ascending j; x + (+0) == x exactly
            float acc = 0.0f;
            if (lane < 3 && !P.steering_overrides) {
                const float *plane = S.contrib[wid][lane];
                for (uint32_t j = 0; j < n; ++j) acc = fadd(acc, plane[j]);
            }
            const V3 a_boids = v3(__shfl_sync(0xffffffffu, acc, 0), __shfl_sync(0xffffffffu, acc, 1),
                                  __shfl_sync(0xffffffffu, acc, 2));
            if (lane == 0) {
                Extras e;
                const V3 a = accel_total(P, self, a_boids, e, flags);
                V3 np, nv;
                euler(P, self.p, self.v, a, np, nv);
                S.pos[cur ^ 1][i] = make_float4(np.x, np.y, np.z, pi4.w);
                S.vel[cur ^ 1][i] = make_float4(nv.x, nv.y, nv.z, 0.0f);
            }
            __syncwarp();
        }
        cur ^= 1;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += SM_THREADS) {
        gpos[i] = S.pos[cur][i];
        gvel[i] = S.vel[cur][i];
    }
    if (flags) atomicOr(status, flags);
}

int launch_small(cudaStream_t st, const DevParams &P, float4 *pos, float4 *vel, uint32_t n,
                 uint32_t nsteps, const float *lead_table, uint32_t lead_rows, unsigned *status) {
    if (!n || !nsteps) return FP_OK;
    if (n > SM_MAX) {
        set_error("flock too large for the single-CTA kernel");
        return FP_ERR_INVALID;
    }
    FP_CUDA(cudaFuncSetAttribute(small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sizeof(SmallSmem)));
    small_kernel<<<1, SM_THREADS, sizeof(SmallSmem), st>>>(P, pos, vel, n, nsteps,
                                                          lead_rows ? lead_table : nullptr, lead_rows,
                                                          status);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

}  // namespace fp

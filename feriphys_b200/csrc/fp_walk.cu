// fp_walk.cu -- K3, production form: shared-memory staged 27-cell walk.
//
// A CTA owns BLOCK consecutive boids of the cell-sorted state (~BLOCK/8 cells
// of one grid row; rows run along z, the fastest key dimension).  A boid's home
// cell is the key its slot was binned under (lazy re-binning: the position may
// have drifted by up to skin / 2 since).  For each of the nine (dx, dy) neighbour rows, everything
// its threads can reach is ONE contiguous slot interval of the sorted position
// array, so the CTA stages those nine intervals into shared memory with 1-D
// TMA bulk copies (cp.async.bulk -> UBLKCP) signalled on an mbarrier: every
// candidate position is fetched from L2/HBM once per CTA and then re-read ~20
// times from shared memory.  Each thread then takes its own boid through two
// warp-convergent phases:
//   1. pre-gate -- every candidate, packed FP32 (two candidates per instruction), fused and
//                approximate on purpose: squared distance against m2_cut (1 + 1e-6) and a
//                conservative field-of-view test (drops only pairs culled with a 1e-5
//                margin) -- a superset of the contributing pairs; the survivors' tile
//                offsets (16 bit) go to a per-thread list in shared memory;
//   2. forces   -- survivors only: the exact squared distance against m2_cut, then the exact
//                pair function (which re-tests the FOV exactly), accumulated in list
//                (= slot) order.
// A full list is drained (phase 2) before the next chunk, warp-uniformly, and
// the row loop is kept rolled so the kernel stays inside the instruction cache.
//
// Arithmetic and summation order are exactly those of the one-phase kernel
// (grid_walk_kernel): bit-identical to the reference loop (flocking.rs:133-151)
// run over the same boids in cell-sorted order.  A boid's own slot is skipped:
// for finite states its self-pair contributes exactly +0 (boid.rs:111,121,132).
//
// CTAs whose nine intervals exceed the tile (very dense clusters) take the
// one-phase path straight from global memory -- slower, same result.
#include "fp_walk_stage.cuh"  // TMA bulk-copy / mbarrier primitives

namespace fp {

template <int BLOCK, int TILE_CAP, int CAP>
struct Walk3Smem {
    // staged candidate positions, SoA (+ padding for masked tail reads).  SoA so that two
    // neighbouring candidates load as one 64-bit pair for the packed FP32 gate.
    alignas(16) float tx[TILE_CAP + 8], ty[TILE_CAP + 8], tz[TILE_CAP + 8];
    uint16_t list[CAP][BLOCK];   // per-thread survivor lists: tile offsets
    uint32_t rng[10][BLOCK];     // per-thread (tile start | len << 16) per row; row 9 = final drain
    uint32_t ub[9], ue[9];       // CTA-wide slot interval of each row, widened to multiples of 4
    uint32_t toff[10];           // tile offset of each interval (prefix sum)
    uint32_t tslot[9];           // slot - tile offset within each interval (ub[r] - toff[r])
    int fallback;
    alignas(8) uint64_t bar;
};

template <int TAP, int BLOCK, int TILE_CAP, int CAP>
__global__ void __launch_bounds__(BLOCK, 768 / BLOCK)
grid_walk3_kernel(const DevParams P, const GridDesc g, const WalkIO io, unsigned *__restrict__ status,
                  TapOut tap) {
    static_assert(TILE_CAP + 8 <= 4096, "list entries carry a 12-bit tile offset");
    if (TAP == TAP_STEP && io.ctl && io.ctl->stale) return;  // lazy re-binning: this step is void
    const float4 *__restrict__ pos_s = io.pos_s;
    const float4 *__restrict__ vel_s = io.vel_s;
    const float *__restrict__ sx = io.soa_in[0];
    const float *__restrict__ sy = io.soa_in[1];
    const float *__restrict__ sz = io.soa_in[2];
    const uint32_t *__restrict__ cell_start = io.cell_start;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = Walk3Smem<BLOCK, TILE_CAP, CAP>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t s = io.first + blockIdx.x * BLOCK + tid;
    const bool active = s < io.last;

    if (tid == 0) {
        mbar_init(&S.bar, 1);
        S.fallback = 0;
    }
    if (tid < 9) {
        S.ub[tid] = 0xffffffffu;
        S.ue[tid] = 0u;
    }

    float4 pi4 = make_float4(0, 0, 0, 0), vi4 = make_float4(0, 0, 0, 0);
    Self self;
    self.p = self.v = self.vhat = v3zero();
    bool work = false;
    int cx = 0, cy = 0, cz = 0;
    if (active) {
        pi4 = pos_s[s];
        vi4 = vel_s[s];
    }
    if (TAP == TAP_STEP && io.ctl) track_motion(io.ctl, active, pi4, vi4);
    if (active) {
        self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
        const bool ghost = __float_as_uint(vi4.w) != 0u;
        work = !ghost && ((TAP != TAP_STEP) || !P.steering_overrides);
        home_cell(g, __ldg(io.home + (s - io.first)), cx, cy, cz);  // the cell it was binned under
    }
    // the nine slot ranges of this boid, rows in ascending key order (dx outer, dy inner)
    uint32_t jb[9], je[9];
    {
        const int z0 = max(cz - g.zspan, 0), z1 = min(cz + g.zspan, g.dim[2] - 1);
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int x = cx + r / 3 - 1, y = cy + r % 3 - 1;
            jb[r] = je[r] = 0;
            if (work && x >= 0 && x < g.dim[0] && y >= 0 && y < g.dim[1]) {
                const uint32_t rowbase = row_base(g, x, y);
                jb[r] = __ldg(cell_start + rowbase + z0);
                je[r] = __ldg(cell_start + rowbase + z1 + 1);
            }
        }
    }
    __syncthreads();  // barrier + ub/ue initialised
    // CTA-wide union of each row's ranges: one contiguous interval per row
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const bool has = je[r] > jb[r];
        const uint32_t lo = __reduce_min_sync(0xffffffffu, has ? jb[r] : 0xffffffffu);
        const uint32_t hi = __reduce_max_sync(0xffffffffu, has ? je[r] : 0u);
        if ((tid & 31) == 0 && hi > 0) {
            atomicMin(&S.ub[r], lo);
            atomicMax(&S.ue[r], hi);
        }
    }
    __syncthreads();
    if (tid < 32) {
        // warp 0 lays the nine intervals out in the tile (lane r = row r) and issues the copies
        uint32_t ub = tid < 9 ? S.ub[tid] : 0u, ue = tid < 9 ? S.ue[tid] : 0u;
        if (ue > ub) {  // 16-byte granules for the 4-byte SoA arrays
            ub &= ~3u;
            ue = (ue + 3u) & ~3u;
        } else {
            ub = ue = 0u;
        }
        const uint32_t len = ue - ub;
        uint32_t inc = len;  // inclusive prefix sum over the lanes
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
            if ((int)tid >= off) inc += t;
        }
        const uint32_t toff = inc - len, total = __shfl_sync(0xffffffffu, inc, 8);
        if (tid < 9) {
            S.ub[tid] = ub;
            S.ue[tid] = ue;
            S.toff[tid] = toff;
            S.tslot[tid] = ub - toff;
        }
        const bool staged = total > 0 && total <= (uint32_t)TILE_CAP;
        if (tid == 0) {
            S.toff[9] = total;
            if (total > (uint32_t)TILE_CAP) S.fallback = 1;
            if (staged) mbar_expect_tx(&S.bar, total * 12u);
        }
        __syncwarp();
        if (staged && tid < 9 && len) {
            bulk_g2s(&S.tx[toff], sx + ub, len * 4u, &S.bar);
            bulk_g2s(&S.ty[toff], sy + ub, len * 4u, &S.bar);
            bulk_g2s(&S.tz[toff], sz + ub, len * 4u, &S.bar);
        }
    }
    __syncthreads();
    V3 acc = v3zero();
    if (S.fallback) {
        // one-phase walk from global memory (dense cluster: the tile would not fit)
        if (work) {
            const int z0 = max(cz - g.zspan, 0), z1 = min(cz + g.zspan, g.dim[2] - 1);
            for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x) {
                for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
                    const uint32_t rowbase = row_base(g, x, y);
                    const uint32_t b = __ldg(cell_start + rowbase + z0);
                    const uint32_t e = __ldg(cell_start + rowbase + z1 + 1);
                    for (uint32_t j = b; j < e; ++j) {
                        if (j == s) continue;
                        const float4 pj = __ldg(pos_s + j);
                        V3 d;
                        const float m2 = pair_m2(self, v3(pj.x, pj.y, pj.z), d);
                        if (m2 >= P.m2_cut) continue;
                        const float4 vj = __ldg(vel_s + j);
                        V3 contrib;
                        if (pair_flock(P, self, d, m2, v3(vj.x, vj.y, vj.z), contrib))
                            acc = vadd(acc, contrib);
                    }
                }
            }
        }
    } else {
        // per-thread ranges as (tile start | len << 16), so the row loop below stays rolled
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            const uint32_t len = je[r] - jb[r];
            const uint32_t t0 = len ? S.toff[r] + (jb[r] - S.ub[r]) : 0u;
            S.rng[r][tid] = t0 | (len << 16);
        }
        S.rng[9][tid] = 0u;
        // tile offset of this boid's own slot (row 4 = its own grid row); 0xffff if not staged
        const uint32_t t_self = (work && je[4] > jb[4]) ? S.toff[4] + (s - S.ub[4]) : 0xffffu;
        if (S.toff[9] > 0) mbar_wait(&S.bar, 0);
        uint16_t *const lst = &S.list[0][tid];  // entry k at lst[k * BLOCK]
        int cnt = 0;
        // -p_i broadcast into both halves: p_j + (-p_i) == p_j - p_i exactly
        const float2 nsx = make_float2(-self.p.x, -self.p.x), nsy = make_float2(-self.p.y, -self.p.y),
                     nsz = make_float2(-self.p.z, -self.p.z);
        const float2 vhx = make_float2(self.vhat.x, self.vhat.x), vhy = make_float2(self.vhat.y, self.vhat.y),
                     vhz = make_float2(self.vhat.z, self.vhat.z);
        const float2 kh2 = make_float2(P.fov_kh, P.fov_kh), kl2 = make_float2(P.fov_kl, P.fov_kl);
#pragma unroll 1
        for (int r = 0; r <= 9; ++r) {
            const uint32_t pk = S.rng[r][tid];
            const uint32_t t0 = pk & 0xffffu, len = pk >> 16;
            uint32_t i = 0;  // this lane's progress through its row range
#pragma unroll 1
            for (;;) {
                // room every lane is guaranteed to have left in its list
                int room = CAP - (int)__reduce_max_sync(0xffffffffu, (unsigned)cnt);
                const bool more = __any_sync(0xffffffffu, i < len);
                if (room == 0 || (r == 9 && room < CAP)) {
                    // drain: exact forces in list (= slot) order.  An entry is (row << 12 | tile
                    // offset); the row turns the offset back into a slot for the velocity gather.
                    const int nb = cnt;
                    auto slot_of = [&](uint32_t e) { return (e & 0xfffu) + S.tslot[e >> 12]; };
                    // Two entries per trip: their (branch-free) force evaluations are independent
                    // and interleave; the two adds into acc stay in list order.
                    // Velocities are gathered TWO trips ahead (four loads in flight): one trip of
                    // arithmetic does not cover an L2 miss (ncu: the single-trip version spent 17 %
                    // of the kernel in long-scoreboard stalls here).
                    uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
                    float4 w0 = make_float4(0, 0, 0, 0), w1 = w0, w2 = w0, w3 = w0;
                    if (nb > 0) { e0 = lst[0]; w0 = __ldg(vel_s + slot_of(e0)); }
                    if (nb > 1) { e1 = lst[BLOCK]; w1 = __ldg(vel_s + slot_of(e1)); }
                    if (nb > 2) { e2 = lst[2 * BLOCK]; w2 = __ldg(vel_s + slot_of(e2)); }
                    if (nb > 3) { e3 = lst[3 * BLOCK]; w3 = __ldg(vel_s + slot_of(e3)); }
                    for (int k = 0; k < nb; k += 2) {
                        const uint32_t ta = e0, tb = e1;
                        const float4 va = w0, vb = w1;
                        const bool hasb = k + 1 < nb;
                        e0 = e2; e1 = e3; w0 = w2; w1 = w3;
                        if (k + 4 < nb) { e2 = lst[(k + 4) * BLOCK]; w2 = __ldg(vel_s + slot_of(e2)); }
                        if (k + 5 < nb) { e3 = lst[(k + 5) * BLOCK]; w3 = __ldg(vel_s + slot_of(e3)); }
                        const uint32_t ia = ta & 0xfffu, ib = tb & 0xfffu;
                        const V3 pa = v3(S.tx[ia], S.ty[ia], S.tz[ia]), pb = v3(S.tx[ib], S.ty[ib], S.tz[ib]);
                        V3 da, db;
                        const float ma = pair_m2(self, v3(pa.x, pa.y, pa.z), da);
                        const float mb = pair_m2(self, v3(pb.x, pb.y, pb.z), db);
                        // what counts: not the boid itself (flocking.rs:137-139), and in range by the
                        // EXACT squared distance (the pre-gate let a sliver too many through)
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
                        } else {  // extreme distances (coincident boids, ...): generic exact path
                            V3 contrib;
                            if (oka &&
                                pair_inrange<false>(P, self, da, ma, v3(va.x, va.y, va.z), 1.0f, P.cstar, contrib))
                                acc = vadd(acc, contrib);
                            if (okb &&
                                pair_inrange<false>(P, self, db, mb, v3(vb.x, vb.y, vb.z), 1.0f, P.cstar, contrib))
                                acc = vadd(acc, contrib);
                        }
                    }
                    cnt = 0;
                    room = CAP;
                }
                if (!more) break;
                const uint32_t i1 = min(i + (uint32_t)room, len);
                // phase 1, the pre-gate.  Four candidates per batch, all tile loads issued ahead of
                // the list stores; full batches carry no tail logic, and the (masked) tail may
                // read up to three records past its range -- the tile is padded for that.
                uint32_t w = (uint32_t)cnt * BLOCK;  // list cursor, in entries
                const uint32_t tag = (uint32_t)r << 12;
                // Batches of four candidates at an even tile index: two neighbours load as one
                // 64-bit pair and go through the packed FP32 pipe (FADD2 / FMUL2 / FFMA2, sm_100).
                // This is a PRE-gate, deliberately fused and approximate -- it only has to keep a
                // superset of the pairs that contribute; the drain re-tests exactly:
                //  * distance: the fused sum of squares is within 4e-7 relative of the reference's
                //    separately rounded one; everything below m2_cut_hi = m2_cut (1 + 1e-6) stays;
                //  * field of view: c~ = q / sqrt(m2), q = vhat . d, is within 1e-6 of the exact
                //    cosine (boid.rs:102-105); a pair is dropped only when it is culled with a 1e-5
                //    margin, -1 + 1e-5 < c~ < cstar - 1e-5, tested without the square root through
                //    the monotone map x -> x |x|:  KL m2 < q |q| < KH m2.  NaN never drops.
                auto gate4 = [&](uint32_t T, uint32_t live) {  // live: bit u set <=> candidate T+u counts
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
                    const float2 q01 = __ffma2_rn(vhz, dz01, __ffma2_rn(vhy, dy01, __fmul2_rn(vhx, dx01)));
                    const float2 q23 = __ffma2_rn(vhz, dz23, __ffma2_rn(vhy, dy23, __fmul2_rn(vhx, dx23)));
                    const float2 s01 = __fmul2_rn(q01, make_float2(fabsf(q01.x), fabsf(q01.y)));
                    const float2 s23 = __fmul2_rn(q23, make_float2(fabsf(q23.x), fabsf(q23.y)));
                    const float2 h01 = __fmul2_rn(kh2, m01), h23 = __fmul2_rn(kh2, m23);
                    const float2 l01 = __fmul2_rn(kl2, m01), l23 = __fmul2_rn(kl2, m23);
                    const float mm[4] = {m01.x, m01.y, m23.x, m23.y};
                    const float ss[4] = {s01.x, s01.y, s23.x, s23.y};
                    const float hh[4] = {h01.x, h01.y, h23.x, h23.y};
                    const float ll[4] = {l01.x, l01.y, l23.x, l23.y};
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if ((live >> u & 1u) && !(mm[u] >= P.m2_cut_hi) && !(ss[u] < hh[u] && ss[u] > ll[u])) {
                            lst[w] = (uint16_t)(tag | (T + u));
                            w += BLOCK;
                        }
                };
                if (i < i1) {
                    const uint32_t A = t0 + i, B = t0 + i1;  // tile index range of this pass
                    uint32_t T = A & ~1u;
                    if (T < A) {  // odd start: first batch drops the slot before the range
                        gate4(T, (B - T >= 4 ? 0xeu : ((1u << (B - T)) - 1u) & 0xeu));
                        T += 4;
                    }
                    for (; T + 4 <= B; T += 4) gate4(T, 0xfu);
                    if (T < B) gate4(T, (1u << (B - T)) - 1u);
                }
                i = i1;
                cnt = (int)(w / BLOCK);
            }
        }
    }
    if (!active) return;
    walk_finish<TAP>(P, s, pi4, vi4, self, acc, 0u, 0ull, io, status, tap);
}

template <int TAP, int BLOCK, int TILE_CAP, int CAP>
static int launch3(cudaStream_t st, const DevParams &P, const GridDesc &g, const WalkIO &io, unsigned *status,
                   const TapOut &tap_out) {
    auto kern = grid_walk3_kernel<TAP, BLOCK, TILE_CAP, CAP>;
    const int smem = (int)sizeof(Walk3Smem<BLOCK, TILE_CAP, CAP>);
    FP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(io.last - io.first + BLOCK - 1) / BLOCK, BLOCK, smem, st>>>(P, g, io, status, tap_out);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

int launch_grid_walk3(cudaStream_t st, const DevParams &P, const GridDesc &g, int tap, const WalkIO &io,
                      unsigned *status, const TapOut &tap_out) {
    // 128 boids, a tile of 1904 candidates, 64-entry survivor lists: 43.6 KB, five CTAs per SM; the
    // tile overflows ~never at the densities lists are built for, and one drain per boid almost
    // always (lists average 17 entries).  Other shapes were measured and lost.
    return tap == TAP_STEP ? launch3<TAP_STEP, 128, 1904, 64>(st, P, g, io, status, tap_out)
                           : launch3<TAP_ACCEL, 128, 1904, 64>(st, P, g, io, status, tap_out);
}

}  // namespace fp

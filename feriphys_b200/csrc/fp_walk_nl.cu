// fp_walk_nl.cu -- K3 on standing candidate lists: the grid step's default form.
//
// With lazy re-binning a binning stands for dozens of steps, yet the plain staged
// walk (fp_walk.cu) pre-gates all ~160 candidates of a boid's 27 home cells in every one of them
// -- 48 % of its instructions (profiles/r1_c4_walk_hotspots.txt) -- to find the ~34 that are within
// reach.  Here that work is done ONCE per binning:
//
//   nl_build_kernel  (after a binning)
//       stages the same nine intervals as the plain walk, keeps every candidate whose squared
//       distance is below (reach + skin)^2 (1 + 1e-5), and writes the survivors' tile offsets
//       (16 bit, row << 12 | offset, ascending slot order, the boid itself left out) to a
//       per-thread list in global memory, [cta][entry / 4][thread][entry % 4]: a thread fetches four
//       entries with one 8-byte load, a warp's batch is one 256-byte run; also the per-boid count
//       and the CTA's tile layout (nine intervals).
//   nl_walk_kernel   (EXACT numerics: every step until the next binning)
//       stages the tile from the cached layout (no cell-table look-ups, no reductions), runs the
//       plain walk's fused pre-gate over the ~34 cached entries instead of ~160 candidates, and
//       drains the survivors exactly as the plain walk does: bit-identical to it.
//   nl_fast_kernel   (FAST numerics, fp_flock_set_numerics)
//       one pass over the cached entries with positions AND velocities staged in shared memory:
//       exact squared distance (so the distance decisions are the reference's), the field-of-view
//       decision on a fused cosine outside a 1e-5 guard band and by the exact sequence inside it,
//       forces with FMA and MUFU.RSQ.  Neighbour sets are bit-exact, accelerations agree with the
//       reference to ~1e-6 relative (bar: 1e-5).  Its lists are built VISIBLE FIRST: the entries the
//       boid is predicted to see (field of view at build time) come before the ones it is not, so a
//       warp's lanes agree far more often on whether an entry contributes -- the force code runs for
//       the first ~half of the rows with most lanes active and is skipped for the rest (in slot
//       order it ran on every row with a third of the lanes: profiles/r2_c4_nl_fast_v1_*).
//
// Exactness of the lists.  While the binning stands every boid is within skin / 2 of where it was
// binned (the device-checked displacement bound D), so a pair closer than reach now was closer
// than reach + skin then: the cached list is a superset of every pair that can contribute, in
// slot order.
//
// Capacity.  A list holds nl.vcap entries.  A CTA in which some boid has more, or whose nine
// intervals do not fit the tile, is marked in its layout record and walks the 27 cells from
// global memory every step (the plain walk's own path for CTAs that overflow the tile): slower,
// same result, no effect on any other CTA.  The build counts such CTAs; the host turns the lists
// off when they stop being rare.
#include "fp_walk_stage.cuh"

namespace fp {

namespace {

constexpr int NL_BLOCK = 128;   // threads (= boids) per CTA, as the plain walk
constexpr int NL_TILE = 1904;   // staged candidates per CTA (build and exact walk), as the plain walk
constexpr int NL_CAP = 48;      // survivor list of the exact walk (shared memory): six CTAs per SM
constexpr int NF_TILE = 1568;   // staged candidates per CTA of the fast walk, form A (positions + velocities)
constexpr int NF_TILE_B = 2560; // ... form B (positions only; the build kernel stages this many too)
constexpr int NL_CTA_WORDS = 20;  // cached tile layout per CTA: ub[9], ue[9], [18] = no lists, 1 spare

// this CTA gets no lists: it walks from global memory every step (counted once per CTA)
__device__ __forceinline__ void nl_no_lists(const NlIO &nl, uint32_t *tab) {
    if (atomicExch(tab + 18, 1u) == 0u) atomicAdd(nl.flag, 1u);
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ uint32_t ld_nc_u32(const uint32_t *p) {  // (volatile: issued where it stands)
    uint32_t v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// What every list walk starts with.  All loads of the prologue are issued before anything waits on
// one of them: the step's validity, the CTA's layout record, the boid, its list length, its first
// two words of entries (the start-up chain layout -> TMA -> tile was a third of the kernel's stall
// samples while each link waited for the one before, profiles/r2_c4_nl_fast_v3_*).
// Measured and dropped: warming L2 for the CTA that will take this one's place (prefetch.global.L2 of
// its boids and entries, cp.async.bulk.prefetch.L2 of its tile): L2 hits rose from 49 to 63 %, the
// wait at the start-up barrier did not move, the step lost 5 % (profiles/r2_bench_c4_l2warm.json).
struct NlPrologue {
    uint32_t stale, no_lists, ub, ue, n_c;
    float4 pi4, vi4;
    const uint2 *vlp;
    uint2 q0, q1;
};
template <bool STEP>
__device__ __forceinline__ NlPrologue nl_prologue(const WalkIO &io, const NlIO &nl, uint32_t tid, uint32_t s,
                                                  bool active) {
    NlPrologue p;
    const uint32_t *const tab = nl.cta_tab + (size_t)blockIdx.x * NL_CTA_WORDS;
    p.stale = (STEP && io.ctl) ? ld_relaxed_u32(&io.ctl->stale) : 0u;
    p.no_lists = ld_nc_u32(tab + 18);
    p.ub = p.ue = 0;
    if (tid < 9) {
        p.ub = ld_nc_u32(tab + tid);
        p.ue = ld_nc_u32(tab + 9 + tid);
    }
    p.pi4 = p.vi4 = make_float4(0, 0, 0, 0);
    p.n_c = 0;
    if (active) {
        p.pi4 = io.pos_s[s];
        p.vi4 = io.vel_s[s];
        p.n_c = __ldg(nl.count + (s - io.first));
    }
    // this CTA's list block: four entries per 8-byte word, [word][thread]
    p.vlp = reinterpret_cast<const uint2 *>(nl.entries) + (size_t)blockIdx.x * (nl.vcap / 4) * NL_BLOCK + tid;
    p.q0 = __ldcs(p.vlp);
    p.q1 = __ldcs(p.vlp + NL_BLOCK);
    return p;
}

// Lays the nine CTA-wide intervals (ub, ue: multiples of 4, empty = 0, 0) out in the tile and
// issues their bulk copies (positions; velocities too when `tv` is given).  Called by warp 0;
// lane r = row r.  Returns the tile total.
template <class Smem>
__device__ __forceinline__ uint32_t nl_stage(Smem &S, uint32_t tid, uint32_t ub, uint32_t ue,
                                             const float *__restrict__ sx, const float *__restrict__ sy,
                                             const float *__restrict__ sz, uint32_t tile_cap,
                                             float4 *tv = nullptr, const float4 *__restrict__ vel_s = nullptr) {
    const uint32_t len = ue - ub;
    uint32_t inc = len;  // inclusive prefix sum over the lanes
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
        if ((int)tid >= off) inc += t;
    }
    const uint32_t toff = inc - len, total = __shfl_sync(0xffffffffu, inc, 8);
    if (tid < 9) {
        S.toff[tid] = toff;
        S.tslot[tid] = ub - toff;
    }
    const bool staged = total > 0 && total <= tile_cap;
    if (tid == 0) {
        S.toff[9] = total;
        if (staged) mbar_expect_tx(&S.bar, total * (tv ? 28u : 12u));
    }
    __syncwarp();
    if (staged && tid < 9 && len) {
        bulk_g2s(&S.tx[toff], sx + ub, len * 4u, &S.bar);
        bulk_g2s(&S.ty[toff], sy + ub, len * 4u, &S.bar);
        bulk_g2s(&S.tz[toff], sz + ub, len * 4u, &S.bar);
        if (tv) bulk_g2s(tv + toff, vel_s + ub, len * 16u, &S.bar);
    }
    return total;
}

// the 27 home cells from global memory, exact pair function: a CTA without lists
__device__ __forceinline__ V3 nl_walk_global(const DevParams &P, const GridDesc &g, const WalkIO &io, uint32_t s,
                                             const Self &self) {
    V3 acc = v3zero();
    int cx, cy, cz;
    home_cell(g, __ldg(io.home + (s - io.first)), cx, cy, cz);
    const int z0 = max(cz - g.zspan, 0), z1 = min(cz + g.zspan, g.dim[2] - 1);
    for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x) {
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
            const uint32_t rowbase = row_base(g, x, y);
            const uint32_t b = __ldg(io.cell_start + rowbase + z0);
            const uint32_t e = __ldg(io.cell_start + rowbase + z1 + 1);
            for (uint32_t j = b; j < e; ++j) {
                if (j == s) continue;
                const float4 pj = __ldg(io.pos_s + j);
                V3 d;
                const float m2 = pair_m2(self, v3(pj.x, pj.y, pj.z), d);
                if (m2 >= P.m2_cut) continue;
                const float4 vj = __ldg(io.vel_s + j);
                V3 contrib;
                if (pair_flock(P, self, d, m2, v3(vj.x, vj.y, vj.z), contrib)) acc = vadd(acc, contrib);
            }
        }
    }
    return acc;
}

template <int TILE>
struct NlBuildSmem {
    alignas(16) float tx[TILE + 8], ty[TILE + 8], tz[TILE + 8];
    uint32_t rng[9][NL_BLOCK];  // per-thread (tile start | len << 16) per row
    uint32_t ub[9], ue[9];
    uint32_t toff[10], tslot[9];
    alignas(8) uint64_t bar;
};

constexpr int NB_TMP = 64;  // VIS_FIRST: entries held back in shared memory (the ones not in view)

// VIS_FIRST (the fast walk's lists): entries the boid is predicted to SEE -- fused cosine of the
// sight angle against nl.vis_c, from the velocities of this moment -- are written first, the others
// after them (each class in ascending slot order).  Any order is a correct list; this one makes the
// lanes of a warp agree on whether row k contributes.
template <bool VIS_FIRST, int TILE>
__global__ void __launch_bounds__(NL_BLOCK)
nl_build_kernel(const GridDesc g, const WalkIO io, const NlIO nl) {
    // (No look at ctl->stale: the lists describe the binning, which stands whether or not the
    // step they are built in turns out void.)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using Smem = NlBuildSmem<TILE>;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    // the walk that will use the lists holds nl.tile_cap candidates, or (fast walk, form B) nl.tile_cap_b
    const uint32_t cap_max = nl.tile_cap_b ? nl.tile_cap_b : nl.tile_cap;
    const uint32_t tid = threadIdx.x;
    const uint32_t s = io.first + blockIdx.x * NL_BLOCK + tid;
    const bool active = s < io.last;
    uint32_t *const tab = nl.cta_tab + (size_t)blockIdx.x * NL_CTA_WORDS;
    if (tid == 0) {
        mbar_init(&S.bar, 1);
        tab[18] = 0u;  // (ordered before any nl_no_lists of this CTA by the barriers below)
    }
    if (tid < 9) {
        S.ub[tid] = 0xffffffffu;
        S.ue[tid] = 0u;
    }
    float4 pi4 = make_float4(0, 0, 0, 0);
    bool work = false;
    int cx = 0, cy = 0, cz = 0;
    float2 vhx = make_float2(0, 0), vhy = vhx, vhz = vhx;  // VIS_FIRST: direction of flight, both halves
    if (active) {
        pi4 = io.pos_s[s];
        const float4 vi4 = io.vel_s[s];
        work = __float_as_uint(vi4.w) == 0u;  // not a ghost record
        if (VIS_FIRST) {
            const float inv = rsqrtf(fmaf(vi4.z, vi4.z, fmaf(vi4.y, vi4.y, vi4.x * vi4.x)));  // (a prediction)
            vhx = make_float2(vi4.x * inv, vi4.x * inv);
            vhy = make_float2(vi4.y * inv, vi4.y * inv);
            vhz = make_float2(vi4.z * inv, vi4.z * inv);
        }
        home_cell(g, __ldg(io.home + (s - io.first)), cx, cy, cz);
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
                jb[r] = __ldg(io.cell_start + rowbase + z0);
                je[r] = __ldg(io.cell_start + rowbase + z1 + 1);
            }
        }
    }
    __syncthreads();
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
        uint32_t ub = tid < 9 ? S.ub[tid] : 0u, ue = tid < 9 ? S.ue[tid] : 0u;
        if (ue > ub) {  // 16-byte granules for the 4-byte SoA arrays
            ub &= ~3u;
            ue = (ue + 3u) & ~3u;
        } else {
            ub = ue = 0u;
        }
        if (tid < 9) {
            S.ub[tid] = ub;
            S.ue[tid] = ue;
            tab[tid] = ub;       // the layout every walk launch of this binning re-uses
            tab[9 + tid] = ue;
        }
        nl_stage(S, tid, ub, ue, io.soa_in[0], io.soa_in[1], io.soa_in[2], cap_max);
    }
    __syncthreads();
    const uint32_t total = S.toff[9];
    if (total > cap_max) {  // dense cluster: the tile does not fit -- no lists for this CTA
        if (tid == 0) nl_no_lists(nl, tab);
        return;
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        const uint32_t len = je[r] - jb[r];
        const uint32_t t0 = len ? S.toff[r] + (jb[r] - S.ub[r]) : 0u;
        S.rng[r][tid] = t0 | (len << 16);
    }
    const uint32_t t_self = work ? s - S.tslot[4] : 0xffffu;  // tile offset of the boid itself (row 4)
    if (total > 0) mbar_wait(&S.bar, 0);
    // entry k of this thread: [cta][k / 4][thread][k % 4] -- four consecutive entries are one 8-byte word
    uint16_t *const out = nl.entries + ((size_t)blockIdx.x * (nl.vcap / 4) * NL_BLOCK + tid) * 4;
    auto out_at = [&](uint32_t k) -> uint16_t & { return out[(size_t)(k >> 2) * (NL_BLOCK * 4) + (k & 3u)]; };
    uint16_t *const tmp = reinterpret_cast<uint16_t *>(smem_raw + sizeof(Smem)) + tid;  // VIS_FIRST: [NB_TMP][BLOCK]
    const uint32_t vcap = nl.vcap;
    uint32_t w = 0;   // entries written to the list
    uint32_t wn = 0;  // VIS_FIRST: entries held back (not in view)
    const float vis_c = nl.vis_c;
    const float2 nsx = make_float2(-pi4.x, -pi4.x), nsy = make_float2(-pi4.y, -pi4.y),
                 nsz = make_float2(-pi4.z, -pi4.z);
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const uint32_t pk = S.rng[r][tid];
        const uint32_t t0 = pk & 0xffffu, len = pk >> 16;
        const uint32_t tag = (uint32_t)r << 12;
        // four candidates at an even tile index per batch, packed FP32 (as the plain walk's pre-gate);
        // the fused sum of squares is within 4e-7 relative of the exact one, the cut carries 1e-5
        auto gate4 = [&](uint32_t T, uint32_t live) {
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
            float qq[4] = {0, 0, 0, 0};
            if (VIS_FIRST) {
                const float2 q01 = __ffma2_rn(vhz, dz01, __ffma2_rn(vhy, dy01, __fmul2_rn(vhx, dx01)));
                const float2 q23 = __ffma2_rn(vhz, dz23, __ffma2_rn(vhy, dy23, __fmul2_rn(vhx, dx23)));
                qq[0] = q01.x; qq[1] = q01.y; qq[2] = q23.x; qq[3] = q23.y;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((live >> u & 1u) && !(mm[u] >= nl.m2_wide) && T + u != t_self) {  // NaN never drops
                    const uint16_t e = (uint16_t)(tag | (T + u));
                    // in view <=> cos > vis_c <=> q > vis_c |d|  (q |q| against vis_c |vis_c| m2: no square root)
                    const bool held = VIS_FIRST && wn < (uint32_t)NB_TMP &&
                                      qq[u] * fabsf(qq[u]) <= vis_c * fabsf(vis_c) * mm[u];
                    if (held) {
                        tmp[(size_t)wn * NL_BLOCK] = e;
                        ++wn;
                    } else {
                        if (w < vcap) out_at(w) = e;
                        ++w;
                    }
                }
        };
        if (len) {
            const uint32_t A = t0, B = t0 + len;
            uint32_t T = A & ~1u;
            if (T < A) {  // odd start: the first batch drops the slot before the range
                gate4(T, (B - T >= 4 ? 0xeu : ((1u << (B - T)) - 1u) & 0xeu));
                T += 4;
            }
            for (; T + 4 <= B; T += 4) gate4(T, 0xfu);
            if (T < B) gate4(T, (1u << (B - T)) - 1u);
        }
    }
    if (VIS_FIRST) {  // the entries not in view, behind the ones in view
        for (uint32_t k = 0; k < wn; ++k) {
            if (w < vcap) out_at(w) = tmp[(size_t)k * NL_BLOCK];
            ++w;
        }
    }
    if (active) nl.count[s - io.first] = (uint16_t)min(w, vcap);
    if (__any_sync(0xffffffffu, w > vcap) && (tid & 31) == 0) nl_no_lists(nl, tab);
    // Rows up to the warp's longest list (whole batches of four) are padded with a sentinel: the tile
    // slot just past the staged candidates, which the walk fills with a position far outside any
    // flock -- so the fast walk needs no per-entry "is this row mine" test.
    const uint32_t wpad = min((__reduce_max_sync(0xffffffffu, min(w, vcap)) + 3u) & ~3u, vcap);
    const uint32_t pad = total > nl.tile_cap ? nl.tile_cap_b : nl.tile_cap;  // the slot just past the walk's tile
    for (uint32_t k = min(w, vcap); k < wpad; ++k) out_at(k) = (uint16_t)pad;
}

struct NlWalkSmem {
    alignas(16) float tx[NL_TILE + 8], ty[NL_TILE + 8], tz[NL_TILE + 8];
    uint16_t list[NL_CAP][NL_BLOCK];  // per-thread survivor lists: tile offsets
    uint32_t toff[10], tslot[9];
    alignas(8) uint64_t bar;
};

// EXACT numerics: 35 KB of shared memory, six CTAs per SM at 80 registers.  With ~17 survivors per
// boid a 48-entry list drains once per boid almost always (measured against 64 entries at five
// CTAs per SM: C4 3.34 against 3.55 ms, profiles/r2_bench_*).
__global__ void __launch_bounds__(NL_BLOCK, 6)
nl_walk_kernel(const DevParams P, const GridDesc g, const WalkIO io, const NlIO nl, unsigned *__restrict__ status) {
    const float4 *__restrict__ vel_s = io.vel_s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NlWalkSmem &S = *reinterpret_cast<NlWalkSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t s = io.first + blockIdx.x * NL_BLOCK + tid;
    const bool active = s < io.last;
    const NlPrologue pro = nl_prologue<true>(io, nl, tid, s, active);
    if (pro.stale) return;  // lazy re-binning: this step is void (nothing staged, nothing written yet)
    const bool no_lists = pro.no_lists != 0u;
    if (tid < 32 && !no_lists) {  // warp 0 gets the tile moving
        if (tid == 0) mbar_init(&S.bar, 1);
        __syncwarp();
        nl_stage(S, tid, pro.ub, pro.ue, io.soa_in[0], io.soa_in[1], io.soa_in[2], NL_TILE);
    }
    const float4 pi4 = pro.pi4, vi4 = pro.vi4;
    uint32_t n_c = pro.n_c;  // cached candidates of this boid
    const uint2 *const vlp = pro.vlp;
    uint2 q0 = pro.q0, q1 = pro.q1;  // (two batches of entries in flight)
    if (io.ctl) track_motion(io.ctl, active, pi4, vi4);
    Self self;
    self.p = self.v = self.vhat = v3zero();
    bool work = false;
    if (active) {
        self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
        const bool ghost = __float_as_uint(vi4.w) != 0u;
        work = !ghost && !P.steering_overrides;
    }
    if (!work) n_c = 0;
    V3 acc = v3zero();
    if (no_lists) {
        // a CTA without lists (tile or list overflow at build time)
        if (work) acc = nl_walk_global(P, g, io, s, self);
        if (active) walk_finish<TAP_STEP>(P, s, pi4, vi4, self, acc, 0u, 0ull, io, status, TapOut{});
        return;
    }
    __syncthreads();  // the barrier is initialised, the layout is in shared memory
    const uint32_t t_self = work ? s - S.tslot[4] : 0xffffu;
    const uint32_t nmax = __reduce_max_sync(0xffffffffu, n_c);
    if (S.toff[9] > 0) mbar_wait(&S.bar, 0);  // (a layout with lists always fits the tile)

    uint16_t *const lst = &S.list[0][tid];  // entry k at lst[k * BLOCK]
    int cnt = 0;
    uint32_t base = 0;  // warp-uniform progress through the cached lists, a multiple of 4
    const float2 kh2 = f2s(P.fov_kh), kl2 = f2s(P.fov_kl);
    const float2 nsx = f2s(-self.p.x), nsy = f2s(-self.p.y), nsz = f2s(-self.p.z);
    const float2 vhx = f2s(self.vhat.x), vhy = f2s(self.vhat.y), vhz = f2s(self.vhat.z);
#pragma unroll 1
    for (;;) {
        int room = NL_CAP - (int)__reduce_max_sync(0xffffffffu, (unsigned)cnt);
        const bool more = base < nmax;
        if (room < 4 || (!more && room < NL_CAP)) {
            drain_list<NL_BLOCK>(P, self, lst, cnt, S.tx, S.ty, S.tz, S.tslot, t_self, vel_s, acc);
            cnt = 0;
            room = NL_CAP;
        }
        if (!more) break;
        const uint32_t end = min(base + ((uint32_t)room & ~3u), (nmax + 3u) & ~3u);
        uint32_t w = (uint32_t)cnt * NL_BLOCK;  // list cursor, in entries
#pragma unroll 1
        for (uint32_t k = base; k < end; k += 4) {
            const uint32_t c[4] = {q0.x & 0xffffu, q0.x >> 16, q0.y & 0xffffu, q0.y >> 16};
            q0 = q1;
            if (k + 8 < nmax) q1 = __ldcs(vlp + (size_t)(k / 4 + 2) * NL_BLOCK);  // (rows < vcap: a multiple of 4)
            // The plain walk's pre-gate (fp_walk.cu) on packed FP32 -- two entries per FADD2 / FMUL2 /
            // FFMA2: fused squared distance against m2_cut_hi, and the conservative FOV test
            // KL m2 < q |q| < KH m2 (drops only pairs culled with a 1e-5 margin; NaN never drops).
            // A superset is all it has to keep: the drain re-tests exactly.
            const uint32_t t0 = c[0] & 0xfffu, t1 = c[1] & 0xfffu, t2 = c[2] & 0xfffu, t3 = c[3] & 0xfffu;
            // (entries past n_c hold stale offsets: any 12-bit offset reads inside the tile arrays,
            //  result unused)
            const float2 dx01 = __fadd2_rn(f2(S.tx[t0], S.tx[t1]), nsx), dx23 = __fadd2_rn(f2(S.tx[t2], S.tx[t3]), nsx);
            const float2 dy01 = __fadd2_rn(f2(S.ty[t0], S.ty[t1]), nsy), dy23 = __fadd2_rn(f2(S.ty[t2], S.ty[t3]), nsy);
            const float2 dz01 = __fadd2_rn(f2(S.tz[t0], S.tz[t1]), nsz), dz23 = __fadd2_rn(f2(S.tz[t2], S.tz[t3]), nsz);
            const float2 m01 = __ffma2_rn(dz01, dz01, __ffma2_rn(dy01, dy01, __fmul2_rn(dx01, dx01)));
            const float2 m23 = __ffma2_rn(dz23, dz23, __ffma2_rn(dy23, dy23, __fmul2_rn(dx23, dx23)));
            const float2 q01 = __ffma2_rn(vhz, dz01, __ffma2_rn(vhy, dy01, __fmul2_rn(vhx, dx01)));
            const float2 q23 = __ffma2_rn(vhz, dz23, __ffma2_rn(vhy, dy23, __fmul2_rn(vhx, dx23)));
            const float2 s01 = __fmul2_rn(q01, f2(fabsf(q01.x), fabsf(q01.y)));
            const float2 s23 = __fmul2_rn(q23, f2(fabsf(q23.x), fabsf(q23.y)));
            const float2 h01 = __fmul2_rn(kh2, m01), h23 = __fmul2_rn(kh2, m23);
            const float2 l01 = __fmul2_rn(kl2, m01), l23 = __fmul2_rn(kl2, m23);
            const float mm[4] = {m01.x, m01.y, m23.x, m23.y};
            const float ss[4] = {s01.x, s01.y, s23.x, s23.y};
            const float hh[4] = {h01.x, h01.y, h23.x, h23.y};
            const float ll[4] = {l01.x, l01.y, l23.x, l23.y};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (k + u < n_c && !(mm[u] >= P.m2_cut_hi) && !(ss[u] < hh[u] && ss[u] > ll[u])) {
                    lst[w] = (uint16_t)c[u];
                    w += NL_BLOCK;
                }
            }
        }
        base = end;
        cnt = (int)(w / NL_BLOCK);
    }
    if (!active) return;
    walk_finish<TAP_STEP>(P, s, pi4, vi4, self, acc, 0u, 0ull, io, status, TapOut{});
}

// ---- FAST numerics ----------------------------------------------------------------------------
// Shared memory of the fast walk, 44 KB either way (five CTAs per SM):
//   form A (the nine intervals hold <= NF_TILE candidates): x, y, z AND velocities staged by TMA;
//   form B (<= NF_TILE_B): positions only, the velocity of a contributing pair comes from global
//          memory (L1 / L2: the CTA's boids share their candidates).  A CTA whose boids straddle
//          two cell columns has a union of intervals ~15 % above the average; with form A alone
//          1.4 % of C4's CTAs overflowed the tile and walked their 27 cells from global memory with
//          the exact pair function -- 9 % of the kernel's time (profiles/r2_c4_nl_fast_v3_*).
constexpr int NF_STRIDE_A = NF_TILE + 8, NF_STRIDE_B = NF_TILE_B + 8;  // floats per coordinate array
constexpr int NF_SMEM_FLOATS = 3 * NF_STRIDE_A + 4 * NF_TILE;          // form A: 44 000 B
static_assert(3 * NF_STRIDE_B <= NF_SMEM_FLOATS, "form B must fit the same buffer");
struct NlFastTail {  // behind the tile
    uint32_t toff[10], tslot[9];
    alignas(8) uint64_t bar;
};


// Two entries of a boid's cached list under FAST numerics, in the two halves of packed FP32
// registers (FADD2 / FMUL2 / FFMA2: one issue slot for both).  The squared distance is the
// reference's own (separately rounded, boid.rs:94-96 through cgmath's dot), so "in range" and
// "weight 1" are its decisions.  The cosine of the sight angle is fused and uses MUFU.RSQ: within
// ~1e-6 of the reference's (boid.rs:102-105); the decision is taken on it when it is further than
// the guard band from both ends of the culled interval [-1, cstar], else `unsure` sends the pair
// down the exact path.
struct FastSelf {
    float px, py, pz;  // position
    float hx, hy, hz;  // direction of flight
    float vx, vy, vz;  // velocity
};
struct FastPair2 {
    float2 dx, dy, dz, m2, r, gc;
};
__device__ __forceinline__ FastPair2 fast_gate2(const DevParams &P, const FastSelf &F, float2 px, float2 py,
                                                float2 pz) {
    FastPair2 f;
    f.dx = __fadd2_rn(px, f2s(-F.px));
    f.dy = __fadd2_rn(py, f2s(-F.py));
    f.dz = __fadd2_rn(pz, f2s(-F.pz));
    // (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 -- the explicit rounding modifiers do
    //  not stop it as they do for scalars -- but never a packed product with a scalar add: the sums
    //  are scalar, each product and each sum separately rounded as in the reference)
    const float2 xx = __fmul2_rn(f.dx, f.dx), yy = __fmul2_rn(f.dy, f.dy), zz = __fmul2_rn(f.dz, f.dz);
    f.m2 = f2(fadd(fadd(xx.x, yy.x), zz.x), fadd(fadd(xx.y, yy.y), zz.y));
    const float2 q = __ffma2_rn(f2s(F.hz), f.dz, __ffma2_rn(f2s(F.hy), f.dy, __fmul2_rn(f2s(F.hx), f.dx)));
    f.r = f2(rsqrt_seed(f.m2.x), rsqrt_seed(f.m2.y));
    const float2 c = __fmul2_rn(q, f.r);
    // <= 0: culled (acosf(c) > max_sight_angle); NaN for coincident positions (0 x inf)
    f.gc = __fmul2_rn(__fadd2_rn(c, f2s(-P.fz_a)), __fadd2_rn(c, f2s(-P.fz_b)));
    return f;
}
// contributions of the two pairs (those that passed), accumulated with FMAs into packed partial sums:
// w_d ((av + ce) + vm)  (boid.rs:162-165)
__device__ __forceinline__ void fast_force2(const DevParams &P, const FastSelf &F, const FastPair2 &f, bool pass_a,
                                            bool pass_b, float4 va, float4 vb, float2 &ax, float2 &ay, float2 &az) {
    // dist: the seed refined to the correctly rounded square root (it enters the ramp by difference)
    const float2 g0 = __fmul2_rn(f.m2, f.r), h = __fmul2_rn(f.r, f2s(0.5f));
    const float2 mag = __ffma2_rn(__ffma2_rn(f2(-g0.x, -g0.y), g0, f.m2), h, g0);
    // ((-f_a / d^2) + f_c d) / d, on d
    const float2 coef = __fmul2_rn(__ffma2_rn(f2s(P.f_c), mag, __fmul2_rn(f2s(P.neg_f_a), __fmul2_rn(f.r, f.r))), f.r);
    const float2 ramp = __fmul2_rn(__fadd2_rn(mag, f2s(-P.thr)), f2s(P.fz_rinv_fall));  // boid.rs:152-161 (F7)
    const float2 w = f2(f.m2.x <= P.m2_one ? 1.0f : ramp.x, f.m2.y <= P.m2_one ? 1.0f : ramp.y);
    const float2 dvx = f2(va.x - F.vx, vb.x - F.vx), dvy = f2(va.y - F.vy, vb.y - F.vy),
                 dvz = f2(va.z - F.vz, vb.z - F.vz);
    // velocity matching unless the velocities agree to EPSILON (boid.rs:132)
    const bool ma = pass_a && !(fmaxf(fmaxf(fabsf(dvx.x), fabsf(dvy.x)), fabsf(dvz.x)) <= FP_F32_EPSILON);
    const bool mb = pass_b && !(fmaxf(fmaxf(fabsf(dvx.y), fabsf(dvy.y)), fabsf(dvz.y)) <= FP_F32_EPSILON);
    float2 cw = __fmul2_rn(coef, w), fw = __fmul2_rn(f2s(P.f_v), w);
    // (a pair that did not pass may hold inf / NaN in coef: selected away, never multiplied away)
    cw = f2(pass_a ? cw.x : 0.0f, pass_b ? cw.y : 0.0f);
    fw = f2(ma ? fw.x : 0.0f, mb ? fw.y : 0.0f);
    ax = __ffma2_rn(cw, f.dx, __ffma2_rn(fw, dvx, ax));
    ay = __ffma2_rn(cw, f.dy, __ffma2_rn(fw, dvy, ay));
    az = __ffma2_rn(cw, f.dz, __ffma2_rn(fw, dvz, az));
}

// The pass over a boid's cached entries.  FORM_B: velocities from global memory.
// Decisions per entry (finite states): in range <=> m2 < m2_cut, the reference's own squared distance;
// contributes <=> gc > tol (visible, outside the guard band); `unsure` <=> |gc| <= tol or NaN (guard
// band; coincident positions).  A batch with a pair closer than 1e-6 (the abs_diff_eq! guards of
// boid.rs:111,121 may apply) is evaluated by the exact sequence altogether.
// (shared-memory loads by 32-bit address: through a generic pointer the compiler rebuilt the shared
//  window base -- S2R CgaCtaId, LEA, an IADD3 per entry -- in every batch)
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ float4 lds_f32x4(uint32_t a) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF));
    return v;
}

template <bool FORM_B>
__device__ __forceinline__ void fast_entries(const DevParams &P, const Self &self, const uint32_t tile_s,
                                             const uint32_t *tslot, const float4 *__restrict__ vel_s,
                                             const uint2 *__restrict__ vlp, uint32_t nmax, uint2 q0, uint2 q1,
                                             float &ax_out, float &ay_out, float &az_out) {
    // tile_s: shared-space address of the tile, obtained AFTER the tile wait (every load below depends on it)
    constexpr int STRIDE = (FORM_B ? NF_STRIDE_B : NF_STRIDE_A) * 4;  // bytes
    constexpr int TV = 3 * NF_STRIDE_A * 4;                           // (form A) velocities
    const FastSelf F{self.p.x, self.p.y, self.p.z, self.vhat.x, self.vhat.y, self.vhat.z, self.v.x, self.v.y, self.v.z};
    // velocity of an entry (form B: only of one that contributes -- a padding entry has no slot to read)
    auto vel_of = [&](uint32_t c, uint32_t a4, bool wanted) -> float4 {  // a4: tile_s + 4 * offset
        if (!FORM_B) return lds_f32x4<TV>(tile_s + 4u * (a4 - tile_s));
        return wanted ? __ldg(vel_s + (tslot[c >> 12] + (c & 0xfffu))) : make_float4(0, 0, 0, 0);
    };
    float2 ax = f2s(0.0f), ay = ax, az = ax;
    const float cut = P.m2_cut, tol = P.fz_gc_tol;
    // one batch of four entries (an 8-byte word of the list)
    auto batch = [&](const uint2 q) {
        // (rows past this lane's list hold the build's padding entry: a candidate 1e18 away)
        const uint32_t a0 = tile_s + ((q.x << 2) & 0x3ffcu), a1 = tile_s + ((q.x >> 14) & 0x3ffcu);
        const uint32_t a2 = tile_s + ((q.y << 2) & 0x3ffcu), a3 = tile_s + ((q.y >> 14) & 0x3ffcu);
        const FastPair2 fa = fast_gate2(P, F, f2(lds_f32<0>(a0), lds_f32<0>(a1)), f2(lds_f32<STRIDE>(a0), lds_f32<STRIDE>(a1)),
                                        f2(lds_f32<2 * STRIDE>(a0), lds_f32<2 * STRIDE>(a1)));
        const FastPair2 fb = fast_gate2(P, F, f2(lds_f32<0>(a2), lds_f32<0>(a3)), f2(lds_f32<STRIDE>(a2), lds_f32<STRIDE>(a3)),
                                        f2(lds_f32<2 * STRIDE>(a2), lds_f32<2 * STRIDE>(a3)));
        // in range <=> m2 < cut (NaN counts as in range and ends up `open`)
        const bool i0 = !(fa.m2.x >= cut), i1 = !(fa.m2.y >= cut), i2 = !(fb.m2.x >= cut), i3 = !(fb.m2.y >= cut);
        // batch-wide: a pair closer than 1e-6, or a cosine inside a guard band / NaN (of ANY entry, in
        // range or not -- the exact evaluation below sorts that out): fmin / fmax drop a NaN operand
        // only next to a number, and a NaN cosine comes from m2 = 0, which `tiny` sees
        const bool tiny = !(fminf(fminf(fa.m2.x, fa.m2.y), fminf(fb.m2.x, fb.m2.y)) >= 1e-12f);
        const bool open = tiny || !(fminf(fminf(fabsf(fa.gc.x), fabsf(fa.gc.y)), fminf(fabsf(fb.gc.x), fabsf(fb.gc.y))) > tol);
        if (!open) {
            const bool p0 = i0 && fa.gc.x > tol, p1 = i1 && fa.gc.y > tol, p2 = i2 && fb.gc.x > tol, p3 = i3 && fb.gc.y > tol;
            if (p0 || p1) fast_force2(P, F, fa, p0, p1, vel_of(q.x & 0xffffu, a0, p0), vel_of(q.x >> 16, a1, p1), ax, ay, az);
            if (p2 || p3) fast_force2(P, F, fb, p2, p3, vel_of(q.y & 0xffffu, a2, p2), vel_of(q.y >> 16, a3, p3), ax, ay, az);
        } else {
            // rare: every in-range entry of the batch goes down the reference's own sequence, which
            // decides (field of view) and evaluates
#pragma unroll 1
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t cu = u == 0 ? (q.x & 0xffffu) : u == 1 ? (q.x >> 16) : u == 2 ? (q.y & 0xffffu) : (q.y >> 16);
                const uint32_t au = tile_s + 4u * (cu & 0xfffu);
                const float dx = fsub(lds_f32<0>(au), self.p.x), dy = fsub(lds_f32<STRIDE>(au), self.p.y),
                            dz = fsub(lds_f32<2 * STRIDE>(au), self.p.z);
                const float m2 = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
                if (m2 >= cut) continue;
                const float4 vj = vel_of(cu, au, true);
                V3 contrib;
                if (pair_inrange<false>(P, self, v3(dx, dy, dz), m2, v3(vj.x, vj.y, vj.z), 1.0f, P.cstar, contrib)) {
                    ax.x += contrib.x;
                    ay.x += contrib.y;
                    az.x += contrib.z;
                }
            }
        }
    };
    // four words of entries in flight; the loop is unrolled by four so that they need no rotation
    uint2 q2 = make_uint2(0, 0), q3 = q2;
    if (8 < nmax) q2 = __ldcs(vlp + 2 * NL_BLOCK);
    if (12 < nmax) q3 = __ldcs(vlp + 3 * NL_BLOCK);
    const uint2 *nxt = vlp + 4 * NL_BLOCK;  // word of entries k + 16 ..
#pragma unroll 1
    for (uint32_t k = 0; k < nmax; k += 16, nxt += 4 * NL_BLOCK) {
        const uint2 w0 = q0, w1 = q1, w2 = q2, w3 = q3;
        if (k + 16 < nmax) q0 = __ldcs(nxt);
        if (k + 20 < nmax) q1 = __ldcs(nxt + NL_BLOCK);
        if (k + 24 < nmax) q2 = __ldcs(nxt + 2 * NL_BLOCK);
        if (k + 28 < nmax) q3 = __ldcs(nxt + 3 * NL_BLOCK);
        batch(w0);
        if (k + 4 < nmax) batch(w1);
        if (k + 8 < nmax) batch(w2);
        if (k + 12 < nmax) batch(w3);
    }
    ax_out = ax.x + ax.y;
    ay_out = ay.x + ay.y;
    az_out = az.x + az.y;
}

template <int TAP>
__global__ void __launch_bounds__(NL_BLOCK, 5)
nl_fast_kernel(const DevParams P, const GridDesc g, const WalkIO io, const NlIO nl, unsigned *__restrict__ status,
               TapOut tap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *const tile = reinterpret_cast<float *>(smem_raw);
    NlFastTail &S = *reinterpret_cast<NlFastTail *>(smem_raw + sizeof(float) * NF_SMEM_FLOATS);
    const uint32_t tid = threadIdx.x;
    const uint32_t s = io.first + blockIdx.x * NL_BLOCK + tid;
    const bool active = s < io.last;
    const NlPrologue pro = nl_prologue<TAP == TAP_STEP>(io, nl, tid, s, active);
    const uint32_t no_lists = pro.no_lists, ub = pro.ub, ue = pro.ue;
    const float4 pi4 = pro.pi4, vi4 = pro.vi4;
    const uint32_t n_c = pro.n_c;
    if (pro.stale) return;  // lazy re-binning: this step is void (nothing staged, nothing written yet)

    bool form_b = false;
    if (tid < 32 && !no_lists) {  // warp 0 gets the tile moving
        if (tid == 0) mbar_init(&S.bar, 1);
        __syncwarp();
        const uint32_t len = ue - ub;
        uint32_t inc = len;  // inclusive prefix sum over the lanes
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
            if ((int)tid >= off) inc += t;
        }
        const uint32_t toff = inc - len, total = __shfl_sync(0xffffffffu, inc, 8);
        form_b = total > (uint32_t)NF_TILE;  // (a layout with lists fits form B: the build checked)
        if (tid < 9) {
            S.toff[tid] = toff;
            S.tslot[tid] = ub - toff;
        }
        if (tid == 0) {
            S.toff[9] = total;
            if (total) mbar_expect_tx(&S.bar, total * (form_b ? 12u : 28u));
        }
        __syncwarp();
        if (tid < 9 && len) {
            const int stride = form_b ? NF_STRIDE_B : NF_STRIDE_A;
            bulk_g2s(tile + toff, io.soa_in[0] + ub, len * 4u, &S.bar);
            bulk_g2s(tile + stride + toff, io.soa_in[1] + ub, len * 4u, &S.bar);
            bulk_g2s(tile + 2 * stride + toff, io.soa_in[2] + ub, len * 4u, &S.bar);
            if (!form_b) bulk_g2s(reinterpret_cast<float4 *>(tile + 3 * NF_STRIDE_A) + toff, io.vel_s + ub, len * 16u, &S.bar);
        }
        // the padding entries' "candidate": the slot just past the form's tile
        if (tid == 9) {
            const int stride = form_b ? NF_STRIDE_B : NF_STRIDE_A;
            tile[stride - 8] = tile[2 * stride - 8] = tile[3 * stride - 8] = 1e18f;
        }
    }
    if (TAP == TAP_STEP && io.ctl) track_motion(io.ctl, active, pi4, vi4);
    Self self;
    self.p = self.v = self.vhat = v3zero();
    bool work = false;
    if (active) {
        self = make_self(v3(pi4.x, pi4.y, pi4.z), v3(vi4.x, vi4.y, vi4.z));
        const bool ghost = __float_as_uint(vi4.w) != 0u;
        work = !ghost && ((TAP != TAP_STEP) || !P.steering_overrides);
    }
    if (no_lists) {  // a CTA without lists
        V3 acc = v3zero();
        if (work) acc = nl_walk_global(P, g, io, s, self);
        if (active) walk_finish<TAP, true>(P, s, pi4, vi4, self, acc, 0u, 0ull, io, status, tap);
        return;
    }
    __syncthreads();  // the barrier is initialised, the layout is in shared memory
    const uint32_t nmax = __reduce_max_sync(0xffffffffu, n_c);
    const uint32_t total = S.toff[9];
    if (total > 0) mbar_wait(&S.bar, 0);
    uint32_t tile_s;  // (volatile: stays behind the wait, and every tile load depends on it)
    asm volatile("mov.u32 %0, %1;" : "=r"(tile_s) : "r"(smem_u32(smem_raw)) : "memory");
    float ax, ay, az;
    if (total > (uint32_t)NF_TILE)
        fast_entries<true>(P, self, tile_s, S.tslot, io.vel_s, pro.vlp, nmax, pro.q0, pro.q1, ax, ay, az);
    else
        fast_entries<false>(P, self, tile_s, S.tslot, io.vel_s, pro.vlp, nmax, pro.q0, pro.q1, ax, ay, az);
    if (!active) return;
    if (!work) ax = ay = az = 0.0f;  // (steering overrides, ghost record: the lists were walked for nothing)
    walk_finish<TAP, true>(P, s, pi4, vi4, self, v3(ax, ay, az), 0u, 0ull, io, status, tap);
}

}  // namespace

size_t nl_entries_elems(uint32_t rows, uint32_t vcap) {
    const size_t ctas = ((size_t)rows + NL_BLOCK - 1) / NL_BLOCK;
    return (ctas * vcap + 8) * NL_BLOCK;  // + slack rows: the walk's first two batches are loaded unconditionally
}
size_t nl_cta_tab_elems(uint32_t rows) {
    return (((size_t)rows + NL_BLOCK - 1) / NL_BLOCK) * NL_CTA_WORDS;
}
uint32_t nl_tile_cap(bool fast) { return fast ? NF_TILE : NL_TILE; }
uint32_t nl_tile_cap_b(bool fast) { return fast ? NF_TILE_B : 0; }
int launch_nl_build(cudaStream_t st, const GridDesc &g, const WalkIO &io, const NlIO &nl) {
    if (io.last <= io.first) return FP_OK;
    const uint32_t ctas = (io.last - io.first + NL_BLOCK - 1) / NL_BLOCK;
    if (nl.tile_cap_b > (uint32_t)NF_TILE_B || (!nl.tile_cap_b && nl.tile_cap > (uint32_t)NL_TILE)) {
        set_error("internal: candidate-list tile capacity");
        return FP_ERR_INVALID;
    }
    if (nl.vis_first) {
        const int smem = (int)(sizeof(NlBuildSmem<NF_TILE_B>) + sizeof(uint16_t) * NB_TMP * NL_BLOCK);
        FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<true, NF_TILE_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        nl_build_kernel<true, NF_TILE_B><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
    } else if (nl.tile_cap_b) {
        const int smem = (int)sizeof(NlBuildSmem<NF_TILE_B>);
        FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<false, NF_TILE_B>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        nl_build_kernel<false, NF_TILE_B><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
    } else {
        const int smem = (int)sizeof(NlBuildSmem<NL_TILE>);
        FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<false, NL_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        nl_build_kernel<false, NL_TILE><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
    }
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

int launch_nl_walk(cudaStream_t st, const DevParams &P, const GridDesc &g, int tap, const WalkIO &io, const NlIO &nl,
                   unsigned *status, const TapOut &tap_out) {
    if (io.last <= io.first) return FP_OK;
    const uint32_t ctas = (io.last - io.first + NL_BLOCK - 1) / NL_BLOCK;
    if (P.numerics_fast) {
        const int smem = (int)(sizeof(float) * NF_SMEM_FLOATS + sizeof(NlFastTail));
        if (tap == TAP_STEP) {
            FP_CUDA(cudaFuncSetAttribute(nl_fast_kernel<TAP_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            nl_fast_kernel<TAP_STEP><<<ctas, NL_BLOCK, smem, st>>>(P, g, io, nl, status, tap_out);
        } else if (tap == TAP_ACCEL) {
            FP_CUDA(cudaFuncSetAttribute(nl_fast_kernel<TAP_ACCEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            nl_fast_kernel<TAP_ACCEL><<<ctas, NL_BLOCK, smem, st>>>(P, g, io, nl, status, tap_out);
        } else {
            set_error("internal: the list walk serves steps and the acceleration tap only");
            return FP_ERR_INVALID;
        }
    } else {
        if (tap != TAP_STEP) {
            set_error("internal: the exact list walk serves steps only");
            return FP_ERR_INVALID;
        }
        const int smem = (int)sizeof(NlWalkSmem);
        FP_CUDA(cudaFuncSetAttribute(nl_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        nl_walk_kernel<<<ctas, NL_BLOCK, smem, st>>>(P, g, io, nl, status);
    }
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

}  // namespace fp

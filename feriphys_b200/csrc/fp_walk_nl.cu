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
constexpr int NF_TILE = 1568;   // staged candidates per CTA of the fast walk (positions + velocities)
constexpr int NL_CTA_WORDS = 20;  // cached tile layout per CTA: ub[9], ue[9], [18] = no lists, 1 spare

// this CTA gets no lists: it walks from global memory every step (counted once per CTA)
__device__ __forceinline__ void nl_no_lists(const NlIO &nl, uint32_t *tab) {
    if (atomicExch(tab + 18, 1u) == 0u) atomicAdd(nl.flag, 1u);
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

struct NlBuildSmem {
    alignas(16) float tx[NL_TILE + 8], ty[NL_TILE + 8], tz[NL_TILE + 8];
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
template <bool VIS_FIRST>
__global__ void __launch_bounds__(NL_BLOCK)
nl_build_kernel(const GridDesc g, const WalkIO io, const NlIO nl) {
    // (No look at ctl->stale: the lists describe the binning, which stands whether or not the
    // step they are built in turns out void.)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NlBuildSmem &S = *reinterpret_cast<NlBuildSmem *>(smem_raw);
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
        nl_stage(S, tid, ub, ue, io.soa_in[0], io.soa_in[1], io.soa_in[2], nl.tile_cap);
    }
    __syncthreads();
    const uint32_t total = S.toff[9];
    if (total > nl.tile_cap) {  // dense cluster: the tile does not fit -- no lists for this CTA
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
    uint16_t *const tmp = reinterpret_cast<uint16_t *>(smem_raw + sizeof(NlBuildSmem)) + tid;  // VIS_FIRST: [NB_TMP][BLOCK]
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
    for (uint32_t k = min(w, vcap); k < wpad; ++k) out_at(k) = (uint16_t)nl.tile_cap;
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
    if (io.ctl && io.ctl->stale) return;  // lazy re-binning: this step is void
    const float4 *__restrict__ vel_s = io.vel_s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NlWalkSmem &S = *reinterpret_cast<NlWalkSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t s = io.first + blockIdx.x * NL_BLOCK + tid;
    const bool active = s < io.last;
    const uint32_t *const tab = nl.cta_tab + (size_t)blockIdx.x * NL_CTA_WORDS;
    const bool no_lists = __ldg(tab + 18) != 0u;
    if (tid < 32 && !no_lists) {  // warp 0 gets the tile moving before anything else
        if (tid == 0) mbar_init(&S.bar, 1);
        __syncwarp();
        const uint32_t ub = tid < 9 ? __ldg(tab + tid) : 0u, ue = tid < 9 ? __ldg(tab + 9 + tid) : 0u;
        nl_stage(S, tid, ub, ue, io.soa_in[0], io.soa_in[1], io.soa_in[2], NL_TILE);
    }

    float4 pi4 = make_float4(0, 0, 0, 0), vi4 = make_float4(0, 0, 0, 0);
    uint32_t n_c = 0;  // cached candidates of this boid
    if (active) {
        pi4 = io.pos_s[s];
        vi4 = vel_s[s];
        n_c = __ldg(nl.count + (s - io.first));
    }
    // this CTA's list block and its first two batches of entries, in flight while the tile is staged
    // (four entries per 8-byte word; two batches in flight)
    const uint2 *const vlp = reinterpret_cast<const uint2 *>(nl.entries) + (size_t)blockIdx.x * (nl.vcap / 4) * NL_BLOCK + tid;
    uint2 q0 = __ldcs(vlp), q1 = __ldcs(vlp + NL_BLOCK);
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
    const float kh = P.fov_kh, kl = P.fov_kl;
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
            // The plain walk's pre-gate (fp_walk.cu), one candidate per lane-slot: fused squared
            // distance against m2_cut_hi, and the conservative FOV test KL m2 < q |q| < KH m2
            // (drops only pairs culled with a 1e-5 margin; NaN never drops).
            float mm[4], ss[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t t = c[u] & 0xfffu;  // (entries past n_c hold stale offsets: any 12-bit
                                                   //  offset reads inside the tile arrays, result unused)
                const float dx = S.tx[t] - self.p.x, dy = S.ty[t] - self.p.y, dz = S.tz[t] - self.p.z;
                mm[u] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const float q = fmaf(self.vhat.z, dz, fmaf(self.vhat.y, dy, self.vhat.x * dx));
                ss[u] = q * fabsf(q);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float hh = kh * mm[u], ll = kl * mm[u];
                if (k + u < n_c && !(mm[u] >= P.m2_cut_hi) && !(ss[u] < hh && ss[u] > ll)) {
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
struct NlFastSmem {
    alignas(16) float tx[NF_TILE + 8], ty[NF_TILE + 8], tz[NF_TILE + 8];
    alignas(16) float4 tv[NF_TILE];  // velocities of the staged candidates (.w: record flag, unused)
    uint32_t toff[10], tslot[9];
    alignas(8) uint64_t bar;
};

// One entry of a boid's cached list under FAST numerics.  The squared distance is the reference's
// own (separately rounded, boid.rs:94-96 through cgmath's dot), so "in range" and "weight 1" are its
// decisions.  The cosine of the sight angle is fused and uses MUFU.RSQ: within ~1e-6 of the
// reference's (boid.rs:102-105); the decision is taken on it when it is further than the guard band
// from both ends of the culled interval [-1, cstar], else `unsure` sends the pair down the exact path.
struct FastPair {
    float dx, dy, dz, m2, r;
    bool pass;    // contributes, decided outside the guard bands
    bool unsure;  // in range but degenerate or inside a guard band: the exact sequence decides
};
__device__ __forceinline__ FastPair fast_gate(const DevParams &P, const Self &self, float px, float py, float pz) {
    FastPair f;
    f.dx = fsub(px, self.p.x);
    f.dy = fsub(py, self.p.y);
    f.dz = fsub(pz, self.p.z);
    f.m2 = fadd(fadd(fmul(f.dx, f.dx), fmul(f.dy, f.dy)), fmul(f.dz, f.dz));
    const float q = fmaf(self.vhat.z, f.dz, fmaf(self.vhat.y, f.dy, self.vhat.x * f.dx));
    f.r = rsqrt_seed(f.m2);
    const float c = q * f.r;
    const float gc = (c - P.fz_a) * (c - P.fz_b);   // <= 0: culled (acosf(c) > max_sight_angle)
    const bool in = !(f.m2 >= P.m2_cut);  // (a padding entry is 1e18 away)
    // coincident positions (abs_diff_eq! guards, boid.rs:111,121), NaN, and cosines in the guard band
    const bool clear = fabsf(gc) > P.fz_gc_tol && f.m2 >= 1e-12f;
    f.unsure = in && !clear;
    f.pass = in && clear && gc > 0.0f;
    return f;
}
// contribution of a pair that passed, accumulated with FMAs: w_d ((av + ce) + vm)  (boid.rs:162-165)
__device__ __forceinline__ void fast_force(const DevParams &P, const Self &self, const FastPair &f, float4 vj,
                                           float &ax, float &ay, float &az) {
    // dist: the seed refined to the correctly rounded square root (it enters the ramp by difference)
    const float g0 = f.m2 * f.r, h = 0.5f * f.r;
    const float mag = fmaf(fmaf(-g0, g0, f.m2), h, g0);
    const float coef = fmaf(P.f_c, mag, P.neg_f_a * (f.r * f.r)) * f.r;  // ((-f_a / d^2) + f_c d) / d, on d
    const float w = f.m2 <= P.m2_one ? 1.0f : (mag - P.thr) * P.fz_rinv_fall;  // boid.rs:152-161 (F7)
    const float dvx = vj.x - self.v.x, dvy = vj.y - self.v.y, dvz = vj.z - self.v.z;
    const bool vsm = fmaxf(fmaxf(fabsf(dvx), fabsf(dvy)), fabsf(dvz)) <= FP_F32_EPSILON;  // boid.rs:132
    const float cw = coef * w, fw = vsm ? 0.0f : P.f_v * w;
    ax = fmaf(cw, f.dx, fmaf(fw, dvx, ax));
    ay = fmaf(cw, f.dy, fmaf(fw, dvy, ay));
    az = fmaf(cw, f.dz, fmaf(fw, dvz, az));
}

template <int TAP>
__global__ void __launch_bounds__(NL_BLOCK, 5)
nl_fast_kernel(const DevParams P, const GridDesc g, const WalkIO io, const NlIO nl, unsigned *__restrict__ status,
               TapOut tap) {
    if (TAP == TAP_STEP && io.ctl && io.ctl->stale) return;  // lazy re-binning: this step is void
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NlFastSmem &S = *reinterpret_cast<NlFastSmem *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t s = io.first + blockIdx.x * NL_BLOCK + tid;
    const bool active = s < io.last;
    const uint32_t *const tab = nl.cta_tab + (size_t)blockIdx.x * NL_CTA_WORDS;
    const bool no_lists = __ldg(tab + 18) != 0u;
    if (tid < 32 && !no_lists) {  // warp 0 gets the tile moving before anything else
        if (tid == 0) mbar_init(&S.bar, 1);
        __syncwarp();
        const uint32_t ub = tid < 9 ? __ldg(tab + tid) : 0u, ue = tid < 9 ? __ldg(tab + 9 + tid) : 0u;
        nl_stage(S, tid, ub, ue, io.soa_in[0], io.soa_in[1], io.soa_in[2], NF_TILE, S.tv, io.vel_s);
    }
    if (tid == 32) S.tx[NF_TILE] = S.ty[NF_TILE] = S.tz[NF_TILE] = 1e18f;  // the padding entries' "candidate"

    float4 pi4 = make_float4(0, 0, 0, 0), vi4 = make_float4(0, 0, 0, 0);
    uint32_t n_c = 0;
    if (active) {
        pi4 = io.pos_s[s];
        vi4 = io.vel_s[s];
        n_c = __ldg(nl.count + (s - io.first));
    }
    // (four entries per 8-byte word; two batches in flight)
    const uint2 *const vlp = reinterpret_cast<const uint2 *>(nl.entries) + (size_t)blockIdx.x * (nl.vcap / 4) * NL_BLOCK + tid;
    uint2 q0 = __ldcs(vlp), q1 = __ldcs(vlp + NL_BLOCK);
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
    if (S.toff[9] > 0) mbar_wait(&S.bar, 0);

    float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll 1
    for (uint32_t k = 0; k < nmax; k += 4) {
        const uint32_t c[4] = {q0.x & 0xffffu, q0.x >> 16, q0.y & 0xffffu, q0.y >> 16};
        q0 = q1;
        if (k + 8 < nmax) q1 = __ldcs(vlp + (size_t)(k / 4 + 2) * NL_BLOCK);
        FastPair f[4];
        uint32_t t[4];
        bool any_unsure = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            t[u] = c[u] & 0xfffu;  // (rows past this lane's list hold the build's padding entry)
            f[u] = fast_gate(P, self, S.tx[t[u]], S.ty[t[u]], S.tz[t[u]]);
            any_unsure |= f[u].unsure;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (f[u].pass) fast_force(P, self, f[u], S.tv[t[u]], ax, ay, az);
        if (any_unsure) {
            // guard band / degenerate pair (rare): the reference's own sequence decides and evaluates
#pragma unroll 1
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t tu = (u == 0 ? c[0] : u == 1 ? c[1] : u == 2 ? c[2] : c[3]) & 0xfffu;
                const FastPair fu = fast_gate(P, self, S.tx[tu], S.ty[tu], S.tz[tu]);
                if (!fu.unsure) continue;
                const float4 vj = S.tv[tu];
                V3 contrib;
                if (pair_inrange<false>(P, self, v3(fu.dx, fu.dy, fu.dz), fu.m2, v3(vj.x, vj.y, vj.z), 1.0f, P.cstar,
                                        contrib)) {
                    ax += contrib.x;
                    ay += contrib.y;
                    az += contrib.z;
                }
            }
        }
    }
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

int launch_nl_build(cudaStream_t st, const GridDesc &g, const WalkIO &io, const NlIO &nl) {
    if (io.last <= io.first) return FP_OK;
    const uint32_t ctas = (io.last - io.first + NL_BLOCK - 1) / NL_BLOCK;
    if (nl.vis_first) {
        const int smem = (int)(sizeof(NlBuildSmem) + sizeof(uint16_t) * NB_TMP * NL_BLOCK);
        FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        nl_build_kernel<true><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
    } else {
        const int smem = (int)sizeof(NlBuildSmem);
        FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        nl_build_kernel<false><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
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
        const int smem = (int)sizeof(NlFastSmem);
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

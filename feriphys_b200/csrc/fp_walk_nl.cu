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
constexpr int NF_TILE_B = 3648; // ... form B (positions only)
constexpr int NL_CTA_WORDS = 20;  // cached tile layout per CTA: ub[9], ue[9], [18] = no lists, 1 spare

// this CTA gets no lists: it walks from global memory every step (counted once per CTA)
__device__ __forceinline__ void nl_no_lists(const NlIO &nl, uint32_t *tab) {
    if (atomicExch(tab + 18, 1u) == 0u) atomicAdd(nl.flag, 1u);
}

__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
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
// one of them -- the step's validity, the CTA's layout record, the boid, its list length, its first
// two words of entries -- and, into L2 only, the same items of the CTA that will take this one's
// place `nl.ahead` blocks on: its prologue then waits for L2, not for HBM (the start-up chain
// layout -> TMA -> tile was a third of the kernel's stall samples, profiles/r2_c4_nl_fast_v3_*).
struct NlPrologue {
    uint32_t stale, no_lists, ub, ue, n_c;
    float4 pi4, vi4;
    const uint2 *vlp;
    uint2 q0, q1;
    bool pf;            // this CTA warms L2 for a successor
    uint32_t ub2, ue2;  // the successor's intervals (threads 32 .. 40)
};
template <bool STEP, bool EARLY2>  // EARLY2: the successor's intervals are loaded here (two registers), else later
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
    p.pf = nl.ahead && blockIdx.x + nl.ahead < gridDim.x;
    p.ub2 = p.ue2 = 0;
    if (p.pf) {
        const uint32_t s2 = s + nl.ahead * NL_BLOCK;
        if (s2 < io.last) {
            prefetch_l2(io.pos_s + s2);
            prefetch_l2(io.vel_s + s2);
        }
        prefetch_l2(p.vlp + (size_t)nl.ahead * (nl.vcap / 4) * NL_BLOCK);
        prefetch_l2(p.vlp + (size_t)nl.ahead * (nl.vcap / 4) * NL_BLOCK + NL_BLOCK);
        if (EARLY2) {
            if (tid >= 32 && tid < 41) {
                p.ub2 = ld_nc_u32(tab + (size_t)nl.ahead * NL_CTA_WORDS + (tid - 32));
                p.ue2 = ld_nc_u32(tab + (size_t)nl.ahead * NL_CTA_WORDS + 9 + (tid - 32));
            }
        } else if (tid == 32) {
            prefetch_l2(tab + (size_t)nl.ahead * NL_CTA_WORDS);
        }
    }
    return p;
}
// the successor's tile, into L2 (threads 32 .. 40, once this CTA's own tile has landed)
template <bool EARLY2>
__device__ __forceinline__ void nl_warm_successor(const NlPrologue &p, const WalkIO &io, const NlIO &nl, uint32_t tid,
                                                  bool vel) {
    if (p.pf && tid >= 32 && tid < 41) {
        uint32_t ub2 = p.ub2, ue2 = p.ue2;
        if (!EARLY2) {
            const uint32_t *const tab2 = nl.cta_tab + ((size_t)blockIdx.x + nl.ahead) * NL_CTA_WORDS;
            ub2 = ld_nc_u32(tab2 + (tid - 32));
            ue2 = ld_nc_u32(tab2 + 9 + (tid - 32));
        }
        if (ue2 > ub2) {
            const uint32_t bytes = (ue2 - ub2) * 4u;
            bulk_prefetch_l2(io.soa_in[0] + ub2, bytes);
            bulk_prefetch_l2(io.soa_in[1] + ub2, bytes);
            bulk_prefetch_l2(io.soa_in[2] + ub2, bytes);
            if (vel) bulk_prefetch_l2(io.vel_s + ub2, bytes * 4u);
        }
    }
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
    const NlPrologue pro = nl_prologue<true, false>(io, nl, tid, s, active);
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
    nl_warm_successor<false>(pro, io, nl, tid, false);

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
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
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
template <bool FORM_B>
__device__ __forceinline__ void fast_entries(const DevParams &P, const Self &self, const uint32_t *tslot,
                                             const float4 *__restrict__ vel_s, const uint2 *__restrict__ vlp,
                                             uint32_t nmax, uint2 q0, uint2 q1, float &ax_out, float &ay_out,
                                             float &az_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int STRIDE = FORM_B ? NF_STRIDE_B : NF_STRIDE_A;
    const float *const tx = reinterpret_cast<const float *>(smem_raw);
    const float *const ty = tx + STRIDE, *const tz = tx + 2 * STRIDE;
    const float4 *const tv = reinterpret_cast<const float4 *>(tx + 3 * NF_STRIDE_A);  // (form A)
    const FastSelf F{self.p.x, self.p.y, self.p.z, self.vhat.x, self.vhat.y, self.vhat.z, self.v.x, self.v.y, self.v.z};
    // velocity of an entry that contributes (form B: a padding entry has no slot to read)
    auto vel_of = [&](uint32_t c, bool wanted) -> float4 {
        const uint32_t t = c & 0xfffu;
        if (!FORM_B) return tv[t];
        return wanted ? __ldg(vel_s + (tslot[c >> 12] + t)) : make_float4(0, 0, 0, 0);
    };
    float2 ax = f2s(0.0f), ay = ax, az = ax;
    const float cut = P.m2_cut, tol = P.fz_gc_tol;
    // one batch of four entries (an 8-byte word of the list)
    auto batch = [&](const uint2 q) {
        const uint32_t c0 = q.x & 0xffffu, c1 = q.x >> 16, c2 = q.y & 0xffffu, c3 = q.y >> 16;
        // (rows past this lane's list hold the build's padding entry: a candidate 1e18 away)
        const uint32_t t0 = c0 & 0xfffu, t1 = c1 & 0xfffu, t2 = c2 & 0xfffu, t3 = c3 & 0xfffu;
        const FastPair2 fa = fast_gate2(P, F, f2(tx[t0], tx[t1]), f2(ty[t0], ty[t1]), f2(tz[t0], tz[t1]));
        const FastPair2 fb = fast_gate2(P, F, f2(tx[t2], tx[t3]), f2(ty[t2], ty[t3]), f2(tz[t2], tz[t3]));
        const bool tiny = !(fminf(fminf(fa.m2.x, fa.m2.y), fminf(fb.m2.x, fb.m2.y)) >= 1e-12f);
        const bool i0 = !(fa.m2.x >= cut) && !tiny, i1 = !(fa.m2.y >= cut) && !tiny;  // (NaN: in, then unsure)
        const bool i2 = !(fb.m2.x >= cut) && !tiny, i3 = !(fb.m2.y >= cut) && !tiny;
        const bool p0 = i0 && fa.gc.x > tol, p1 = i1 && fa.gc.y > tol, p2 = i2 && fb.gc.x > tol, p3 = i3 && fb.gc.y > tol;
        const bool u0 = i0 && !(fabsf(fa.gc.x) > tol), u1 = i1 && !(fabsf(fa.gc.y) > tol);
        const bool u2 = i2 && !(fabsf(fb.gc.x) > tol), u3 = i3 && !(fabsf(fb.gc.y) > tol);
        if (p0 || p1) fast_force2(P, F, fa, p0, p1, vel_of(c0, p0), vel_of(c1, p1), ax, ay, az);
        if (p2 || p3) fast_force2(P, F, fb, p2, p3, vel_of(c2, p2), vel_of(c3, p3), ax, ay, az);
        if (tiny || u0 || u1 || u2 || u3) {
            // guard band / degenerate pair (rare): the reference's own sequence decides and evaluates
            // (which entries: the flags taken above, not a second evaluation of the cosine -- the
            //  compiler may fuse the two differently, and an entry must go exactly one way)
            const uint32_t open_mask = tiny ? 0xfu : (u0 ? 1u : 0u) | (u1 ? 2u : 0u) | (u2 ? 4u : 0u) | (u3 ? 8u : 0u);
#pragma unroll 1
            for (uint32_t u = 0; u < 4; ++u) {
                if (!(open_mask >> u & 1u)) continue;
                const uint32_t cu = u == 0 ? c0 : u == 1 ? c1 : u == 2 ? c2 : c3;
                const uint32_t tu = cu & 0xfffu;
                const float dx = fsub(tx[tu], self.p.x), dy = fsub(ty[tu], self.p.y), dz = fsub(tz[tu], self.p.z);
                const float m2 = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));  // (the same bits as above)
                if (m2 >= cut) continue;
                const float4 vj = vel_of(cu, true);
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
    const NlPrologue pro = nl_prologue<TAP == TAP_STEP, true>(io, nl, tid, s, active);
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
    nl_warm_successor<true>(pro, io, nl, tid, true);
    float ax, ay, az;
    if (total > (uint32_t)NF_TILE)
        fast_entries<true>(P, self, S.tslot, io.vel_s, pro.vlp, nmax, pro.q0, pro.q1, ax, ay, az);
    else
        fast_entries<false>(P, self, S.tslot, io.vel_s, pro.vlp, nmax, pro.q0, pro.q1, ax, ay, az);
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
// CTAs of a list walk that are resident on the current device at once (0 if that cannot be told)
uint32_t nl_walk_resident_ctas(bool fast) {
    static uint32_t cached[2] = {~0u, ~0u};
    uint32_t &c = cached[fast ? 1 : 0];
    if (c == ~0u) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) {
            if (fast) {
                const int smem = (int)(sizeof(float) * NF_SMEM_FLOATS + sizeof(NlFastTail));
                e = cudaFuncSetAttribute(nl_fast_kernel<TAP_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                if (e == cudaSuccess)
                    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nl_fast_kernel<TAP_STEP>, NL_BLOCK, smem);
            } else {
                const int smem = (int)sizeof(NlWalkSmem);
                e = cudaFuncSetAttribute(nl_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nl_walk_kernel, NL_BLOCK, smem);
            }
        }
        if (e != cudaSuccess) (void)cudaGetLastError();
        c = e == cudaSuccess ? (uint32_t)(sms * per_sm) : 0u;
    }
    return c;
}

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

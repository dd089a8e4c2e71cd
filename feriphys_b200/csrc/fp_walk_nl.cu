// fp_walk_nl.cu -- K3 with standing candidate lists (EXPERIMENTAL, TAP_STEP only, not the
// default: FP_WALK_VARIANT=41 .. 46, see nl_prepare in fp_api.cu).  The plain form (41) has run on
// a B200: bit-identical to the production walk in every run compared, C4 5.66 -> 3.85 ms/step
// (profiles/r1_nl_*.log); the other forms are written but have not run yet.
//
// With lazy re-binning a binning stands for ~20 steps, yet the production
// walk (fp_walk.cu) pre-gates all ~160 candidates of a boid's 27 home cells in every one of
// them -- 48 % of its instructions (profiles/r1_c4_walk_hotspots.txt) -- to find the ~34
// that are within reach.  Here that work is done ONCE per binning:
//
//   nl_build_kernel  (right after a binning, positions = the binned positions)
//       stages the same nine intervals as the production walk, keeps every candidate whose
//       squared distance is below (reach + skin)^2 (1 + 1e-5), and writes the survivors' tile
//       offsets (16 bit, row << 12 | offset, ascending slot order) to a per-thread list in
//       global memory, [cta][entry][thread] so that entry k of a warp is one 64-byte run;
//       also the per-boid count and the CTA's tile layout (nine intervals).
//   nl_walk_kernel   (every step until the next binning)
//       stages the tile from the cached layout (no cell-table look-ups, no reductions), runs
//       the production pre-gate -- fused squared distance against m2_cut_hi and the
//       conservative FOV test -- over the ~34 cached entries instead of ~160 candidates, and
//       drains the survivors exactly as the production kernel does.
//
// Exactness.  While the binning stands every boid is within skin / 2 of where it was binned
// (the device-checked displacement bound D), so a pair closer than reach now was closer than
// reach + skin then: the cached list is a superset of every pair the production pre-gate
// keeps, in the same order.  The survivor list the drain sees is therefore the same list,
// and the result is bit-identical to the production kernel's (and to the oracle's on the
// standing listing).
//
// Capacity.  A list holds nl.vcap entries.  A CTA in which some boid has more, or whose nine
// intervals do not fit the tile, is marked in its layout record and walks the 27 cells from
// global memory every step (the production kernel's own path for CTAs that overflow the tile):
// slower, same result, no effect on any other CTA.  The build counts such CTAs; the host turns
// the lists off when they stop being rare.
#include "fp_walk_stage.cuh"

namespace fp {

namespace {

constexpr int NL_BLOCK = 128;   // threads (= boids) per CTA, as the production walk
constexpr int NL_TILE = 1904;   // staged candidates per CTA, as the production walk
constexpr int NL_CAP = 64;      // survivor list (shared memory), as the production walk
constexpr int NL_CTA_WORDS = 20;  // cached tile layout per CTA: ub[9], ue[9], [18] = no lists, 1 spare

// this CTA gets no lists: it walks from global memory every step (counted once per CTA)
__device__ __forceinline__ void nl_no_lists(const NlIO &nl, uint32_t *tab) {
    if (atomicExch(tab + 18, 1u) == 0u) atomicAdd(nl.flag, 1u);
}

// Lays the nine CTA-wide intervals (ub, ue: multiples of 4, empty = 0, 0) out in the tile and
// issues their bulk copies.  Called by warp 0; lane r = row r.  Returns the tile total.
template <class Smem>
__device__ __forceinline__ uint32_t nl_stage(Smem &S, uint32_t tid, uint32_t ub, uint32_t ue,
                                             const float *__restrict__ sx, const float *__restrict__ sy,
                                             const float *__restrict__ sz, uint32_t tile_cap) {
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
        if (staged) mbar_expect_tx(&S.bar, total * 12u);
    }
    __syncwarp();
    if (staged && tid < 9 && len) {
        bulk_g2s(&S.tx[toff], sx + ub, len * 4u, &S.bar);
        bulk_g2s(&S.ty[toff], sy + ub, len * 4u, &S.bar);
        bulk_g2s(&S.tz[toff], sz + ub, len * 4u, &S.bar);
    }
    return total;
}

struct NlBuildSmem {
    alignas(16) float tx[NL_TILE + 8], ty[NL_TILE + 8], tz[NL_TILE + 8];
    uint32_t rng[9][NL_BLOCK];  // per-thread (tile start | len << 16) per row
    uint32_t ub[9], ue[9];
    uint32_t toff[10], tslot[9];
    alignas(8) uint64_t bar;
};
constexpr int NL_STAGE_ROWS = 96;  // STAGED: rows of the shared-memory staging block (>= vcap)

// STAGED = false: every entry goes straight to global memory -- a warp's 32 lanes are at 32
// different list positions, so each 2-byte store touches its own sector (the form checked on
// hardware, FP_WALK_VARIANT=41: a build costs 4.9 ms at C4).  STAGED = true (variant 43, not yet
// run on hardware): entries are collected in shared memory, [entry][thread], and the CTA's
// block is written row by row, 256 contiguous bytes per row.
// SORTED (needs STAGED; variant 45, not yet run on hardware): the CTA's 128 boids are handed to its
// threads in descending order of the work they will cost the walk, so that the lanes of a warp
// finish together (unsorted, a warp runs as long as its longest gate list, ~44 entries against 34
// on average, and its longest drain list, ~24 against 16).  Lane l of the CTA then holds boid (count[cta * 128 + l] >> 8) of the
// CTA's slot window, with its list in column l; the count stays in the low byte.  Any
// assignment of boids to threads gives the same result: each boid still sums its own
// contributions in slot order.
template <bool STAGED, bool SORTED>
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
    float2 vhx = make_float2(0, 0), vhy = vhx, vhz = vhx;  // SORTED: direction of flight, both halves
    if (active) {
        pi4 = io.pos_s[s];
        if (SORTED) {
            const float4 vi4 = io.vel_s[s];
            work = __float_as_uint(vi4.w) == 0u;
            const float inv = rsqrtf(fmaf(vi4.z, vi4.z, fmaf(vi4.y, vi4.y, vi4.x * vi4.x)));  // a prediction: no need to be exact
            vhx = make_float2(vi4.x * inv, vi4.x * inv);
            vhy = make_float2(vi4.y * inv, vi4.y * inv);
            vhz = make_float2(vi4.z * inv, vi4.z * inv);
        } else {
            work = __float_as_uint(io.vel_s[s].w) == 0u;  // not a ghost record
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
            tab[tid] = ub;       // the layout every nl_walk_kernel launch of this binning re-uses
            tab[9 + tid] = ue;
        }
        nl_stage(S, tid, ub, ue, io.soa_in[0], io.soa_in[1], io.soa_in[2], NL_TILE);
    }
    __syncthreads();
    const uint32_t total = S.toff[9];
    if (total > (uint32_t)NL_TILE) {  // dense cluster: the tile does not fit -- no lists for this CTA
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
    uint16_t *const out = nl.entries + (size_t)blockIdx.x * nl.vcap * NL_BLOCK + tid;  // entry k at out[k * BLOCK]
    uint16_t *const stg = reinterpret_cast<uint16_t *>(smem_raw + sizeof(NlBuildSmem)) + tid;  // STAGED: same layout
    const uint32_t vcap = STAGED ? min(nl.vcap, (uint32_t)NL_STAGE_ROWS) : nl.vcap;
    uint32_t w = 0;  // entries found
    uint32_t pred = 0;  // SORTED: entries that would pass the walk's pre-gate as things stand now
    // (what that takes of DevParams -- m2_cut_hi, fov_kh, fov_kl -- sits behind the counter in nl.flag)
    const float sp_cut = SORTED ? __uint_as_float(__ldg(nl.flag + 1)) : 0.0f;
    const float sp_kh = SORTED ? __uint_as_float(__ldg(nl.flag + 2)) : 0.0f;
    const float sp_kl = SORTED ? __uint_as_float(__ldg(nl.flag + 3)) : 0.0f;
    const float2 nsx = make_float2(-pi4.x, -pi4.x), nsy = make_float2(-pi4.y, -pi4.y),
                 nsz = make_float2(-pi4.z, -pi4.z);
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        const uint32_t pk = S.rng[r][tid];
        const uint32_t t0 = pk & 0xffffu, len = pk >> 16;
        const uint32_t tag = (uint32_t)r << 12;
        // four candidates at an even tile index per batch, packed FP32 (as the production pre-gate);
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
            float ss[4] = {0, 0, 0, 0};
            if (SORTED) {  // the walk's conservative FOV test (fp_walk.cu), for the prediction
                const float2 q01 = __ffma2_rn(vhz, dz01, __ffma2_rn(vhy, dy01, __fmul2_rn(vhx, dx01)));
                const float2 q23 = __ffma2_rn(vhz, dz23, __ffma2_rn(vhy, dy23, __fmul2_rn(vhx, dx23)));
                ss[0] = q01.x * fabsf(q01.x); ss[1] = q01.y * fabsf(q01.y);
                ss[2] = q23.x * fabsf(q23.x); ss[3] = q23.y * fabsf(q23.y);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if ((live >> u & 1u) && !(mm[u] >= nl.m2_wide) && T + u != t_self) {  // NaN never drops
                    if (w < vcap) (STAGED ? stg : out)[(size_t)w * NL_BLOCK] = (uint16_t)(tag | (T + u));
                    ++w;
                    if (SORTED && !(mm[u] >= sp_cut) && !(ss[u] < sp_kh * mm[u] && ss[u] > sp_kl * mm[u])) ++pred;
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
    if (!SORTED && active) nl.count[s - io.first] = (uint16_t)min(w, vcap);
    if (__any_sync(0xffffffffu, w > vcap) && (tid & 31) == 0) nl_no_lists(nl, tab);
    if (STAGED && SORTED) {
        static_assert(!SORTED || STAGED, "the sorted layout is written from the staging block");
        // Sort key: the work a boid will cost the walk -- mostly its drain (entries that survive the
        // pre-gate: predicted from the velocities of this moment, the field of view turns slowly),
        // a little its gate (all entries).  Emulated on a C3-density flock: sorting by
        // this key takes ~14 % off the gate + drain work of a warp, by the list length alone 9 %.
        constexpr int NKEY = 128;
        __shared__ uint32_t hist[NKEY];  // threads per key, then running ranks
        __shared__ uint32_t wmax_s;
        const uint32_t c = min(w, vcap);
        const uint32_t key = min(pred + c / 4u, (uint32_t)NKEY - 1u);
        hist[tid] = 0u;  // (NKEY == NL_BLOCK)
        if (tid == 0) wmax_s = 0u;
        __syncthreads();
        atomicAdd(&hist[key], 1u);
        atomicMax(&wmax_s, c);
        __syncthreads();
        if (tid < 32) {  // first rank of each key, costliest first: lane l owns keys 127 - 4 l .. 124 - 4 l
            uint32_t t[4], sum = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                t[u] = hist[NKEY - 1 - (4 * tid + u)];
                sum += t[u];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
                if ((int)tid >= off) inc += o;
            }
            uint32_t run = inc - sum;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                hist[NKEY - 1 - (4 * tid + u)] = run;
                run += t[u];
            }
        }
        __syncthreads();
        const uint32_t rank = atomicAdd(&hist[key], 1u);  // (order within one key: whoever comes first)
        nl.count[(size_t)blockIdx.x * NL_BLOCK + rank] = (uint16_t)(c | (tid << 8));
        uint16_t *const col = nl.entries + (size_t)blockIdx.x * nl.vcap * NL_BLOCK + rank;
        const uint32_t rows = wmax_s;
        for (uint32_t k = 0; k < rows; ++k) col[(size_t)k * NL_BLOCK] = stg[(size_t)k * NL_BLOCK];
    } else if (STAGED) {
        // rows 0 .. (longest list of the CTA) of the staging block, as they are: a thread's entries
        // past its own count are whatever the shared memory held, and are never read as entries
        __shared__ uint32_t wmax;  // longest list of the CTA
        if (tid == 0) wmax = 0u;
        __syncthreads();
        const uint32_t wm = __reduce_max_sync(0xffffffffu, min(w, vcap));
        if ((tid & 31) == 0) atomicMax(&wmax, wm);
        __syncthreads();
        const uint32_t rows = wmax;
        for (uint32_t k = 0; k < rows; ++k) out[(size_t)k * NL_BLOCK] = stg[(size_t)k * NL_BLOCK];
    }
}

template <int CAP>
struct NlWalkSmem {
    alignas(16) float tx[NL_TILE + 8], ty[NL_TILE + 8], tz[NL_TILE + 8];
    uint16_t list[CAP][NL_BLOCK];  // per-thread survivor lists: tile offsets
    uint32_t toff[10], tslot[9];
    alignas(8) uint64_t bar;
};

// <CAP, CTAS>: survivor-list capacity and the CTAs per SM the registers are budgeted for.
// <64, 5> is the form checked on hardware (39 KB of shared memory: five CTAs per SM, 102
// registers).  <48, 6> (variant 44, not yet run): 35 KB, six CTAs per SM at 80 registers -- with
// ~17 survivors per boid a 48-entry list still drains once per boid almost always.
template <int CAP, int CTAS, bool SORTED>
__global__ void __launch_bounds__(NL_BLOCK, CTAS)
nl_walk_kernel(const DevParams P, const GridDesc g, const WalkIO io, const NlIO nl, unsigned *__restrict__ status) {
    if (io.ctl && io.ctl->stale) return;  // lazy re-binning: this step is void
    const float4 *__restrict__ vel_s = io.vel_s;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    NlWalkSmem<CAP> &S = *reinterpret_cast<NlWalkSmem<CAP> *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    // SORTED: which boid of the CTA's window this thread holds, and its count, come packed from the build
    // (a CTA without lists keeps the plain assignment: its build may not have got as far as the ranking)
    const bool ranked = SORTED && !__ldg(nl.cta_tab + (size_t)blockIdx.x * NL_CTA_WORDS + 18);
    const uint32_t packed = ranked ? __ldg(nl.count + (size_t)blockIdx.x * NL_BLOCK + tid) : 0u;
    const uint32_t s = io.first + blockIdx.x * NL_BLOCK + (ranked ? packed >> 8 : tid);
    const bool active = s < io.last;
    if (tid == 0) mbar_init(&S.bar, 1);

    float4 pi4 = make_float4(0, 0, 0, 0), vi4 = make_float4(0, 0, 0, 0);
    uint32_t n_c = 0;  // cached candidates of this boid
    if (active) {
        pi4 = io.pos_s[s];
        vi4 = vel_s[s];
        n_c = ranked ? (packed & 0xffu) : __ldg(nl.count + (s - io.first));
    }
    // this CTA's list block and the first batch of entries, in flight while the tile is staged
    const uint16_t *const vlp = nl.entries + (size_t)blockIdx.x * nl.vcap * NL_BLOCK + tid;
    uint32_t e0 = __ldcs(vlp), e1 = __ldcs(vlp + NL_BLOCK), e2 = __ldcs(vlp + 2 * NL_BLOCK),
             e3 = __ldcs(vlp + 3 * NL_BLOCK);
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
    const uint32_t *const tab = nl.cta_tab + (size_t)blockIdx.x * NL_CTA_WORDS;
    V3 acc = v3zero();
    if (__ldg(tab + 18)) {
        // a CTA without lists (tile or list overflow at build time): the one-phase walk of the 27
        // home cells from global memory, as the production kernel does for its overflowing CTAs
        if (work) {
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
                        const float4 vj = __ldg(vel_s + j);
                        V3 contrib;
                        if (pair_flock(P, self, d, m2, v3(vj.x, vj.y, vj.z), contrib)) acc = vadd(acc, contrib);
                    }
                }
            }
        }
        if (active) walk_finish<TAP_STEP>(P, s, pi4, vi4, self, acc, 0u, 0ull, io, status, TapOut{});
        return;
    }
    __syncthreads();  // the barrier is initialised
    if (tid < 32) {
        const uint32_t ub = tid < 9 ? __ldg(tab + tid) : 0u, ue = tid < 9 ? __ldg(tab + 9 + tid) : 0u;
        nl_stage(S, tid, ub, ue, io.soa_in[0], io.soa_in[1], io.soa_in[2], NL_TILE);
    }
    __syncthreads();
    const uint32_t t_self = work ? s - S.tslot[4] : 0xffffu;
    const uint32_t nmax = __reduce_max_sync(0xffffffffu, n_c);
    if (S.toff[9] > 0) mbar_wait(&S.bar, 0);  // (a layout with lists always fits the tile)

    uint16_t *const lst = &S.list[0][tid];  // entry k at lst[k * BLOCK]
    int cnt = 0;
    uint32_t base = 0;  // warp-uniform progress through the cached lists, a multiple of 4
    const float kh = P.fov_kh, kl = P.fov_kl;
#pragma unroll 1
    for (;;) {
        int room = CAP - (int)__reduce_max_sync(0xffffffffu, (unsigned)cnt);
        const bool more = base < nmax;
        if (room < 4 || (!more && room < CAP)) {
            drain_list<NL_BLOCK>(P, self, lst, cnt, S.tx, S.ty, S.tz, S.tslot, t_self, vel_s, acc);
            cnt = 0;
            room = CAP;
        }
        if (!more) break;
        const uint32_t end = min(base + ((uint32_t)room & ~3u), (nmax + 3u) & ~3u);
        uint32_t w = (uint32_t)cnt * NL_BLOCK;  // list cursor, in entries
#pragma unroll 1
        for (uint32_t k = base; k < end; k += 4) {
            const uint32_t c[4] = {e0, e1, e2, e3};
            if (k + 4 < nmax) {  // the next batch: rows k + 4 .. k + 7 < vcap (vcap is a multiple of 4)
                e0 = __ldcs(vlp + (size_t)(k + 4) * NL_BLOCK);
                e1 = __ldcs(vlp + (size_t)(k + 5) * NL_BLOCK);
                e2 = __ldcs(vlp + (size_t)(k + 6) * NL_BLOCK);
                e3 = __ldcs(vlp + (size_t)(k + 7) * NL_BLOCK);
            }
            // The production pre-gate (fp_walk.cu), one candidate per lane-slot: fused squared
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

}  // namespace

size_t nl_entries_elems(uint32_t rows, uint32_t vcap) {
    const size_t ctas = ((size_t)rows + NL_BLOCK - 1) / NL_BLOCK;
    return (ctas * vcap + 8) * NL_BLOCK;  // + slack rows: the walk's first batch is loaded unconditionally
}
size_t nl_cta_tab_elems(uint32_t rows) {
    return (((size_t)rows + NL_BLOCK - 1) / NL_BLOCK) * NL_CTA_WORDS;
}

int launch_nl_build(cudaStream_t st, const GridDesc &g, const WalkIO &io, const NlIO &nl, int form,
                    const float sort_params[3]) {
    if (io.last <= io.first) return FP_OK;
    const uint32_t ctas = (io.last - io.first + NL_BLOCK - 1) / NL_BLOCK;
    if (form & (NL_FORM_STAGED | NL_FORM_SORTED)) {
        const int smem = (int)(sizeof(NlBuildSmem) + sizeof(uint16_t) * NL_STAGE_ROWS * NL_BLOCK);
        if (form & NL_FORM_SORTED) {
            // (pageable source: the copy has left the host buffer when the call returns)
            FP_CUDA(cudaMemcpyAsync(nl.flag + 1, sort_params, 3 * sizeof(float), cudaMemcpyHostToDevice, st));
            FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            nl_build_kernel<true, true><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
        } else {
            FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            nl_build_kernel<true, false><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
        }
    } else {
        const int smem = (int)sizeof(NlBuildSmem);
        FP_CUDA(cudaFuncSetAttribute(nl_build_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        nl_build_kernel<false, false><<<ctas, NL_BLOCK, smem, st>>>(g, io, nl);
    }
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

template <int CAP, int CTAS, bool SORTED>
static int launch_nl_walk_as(cudaStream_t st, const DevParams &P, const GridDesc &g, const WalkIO &io,
                             const NlIO &nl, unsigned *status) {
    const int smem = (int)sizeof(NlWalkSmem<CAP>);
    auto kern = nl_walk_kernel<CAP, CTAS, SORTED>;
    FP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<(io.last - io.first + NL_BLOCK - 1) / NL_BLOCK, NL_BLOCK, smem, st>>>(P, g, io, nl, status);
    return FP_OK;
}

int launch_nl_walk(cudaStream_t st, const DevParams &P, const GridDesc &g, const WalkIO &io, const NlIO &nl,
                   unsigned *status, int form) {
    if (io.last <= io.first) return FP_OK;
    const bool six = form & NL_FORM_SIX_CTAS, sorted = form & NL_FORM_SORTED;
    const int rc = six ? (sorted ? launch_nl_walk_as<48, 6, true>(st, P, g, io, nl, status)
                                 : launch_nl_walk_as<48, 6, false>(st, P, g, io, nl, status))
                       : (sorted ? launch_nl_walk_as<NL_CAP, 5, true>(st, P, g, io, nl, status)
                                 : launch_nl_walk_as<NL_CAP, 5, false>(st, P, g, io, nl, status));
    if (rc) return rc;
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

}  // namespace fp

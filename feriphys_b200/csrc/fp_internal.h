// fp_internal.h -- host-side declarations shared by the translation units of
// libferiphys_cuda.so.  Nothing here is part of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/feriphys_cuda.h"
#include "fp_device.cuh"

namespace fp {

extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t k = 1) { g_launches.fetch_add(k, std::memory_order_relaxed); }

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FP_CUDA(expr)                                                          \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return ::fp::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)

// Tap selector for the influence kernels.
enum Tap {
    TAP_STEP = 0,       // advance the state (flocking.rs:97-122)
    TAP_ACCEL = 1,      // write total acceleration (+ components) by caller index
    TAP_NEIGHBORS = 2,  // write |N(i)| and hash by caller index
    TAP_CENSUS = 3      // accumulate pair-outcome counts
};

struct TapOut {
    float *accel3;               // n x 3 (TAP_ACCEL)
    float *comp15;               // n x 15 or NULL
    uint32_t *nbr_count;         // n (TAP_NEIGHBORS)
    unsigned long long *nbr_hash;  // n
    unsigned long long *census;  // 4 counters (TAP_CENSUS)
};

// ---- uniform grid description (host + device) ------------------------------
struct GridDesc {
    float origin[3];
    float inv_cell;
    float cell;
    int dim[3];        // cells of THIS rank's grid (slab + one halo layer each side when sharded)
    uint32_t ncells;   // dim[0] * dim[1] * dim[2]
    uint32_t key_bits; // bits of ncells (the value ncells itself is the "dead record" key)
    int gdimx;         // global cell layers along x (== dim[0] on a single GPU)
    int xoff;          // global layer index of local layer 0 (0 on a single GPU)
    float skin;        // cell = reach * (1 + 1/512) + skin: a binning stays exact while every
                       // boid is within skin / 2 of the position it was binned at
    // Along z (the fastest key dimension, the direction of a walk "row") cells are split into
    // zspan slices: a boid's row range is cz - zspan .. cz + zspan, i.e. (2 + 1/zspan) cell
    // edges instead of 3 -- fewer candidates at no extra rows.
    float inv_cell_z;  // zspan / cell
    int zspan;         // 1, 2 or 4
};

// Device-side control block of the lazy re-binning (fp_misc.cu: skin_gate_kernel).
struct SkinCtl {
    float D;               // upper bound on any boid's displacement since the last binning
    uint32_t v2max;        // bits of max |v|^2 over the inputs of the last walk
    uint32_t pmax;         // bits of max |coordinate| over the inputs of the last walk
    uint32_t stale;        // sticky: D exceeded skin / 2 -- every later kernel is a no-op
    uint32_t first_stale;  // ordinal of the first step that was not performed
    uint32_t g_v2max;      // what the last gate folded into D (sharded: the maximum over all ranks,
    uint32_t g_pmax;       //  identical on every rank -- the host plans from these)
    uint32_t fault;        // a peer never posted its step (time-out): the run cannot continue
};

// Sharded grid: after each walk every rank posts (tag = ordinal + 1, its max |v|^2, max |coord|)
// into slot [its rank] of EVERY rank's mailbox with peer stores over NVLink.  The next step's
// gate waits for all tags: that is the step barrier (ghosts pushed by the neighbours' walks have
// landed) and the all-reduce of the speed bound in one.
// A rank that has seen all tags may walk and post its NEXT step before a slower rank has read
// this one, so a mailbox holds two sets of slots, used alternately by tag parity (no rank can
// be two steps ahead: its next gate needs the slow rank's next post).
struct Mail {
    uint32_t tag, v2max, pmax, pad;
};
constexpr int FP_MAX_WORLD = 16;
constexpr int FP_MAIL_SLOTS = 2 * FP_MAX_WORLD;  // [tag & 1][source rank]
struct MailPeers {
    Mail *box[FP_MAX_WORLD];  // box[q] = rank q's mailbox (peer-mapped; own entry = local)
};
enum { GATE_LOCAL = 0, GATE_REDUCED = 1, GATE_MAILBOX = 2 };

// Sharded grid: a boundary layer of this rank's slab is the neighbour's ghost layer.  The walk
// writes the advanced state of slots [begin, end) straight into the neighbour's next-step
// buffers (peer memory over NVLink), record (slot - begin) of its ghost block.  end == 0: off.
struct PeerFace {
    uint32_t begin, end;
    float4 *pos, *vel;
    float *sx, *sy, *sz;
};

// Everything one walk launch reads and writes.
struct WalkIO {
    const float4 *pos_s, *vel_s;   // cell-sorted records (owned, and ghost copies when sharded)
    const float *soa_in[3];        // x, y, z of pos_s: the TMA source of the staged walk
    const uint32_t *home;          // cell key each slot was binned under (owned slots)
    const uint32_t *cell_start;    // ncells + 1
    uint32_t first, last;          // slots this launch steps / taps (owned); candidates are any slot
    float4 *pos_out, *vel_out;     // TAP_STEP: the advanced state, same slots
    float *soa_out[3];             // TAP_STEP: x, y, z of pos_out
    SkinCtl *ctl;                  // TAP_STEP: speed tracking + stale check (may be null)
    PeerFace push[2];              // TAP_STEP, sharded: halo push to the left / right neighbour
};

// ---- all-pairs (fp_allpairs.cu) --------------------------------------------
// rows [row0, row0 + nrows) of the local output against all n_all boids of
// pos_all/vel_all.  For TAP_STEP writes pos_out/vel_out[0..nrows).
int launch_allpairs(cudaStream_t st, const DevParams &P, int tap, const float4 *pos_all,
                    const float4 *vel_all, uint32_t n_all, uint32_t row0, uint32_t nrows,
                    float4 *pos_out, float4 *vel_out, unsigned *status, const TapOut &tap_out, int variant = 0,
                    const float *bounds8 = nullptr);  // FAST numerics: launch_bounds() of pos_all (device)

// One CTA, `nsteps` steps in one launch, reference summation order (fp_small.cu).
// lead_table: nsteps rows of n_leads x 8 floats, or NULL to use P.leads for every step.
// aos_in (may be null): the state comes from these caller-order rows [px py pz vx vy vz] (mapped host
// memory) instead of pos / vel; aos_out (may be null): the advanced rows are also written there.
int launch_small(cudaStream_t st, const DevParams &P, float4 *pos, float4 *vel, uint32_t n,
                 uint32_t nsteps, const float *lead_table, uint32_t lead_rows, unsigned *status,
                 const float *aos_in = nullptr, float *aos_out = nullptr, uint32_t first_index = 0);
uint32_t small_max_boids();

// ---- grid (fp_grid.cu, fp_sort.cu) -----------------------------------------
struct GridWork {  // device scratch owned by the handle
    uint32_t *keys[2];
    uint32_t *vals[2];
    uint32_t *cell_start;  // ncells + 1
    uint32_t *tile_hist;   // radix-sort per-tile digit histograms
    uint32_t *scan_tmp;    // block partials for the scans
    size_t tile_hist_elems, scan_tmp_elems, cell_cap;
    uint32_t cap;          // boid capacity of keys/vals
    float *soa[2][3];      // x, y, z of the sorted state (cap + 8 floats each), double-buffered like
                           // pos/vel: the walk's TMA source, rewritten by the walk for the next step
    uint32_t soa_cap;
    int soa_cur;           // which SoA copy belongs to the current sorted state
    const uint32_t *home;  // sorted cell keys of the current binning (one of keys[])
    SkinCtl *ctl;          // device control block
};

// keys[0][i] = cell key of pos[i], vals[0][i] = i; cell_start <- exclusive scan of counts
int launch_grid_keys(cudaStream_t st, const GridDesc &g, const float4 *pos, uint32_t n, GridWork &w);
// stable LSD radix sort of (keys[0], vals[0]) on key bits [0, key_bits); result index in *out_buf
int launch_radix_sort(cudaStream_t st, GridWork &w, uint32_t n, uint32_t key_bits, int *out_buf);
// pos_out[i] = pos_in[vals[i]] (same for vel)
// (no-op when ctl->stale)
int launch_grid_reorder(cudaStream_t st, const uint32_t *vals, const float4 *pos_in,
                        const float4 *vel_in, float4 *pos_out, float4 *vel_out, float *const *soa,
                        uint32_t n, const SkinCtl *ctl);
// 27-cell walk over the n_all sorted records (owned boids are stepped / tapped; ghost
// and dead records of a sharded flock only serve as candidates).
int launch_grid_walk(cudaStream_t st, const DevParams &P, const GridDesc &g, int tap, const WalkIO &io,
                     unsigned *status, const TapOut &tap_out);

// TMA-staged two-phase walk (fp_walk.cu): TAP_STEP / TAP_ACCEL without candidate lists
int launch_grid_walk3(cudaStream_t st, const DevParams &P, const GridDesc &g, int tap, const WalkIO &io,
                      unsigned *status, const TapOut &tap_out);

// Standing candidate lists (fp_walk_nl.cu): the distance pre-gate of the walk done once per
// binning instead of once per step.  The default form of a grid step.
struct NlIO {
    uint16_t *entries;   // [cta][vcap][128]: tile offsets (row << 12 | offset) of the candidates in reach + skin
    uint16_t *count;     // [slot - first]: entries of each boid
    uint32_t *cta_tab;   // [cta][20]: the nine staged intervals of each CTA, [18] != 0: this CTA has no lists
    unsigned *flag;      // [0]: CTAs the last build left without lists (tile or list overflow)
    uint32_t vcap;       // list capacity, a multiple of 4
    uint32_t tile_cap;   // staged candidates the walk that will use the lists can hold per CTA
    uint32_t tile_cap_b; // ... in its positions-only form (fast walk; 0: there is no such form)
    float m2_wide;       // build cut: (reach + skin)^2 (1 + 1e-5)
    int vis_first;       // order each list with the entries predicted to be in view first (fast walk)
    float vis_c;         // ... in view <=> cosine of the sight angle > vis_c
};
size_t nl_entries_elems(uint32_t rows, uint32_t vcap);
size_t nl_cta_tab_elems(uint32_t rows);
uint32_t nl_tile_cap(bool fast);
uint32_t nl_tile_cap_b(bool fast);
// after a binning, before the first walk that uses the lists
int launch_nl_build(cudaStream_t st, const GridDesc &g, const WalkIO &io, const NlIO &nl);
// a step (TAP_STEP) -- or, under FAST numerics, the acceleration tap -- on the standing lists.
// EXACT numerics: same bits as launch_grid_walk.  FAST numerics (P.numerics_fast): same neighbour
// sets, accelerations within ~1e-6 relative.
int launch_nl_walk(cudaStream_t st, const DevParams &P, const GridDesc &g, int tap, const WalkIO &io, const NlIO &nl,
                   unsigned *status, const TapOut &tap_out);

// Lazy re-binning control (fp_misc.cu).  Runs before each grid step: on a re-binning step it
// resets the displacement bound, otherwise it adds the last walk's bound
// dt * max|v| + rounding and marks the flock stale when the bound exceeds `budget`.
// mode GATE_LOCAL: the bound comes from this GPU's walk; GATE_REDUCED: from ctl->g_v2max / g_pmax
// (an NCCL max-all-reduce put them there); GATE_MAILBOX: wait for `world` tags == ordinal in
// `mail`, then take the maximum over the posts.
int launch_skin_gate(cudaStream_t st, SkinCtl *ctl, uint32_t ordinal, int rebin, float dt, float budget,
                     int mode = GATE_LOCAL, const Mail *mail = nullptr, int world = 1, unsigned *status = nullptr);
int launch_mail_post(cudaStream_t st, const SkinCtl *ctl, const MailPeers &peers, int world, int rank,
                     uint32_t tag);

// ---- misc kernels (fp_misc.cu) ----------------------------------------------
int launch_aos6_to_soa(cudaStream_t st, const float *aos6, float4 *pos, float4 *vel, uint32_t n,
                       uint32_t first_index);
// scatter by the caller index carried in pos.w (relative to first_index)
int launch_soa_to_aos6(cudaStream_t st, const float4 *pos, const float4 *vel, float *aos6,
                       uint32_t n, uint32_t first_index, int by_index);
int launch_unpermute(cudaStream_t st, const float4 *pos_in, const float4 *vel_in, float4 *pos_out,
                     float4 *vel_out, uint32_t n, uint32_t first_index);
int launch_instances(cudaStream_t st, const float4 *pos, const float4 *vel, float *out, uint32_t n,
                     uint32_t first_index, int raw);
int launch_state_combine_euler(cudaStream_t st, const float *s, const float *ds, float h, float *out,
                               size_t n);
int launch_state_combine_rk4(cudaStream_t st, const float *s, const float *k1, const float *k2,
                             const float *k3, const float *k4, float h, float *out, size_t n);
// State<boid>::euler_step / rk4_step with frozen acceleration: accel3 by internal order
int launch_flock_state_step(cudaStream_t st, float4 *pos, float4 *vel, const float *accel3_by_index,
                            uint32_t n, uint32_t first_index, float h, int rk4);
int launch_fastmath_check(uint64_t n, uint64_t seed, uint64_t out_mismatch[2]);
// out8: min xyz, max xyz, max |v|^2, unused
int launch_bounds(cudaStream_t st, const float4 *pos, const float4 *vel, uint32_t n, float *out8 /*device*/);
// generic exclusive scan in place over n uint32 (n <= 2^28); tmp >= (n/4096 + 2) elements
int launch_exclusive_scan(cudaStream_t st, uint32_t *data, size_t n, uint32_t *tmp);

}  // namespace fp

// fp_api.cu -- the C ABI of libferiphys_cuda.so (include/feriphys_cuda.h):
// handle management, host-side threshold derivation, step orchestration.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>

#include "fp_internal.h"
#include "fp_flock.h"
#include "fp_shard.h"

namespace fp {

std::atomic<uint64_t> g_launches{0};
static thread_local std::string t_error;

void set_error(const std::string &msg) { t_error = msg; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " (" + file + ":" +
              std::to_string(line) + ")";
    return FP_ERR_CUDA;
}

// ---- host-side thresholds ----------------------------------------------------
// Ordinals of non-negative floats are their bit patterns.
static inline uint32_t f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
static inline float u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// smallest m2 >= 0 with pred(sqrtf(m2)) true, for a predicate monotone (false..true) in m2;
// returns NaN when no finite-or-inf m2 satisfies it.
template <class Pred>
static float smallest_m2(Pred pred) {
    const uint32_t inf = 0x7f800000u;
    if (pred(sqrtf(0.0f))) return 0.0f;
    if (!pred(sqrtf(u2f(inf)))) return NAN;
    uint32_t lo = 0, hi = inf;  // pred(lo) false, pred(hi) true
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (pred(sqrtf(u2f(mid)))) hi = mid; else lo = mid;
    }
    return u2f(hi);
}

// largest c in [-1, 1] with acosf(c) > theta; -2 when none.  acosf here is the
// platform libm -- the function Rust's f32::acos calls.
static float acos_threshold(float theta) {
    if (!(acosf(-1.0f) > theta)) return -2.0f;
    if (acosf(1.0f) > theta) return 1.0f;
    auto ord = [](float f) -> int64_t {
        const uint32_t u = f2u(f);
        return (u & 0x80000000u) ? -(int64_t)(u & 0x7fffffffu) : (int64_t)u;
    };
    auto unord = [](int64_t o) -> float {
        return o < 0 ? u2f(0x80000000u | (uint32_t)(-o)) : u2f((uint32_t)o);
    };
    int64_t lo = ord(-1.0f), hi = ord(1.0f);
    while (hi - lo > 1) {
        const int64_t mid = lo + (hi - lo) / 2;
        if (acosf(unord(mid)) > theta) lo = mid; else hi = mid;
    }
    return unord(lo);
}

void derive_params(const fp_config &c, DevParams &P, bool want_fast) {
    P.dt = c.dt;
    P.f_c = c.centering_factor;
    P.f_v = c.velocity_matching_factor;
    P.neg_f_a = -1.0f * c.avoidance_factor;
    P.thr = c.distance_weight_threshold;
    P.fall = c.distance_weight_threshold_falloff;
    const float thr = P.thr;
    const float R = thr + P.fall;  // f32 sum, as boid.rs:155
    // rejected  <=>  !(dist <= thr) && dist >= R
    const float m2_ge_R = smallest_m2([R](float d) { return d >= R; });
    const float m2_gt_thr = smallest_m2([thr](float d) { return !(d <= thr); });
    P.m2_cut = (isnan(m2_ge_R) || isnan(m2_gt_thr)) ? NAN : std::max(m2_ge_R, m2_gt_thr);
    // the FMA form of dx^2 + dy^2 + dz^2 is within 4e-7 relative of the separately rounded one
    P.m2_cut_hi = std::isfinite(P.m2_cut) ? nextafterf(P.m2_cut * 1.000001f, INFINITY) + 1e-37f : P.m2_cut;
    // weight 1  <=>  dist <= thr  <=>  m2 <= m2_one
    if (isnan(m2_gt_thr)) P.m2_one = INFINITY;           // every distance is <= thr
    else if (m2_gt_thr == 0.0f) P.m2_one = -1.0f;        // none is
    else P.m2_one = u2f(f2u(m2_gt_thr) - 1);             // predecessor
    P.cstar = acos_threshold(c.max_sight_angle);
    P.cstar_lead = acos_threshold(c.max_sight_angle_to_lead_boid);
    {
        const float h = P.cstar - 1e-5f, l = -1.0f + 1e-5f;
        P.fov_kh = h * fabsf(h);
        P.fov_kl = l * fabsf(l);
    }
    P.steer_secs = c.time_to_start_steering_secs;
    P.steer_nanos = c.time_to_start_steering_nanos;
    P.steering_overrides = c.steering_overrides ? 1 : 0;
    // branch-free exact path: every scalar a normal number of moderate exponent
    auto moderate = [](float x) { return std::isfinite(x) && fabsf(x) >= 0x1p-40f && fabsf(x) <= 0x1p40f; };
    auto bounded = [](float x) { return std::isfinite(x) && fabsf(x) <= 0x1p40f; };
    P.fast_ok = moderate(P.neg_f_a) && moderate(P.fall) && P.fall > 0.0f && bounded(P.thr) && bounded(P.f_c) &&
                bounded(P.f_v);
    int e = 0;
    P.fall_pow2 = P.fast_ok && frexpf(P.fall, &e) == 0.5f;
    P.inv_fall = P.fall_pow2 ? 1.0f / P.fall : 0.0f;
    // FAST numerics: usable when the thresholds are ordinary numbers (else the exact kernels run)
    const bool fz_ok = P.fast_ok && std::isfinite(P.m2_cut) && P.m2_cut > 0.0f && P.m2_one < P.m2_cut &&
                       P.cstar >= -2.0f && P.cstar <= 1.0f && P.cstar_lead >= -2.0f;
    P.numerics_fast = want_fast && fz_ok;
    P.fz_one = P.m2_one;
    if (P.cstar < -1.0f) P.fz_a = P.fz_b = -3.0f;  // nothing is ever culled
    else { P.fz_a = P.cstar; P.fz_b = -1.0f; }
    // the fused cosine is within 1e-6 of the reference's; decisions within 1e-5 of either end of
    // the culled interval are re-taken exactly: |(c - a)(c - b)| <= 1e-5 (|a - b| + 1e-5) covers them
    P.fz_gc_tol = 1e-5f * (fabsf(P.fz_a - P.fz_b) + 1e-5f);
    P.fz_gm_tol = 0.0f;  // (the fast walk computes the squared distance exactly: no band needed)
    P.fz_rinv_fall = P.fall != 0.0f ? 1.0f / P.fall : 0.0f;
    {
        // obstacles farther than radius + |v| t_start (1 + 1e-4) cannot start steering: skipped.
        // Safe against Duration's ns rounding when t_start 1e-4 >> 1 ns.
        const double ts = (double)c.time_to_start_steering_secs + 1e-9 * (double)c.time_to_start_steering_nanos;
        P.fz_steer_reach = (ts >= 1e-4 && ts < 1e30) ? (float)(ts * 1.0001) : -1.0f;
    }
}

}  // namespace fp

using namespace fp;

// struct fp_flock lives in fp_flock.h (shared with fp_shard.cu)

namespace {

template <class T>
int dev_alloc(T **p, size_t count) {
    *p = nullptr;
    if (!count) count = 1;
    FP_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    return FP_OK;
}
template <class T>
void dev_free(T *&p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// bookkeeping of the candidate lists: the flock has just been binned
void nl_binned(fp_flock *f) {
    f->nl_fresh = true;
    f->nl_prev_bin_steps = f->nl_bin_steps;
    f->nl_bin_steps = 0;
}

int ensure_stage(fp_flock *f, size_t bytes) {
    if (bytes <= f->stage_bytes) return FP_OK;
    if (f->d_stage) cudaFree(f->d_stage);
    f->d_stage = nullptr;
    f->stage_bytes = 0;
    FP_CUDA(cudaMalloc(&f->d_stage, bytes));
    f->stage_bytes = bytes;
    return FP_OK;
}

int settle(fp_flock *f);
bool nl_enabled();  // candidate lists (below): FP_NL=0 turns them off

// every entry point except fp_flock_step: the handle must be valid and every enqueued step
// must be known to have happened (lazy re-binning may have voided some: settle replays them)
// rows left in a pinned slot by fp_flock_write_state become the device state (one kernel reading
// mapped host memory: no copy engine involved)
int ingest_pending(fp_flock *f) {
    if (!f->pending_aos) return FP_OK;
    int rc = launch_aos6_to_soa(f->stream, f->pending_aos, f->pos[f->cur], f->vel[f->cur], f->n, f->first_index);
    if (rc) return rc;
    FP_CUDA(cudaEventRecord(f->up_ev[f->pending_slot], f->stream));
    f->pending_aos = nullptr;
    return FP_OK;
}

// keep_io: the caller deals with rows pending in a pinned slot and with the mapped output rows
// itself (fp_flock_step, fp_flock_read_state); everybody else sees plain device state
int check(fp_flock *f, bool settled = true, bool keep_io = false) {
    if (!f) {
        set_error("null flock handle");
        return FP_ERR_INVALID;
    }
    FP_CUDA(cudaSetDevice(f->device));
    if (!keep_io) {
        f->out_valid = false;
        int rc = ingest_pending(f);
        if (rc) return rc;
    }
    return settled ? settle(f) : FP_OK;
}

// (the allocation is kept when the new table fits it)
int upload_table(fp_flock *f, float **dst, size_t *cap, const float *src, size_t floats) {
    if (!floats) return FP_OK;  // (the count the kernels see is 0; the old table is never read)
    if (!*dst || floats > *cap) {
        dev_free(*dst);
        *cap = 0;
        int rc = dev_alloc(dst, floats);
        if (rc) return rc;
        *cap = floats;
    }
    FP_CUDA(cudaMemcpyAsync(*dst, src, floats * sizeof(float), cudaMemcpyHostToDevice, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));  // caller's buffer is not retained
    return FP_OK;
}

void refresh_tables(fp_flock *f) {
    f->P.leads = f->n_leads ? f->d_lead_ring + (size_t)(f->lead_ver % fp_flock::LEAD_RING) * f->n_leads * 8 : nullptr;
    f->P.n_leads = (int)f->n_leads;
    f->P.attractors = f->d_attr;
    f->P.n_attractors = (int)f->n_attr;
    f->P.obstacles = f->d_obs;
    f->P.n_obstacles = (int)f->n_obs;
}

float reach_of(const fp_config &c) {
    const float thr = c.distance_weight_threshold;
    const float R = thr + c.distance_weight_threshold_falloff;
    return std::max(thr, R);
}

void free_grid_work(fp_flock *f) {
    dev_free(f->work.keys[0]);
    dev_free(f->work.keys[1]);
    dev_free(f->work.vals[0]);
    dev_free(f->work.vals[1]);
    dev_free(f->work.cell_start);
    dev_free(f->work.tile_hist);
    dev_free(f->work.scan_tmp);
    for (auto &b : f->work.soa) for (auto &p : b) dev_free(p);
    dev_free(f->work.ctl);
    f->work = GridWork{};
}

// Planning estimate of the displacement bound one step adds (skin_gate_kernel computes the
// real one): 25 % margin on the speed, so a plan only fails when the fastest boid gains more.
float plan_delta(float v2max, float pmax, float dt) {
    if (!(v2max >= 0.0f)) v2max = INFINITY;
    return 1.25f * sqrtf(v2max) * fabsf(dt) + pmax * 2.1e-7f + 1e-30f;
}
// steps a fresh binning is planned to serve
int64_t plan_steps(const fp_flock *f, float D, float first_delta) {
    const float room = f->skin_budget - D - first_delta;
    if (!(room >= 0.0f)) return 0;
    const double m = 1.0 + floor((double)room / (double)f->delta_est);
    return (int64_t)std::min(1.0e6, m * (double)f->plan_scale);
}

// Fit the uniform grid to the current positions (or the user's domain).
int fit_grid(fp_flock *f) {
    const float reach = reach_of(f->cfg);
    if (!(reach > 0.0f) || !std::isfinite(reach)) {
        set_error("grid method needs a finite positive distance_weight_threshold + falloff");
        return FP_ERR_UNSUPPORTED;
    }
    float lo[3], hi[3], v2max = 0.0f;
    {
        int rc = launch_bounds(f->stream, f->pos[f->cur], f->vel[f->cur], f->n, f->d_bounds);
        if (rc) return rc;
        float b[8];
        FP_CUDA(cudaMemcpyAsync(b, f->d_bounds, sizeof(b), cudaMemcpyDeviceToHost, f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
        for (int a = 0; a < 3; ++a) {
            lo[a] = b[a];
            hi[a] = b[3 + a];
            if (!(lo[a] <= hi[a])) lo[a] = hi[a] = 0.0f;
        }
        v2max = b[6];
        if (f->shard) shard_reduce_bounds(f->shard, f->stream, lo, hi, &v2max);
    }
    float pmax = 0.0f;
    for (int a = 0; a < 3; ++a) pmax = std::max(pmax, std::max(fabsf(lo[a]), fabsf(hi[a])));
    if (f->domain_user) {
        memcpy(lo, f->user_lo, sizeof(lo));
        memcpy(hi, f->user_hi, sizeof(hi));
    }
    // Lazy re-binning: a binning stays exact while every boid is within skin / 2 of where it
    // was binned, at the price of (1 + skin / reach)^3 more candidates.  The skin that balances
    // the two for this flock's speed: skin = sqrt(1.9 * (binning cost / step cost) * delta * reach)
    // (DESIGN.md 4.1), delta = the per-step displacement bound.  Flocks too fast for two steps
    // per binning get none.
    const float delta = plan_delta(v2max, pmax, f->cfg.dt);
    // cost ratio binning : step.  Both grow with the boids a GPU holds, but a binning also has a
    // fixed part -- ~20 launches, and on a sharded flock four NCCL groups and three host syncs
    // (0.5 ms fitted the 8-GPU measurements of round 1) -- that a small or sharded flock
    // amortises over more steps with a larger skin.
    const double n_here = std::max<double>(1.0, f->shard ? (double)f->n_global / shard_world(f->shard) : f->n);
    // per-boid costs measured at C4 (ns): a binning 0.06 and, with candidate lists, their build 0.23;
    // a step 0.34 (staged walk), 0.20 (exact list walk) or 0.11 (fast list walk)
    const bool lists = nl_enabled() && !f->nl_off;
    const double c_bin = 0.06e-9 + (lists ? 0.23e-9 : 0.0);
    const double c_step = !lists ? 0.34e-9 : (f->P.numerics_fast ? 0.11e-9 : 0.20e-9);
    const double ratio = (c_bin * n_here + (f->shard ? 0.5e-3 : 1.0e-4)) / (c_step * n_here);
    float skin = std::min(sqrtf((float)(1.9 * ratio) * delta * reach), reach / 8.0f);
    if (!(skin > 0.0f) || !std::isfinite(skin) || skin / 2.0f / delta < 2.0f) skin = 0.0f;
    {
        static const char *env = getenv("FP_SKIN");  // tuning: "0" disables, else a fixed skin
        if (env && *env) skin = std::max(0.0f, (float)atof(env));
        if (f->skin_override >= 0.0f) skin = f->skin_override;  // fp_flock_set_rebin
        skin = std::min(skin, 64.0f * reach);
    }
    if (f->shard && shard_is_slab(f->shard) && f->grid_valid && !f->domain_user && f->grid.skin > 0.0f) {
        // A periodic re-fit of a SHARDED flock repartitions everything through the index-ordered
        // interchange form (two flock-sized all-reduces, a fresh selection, new peer mappings: ~80 ms
        // at C4 on two GPUs against 1.4 ms steps).  It only matters for speed -- a boid outside the
        // grid is clamped into an edge cell, still exact -- so it is skipped while the standing
        // grid still covers the flock (half a cell of overhang allowed) and its skin is within a
        // factor 1.6 of the one wanted now.  (Every rank sees the same reduced bounds: same decision.)
        const GridDesc &o = f->grid;
        const double dims[3] = {(double)o.gdimx, (double)o.dim[1], (double)o.dim[2] / o.zspan};
        bool covers = true;
        for (int a = 0; a < 3; ++a) {
            const double glo = (double)o.origin[a] - 0.5 * o.cell, ghi = (double)o.origin[a] + (dims[a] + 0.5) * o.cell;
            covers = covers && (double)lo[a] >= glo && (double)hi[a] <= ghi;
        }
        if (covers && skin > 0.0f && skin <= 1.6f * o.skin && o.skin <= 1.6f * skin) {
            f->delta_est = delta;
            f->steps_since_fit = 0;
            return FP_OK;
        }
    }
    f->skin_budget = skin / 2.0f;
    f->delta_est = delta;
    // cell edge > reach by a margin that covers the f32 rounding of the cell coordinate
    // (relative 2^-23 of a coordinate < 4096 cells) and of the distance itself.
    double cell = (double)reach * (1.0 + 1.0 / 512.0) + (double)skin;
    GridDesc g{};
    static const int zspan_env = [] {
        const char *e = getenv("FP_GRID_ZSPAN");  // tuning: 1, 2 or 4 (default: as fine as the table allows)
        return e && *e ? atoi(e) : 0;
    }();
    for (;;) {
        uint64_t nxy = 1;
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            const double ext = (double)hi[a] - (double)lo[a];
            const double d = floor(ext / cell) + 1.0;
            if (!(d <= 4096.0)) { ok = false; break; }
            g.dim[a] = (int)d;
            if (a < 2) nxy *= (uint64_t)g.dim[a];
        }
        if (ok && nxy * (uint64_t)g.dim[2] <= (1ull << 24)) {
            // slices along z: as many as keep the cell table within 2^23 entries
            g.zspan = 1;
            for (int zs : {4, 2}) {
                if (zspan_env && zs != zspan_env) continue;
                const double dz = floor(((double)hi[2] - (double)lo[2]) / (cell / zs)) + 1.0;
                if (dz <= 16384.0 && (double)nxy * dz <= (double)(1u << 23)) {
                    g.zspan = zs;
                    g.dim[2] = (int)dz;
                    break;
                }
            }
            if (zspan_env == 1) g.zspan = 1;
            g.ncells = (uint32_t)(nxy * (uint64_t)g.dim[2]);
            break;
        }
        cell *= 1.25;
    }
    g.cell = (float)cell;
    g.inv_cell = 1.0f / g.cell;
    // z slices: an edge of cell / zspan, still exact -- two boids closer than `cell` along z are
    // at most zspan slices apart (the 1/512 margin of the edge covers the rounding of zspan / cell)
    g.inv_cell_z = (float)((double)g.zspan / cell);
    for (int a = 0; a < 3; ++a) g.origin[a] = lo[a];
    {
        // The grid is centred on the flock rather than anchored at its minimum corner.  When the extent
        // is a hair over a whole number of cells an anchored grid ends in a sliver row holding a handful
        // of boids spread over all of z; the CTA that holds them has nine intervals spanning whole rows,
        // overflows the tile and takes the slow global-memory path (C3 at skin 0.28: 53 of 8192 CTAs,
        // emulated on the CPU and seen on the GPU; 0 when centred: every edge cell is then at least half
        // a cell wide.  Measured: C3 0.261 -> 0.253 ms/step).  FP_GRID_CENTER=0 anchors it (tuning).
        static const bool center = [] {
            const char *e = getenv("FP_GRID_CENTER");
            return !(e && *e == '0');
        }();
        if (center)
            for (int a = 0; a < 3; ++a) {
                const double edge = a == 2 ? cell / g.zspan : cell;
                const double slack = (double)g.dim[a] * edge - ((double)hi[a] - (double)lo[a]);
                if (slack > 0.0) g.origin[a] = (float)((double)lo[a] - 0.5 * slack);
            }
    }
    uint32_t bits = 1;
    while ((1ull << bits) < g.ncells) ++bits;
    g.key_bits = bits;
    g.gdimx = g.dim[0];
    g.xoff = 0;
    g.skin = skin;
    f->grid = g;
    f->bin_valid = false;
    if (!f->work.ctl) {
        int rc = dev_alloc(&f->work.ctl, 1);
        if (rc) return rc;
        FP_CUDA(cudaMemsetAsync(f->work.ctl, 0, sizeof(SkinCtl), f->stream));
    }
    if (!f->h_ctl) FP_CUDA(cudaMallocHost((void **)&f->h_ctl, sizeof(SkinCtl)));
    if (f->shard) return shard_grid_fitted(f->shard, f);  // slab layout + scratch are the shard's

    // scratch
    GridWork &w = f->work;
    const uint32_t cap = f->n;  // (a sharded flock sizes its scratch in shard_grid_fitted)
    const size_t ntiles = ((size_t)cap + 4095) / 4096 + 1;
    const size_t hist = 256 * ntiles;
    const size_t scan_n = std::max(hist, (size_t)g.ncells + 1);
    const size_t scan_tmp = scan_n / 4096 + 2;
    if (cap + 8 > w.soa_cap) {
        for (auto &b : w.soa)
            for (auto &p : b) {
                dev_free(p);
                int rc = dev_alloc(&p, (size_t)cap + 8);
                if (rc) return rc;
                FP_CUDA(cudaMemsetAsync(p, 0, ((size_t)cap + 8) * sizeof(float), f->stream));
            }
        w.soa_cap = cap + 8;
    }
    if (cap > w.cap) {
        dev_free(w.keys[0]); dev_free(w.keys[1]); dev_free(w.vals[0]); dev_free(w.vals[1]);
        int rc;
        if ((rc = dev_alloc(&w.keys[0], cap)) || (rc = dev_alloc(&w.keys[1], cap)) ||
            (rc = dev_alloc(&w.vals[0], cap)) || (rc = dev_alloc(&w.vals[1], cap)))
            return rc;
        w.cap = cap;
    }
    if (hist > w.tile_hist_elems) {
        dev_free(w.tile_hist);
        int rc = dev_alloc(&w.tile_hist, hist);
        if (rc) return rc;
        w.tile_hist_elems = hist;
    }
    if ((size_t)g.ncells + 1 > w.cell_cap) {
        dev_free(w.cell_start);
        int rc = dev_alloc(&w.cell_start, (size_t)g.ncells + 1);
        if (rc) return rc;
        w.cell_cap = (size_t)g.ncells + 1;
    }
    if (scan_tmp > w.scan_tmp_elems) {
        dev_free(w.scan_tmp);
        int rc = dev_alloc(&w.scan_tmp, scan_tmp);
        if (rc) return rc;
        w.scan_tmp_elems = scan_tmp;
    }
    f->grid_valid = true;
    f->steps_since_fit = 0;
    return FP_OK;
}

int resolve_method(fp_flock *f) {
    int m = f->method;
    if (m == FP_METHOD_AUTO) {
        const float reach = reach_of(f->cfg);
        const bool grid_ok = reach > 0.0f && std::isfinite(reach);
        const uint64_t ng = f->shard ? f->n_global : f->n;
        if (ng <= small_max_boids() && !f->shard) m = FP_METHOD_SMALL;
        else if (ng >= 32768 && grid_ok) m = FP_METHOD_GRID;
        else m = FP_METHOD_ALLPAIRS;
    }
    if (m == FP_METHOD_SMALL && (f->n > small_max_boids() || f->shard)) m = FP_METHOD_ALLPAIRS;
    f->method_in_use = m;
    return m;
}

// bring the state back to caller index order (all-pairs sums in that order)
int ensure_caller_order(fp_flock *f) {
    if (!f->permuted) return FP_OK;
    int rc = launch_unpermute(f->stream, f->pos[f->cur], f->vel[f->cur], f->pos[f->cur ^ 1],
                              f->vel[f->cur ^ 1], f->n, f->first_index);
    if (rc) return rc;
    f->cur ^= 1;
    f->permuted = false;
    f->bin_valid = false;
    return FP_OK;
}

// the lead rows of the step about to be enqueued -> f->P; returns the version it runs with
// (a replayed step takes the version it was first enqueued under)
uint32_t select_leads(fp_flock *f) {
    uint32_t ver = f->lead_ver;
    if (!f->replay_lead_vers.empty()) {
        ver = f->replay_lead_vers.front();
        f->replay_lead_vers.erase(f->replay_lead_vers.begin());
    }
    if (f->d_lead_table && f->table_rows) {
        const uint32_t row = std::min(f->table_cursor, f->table_rows - 1);
        f->P.leads = f->d_lead_table + (size_t)row * f->table_leads * 8;
        f->P.n_leads = (int)f->table_leads;
    } else {
        f->P.leads = f->n_leads ? f->d_lead_ring + (size_t)(ver % fp_flock::LEAD_RING) * f->n_leads * 8 : nullptr;
        f->P.n_leads = (int)f->n_leads;
    }
    return ver;
}

bool refit_due(const fp_flock *f) {
    return !f->grid_valid || (!f->domain_user && f->steps_since_fit >= 256);
}

// Bin the current state: sort it by cell key.  The sorted copy becomes the state (same
// boids, new order), with its SoA positions, home keys and cell table.  Enqueues only.
int grid_rebin(fp_flock *f) {
    GridWork &w = f->work;
    int rc = launch_skin_gate(f->stream, w.ctl, f->ordinal, 1, f->P.dt, f->skin_budget);
    if (rc) return rc;
    if ((rc = launch_grid_keys(f->stream, f->grid, f->pos[f->cur], f->n, w))) return rc;
    int buf = 0;
    if ((rc = launch_radix_sort(f->stream, w, f->n, f->grid.key_bits, &buf))) return rc;
    if ((rc = launch_grid_reorder(f->stream, w.vals[buf], f->pos[f->cur], f->vel[f->cur], f->pos[f->cur ^ 1],
                                  f->vel[f->cur ^ 1], w.soa[w.soa_cur ^ 1], f->n, w.ctl)))
        return rc;
    f->cur ^= 1;
    w.soa_cur ^= 1;
    w.home = w.keys[buf];
    f->permuted = true;
    f->bin_valid = true;
    f->plan_left = plan_steps(f, 0.0f, 0.0f);
    ++f->stat_rebins;
    nl_binned(f);
    return FP_OK;
}

WalkIO walk_io(fp_flock *f, bool stepping) {
    GridWork &w = f->work;
    WalkIO io{};
    io.pos_s = f->pos[f->cur];
    io.vel_s = f->vel[f->cur];
    for (int a = 0; a < 3; ++a) io.soa_in[a] = w.soa[w.soa_cur][a];
    io.home = w.home;
    io.cell_start = w.cell_start;
    io.first = 0;
    io.last = f->n;
    if (stepping) {
        io.pos_out = f->pos[f->cur ^ 1];
        io.vel_out = f->vel[f->cur ^ 1];
        for (int a = 0; a < 3; ++a) io.soa_out[a] = w.soa[w.soa_cur ^ 1][a];
        io.ctl = w.ctl;
    }
    return io;
}

// ---- standing candidate lists (fp_walk_nl.cu) ------------------------------------------------
// The default form of a grid step, single GPU and sharded alike: FP_NL=0 turns them off (plain
// staged walk every step; same bits under EXACT numerics).
constexpr uint32_t NL_VCAP = 96;  // C3 / C4 density: 34 candidates per boid on average, ~70 at most

bool nl_enabled() {
    static const bool on = [] {
        const char *e = getenv("FP_NL");
        return !(e && *e == '0');
    }();
    return on;
}
bool nl_trace() {
    static const bool on = getenv("FP_NL_TRACE") != nullptr;
    return on;
}

bool nl_wanted(const fp_flock *f) {
    return nl_enabled() && !f->nl_off && f->grid.skin > 0.0f;  // (no skin = a binning per step: nothing to re-use)
}

void nl_free(fp_flock *f) {
    dev_free(f->nl_entries);
    dev_free(f->nl_count);
    dev_free(f->nl_cta_tab);
    dev_free(f->nl_flag);
    f->nl_rows = 0;
    f->nl_serial = ~0ull;
}

int nl_ensure(fp_flock *f, uint32_t rows) {
    if (f->nl_entries && f->nl_rows >= rows) return FP_OK;
    nl_free(f);
    // a slab's owned count changes from binning to binning: leave room so that it rarely re-allocates
    const uint32_t cap = f->shard ? rows + rows / 8 + 4096 : rows;
    int rc;
    const size_t entries = nl_entries_elems(cap, NL_VCAP);
    if ((rc = dev_alloc(&f->nl_entries, entries)) || (rc = dev_alloc(&f->nl_count, (size_t)cap + 128)) ||
        (rc = dev_alloc(&f->nl_cta_tab, nl_cta_tab_elems(cap))) || (rc = dev_alloc(&f->nl_flag, 4)))
        return rc;
    FP_CUDA(cudaMemsetAsync(f->nl_entries, 0, entries * sizeof(uint16_t), f->stream));
    FP_CUDA(cudaMemsetAsync(f->nl_flag, 0, sizeof(unsigned), f->stream));
    f->nl_rows = cap;
    if (nl_trace()) fprintf(stderr, "fp: candidate lists on (%u boids, %u entries each)\n", rows, NL_VCAP);
    return FP_OK;
}

// Before a new build: how did the last one go?  CTAs without lists walk from global memory, which
// is far slower than the staged walk; when more than 1 in 32 had none (and more than a handful),
// the flock is too dense for lists of this size and the staged walk takes over (until a new
// state / config arrives).
int nl_review(fp_flock *f) {
    if (!f->nl_flag) return FP_OK;
    unsigned none = 0;
    FP_CUDA(cudaMemcpyAsync(&none, f->nl_flag, sizeof(none), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));  // (a binning has just settled: the stream is all but idle)
    FP_CUDA(cudaMemsetAsync(f->nl_flag, 0, sizeof(none), f->stream));
    if (nl_trace() && none) fprintf(stderr, "fp: %u CTAs without candidate lists in the last build\n", none);
    if (none > 8 && (uint64_t)none * 32u > ((uint64_t)f->nl_built_rows + 127u) / 128u) {
        if (nl_trace()) fprintf(stderr, "fp: candidate lists off\n");
        f->nl_off = true;
    }
    return FP_OK;
}

NlIO nl_io(const fp_flock *f, double skins = 1.0) {
    NlIO nl{};
    nl.entries = f->nl_entries;
    nl.count = f->nl_count;
    nl.cta_tab = f->nl_cta_tab;
    nl.flag = f->nl_flag;
    nl.vcap = NL_VCAP;
    nl.tile_cap = nl_tile_cap(f->P.numerics_fast != 0);
    nl.tile_cap_b = nl_tile_cap_b(f->P.numerics_fast != 0);
    // every pair within reach while the binning stands was within reach + skin when it was made
    // (skins = 2: within reach + 2 skin at any other moment of the binning's life)
    const double R = (double)reach_of(f->cfg) + skins * (double)f->grid.skin;
    nl.m2_wide = nextafterf((float)(R * R * (1.0 + 1e-5)), INFINITY);
    // the fast walk sums in list order whatever it is: its lists put the entries in view first
    nl.vis_first = f->P.numerics_fast && f->P.fz_a > -2.5f && f->P.fz_a < 1.0f ? 1 : 0;
    nl.vis_c = f->P.fz_a;
    return nl;
}

// lists for the standing binning, from the positions as they are now
int nl_build_now(fp_flock *f, const GridDesc &g, const WalkIO &io, double skins) {
    int rc = nl_review(f);  // (may turn the lists off)
    if (rc || !nl_wanted(f)) return rc;
    const uint32_t rows = io.last - io.first;
    if ((rc = nl_ensure(f, rows)) || (rc = launch_nl_build(f->stream, g, io, nl_io(f, skins)))) return rc;
    f->nl_serial = f->stat_rebins;
    f->nl_built_rows = rows;
    return FP_OK;
}

// Called once per step, after the binning / gate and before the walk.  Lists are built only for
// binnings that live: a caller that hands over a new state every step bins every step, and a build
// plus a list walk is slower than the staged walk alone.  At once when the previous binning served
// >= 8 steps; otherwise at the binning's SECOND step, from the positions of that moment -- every
// boid is within skin / 2 of its binned position throughout, hence within skin of where it stands
// at the build: the cut becomes reach + 2 skin.
int nl_prepare(fp_flock *f, const GridDesc &g, const WalkIO &io) {
    if (!nl_wanted(f) || f->nl_serial == f->stat_rebins) return FP_OK;
    double skins = 1.0;
    if (f->nl_fresh) {
        if (f->nl_prev_bin_steps < 8) return FP_OK;  // short-lived so far: see whether a second step comes
    } else if (f->nl_bin_steps == 1) {
        skins = 2.0;
    } else {
        return FP_OK;
    }
    return nl_build_now(f, g, io, skins);
}

// the step's walk: on the lists when they describe the standing binning, else the staged walk
int nl_or_plain_walk(fp_flock *f, const GridDesc &g, const WalkIO &io) {
    f->nl_fresh = false;
    ++f->nl_bin_steps;
    if (nl_wanted(f) && f->nl_serial == f->stat_rebins)
        return launch_nl_walk(f->stream, f->P, g, TAP_STEP, io, nl_io(f), f->d_status, TapOut{});
    return launch_grid_walk(f->stream, f->P, g, TAP_STEP, io, f->d_status, TapOut{});
}

// a tap's walk over a flock that has just been binned.  The acceleration tap of a FAST-numerics
// flock runs the arithmetic its steps run (lists built on the spot), so that the parity checks
// measure what is shipped; everything else is evaluated by the exact kernels.
int tap_walk(fp_flock *f, const GridDesc &g, int tap, const WalkIO &io, const TapOut &out) {
    if (tap == TAP_ACCEL && f->P.numerics_fast && nl_wanted(f)) {
        int rc = FP_OK;
        if (f->nl_serial != f->stat_rebins && (rc = nl_build_now(f, g, io, 1.0))) return rc;
        if (nl_wanted(f) && f->nl_serial == f->stat_rebins)
            return launch_nl_walk(f->stream, f->P, g, TAP_ACCEL, io, nl_io(f), f->d_status, out);
    }
    return launch_grid_walk(f->stream, f->P, g, tap, io, f->d_status, out);
}

// timing hook: record the next pooled event on the stream (no-op unless timing)
int mark_event(fp_flock *f) {
    if (!f->timing) return FP_OK;
    if (f->ev_used == f->ev_pool.size()) {
        cudaEvent_t e;
        FP_CUDA(cudaEventCreate(&e));
        f->ev_pool.push_back(e);
    }
    FP_CUDA(cudaEventRecord(f->ev_pool[f->ev_used++], f->stream));
    return FP_OK;
}

// Enqueue `nsteps` grid steps (single GPU).  A step re-bins first when there is no valid
// binning or the plan says the skin is used up; otherwise it walks the standing binning.
// Every step is logged as pending until settle() has seen that the device performed it.
int grid_steps(fp_flock *f, uint32_t nsteps) {
    int rc;
    for (uint32_t s = 0; s < nsteps; ++s) {
        if (f->pending.size() >= 256 && (rc = settle(f))) return rc;
        if (refit_due(f)) {
            if ((rc = settle(f)) || (rc = fit_grid(f))) return rc;  // the fit reads the positions
        }
        // a binning never runs inside a window the device may have voided: settle first (one
        // host sync per binning, i.e. every few dozen steps; it also refreshes the plan)
        if ((!f->bin_valid || f->plan_left <= 0) && (rc = settle(f))) return rc;
        const uint32_t lead_ver = select_leads(f);
        if ((rc = mark_event(f))) return rc;
        f->pending.push_back({f->ordinal, f->cur, f->work.soa_cur, f->table_cursor, f->steps_since_fit, 0u, lead_ver});
        if (!f->bin_valid || f->plan_left <= 0) {
            if ((rc = grid_rebin(f))) return rc;
        } else if ((rc = launch_skin_gate(f->stream, f->work.ctl, f->ordinal, 0, f->P.dt, f->skin_budget))) {
            return rc;
        }
        if ((rc = nl_prepare(f, f->grid, walk_io(f, true)))) return rc;  // (its build counts as sort phase)
        if ((rc = mark_event(f))) return rc;
        if ((rc = nl_or_plain_walk(f, f->grid, walk_io(f, true)))) return rc;
        f->cur ^= 1;
        f->work.soa_cur ^= 1;
        if ((rc = mark_event(f))) return rc;
        --f->plan_left;
        ++f->ordinal;
        ++f->steps_since_fit;
        ++f->table_cursor;
        ++f->stat_grid_steps;
    }
    return FP_OK;
}

// Make every enqueued step a fact.  If the device found the skin used up before a step (the
// flock got faster than planned), that step and all later ones were no-ops: restore the host
// view to just before it, re-bin and enqueue them again.
int settle(fp_flock *f) {
    if (f->shard) return shard_settle(f->shard, f);
    int rc;
    while (!f->pending.empty()) {
        FP_CUDA(cudaMemcpyAsync(f->h_ctl, f->work.ctl, sizeof(SkinCtl), cudaMemcpyDeviceToHost, f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
        const SkinCtl c = *f->h_ctl;
        if (!c.stale) {
            // refresh the plan from what the device measured
            float v2, pm;
            memcpy(&v2, &c.v2max, 4);
            memcpy(&pm, &c.pmax, 4);
            f->delta_est = plan_delta(v2, pm, f->cfg.dt);
            const float last = sqrtf(v2) * fabsf(f->cfg.dt) * 1.000001f + pm * 2.1e-7f + 1e-30f;
            if (f->bin_valid) f->plan_left = std::min(f->plan_left, plan_steps(f, c.D, last));
            f->pending.clear();
            break;
        }
        size_t k = 0;
        while (k < f->pending.size() && f->pending[k].ordinal != c.first_stale) ++k;
        if (k == f->pending.size()) {
            set_error("internal: stale step not in the pending log");
            return FP_ERR_INVALID;
        }
        const fp_flock::Pending at = f->pending[k];
        const uint32_t redo = (uint32_t)(f->pending.size() - k);
        {   // versions of the steps to redo, ahead of whatever an enclosing replay still has to enqueue
            std::vector<uint32_t> vers;
            for (size_t q = k; q < f->pending.size(); ++q) vers.push_back(f->pending[q].lead_ver);
            vers.insert(vers.end(), f->replay_lead_vers.begin(), f->replay_lead_vers.end());
            f->replay_lead_vers.swap(vers);
        }
        f->pending.clear();
        f->cur = at.cur;
        f->work.soa_cur = at.soa_cur;
        f->table_cursor = at.table_cursor;
        f->steps_since_fit = at.steps_since_fit;
        f->bin_valid = false;
        f->stat_replayed += redo;
        f->stat_grid_steps -= redo;
        FP_CUDA(cudaMemsetAsync(f->work.ctl, 0, sizeof(SkinCtl), f->stream));
        // the fastest boid outran the plan: let the next fit size the skin for it
        if (!f->domain_user) f->grid_valid = false;
        if ((rc = grid_steps(f, redo))) return rc;
    }
    return FP_OK;
}

int run_tap(fp_flock *f, int tap, const TapOut &out) {
    select_leads(f);
    if (f->shard) return shard_tap(f->shard, f, tap, out);
    const int m = resolve_method(f);
    if (m == FP_METHOD_GRID) {
        // (the caller has settled.)  Taps always bin the state where it stands: the listing they
        // report against is the cell-sorted order of the current positions.
        int rc;
        if (refit_due(f) && (rc = fit_grid(f))) return rc;
        if ((rc = grid_rebin(f))) return rc;
        return tap_walk(f, f->grid, tap, walk_io(f, false), out);
    }
    int rc = ensure_caller_order(f);
    if (rc) return rc;
    if (f->P.numerics_fast && (rc = launch_bounds(f->stream, f->pos[f->cur], nullptr, f->n, f->d_bounds))) return rc;
    return launch_allpairs(f->stream, f->P, tap, f->pos[f->cur], f->vel[f->cur], f->n, 0, f->n, nullptr,
                           nullptr, f->d_status, out, 0, f->d_bounds);
}

}  // namespace

// entry points fp_shard.cu needs from this file
namespace fp {
int flock_fit_grid(fp_flock *f) { return fit_grid(f); }
int flock_nl_prepare(fp_flock *f, const GridDesc &g, const WalkIO &io) { return nl_prepare(f, g, io); }
void flock_nl_binned(fp_flock *f) { nl_binned(f); }
int flock_step_walk(fp_flock *f, const GridDesc &g, const WalkIO &io) { return nl_or_plain_walk(f, g, io); }
int flock_tap_walk(fp_flock *f, const GridDesc &g, int tap, const WalkIO &io, const TapOut &out) {
    return tap_walk(f, g, tap, io, out);
}
uint32_t flock_select_leads(fp_flock *f) { return select_leads(f); }
int64_t flock_plan_steps(const fp_flock *f, float D, float first_delta) { return plan_steps(f, D, first_delta); }
float flock_plan_delta(float v2max, float pmax, float dt) { return plan_delta(v2max, pmax, dt); }
}  // namespace fp

namespace fp {
int flock_allpairs_step(fp_flock *f, const float4 *pos_all, const float4 *vel_all, uint32_t n_all, uint32_t row0,
                        uint32_t nrows, float4 *pos_out, float4 *vel_out) {
    if (f->P.numerics_fast) {  // one kernel; its pre-gate needs the flock's bounds (device side, no host sync)
        int rc = launch_bounds(f->stream, pos_all, nullptr, n_all, f->d_bounds);
        if (rc) return rc;
        return launch_allpairs(f->stream, f->P, TAP_STEP, pos_all, vel_all, n_all, row0, nrows, pos_out, vel_out,
                               f->d_status, TapOut{}, 0, f->d_bounds);
    }
    if (f->ap_choice >= 0 && ++f->ap_age >= 1024) {  // the flock may have changed density: measure again
        f->ap_choice = -1;
        f->ap_probe = 0;
    }
    int variant = f->ap_choice;
    int probing = -1;
    if (variant < 0) {
        if (f->ap_probe == 2) {  // both timings are in flight: wait for them once and decide
            float t0 = 0, t1 = 0;
            FP_CUDA(cudaEventSynchronize(f->ap_ev[3]));
            FP_CUDA(cudaEventElapsedTime(&t0, f->ap_ev[0], f->ap_ev[1]));
            FP_CUDA(cudaEventElapsedTime(&t1, f->ap_ev[2], f->ap_ev[3]));
            f->ap_choice = variant = t1 < t0 ? 1 : 0;
            f->ap_age = 0;
        } else {
            probing = variant = (int)f->ap_probe;
            for (auto &e : f->ap_ev)
                if (!e) FP_CUDA(cudaEventCreate(&e));
            FP_CUDA(cudaEventRecord(f->ap_ev[2 * probing], f->stream));
        }
    }
    int rc = launch_allpairs(f->stream, f->P, TAP_STEP, pos_all, vel_all, n_all, row0, nrows, pos_out, vel_out,
                             f->d_status, TapOut{}, variant);
    if (rc) return rc;
    if (probing >= 0) {
        FP_CUDA(cudaEventRecord(f->ap_ev[2 * probing + 1], f->stream));
        ++f->ap_probe;
    }
    return FP_OK;
}
}  // namespace fp

// ---- C ABI ----------------------------------------------------------------------
extern "C" {

const char *fp_last_error(void) { return t_error.c_str(); }
const char *fp_version(void) { return "feriphys-cuda 0.1.0 sm_100a"; }
uint64_t fp_launch_count(void) { return g_launches.load(); }

int fp_config_default(fp_config *cfg) {
    if (!cfg) { set_error("null config"); return FP_ERR_INVALID; }
    // Duration::from_millis(1).as_secs_f32() = 0f32 + 1_000_000f32 / 1e9f32
    cfg->dt = 0.0f + 1000000.0f / 1000000000.0f;
    cfg->avoidance_factor = 1.0f;
    cfg->centering_factor = 0.1f;
    cfg->velocity_matching_factor = 0.5f;
    cfg->distance_weight_threshold = 15.0f;
    cfg->distance_weight_threshold_falloff = 1.0f;
    cfg->max_sight_angle = 3.14159274101257324f / 2.0f;
    cfg->max_sight_angle_to_lead_boid = 3.14159274101257324f;
    cfg->time_to_start_steering_secs = 4;
    cfg->time_to_start_steering_nanos = 0;
    cfg->steering_overrides = 0;
    return FP_OK;
}

static int create_common(fp_flock **out, const fp_config *cfg, uint64_t n_global, uint64_t first,
                         uint64_t n_local, const float *state, int device, uint32_t cap) {
    if (!out) { set_error("null out pointer"); return FP_ERR_INVALID; }
    *out = nullptr;
    if (n_global >= (1ull << 31) || (n_local && !state)) {
        set_error("flock size must be < 2^31 and state must be non-null");
        return FP_ERR_INVALID;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        set_error(std::string("no usable CUDA device ") + std::to_string(device) + ": " +
                  (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range") +
                  " -- this library has no CPU fallback");
        return FP_ERR_CUDA;
    }
    FP_CUDA(cudaSetDevice(device));
    fp_flock *f = new (std::nothrow) fp_flock();
    if (!f) { set_error("out of host memory"); return FP_ERR_INVALID; }
    f->device = device;
    f->n = (uint32_t)n_local;
    f->cap = std::max<uint32_t>(cap, (uint32_t)n_local);
    f->n_global = n_global;
    f->first_index = (uint32_t)first;
    if (cfg) f->cfg = *cfg; else fp_config_default(&f->cfg);
    {
        static const int env_numerics = [] {  // FP_NUMERICS=fast|exact: the default of new handles (tuning)
            const char *e = getenv("FP_NUMERICS");
            return e && (*e == 'f' || *e == 'F' || *e == '1') ? FP_NUMERICS_FAST : FP_NUMERICS_EXACT;
        }();
        f->numerics = env_numerics;
    }
    derive_params(f->cfg, f->P, f->numerics == FP_NUMERICS_FAST);
    int rc = FP_OK;
    auto fail = [&](int code) { fp_flock_destroy(f); return code; };
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaStreamCreate", __FILE__, __LINE__));
    for (int b = 0; b < 2; ++b) {
        if ((rc = dev_alloc(&f->pos[b], f->cap)) || (rc = dev_alloc(&f->vel[b], f->cap))) return fail(rc);
    }
    if ((rc = dev_alloc(&f->d_status, 1)) || (rc = dev_alloc(&f->d_census, 4)) ||
        (rc = dev_alloc(&f->d_bounds, 8)))
        return fail(rc);
    cudaMemsetAsync(f->d_status, 0, sizeof(unsigned), f->stream);
    refresh_tables(f);
    if (n_local) {
        if ((rc = ensure_stage(f, n_local * 6 * sizeof(float)))) return fail(rc);
        if (cudaMemcpyAsync(f->d_stage, state, n_local * 6 * sizeof(float), cudaMemcpyHostToDevice,
                            f->stream) != cudaSuccess)
            return fail(cuda_fail(cudaGetLastError(), "cudaMemcpyAsync", __FILE__, __LINE__));
        if ((rc = launch_aos6_to_soa(f->stream, (const float *)f->d_stage, f->pos[0], f->vel[0], f->n,
                                     f->first_index)))
            return fail(rc);
    }
    if (cudaStreamSynchronize(f->stream) != cudaSuccess)
        return fail(cuda_fail(cudaGetLastError(), "cudaStreamSynchronize", __FILE__, __LINE__));
    *out = f;
    return FP_OK;
}

int fp_flock_create(fp_flock **out, const fp_config *cfg, uint64_t n, const float *state_aos6,
                    int device) {
    return create_common(out, cfg, n, 0, n, state_aos6, device, (uint32_t)n);
}

int fp_flock_create_sharded(fp_flock **out, const fp_config *cfg, uint64_t n_global,
                            uint64_t first_index, uint64_t n_local, const float *state_aos6,
                            int device, int rank, int world, const uint8_t nccl_unique_id[128]) {
    if (world < 1 || rank < 0 || rank >= world || first_index + n_local > n_global) {
        set_error("bad rank/world/first_index");
        return FP_ERR_INVALID;
    }
    if (world == 1) return create_common(out, cfg, n_global, 0, n_local, state_aos6, device,
                                         (uint32_t)n_local);
    // slabs are unbalanced by nature: leave head-room for migration and halos
    // (at least 2 MB per buffer: cudaIpc exports whole allocations, small ones share blocks)
    const uint64_t cap64 = std::max<uint64_t>(
        std::min<uint64_t>(n_global, (n_global / world) * 3 / 2 + (1u << 16)), 1u << 19);
    int rc = create_common(out, cfg, n_global, first_index, n_local, state_aos6, device,
                           (uint32_t)std::max<uint64_t>(cap64, n_local));
    if (rc) return rc;
    rc = shard_create(&(*out)->shard, *out, rank, world, nccl_unique_id);
    if (rc) {
        fp_flock_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

int fp_flock_destroy(fp_flock *f) {
    if (!f) return FP_OK;
    cudaSetDevice(f->device);
    if (f->stream) cudaStreamSynchronize(f->stream);
    if (f->shard) shard_destroy(f->shard);
    for (int b = 0; b < 2; ++b) { dev_free(f->pos[b]); dev_free(f->vel[b]); }
    dev_free(f->d_lead_ring); dev_free(f->d_attr); dev_free(f->d_obs); dev_free(f->d_lead_table);
    if (f->h_lead_stage) cudaFreeHost(f->h_lead_stage);
    for (auto &ev : f->lead_ev) if (ev) cudaEventDestroy(ev);
    dev_free(f->d_status); dev_free(f->d_census); dev_free(f->d_bounds);
    free_grid_work(f);
    nl_free(f);
    if (f->d_stage) cudaFree(f->d_stage);
    for (int k = 0; k < 2; ++k) {
        if (f->h_up[k]) cudaFreeHost(f->h_up[k]);
        if (k == 0 && f->h_out) cudaFreeHost(f->h_out);
        if (f->up_ev[k]) cudaEventDestroy(f->up_ev[k]);
    }
    if (f->h_ctl) cudaFreeHost(f->h_ctl);
    for (auto &ev : f->ev_pool) if (ev) cudaEventDestroy(ev);
    for (auto &ev : f->ap_ev) if (ev) cudaEventDestroy(ev);
    if (f->stream) cudaStreamDestroy(f->stream);
    delete f;
    return FP_OK;
}

uint64_t fp_flock_len(const fp_flock *f) { return f ? (f->shard ? f->n_global : f->n) : 0; }

int fp_flock_set_config(fp_flock *f, const fp_config *cfg) {
    int rc = check(f);
    if (rc) return rc;
    if (!cfg) { set_error("null config"); return FP_ERR_INVALID; }
    const float old_reach = reach_of(f->cfg);
    f->cfg = *cfg;
    derive_params(f->cfg, f->P, f->numerics == FP_NUMERICS_FAST);
    if (reach_of(f->cfg) != old_reach) f->grid_valid = false;
    f->bin_valid = false;  // dt and reach enter the skin accounting
    f->nl_off = false;
    f->ap_choice = -1;     // a new reach changes which all-pairs kernel wins
    f->ap_probe = 0;
    return FP_OK;
}

int fp_flock_get_config(fp_flock *f, fp_config *cfg) {
    if (!f || !cfg) { set_error("null argument"); return FP_ERR_INVALID; }
    *cfg = f->cfg;
    return FP_OK;
}

int fp_flock_set_method(fp_flock *f, int method) {
    if (!f || method < FP_METHOD_AUTO || method > FP_METHOD_SMALL) {
        set_error("bad method");
        return FP_ERR_INVALID;
    }
    int rc = check(f);
    if (rc) return rc;
    f->method = method;
    return FP_OK;
}

int fp_flock_set_numerics(fp_flock *f, int numerics) {
    if (!f || (numerics != FP_NUMERICS_EXACT && numerics != FP_NUMERICS_FAST)) {
        set_error("bad numerics");
        return FP_ERR_INVALID;
    }
    int rc = check(f);
    if (rc) return rc;
    if (numerics == f->numerics) return FP_OK;
    f->numerics = numerics;
    derive_params(f->cfg, f->P, numerics == FP_NUMERICS_FAST);
    f->bin_valid = false;  // the lists on hand were sized for the other walk's tile
    f->nl_serial = ~0ull;
    f->ap_choice = -1;
    f->ap_probe = 0;
    return FP_OK;
}

int fp_flock_get_numerics(fp_flock *f, int *numerics, int *in_use) {
    if (!f) { set_error("null flock handle"); return FP_ERR_INVALID; }
    if (numerics) *numerics = f->numerics;
    if (in_use) *in_use = f->P.numerics_fast ? FP_NUMERICS_FAST : FP_NUMERICS_EXACT;
    return FP_OK;
}

int fp_flock_get_method(fp_flock *f, int *method_in_use) {
    if (!f || !method_in_use) { set_error("null argument"); return FP_ERR_INVALID; }
    if (f->shard) *method_in_use = shard_method(f->shard, f->method, f->cfg);
    else *method_in_use = resolve_method(f);
    return FP_OK;
}

int fp_flock_set_leads(fp_flock *f, uint32_t n_leads, const float *leads7) {
    // (no settle: the steps in flight keep the rows they were enqueued with -- see fp_flock.h;
    //  the boid state is not touched: rows pending in a pinned slot / mapped output rows stay)
    int rc = check(f, false, true);
    if (rc) return rc;
    if (n_leads && !leads7) { set_error("null leads"); return FP_ERR_INVALID; }
    constexpr uint32_t RING = fp_flock::LEAD_RING, STAGES = fp_flock::LEAD_STAGES;
    const size_t stride = (size_t)n_leads * 8;
    const bool relayout = n_leads != f->n_leads || (n_leads && !f->d_lead_ring) || f->d_lead_table;
    if (relayout || (!f->pending.empty() && f->lead_ver + 1 - f->pending.front().lead_ver >= RING - 1)) {
        if ((rc = settle(f))) return rc;  // nothing in flight refers to the rows any more
    }
    if (relayout) {
        FP_CUDA(cudaStreamSynchronize(f->stream));
        dev_free(f->d_lead_ring);
        if (f->h_lead_stage) cudaFreeHost(f->h_lead_stage);
        f->h_lead_stage = nullptr;
        if (n_leads) {
            if ((rc = dev_alloc(&f->d_lead_ring, stride * RING))) return rc;
            FP_CUDA(cudaMallocHost((void **)&f->h_lead_stage, stride * STAGES * sizeof(float)));
            for (auto &e : f->lead_ev)
                if (!e) FP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        dev_free(f->d_lead_table);  // a plain table replaces any per-step table
        f->table_rows = f->table_leads = f->table_cursor = 0;
        f->n_leads = n_leads;
    }
    if (n_leads) {
        ++f->lead_ver;
        const uint32_t st = f->lead_stage_cur++ % STAGES;
        FP_CUDA(cudaEventSynchronize(f->lead_ev[st]));  // the copy that last used this staging slot
        float *h = f->h_lead_stage + (size_t)st * stride;
        for (uint32_t k = 0; k < n_leads; ++k) {
            memcpy(h + 8 * k, leads7 + 7 * k, 7 * sizeof(float));
            h[8 * k + 7] = 0.0f;
        }
        FP_CUDA(cudaMemcpyAsync(f->d_lead_ring + (size_t)(f->lead_ver % RING) * stride, h, stride * sizeof(float),
                                cudaMemcpyHostToDevice, f->stream));
        FP_CUDA(cudaEventRecord(f->lead_ev[st], f->stream));
    }
    f->d_leads = n_leads ? f->d_lead_ring + (size_t)(f->lead_ver % RING) * stride : nullptr;
    refresh_tables(f);
    return FP_OK;
}

int fp_flock_set_lead_table(fp_flock *f, uint32_t steps, uint32_t n_leads, const float *table7) {
    int rc = check(f);
    if (rc) return rc;
    if (steps && n_leads && !table7) { set_error("null lead table"); return FP_ERR_INVALID; }
    const size_t rows = (size_t)steps * n_leads;
    std::vector<float> padded(rows * 8, 0.0f);
    for (size_t k = 0; k < rows; ++k) memcpy(&padded[8 * k], table7 + 7 * k, 7 * sizeof(float));
    dev_free(f->d_lead_table);
    size_t table_cap = 0;
    rc = upload_table(f, &f->d_lead_table, &table_cap, padded.data(), padded.size());
    if (rc) return rc;
    f->table_rows = n_leads ? steps : 0;
    f->table_leads = n_leads;
    f->table_cursor = 0;
    return FP_OK;
}

int fp_flock_set_attractors(fp_flock *f, uint32_t n, const float *a4) {
    int rc = check(f);
    if (rc) return rc;
    if (n && !a4) { set_error("null attractors"); return FP_ERR_INVALID; }
    rc = upload_table(f, &f->d_attr, &f->attr_cap, a4, (size_t)n * 4);
    if (rc) return rc;
    f->n_attr = n;
    refresh_tables(f);
    return FP_OK;
}

int fp_flock_set_obstacles(fp_flock *f, uint32_t n, const float *o4) {
    int rc = check(f);
    if (rc) return rc;
    if (n && !o4) { set_error("null obstacles"); return FP_ERR_INVALID; }
    rc = upload_table(f, &f->d_obs, &f->obs_cap, o4, (size_t)n * 4);
    if (rc) return rc;
    f->n_obs = n;
    refresh_tables(f);
    return FP_OK;
}

int fp_flock_set_bbox(fp_flock *f, const float *bbox6) {
    int rc = check(f);
    if (rc) return rc;
    f->P.has_bbox = bbox6 ? 1 : 0;
    if (bbox6) memcpy(f->P.bbox, bbox6, 6 * sizeof(float));
    return FP_OK;
}

int fp_flock_set_grid_domain(fp_flock *f, const float lo3[3], const float hi3[3]) {
    int rc = check(f);
    if (rc) return rc;
    if (!lo3 || !hi3) {
        f->domain_user = false;
    } else {
        for (int a = 0; a < 3; ++a)
            if (!(lo3[a] <= hi3[a]) || !std::isfinite(lo3[a]) || !std::isfinite(hi3[a])) {
                set_error("grid domain must be finite with lo <= hi");
                return FP_ERR_INVALID;
            }
        memcpy(f->user_lo, lo3, sizeof(f->user_lo));
        memcpy(f->user_hi, hi3, sizeof(f->user_hi));
        f->domain_user = true;
    }
    f->grid_valid = false;
    return FP_OK;
}

int fp_flock_grid_info(fp_flock *f, uint32_t dims3[3], float *cell_size, uint32_t *key_bits) {
    int rc = check(f);
    if (rc) return rc;
    if (!f->grid_valid && (rc = fit_grid(f))) return rc;
    if (dims3) for (int a = 0; a < 3; ++a) dims3[a] = (uint32_t)f->grid.dim[a];
    if (cell_size) *cell_size = f->grid.cell;
    if (key_bits) *key_bits = f->grid.key_bits;
    return FP_OK;
}

int fp_flock_set_rebin(fp_flock *f, float skin, float plan_scale) {
    int rc = check(f);
    if (rc) return rc;
    if (!(plan_scale > 0.0f) || std::isnan(skin)) { set_error("bad re-binning policy"); return FP_ERR_INVALID; }
    f->skin_override = skin;
    f->plan_scale = plan_scale;
    f->grid_valid = false;  // the skin is part of the cell edge
    f->bin_valid = false;
    return FP_OK;
}

int fp_flock_rebin_info(fp_flock *f, float *skin, uint64_t *grid_steps, uint64_t *rebins, uint64_t *replayed) {
    int rc = check(f);
    if (rc) return rc;
    if (skin) *skin = f->grid_valid ? f->grid.skin : 0.0f;
    if (grid_steps) *grid_steps = f->stat_grid_steps;
    if (rebins) *rebins = f->stat_rebins;
    if (replayed) *replayed = f->stat_replayed;
    return FP_OK;
}

static int mark(fp_flock *f) { return mark_event(f); }

}  // extern "C"
namespace fp {
int flock_mark(fp_flock *f) { return mark_event(f); }
int flock_settle_local(fp_flock *f);
}
extern "C" {

int fp_flock_step(fp_flock *f, uint32_t nsteps) {
    // grid steps keep running ahead of the host: no settle here (grid_steps / shard_step do it
    // when they need to); the other methods start from a settled state
    int rc = check(f, false, true);
    if (rc) return rc;
    if (nsteps == 0) return FP_OK;
    f->out_valid = false;
    if (f->timing) f->timed_steps += nsteps;
    const int m = f->shard ? FP_METHOD_AUTO : resolve_method(f);
    if (m != FP_METHOD_SMALL && (rc = ingest_pending(f))) return rc;
    if (f->shard) return shard_step(f->shard, f, nsteps);
    if (f->n == 0) return FP_OK;
    if (m == FP_METHOD_GRID) return grid_steps(f, nsteps);
    if ((rc = settle(f))) return rc;
    if (m == FP_METHOD_SMALL) {
        rc = ensure_caller_order(f);
        if (rc) return rc;
        const bool table = f->d_lead_table && f->table_rows;
        const float *rows = nullptr;
        uint32_t nrows = 0;
        DevParams P = f->P;
        if (table) {
            const uint32_t row = std::min(f->table_cursor, f->table_rows - 1);
            rows = f->d_lead_table + (size_t)row * f->table_leads * 8;
            nrows = f->table_rows - row;
            P.n_leads = (int)f->table_leads;
            P.leads = rows;
        }
        if ((rc = mark(f)) || (rc = mark(f))) return rc;
        const size_t out_bytes = (size_t)f->n * 6 * sizeof(float);
        if (f->h_out_bytes < out_bytes) {
            if (f->h_out) cudaFreeHost(f->h_out);
            f->h_out = nullptr;
            f->h_out_bytes = 0;
            FP_CUDA(cudaHostAlloc((void **)&f->h_out, out_bytes, cudaHostAllocMapped));
            f->h_out_bytes = out_bytes;
        }
        float *d_out = nullptr;
        FP_CUDA(cudaHostGetDevicePointer((void **)&d_out, f->h_out, 0));
        rc = launch_small(f->stream, P, f->pos[f->cur], f->vel[f->cur], f->n, nsteps, rows, nrows,
                          f->d_status, f->pending_aos, d_out, f->first_index);
        if (rc) return rc;
        if (f->pending_aos) {  // the slot is free again once this launch has run
            FP_CUDA(cudaEventRecord(f->up_ev[f->pending_slot], f->stream));
            f->pending_aos = nullptr;
        }
        f->out_valid = true;
        if ((rc = mark(f))) return rc;
        f->table_cursor += nsteps;
        return FP_OK;
    }
    for (uint32_t s = 0; s < nsteps; ++s) {
        select_leads(f);
        if ((rc = mark(f))) return rc;
        rc = ensure_caller_order(f);
        if (rc) return rc;
        if ((rc = mark(f))) return rc;
        rc = flock_allpairs_step(f, f->pos[f->cur], f->vel[f->cur], f->n, 0, f->n, f->pos[f->cur ^ 1],
                                 f->vel[f->cur ^ 1]);
        if (rc) return rc;
        f->cur ^= 1;
        if ((rc = mark(f))) return rc;
        ++f->table_cursor;
    }
    return FP_OK;
}

int fp_flock_sync(fp_flock *f) {
    int rc = check(f);
    if (rc) return rc;
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

int fp_flock_timing_begin(fp_flock *f) {
    int rc = check(f);
    if (rc) return rc;
    f->ev_used = 0;
    f->timed_steps = 0;
    f->timing = true;
    return FP_OK;
}

int fp_flock_timing_end(fp_flock *f, uint32_t *steps, float *span_ms, float *sort_ms,
                        float *influence_ms) {
    int rc = check(f);  // settles: steps the device voided are replayed inside the timed span
    if (rc) return rc;
    const size_t k = f->ev_used / 3;  // step triples, replays included
    if ((rc = mark(f))) return rc;    // closing event
    f->timing = false;
    FP_CUDA(cudaStreamSynchronize(f->stream));
    float span = 0, so = 0, in = 0;
    for (size_t s = 0; s < k; ++s) {
        float a = 0, b = 0;
        FP_CUDA(cudaEventElapsedTime(&a, f->ev_pool[3 * s], f->ev_pool[3 * s + 1]));
        FP_CUDA(cudaEventElapsedTime(&b, f->ev_pool[3 * s + 1], f->ev_pool[3 * s + 2]));
        so += a;
        in += b;
    }
    if (k) FP_CUDA(cudaEventElapsedTime(&span, f->ev_pool[0], f->ev_pool[3 * k]));
    f->ev_used = 0;
    if (steps) *steps = f->timed_steps;
    if (span_ms) *span_ms = span;
    if (sort_ms) *sort_ms = so;
    if (influence_ms) *influence_ms = in;
    return FP_OK;
}

int fp_flock_status(fp_flock *f, uint32_t *flags) {
    int rc = check(f);
    if (rc) return rc;
    unsigned v = 0;
    FP_CUDA(cudaMemcpyAsync(&v, f->d_status, sizeof(v), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaMemsetAsync(f->d_status, 0, sizeof(v), f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    if (flags) *flags = v;
    return FP_OK;
}

int fp_flock_read_state(fp_flock *f, float *out) {
    int rc = check(f, true, true);
    if (rc) return rc;
    if ((rc = ingest_pending(f))) return rc;
    if (f->shard) return shard_read_state(f->shard, f, out);
    if (!f->n) return FP_OK;
    if (!out) { set_error("null output"); return FP_ERR_INVALID; }
    const size_t bytes = (size_t)f->n * 6 * sizeof(float);
    if (f->out_valid && f->h_out) {  // the last small step left the rows in mapped memory
        FP_CUDA(cudaStreamSynchronize(f->stream));
        memcpy(out, f->h_out, bytes);
        return FP_OK;
    }
    if ((rc = ensure_stage(f, bytes))) return rc;
    if ((rc = launch_soa_to_aos6(f->stream, f->pos[f->cur], f->vel[f->cur], (float *)f->d_stage, f->n,
                                 f->first_index, 1)))
        return rc;
    FP_CUDA(cudaMemcpyAsync(out, f->d_stage, bytes, cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

int fp_flock_write_state(fp_flock *f, const float *state) {
    int rc = check(f);
    if (rc) return rc;
    if (f->shard) { set_error("write_state is not supported on a sharded flock"); return FP_ERR_UNSUPPORTED; }
    if (!f->n) return FP_OK;
    if (!state) { set_error("null state"); return FP_ERR_INVALID; }
    const size_t bytes = (size_t)f->n * 6 * sizeof(float);
    if ((rc = ensure_stage(f, bytes))) return rc;
    if (bytes <= fp_flock::SMALL_UPLOAD) {
        // copy the caller's buffer out now; the device picks the rows up from the pinned slot when it
        // gets there -- the single-CTA step kernel itself, or a conversion kernel before anything else
        const uint32_t slot = f->up_cur++ & 1u;
        if (!f->h_up[slot]) {
            FP_CUDA(cudaHostAlloc(&f->h_up[slot], fp_flock::SMALL_UPLOAD, cudaHostAllocMapped));
            FP_CUDA(cudaEventCreateWithFlags(&f->up_ev[slot], cudaEventDisableTiming));
        }
        FP_CUDA(cudaEventSynchronize(f->up_ev[slot]));  // the kernel that last read this slot
        memcpy(f->h_up[slot], state, bytes);
        void *dptr = nullptr;
        FP_CUDA(cudaHostGetDevicePointer(&dptr, f->h_up[slot], 0));
        f->pending_aos = (const float *)dptr;
        f->pending_slot = slot;
    } else {
        FP_CUDA(cudaMemcpyAsync(f->d_stage, state, bytes, cudaMemcpyHostToDevice, f->stream));
        if ((rc = launch_aos6_to_soa(f->stream, (const float *)f->d_stage, f->pos[f->cur], f->vel[f->cur],
                                     f->n, f->first_index)))
            return rc;
        FP_CUDA(cudaStreamSynchronize(f->stream));  // the caller's buffer is not retained
    }
    f->permuted = false;
    f->bin_valid = false;
    f->nl_off = false;
    if (!f->domain_user) f->grid_valid = false;
    return FP_OK;
}

static int read_instances(fp_flock *f, float *out, int raw) {
    int rc = check(f);
    if (rc) return rc;
    if (f->shard) { set_error("instance export of a sharded flock: use fp_flock_read_local"); return FP_ERR_UNSUPPORTED; }
    if (!f->n) return FP_OK;
    if (!out) { set_error("null output"); return FP_ERR_INVALID; }
    const size_t bytes = (size_t)f->n * (raw ? 25 : 8) * sizeof(float);
    if ((rc = ensure_stage(f, bytes))) return rc;
    if ((rc = launch_instances(f->stream, f->pos[f->cur], f->vel[f->cur], (float *)f->d_stage, f->n,
                               f->first_index, raw)))
        return rc;
    FP_CUDA(cudaMemcpyAsync(out, f->d_stage, bytes, cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}
int fp_flock_export_instances(fp_flock *f, void *dst, int raw) {
    int rc = check(f);
    if (rc) return rc;
    if (f->shard) { set_error("instance export of a sharded flock: use fp_flock_read_local"); return FP_ERR_UNSUPPORTED; }
    if (!f->n) return FP_OK;
    if (!dst) { set_error("null output"); return FP_ERR_INVALID; }
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, dst) != cudaSuccess || at.type == cudaMemoryTypeUnregistered ||
        !at.devicePointer) {
        cudaGetLastError();
        set_error("export_instances: the destination is not addressable by the device (pageable host memory?)");
        return FP_ERR_INVALID;
    }
    if ((rc = launch_instances(f->stream, f->pos[f->cur], f->vel[f->cur], (float *)at.devicePointer, f->n,
                               f->first_index, raw ? 1 : 0)))
        return rc;
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}
int fp_flock_read_instances(fp_flock *f, float *out8) { return read_instances(f, out8, 0); }
int fp_flock_read_instances_raw(fp_flock *f, float *out25) { return read_instances(f, out25, 1); }

int fp_flock_read_accel(fp_flock *f, float *out_accel3, float *out_comp15) {
    int rc = check(f);
    if (rc) return rc;
    const uint64_t n = f->shard ? f->n_global : f->n;
    if (!n) return FP_OK;
    if (!out_accel3) { set_error("null output"); return FP_ERR_INVALID; }
    const size_t fl = (size_t)n * (out_comp15 ? 18 : 3);
    if ((rc = ensure_stage(f, fl * sizeof(float)))) return rc;
    FP_CUDA(cudaMemsetAsync(f->d_stage, 0, fl * sizeof(float), f->stream));
    TapOut t{};
    t.accel3 = (float *)f->d_stage;
    t.comp15 = out_comp15 ? (float *)f->d_stage + 3 * n : nullptr;
    if ((rc = run_tap(f, TAP_ACCEL, t))) return rc;
    FP_CUDA(cudaMemcpyAsync(out_accel3, t.accel3, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
    if (out_comp15)
        FP_CUDA(cudaMemcpyAsync(out_comp15, t.comp15, n * 15 * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

int fp_flock_read_neighbors(fp_flock *f, uint32_t *out_count, uint64_t *out_hash) {
    int rc = check(f);
    if (rc) return rc;
    const uint64_t n = f->shard ? f->n_global : f->n;
    if (!n) return FP_OK;
    if (!out_count || !out_hash) { set_error("null output"); return FP_ERR_INVALID; }
    const size_t bytes = (size_t)n * 12;
    if ((rc = ensure_stage(f, bytes))) return rc;
    FP_CUDA(cudaMemsetAsync(f->d_stage, 0, bytes, f->stream));
    TapOut t{};
    t.nbr_hash = (unsigned long long *)f->d_stage;
    t.nbr_count = (uint32_t *)((char *)f->d_stage + (size_t)n * 8);
    if ((rc = run_tap(f, TAP_NEIGHBORS, t))) return rc;
    FP_CUDA(cudaMemcpyAsync(out_hash, t.nbr_hash, n * 8, cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaMemcpyAsync(out_count, t.nbr_count, n * 4, cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

int fp_flock_pair_census(fp_flock *f, uint64_t out4[4]) {
    int rc = check(f);
    if (rc) return rc;
    if (!out4) { set_error("null output"); return FP_ERR_INVALID; }
    FP_CUDA(cudaMemsetAsync(f->d_census, 0, 4 * sizeof(unsigned long long), f->stream));
    TapOut t{};
    t.census = f->d_census;
    if (f->n || f->shard) {
        if ((rc = run_tap(f, TAP_CENSUS, t))) return rc;
    }
    FP_CUDA(cudaMemcpyAsync(out4, f->d_census, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

int fp_flock_device_state(fp_flock *f, const void **pos4, const void **vel4) {
    int rc = check(f);  // settles: the pointers describe steps that have happened
    if (rc) return rc;
    if (pos4) *pos4 = f->pos[f->cur];
    if (vel4) *vel4 = f->vel[f->cur];
    return FP_OK;
}

// State<boid>::euler_step / rk4_step with the acceleration accumulated first
static int flock_state_step(fp_flock *f, float h, int rk4) {
    int rc = check(f);
    if (rc) return rc;
    if (f->shard) { set_error("State integrators are single-GPU"); return FP_ERR_UNSUPPORTED; }
    if (!f->n) return FP_OK;
    if ((rc = ensure_stage(f, (size_t)f->n * 3 * sizeof(float)))) return rc;
    TapOut t{};
    t.accel3 = (float *)f->d_stage;
    if ((rc = run_tap(f, TAP_ACCEL, t))) return rc;
    f->bin_valid = false;  // positions move outside the walk's displacement accounting
    return launch_flock_state_step(f->stream, f->pos[f->cur], f->vel[f->cur], t.accel3, f->n,
                                   f->first_index, h, rk4);
}
int fp_flock_state_euler(fp_flock *f, float h) { return flock_state_step(f, h, 0); }
int fp_flock_state_rk4(fp_flock *f, float h) { return flock_state_step(f, h, 1); }

static int combine(int device, size_t n, const float *const *in, int nin, float h, float *out, int rk4) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_error("no usable CUDA device -- this library has no CPU fallback");
        return FP_ERR_CUDA;
    }
    if (!n) return FP_OK;
    for (int k = 0; k < nin; ++k)
        if (!in[k]) { set_error("null input"); return FP_ERR_INVALID; }
    if (!out) { set_error("null output"); return FP_ERR_INVALID; }
    FP_CUDA(cudaSetDevice(device));
    float *d = nullptr;
    FP_CUDA(cudaMalloc((void **)&d, (size_t)(nin + 1) * n * sizeof(float)));
    int rc = FP_OK;
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < nin && e == cudaSuccess; ++k)
        e = cudaMemcpy(d + (size_t)k * n, in[k], n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        float *o = d + (size_t)nin * n;
        rc = rk4 ? launch_state_combine_rk4(0, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, h, o, n)
                 : launch_state_combine_euler(0, d, d + n, h, o, n);
        if (!rc) e = cudaMemcpy(out, o, n * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (e != cudaSuccess) return cuda_fail(e, "state combine", __FILE__, __LINE__);
    return rc;
}

int fp_state_euler_combine(int device, size_t n, const float *s, const float *ds, float h, float *out) {
    const float *in[2] = {s, ds};
    return combine(device, n, in, 2, h, out, 0);
}
int fp_state_rk4_combine(int device, size_t n, const float *s, const float *k1, const float *k2,
                         const float *k3, const float *k4, float h, float *out) {
    const float *in[5] = {s, k1, k2, k3, k4};
    return combine(device, n, in, 5, h, out, 1);
}

int fp_nccl_unique_id(uint8_t out128[128]) { return shard_unique_id(out128); }

int fp_debug_fastmath_check(int device, uint64_t n, uint64_t seed, uint64_t out_mismatch[2]) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_error("no usable CUDA device -- this library has no CPU fallback");
        return FP_ERR_CUDA;
    }
    if (!out_mismatch) { set_error("null output"); return FP_ERR_INVALID; }
    FP_CUDA(cudaSetDevice(device));
    return launch_fastmath_check(n, seed, out_mismatch);
}

// host copy of the resident records; owned ones only (ghost / dead halo records are skipped)
static int fetch_owned(fp_flock *f, std::vector<float4> &p, std::vector<float4> &v) {
    p.resize(f->n);
    v.resize(f->n);
    if (f->n) {
        FP_CUDA(cudaMemcpyAsync(p.data(), f->pos[f->cur], (size_t)f->n * sizeof(float4), cudaMemcpyDeviceToHost,
                                f->stream));
        FP_CUDA(cudaMemcpyAsync(v.data(), f->vel[f->cur], (size_t)f->n * sizeof(float4), cudaMemcpyDeviceToHost,
                                f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
    }
    size_t k = 0;
    for (size_t i = 0; i < p.size(); ++i) {
        uint32_t flag;
        memcpy(&flag, &v[i].w, 4);
        if (flag != 0) continue;
        p[k] = p[i];
        v[k] = v[i];
        ++k;
    }
    p.resize(k);
    v.resize(k);
    return FP_OK;
}

int fp_flock_shard_info(fp_flock *f, int *rank, int *world, int *peer_mapped) {
    int rc = check(f);
    if (rc) return rc;
    if (rank) *rank = 0;
    if (world) *world = 1;
    if (peer_mapped) *peer_mapped = 0;
    if (f->shard) shard_info(f->shard, rank, world, peer_mapped);
    return FP_OK;
}

int fp_flock_local_len(fp_flock *f, uint64_t *n_local) {
    int rc = check(f);
    if (rc) return rc;
    if (!n_local) { set_error("null argument"); return FP_ERR_INVALID; }
    if (!f->shard) {
        *n_local = f->n;
        return FP_OK;
    }
    return shard_read_local(f->shard, f, n_local, nullptr, nullptr);
}

int fp_flock_read_local(fp_flock *f, uint64_t *out_index, float *out_aos6) {
    int rc = check(f);
    if (rc) return rc;
    if (f->shard) {
        uint64_t n_own = 0;
        if (!out_index || !out_aos6) { set_error("null output"); return FP_ERR_INVALID; }
        return shard_read_local(f->shard, f, &n_own, out_index, out_aos6);
    }
    std::vector<float4> p, v;
    if ((rc = fetch_owned(f, p, v))) return rc;
    if (p.empty()) return FP_OK;
    if (!out_index || !out_aos6) { set_error("null output"); return FP_ERR_INVALID; }
    for (size_t i = 0; i < p.size(); ++i) {
        uint32_t u;
        memcpy(&u, &p[i].w, 4);
        out_index[i] = u;
        float *o = out_aos6 + 6 * i;
        o[0] = p[i].x; o[1] = p[i].y; o[2] = p[i].z;
        o[3] = v[i].x; o[4] = v[i].y; o[5] = v[i].z;
    }
    return FP_OK;
}


int fp_flock_write_local(fp_flock *f, uint64_t n_local, const uint64_t *index, const float *state_aos6) {
    int rc = check(f);
    if (rc) return rc;
    if (f->shard) return shard_write_local(f->shard, f, n_local, index, state_aos6);
    // single GPU: the same contract over the resident order
    if (n_local != f->n) { set_error("write_local: row count differs from fp_flock_len"); return FP_ERR_INVALID; }
    if (!f->n) return FP_OK;
    if (!index || !state_aos6) { set_error("null input"); return FP_ERR_INVALID; }
    std::vector<float4> p, v;
    if ((rc = fetch_owned(f, p, v))) return rc;
    for (size_t i = 0; i < p.size(); ++i) {
        uint32_t u;
        memcpy(&u, &p[i].w, 4);
        if (u != index[i]) { set_error("write_local: index order differs from fp_flock_read_local"); return FP_ERR_INVALID; }
        const float *s = state_aos6 + 6 * i;
        p[i] = make_float4(s[0], s[1], s[2], p[i].w);
        v[i] = make_float4(s[3], s[4], s[5], 0.0f);
    }
    FP_CUDA(cudaMemcpyAsync(f->pos[f->cur], p.data(), p.size() * sizeof(float4), cudaMemcpyHostToDevice, f->stream));
    FP_CUDA(cudaMemcpyAsync(f->vel[f->cur], v.data(), v.size() * sizeof(float4), cudaMemcpyHostToDevice, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    f->bin_valid = false;
    if (!f->domain_user) f->grid_valid = false;
    return FP_OK;
}

}  // extern "C"

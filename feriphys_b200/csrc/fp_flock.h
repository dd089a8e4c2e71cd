// fp_flock.h -- the handle behind the opaque fp_flock of the C ABI.
#pragma once

#include <vector>

#include "fp_internal.h"

namespace fp {
struct Shard;
}

struct fp_flock {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t n = 0;        // boids held by this handle (local rows when sharded)
    uint32_t cap = 0;      // capacity of pos/vel
    uint64_t n_global = 0;
    uint32_t first_index = 0;
    fp_config cfg{};
    fp::DevParams P{};
    float4 *pos[2] = {nullptr, nullptr}, *vel[2] = {nullptr, nullptr};
    int cur = 0;
    bool permuted = false;
    int method = FP_METHOD_AUTO, method_in_use = FP_METHOD_ALLPAIRS;
    int numerics = FP_NUMERICS_EXACT;
    float *d_leads = nullptr, *d_attr = nullptr, *d_obs = nullptr, *d_lead_table = nullptr;
    uint32_t n_leads = 0, n_attr = 0, n_obs = 0, table_rows = 0, table_leads = 0, table_cursor = 0;
    // fp_flock_set_leads keeps the last LEAD_RING row sets on the device: a step enqueued under
    // version v reads slot v % LEAD_RING, so the caller can hand over new rows before every step
    // without the host waiting for the steps in flight (a replayed step finds the rows it ran with)
    static constexpr uint32_t LEAD_RING = 512, LEAD_STAGES = 8;
    float *d_lead_ring = nullptr;          // LEAD_RING x n_leads x 8 floats; d_leads points into it
    float *h_lead_stage = nullptr;         // pinned, LEAD_STAGES x n_leads x 8
    cudaEvent_t lead_ev[LEAD_STAGES] = {}; // the copy out of each staging slot
    uint32_t lead_ver = 0, lead_stage_cur = 0;
    std::vector<uint32_t> replay_lead_vers;  // versions of the steps being replayed (front first)
    size_t attr_cap = 0, obs_cap = 0;      // floats allocated behind d_attr / d_obs
    unsigned *d_status = nullptr;
    unsigned long long *d_census = nullptr;
    float *d_bounds = nullptr;
    // grid
    fp::GridDesc grid{};
    bool grid_valid = false, domain_user = false;
    float user_lo[3]{}, user_hi[3]{};
    uint64_t steps_since_fit = 0;
    fp::GridWork work{};
    // lazy re-binning (single GPU and sharded): see settle() in fp_api.cu
    bool bin_valid = false;      // pos[cur] is cell-sorted under `grid`; work.home / cell_start describe it
    float skin_budget = 0.0f;    // skin / 2: the displacement bound a binning tolerates
    float delta_est = 0.0f;      // planning estimate of the per-step displacement bound (with margin)
    int64_t plan_left = 0;       // steps that may still be enqueued before the planned re-binning
    uint32_t steps_since_bin = 0;  // sharded: parity of the neighbours' buffers
    uint32_t ordinal = 0;        // ordinal of the next step (what the gate kernel records when stale)
    struct Pending {             // a step that is enqueued but not yet known to have happened
        uint32_t ordinal;
        int cur, soa_cur;
        uint32_t table_cursor;
        uint64_t steps_since_fit;
        uint32_t steps_since_bin;
        uint32_t lead_ver;
    };
    std::vector<Pending> pending;
    fp::SkinCtl *h_ctl = nullptr;  // pinned read-back of work.ctl
    uint64_t stat_rebins = 0, stat_replayed = 0, stat_grid_steps = 0;
    uint32_t timed_steps = 0;
    float skin_override = -1.0f;  // < 0: sized from the flock's speed at every fit
    float plan_scale = 1.0f;      // stretches the planned steps per binning (tests)
    // standing candidate lists (fp_walk_nl.cu)
    uint16_t *nl_entries = nullptr, *nl_count = nullptr;
    uint32_t *nl_cta_tab = nullptr;
    unsigned *nl_flag = nullptr;
    uint32_t nl_rows = 0;        // boids the buffers are sized for
    uint32_t nl_built_rows = 0;  // boids of the last build
    uint64_t nl_serial = ~0ull;  // stat_rebins of the binning the lists describe
    bool nl_fresh = false;       // the flock was binned and has not stepped since: lists may be built
    uint32_t nl_bin_steps = 0, nl_prev_bin_steps = 0;  // steps walked on this binning / on the one before
    bool nl_off = false;         // a list overflowed: production walk until a new state / config arrives
    // staging for host transfers
    void *d_stage = nullptr;
    size_t stage_bytes = 0;
    // small uploads (demo-sized flocks) go through two pinned slots so that fp_flock_write_state
    // can return without waiting for the device: the caller's buffer is copied out at once
    static constexpr size_t SMALL_UPLOAD = 256 * 1024;
    void *h_up[2] = {nullptr, nullptr};
    cudaEvent_t up_ev[2] = {nullptr, nullptr};
    uint32_t up_cur = 0;
    // Demo-sized flocks on the single-CTA kernel skip the copies altogether: fp_flock_write_state
    // leaves the caller's rows in the pinned (device-mapped) slot, the step kernel reads them from
    // there and writes the advanced rows to a second mapped buffer, fp_flock_read_state waits for the
    // stream and copies them out -- one launch and one synchronisation per write / step / read round.
    const float *pending_aos = nullptr;  // rows handed over by write_state that no kernel has ingested yet
    uint32_t pending_slot = 0;
    float *h_out = nullptr;              // pinned, mapped: caller-order rows written by the last small step
    size_t h_out_bytes = 0;
    bool out_valid = false;              // h_out describes the current state
    // timing hook: three events per step (before sort phase, before influence, after)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    bool timing = false;
    // all-pairs kernel choice (staged / one-phase: bit-identical, speed depends on density):
    // steps 0 and 1 after a (re)start are timed, one kernel each, then the faster one is kept
    int ap_choice = -1;
    uint32_t ap_probe = 0, ap_age = 0;
    cudaEvent_t ap_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    fp::Shard *shard = nullptr;  // multi-GPU state (fp_shard.cu)
};

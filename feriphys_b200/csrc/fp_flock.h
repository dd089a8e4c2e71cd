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
    float *d_leads = nullptr, *d_attr = nullptr, *d_obs = nullptr, *d_lead_table = nullptr;
    uint32_t n_leads = 0, n_attr = 0, n_obs = 0, table_rows = 0, table_leads = 0, table_cursor = 0;
    unsigned *d_status = nullptr;
    unsigned long long *d_census = nullptr;
    float *d_bounds = nullptr;
    // grid
    fp::GridDesc grid{};
    bool grid_valid = false, domain_user = false;
    float user_lo[3]{}, user_hi[3]{};
    uint64_t steps_since_fit = 0;
    fp::GridWork work{};
    // staging for host transfers
    void *d_stage = nullptr;
    size_t stage_bytes = 0;
    // timing hook: three events per step (before sort phase, before influence, after)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    bool timing = false;
    fp::Shard *shard = nullptr;  // multi-GPU state (fp_shard.cu)
};

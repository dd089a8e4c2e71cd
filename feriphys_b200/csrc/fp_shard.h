// fp_shard.h -- multi-GPU sharding of one flock (one process per GPU, NCCL).
#pragma once

#include "fp_flock.h"

namespace fp {

// implemented in fp_api.cu
int flock_fit_grid(fp_flock *f);
uint32_t flock_select_leads(fp_flock *f);
int flock_mark(fp_flock *f);  // timing-hook event
// candidate lists (fp_walk_nl.cu): build after a binning if wanted; the step's walk; a tap's walk
int flock_nl_prepare(fp_flock *f, const GridDesc &g, const WalkIO &io);
void flock_nl_binned(fp_flock *f);  // bookkeeping: the flock has just been binned
int flock_step_walk(fp_flock *f, const GridDesc &g, const WalkIO &io);  // lists if on hand, else the staged walk
int flock_tap_walk(fp_flock *f, const GridDesc &g, int tap, const WalkIO &io, const TapOut &out);
// one all-pairs step launch with the staged / one-phase choice made by measurement
int flock_allpairs_step(fp_flock *f, const float4 *pos_all, const float4 *vel_all, uint32_t n_all, uint32_t row0,
                        uint32_t nrows, float4 *pos_out, float4 *vel_out);
int64_t flock_plan_steps(const fp_flock *f, float D, float first_delta);
float flock_plan_delta(float v2max, float pmax, float dt);

// implemented in fp_shard.cu
int shard_unique_id(uint8_t out128[128]);
int shard_create(Shard **out, fp_flock *f, int rank, int world, const uint8_t id[128]);
void shard_destroy(Shard *s);
int shard_method(Shard *s, int requested, const fp_config &cfg);
int shard_reduce_bounds(Shard *s, cudaStream_t st, float lo[3], float hi[3], float *v2max);
int shard_settle(Shard *s, fp_flock *f);
int shard_world(const Shard *s);
bool shard_is_slab(const Shard *s);  // the flock currently lives in x-slabs (grid partition)
void shard_info(const Shard *s, int *rank, int *world, int *peer_mapped);
// called by fit_grid once the GLOBAL grid is known: lay out this rank's slab
int shard_grid_fitted(Shard *s, fp_flock *f);
int shard_step(Shard *s, fp_flock *f, uint32_t nsteps);
int shard_tap(Shard *s, fp_flock *f, int tap, const TapOut &out);
int shard_read_state(Shard *s, fp_flock *f, float *out_aos6);
// owned records: count always, contents when the output pointers are non-null
int shard_read_local(Shard *s, fp_flock *f, uint64_t *n_local, uint64_t *out_index, float *out_aos6);
int shard_write_local(Shard *s, fp_flock *f, uint64_t n_local, const uint64_t *index, const float *aos6);
// index range [first, first + count) rank owns under the boid-index partition
void shard_index_range(uint64_t n_global, int rank, int world, uint64_t *first, uint64_t *count);

}  // namespace fp

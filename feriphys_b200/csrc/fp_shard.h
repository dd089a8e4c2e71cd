// fp_shard.h -- multi-GPU sharding of one flock (one process per GPU).
#pragma once

#include "fp_internal.h"

struct fp_flock;

namespace fp {

struct Shard;

// The pieces of a handle the sharding layer works on.
struct FlockView {
    cudaStream_t stream;
    DevParams *P;
    float4 *pos[2], *vel[2];
    int *cur;
    uint32_t *n;
    uint32_t cap;
    bool *permuted;
    unsigned *status;
    GridDesc *grid;
    GridWork *work;
    uint32_t first_index;
    int method;
    const fp_config *cfg;
};

FlockView flock_view(fp_flock *f);
int flock_grid_prepare_fit(fp_flock *f);
void flock_count_steps(fp_flock *f, uint64_t k);
int flock_mark(fp_flock *f);  // timing hook event

int shard_unique_id(uint8_t out128[128]);
int shard_create(Shard **out, fp_flock *f, int rank, int world, const uint8_t id[128]);
void shard_destroy(Shard *s);
uint32_t shard_capacity(Shard *s);
int shard_method(Shard *s, int requested, const fp_config &cfg);
int shard_reduce_bounds(Shard *s, cudaStream_t st, float lo[3], float hi[3]);
int shard_step(Shard *s, fp_flock *f, uint32_t nsteps);
int shard_tap(Shard *s, fp_flock *f, int tap, const TapOut &out);
int shard_read_state(Shard *s, fp_flock *f, float *out_aos6);

}  // namespace fp

// fp_shard.cu -- multi-GPU sharding (placeholder until the NCCL layer lands).
#include "fp_shard.h"

namespace fp {

struct Shard {};

int shard_unique_id(uint8_t *) {
    set_error("multi-GPU sharding is not built yet");
    return FP_ERR_UNSUPPORTED;
}
int shard_create(Shard **out, fp_flock *, int, int, const uint8_t *) {
    *out = nullptr;
    set_error("multi-GPU sharding is not built yet");
    return FP_ERR_UNSUPPORTED;
}
void shard_destroy(Shard *) {}
uint32_t shard_capacity(Shard *) { return 0; }
int shard_method(Shard *, int requested, const fp_config &) { return requested; }
int shard_reduce_bounds(Shard *, cudaStream_t, float *, float *) { return FP_OK; }
int shard_step(Shard *, fp_flock *, uint32_t) { return FP_ERR_UNSUPPORTED; }
int shard_tap(Shard *, fp_flock *, int, const TapOut &) { return FP_ERR_UNSUPPORTED; }
int shard_read_state(Shard *, fp_flock *, float *) { return FP_ERR_UNSUPPORTED; }

}  // namespace fp

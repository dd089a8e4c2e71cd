// fp_sort.cu -- K2a: stable LSD radix sort of (cell key, boid slot) pairs.
//
// HBM-bound integer work: per pass the keys are read once for the per-tile
// digit histogram (4 B), then keys+values are read and written once by the
// scatter (16 B).  Stability (equal keys keep their input order) makes the
// within-cell order -- and therefore every f32 sum in the walk kernel --
// deterministic.  Ranking is warp-synchronous: __match_any_sync groups equal
// digits, the group leader bumps a per-warp shared-memory counter.
#include "fp_internal.h"

namespace fp {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 keys per CTA
constexpr int RS_MAX_BITS = 8;
constexpr int RS_BINS = 1 << RS_MAX_BITS;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t mask,
                  uint32_t *__restrict__ tile_hist, uint32_t ntiles) {
    __shared__ uint32_t hist[RS_BINS];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        const uint32_t d = valid ? ((keys[i] >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[d], __popc(peers));
    }
    __syncthreads();
    if (threadIdx.x <= mask) tile_hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                     int shift, uint32_t mask, const uint32_t *__restrict__ tile_offsets,
                     uint32_t ntiles) {
    __shared__ uint32_t wcount[RS_WARPS][RS_BINS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * RS_BINS; k += RS_THREADS) (&wcount[0][0])[k] = 0;
    __syncthreads();

    // warp w owns the contiguous slice [w*512, (w+1)*512) of the tile; round r covers 32
    // consecutive keys, so (warp, round, lane) order is input order.
    const uint32_t wbase = blockIdx.x * RS_TILE + wid * (RS_ITEMS * 32);
    uint32_t key[RS_ITEMS], rank[RS_ITEMS];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0xffffffffu;
        const uint32_t d = valid ? ((key[r] >> shift) & mask) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        uint32_t base = 0;
        if (valid) base = wcount[wid][d];
        __syncwarp();
        if (valid && lane == (__ffs(peers) - 1)) wcount[wid][d] = base + __popc(peers);
        __syncwarp();
        rank[r] = base + __popc(peers & lt);
    }
    __syncthreads();
    // digit d = threadIdx.x: exclusive prefix over the warps + global base of (digit, tile)
    if (threadIdx.x <= mask) {
        uint32_t run = tile_offsets[(size_t)threadIdx.x * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = wcount[w][threadIdx.x];
            wcount[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & mask;
            const uint32_t dst = wcount[wid][d] + rank[r];
            keys_out[dst] = key[r];
            vals_out[dst] = vals_in ? vals_in[i] : i;  // first pass: values are the slots 0..n-1
        }
    }
}

int launch_radix_sort(cudaStream_t st, GridWork &w, uint32_t n, uint32_t key_bits, int *out_buf) {
    *out_buf = 0;
    if (n == 0) return FP_OK;
    if (key_bits == 0) key_bits = 1;
    const int passes = (int)((key_bits + RS_MAX_BITS - 1) / RS_MAX_BITS);
    const int bits = (int)((key_bits + passes - 1) / passes);  // even split, <= 8
    const uint32_t mask = (1u << bits) - 1u;
    const uint32_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    const size_t hist_elems = (size_t)(mask + 1) * ntiles;
    if (hist_elems > w.tile_hist_elems) {
        set_error("radix sort scratch too small");
        return FP_ERR_INVALID;
    }
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * bits;
        radix_hist_kernel<<<ntiles, RS_THREADS, 0, st>>>(w.keys[cur], n, shift, mask, w.tile_hist, ntiles);
        count_launch();
        int rc = launch_exclusive_scan(st, w.tile_hist, hist_elems, w.scan_tmp);
        if (rc) return rc;
        radix_scatter_kernel<<<ntiles, RS_THREADS, 0, st>>>(
            w.keys[cur], (p == 0) ? nullptr : w.vals[cur], w.keys[cur ^ 1], w.vals[cur ^ 1], n, shift,
            mask, w.tile_hist, ntiles);
        count_launch();
        cur ^= 1;
    }
    FP_CUDA(cudaGetLastError());
    *out_buf = cur;
    return FP_OK;
}

}  // namespace fp

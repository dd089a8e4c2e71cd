// fp_shard.cu -- one flock over several GPUs (one process per GPU, NCCL over NVLink).
//
// Two partitions, both exact:
//  * all-pairs: rows i are independent given the full old state (Jacobi update,
//    flocking.rs:101-122).  Ranks own contiguous boid-index ranges and exchange
//    the new positions/velocities with ONE ncclAllGather per step; every rank
//    sums j = 0..N-1 in index order, so results are bit-identical to one GPU.
//  * grid: ranks own slabs of whole cell layers along x.  Per step each rank
//    sends its two boundary layers to the +-1 neighbours as read-only "ghost"
//    copies, together with the boids whose Euler update carried them across the
//    slab face (ownership transfer) -- one fixed-size ncclSend/ncclRecv pair per
//    face, counts travelling in a header record, no host round trip for sizes.
//    Ghosts, arrivals and residents are merged by the same stable radix sort that
//    builds the cell table; dead records get the key `ncells` and fall off the end.
//
// The interchange form between partitions, and for read-back, is the flock in
// global index order on every rank (all_pos / all_vel).
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is resolved at run time with dlopen
#include <string.h>

#include <algorithm>

#include "fp_grid.cuh"
#include "fp_shard.h"

namespace fp {

// ---- NCCL, resolved lazily so the single-GPU library has no link-time dependency -----
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static int load_nccl(NcclApi &a) {
    if (a.lib) return FP_OK;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) {
        set_error(std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror());
        return FP_ERR_NCCL;
    }
#define FP_SYM(field, sym)                                            \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym)); \
    if (!a.field) {                                                   \
        set_error(std::string("NCCL symbol missing: ") + sym);        \
        return FP_ERR_NCCL;                                           \
    }
    FP_SYM(GetUniqueId, "ncclGetUniqueId")
    FP_SYM(CommInitRank, "ncclCommInitRank")
    FP_SYM(CommDestroy, "ncclCommDestroy")
    FP_SYM(AllGather, "ncclAllGather")
    FP_SYM(AllReduce, "ncclAllReduce")
    FP_SYM(Send, "ncclSend")
    FP_SYM(Recv, "ncclRecv")
    FP_SYM(GroupStart, "ncclGroupStart")
    FP_SYM(GroupEnd, "ncclGroupEnd")
    FP_SYM(GetErrorString, "ncclGetErrorString")
#undef FP_SYM
    return FP_OK;
}

static NcclApi g_nccl;

#define FP_NCCL(s, expr)                                                                      \
    do {                                                                                      \
        ncclResult_t _r = (expr);                                                             \
        if (_r != ncclSuccess) {                                                              \
            set_error(std::string("NCCL error: ") + (s)->api.GetErrorString(_r) + " in " #expr); \
            return FP_ERR_NCCL;                                                               \
        }                                                                                     \
    } while (0)

enum Rep { REP_SLICE = 0, REP_SLAB = 1 };

struct Shard {
    int rank = 0, world = 1;
    NcclApi api;
    ncclComm_t comm = nullptr;
    uint64_t n_global = 0;
    uint32_t per = 0;          // index partition: rows per rank (last rank may hold fewer)
    uint32_t first = 0, n_slice = 0;
    int rep = REP_SLICE;       // what f->pos[f->cur][0 .. f->n) currently holds
    bool global_valid = false; // all_*[acur] is the current flock in index order
    float4 *all_pos[2] = {nullptr, nullptr}, *all_vel[2] = {nullptr, nullptr};
    int acur = 0;
    // slab
    int xs0 = 0, xs1 = 0;      // owned global cell layers [xs0, xs1)
    GridDesc lgrid{};          // this rank's grid: slab + halo layers
    uint32_t halo_cap = 0;     // records per face message (plus one header record)
    float4 *send_pos[2] = {nullptr, nullptr}, *send_vel[2] = {nullptr, nullptr};  // [0] left, [1] right
    uint32_t *blk_cnt[2] = {nullptr, nullptr};
    size_t blk_cap = 0;
    uint32_t *h_live = nullptr;  // pinned
    cudaEvent_t ev_live = nullptr;  // completion of the live-count read-back
};

void shard_index_range(uint64_t n, int rank, int world, uint64_t *first, uint64_t *count) {
    const uint64_t per = (n + world - 1) / world;
    const uint64_t f = std::min<uint64_t>((uint64_t)rank * per, n);
    *first = f;
    *count = std::min<uint64_t>(per, n - f);
}

int shard_unique_id(uint8_t out128[128]) {
    int rc = load_nccl(g_nccl);
    if (rc) return rc;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) {
        set_error("ncclGetUniqueId failed");
        return FP_ERR_NCCL;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return FP_OK;
}

template <class T>
static int dalloc(T **p, size_t count) {
    *p = nullptr;
    FP_CUDA(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
    return FP_OK;
}

int shard_create(Shard **out, fp_flock *f, int rank, int world, const uint8_t id128[128]) {
    *out = nullptr;
    int rc = load_nccl(g_nccl);
    if (rc) return rc;
    Shard *s = new Shard();
    s->api = g_nccl;
    s->rank = rank;
    s->world = world;
    s->n_global = f->n_global;
    uint64_t first, count;
    shard_index_range(f->n_global, rank, world, &first, &count);
    s->per = (uint32_t)((f->n_global + world - 1) / world);
    if (first != f->first_index || count != f->n) {
        set_error("sharded create: rank r must hold the index range [r*per, (r+1)*per) with per = ceil(n/world)");
        delete s;
        return FP_ERR_INVALID;
    }
    s->first = (uint32_t)first;
    s->n_slice = (uint32_t)count;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = s->api.CommInitRank(&s->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error(std::string("ncclCommInitRank: ") + s->api.GetErrorString(r));
        delete s;
        return FP_ERR_NCCL;
    }
    const size_t total = (size_t)s->per * world;
    for (int b = 0; b < 2; ++b)
        if ((rc = dalloc(&s->all_pos[b], total)) || (rc = dalloc(&s->all_vel[b], total))) return rc;
    if (cudaMallocHost((void **)&s->h_live, 4 * sizeof(uint32_t)) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "cudaMallocHost", __FILE__, __LINE__);
    if (cudaEventCreateWithFlags(&s->ev_live, cudaEventDisableTiming) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "cudaEventCreate", __FILE__, __LINE__);
    *out = s;
    return FP_OK;
}

void shard_destroy(Shard *s) {
    if (!s) return;
    for (int b = 0; b < 2; ++b) {
        cudaFree(s->all_pos[b]);
        cudaFree(s->all_vel[b]);
        cudaFree(s->send_pos[b]);
        cudaFree(s->send_vel[b]);
        cudaFree(s->blk_cnt[b]);
    }
    if (s->h_live) cudaFreeHost(s->h_live);
    if (s->ev_live) cudaEventDestroy(s->ev_live);
    if (s->comm) s->api.CommDestroy(s->comm);
    delete s;
}


int shard_method(Shard *, int requested, const fp_config &cfg) {
    const float thr = cfg.distance_weight_threshold;
    const float reach = std::max(thr, thr + cfg.distance_weight_threshold_falloff);
    const bool grid_ok = reach > 0.0f && std::isfinite(reach);
    if (requested == FP_METHOD_GRID) return FP_METHOD_GRID;
    if (requested == FP_METHOD_AUTO && grid_ok) return FP_METHOD_GRID;
    return FP_METHOD_ALLPAIRS;
}

// lo/hi <- min/max over ranks (device round trip through the flock's bounds scratch is the caller's)
int shard_reduce_bounds(Shard *s, cudaStream_t st, float lo[3], float hi[3], float *v2max) {
    float *d = nullptr;
    FP_CUDA(cudaMalloc((void **)&d, 7 * sizeof(float)));
    float h[7] = {lo[0], lo[1], lo[2], -hi[0], -hi[1], -hi[2], -*v2max};  // one ncclMin covers all
    cudaError_t e = cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, st);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess) r = s->api.AllReduce(d, d, 7, ncclFloat32, ncclMin, s->comm, st);
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (r != ncclSuccess) {
        set_error(std::string("NCCL allreduce(bounds): ") + s->api.GetErrorString(r));
        return FP_ERR_NCCL;
    }
    if (e != cudaSuccess) return cuda_fail(e, "bounds reduce", __FILE__, __LINE__);
    for (int a = 0; a < 3; ++a) {
        lo[a] = h[a];
        hi[a] = -h[3 + a];
    }
    *v2max = -h[6];
    return FP_OK;
}

// ---- kernels ---------------------------------------------------------------------------
constexpr int SB = 256;
static inline unsigned nblk(size_t n) { return (unsigned)((n + SB - 1) / SB); }

enum { REC_OWNED = 0u, REC_GHOST = 1u, REC_DEAD = 2u };

// owned boids of a slab-resident array -> their rows of a zeroed index-ordered flock
__global__ void scatter_owned_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                     uint32_t n, float4 *__restrict__ all_pos,
                                     float4 *__restrict__ all_vel) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    if (i >= n) return;
    const float4 v = vel[i];
    if (__float_as_uint(v.w) != REC_OWNED) return;
    const float4 p = pos[i];
    const uint32_t g = __float_as_uint(p.w);
    all_pos[g] = p;
    all_vel[g] = make_float4(v.x, v.y, v.z, 0.0f);
}

// select this rank's slab from the index-ordered flock, keeping index order (deterministic):
// pass 1 counts per block, pass 2 writes at scanned offsets
__device__ __forceinline__ bool in_slab(const GridDesc &g, int xs0, int xs1, float x) {
    const int gx = cell_coord(x, g.origin[0], g.inv_cell, g.gdimx);
    return gx >= xs0 && gx < xs1;
}
__global__ void slab_select_count_kernel(const GridDesc g, int xs0, int xs1, const float4 *__restrict__ all_pos,
                                         uint32_t n, uint32_t *__restrict__ blk) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    const bool take = i < n && in_slab(g, xs0, xs1, all_pos[i].x);
    const int c = __syncthreads_count(take);
    if (threadIdx.x == 0) blk[blockIdx.x] = (uint32_t)c;
}
__device__ __forceinline__ uint32_t block_rank(bool flag) {
    // exclusive rank of this thread among the flagged threads of its block (thread order)
    __shared__ uint32_t wsum[SB / 32];
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    uint32_t base = 0;
    for (int k = 0; k < w; ++k) base += wsum[k];
    __syncthreads();
    return base + __popc(b & ((1u << lane) - 1u));
}
__global__ void slab_select_emit_kernel(const GridDesc g, int xs0, int xs1, const float4 *__restrict__ all_pos,
                                        const float4 *__restrict__ all_vel, uint32_t n,
                                        const uint32_t *__restrict__ blk_off, float4 *__restrict__ pos,
                                        float4 *__restrict__ vel, uint32_t cap, unsigned *status) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    const bool take = i < n && in_slab(g, xs0, xs1, all_pos[i].x);
    const uint32_t r = block_rank(take);
    if (!take) return;
    const uint32_t dst = blk_off[blockIdx.x] + r;
    if (dst >= cap) {
        atomicOr(status, 4u);
        return;
    }
    const float4 v = all_vel[i];
    pos[dst] = all_pos[i];
    vel[dst] = make_float4(v.x, v.y, v.z, __uint_as_float(REC_OWNED));
}

// What happens to record i of the slab-resident array at the start of a step.
struct Fate {
    bool send[2];      // a copy goes to the left / right neighbour ...
    uint32_t as[2];    // ... as REC_OWNED (ownership transfer) or REC_GHOST
    uint32_t keep;     // what the local record becomes
    bool jumped;       // crossed more than one layer in a step
};
__device__ __forceinline__ Fate fate_of(const GridDesc &g, int xs0, int xs1, int rank, int world, float x,
                                        uint32_t flag) {
    Fate f;
    f.send[0] = f.send[1] = false;
    f.as[0] = f.as[1] = REC_GHOST;
    f.keep = REC_DEAD;
    f.jumped = false;
    if (flag != REC_OWNED) return f;  // last step's ghosts and dead records are dropped
    const int gx = cell_coord(x, g.origin[0], g.inv_cell, g.gdimx);
    f.keep = REC_OWNED;
    if (gx < xs0) {  // left the slab to the left: the neighbour owns it now
        f.send[0] = true;
        f.as[0] = REC_OWNED;
        f.keep = (gx == xs0 - 1) ? REC_GHOST : REC_DEAD;
        f.jumped = gx < xs0 - 1;
    } else if (gx >= xs1) {
        f.send[1] = true;
        f.as[1] = REC_OWNED;
        f.keep = (gx == xs1) ? REC_GHOST : REC_DEAD;
        f.jumped = gx > xs1;
    } else {
        if (gx == xs0 && rank > 0) f.send[0] = true;            // boundary layers: ghost copies
        if (gx == xs1 - 1 && rank < world - 1) f.send[1] = true;
    }
    return f;
}

__global__ void slab_fate_count_kernel(const GridDesc g, int xs0, int xs1, int rank, int world,
                                       const float4 *__restrict__ pos, const float4 *__restrict__ vel,
                                       uint32_t n, uint32_t *__restrict__ blk_l, uint32_t *__restrict__ blk_r) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    bool l = false, r = false;
    if (i < n) {
        const Fate f = fate_of(g, xs0, xs1, rank, world, pos[i].x, __float_as_uint(vel[i].w));
        l = f.send[0];
        r = f.send[1];
    }
    const int cl = __syncthreads_count(l), cr = __syncthreads_count(r);
    if (threadIdx.x == 0) {
        blk_l[blockIdx.x] = (uint32_t)cl;
        blk_r[blockIdx.x] = (uint32_t)cr;
    }
}

// exclusive scan of two short arrays (per-CTA face counts) in one single-CTA launch
__global__ void __launch_bounds__(1024) scan2_kernel(uint32_t *a, uint32_t *b, uint32_t n) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    for (int which = 0; which < 2; ++which) {
        uint32_t *x = which ? b : a;
        uint32_t carry = 0;
        for (uint32_t base = 0; base < n; base += 1024) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t v = i < n ? x[i] : 0u;
            uint32_t inc = v;
            const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off);
                if (lane >= off) inc += t;
            }
            if (lane == 31) wsum[w] = inc;
            __syncthreads();
            if (w == 0) {
                uint32_t t = wsum[lane];
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, t, off);
                    if (lane >= off) t += u;
                }
                wsum[lane] = t;
                if (lane == 31) carry_s = t;
            }
            __syncthreads();
            const uint32_t before = w ? wsum[w - 1] : 0u;
            if (i < n) x[i] = carry + before + inc - v;
            carry += carry_s;
            __syncthreads();
        }
    }
}

// send buffers: record 0 is a header whose .x carries the record count
__global__ void slab_fate_emit_kernel(const GridDesc g, int xs0, int xs1, int rank, int world,
                                      const float4 *__restrict__ pos, float4 *__restrict__ vel, uint32_t n,
                                      const uint32_t *__restrict__ off_l, const uint32_t *__restrict__ off_r,
                                      uint32_t nblocks, float4 *__restrict__ sl_pos, float4 *__restrict__ sl_vel,
                                      float4 *__restrict__ sr_pos, float4 *__restrict__ sr_vel,
                                      uint32_t halo_cap, unsigned *__restrict__ status) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    Fate f;
    f.send[0] = f.send[1] = false;
    float4 p = make_float4(0, 0, 0, 0), v = p;
    if (i < n) {
        p = pos[i];
        v = vel[i];
        f = fate_of(g, xs0, xs1, rank, world, p.x, __float_as_uint(v.w));
    }
    const uint32_t rl = block_rank(f.send[0]);
    const uint32_t rr = block_rank(f.send[1]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // headers: totals live one past the last block offset
        const uint32_t tl = min(off_l[nblocks], halo_cap), tr = min(off_r[nblocks], halo_cap);
        if (off_l[nblocks] > halo_cap || off_r[nblocks] > halo_cap) atomicOr(status, 8u);
        sl_pos[0] = make_float4(__uint_as_float(tl), 0, 0, 0);
        sr_pos[0] = make_float4(__uint_as_float(tr), 0, 0, 0);
        sl_vel[0] = sr_vel[0] = make_float4(0, 0, 0, __uint_as_float(REC_DEAD));
    }
    if (i >= n) return;
    if (f.send[0]) {
        const uint32_t d = off_l[blockIdx.x] + rl;
        if (d < halo_cap) {
            sl_pos[1 + d] = p;
            sl_vel[1 + d] = make_float4(v.x, v.y, v.z, __uint_as_float(f.as[0]));
        }
    }
    if (f.send[1]) {
        const uint32_t d = off_r[blockIdx.x] + rr;
        if (d < halo_cap) {
            sr_pos[1 + d] = p;
            sr_vel[1 + d] = make_float4(v.x, v.y, v.z, __uint_as_float(f.as[1]));
        }
    }
    if (f.jumped) atomicOr(status, 16u);
    vel[i] = make_float4(v.x, v.y, v.z, __uint_as_float(f.keep));
}

// keys of residents + the two receive regions (each: header record, then halo_cap records)
__global__ void slab_keys_kernel(const GridDesc g, const float4 *__restrict__ pos, float4 *__restrict__ vel,
                                 uint32_t n_res, uint32_t halo_cap, uint32_t *__restrict__ keys,
                                 uint32_t *__restrict__ cell_count) {
    const uint32_t m = n_res + 2 * (halo_cap + 1);
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    const bool valid = i < m;
    uint32_t key = 0xffffffffu;
    if (valid) {
        float4 v = vel[i];
        uint32_t flag = __float_as_uint(v.w);
        if (i >= n_res) {
            const uint32_t region = (i - n_res) / (halo_cap + 1), k = (i - n_res) % (halo_cap + 1);
            const uint32_t count = __float_as_uint(pos[n_res + region * (halo_cap + 1)].x);
            if (k == 0 || k - 1 >= count || flag > REC_GHOST) flag = REC_DEAD;
        }
        if (flag >= REC_DEAD) {
            key = g.ncells;
            vel[i] = make_float4(v.x, v.y, v.z, __uint_as_float(REC_DEAD));
        } else {
            key = cell_key_of(g, pos[i]);
        }
        keys[i] = key;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (valid && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(cell_count + key, __popc(peers));
}

// ---- representation changes --------------------------------------------------------------
static int gather_slice(Shard *s, fp_flock *f) {  // REP_SLICE -> all_*[acur]
    const size_t off = (size_t)s->rank * s->per;
    FP_CUDA(cudaMemcpyAsync(s->all_pos[s->acur] + off, f->pos[f->cur], (size_t)s->n_slice * sizeof(float4),
                            cudaMemcpyDeviceToDevice, f->stream));
    FP_CUDA(cudaMemcpyAsync(s->all_vel[s->acur] + off, f->vel[f->cur], (size_t)s->n_slice * sizeof(float4),
                            cudaMemcpyDeviceToDevice, f->stream));
    FP_NCCL(s, s->api.GroupStart());
    FP_NCCL(s, s->api.AllGather(s->all_pos[s->acur] + off, s->all_pos[s->acur], (size_t)s->per * 4, ncclFloat32,
                                s->comm, f->stream));
    FP_NCCL(s, s->api.AllGather(s->all_vel[s->acur] + off, s->all_vel[s->acur], (size_t)s->per * 4, ncclFloat32,
                                s->comm, f->stream));
    FP_NCCL(s, s->api.GroupEnd());
    return FP_OK;
}

// make all_*[acur] the current flock in global index order on every rank
static int to_global(Shard *s, fp_flock *f) {
    if (s->global_valid) return FP_OK;
    if (s->rep == REP_SLICE) {
        if (f->permuted) {
            int rc = launch_unpermute(f->stream, f->pos[f->cur], f->vel[f->cur], f->pos[f->cur ^ 1],
                                      f->vel[f->cur ^ 1], f->n, s->first);
            if (rc) return rc;
            f->cur ^= 1;
            f->permuted = false;
        }
        int rc = gather_slice(s, f);
        if (rc) return rc;
    } else {
        // owned boids into a zeroed flock, then a bit-preserving sum (every row has one writer)
        const size_t total = (size_t)s->per * s->world;
        FP_CUDA(cudaMemsetAsync(s->all_pos[s->acur], 0, total * sizeof(float4), f->stream));
        FP_CUDA(cudaMemsetAsync(s->all_vel[s->acur], 0, total * sizeof(float4), f->stream));
        if (f->n) {
            scatter_owned_kernel<<<nblk(f->n), SB, 0, f->stream>>>(f->pos[f->cur], f->vel[f->cur], f->n,
                                                                   s->all_pos[s->acur], s->all_vel[s->acur]);
            count_launch();
            FP_CUDA(cudaGetLastError());
        }
        FP_NCCL(s, s->api.GroupStart());
        FP_NCCL(s, s->api.AllReduce(s->all_pos[s->acur], s->all_pos[s->acur], total * 4, ncclUint32, ncclSum,
                                    s->comm, f->stream));
        FP_NCCL(s, s->api.AllReduce(s->all_vel[s->acur], s->all_vel[s->acur], total * 4, ncclUint32, ncclSum,
                                    s->comm, f->stream));
        FP_NCCL(s, s->api.GroupEnd());
    }
    s->global_valid = true;
    return FP_OK;
}

static int ensure_local_cap(fp_flock *f, uint32_t cap) {
    if (cap <= f->cap) return FP_OK;
    FP_CUDA(cudaStreamSynchronize(f->stream));
    for (int b = 0; b < 2; ++b) {
        float4 *np = nullptr, *nv = nullptr;
        FP_CUDA(cudaMalloc((void **)&np, (size_t)cap * sizeof(float4)));
        FP_CUDA(cudaMalloc((void **)&nv, (size_t)cap * sizeof(float4)));
        FP_CUDA(cudaMemcpy(np, f->pos[b], (size_t)f->cap * sizeof(float4), cudaMemcpyDeviceToDevice));
        FP_CUDA(cudaMemcpy(nv, f->vel[b], (size_t)f->cap * sizeof(float4), cudaMemcpyDeviceToDevice));
        cudaFree(f->pos[b]);
        cudaFree(f->vel[b]);
        f->pos[b] = np;
        f->vel[b] = nv;
    }
    f->cap = cap;
    return FP_OK;
}

static int to_slice(Shard *s, fp_flock *f) {  // any -> REP_SLICE in f->pos[cur]
    if (s->rep == REP_SLICE) return FP_OK;
    int rc = to_global(s, f);
    if (rc) return rc;
    if ((rc = ensure_local_cap(f, s->n_slice))) return rc;
    const size_t off = (size_t)s->rank * s->per;
    FP_CUDA(cudaMemcpyAsync(f->pos[f->cur], s->all_pos[s->acur] + off, (size_t)s->n_slice * sizeof(float4),
                            cudaMemcpyDeviceToDevice, f->stream));
    FP_CUDA(cudaMemcpyAsync(f->vel[f->cur], s->all_vel[s->acur] + off, (size_t)s->n_slice * sizeof(float4),
                            cudaMemcpyDeviceToDevice, f->stream));
    f->n = s->n_slice;
    f->permuted = false;
    s->rep = REP_SLICE;
    return FP_OK;
}

static int ensure_blk(Shard *s, size_t blocks) {
    if (blocks + 1 <= s->blk_cap) return FP_OK;
    for (int b = 0; b < 2; ++b) {
        cudaFree(s->blk_cnt[b]);
        int rc = dalloc(&s->blk_cnt[b], blocks + 1);
        if (rc) return rc;
    }
    s->blk_cap = blocks + 1;
    return FP_OK;
}

// Called by fit_grid with f->grid = the GLOBAL grid.  Splits its x layers into slabs, builds
// this rank's local grid (slab + one halo layer on each side) and repartitions the flock.
int shard_grid_fitted(Shard *s, fp_flock *f) {
    int rc = to_global(s, f);  // from the old representation, before anything is re-laid out
    if (rc) return rc;
    const GridDesc G = f->grid;
    if (G.dim[0] < s->world) {
        set_error("slab sharding needs at least one cell layer along x per rank");
        return FP_ERR_UNSUPPORTED;
    }
    s->xs0 = (int)((int64_t)G.dim[0] * s->rank / s->world);
    s->xs1 = (int)((int64_t)G.dim[0] * (s->rank + 1) / s->world);
    GridDesc L = G;
    L.gdimx = G.dim[0];
    L.xoff = s->xs0 - 1;
    L.dim[0] = std::max(s->xs1 - s->xs0, 0) + 2;
    L.ncells = (uint32_t)((uint64_t)L.dim[0] * L.dim[1] * L.dim[2]);
    uint32_t bits = 1;
    while ((1ull << bits) < (uint64_t)L.ncells + 1) ++bits;  // + the dead-record key
    L.key_bits = bits;
    s->lgrid = L;
    f->grid = L;

    // select this rank's slab from the global flock (index order)
    const uint32_t n = (uint32_t)s->n_global;
    const unsigned nb = nblk(n);
    if ((rc = ensure_blk(s, nb))) return rc;
    // scratch for the scan: reuse the grid work area, sized below; a private temp is simpler here
    uint32_t *tmp = nullptr;
    if ((rc = dalloc(&tmp, (size_t)nb / 4096 + 2))) return rc;
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[0], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    slab_select_count_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, s->all_pos[s->acur], n, s->blk_cnt[0]);
    count_launch();
    rc = launch_exclusive_scan(f->stream, s->blk_cnt[0], (size_t)nb + 1, tmp);
    if (rc) { cudaFree(tmp); return rc; }
    FP_CUDA(cudaMemcpyAsync(s->h_live, s->blk_cnt[0] + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    const uint32_t n_own = s->h_live[0];
    const uint64_t need = (uint64_t)n_own + n_own / 4 + 65536;  // halo room is added once it is sized
    if ((rc = ensure_local_cap(f, (uint32_t)std::min<uint64_t>(need, 0x7fffffffu)))) { cudaFree(tmp); return rc; }
    slab_select_emit_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, s->all_pos[s->acur], s->all_vel[s->acur],
                                                     n, s->blk_cnt[0], f->pos[f->cur], f->vel[f->cur], f->cap,
                                                     f->d_status);
    count_launch();
    FP_CUDA(cudaGetLastError());
    FP_CUDA(cudaStreamSynchronize(f->stream));
    cudaFree(tmp);
    f->n = n_own;
    f->permuted = true;
    s->rep = REP_SLAB;

    // Face messages: size them from the boundary layers as they are now -- the largest face of
    // any rank (both ends of a message must agree on the size), with 50 % head-room for drift
    // until the next re-fit.  An overflow raises FP_STATUS_HALO_OVERFLOW.
    {
        const unsigned nbo = nblk(std::max(n_own, 1u));
        if ((rc = ensure_blk(s, nbo))) return rc;
        FP_CUDA(cudaMemsetAsync(s->blk_cnt[0], 0, ((size_t)nbo + 1) * sizeof(uint32_t), f->stream));
        FP_CUDA(cudaMemsetAsync(s->blk_cnt[1], 0, ((size_t)nbo + 1) * sizeof(uint32_t), f->stream));
        slab_fate_count_kernel<<<nbo, SB, 0, f->stream>>>(L, s->xs0, s->xs1, s->rank, s->world, f->pos[f->cur],
                                                        f->vel[f->cur], n_own, s->blk_cnt[0], s->blk_cnt[1]);
        scan2_kernel<<<1, 1024, 0, f->stream>>>(s->blk_cnt[0], s->blk_cnt[1], nbo + 1);
        count_launch(2);
        // max(left face, right face) on this rank, then max over ranks, on the device
        uint32_t *d2 = s->blk_cnt[0] + nbo;  // total of array 0; fold array 1's total into it
        FP_CUDA(cudaMemcpyAsync(s->h_live, s->blk_cnt[0] + nbo, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
        FP_CUDA(cudaMemcpyAsync(s->h_live + 1, s->blk_cnt[1] + nbo, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
        s->h_live[2] = std::max(s->h_live[0], s->h_live[1]);
        FP_CUDA(cudaMemcpyAsync(d2, s->h_live + 2, sizeof(uint32_t), cudaMemcpyHostToDevice, f->stream));
        FP_NCCL(s, s->api.AllReduce(d2, d2, 1, ncclUint32, ncclMax, s->comm, f->stream));
        FP_CUDA(cudaMemcpyAsync(s->h_live + 3, d2, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
        const uint64_t face = s->h_live[3];
        s->halo_cap = (uint32_t)std::min<uint64_t>(s->n_global, face + face / 2 + 4096);
        for (int b = 0; b < 2; ++b) {
            cudaFree(s->send_pos[b]);
            cudaFree(s->send_vel[b]);
            if ((rc = dalloc(&s->send_pos[b], (size_t)s->halo_cap + 1)) ||
                (rc = dalloc(&s->send_vel[b], (size_t)s->halo_cap + 1)))
                return rc;
        }
        const uint64_t need2 = (uint64_t)n_own + n_own / 4 + 4ull * (s->halo_cap + 1) + 65536;
        if ((rc = ensure_local_cap(f, (uint32_t)std::min<uint64_t>(need2, 0x7fffffffu)))) return rc;
    }

    // grid scratch for the largest array a step can sort
    GridWork &w = f->work;
    const uint32_t cap = f->cap;
    const size_t ntiles = ((size_t)cap + 4095) / 4096 + 1;
    const size_t hist = 256 * ntiles;
    const size_t cells = (size_t)L.ncells + 2;
    const size_t scan_tmp = std::max(hist, cells) / 4096 + 2;
    auto grow = [&](uint32_t *&p, size_t have, size_t want) -> int {
        if (want <= have) return FP_OK;
        cudaFree(p);
        return dalloc(&p, want);
    };
    if ((rc = grow(w.keys[0], w.cap, cap)) || (rc = grow(w.keys[1], w.cap, cap)) ||
        (rc = grow(w.vals[0], w.cap, cap)) || (rc = grow(w.vals[1], w.cap, cap)))
        return rc;
    w.cap = std::max(w.cap, cap);
    if (cap + 8 > w.soa_cap) {
        for (auto &b : w.soa)
            for (auto &p : b) {
                cudaFree(p);
                if ((rc = dalloc(&p, (size_t)cap + 8))) return rc;
                FP_CUDA(cudaMemsetAsync(p, 0, ((size_t)cap + 8) * sizeof(float), f->stream));
            }
        w.soa_cap = cap + 8;
    }
    if ((rc = grow(w.tile_hist, w.tile_hist_elems, hist))) return rc;
    w.tile_hist_elems = std::max(w.tile_hist_elems, hist);
    if ((rc = grow(w.cell_start, w.cell_cap, cells))) return rc;
    w.cell_cap = std::max(w.cell_cap, cells);
    if ((rc = grow(w.scan_tmp, w.scan_tmp_elems, scan_tmp))) return rc;
    w.scan_tmp_elems = std::max(w.scan_tmp_elems, scan_tmp);
    f->grid_valid = true;
    f->steps_since_fit = 0;
    return FP_OK;
}

// ---- one slab step up to the sorted arrays ------------------------------------------------
// in:  residents in f->pos[cur][0..f->n).  out: sorted live records in f->pos[cur^1][0..*n_live),
// cell table in f->work.cell_start.
static int slab_exchange_and_sort(Shard *s, fp_flock *f, uint32_t *n_live) {
    const GridDesc &L = s->lgrid;
    const uint32_t n = f->n, hc = s->halo_cap;
    const uint32_t m = n + 2 * (hc + 1);
    if (m > f->cap) {
        set_error("slab capacity exceeded (flock too clustered for this many ranks)");
        return FP_ERR_UNSUPPORTED;
    }
    float4 *pos = f->pos[f->cur], *vel = f->vel[f->cur];
    const unsigned nb = nblk(std::max(n, 1u));
    int rc = ensure_blk(s, nb);
    if (rc) return rc;
    // fates: deterministic two-pass compaction into the face messages
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[0], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[1], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    slab_fate_count_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, s->rank, s->world, pos, vel, n,
                                                    s->blk_cnt[0], s->blk_cnt[1]);
    count_launch();
    scan2_kernel<<<1, 1024, 0, f->stream>>>(s->blk_cnt[0], s->blk_cnt[1], nb + 1);
    count_launch();
    slab_fate_emit_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, s->rank, s->world, pos, vel, n,
                                                   s->blk_cnt[0], s->blk_cnt[1], nb, s->send_pos[0],
                                                   s->send_vel[0], s->send_pos[1], s->send_vel[1], hc,
                                                   f->d_status);
    count_launch();
    FP_CUDA(cudaGetLastError());
    // receive regions sit right behind the residents: [left: header + hc][right: header + hc]
    float4 *rl_pos = pos + n, *rl_vel = vel + n, *rr_pos = pos + n + (hc + 1), *rr_vel = vel + n + (hc + 1);
    FP_CUDA(cudaMemsetAsync(rl_pos, 0, sizeof(float4), f->stream));  // count 0 unless a neighbour says otherwise
    FP_CUDA(cudaMemsetAsync(rr_pos, 0, sizeof(float4), f->stream));
    const size_t words = ((size_t)hc + 1) * 4;
    FP_NCCL(s, s->api.GroupStart());
    if (s->rank > 0) {
        FP_NCCL(s, s->api.Send(s->send_pos[0], words, ncclFloat32, s->rank - 1, s->comm, f->stream));
        FP_NCCL(s, s->api.Send(s->send_vel[0], words, ncclFloat32, s->rank - 1, s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(rl_pos, words, ncclFloat32, s->rank - 1, s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(rl_vel, words, ncclFloat32, s->rank - 1, s->comm, f->stream));
    }
    if (s->rank < s->world - 1) {
        FP_NCCL(s, s->api.Send(s->send_pos[1], words, ncclFloat32, s->rank + 1, s->comm, f->stream));
        FP_NCCL(s, s->api.Send(s->send_vel[1], words, ncclFloat32, s->rank + 1, s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(rr_pos, words, ncclFloat32, s->rank + 1, s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(rr_vel, words, ncclFloat32, s->rank + 1, s->comm, f->stream));
    }
    FP_NCCL(s, s->api.GroupEnd());
    // keys (dead records -> key ncells), cell table, sort, gather
    GridWork &w = f->work;
    FP_CUDA(cudaMemsetAsync(w.cell_start, 0, ((size_t)L.ncells + 2) * sizeof(uint32_t), f->stream));
    slab_keys_kernel<<<nblk(m), SB, 0, f->stream>>>(L, pos, vel, n, hc, w.keys[0], w.cell_start);
    count_launch();
    FP_CUDA(cudaGetLastError());
    if ((rc = launch_exclusive_scan(f->stream, w.cell_start, (size_t)L.ncells + 2, w.scan_tmp))) return rc;
    FP_CUDA(cudaMemcpyAsync(s->h_live, w.cell_start + L.ncells, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                            f->stream));
    FP_CUDA(cudaEventRecord(s->ev_live, f->stream));
    int buf = 0;
    if ((rc = launch_radix_sort(f->stream, w, m, L.key_bits, &buf))) return rc;
    // The live count sizes the gather and the walk.  Wait for its copy only: the sort queued
    // behind it keeps the GPU busy while the host enqueues what follows.
    FP_CUDA(cudaEventSynchronize(s->ev_live));
    *n_live = s->h_live[0];
    w.home = w.keys[buf];
    return launch_grid_reorder(f->stream, w.vals[buf], pos, vel, f->pos[f->cur ^ 1], f->vel[f->cur ^ 1], w.soa[0],
                               *n_live, nullptr);
}

// sorted records in pos/vel[which], SoA copy 0; TAP_STEP output into the other buffer
static WalkIO slab_walk_io(fp_flock *f, int which, uint32_t n_live, bool stepping) {
    GridWork &w = f->work;
    WalkIO io{};
    io.pos_s = f->pos[which];
    io.vel_s = f->vel[which];
    for (int a = 0; a < 3; ++a) io.soa_in[a] = w.soa[0][a];
    io.home = w.home;
    io.cell_start = w.cell_start;
    io.n_all = n_live;
    if (stepping) {
        io.pos_out = f->pos[which ^ 1];
        io.vel_out = f->vel[which ^ 1];
        for (int a = 0; a < 3; ++a) io.soa_out[a] = w.soa[1][a];
    }
    return io;
}

static int ensure_slab(Shard *s, fp_flock *f) {
    if (!f->grid_valid || (!f->domain_user && f->steps_since_fit >= 256) || s->rep != REP_SLAB) {
        // bounds come from whatever the rank holds now (ghosts are real boids elsewhere: harmless)
        int rc = flock_fit_grid(f);  // -> shard_grid_fitted -> repartition
        if (rc) return rc;
    }
    return FP_OK;
}

static int allpairs_prepare(Shard *s, fp_flock *f) {
    int rc = to_slice(s, f);
    if (rc) return rc;
    return to_global(s, f);
}

int shard_settle(Shard *, fp_flock *) { return FP_OK; }

int shard_step(Shard *s, fp_flock *f, uint32_t nsteps) {
    const int m = shard_method(s, f->method, f->cfg);
    int rc;
    if (m == FP_METHOD_GRID) {
        if ((rc = ensure_slab(s, f))) return rc;
        for (uint32_t k = 0; k < nsteps; ++k) {
            flock_select_leads(f);
            if ((rc = flock_mark(f))) return rc;
            uint32_t n_live = 0;
            if ((rc = slab_exchange_and_sort(s, f, &n_live))) return rc;
            if ((rc = flock_mark(f))) return rc;
            rc = launch_grid_walk(f->stream, f->P, s->lgrid, TAP_STEP, slab_walk_io(f, f->cur ^ 1, n_live, true),
                                  f->d_status, TapOut{});
            if (rc) return rc;
            if ((rc = flock_mark(f))) return rc;
            f->n = n_live;
            s->global_valid = false;
            ++f->steps_since_fit;
            ++f->table_cursor;
        }
        return FP_OK;
    }
    if ((rc = allpairs_prepare(s, f))) return rc;
    const size_t off = (size_t)s->rank * s->per;
    for (uint32_t k = 0; k < nsteps; ++k) {
        flock_select_leads(f);
        if ((rc = flock_mark(f)) || (rc = flock_mark(f))) return rc;
        const int a = s->acur;
        rc = launch_allpairs(f->stream, f->P, TAP_STEP, s->all_pos[a], s->all_vel[a], (uint32_t)s->n_global,
                             s->first, s->n_slice, s->all_pos[a ^ 1] + off, s->all_vel[a ^ 1] + off, f->d_status,
                             TapOut{});
        if (rc) return rc;
        if ((rc = flock_mark(f))) return rc;
        FP_NCCL(s, s->api.GroupStart());
        FP_NCCL(s, s->api.AllGather(s->all_pos[a ^ 1] + off, s->all_pos[a ^ 1], (size_t)s->per * 4, ncclFloat32,
                                    s->comm, f->stream));
        FP_NCCL(s, s->api.AllGather(s->all_vel[a ^ 1] + off, s->all_vel[a ^ 1], (size_t)s->per * 4, ncclFloat32,
                                    s->comm, f->stream));
        FP_NCCL(s, s->api.GroupEnd());
        s->acur ^= 1;
        ++f->table_cursor;
    }
    // the slice is the authoritative local form between calls
    FP_CUDA(cudaMemcpyAsync(f->pos[f->cur], s->all_pos[s->acur] + off, (size_t)s->n_slice * sizeof(float4),
                            cudaMemcpyDeviceToDevice, f->stream));
    FP_CUDA(cudaMemcpyAsync(f->vel[f->cur], s->all_vel[s->acur] + off, (size_t)s->n_slice * sizeof(float4),
                            cudaMemcpyDeviceToDevice, f->stream));
    s->global_valid = true;
    return FP_OK;
}

// sum the per-rank tap outputs (each row has exactly one writer; the rest are zero bits)
static int reduce_tap(Shard *s, fp_flock *f, int tap, const TapOut &out) {
    const size_t n = s->n_global;
    FP_NCCL(s, s->api.GroupStart());
    if (tap == TAP_ACCEL) {
        FP_NCCL(s, s->api.AllReduce(out.accel3, out.accel3, n * 3, ncclUint32, ncclSum, s->comm, f->stream));
        if (out.comp15)
            FP_NCCL(s, s->api.AllReduce(out.comp15, out.comp15, n * 15, ncclUint32, ncclSum, s->comm, f->stream));
    } else if (tap == TAP_NEIGHBORS) {
        FP_NCCL(s, s->api.AllReduce(out.nbr_count, out.nbr_count, n, ncclUint32, ncclSum, s->comm, f->stream));
        FP_NCCL(s, s->api.AllReduce(out.nbr_hash, out.nbr_hash, n, ncclUint64, ncclSum, s->comm, f->stream));
    } else if (tap == TAP_CENSUS) {
        FP_NCCL(s, s->api.AllReduce(out.census, out.census, 4, ncclUint64, ncclSum, s->comm, f->stream));
    }
    FP_NCCL(s, s->api.GroupEnd());
    return FP_OK;
}

int shard_tap(Shard *s, fp_flock *f, int tap, const TapOut &out) {
    const int m = shard_method(s, f->method, f->cfg);
    int rc;
    if (m == FP_METHOD_GRID) {
        if ((rc = ensure_slab(s, f))) return rc;
        uint32_t n_live = 0;
        if ((rc = slab_exchange_and_sort(s, f, &n_live))) return rc;
        f->cur ^= 1;  // the sorted copy (with this step's ghosts) becomes the resident array
        f->n = n_live;
        rc = launch_grid_walk(f->stream, f->P, s->lgrid, tap, slab_walk_io(f, f->cur, n_live, false), f->d_status,
                              out);
        if (rc) return rc;
    } else {
        if ((rc = allpairs_prepare(s, f))) return rc;
        rc = launch_allpairs(f->stream, f->P, tap, s->all_pos[s->acur], s->all_vel[s->acur],
                             (uint32_t)s->n_global, s->first, s->n_slice, nullptr, nullptr, f->d_status, out);
        if (rc) return rc;
    }
    return reduce_tap(s, f, tap, out);
}

// ---- the boids this rank owns, compacted on the device (deterministic order) ----------------
__global__ void owned_count_kernel(const float4 *__restrict__ vel, uint32_t n, uint32_t *__restrict__ blk) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    const bool own = i < n && __float_as_uint(vel[i].w) == REC_OWNED;
    const int c = __syncthreads_count(own);
    if (threadIdx.x == 0) blk[blockIdx.x] = (uint32_t)c;
}
__global__ void owned_emit_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ vel, uint32_t n,
                                  const uint32_t *__restrict__ blk_off, unsigned long long *__restrict__ idx,
                                  float *__restrict__ aos6) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    float4 p = make_float4(0, 0, 0, 0), v = p;
    bool own = false;
    if (i < n) {
        p = pos[i];
        v = vel[i];
        own = __float_as_uint(v.w) == REC_OWNED;
    }
    const uint32_t r = block_rank(own);
    if (!own) return;
    const size_t d = blk_off[blockIdx.x] + r;
    idx[d] = __float_as_uint(p.w);
    float *o = aos6 + 6 * d;
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
    o[3] = v.x; o[4] = v.y; o[5] = v.z;
}

// count (and, with outputs, fetch) the owned records of the resident array
int shard_read_local(Shard *s, fp_flock *f, uint64_t *n_local, uint64_t *out_index, float *out_aos6) {
    const uint32_t n = f->n;
    *n_local = 0;
    if (!n) return FP_OK;
    const unsigned nb = nblk(n);
    int rc = ensure_blk(s, nb);
    if (rc) return rc;
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[0], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[1], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    owned_count_kernel<<<nb, SB, 0, f->stream>>>(f->vel[f->cur], n, s->blk_cnt[0]);
    scan2_kernel<<<1, 1024, 0, f->stream>>>(s->blk_cnt[0], s->blk_cnt[1], nb + 1);
    count_launch(2);
    FP_CUDA(cudaMemcpyAsync(s->h_live, s->blk_cnt[0] + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    const uint32_t own = s->h_live[0];
    *n_local = own;
    if (!out_index || !out_aos6 || !own) return FP_OK;
    const size_t bytes = (size_t)own * (6 * sizeof(float) + sizeof(unsigned long long));
    if (bytes > f->stage_bytes) {
        if (f->d_stage) cudaFree(f->d_stage);
        f->d_stage = nullptr;
        f->stage_bytes = 0;
        FP_CUDA(cudaMalloc(&f->d_stage, bytes));
        f->stage_bytes = bytes;
    }
    unsigned long long *d_idx = (unsigned long long *)f->d_stage;
    float *d_aos = (float *)(d_idx + own);
    owned_emit_kernel<<<nb, SB, 0, f->stream>>>(f->pos[f->cur], f->vel[f->cur], n, s->blk_cnt[0], d_idx, d_aos);
    count_launch();
    FP_CUDA(cudaGetLastError());
    FP_CUDA(cudaMemcpyAsync(out_index, d_idx, (size_t)own * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            f->stream));
    FP_CUDA(cudaMemcpyAsync(out_aos6, d_aos, (size_t)own * 6 * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

int shard_read_state(Shard *s, fp_flock *f, float *out) {
    int rc = to_global(s, f);
    if (rc) return rc;
    if (!s->n_global) return FP_OK;
    if (!out) {
        set_error("null output");
        return FP_ERR_INVALID;
    }
    const size_t bytes = (size_t)s->n_global * 6 * sizeof(float);
    if (bytes > f->stage_bytes) {
        if (f->d_stage) cudaFree(f->d_stage);
        f->d_stage = nullptr;
        f->stage_bytes = 0;
        FP_CUDA(cudaMalloc(&f->d_stage, bytes));
        f->stage_bytes = bytes;
    }
    if ((rc = launch_soa_to_aos6(f->stream, s->all_pos[s->acur], s->all_vel[s->acur], (float *)f->d_stage,
                                 (uint32_t)s->n_global, 0, 0)))
        return rc;
    FP_CUDA(cudaMemcpyAsync(out, f->d_stage, bytes, cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

}  // namespace fp

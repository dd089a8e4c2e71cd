// fp_shard.cu -- one flock over several GPUs (one process per GPU, NCCL over NVLink).
//
// Two partitions, both exact:
//  * all-pairs: rows i are independent given the full old state (Jacobi update,
//    flocking.rs:101-122).  Ranks own contiguous boid-index ranges and exchange
//    the new positions/velocities with ONE ncclAllGather per step; every rank
//    sums j = 0..N-1 in index order, so results are bit-identical to one GPU.
//  * grid: ranks own slabs of whole cell layers along x.  Cell keys are x-slowest, so
//    every layer is one contiguous slot range of the sorted arrays and a rank's records are
//    laid out [left ghost layer | owned layers | right ghost layer].  A BINNING (every few dozen
//    steps, see "lazy re-binning" in fp_api.cu) moves the boids that left the slab to
//    their new owner (fixed-size ncclSend/ncclRecv, counts in a header record), sorts the
//    owned records, and copies each boundary layer verbatim into the neighbour's ghost block
//    (per-cell counts + records, contiguous, no packing).  Between binnings the owner's walk
//    kernel writes the advanced state of its boundary boids straight into the neighbours'
//    ghost blocks with peer stores over NVLink (cudaIpc-mapped buffers): the halo exchange
//    is fused into the influence kernel.  A mailbox of (step tag, max speed) per rank, posted
//    with peer stores after each walk, is the step barrier and the all-reduce of the
//    displacement bound at once; no host round trip, no NCCL call in a step.  When peer
//    mapping is unavailable the same blocks travel by ncclSend/ncclRecv after each walk.
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
    uint32_t mig_cap = 0;      // records per migration message (plus one header record)
    float4 *send_pos[2] = {nullptr, nullptr}, *send_vel[2] = {nullptr, nullptr};  // [0] left, [1] right
    uint32_t *blk_cnt[2] = {nullptr, nullptr};
    size_t blk_cap = 0;
    uint32_t *h_live = nullptr;  // pinned scratch (16 words)
    cudaEvent_t ev_live = nullptr;  // completion of a count read-back
    // current binning: records [0, n_all) = [ghost L | owned | ghost R]
    uint32_t own0 = 0, own_n = 0, n_all = 0;
    uint32_t lay_first[2] = {0, 0}, lay_last[2] = {0, 0};  // slot ranges of my first / last owned layer
    struct Layout { uint32_t nL, nO, nR, cur, soa_cur, next_fits, pad[2]; };
    bool next_fits = true;     // every rank's NEXT binning fits its buffers (agreed at the last one)
    Layout *h_layout = nullptr;  // pinned, one per rank (all-gathered at every binning)
    uint32_t *d_layout = nullptr;
    // peer mapping (cudaIpc): neighbours' state buffers, every rank's mailbox
    bool peer_ok = false;
    Mail *mail = nullptr;                 // my mailbox, FP_MAIL_SLOTS entries
    MailPeers mail_peers{};
    void *nbr_buf[2][10] = {};            // [face][pos0 pos1 vel0 vel1 soa0xyz soa1xyz] of the neighbour
    std::vector<void *> opened;           // every pointer cudaIpcOpenMemHandle returned
    unsigned char *d_ipc = nullptr;       // all-gather staging for the handles
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
    if (cudaMallocHost((void **)&s->h_live, 16 * sizeof(uint32_t)) != cudaSuccess ||
        cudaMallocHost((void **)&s->h_layout, world * sizeof(Shard::Layout)) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "cudaMallocHost", __FILE__, __LINE__);
    if ((rc = dalloc(&s->d_layout, (size_t)world * sizeof(Shard::Layout) / 4)) ||
        (rc = dalloc(&s->mail, (size_t)FP_MAIL_SLOTS)))
        return rc;
    if (cudaMemset(s->mail, 0, FP_MAIL_SLOTS * sizeof(Mail)) != cudaSuccess)
        return cuda_fail(cudaGetLastError(), "cudaMemset", __FILE__, __LINE__);
    if (world > FP_MAX_WORLD) {
        set_error("at most 16 ranks");
        return FP_ERR_UNSUPPORTED;
    }
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
    for (void *p : s->opened) cudaIpcCloseMemHandle(p);
    cudaFree(s->d_layout);
    cudaFree(s->mail);
    cudaFree(s->d_ipc);
    if (s->h_live) cudaFreeHost(s->h_live);
    if (s->h_layout) cudaFreeHost(s->h_layout);
    if (s->ev_live) cudaEventDestroy(s->ev_live);
    if (s->comm) s->api.CommDestroy(s->comm);
    delete s;
}


int shard_world(const Shard *s) { return s->world; }
bool shard_is_slab(const Shard *s) { return s->rep == REP_SLAB; }

void shard_info(const Shard *s, int *rank, int *world, int *peer_mapped) {
    if (rank) *rank = s->rank;
    if (world) *world = s->world;
    if (peer_mapped) *peer_mapped = s->peer_ok ? 1 : 0;
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

enum { REC_OWNED = 0u, REC_DEAD = 2u };  // vel.w inside a binning: still here / migrated away or unused

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

// Where an owned record goes at a binning: it stays, or its owner becomes the left / right
// neighbour because its position now lies in that rank's layers.
struct Fate {
    int dir;      // -1 stays, 0 to the left neighbour, 1 to the right
    bool jumped;  // crossed more than one slab since the last binning
};
__device__ __forceinline__ Fate fate_of(const GridDesc &g, int xs0, int xs1, float x, uint32_t flag) {
    Fate f;
    f.dir = -1;
    f.jumped = false;
    if (flag != REC_OWNED) return f;
    const int gx = cell_coord(x, g.origin[0], g.inv_cell, g.gdimx);
    if (gx < xs0) {
        f.dir = 0;
        f.jumped = false;  // the neighbour re-examines it at ITS next binning if it is further left
    } else if (gx >= xs1) {
        f.dir = 1;
    }
    return f;
}

__global__ void slab_fate_count_kernel(const GridDesc g, int xs0, int xs1, const float4 *__restrict__ pos,
                                       const float4 *__restrict__ vel, uint32_t n,
                                       uint32_t *__restrict__ blk_l, uint32_t *__restrict__ blk_r) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    bool l = false, r = false;
    if (i < n) {
        const Fate f = fate_of(g, xs0, xs1, pos[i].x, __float_as_uint(vel[i].w));
        l = f.dir == 0;
        r = f.dir == 1;
    }
    const int cl = __syncthreads_count(l), cr = __syncthreads_count(r);
    if (threadIdx.x == 0) {
        blk_l[blockIdx.x] = (uint32_t)cl;
        blk_r[blockIdx.x] = (uint32_t)cr;
    }
}

// records of the first / last owned layer (sizes the buffers at a fit)
__global__ void slab_face_count_kernel(const GridDesc g, int xs0, int xs1, const float4 *__restrict__ pos,
                                       uint32_t n, uint32_t *__restrict__ out2) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    bool l = false, r = false;
    if (i < n) {
        const int gx = cell_coord(pos[i].x, g.origin[0], g.inv_cell, g.gdimx);
        l = gx == xs0;
        r = gx == xs1 - 1;
    }
    const int cl = __syncthreads_count(l), cr = __syncthreads_count(r);
    if (threadIdx.x == 0) {
        if (cl) atomicAdd(out2, (uint32_t)cl);
        if (cr) atomicAdd(out2 + 1, (uint32_t)cr);
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

// migration messages: record 0 is a header whose .x carries the record count
__global__ void slab_fate_emit_kernel(const GridDesc g, int xs0, int xs1, const float4 *__restrict__ pos,
                                      float4 *__restrict__ vel, uint32_t n,
                                      const uint32_t *__restrict__ off_l, const uint32_t *__restrict__ off_r,
                                      uint32_t nblocks, float4 *__restrict__ sl_pos, float4 *__restrict__ sl_vel,
                                      float4 *__restrict__ sr_pos, float4 *__restrict__ sr_vel,
                                      uint32_t cap, unsigned *__restrict__ status) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    Fate f;
    f.dir = -1;
    f.jumped = false;
    float4 p = make_float4(0, 0, 0, 0), v = p;
    if (i < n) {
        p = pos[i];
        v = vel[i];
        f = fate_of(g, xs0, xs1, p.x, __float_as_uint(v.w));
    }
    const uint32_t rl = block_rank(f.dir == 0);
    const uint32_t rr = block_rank(f.dir == 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // headers: totals live one past the last block offset
        const uint32_t tl = min(off_l[nblocks], cap), tr = min(off_r[nblocks], cap);
        if (off_l[nblocks] > cap || off_r[nblocks] > cap) atomicOr(status, 8u);
        sl_pos[0] = make_float4(__uint_as_float(tl), 0, 0, 0);
        sr_pos[0] = make_float4(__uint_as_float(tr), 0, 0, 0);
        sl_vel[0] = sr_vel[0] = make_float4(0, 0, 0, __uint_as_float(REC_DEAD));
    }
    if (i >= n || f.dir < 0) return;
    float4 *dp = f.dir ? sr_pos : sl_pos, *dv = f.dir ? sr_vel : sl_vel;
    const uint32_t d = f.dir ? off_r[blockIdx.x] + rr : off_l[blockIdx.x] + rl;
    if (d < cap) {
        dp[1 + d] = p;
        dv[1 + d] = make_float4(v.x, v.y, v.z, __uint_as_float(REC_OWNED));
        vel[i] = make_float4(v.x, v.y, v.z, __uint_as_float(REC_DEAD));  // it lives on the neighbour now
    }
}

// keys of residents + the two receive regions (each: header record, then cap records).
// Residents that migrated out and unused receive records get the key `ncells` (dead).
__global__ void slab_keys_kernel(const GridDesc g, int xs0, int xs1, const float4 *__restrict__ pos,
                                 float4 *__restrict__ vel, uint32_t n_res, uint32_t cap,
                                 uint32_t *__restrict__ keys, uint32_t *__restrict__ cell_count,
                                 unsigned *__restrict__ status) {
    const uint32_t m = n_res + 2 * (cap + 1);
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    const bool valid = i < m;
    uint32_t key = 0xffffffffu;
    if (valid) {
        float4 v = vel[i];
        uint32_t flag = __float_as_uint(v.w);
        if (i >= n_res) {
            const uint32_t region = (i - n_res) / (cap + 1), k = (i - n_res) % (cap + 1);
            const uint32_t count = __float_as_uint(pos[n_res + region * (cap + 1)].x);
            if (k == 0 || k - 1 >= count || flag != REC_OWNED) flag = REC_DEAD;
        }
        if (flag != REC_OWNED) {
            key = g.ncells;
        } else {
            const float4 p = pos[i];
            const int gx = cell_coord(p.x, g.origin[0], g.inv_cell, g.gdimx);
            if (gx < xs0 || gx >= xs1) atomicOr(status, 16u);  // crossed more than one slab per binning
            // (a flagged stray is kept in the nearest owned layer: ghost layers hold no owned record)
            const int cx = min(max(cell_coord_x(g, p.x), 1), g.dim[0] - 2);
            key = cell_key(g, cx, cell_coord(p.y, g.origin[1], g.inv_cell, g.dim[1]),
                           cell_coord(p.z, g.origin[2], g.inv_cell_z, g.dim[2]));
        }
        keys[i] = key;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (valid && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(cell_count + key, __popc(peers));
}

// sorted owned record i -> slot (ghost L count) + i of the new arrays; flags cleared
__global__ void slab_reorder_kernel(const uint32_t *__restrict__ vals, const float4 *__restrict__ pos_in,
                                    const float4 *__restrict__ vel_in, float4 *__restrict__ pos_out,
                                    float4 *__restrict__ vel_out, float *__restrict__ sx, float *__restrict__ sy,
                                    float *__restrict__ sz, const uint32_t *__restrict__ cell_start,
                                    uint32_t own_first_cell, uint32_t own_end_cell, uint32_t m, uint32_t cap) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    const uint32_t nL = cell_start[own_first_cell];
    const uint32_t nO = cell_start[own_end_cell] - nL;
    if (i >= m || i >= nO) return;
    const uint32_t src = vals[i];
    const float4 p = pos_in[src], v = vel_in[src];
    const uint32_t d = nL + i;
    if (d >= cap) return;  // (the host reports the overflow once it has read the counts)
    pos_out[d] = p;
    vel_out[d] = make_float4(v.x, v.y, v.z, 0.0f);
    sx[d] = p.x;
    sy[d] = p.y;
    sz[d] = p.z;
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
        if (s->own_n) {  // slab: the owned records are the slots [own0, own0 + own_n)
            scatter_owned_kernel<<<nblk(s->own_n), SB, 0, f->stream>>>(f->pos[f->cur] + s->own0,
                                                                       f->vel[f->cur] + s->own0, s->own_n,
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

// ---- peer mapping ------------------------------------------------------------------------
// Each rank exports its ten state buffers and its mailbox with cudaIpcGetMemHandle; the handles
// are all-gathered through NCCL and the neighbours' buffers / every rank's mailbox opened.
// Collective.  Falls back to NCCL halo messages (peer_ok = false) when any rank fails.
static int ipc_close(Shard *s) {
    for (void *p : s->opened) cudaIpcCloseMemHandle(p);
    s->opened.clear();
    memset(s->nbr_buf, 0, sizeof(s->nbr_buf));
    memset(&s->mail_peers, 0, sizeof(s->mail_peers));
    s->peer_ok = false;
    return FP_OK;
}

static int ipc_setup(Shard *s, fp_flock *f) {
    constexpr int NH = 11;
    struct Pack { cudaIpcMemHandle_t h[NH]; };
    static_assert(sizeof(Pack) % 4 == 0, "handles travel as uint32 words");
    static const bool want = [] {
        const char *e = getenv("FP_SHARD_PEER");
        return !(e && *e == '0');
    }();
    int ok = want ? 1 : 0;
    GridWork &w = f->work;
    void *mine[NH] = {f->pos[0], f->pos[1], f->vel[0], f->vel[1], w.soa[0][0], w.soa[0][1], w.soa[0][2],
                      w.soa[1][0], w.soa[1][1], w.soa[1][2], s->mail};
    std::vector<Pack> all(s->world);
    Pack me{};
    for (int k = 0; k < NH && ok; ++k)
        if (cudaIpcGetMemHandle(&me.h[k], mine[k]) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
        }
    if (!s->d_ipc) {
        int rc = dalloc(&s->d_ipc, sizeof(Pack) * (size_t)s->world);
        if (rc) return rc;
    }
    FP_CUDA(cudaMemcpyAsync(s->d_ipc + sizeof(Pack) * s->rank, &me, sizeof(Pack), cudaMemcpyHostToDevice, f->stream));
    FP_NCCL(s, s->api.AllGather(s->d_ipc + sizeof(Pack) * s->rank, s->d_ipc, sizeof(Pack) / 4, ncclUint32, s->comm,
                                f->stream));
    FP_CUDA(cudaMemcpyAsync(all.data(), s->d_ipc, sizeof(Pack) * s->world, cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    auto open = [&](const cudaIpcMemHandle_t &h) -> void * {
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        s->opened.push_back(p);
        return p;
    };
    if (ok) {
        for (int fc = 0; fc < 2 && ok; ++fc) {
            const int q = s->rank + (fc ? 1 : -1);
            if (q < 0 || q >= s->world) continue;
            for (int k = 0; k < 10 && ok; ++k)
                if (!(s->nbr_buf[fc][k] = open(all[q].h[k]))) ok = 0;
        }
        for (int q = 0; q < s->world && ok; ++q) {
            if (q == s->rank) s->mail_peers.box[q] = s->mail;
            else if (!(s->mail_peers.box[q] = (Mail *)open(all[q].h[10]))) ok = 0;
        }
    }
    // everyone or no one
    s->h_live[8] = (uint32_t)ok;
    uint32_t *d = s->d_layout;
    FP_CUDA(cudaMemcpyAsync(d, s->h_live + 8, sizeof(uint32_t), cudaMemcpyHostToDevice, f->stream));
    FP_NCCL(s, s->api.AllReduce(d, d, 1, ncclUint32, ncclMin, s->comm, f->stream));
    FP_CUDA(cudaMemcpyAsync(s->h_live + 9, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    s->peer_ok = s->h_live[9] != 0;
    if (!s->peer_ok) ipc_close(s);
    return FP_OK;
}

// Called by fit_grid with f->grid = the GLOBAL grid.  Splits its x layers into slabs, builds
// this rank's local grid (slab + one ghost layer on each side) and repartitions the flock.
int shard_grid_fitted(Shard *s, fp_flock *f) {
    int rc = to_global(s, f);  // from the old representation, before anything is re-laid out
    if (rc) return rc;
    // nobody may still map a buffer that is about to be re-allocated: close, then a barrier
    ipc_close(s);
    {
        uint32_t *d = s->d_layout;
        FP_NCCL(s, s->api.AllReduce(d, d, 1, ncclUint32, ncclMax, s->comm, f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
    }
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
    uint32_t *tmp = nullptr;
    if ((rc = dalloc(&tmp, (size_t)nb / 4096 + 2))) return rc;
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[0], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    slab_select_count_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, s->all_pos[s->acur], n, s->blk_cnt[0]);
    count_launch();
    rc = launch_exclusive_scan(f->stream, s->blk_cnt[0], (size_t)nb + 1, tmp);
    if (rc) { cudaFree(tmp); return rc; }
    // faces: records of the first / last owned layer of ANY rank bound the ghost blocks
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[1], 0, 2 * sizeof(uint32_t), f->stream));
    slab_face_count_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, s->all_pos[s->acur], n, s->blk_cnt[1]);
    count_launch();
    FP_CUDA(cudaMemcpyAsync(s->h_live, s->blk_cnt[0] + nb, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaMemcpyAsync(s->h_live + 1, s->blk_cnt[1], 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    const uint32_t n_own = s->h_live[0];
    s->h_live[3] = std::max(s->h_live[1], s->h_live[2]);
    {
        uint32_t *d = s->d_layout;
        FP_CUDA(cudaMemcpyAsync(d, s->h_live + 3, sizeof(uint32_t), cudaMemcpyHostToDevice, f->stream));
        FP_NCCL(s, s->api.AllReduce(d, d, 1, ncclUint32, ncclMax, s->comm, f->stream));
        FP_CUDA(cudaMemcpyAsync(s->h_live + 4, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
    }
    const uint64_t face = s->h_live[4];
    // migration messages: what crosses a face between two binnings is a sliver of a layer
    s->mig_cap = (uint32_t)std::min<uint64_t>(s->n_global, face / 4 + 4096);
    for (int b = 0; b < 2; ++b) {
        cudaFree(s->send_pos[b]);
        cudaFree(s->send_vel[b]);
        if ((rc = dalloc(&s->send_pos[b], (size_t)s->mig_cap + 1)) ||
            (rc = dalloc(&s->send_vel[b], (size_t)s->mig_cap + 1)))
            return rc;
    }
    // room for [ghost L | owned | ghost R] with drift, and for the binning's sort input
    // [ghost L | owned | recv L | recv R]
    const uint64_t need = (uint64_t)n_own + n_own / 4 + 3 * face + 2ull * (s->mig_cap + 1) + 65536;
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
    s->own0 = 0;
    s->own_n = n_own;
    s->n_all = n_own;
    s->next_fits = true;  // `need` above covers the first binning's sort input on every rank
    f->bin_valid = false;

    // grid scratch for the largest array a binning can sort
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
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return ipc_setup(s, f);
}

// A failure decided on one rank in the middle of a collective protocol must become every rank's
// failure before anybody stops talking (NCCL has no time-out: the peers would wait for ever).
// Max-all-reduce of one word, read back; ~40 us.
static int slab_vote(Shard *s, fp_flock *f, uint32_t mine, uint32_t *any) {
    s->h_live[10] = mine;
    uint32_t *d = s->d_layout;
    FP_CUDA(cudaMemcpyAsync(d, s->h_live + 10, sizeof(uint32_t), cudaMemcpyHostToDevice, f->stream));
    FP_NCCL(s, s->api.AllReduce(d, d, 1, ncclUint32, ncclMax, s->comm, f->stream));
    FP_CUDA(cudaMemcpyAsync(s->h_live + 11, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    *any = s->h_live[11];
    return FP_OK;
}
static int slab_capacity_error(fp_flock *f, const char *what) {
    // (every rank takes this exit together; the status bit tells a caller polling fp_flock_status)
    const unsigned bit = FP_STATUS_SLAB_CAPACITY;
    unsigned cur = 0;
    cudaMemcpyAsync(&cur, f->d_status, sizeof(cur), cudaMemcpyDeviceToHost, f->stream);
    cudaStreamSynchronize(f->stream);
    cur |= bit;
    cudaMemcpyAsync(f->d_status, &cur, sizeof(cur), cudaMemcpyHostToDevice, f->stream);
    cudaStreamSynchronize(f->stream);
    set_error(std::string("slab capacity exceeded on some rank (") + what +
              "): the flock is too clustered for this many ranks");
    return FP_ERR_UNSUPPORTED;
}

// ---- a binning of the slab ------------------------------------------------------------------
// in:  owned records in f->pos[cur][own0 .. own0 + own_n) (any order).  Collective; the caller
// has settled.  out: [ghost L | owned | ghost R] cell-sorted in the other buffer, which becomes
// current; cell table, home keys, layouts of all ranks.
static int slab_rebin(Shard *s, fp_flock *f) {
    const GridDesc &L = s->lgrid;
    GridWork &w = f->work;
    const uint32_t n = s->own_n, mc = s->mig_cap;
    const uint32_t m = n + 2 * (mc + 1);
    // (whether this binning's sort input fits was agreed by all ranks at the previous binning, or is
    // guaranteed by the sizing of a fresh fit: nobody leaves the protocol alone)
    if (!s->next_fits) return slab_capacity_error(f, "sort input");
    if ((uint64_t)s->own0 + m > f->cap) {
        set_error("internal: slab sort input exceeds the buffers although every rank agreed it fits");
        return FP_ERR_INVALID;
    }
    const bool has[2] = {s->rank > 0, s->rank < s->world - 1};
    const int nbr[2] = {s->rank - 1, s->rank + 1};
    float4 *pos = f->pos[f->cur] + s->own0, *vel = f->vel[f->cur] + s->own0;
    int rc = launch_skin_gate(f->stream, w.ctl, f->ordinal, 1, f->P.dt, f->skin_budget);
    if (rc) return rc;
    const unsigned nb = nblk(std::max(n, 1u));
    if ((rc = ensure_blk(s, nb))) return rc;
    // 1. boids that left the slab: deterministic two-pass compaction into the two messages
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[0], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[1], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    slab_fate_count_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, pos, vel, n, s->blk_cnt[0], s->blk_cnt[1]);
    scan2_kernel<<<1, 1024, 0, f->stream>>>(s->blk_cnt[0], s->blk_cnt[1], nb + 1);
    slab_fate_emit_kernel<<<nb, SB, 0, f->stream>>>(L, s->xs0, s->xs1, pos, vel, n, s->blk_cnt[0], s->blk_cnt[1], nb,
                                                   s->send_pos[0], s->send_vel[0], s->send_pos[1], s->send_vel[1],
                                                   mc, f->d_status);
    count_launch(3);
    FP_CUDA(cudaGetLastError());
    // receive regions sit right behind the residents: [left: header + mc][right: header + mc]
    float4 *rpos[2] = {pos + n, pos + n + (mc + 1)}, *rvel[2] = {vel + n, vel + n + (mc + 1)};
    FP_CUDA(cudaMemsetAsync(rpos[0], 0, sizeof(float4), f->stream));  // count 0 unless a neighbour says otherwise
    FP_CUDA(cudaMemsetAsync(rpos[1], 0, sizeof(float4), f->stream));
    const size_t words = ((size_t)mc + 1) * 4;
    FP_NCCL(s, s->api.GroupStart());
    for (int fc = 0; fc < 2; ++fc)
        if (has[fc]) {
            FP_NCCL(s, s->api.Send(s->send_pos[fc], words, ncclFloat32, nbr[fc], s->comm, f->stream));
            FP_NCCL(s, s->api.Send(s->send_vel[fc], words, ncclFloat32, nbr[fc], s->comm, f->stream));
            FP_NCCL(s, s->api.Recv(rpos[fc], words, ncclFloat32, nbr[fc], s->comm, f->stream));
            FP_NCCL(s, s->api.Recv(rvel[fc], words, ncclFloat32, nbr[fc], s->comm, f->stream));
        }
    FP_NCCL(s, s->api.GroupEnd());
    // 2. keys + per-cell counts of the owned records (ghost layers still empty)
    const uint32_t CL = (uint32_t)(L.dim[1] * L.dim[2]);
    const uint32_t c_first = CL, c_last = (uint32_t)(L.dim[0] - 2) * CL, c_end = (uint32_t)(L.dim[0] - 1) * CL;
    uint32_t *cnt = w.cell_start;
    FP_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)L.ncells + 2) * sizeof(uint32_t), f->stream));
    slab_keys_kernel<<<nblk(m), SB, 0, f->stream>>>(L, s->xs0, s->xs1, pos, vel, n, mc, w.keys[0], cnt, f->d_status);
    count_launch();
    FP_CUDA(cudaGetLastError());
    // 3. my boundary layers' cell counts are the neighbours' ghost layers' counts
    FP_NCCL(s, s->api.GroupStart());
    if (has[0]) {
        FP_NCCL(s, s->api.Send(cnt + c_first, CL, ncclUint32, nbr[0], s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(cnt, CL, ncclUint32, nbr[0], s->comm, f->stream));
    }
    if (has[1]) {
        FP_NCCL(s, s->api.Send(cnt + c_last, CL, ncclUint32, nbr[1], s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(cnt + c_end, CL, ncclUint32, nbr[1], s->comm, f->stream));
    }
    FP_NCCL(s, s->api.GroupEnd());
    if ((rc = launch_exclusive_scan(f->stream, cnt, (size_t)L.ncells + 2, w.scan_tmp))) return rc;
    const uint32_t probe[5] = {c_first, 2 * CL, c_last, c_end, L.ncells};
    for (int k = 0; k < 5; ++k)
        FP_CUDA(cudaMemcpyAsync(s->h_live + k, cnt + probe[k], sizeof(uint32_t), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaEventRecord(s->ev_live, f->stream));
    // 4. sort (the count read-back overlaps it), gather the owned records behind the left ghosts
    int buf = 0;
    if ((rc = launch_radix_sort(f->stream, w, m, L.key_bits, &buf))) return rc;
    float4 *npos = f->pos[f->cur ^ 1], *nvel = f->vel[f->cur ^ 1];
    float *const *nsoa = w.soa[w.soa_cur ^ 1];
    slab_reorder_kernel<<<nblk(m), SB, 0, f->stream>>>(w.vals[buf], pos, vel, npos, nvel, nsoa[0], nsoa[1], nsoa[2],
                                                      cnt, c_first, c_end, m, f->cap);
    count_launch();
    FP_CUDA(cudaGetLastError());
    FP_CUDA(cudaEventSynchronize(s->ev_live));
    const uint32_t nL = s->h_live[0], f_end = s->h_live[1], l_beg = s->h_live[2], o_end = s->h_live[3],
                   n_all = s->h_live[4];
    {
        uint32_t any = 0;
        const uint32_t mine = (n_all > f->cap || o_end < nL || n_all < o_end) ? 1u : 0u;
        if ((rc = slab_vote(s, f, mine, &any))) return rc;
        if (any) return slab_capacity_error(f, "ghost layers");
    }
    // 5. boundary layers -> the neighbours' ghost blocks, verbatim (same order on both sides)
    const uint32_t sb[2] = {nL, l_beg}, se[2] = {std::min(f_end, o_end), o_end};  // what I send: first / last layer
    const uint32_t rb[2] = {0, o_end}, re[2] = {nL, n_all};                       // where I receive
    FP_NCCL(s, s->api.GroupStart());
    for (int fc = 0; fc < 2; ++fc)
        if (has[fc]) {
            const size_t ns = se[fc] - sb[fc], nr = re[fc] - rb[fc];
            FP_NCCL(s, s->api.Send(npos + sb[fc], ns * 4, ncclFloat32, nbr[fc], s->comm, f->stream));
            FP_NCCL(s, s->api.Send(nvel + sb[fc], ns * 4, ncclFloat32, nbr[fc], s->comm, f->stream));
            FP_NCCL(s, s->api.Recv(npos + rb[fc], nr * 4, ncclFloat32, nbr[fc], s->comm, f->stream));
            FP_NCCL(s, s->api.Recv(nvel + rb[fc], nr * 4, ncclFloat32, nbr[fc], s->comm, f->stream));
            for (int a = 0; a < 3; ++a) {
                FP_NCCL(s, s->api.Send(nsoa[a] + sb[fc], ns, ncclFloat32, nbr[fc], s->comm, f->stream));
                FP_NCCL(s, s->api.Recv(nsoa[a] + rb[fc], nr, ncclFloat32, nbr[fc], s->comm, f->stream));
            }
        }
    FP_NCCL(s, s->api.GroupEnd());
    // 6. every rank learns every layout (where my boundary layers land in the neighbours' arrays)
    f->cur ^= 1;
    w.soa_cur ^= 1;
    // does the NEXT binning's sort input ([ghost L | owned | two receive regions]) fit my buffers?
    const uint32_t fits = (uint64_t)nL + (o_end - nL) + 2ull * (mc + 1) <= f->cap ? 1u : 0u;
    Shard::Layout me{nL, o_end - nL, n_all - o_end, (uint32_t)f->cur, (uint32_t)w.soa_cur, fits, {0, 0}};
    constexpr size_t LW = sizeof(Shard::Layout) / 4;
    s->h_layout[s->rank] = me;
    FP_CUDA(cudaMemcpyAsync(s->d_layout + LW * s->rank, &s->h_layout[s->rank], sizeof(me), cudaMemcpyHostToDevice,
                            f->stream));
    FP_NCCL(s, s->api.AllGather(s->d_layout + LW * s->rank, s->d_layout, LW, ncclUint32, s->comm, f->stream));
    FP_CUDA(cudaMemcpyAsync(s->h_layout, s->d_layout, sizeof(me) * s->world, cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    s->next_fits = true;
    for (int q = 0; q < s->world; ++q) s->next_fits = s->next_fits && s->h_layout[q].next_fits != 0;
    w.home = w.keys[buf];
    s->own0 = nL;
    s->own_n = o_end - nL;
    s->n_all = n_all;
    s->lay_first[0] = sb[0]; s->lay_first[1] = se[0];
    s->lay_last[0] = sb[1];  s->lay_last[1] = se[1];
    f->n = n_all;
    f->permuted = true;
    f->bin_valid = true;
    f->steps_since_bin = 0;
    f->plan_left = flock_plan_steps(f, 0.0f, 0.0f);
    ++f->stat_rebins;
    flock_nl_binned(f);
    s->global_valid = false;
    return FP_OK;
}

// owned slots of the current buffers; TAP_STEP output (and halo push) into the other ones
static WalkIO slab_walk_io(Shard *s, fp_flock *f, bool stepping) {
    GridWork &w = f->work;
    WalkIO io{};
    io.pos_s = f->pos[f->cur];
    io.vel_s = f->vel[f->cur];
    for (int a = 0; a < 3; ++a) io.soa_in[a] = w.soa[w.soa_cur][a];
    io.home = w.home;
    io.cell_start = w.cell_start;
    io.first = s->own0;
    io.last = s->own0 + s->own_n;
    if (!stepping) return io;
    io.pos_out = f->pos[f->cur ^ 1];
    io.vel_out = f->vel[f->cur ^ 1];
    for (int a = 0; a < 3; ++a) io.soa_out[a] = w.soa[w.soa_cur ^ 1][a];
    io.ctl = w.ctl;
    if (s->peer_ok) {
        // The neighbour reads buffer (its cur at the binning) ^ (k & 1) in step k and writes the
        // other one: my boundary boids go where its step k + 1 will read them.
        const uint32_t k = f->steps_since_bin;
        for (int fc = 0; fc < 2; ++fc) {
            const int q = s->rank + (fc ? 1 : -1);
            if (q < 0 || q >= s->world) continue;
            const Shard::Layout &lq = s->h_layout[q];
            const uint32_t pb = (lq.cur ^ (k & 1u)) ^ 1u, sb = (lq.soa_cur ^ (k & 1u)) ^ 1u;
            // my first layer is the left neighbour's RIGHT ghost block, my last layer the right one's LEFT
            const uint32_t dst = fc ? 0u : lq.nL + lq.nO;
            PeerFace &pf = io.push[fc];
            pf.begin = fc ? s->lay_last[0] : s->lay_first[0];
            pf.end = fc ? s->lay_last[1] : s->lay_first[1];
            pf.pos = (float4 *)s->nbr_buf[fc][0 + pb] + dst;
            pf.vel = (float4 *)s->nbr_buf[fc][2 + pb] + dst;
            pf.sx = (float *)s->nbr_buf[fc][4 + 3 * sb + 0] + dst;
            pf.sy = (float *)s->nbr_buf[fc][4 + 3 * sb + 1] + dst;
            pf.sz = (float *)s->nbr_buf[fc][4 + 3 * sb + 2] + dst;
        }
    }
    return io;
}

// after a walk: tell everyone (peer mode), or ship the boundary layers and reduce the speed
// bound with NCCL (fallback)
static int slab_post(Shard *s, fp_flock *f) {
    GridWork &w = f->work;
    if (s->peer_ok) return launch_mail_post(f->stream, w.ctl, s->mail_peers, s->world, s->rank, f->ordinal + 1);
    float4 *npos = f->pos[f->cur ^ 1], *nvel = f->vel[f->cur ^ 1];
    float *const *nsoa = w.soa[w.soa_cur ^ 1];
    const uint32_t sb[2] = {s->lay_first[0], s->lay_last[0]}, se[2] = {s->lay_first[1], s->lay_last[1]};
    const uint32_t rb[2] = {0, s->own0 + s->own_n}, re[2] = {s->own0, s->n_all};
    FP_NCCL(s, s->api.GroupStart());
    for (int fc = 0; fc < 2; ++fc) {
        const int q = s->rank + (fc ? 1 : -1);
        if (q < 0 || q >= s->world) continue;
        const size_t ns = se[fc] - sb[fc], nr = re[fc] - rb[fc];
        FP_NCCL(s, s->api.Send(npos + sb[fc], ns * 4, ncclFloat32, q, s->comm, f->stream));
        FP_NCCL(s, s->api.Send(nvel + sb[fc], ns * 4, ncclFloat32, q, s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(npos + rb[fc], nr * 4, ncclFloat32, q, s->comm, f->stream));
        FP_NCCL(s, s->api.Recv(nvel + rb[fc], nr * 4, ncclFloat32, q, s->comm, f->stream));
        for (int a = 0; a < 3; ++a) {
            FP_NCCL(s, s->api.Send(nsoa[a] + sb[fc], ns, ncclFloat32, q, s->comm, f->stream));
            FP_NCCL(s, s->api.Recv(nsoa[a] + rb[fc], nr, ncclFloat32, q, s->comm, f->stream));
        }
    }
    // max over ranks of (max |v|^2, max |coordinate|): bit patterns of non-negative floats
    FP_NCCL(s, s->api.AllReduce(&w.ctl->v2max, &w.ctl->g_v2max, 1, ncclUint32, ncclMax, s->comm, f->stream));
    FP_NCCL(s, s->api.AllReduce(&w.ctl->pmax, &w.ctl->g_pmax, 1, ncclUint32, ncclMax, s->comm, f->stream));
    FP_NCCL(s, s->api.GroupEnd());
    return FP_OK;
}

static int ensure_slab(Shard *s, fp_flock *f) {
    if (!f->grid_valid || (!f->domain_user && f->steps_since_fit >= 256) || s->rep != REP_SLAB) {
        int rc = shard_settle(s, f);
        if (rc) return rc;
        // bounds come from whatever the rank holds now (ghosts are real boids elsewhere: harmless)
        if ((rc = flock_fit_grid(f))) return rc;  // -> shard_grid_fitted -> repartition
    }
    return FP_OK;
}

static int slab_steps(Shard *s, fp_flock *f, uint32_t nsteps) {
    int rc;
    GridWork &w = f->work;
    for (uint32_t k = 0; k < nsteps; ++k) {
        if (f->pending.size() >= 256 && (rc = shard_settle(s, f))) return rc;
        if ((rc = ensure_slab(s, f))) return rc;
        if ((!f->bin_valid || f->plan_left <= 0) && (rc = shard_settle(s, f))) return rc;
        const uint32_t lead_ver = flock_select_leads(f);
        if ((rc = flock_mark(f))) return rc;
        f->pending.push_back({f->ordinal, f->cur, w.soa_cur, f->table_cursor, f->steps_since_fit, f->steps_since_bin,
                              lead_ver});
        if (!f->bin_valid || f->plan_left <= 0) {
            if ((rc = slab_rebin(s, f))) return rc;
        } else if ((rc = launch_skin_gate(f->stream, w.ctl, f->ordinal, 0, f->P.dt, f->skin_budget,
                                          s->peer_ok ? GATE_MAILBOX : GATE_REDUCED, s->mail, s->world,
                                          f->d_status))) {
            return rc;
        }
        if ((rc = flock_nl_prepare(f, s->lgrid, slab_walk_io(s, f, true)))) return rc;
        if ((rc = flock_mark(f))) return rc;
        if ((rc = flock_step_walk(f, s->lgrid, slab_walk_io(s, f, true)))) return rc;
        if ((rc = slab_post(s, f))) return rc;
        f->cur ^= 1;
        w.soa_cur ^= 1;
        if ((rc = flock_mark(f))) return rc;
        --f->plan_left;
        ++f->ordinal;
        ++f->steps_since_bin;
        ++f->steps_since_fit;
        ++f->table_cursor;
        ++f->stat_grid_steps;
        s->global_valid = false;
    }
    return FP_OK;
}

// Sharded counterpart of settle() in fp_api.cu.  Every rank takes the same decisions: the
// device-side bound is computed from all-reduced maxima, so `stale` and `first_stale` agree.
int shard_settle(Shard *s, fp_flock *f) {
    int rc;
    while (!f->pending.empty()) {
        FP_CUDA(cudaMemcpyAsync(f->h_ctl, f->work.ctl, sizeof(SkinCtl), cudaMemcpyDeviceToHost, f->stream));
        FP_CUDA(cudaStreamSynchronize(f->stream));
        const SkinCtl c = *f->h_ctl;
        if (c.fault) {
            set_error("a peer rank never posted its step (time-out in the step barrier)");
            return FP_ERR_NCCL;
        }
        if (!c.stale) {
            float v2, pm;
            memcpy(&v2, &c.g_v2max, 4);
            memcpy(&pm, &c.g_pmax, 4);
            if (c.g_v2max || c.g_pmax) f->delta_est = flock_plan_delta(v2, pm, f->cfg.dt);
            if (f->bin_valid) f->plan_left = std::min(f->plan_left, flock_plan_steps(f, c.D, f->delta_est));
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
        f->steps_since_bin = at.steps_since_bin;
        f->bin_valid = false;
        f->stat_replayed += redo;
        f->stat_grid_steps -= redo;
        FP_CUDA(cudaMemsetAsync(f->work.ctl, 0, sizeof(SkinCtl), f->stream));
        if (!f->domain_user) f->grid_valid = false;  // re-size the skin for the speeds seen now
        if ((rc = slab_steps(s, f, redo))) return rc;
    }
    return FP_OK;
}

static int allpairs_prepare(Shard *s, fp_flock *f) {
    int rc = to_slice(s, f);
    if (rc) return rc;
    return to_global(s, f);
}

int shard_step(Shard *s, fp_flock *f, uint32_t nsteps) {
    const int m = shard_method(s, f->method, f->cfg);
    int rc;
    if (m == FP_METHOD_GRID) return slab_steps(s, f, nsteps);
    if ((rc = shard_settle(s, f))) return rc;
    if ((rc = allpairs_prepare(s, f))) return rc;
    const size_t off = (size_t)s->rank * s->per;
    for (uint32_t k = 0; k < nsteps; ++k) {
        flock_select_leads(f);
        if ((rc = flock_mark(f)) || (rc = flock_mark(f))) return rc;
        const int a = s->acur;
        rc = flock_allpairs_step(f, s->all_pos[a], s->all_vel[a], (uint32_t)s->n_global, s->first, s->n_slice,
                                 s->all_pos[a ^ 1] + off, s->all_vel[a ^ 1] + off);
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
    if ((rc = shard_settle(s, f))) return rc;
    if (m == FP_METHOD_GRID) {
        // a tap bins the flock where it stands (collective), then walks the owned slots
        if ((rc = ensure_slab(s, f))) return rc;
        if ((rc = slab_rebin(s, f))) return rc;
        if ((rc = flock_tap_walk(f, s->lgrid, tap, slab_walk_io(s, f, false), out))) return rc;
    } else {
        if ((rc = allpairs_prepare(s, f))) return rc;
        if (f->P.numerics_fast &&
            (rc = launch_bounds(f->stream, s->all_pos[s->acur], nullptr, (uint32_t)s->n_global, f->d_bounds)))
            return rc;
        rc = launch_allpairs(f->stream, f->P, tap, s->all_pos[s->acur], s->all_vel[s->acur],
                             (uint32_t)s->n_global, s->first, s->n_slice, nullptr, nullptr, f->d_status, out, 0,
                             f->d_bounds);
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
    *n_local = 0;
    int rc = shard_settle(s, f);
    if (rc) return rc;
    // slab: owned = slots [own0, own0 + own_n); index slice: everything the handle holds
    const uint32_t first = s->rep == REP_SLAB ? s->own0 : 0u;
    const uint32_t n = s->rep == REP_SLAB ? s->own_n : f->n;
    if (!n) return FP_OK;
    const float4 *pos = f->pos[f->cur] + first, *vel = f->vel[f->cur] + first;
    const unsigned nb = nblk(n);
    if ((rc = ensure_blk(s, nb))) return rc;
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[0], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    FP_CUDA(cudaMemsetAsync(s->blk_cnt[1], 0, ((size_t)nb + 1) * sizeof(uint32_t), f->stream));
    owned_count_kernel<<<nb, SB, 0, f->stream>>>(vel, n, s->blk_cnt[0]);
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
    owned_emit_kernel<<<nb, SB, 0, f->stream>>>(pos, vel, n, s->blk_cnt[0], d_idx, d_aos);
    count_launch();
    FP_CUDA(cudaGetLastError());
    FP_CUDA(cudaMemcpyAsync(out_index, d_idx, (size_t)own * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                            f->stream));
    FP_CUDA(cudaMemcpyAsync(out_aos6, d_aos, (size_t)own * 6 * sizeof(float), cudaMemcpyDeviceToHost, f->stream));
    FP_CUDA(cudaStreamSynchronize(f->stream));
    return FP_OK;
}

// rows k of (idx, aos6) -> records first + k; a row whose index differs from the record's flags 64
__global__ void local_write_kernel(float4 *__restrict__ pos, float4 *__restrict__ vel, uint32_t n,
                                   const unsigned long long *__restrict__ idx, const float *__restrict__ aos6,
                                   unsigned *__restrict__ status) {
    const uint32_t i = blockIdx.x * SB + threadIdx.x;
    if (i >= n) return;
    const float4 p = pos[i];
    if ((unsigned long long)__float_as_uint(p.w) != idx[i]) {
        atomicOr(status, 64u);
        return;
    }
    const float *s = aos6 + 6ull * i;
    pos[i] = make_float4(s[0], s[1], s[2], p.w);
    vel[i] = make_float4(s[3], s[4], s[5], 0.0f);
}

int shard_write_local(Shard *s, fp_flock *f, uint64_t n_local, const uint64_t *index, const float *aos6) {
    int rc = shard_settle(s, f);
    if (rc) return rc;
    const uint32_t first = s->rep == REP_SLAB ? s->own0 : 0u;
    const uint32_t n = s->rep == REP_SLAB ? s->own_n : f->n;
    if (n_local != n) {
        set_error("write_local: the row count differs from what fp_flock_read_local lists");
        return FP_ERR_INVALID;
    }
    if (!n) return FP_OK;
    if (!index || !aos6) {
        set_error("null input");
        return FP_ERR_INVALID;
    }
    const size_t bytes = (size_t)n * (6 * sizeof(float) + sizeof(unsigned long long));
    if (bytes > f->stage_bytes) {
        if (f->d_stage) cudaFree(f->d_stage);
        f->d_stage = nullptr;
        f->stage_bytes = 0;
        FP_CUDA(cudaMalloc(&f->d_stage, bytes));
        f->stage_bytes = bytes;
    }
    unsigned long long *d_idx = (unsigned long long *)f->d_stage;
    float *d_aos = (float *)(d_idx + n);
    FP_CUDA(cudaMemcpyAsync(d_idx, index, (size_t)n * sizeof(unsigned long long), cudaMemcpyHostToDevice, f->stream));
    FP_CUDA(cudaMemcpyAsync(d_aos, aos6, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, f->stream));
    local_write_kernel<<<nblk(n), SB, 0, f->stream>>>(f->pos[f->cur] + first, f->vel[f->cur] + first, n, d_idx, d_aos,
                                                     f->d_status);
    count_launch();
    FP_CUDA(cudaGetLastError());
    FP_CUDA(cudaStreamSynchronize(f->stream));  // the caller's buffers are not retained
    f->bin_valid = false;      // positions moved outside the walk's displacement accounting
    s->global_valid = false;
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

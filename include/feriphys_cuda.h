/*
 * feriphys_cuda.h -- C ABI of the B200-native flocking step.
 *
 * This is the drop-in boundary for feriphys's flocking hot path: the entry
 * points a `feriphys-cuda` Rust crate binds with `extern "C"` (see
 * INTEGRATION.md and rust/feriphys-cuda/src/ffi.rs).  Plain pointers and
 * sizes only; every function returns FP_OK (0) or a negative FP_ERR_* code and
 * never unwinds; fp_last_error() gives the message for the calling thread.
 * The caller owns every host buffer; no pointer is retained after a call
 * returns.  One host thread per handle at a time.
 *
 * Each entry cites the reference interface it replaces (paths relative to
 * jalberse/feriphys).  Entries marked ADDITION have no reference counterpart
 * and exist because the reference's state is neither injectable nor readable
 * (SURVEY.md F3, F4).
 */
#ifndef FERIPHYS_CUDA_H
#define FERIPHYS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP_OK 0
#define FP_ERR_INVALID (-1)     /* bad argument */
#define FP_ERR_CUDA (-2)        /* CUDA runtime / driver error, or no device */
#define FP_ERR_UNSUPPORTED (-3) /* configuration outside what the path supports */
#define FP_ERR_NCCL (-4)        /* NCCL unavailable or failed */

/* flocking::Config, src/simulation/flocking/flocking.rs:15-34.
 * time_to_start_steering is a std::time::Duration: whole seconds + nanoseconds. */
typedef struct fp_config {
    float dt;
    float avoidance_factor;
    float centering_factor;
    float velocity_matching_factor;
    float distance_weight_threshold;
    float distance_weight_threshold_falloff;
    float max_sight_angle;
    float max_sight_angle_to_lead_boid;
    uint64_t time_to_start_steering_secs;
    uint32_t time_to_start_steering_nanos;
    int32_t steering_overrides;
} fp_config;

/* How the boid-boid influence pass (flocking.rs:133-151) is evaluated. */
enum {
    FP_METHOD_AUTO = 0,
    FP_METHOD_ALLPAIRS = 1, /* shared-memory tiled O(N^2), reference summation order */
    FP_METHOD_GRID = 2,     /* uniform grid, cell-key radix sort, 27-cell walk (exact) */
    FP_METHOD_SMALL = 3     /* one-CTA kernel for demo-sized flocks, multi-step in one launch */
};

/* Arithmetic of the boid-boid forces (ADDITION; the reference has one arithmetic, separately
 * rounded f32).  Under either setting the neighbour sets -- the distance gate and the sight-angle
 * test of boid.rs:149-157 -- are the reference's, bit for bit.
 *   EXACT: every operation separately rounded in the reference's order; all-pairs and small
 *          flocks are bit-identical to the Rust loop, the grid path to that loop run over the
 *          boids in cell-sorted order (its summation order is the slot order of the binning, so
 *          against the caller's order accelerations agree to ~1e-6 relative, and results depend
 *          on the binning -- skin, re-binning plan, GPU count -- in the last bits).
 *   FAST:  forces, attractor and bounding-box terms use fused multiply-add and MUFU.RSQ / RCP;
 *          the sight-angle decision is taken on a fused cosine outside a 1e-5 guard band and by the
 *          exact sequence inside it.  Accelerations agree with the reference to ~1e-6 relative
 *          (north-star bar: 1e-5); all-pairs splits j across lanes with warp-shuffle reductions. */
enum {
    FP_NUMERICS_EXACT = 0,
    FP_NUMERICS_FAST = 1
};

/* Status bits (fp_flock_status): raised where the reference would panic. */
#define FP_STATUS_STEER_NEGATIVE 1u /* Duration::from_secs_f32(negative), obstacle.rs:25 */
#define FP_STATUS_STEER_NAN_OVF 2u  /* Duration::from_secs_f32(NaN / overflow) */
/* Sharded (multi-GPU) grid runs only; any of these means the step was NOT exact: */
#define FP_STATUS_SLAB_CAPACITY 4u  /* a rank's slab outgrew its buffers */
#define FP_STATUS_HALO_OVERFLOW 8u  /* more boids changed slab between two binnings than a message holds */
#define FP_STATUS_SLAB_JUMP 16u     /* a boid crossed more than one slab between two binnings */
#define FP_STATUS_PEER_TIMEOUT 32u  /* a rank never posted its step in the device-side barrier */

typedef struct fp_flock fp_flock; /* opaque; owns all device memory */

/* Config::default(), flocking.rs:36-51 */
int fp_config_default(fp_config *cfg);

/* Simulation::new (flocking.rs:63-95) with the random jitter made explicit
 * (ADDITION, F3): state_aos6 = n x [px py pz vx vy vz].  device = CUDA
 * ordinal.  Tables start empty (None); config may be NULL for the default. */
int fp_flock_create(fp_flock **out, const fp_config *cfg, uint64_t n, const float *state_aos6,
                    int device);
int fp_flock_destroy(fp_flock *f);

/* sync_sim_config_from_ui (flocking.rs:215-228): legal between any two steps. */
int fp_flock_set_config(fp_flock *f, const fp_config *cfg);
int fp_flock_get_config(fp_flock *f, fp_config *cfg);
int fp_flock_set_method(fp_flock *f, int method);
int fp_flock_get_method(fp_flock *f, int *method_in_use);
/* FP_NUMERICS_*; in_use differs from the request when the configuration has thresholds FAST
 * cannot filter (non-finite or inverted distance thresholds): the exact kernels run then. */
int fp_flock_set_numerics(fp_flock *f, int numerics);
int fp_flock_get_numerics(fp_flock *f, int *numerics, int *in_use);

/* Simulation's Option<...> tables (flocking.rs:56-59); count 0 / NULL = None.
 * leads: n x [px py pz vx vy vz weight] (LeadBoid, boid.rs:13-18);
 * attractors: n x [px py pz mass] (point_attractor.rs:9-12);
 * obstacles:  n x [px py pz radius] (obstacle.rs:11-14);
 * bbox: [x.start x.end y.start y.end z.start z.end] (bounding_box.rs:5-9). */
int fp_flock_set_leads(fp_flock *f, uint32_t n_leads, const float *leads7);
int fp_flock_set_attractors(fp_flock *f, uint32_t n, const float *attractors4);
int fp_flock_set_obstacles(fp_flock *f, uint32_t n, const float *obstacles4);
int fp_flock_set_bbox(fp_flock *f, const float *bbox6);
/* Lead boids follow Rust fn pointers (boid.rs:35, parametric.rs:5) that cannot
 * cross an FFI; the host evaluates them (in the reference's order, F9) and
 * uploads steps x n_leads x 7 floats.  Step k of the following fp_flock_step
 * calls uses row k; when the table is exhausted the last row stays in force. */
int fp_flock_set_lead_table(fp_flock *f, uint32_t steps, uint32_t n_leads, const float *table7);

/* Simulation::step (flocking.rs:97-131), nsteps times, asynchronously on the
 * handle's stream.  The Rust shim returns Duration::from_secs_f32(dt). */
int fp_flock_step(fp_flock *f, uint32_t nsteps);
int fp_flock_sync(fp_flock *f);

/* ADDITIONS (F4): state in caller index order, n x 6 floats.  Neither call retains the caller's
 * buffer.  Flocks up to 256 KB: write_state copies the rows into a pinned, device-mapped slot and
 * returns without touching the device; on the single-CTA kernel (FP_METHOD_SMALL) the next step reads
 * them from there and leaves the advanced rows in mapped memory for read_state -- a write / step /
 * read round is one launch and one stream synchronisation. */
int fp_flock_read_state(fp_flock *f, float *out_aos6);
int fp_flock_write_state(fp_flock *f, const float *state_aos6);
uint64_t fp_flock_len(const fp_flock *f);
/* OR of the per-boid panic flags raised since the last call; clears them. */
int fp_flock_status(fp_flock *f, uint32_t *flags);

/* get_boid_instances (flocking.rs:230-245): n x [px py pz  qs qx qy qz  scale]
 * (graphics/instance.rs:7-11; rotation = Quaternion::from_arc(unit_z, v^)). */
int fp_flock_read_instances(fp_flock *f, float *out8);
/* Instance::to_raw (instance.rs:14-22, :39-44): n x 25 floats, column-major
 * 4x4 model then 3x3 normal matrix -- the 100-byte InstanceRaw record. */
int fp_flock_read_instances_raw(fp_flock *f, float *out25);
/* The same records written by the GPU straight into memory the caller maps: a device allocation
 * (a mapped graphics-interop vertex buffer, cudaMalloc), or pinned host memory that the device can
 * address (cudaHostAlloc / cudaHostRegister: the kernel's stores cross PCIe, no staging copy).
 * Replaces the reference's per-instance queue.write_buffer loop (graphics/instance.rs:131-141).
 * raw = 0: n x 8 floats (Instance), raw != 0: n x 25 floats (InstanceRaw).  Returns when the
 * records are in place.  A pointer the device cannot address is FP_ERR_INVALID. */
int fp_flock_export_instances(fp_flock *f, void *dst_device_visible, int raw);

/* Debug taps for the parity tests (ADDITIONS).  All describe the CURRENT
 * state without advancing it.
 * accel: n x 3 total acceleration; comp15 (may be NULL): n x 5 x 3 --
 *   boids, leads, attractors, bbox, steering (flocking.rs:105-113).
 * neighbors: per boid, |N(i)| and the order-independent hash sum(mix64(j))
 *   of N(i) = { j : boid_j != boid_i, not FOV-culled, dist <= thr or
 *   dist < thr + falloff } -- the bit-exact predicate check.
 * census: [rejected by distance, FOV-culled in range, contributing] ordered
 *   pairs over all i, plus candidates examined (== N(N-1) for all-pairs). */
int fp_flock_read_accel(fp_flock *f, float *out_accel3, float *out_comp15);
int fp_flock_read_neighbors(fp_flock *f, uint32_t *out_count, uint64_t *out_hash);
int fp_flock_pair_census(fp_flock *f, uint64_t out4[4]);

/* Uniform-grid controls (ADDITION; the reference is all-pairs only, F6).
 * The domain is fitted to the state at create/write time; boids outside it are
 * clamped into the edge cells (still exact).  fp_flock_grid_info reports
 * dims[3], cell size and key bits of the grid in use. */
int fp_flock_set_grid_domain(fp_flock *f, const float lo3[3], const float hi3[3]);
int fp_flock_grid_info(fp_flock *f, uint32_t dims3[3], float *cell_size, uint32_t *key_bits);
/* Lazy re-binning of the grid path (ADDITION; the reference has no spatial structure).  The
 * cell edge carries a skin on top of the interaction reach, so one binning (sort by cell)
 * serves every step until some boid could have moved skin / 2 from where it was binned --
 * the device bounds that with max|v| * dt per step and voids a step that would exceed it; the
 * host re-bins and replays voided steps, so results never depend on the plan.  Neighbour
 * sets stay exact: the same f32 predicates decide over a superset of candidates.
 * Reports the skin in use and counters since creation: grid steps performed, binnings,
 * steps replayed after the device voided them. */
/* Policy knobs: skin < 0 (default) sizes the skin from the flock's speed at every grid fit,
 * 0 bins on every step, > 0 fixes it.  plan_scale (default 1) stretches the number of steps
 * the host plans per binning; values > 1 make the device-side check and the replay do the
 * work (used by the tests). */
int fp_flock_set_rebin(fp_flock *f, float skin, float plan_scale);
int fp_flock_rebin_info(fp_flock *f, float *skin, uint64_t *grid_steps, uint64_t *rebins,
                        uint64_t *replayed);

/* Device-resident access for callers that already hold device memory
 * (ADDITION).  pos4/vel4 are the SoA float4 arrays of the current state in
 * INTERNAL order; pos4[i].w carries the caller index as uint32 bits. */
int fp_flock_device_state(fp_flock *f, const void **pos4, const void **vel4);
/* Timing hook (ADDITION): between _begin and _end every step records CUDA
 * events on the handle's stream around its sort phase (keys, scan, radix sort,
 * gather) and its influence kernel.  _end synchronises and returns the number
 * of steps seen, the device time from the first event to the last (span_ms)
 * and the summed durations of the two phases. */
int fp_flock_timing_begin(fp_flock *f);
int fp_flock_timing_end(fp_flock *f, uint32_t *steps, float *span_ms, float *sort_ms,
                        float *influence_ms);

/* State<T>::euler_step / rk4_step (src/simulation/state.rs:75-106) over the
 * flock viewed as a Stateful with 6 elements [px py pz vx vy vz] whose
 * derivative is [v, a] with the acceleration accumulated beforehand and held
 * frozen across stages (the springy pattern, springy_mesh.rs:199-257). */
int fp_flock_state_euler(fp_flock *f, float h);
int fp_flock_state_rk4(fp_flock *f, float h);

/* Generic State<T> integrators on flat host vectors of n floats with a
 * caller-supplied derivative already evaluated on the device side is not
 * expressible over a C ABI; these two take the k-vectors explicitly:
 * out = s + ds*h  and  out = s + (((h/6*k1 + h/3*k2) + h/3*k3) + h/6*k4). */
int fp_state_euler_combine(int device, size_t n, const float *s, const float *ds, float h,
                           float *out);
int fp_state_rk4_combine(int device, size_t n, const float *s, const float *k1, const float *k2,
                         const float *k3, const float *k4, float h, float *out);

/* Device-resident State<T> (src/simulation/state.rs:37-113) for the reference's Stateful types
 * (state.rs:10-16).  The flat state vector -- State::as_vector, n x num_state_elements floats in
 * the element order of each type's as_state -- lives on the device between steps; a step is one
 * streaming pass in which every element evaluates all stages of the integrator in registers
 * (Stateful::derivative sees only its own element, state.rs:14), 2 x 4 bytes of HBM traffic per
 * float and step.  Same arithmetic, operation for operation, as State::euler_step / rk4_step. */
enum {
    FP_STATEFUL_TEST_POINT = 1,     /* state.rs:120-152: [p3 v3], derivative [v, (1, -1, 0)]          */
    FP_STATEFUL_TEST_EXAMPLEFN = 2, /* state.rs:187-216: [y t timestep], y' = y - t^2 + 1              */
    FP_STATEFUL_SPRINGY_POINT = 3,  /* springy_mesh.rs:199-257: [mass p3 v3 accumulated_force3]        */
    FP_STATEFUL_RIGIDBODY = 4,      /* rigidbody.rs:53-190: [p3 q(v3 s) P3 L3 mass Iinv0(9) F3 T3]     */
    FP_STATEFUL_BOID = 5            /* [p3 v3 a3]: a FlockingBoid with its acceleration frozen         */
};
typedef struct fp_state fp_state;
int fp_state_num_state_elements(int kind);                 /* Stateful::num_state_elements (0: unknown kind) */
/* State::new / State::from_state_vector: state = n_elements x k floats (NULL: zeros) */
int fp_state_create(fp_state **out, int device, int kind, uint64_t n_elements, const float *state);
int fp_state_destroy(fp_state *s);
uint64_t fp_state_len(const fp_state *s);
int fp_state_write(fp_state *s, const float *state);       /* from_state_vector */
int fp_state_read(fp_state *s, float *out);                /* as_vector / get_elements */
int fp_state_derivative(fp_state *s, float *out);          /* State::derivative */
/* nsteps x State::euler_step(h) / State::rk4_step(h); asynchronous on the handle's stream */
int fp_state_euler_step(fp_state *s, float h, uint32_t nsteps);
int fp_state_rk4_step(fp_state *s, float h, uint32_t nsteps);
int fp_state_sync(fp_state *s);
int fp_state_device_vector(fp_state *s, const float **dev); /* valid until the next step */
/* bench hook: `launches` single-step passes timed with CUDA events on the handle's stream */
int fp_state_time_steps(fp_state *s, float h, int rk4, uint32_t launches, float *ms_total);

/* The neighbour pass of sph::Simulation::step (src/simulation/sph/mod.rs:89-121) on the flocking
 * path's grid infrastructure (cell keys, radix sort, cell table) instead of a kd-tree rebuilt every
 * step: for every particle its k nearest (itself included, ascending squared_euclidean distance;
 * equal distances by particle id -- declared, kiddo leaves it unspecified) closer than
 * kernal_max_distance, and the density sum(particle_mass * monaghan(r, s)) over them in that order
 * (kernals.rs:6-16).  pos3: n x 3 (host); out_index: n x k particle ids, 0xffffffff past out_count[i].
 * 1 <= k <= 32.  kernel_ms (may be NULL): device time of binning + query. */
int fp_sph_neighbors(int device, uint64_t n, const float *pos3, uint32_t k, float kernal_max_distance,
                     float particle_mass, uint32_t *out_index, uint32_t *out_count, float *out_density,
                     float *kernel_ms);

/* Multi-GPU (one process per GPU).  The flock is sharded by boid index
 * (all-pairs: NCCL all-gather of pos/vel each step) or by x-slab (grid: halo
 * exchange + migration).  nccl_unique_id is the 128-byte ncclUniqueId made by
 * fp_nccl_unique_id on rank 0 and broadcast by the caller (torch.distributed,
 * MPI, ...).  state_aos6 holds this rank's n_local boids; global indices are
 * first_index .. first_index + n_local. */
int fp_nccl_unique_id(uint8_t out128[128]);
int fp_flock_create_sharded(fp_flock **out, const fp_config *cfg, uint64_t n_global,
                            uint64_t first_index, uint64_t n_local, const float *state_aos6,
                            int device, int rank, int world, const uint8_t nccl_unique_id[128]);
/* Rank, world size, and whether the grid path's halo exchange runs over peer-mapped memory
 * (cudaIpc + NVLink stores fused into the walk kernel, mailbox step barrier) or, when the
 * mapping is unavailable, over ncclSend/ncclRecv.  Unsharded handles report 0, 1, 0. */
int fp_flock_shard_info(fp_flock *f, int *rank, int *world, int *peer_mapped);
/* Rows this rank currently owns (grid slabs migrate boids between ranks). */
int fp_flock_local_len(fp_flock *f, uint64_t *n_local);
/* Local state with global indices: out_index n_local x u64, out_aos6 n_local x 6. */
int fp_flock_read_local(fp_flock *f, uint64_t *out_index, float *out_aos6);
/* New values for the boids this rank holds, in the order fp_flock_read_local lists them
 * (index[k] must be the k-th index it returned): the sharded counterpart of
 * fp_flock_write_state.  Every rank calls it (SPMD); the next step re-bins, and a boid whose new
 * position lies in a neighbouring slab migrates there. */
int fp_flock_write_local(fp_flock *f, uint64_t n_local, const uint64_t *index, const float *state_aos6);

/* Self-test (ADDITION): compares the branch-free sqrt / division sequences of the grid walk
 * with __fsqrt_rn / __fdiv_rn bit for bit on n pseudo-random operand sets drawn from the
 * exponent ranges the kernel admits; out_mismatch = {sqrt mismatches, div mismatches}. */
int fp_debug_fastmath_check(int device, uint64_t n, uint64_t seed, uint64_t out_mismatch[2]);

const char *fp_last_error(void);
/* "feriphys-cuda <version> sm_100a" */
const char *fp_version(void);
/* Number of kernel launches issued by this library in this process (bench). */
uint64_t fp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif

// fp_state.cu -- device-resident State<T> (src/simulation/state.rs) for the reference's Stateful types.
//
// The reference integrates a `Vec<T: Stateful>` by flattening it to a `Vec<f32>` (as_vector),
// evaluating `derivative()` element by element and rebuilding the elements from the new vector
// (state.rs:42-106) -- five temporaries per Euler step, seventeen per RK4 step.  A Stateful's
// derivative sees only its own element (state.rs:14), so on the GPU a whole step is ONE streaming
// pass: a thread holds its element's k floats in registers, evaluates every stage there, and
// writes the element back.  No k-vector ever touches memory: 2 * k * 4 bytes of HBM traffic per
// element and step, whatever the integrator.  The flat vector stays on the device between steps
// (fp_state handle); the element types are those of the reference:
//   springy Point     (10 floats, springy_mesh.rs:199-257): forces accumulated beforehand and frozen
//   rigid-body State  (29 floats, rigidbody.rs:53-190)
//   boid              ( 9 floats: position, velocity, frozen acceleration -- the same pattern
//                      applied to flocking::FlockingBoid)
//   the two Stateful types of the reference's own State tests (state.rs:120-216): their golden
//   values are what pins this file.
// Arithmetic: every operation separately rounded in the source's order (round-to-nearest
// intrinsics), as everywhere else on the exact path.
#include <new>

#include "fp_internal.h"

struct fp_state {
    int device = 0;
    int kind = 0;
    uint32_t k = 0;        // floats per element
    uint64_t n = 0;        // elements
    float *vec[2] = {nullptr, nullptr};  // the flat state vector, double-buffered
    int cur = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
};

namespace fp {

namespace {

constexpr int ST_BLOCK = 256;  // elements per CTA

// ---- Stateful::derivative, one function per element type ----------------------------------
template <int KIND>
struct Elem;

// state.rs:120-152 (test type): [p, v], derivative [v, (1, -1, 0)]
template <>
struct Elem<FP_STATEFUL_TEST_POINT> {
    static constexpr int K = 6;
    __device__ static void derivative(const float *s, float *d) {
        d[0] = s[3]; d[1] = s[4]; d[2] = s[5];
        d[3] = 1.0f; d[4] = -1.0f; d[5] = 0.0f;
    }
};
// state.rs:187-216 (test type): [y, t, timestep], y' = y - t^2 + 1, t' = 1
template <>
struct Elem<FP_STATEFUL_TEST_EXAMPLEFN> {
    static constexpr int K = 3;
    __device__ static void derivative(const float *s, float *d) {
        d[0] = fadd(fsub(s[0], fmul(s[1], s[1])), 1.0f);  // self.y - f32::powi(self.t, 2) + 1.0
        d[1] = 1.0f;
        d[2] = 0.0f;
    }
};
// springy_mesh.rs:199-257: [mass, p, v, accumulated_force]
template <>
struct Elem<FP_STATEFUL_SPRINGY_POINT> {
    static constexpr int K = 10;
    __device__ static void derivative(const float *s, float *d) {
        d[0] = 0.0f;                                   // mass does not change
        d[1] = s[4]; d[2] = s[5]; d[3] = s[6];         // position' = velocity
        d[4] = fdiv(s[7], s[0]);                       // velocity' = accumulated_force / mass
        d[5] = fdiv(s[8], s[0]);
        d[6] = fdiv(s[9], s[0]);
        d[7] = d[8] = d[9] = 0.0f;                     // the accumulated force is frozen
    }
};
// position, velocity, frozen acceleration: the Point pattern for a FlockingBoid (mass 1)
template <>
struct Elem<FP_STATEFUL_BOID> {
    static constexpr int K = 9;
    __device__ static void derivative(const float *s, float *d) {
        d[0] = s[3]; d[1] = s[4]; d[2] = s[5];
        d[3] = s[6]; d[4] = s[7]; d[5] = s[8];
        d[6] = d[7] = d[8] = 0.0f;
    }
};
// rigidbody.rs:53-140.  cgmath 0.18 [ext]: Matrix3::from(Quaternion), Matrix3 * Matrix3 (rows of the
// left times columns of the right, dot = (x x' + y y') + z z'), Matrix3 * Vector3 (columns scaled
// and added left to right), f32 * Quaternion, Quaternion * Quaternion.
template <>
struct Elem<FP_STATEFUL_RIGIDBODY> {
    static constexpr int K = 29;
    __device__ static float dot3(float ax, float ay, float az, float bx, float by, float bz) {
        return fadd(fadd(fmul(ax, bx), fmul(ay, by)), fmul(az, bz));
    }
    // column-major 3x3: m[3 * col + row]
    __device__ static void matmul(const float *a, const float *b, float *o) {
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r)
                o[3 * c + r] = dot3(a[r], a[3 + r], a[6 + r], b[3 * c], b[3 * c + 1], b[3 * c + 2]);
    }
    __device__ static void derivative(const float *s, float *d) {
        const float qx = s[3], qy = s[4], qz = s[5], qs = s[6];  // rotation: v then s
        const float mass = s[13];
        const float *I0 = s + 14;                                // initial inverted inertia, columns
        // velocity() = linear_momentum / mass
        d[0] = fdiv(s[7], mass); d[1] = fdiv(s[8], mass); d[2] = fdiv(s[9], mass);
        // Matrix3::from(rotation)
        const float x2 = fadd(qx, qx), y2 = fadd(qy, qy), z2 = fadd(qz, qz);
        const float xx2 = fmul(x2, qx), xy2 = fmul(x2, qy), xz2 = fmul(x2, qz);
        const float yy2 = fmul(y2, qy), yz2 = fmul(y2, qz), zz2 = fmul(z2, qz);
        const float sy2 = fmul(y2, qs), sz2 = fmul(z2, qs), sx2 = fmul(x2, qs);
        float R[9], Rt[9], RI[9], Iinv[9];
        R[0] = fsub(fsub(1.0f, yy2), zz2); R[1] = fadd(xy2, sz2); R[2] = fsub(xz2, sy2);
        R[3] = fsub(xy2, sz2); R[4] = fsub(fsub(1.0f, xx2), zz2); R[5] = fadd(yz2, sx2);
        R[6] = fadd(xz2, sy2); R[7] = fsub(yz2, sx2); R[8] = fsub(fsub(1.0f, xx2), yy2);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) Rt[3 * c + r] = R[3 * r + c];
        matmul(R, I0, RI);        // rotation_matrix * initial_moment_of_intertia_inverted
        matmul(RI, Rt, Iinv);     //   * rotation_matrix.transpose()
        // angular_velocity() = Iinv * angular_momentum
        const float Lx = s[10], Ly = s[11], Lz = s[12];
        float w[3];
        for (int r = 0; r < 3; ++r)
            w[r] = fadd(fadd(fmul(Iinv[r], Lx), fmul(Iinv[3 + r], Ly)), fmul(Iinv[6 + r], Lz));
        // 0.5 * Quaternion::from_sv(0.0, w) * rotation
        const float as = fmul(0.5f, 0.0f), ax = fmul(0.5f, w[0]), ay = fmul(0.5f, w[1]), az = fmul(0.5f, w[2]);
        const float rs = fsub(fsub(fsub(fmul(as, qs), fmul(ax, qx)), fmul(ay, qy)), fmul(az, qz));
        const float rx = fsub(fadd(fadd(fmul(as, qx), fmul(ax, qs)), fmul(ay, qz)), fmul(az, qy));
        const float ry = fsub(fadd(fadd(fmul(as, qy), fmul(ay, qs)), fmul(az, qx)), fmul(ax, qz));
        const float rz = fsub(fadd(fadd(fmul(as, qz), fmul(az, qs)), fmul(ax, qy)), fmul(ay, qx));
        d[3] = rx; d[4] = ry; d[5] = rz; d[6] = rs;
        d[7] = s[23]; d[8] = s[24]; d[9] = s[25];      // linear momentum' = accumulated force
        d[10] = s[26]; d[11] = s[27]; d[12] = s[28];   // angular momentum' = accumulated torque
        for (int i = 13; i < 29; ++i) d[i] = 0.0f;     // constants carried in the state
    }
};

// State::euler_step (state.rs:75-83) / State::rk4_step (state.rs:86-106) of one element, in registers
template <int KIND, bool RK4>
__device__ __forceinline__ void step_element(float *s, float h) {
    constexpr int K = Elem<KIND>::K;
    float k1[K];
    Elem<KIND>::derivative(s, k1);
    if (!RK4) {
#pragma unroll
        for (int i = 0; i < K; ++i) s[i] = fadd(s[i], fmul(k1[i], h));  // x * timestep, then vec_add
        return;
    }
    float k2[K], k3[K], k4[K], t[K];
    const float hh = fmul(h, 0.5f);
#pragma unroll
    for (int i = 0; i < K; ++i) t[i] = fadd(s[i], fmul(k1[i], hh));
    Elem<KIND>::derivative(t, k2);
#pragma unroll
    for (int i = 0; i < K; ++i) t[i] = fadd(s[i], fmul(k2[i], hh));
    Elem<KIND>::derivative(t, k3);
#pragma unroll
    for (int i = 0; i < K; ++i) t[i] = fadd(s[i], fmul(k3[i], h));
    Elem<KIND>::derivative(t, k4);
    const float h6 = fdiv(h, 6.0f), h3 = fdiv(h, 3.0f);
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float delta = fadd(fadd(fadd(fmul(h6, k1[i]), fmul(h3, k2[i])), fmul(h3, k3[i])), fmul(h6, k4[i]));
        s[i] = fadd(s[i], delta);
    }
}

// MODE 0: Euler step, 1: RK4 step, 2: derivative only (State::derivative).  The CTA's slice of the
// flat vector is staged through shared memory so that global traffic is fully coalesced whatever k.
template <int KIND, int MODE>
__global__ void __launch_bounds__(ST_BLOCK)
state_kernel(const float *__restrict__ in, float *__restrict__ out, uint64_t n, float h, uint32_t nsteps) {
    constexpr int K = Elem<KIND>::K;
    __shared__ float tile[ST_BLOCK * K];
    const uint64_t e0 = (uint64_t)blockIdx.x * ST_BLOCK;
    const uint32_t cnt = (uint32_t)min((uint64_t)ST_BLOCK, n - e0);
    const float *src = in + e0 * K;
    for (uint32_t i = threadIdx.x; i < cnt * K; i += ST_BLOCK) tile[i] = src[i];
    __syncthreads();
    if (threadIdx.x < cnt) {
        float s[K];
#pragma unroll
        for (int i = 0; i < K; ++i) s[i] = tile[threadIdx.x * K + i];
        if (MODE == 2) {
            float d[K];
            Elem<KIND>::derivative(s, d);
#pragma unroll
            for (int i = 0; i < K; ++i) tile[threadIdx.x * K + i] = d[i];
        } else {
            // (several steps of one element are independent of every other element: they can run
            //  back to back in registers -- nsteps > 1 is the whole trajectory in one pass)
            for (uint32_t st = 0; st < nsteps; ++st) step_element<KIND, MODE == 1>(s, h);
#pragma unroll
            for (int i = 0; i < K; ++i) tile[threadIdx.x * K + i] = s[i];
        }
    }
    __syncthreads();
    float *dst = out + e0 * K;
    for (uint32_t i = threadIdx.x; i < cnt * K; i += ST_BLOCK) dst[i] = tile[i];
}

template <int KIND>
int launch_kind(fp_state *s, int mode, const float *in, float *out, float h, uint32_t nsteps) {
    const unsigned grid = (unsigned)((s->n + ST_BLOCK - 1) / ST_BLOCK);
    if (!grid) return FP_OK;
    if (mode == 0) state_kernel<KIND, 0><<<grid, ST_BLOCK, 0, s->stream>>>(in, out, s->n, h, nsteps);
    else if (mode == 1) state_kernel<KIND, 1><<<grid, ST_BLOCK, 0, s->stream>>>(in, out, s->n, h, nsteps);
    else state_kernel<KIND, 2><<<grid, ST_BLOCK, 0, s->stream>>>(in, out, s->n, h, nsteps);
    count_launch();
    FP_CUDA(cudaGetLastError());
    return FP_OK;
}

int launch_state(fp_state *s, int mode, const float *in, float *out, float h, uint32_t nsteps) {
    switch (s->kind) {
        case FP_STATEFUL_TEST_POINT: return launch_kind<FP_STATEFUL_TEST_POINT>(s, mode, in, out, h, nsteps);
        case FP_STATEFUL_TEST_EXAMPLEFN: return launch_kind<FP_STATEFUL_TEST_EXAMPLEFN>(s, mode, in, out, h, nsteps);
        case FP_STATEFUL_SPRINGY_POINT: return launch_kind<FP_STATEFUL_SPRINGY_POINT>(s, mode, in, out, h, nsteps);
        case FP_STATEFUL_BOID: return launch_kind<FP_STATEFUL_BOID>(s, mode, in, out, h, nsteps);
        case FP_STATEFUL_RIGIDBODY: return launch_kind<FP_STATEFUL_RIGIDBODY>(s, mode, in, out, h, nsteps);
    }
    set_error("unknown Stateful kind");
    return FP_ERR_INVALID;
}

int kind_elements(int kind) {
    switch (kind) {
        case FP_STATEFUL_TEST_POINT: return 6;
        case FP_STATEFUL_TEST_EXAMPLEFN: return 3;
        case FP_STATEFUL_SPRINGY_POINT: return 10;
        case FP_STATEFUL_BOID: return 9;
        case FP_STATEFUL_RIGIDBODY: return 29;
    }
    return 0;
}

int check_state(fp_state *s) {
    if (!s) {
        set_error("null state handle");
        return FP_ERR_INVALID;
    }
    FP_CUDA(cudaSetDevice(s->device));
    return FP_OK;
}

}  // namespace
}  // namespace fp

using namespace fp;

extern "C" {

int fp_state_num_state_elements(int kind) { return kind_elements(kind); }

int fp_state_create(fp_state **out, int device, int kind, uint64_t n_elements, const float *state) {
    if (!out) { set_error("null out pointer"); return FP_ERR_INVALID; }
    *out = nullptr;
    const int k = kind_elements(kind);
    if (!k) { set_error("unknown Stateful kind"); return FP_ERR_INVALID; }
    if (n_elements >= (1ull << 40) / (uint64_t)k) { set_error("state too large"); return FP_ERR_INVALID; }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        set_error("no usable CUDA device -- this library has no CPU fallback");
        return FP_ERR_CUDA;
    }
    FP_CUDA(cudaSetDevice(device));
    fp_state *s = new (std::nothrow) fp_state();
    if (!s) { set_error("out of host memory"); return FP_ERR_INVALID; }
    s->device = device;
    s->kind = kind;
    s->k = (uint32_t)k;
    s->n = n_elements;
    const size_t bytes = std::max<size_t>(1, (size_t)n_elements * k) * sizeof(float);
    cudaError_t err = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    for (int b = 0; b < 2 && err == cudaSuccess; ++b) err = cudaMalloc((void **)&s->vec[b], bytes);
    for (int b = 0; b < 2 && err == cudaSuccess; ++b) err = cudaEventCreate(&s->ev[b]);
    if (err == cudaSuccess && state && n_elements)
        err = cudaMemcpyAsync(s->vec[0], state, (size_t)n_elements * k * sizeof(float), cudaMemcpyHostToDevice, s->stream);
    else if (err == cudaSuccess)
        err = cudaMemsetAsync(s->vec[0], 0, bytes, s->stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s->stream);
    if (err != cudaSuccess) {
        fp_state_destroy(s);
        return cuda_fail(err, "fp_state_create", __FILE__, __LINE__);
    }
    *out = s;
    return FP_OK;
}

int fp_state_destroy(fp_state *s) {
    if (!s) return FP_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int b = 0; b < 2; ++b) {
        if (s->vec[b]) cudaFree(s->vec[b]);
        if (s->ev[b]) cudaEventDestroy(s->ev[b]);
    }
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return FP_OK;
}

uint64_t fp_state_len(const fp_state *s) { return s ? s->n : 0; }

int fp_state_write(fp_state *s, const float *state) {
    int rc = check_state(s);
    if (rc) return rc;
    if (!s->n) return FP_OK;
    if (!state) { set_error("null state"); return FP_ERR_INVALID; }
    FP_CUDA(cudaMemcpyAsync(s->vec[s->cur], state, (size_t)s->n * s->k * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    FP_CUDA(cudaStreamSynchronize(s->stream));
    return FP_OK;
}

int fp_state_read(fp_state *s, float *out) {
    int rc = check_state(s);
    if (rc) return rc;
    if (!s->n) return FP_OK;
    if (!out) { set_error("null output"); return FP_ERR_INVALID; }
    FP_CUDA(cudaMemcpyAsync(out, s->vec[s->cur], (size_t)s->n * s->k * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    FP_CUDA(cudaStreamSynchronize(s->stream));
    return FP_OK;
}

int fp_state_derivative(fp_state *s, float *out) {
    int rc = check_state(s);
    if (rc) return rc;
    if (!s->n) return FP_OK;
    if (!out) { set_error("null output"); return FP_ERR_INVALID; }
    if ((rc = launch_state(s, 2, s->vec[s->cur], s->vec[s->cur ^ 1], 0.0f, 0))) return rc;
    FP_CUDA(cudaMemcpyAsync(out, s->vec[s->cur ^ 1], (size_t)s->n * s->k * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    FP_CUDA(cudaStreamSynchronize(s->stream));
    return FP_OK;
}

static int state_step(fp_state *s, float h, uint32_t nsteps, int rk4) {
    int rc = check_state(s);
    if (rc) return rc;
    if (!s->n || !nsteps) return FP_OK;
    if ((rc = launch_state(s, rk4 ? 1 : 0, s->vec[s->cur], s->vec[s->cur ^ 1], h, nsteps))) return rc;
    s->cur ^= 1;
    return FP_OK;
}
int fp_state_euler_step(fp_state *s, float h, uint32_t nsteps) { return state_step(s, h, nsteps, 0); }
int fp_state_rk4_step(fp_state *s, float h, uint32_t nsteps) { return state_step(s, h, nsteps, 1); }

int fp_state_sync(fp_state *s) {
    int rc = check_state(s);
    if (rc) return rc;
    FP_CUDA(cudaStreamSynchronize(s->stream));
    return FP_OK;
}

int fp_state_device_vector(fp_state *s, const float **dev) {
    int rc = check_state(s);
    if (rc) return rc;
    FP_CUDA(cudaStreamSynchronize(s->stream));
    if (dev) *dev = s->vec[s->cur];
    return FP_OK;
}

// `launches` passes of the chosen integrator, timed with CUDA events on the handle's stream
int fp_state_time_steps(fp_state *s, float h, int rk4, uint32_t launches, float *ms_total) {
    int rc = check_state(s);
    if (rc) return rc;
    FP_CUDA(cudaEventRecord(s->ev[0], s->stream));
    for (uint32_t i = 0; i < launches; ++i)
        if ((rc = state_step(s, h, 1, rk4))) return rc;
    FP_CUDA(cudaEventRecord(s->ev[1], s->stream));
    FP_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0.0f;
    FP_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    if (ms_total) *ms_total = ms;
    return FP_OK;
}

}  // extern "C"

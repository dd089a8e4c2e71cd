// fp_sph.cu -- SPH neighbour pass (src/simulation/sph/mod.rs:88-125) on the uniform-grid
// infrastructure of the flocking path (cell keys, radix sort, cell table: fp_grid.cu, fp_sort.cu).
//
// The reference builds a kd-tree (kiddo) of all particles every step, asks it for each particle's
// k = 8 nearest by squared_euclidean (the particle itself included), keeps those closer than
// kernal_max_distance, and sums particle_mass * monaghan(r, s) over them for the density
// (mod.rs:89-121, kernals.rs:6-16).  Everything kept lies within kernal_max_distance, so a grid of
// cells no smaller than that finds the same set in the 27 cells around a particle -- exactly, with
// the same f32 distances:
//     d2 = ((0 + dx dx) + dy dy) + dz dz          kiddo::distance::squared_euclidean [ext]
// Order: ascending d2, as kiddo returns them; equal distances -- which the reference's own initial
// lattice is full of -- are ordered by particle id here (kiddo's order among ties is an accident of
// its tree layout: DECLARED, parity unpinned; the reference holds no test for this path).
// The density is summed in that order.
#include <math.h>

#include <algorithm>
#include <vector>

#include "fp_grid.cuh"

namespace fp {

namespace {

constexpr int SPH_BLOCK = 128;

__global__ void sph_load_kernel(const float *__restrict__ pos3, float4 *__restrict__ pos, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    pos[i] = make_float4(pos3[3ull * i], pos3[3ull * i + 1], pos3[3ull * i + 2], __uint_as_float(i));
}
__global__ void sph_gather_kernel(const uint32_t *__restrict__ vals, const float4 *__restrict__ in,
                                  float4 *__restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[vals[i]];
}

// kernals.rs:6-16 -- powi(2) = x x, powi(3) = (x x) x [ext: LLVM expands small constant powers]
__device__ __forceinline__ float monaghan(float r, float s) {
    const float q = fdiv(r, s);
    float num;
    if (q >= 0.0f && q <= 1.0f)
        num = fadd(fsub(1.0f, fmul(1.5f, fmul(q, q))), fmul(0.75f, fmul(fmul(q, q), q)));
    else if (q >= 1.0f && q <= 2.0f) {
        const float t = fsub(2.0f, q);
        num = fmul(0.25f, fmul(fmul(t, t), t));
    } else
        num = 0.0f;
    return fdiv(num, fmul(3.14159274101257324f, fmul(fmul(s, s), s)));
}

template <int KMAX>
__global__ void __launch_bounds__(SPH_BLOCK)
sph_knn_kernel(const GridDesc g, const float4 *__restrict__ pos_s, const uint32_t *__restrict__ cell_start,
               uint32_t n, uint32_t k, float s, float mass, uint32_t *__restrict__ out_index,
               uint32_t *__restrict__ out_count, float *__restrict__ out_density) {
    const uint32_t slot = blockIdx.x * SPH_BLOCK + threadIdx.x;
    if (slot >= n) return;
    const float4 p = pos_s[slot];
    const uint32_t id = __float_as_uint(p.w);
    const float s2 = fmul(s, s);  // kernal_max_distance.powi(2)
    float bd[KMAX];
    uint32_t bi[KMAX];
    uint32_t cnt = 0;
    const int cx = cell_coord_x(g, p.x), cy = cell_coord(p.y, g.origin[1], g.inv_cell, g.dim[1]),
              cz = cell_coord(p.z, g.origin[2], g.inv_cell_z, g.dim[2]);
    for (int x = max(cx - 1, 0); x <= min(cx + 1, g.dim[0] - 1); ++x)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dim[1] - 1); ++y) {
            const uint32_t rowbase = row_base(g, x, y);
            const uint32_t b = __ldg(cell_start + rowbase + max(cz - 1, 0));
            const uint32_t e = __ldg(cell_start + rowbase + min(cz + 1, g.dim[2] - 1) + 1);
            for (uint32_t j = b; j < e; ++j) {
                const float4 o = __ldg(pos_s + j);
                const float dx = fsub(p.x, o.x), dy = fsub(p.y, o.y), dz = fsub(p.z, o.z);
                const float d2 = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
                if (!(d2 < s2)) continue;  // .filter(|neighbor| neighbor.0 < s^2)  (mod.rs:101)
                const uint32_t oid = __float_as_uint(o.w);
                // insert into the k best by (d2, id)
                if (cnt == k && !(d2 < bd[k - 1] || (d2 == bd[k - 1] && oid < bi[k - 1]))) continue;
                uint32_t pos = cnt < k ? cnt : k - 1;
#pragma unroll
                for (int t = KMAX - 1; t > 0; --t) {
                    if ((uint32_t)t <= pos && (d2 < bd[t - 1] || (d2 == bd[t - 1] && oid < bi[t - 1]))) {
                        bd[t] = bd[t - 1];
                        bi[t] = bi[t - 1];
                        pos = t - 1;
                    }
                }
                bd[pos] = d2;
                bi[pos] = oid;
                if (cnt < k) ++cnt;
            }
        }
    // density: particle_mass * monaghan(r, s) summed in neighbour order (mod.rs:106-118)
    float density = 0.0f;
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
        if ((uint32_t)t < cnt) {
            const float r = bd[t] == 0.0f ? 0.0f : fsqrt(bd[t]);  // r_ij.is_zero() <=> d2 == 0 (no underflow at these scales)
            density = fadd(density, fmul(mass, monaghan(r, s)));
            out_index[(size_t)id * k + t] = bi[t];
        } else if ((uint32_t)t < k) {
            out_index[(size_t)id * k + t] = 0xffffffffu;
        }
    }
    out_count[id] = cnt;
    out_density[id] = density;
}

struct Scratch {
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <class T>
    int alloc(T **p, size_t count) {
        *p = nullptr;
        FP_CUDA(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
        ptrs.push_back(*p);
        return FP_OK;
    }
};

}  // namespace
}  // namespace fp

using namespace fp;

extern "C" int fp_sph_neighbors(int device, uint64_t n64, const float *pos3, uint32_t k, float kernal_max_distance,
                                float particle_mass, uint32_t *out_index, uint32_t *out_count, float *out_density,
                                float *kernel_ms) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_error("no usable CUDA device -- this library has no CPU fallback");
        return FP_ERR_CUDA;
    }
    if (k == 0 || k > 32 || n64 >= (1ull << 31) || !(kernal_max_distance > 0.0f) || !std::isfinite(kernal_max_distance)) {
        set_error("sph neighbours: need 1 <= k <= 32, n < 2^31 and a finite positive kernal_max_distance");
        return FP_ERR_INVALID;
    }
    if (!n64) return FP_OK;
    if (!pos3 || !out_index || !out_count || !out_density) { set_error("null argument"); return FP_ERR_INVALID; }
    FP_CUDA(cudaSetDevice(device));
    const uint32_t n = (uint32_t)n64;
    Scratch sc;
    cudaStream_t st = nullptr;  // (a one-shot call: the default stream)
    float *d_pos3 = nullptr, *d_bounds = nullptr, *d_density = nullptr;
    float4 *d_pos = nullptr, *d_sorted = nullptr;
    uint32_t *d_index = nullptr, *d_count = nullptr;
    int rc;
    if ((rc = sc.alloc(&d_pos3, (size_t)n * 3)) || (rc = sc.alloc(&d_pos, n)) || (rc = sc.alloc(&d_sorted, n)) ||
        (rc = sc.alloc(&d_bounds, 8)) || (rc = sc.alloc(&d_index, (size_t)n * k)) || (rc = sc.alloc(&d_count, n)) ||
        (rc = sc.alloc(&d_density, n)))
        return rc;
    FP_CUDA(cudaMemcpyAsync(d_pos3, pos3, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    sph_load_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_pos3, d_pos, n);
    count_launch();
    if ((rc = launch_bounds(st, d_pos, nullptr, n, d_bounds))) return rc;
    float b[8];
    FP_CUDA(cudaMemcpyAsync(b, d_bounds, sizeof(b), cudaMemcpyDeviceToHost, st));
    FP_CUDA(cudaStreamSynchronize(st));
    // grid: cell edge a hair over the kernel support (covers the f32 rounding of the cell coordinate)
    GridDesc g{};
    double cell = (double)kernal_max_distance * (1.0 + 1.0 / 512.0);
    for (;;) {
        uint64_t cells = 1;
        bool ok = true;
        for (int a = 0; a < 3; ++a) {
            const double lo = b[a] <= b[3 + a] ? b[a] : 0.0, hi = b[a] <= b[3 + a] ? b[3 + a] : 0.0;
            const double d = floor((hi - lo) / cell) + 1.0;
            if (!(d <= 4096.0)) { ok = false; break; }
            g.dim[a] = (int)d;
            g.origin[a] = (float)lo;
            cells *= (uint64_t)g.dim[a];
        }
        if (ok && cells <= (1ull << 24)) { g.ncells = (uint32_t)cells; break; }
        cell *= 1.25;
    }
    g.cell = (float)cell;
    g.inv_cell = g.inv_cell_z = 1.0f / g.cell;
    g.zspan = 1;
    g.gdimx = g.dim[0];
    g.xoff = 0;
    uint32_t bits = 1;
    while ((1ull << bits) < g.ncells) ++bits;
    g.key_bits = bits;
    GridWork w{};
    const size_t ntiles = ((size_t)n + 4095) / 4096 + 1, hist = 256 * ntiles;
    const size_t scan_n = std::max(hist, (size_t)g.ncells + 1);
    if ((rc = sc.alloc(&w.keys[0], n)) || (rc = sc.alloc(&w.keys[1], n)) || (rc = sc.alloc(&w.vals[0], n)) ||
        (rc = sc.alloc(&w.vals[1], n)) || (rc = sc.alloc(&w.cell_start, (size_t)g.ncells + 1)) ||
        (rc = sc.alloc(&w.tile_hist, hist)) || (rc = sc.alloc(&w.scan_tmp, scan_n / 4096 + 2)))
        return rc;
    w.cap = n;
    w.tile_hist_elems = hist;
    w.cell_cap = (size_t)g.ncells + 1;
    w.scan_tmp_elems = scan_n / 4096 + 2;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    FP_CUDA(cudaEventCreate(&e0));
    FP_CUDA(cudaEventCreate(&e1));
    FP_CUDA(cudaEventRecord(e0, st));
    int buf = 0;
    rc = launch_grid_keys(st, g, d_pos, n, w);
    if (!rc) rc = launch_radix_sort(st, w, n, g.key_bits, &buf);
    if (!rc) {
        sph_gather_kernel<<<(n + 255) / 256, 256, 0, st>>>(w.vals[buf], d_pos, d_sorted, n);
        const unsigned grid = (n + SPH_BLOCK - 1) / SPH_BLOCK;
        if (k <= 8)
            sph_knn_kernel<8><<<grid, SPH_BLOCK, 0, st>>>(g, d_sorted, w.cell_start, n, k, kernal_max_distance,
                                                         particle_mass, d_index, d_count, d_density);
        else if (k <= 16)
            sph_knn_kernel<16><<<grid, SPH_BLOCK, 0, st>>>(g, d_sorted, w.cell_start, n, k, kernal_max_distance,
                                                          particle_mass, d_index, d_count, d_density);
        else
            sph_knn_kernel<32><<<grid, SPH_BLOCK, 0, st>>>(g, d_sorted, w.cell_start, n, k, kernal_max_distance,
                                                          particle_mass, d_index, d_count, d_density);
        count_launch(2);
    }
    cudaError_t err = cudaEventRecord(e1, st);
    if (!rc && err == cudaSuccess) err = cudaGetLastError();
    if (!rc && err == cudaSuccess)
        err = cudaMemcpyAsync(out_index, d_index, (size_t)n * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (!rc && err == cudaSuccess) err = cudaMemcpyAsync(out_count, d_count, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (!rc && err == cudaSuccess) err = cudaMemcpyAsync(out_density, d_density, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    float ms = 0.0f;
    if (err == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    if (rc) return rc;
    if (err != cudaSuccess) return cuda_fail(err, "fp_sph_neighbors", __FILE__, __LINE__);
    return FP_OK;
}

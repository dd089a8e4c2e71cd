// fp32_rates.cu -- instruction-rate microbenchmark for the pipes the flocking kernels live on
// (B200, sm_100a): scalar FFMA / FADD / FMUL, packed FFMA2 / FADD2, FMNMX, FSETP+SEL, MUFU.RSQ,
// and broadcast LDS.128.  Prints results per SM per clock.  Build and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32_rates tools/micro/fp32_rates.cu && /tmp/fp32_rates
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int ITER = 4096, CHAINS = 8;

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(float *out, float seed) {
    float a[CHAINS], b = seed, c = seed * 0.5f;
    float2 p[CHAINS];
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed, seed, seed);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
        a[k] = seed + k;
        p[k] = make_float2(seed + k, seed - k);
    }
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
    for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (OP == 0) a[k] = fmaf(a[k], b, c);
            if (OP == 1) a[k] = __fadd_rn(a[k], b);
            if (OP == 2) a[k] = __fmul_rn(a[k], b);
            if (OP == 3) p[k] = __ffma2_rn(p[k], b2, c2);
            if (OP == 4) p[k] = __fadd2_rn(p[k], b2);
            if (OP == 5) a[k] = fminf(a[k], b + i);
            if (OP == 6) a[k] = (a[k] > b) ? c : a[k] + 1.0f;     // FSETP + FSEL + FADD
            if (OP == 7) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
            if (OP == 8) {
                const float4 v = sm[(i + k) & 63];                 // broadcast LDS.128
                a[k] += v.x;
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += a[k] + p[k].x + p[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
double run(const char *name, double ops_per_iter_thread) {
    float *out;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    rate_kernel<OP><<<blocks, threads>>>(out, 1.0001f);
    cudaEventRecord(e0);
    rate_kernel<OP><<<blocks, threads>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double total = (double)blocks * threads * ITER * CHAINS * ops_per_iter_thread;
    const double per_sm_clk = total / (ms * 1e-3) / sms / (khz * 1e3);
    printf("%-28s %8.3f ms  %7.1f lane-ops / SM / clk (at %d MHz nominal)\n", name, ms, per_sm_clk, khz / 1000);
    cudaFree(out);
    return per_sm_clk;
}

int main() {
    run<0>("FFMA (scalar)", 1);
    run<1>("FADD (scalar)", 1);
    run<2>("FMUL (scalar)", 1);
    run<3>("FFMA2 (2 FMAs per lane-op)", 2);
    run<4>("FADD2 (2 adds per lane-op)", 2);
    run<5>("FMNMX (+FADD)", 1);
    run<6>("FSETP + FSEL + FADD", 1);
    run<7>("MUFU.RSQ", 1);
    run<8>("LDS.128 broadcast (+FADD)", 1);
    return 0;
}

// Pipe-throughput microbenchmarks for the softmax inner loop of the head-dim-40 attention kernel (sm_100a).
// Answers, on the real part: how many clocks per warp instruction per SM sub-partition do MUFU.EX2, F2FP (bf16x2 pack),
// FFMA, FMNMX and the polynomial exp2 cost, alone and in the mixes the kernel can choose between.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_pipes tools/ubench_pipes.cu && tools/ubench_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack(float a, float b) {
    uint32_t r;
    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// round-to-nearest (ties up) bf16x2 pack on the integer pipes: 2 IADD + 1 PRMT
__device__ __forceinline__ uint32_t pack_int(float a, float b) {
    const uint32_t ua = __float_as_uint(a) + 0x8000u, ub = __float_as_uint(b) + 0x8000u;
    return __byte_perm(ua, ub, 0x7632);
}
__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -126.0f);
    const float r = x + 12582912.0f;
    const float f = x - (r - 12582912.0f);
    float p = fmaf(0.054592825f, f, 0.24221784f);
    p = fmaf(p, f, 0.6933686f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

constexpr int kIters = 512;

template <int MODE>
__global__ void bench(float* out, long long* clk, float seed) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = seed + 0.001f * (threadIdx.x + k);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0) {  // MUFU.EX2 only
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ex2a(v[k]);
        } else if (MODE == 1) {  // F2FP pack only (4 per 8 values)
#pragma unroll
            for (int k = 0; k < 8; k += 2) acc ^= pack(v[k] + it, v[k + 1]);
        } else if (MODE == 2) {  // FFMA only
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaf(v[k], 1.0001f, seed);
        } else if (MODE == 3) {  // FMNMX only
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = fmaxf(v[k], seed + k + it);
        } else if (MODE == 4) {  // polynomial exp2 only
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = ex2_poly(v[k] - 1.0f);
        } else if (MODE == 5) {  // integer pack only
#pragma unroll
            for (int k = 0; k < 8; k += 2) acc ^= pack_int(v[k] + it, v[k + 1]);
        } else if (MODE >= 10 && MODE < 20) {  // softmax body: 8 x (ffma, exp) + 4 packs; MODE-10 of 8 exps on the polynomial
            float e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float a = fmaf(v[k], 0.5f, -seed);
                e[k] = (k < MODE - 10) ? ex2_poly(a) : ex2a(a);
            }
#pragma unroll
            for (int k = 0; k < 8; k += 2) acc ^= pack(e[k], e[k + 1]);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += 1e-3f;
        } else if (MODE >= 20 && MODE < 30) {  // same with the integer pack
            float e[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float a = fmaf(v[k], 0.5f, -seed);
                e[k] = (k < MODE - 20) ? ex2_poly(a) : ex2a(a);
            }
#pragma unroll
            for (int k = 0; k < 8; k += 2) acc ^= pack_int(e[k], e[k + 1]);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += 1e-3f;
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc & 0xff);
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int per_iter_units) {
    float* out;
    long long* clk;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&clk, 148 * sizeof(long long));
    for (int threads : {128, 256, 512, 1024}) {
        bench<MODE><<<148, threads>>>(out, clk, 0.25f);
        cudaDeviceSynchronize();
        bench<MODE><<<148, threads>>>(out, clk, 0.25f);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        const double warps_per_smsp = threads / 32 / 4.0;
        // clocks per (unit x warp) per SMSP
        printf("%-34s threads/SM %4d  clk/iter %8.1f  clk per unit-warp-instr per SMSP %6.2f\n", name, threads,
               avg / kIters, avg / kIters / (per_iter_units * warps_per_smsp));
    }
    cudaFree(out);
    cudaFree(clk);
}

int main() {
    run<0>("MUFU.EX2 (8/iter)", 8);
    run<1>("F2FP bf16x2 pack (4/iter, +4 FADD)", 4);
    run<2>("FFMA (8/iter)", 8);
    run<3>("FMNMX (8/iter)", 8);
    run<4>("poly exp2 (8/iter)", 8);
    run<5>("int pack (4/iter, +4 FADD)", 4);
    run<10>("softmax body 8 elem, poly 0, F2FP", 8);
    run<11>("softmax body 8 elem, poly 1, F2FP", 8);
    run<12>("softmax body 8 elem, poly 2, F2FP", 8);
    run<13>("softmax body 8 elem, poly 3, F2FP", 8);
    run<14>("softmax body 8 elem, poly 4, F2FP", 8);
    run<20>("softmax body 8 elem, poly 0, intpack", 8);
    run<22>("softmax body 8 elem, poly 2, intpack", 8);
    run<23>("softmax body 8 elem, poly 3, intpack", 8);
    run<24>("softmax body 8 elem, poly 4, intpack", 8);
    return 0;
}

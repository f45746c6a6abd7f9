// GroupNorm (+SiLU) and LayerNorm over token-major bf16 activations. HBM-bound: every element is read with
// 128-bit loads, statistics are fp32, and the grid is sized to cover all 148 SMs even at batch 1.
#include "common.cuh"
#include "ptx.cuh"

namespace imagd {

constexpr int kGnMaxC = 2560;
constexpr int kGnThreads = 512;
constexpr int kGnCounters = 1024;  // max samples per call

__host__ __device__ inline int gn_chunks(int HW) {
    int c = (HW + 31) / 32;
    return c < 1 ? 1 : (c > 64 ? 64 : c);
}

// Workspace layout: [kGnCounters uint arrival counters (zero-initialised once by the caller; re-armed by the kernel)]
// | [NB*chunks*groups*2] partial {sum, sum of squares} | [NB*groups*2] final {mean, rstd}.
// The last block of a sample to finish folds the partials in a fixed order
// (bit-reproducible) so the apply kernel reads 2 numbers per group instead of re-reducing `chunks` partials per CTA.
__global__ void __launch_bounds__(kGnThreads) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int HW,
                                                              int C, int groups, int chunks, float* __restrict__ ws,
                                                              int NB, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    // per (pixel-row slot, channel) partials, reduced in a fixed order below -> bit-reproducible statistics
    __shared__ float s_sum[kGnThreads * 8];
    __shared__ float s_sq[kGnThreads * 8];
    const int n = blockIdx.y, chunk = blockIdx.x;
    const int CV = C / 8;
    const int rows = kGnThreads / CV;  // pixel rows processed per iteration (>= 1 since C <= 2560 < 8*512)
    const int ppc = (HW + chunks - 1) / chunks;
    const int p_begin = chunk * ppc;
    const int p_end = min(HW, p_begin + ppc);

    const int cv = threadIdx.x % CV;
    const int prow = threadIdx.x / CV;
    if (prow < rows) {
        float sum[8], sq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sum[k] = sq[k] = 0.f;
        const __nv_bfloat16* base = x + (static_cast<int64_t>(n) * HW) * ldx + cv * 8;
        for (int pix = p_begin + prow; pix < p_end; pix += rows * 4) {
            uint4 v[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                v[t] = make_uint4(0u, 0u, 0u, 0u);  // zeros add nothing to either sum
                if (pix + t * rows < p_end)
                    v[t] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<int64_t>(pix + t * rows) * ldx));
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t u[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float a = bf16lo(u[k]), b = bf16hi(u[k]);
                    sum[2 * k] += a;
                    sq[2 * k] += a * a;
                    sum[2 * k + 1] += b;
                    sq[2 * k + 1] += b * b;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s_sum[prow * C + cv * 8 + k] = sum[k];
            s_sq[prow * C + cv * 8 + k] = sq[k];
        }
    }
    __syncthreads();
    // fold the row slots into slot 0 (fixed order)
    for (int c = threadIdx.x; c < C; c += kGnThreads) {
        float a = s_sum[c], b = s_sq[c];
        for (int r = 1; r < rows; ++r) {
            a += s_sum[r * C + c];
            b += s_sq[r * C + c];
        }
        s_sum[c] = a;
        s_sq[c] = b;
    }
    __syncthreads();
    const int cpg = C / groups;
    for (int g = threadIdx.x; g < groups; g += kGnThreads) {
        float a = 0.f, b = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            a += s_sum[c];
            b += s_sq[c];
        }
        float* dst = ws + ((static_cast<int64_t>(n) * chunks + chunk) * groups + g) * 2;
        __stcg(dst, a);
        __stcg(dst + 1, b);
    }
    // ---- last block of this sample finalizes
    float* fin = ws + static_cast<int64_t>(NB) * chunks * groups * 2;
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws) - kGnCounters;
    __shared__ unsigned int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(&counters[n], 1u);
        s_last = (old == static_cast<unsigned int>(chunks - 1)) ? 1u : 0u;
        if (s_last) counters[n] = 0u;  // re-arm
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int g = threadIdx.x; g < groups; g += kGnThreads) {
        float a = 0.f, b = 0.f;
        for (int c = 0; c < chunks; ++c) {
            const float* src = ws + ((static_cast<int64_t>(n) * chunks + c) * groups + g) * 2;
            a += __ldcg(src);
            b += __ldcg(src + 1);
        }
        const float cnt = static_cast<float>(cpg) * static_cast<float>(HW);
        const float mean = a / cnt;
        const float var = fmaxf(b / cnt - mean * mean, 0.f);
        fin[(static_cast<int64_t>(n) * groups + g) * 2] = mean;
        fin[(static_cast<int64_t>(n) * groups + g) * 2 + 1] = rsqrtf(var + eps);
    }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                       __nv_bfloat16* __restrict__ y, int64_t ldy, int HW, int C,
                                                       int groups, int chunks, const float* __restrict__ ws,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int fuse_silu, int apply_chunks, int NB) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_scale[kGnMaxC];
    __shared__ float s_shift[kGnMaxC];
    const int n = blockIdx.y;
    const int cpg = C / groups;
    const float* fin = ws + static_cast<int64_t>(NB) * chunks * groups * 2 + static_cast<int64_t>(n) * groups * 2;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = (gamma ? gamma[c] : 1.f) * fin[2 * g + 1];
        s_scale[c] = sc;
        s_shift[c] = (beta ? beta[c] : 0.f) - fin[2 * g] * sc;
    }
    __syncthreads();

    const int CV = C / 8;
    const int ppc = (HW + apply_chunks - 1) / apply_chunks;
    const int p_begin = blockIdx.x * ppc;
    const int p_end = min(HW, p_begin + ppc);
    const int total = (p_end - p_begin) * CV;
    constexpr int U = 4;  // independent 128-bit loads in flight per thread
    for (int base = threadIdx.x; base < total; base += blockDim.x * U) {
        uint4 v[U];
        int c0s[U];
        int64_t rowsv[U];
#pragma unroll
        for (int t = 0; t < U; ++t) {
            const int idx = base + t * blockDim.x;
            v[t] = make_uint4(0u, 0u, 0u, 0u);
            c0s[t] = 0;
            rowsv[t] = 0;
            if (idx < total) {
                c0s[t] = (idx % CV) * 8;
                rowsv[t] = static_cast<int64_t>(n) * HW + p_begin + idx / CV;
                v[t] = __ldg(reinterpret_cast<const uint4*>(x + rowsv[t] * ldx + c0s[t]));
            }
        }
#pragma unroll
        for (int t = 0; t < U; ++t) {
            if (base + t * static_cast<int>(blockDim.x) >= total) break;
            const uint32_t u[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
            const int c0 = c0s[t];
            float f[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                f[2 * k] = bf16lo(u[k]) * s_scale[c0 + 2 * k] + s_shift[c0 + 2 * k];
                f[2 * k + 1] = bf16hi(u[k]) * s_scale[c0 + 2 * k + 1] + s_shift[c0 + 2 * k + 1];
            }
            if (fuse_silu) {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = silu(f[k]);
            }
            *reinterpret_cast<uint4*>(y + rowsv[t] * ldy + c0) = make_uint4(
                pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
    }
}

// LayerNorm: one warp normalises R rows at a time (R x VPL independent 128-bit loads in flight per lane — the rows are
// only 640 B..2.5 KB, so memory-level parallelism, not arithmetic, sets the speed). C <= 2048, C % 8 == 0.
template <int VPL, int R>
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                        __nv_bfloat16* __restrict__ y, int64_t ldy, int rows, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int row0 = (blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5)) * R;
    const int lane = threadIdx.x & 31;
    if (row0 >= rows) return;
    const int CV = C / 8;
    uint4 raw[R][VPL];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int cv = lane + i * 32;
            raw[r][i] = make_uint4(0u, 0u, 0u, 0u);
            if (cv < CV && row0 + r < rows)
                raw[r][i] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row0 + r) * ldx + cv * 8));
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (row0 + r >= rows) break;
        float f[VPL][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const uint32_t u[4] = {raw[r][i].x, raw[r][i].y, raw[r][i].z, raw[r][i].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                f[i][2 * k] = bf16lo(u[k]);
                f[i][2 * k + 1] = bf16hi(u[k]);
                sum += f[i][2 * k] + f[i][2 * k + 1];  // lanes beyond CV hold zeros
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / static_cast<float>(C);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            if (lane + i * 32 < CV) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float d = f[i][k] - mean;
                    sq += d * d;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq / static_cast<float>(C) + eps);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int cv = lane + i * 32;
            if (cv < CV) {
                float o[8];
                float g[8], b[8];
                if (gamma) {
                    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8));
                    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8 + 4));
                    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[k] = 1.f;
                }
                if (beta) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8 + 4));
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) b[k] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = (f[i][k] - mean) * rstd * g[k] + b[k];
                *reinterpret_cast<uint4*>(y + static_cast<int64_t>(row0 + r) * ldy + cv * 8) =
                    make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            }
        }
    }
}

template <int VPL, int R>
static int launch_ln(const void* x, int64_t ldx, void* y, int64_t ldy, int rows, int C, const float* gamma,
                     const float* beta, float eps, cudaStream_t st) {
    const int rows_per_block = 8 * R;
    IMAGD_CUDA(launch_pdl(layernorm_kernel<VPL, R>, dim3((rows + rows_per_block - 1) / rows_per_block), dim3(256), 0, st,
                          reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, rows, C,
                          gamma, beta, eps));
    return IMAGD_OK;
}

}  // namespace imagd

extern "C" {

int64_t imagd_groupnorm_ws_bytes(int NB, int HW, int C, int groups) {
    (void)C;
    return (imagd::kGnCounters + static_cast<int64_t>(NB) * imagd::gn_chunks(HW) * groups * 2 +
            static_cast<int64_t>(NB) * groups * 2) * sizeof(float);
}

int imagd_groupnorm_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int NB, int HW, int C, int groups,
                         const float* gamma, const float* beta, float eps, int fuse_silu, void* ws, imagd_stream stream) {
    using namespace imagd;
    IMAGD_CHECK_ARG(x && y && ws, "groupnorm: null pointer");
    IMAGD_CHECK_ARG(NB > 0 && NB <= kGnCounters && HW > 0 && C > 0 && C % 8 == 0 && C <= kGnMaxC,
                    "groupnorm: NB=%d C=%d unsupported", NB, C);
    IMAGD_CHECK_ARG(groups > 0 && groups <= 64 && C % groups == 0, "groupnorm: groups=%d", groups);
    IMAGD_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y), "groupnorm: alignment");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int chunks = gn_chunks(HW);
    IMAGD_CUDA(launch_pdl(gn_stats_kernel, dim3(dim3(chunks, NB)), dim3(kGnThreads), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), ldx, HW, C,
                                                             groups, chunks, reinterpret_cast<float*>(ws) + kGnCounters, NB, eps));
    int apply_chunks = (HW + 15) / 16;
    if (apply_chunks > 128) apply_chunks = 128;
    IMAGD_CUDA(launch_pdl(gn_apply_kernel, dim3(dim3(apply_chunks, NB)), dim3(256), 0, st, 
        reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, HW, C, groups, chunks,
        reinterpret_cast<const float*>(ws) + kGnCounters, gamma, beta, fuse_silu, apply_chunks, NB));
    return IMAGD_OK;
}

int imagd_layernorm_bf16(const void* x, int64_t ldx, void* y, int64_t ldy, int rows, int C, const float* gamma,
                         const float* beta, float eps, imagd_stream stream) {
    using namespace imagd;
    IMAGD_CHECK_ARG(x && y, "layernorm: null pointer");
    IMAGD_CHECK_ARG(rows > 0 && C > 0 && C % 8 == 0 && C <= 2048, "layernorm: C=%d unsupported", C);
    IMAGD_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y), "layernorm: alignment");
    IMAGD_CHECK_ARG((!gamma || aligned16(gamma)) && (!beta || aligned16(beta)), "layernorm: gamma/beta alignment");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int vpl = (C / 8 + 31) / 32;
    if (vpl <= 2) return launch_ln<2, 4>(x, ldx, y, ldy, rows, C, gamma, beta, eps, st);
    if (vpl <= 3) return launch_ln<3, 4>(x, ldx, y, ldy, rows, C, gamma, beta, eps, st);
    if (vpl <= 5) return launch_ln<5, 2>(x, ldx, y, ldy, rows, C, gamma, beta, eps, st);
    return launch_ln<8, 1>(x, ldx, y, ldy, rows, C, gamma, beta, eps, st);

}

}  // extern "C"

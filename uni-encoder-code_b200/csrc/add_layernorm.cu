// add_layernorm.cu -- y = LayerNorm(x + residual) * gamma + beta over the last dimension, fp32.
//
// the two post-norm steps of every encoder layer, `norm1(src + dropout1(attn))` and
// `norm2(src + dropout3(ffn))` (msdeformattn.py:126-142).  The reference (torch eager) runs an add
// kernel and a LayerNorm kernel: 5 passes over [rows, 256] where 3 suffice.  HBM-bound: one warp per
// row, the row lives in registers between the statistics and the normalisation.
//   mean = sum(v) / C;  var = sum((v - mean)^2) / C  (two-pass, in registers);  rstd = rsqrt(var + eps)
//   y = (v - mean) * rstd * gamma + beta             (torch.nn.functional.layer_norm semantics)
#include <cuda_runtime.h>
#include <cstdint>

#include "msda_common.cuh"

namespace msda {

namespace {

// V = float4 chunks per lane: C = 128 * V columns
template <int V>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const float4 *__restrict__ x, const float4 *__restrict__ res,
                     const float4 *__restrict__ gamma, const float4 *__restrict__ beta,
                     float4 *__restrict__ y, long long rows, float eps) {
    constexpr int C4 = 32 * V;                      // float4 per row
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    float4 g[V], b[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        g[i] = gamma[i * 32 + lane];
        b[i] = beta[i * 32 + lane];
    }
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        float4 v[V];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float4 a = ldg_stream_f4(x + row * C4 + i * 32 + lane);
            v[i] = a;
            if (res != nullptr) {
                const float4 r = ldg_stream_f4(res + row * C4 + i * 32 + lane);
                v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
            }
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFullMask, sum, o);
        const float mean = sum * (1.f / (4 * C4));
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
            sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFullMask, sq, o);
        const float rstd = rsqrtf(sq * (1.f / (4 * C4)) + eps);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float4 o;
            o.x = (v[i].x - mean) * rstd * g[i].x + b[i].x;
            o.y = (v[i].y - mean) * rstd * g[i].y + b[i].y;
            o.z = (v[i].z - mean) * rstd * g[i].z + b[i].z;
            o.w = (v[i].w - mean) * rstd * g[i].w + b[i].w;
            y[row * C4 + i * 32 + lane] = o;
        }
    }
}

template <int V>
cudaError_t launch(const float *x, const float *res, const float *gamma, const float *beta, float *y,
                   long long rows, float eps, cudaStream_t stream) {
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    add_layernorm_kernel<V><<<(unsigned)blocks, 256, 0, stream>>>(
        reinterpret_cast<const float4 *>(x), reinterpret_cast<const float4 *>(res),
        reinterpret_cast<const float4 *>(gamma), reinterpret_cast<const float4 *>(beta),
        reinterpret_cast<float4 *>(y), rows, eps);
    note_launch();
    return cudaGetLastError();
}

}  // namespace

// cols in {128, 256, 384, 512}; 16-byte aligned pointers; residual may be null (plain LayerNorm)
cudaError_t launch_add_layernorm(const float *x, const float *res, const float *gamma, const float *beta, float *y,
                                 long long rows, int cols, float eps, cudaStream_t stream, bool *handled) {
    *handled = true;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(res) |
                           reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                           reinterpret_cast<uintptr_t>(y);
    if (bits % 16 != 0) {
        *handled = false;
        return cudaSuccess;
    }
    switch (cols) {
        case 128: return launch<1>(x, res, gamma, beta, y, rows, eps, stream);
        case 256: return launch<2>(x, res, gamma, beta, y, rows, eps, stream);
        case 384: return launch<3>(x, res, gamma, beta, y, rows, eps, stream);
        case 512: return launch<4>(x, res, gamma, beta, y, rows, eps, stream);
        default: *handled = false; return cudaSuccess;
    }
}

}  // namespace msda

// ---------------------------------------------------------------------------
// Backward of y = LayerNorm(x + residual) * gamma + beta (training): with v = x + residual,
// xhat = (v - mean) * rstd and g = grad_y * gamma,
//     grad_v = rstd * (g - mean_c(g) - xhat * mean_c(g * xhat))         (= grad_x = grad_residual)
//     grad_gamma = sum_rows grad_y * xhat,   grad_beta = sum_rows grad_y
// v, mean and rstd are recomputed from x and residual (one extra read of residual instead of a stored copy
// of v).  One warp per row as in the forward; each warp keeps its partial grad_gamma / grad_beta in
// registers over all its rows and adds them to the zero-filled outputs once at the end.
// ---------------------------------------------------------------------------
namespace msda {
namespace {

template <int V>
__global__ void __launch_bounds__(256)
add_layernorm_bwd_kernel(const float4 *__restrict__ gy, const float4 *__restrict__ x, const float4 *__restrict__ res,
                         const float4 *__restrict__ gamma, float4 *__restrict__ gv, float *__restrict__ ggamma,
                         float *__restrict__ gbeta, long long rows, float eps) {
    constexpr int C4 = 32 * V;
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    float4 gm[V], dg[V], db[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        gm[i] = gamma[i * 32 + lane];
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += warps) {
        float4 v[V], g[V];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i] = ldg_stream_f4(x + row * C4 + i * 32 + lane);
            if (res != nullptr) {
                const float4 r = ldg_stream_f4(res + row * C4 + i * 32 + lane);
                v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
            }
            g[i] = ldg_stream_f4(gy + row * C4 + i * 32 + lane);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFullMask, sum, o);
        const float mean = sum * (1.f / (4 * C4));
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFullMask, sq, o);
        const float rstd = rsqrtf(sq * (1.f / (4 * C4)) + eps);
        float s1 = 0.f, s2 = 0.f;                     // sum_c g,  sum_c g * xhat   (g = grad_y * gamma)
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;      // xhat
            dg[i].x += g[i].x * v[i].x; dg[i].y += g[i].y * v[i].y; dg[i].z += g[i].z * v[i].z; dg[i].w += g[i].w * v[i].w;
            db[i].x += g[i].x; db[i].y += g[i].y; db[i].z += g[i].z; db[i].w += g[i].w;
            g[i].x *= gm[i].x; g[i].y *= gm[i].y; g[i].z *= gm[i].z; g[i].w *= gm[i].w;
            s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
            s2 += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(kFullMask, s1, o);
            s2 += __shfl_xor_sync(kFullMask, s2, o);
        }
        s1 *= (1.f / (4 * C4));
        s2 *= (1.f / (4 * C4));
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float4 o;
            o.x = rstd * (g[i].x - s1 - v[i].x * s2);
            o.y = rstd * (g[i].y - s1 - v[i].y * s2);
            o.z = rstd * (g[i].z - s1 - v[i].z * s2);
            o.w = rstd * (g[i].w - s1 - v[i].w * s2);
            gv[row * C4 + i * 32 + lane] = o;
        }
    }
    // per-CTA reduction of the 8 warps' partials in shared memory, then one reduction per column per CTA
    __shared__ float part[2][8][128 * V];
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        float *pg = &part[0][w][(i * 32 + lane) * 4], *pb = &part[1][w][(i * 32 + lane) * 4];
        pg[0] = dg[i].x; pg[1] = dg[i].y; pg[2] = dg[i].z; pg[3] = dg[i].w;
        pb[0] = db[i].x; pb[1] = db[i].y; pb[2] = db[i].z; pb[3] = db[i].w;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 128 * V; c += 256) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a += part[0][k][c];
            b += part[1][k][c];
        }
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(ggamma + c), "f"(a) : "memory");
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(gbeta + c), "f"(b) : "memory");
    }
}

template <int V>
cudaError_t launch_bwd(const float *gy, const float *x, const float *res, const float *gamma, float *gv,
                       float *ggamma, float *gbeta, long long rows, float eps, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(ggamma, 0, 128 * V * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(gbeta, 0, 128 * V * sizeof(float), stream);
    if (e != cudaSuccess) return e;
    long long blocks = (rows + 7) / 8;
    const long long cap = (long long)sm_count() * 4;      // few CTAs: each ends with 2 x C reductions
    if (blocks > cap) blocks = cap;
    add_layernorm_bwd_kernel<V><<<(unsigned)blocks, 256, 0, stream>>>(
        reinterpret_cast<const float4 *>(gy), reinterpret_cast<const float4 *>(x),
        reinterpret_cast<const float4 *>(res), reinterpret_cast<const float4 *>(gamma),
        reinterpret_cast<float4 *>(gv), ggamma, gbeta, rows, eps);
    note_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_add_layernorm_bwd(const float *gy, const float *x, const float *res, const float *gamma, float *gv,
                                     float *ggamma, float *gbeta, long long rows, int cols, float eps,
                                     cudaStream_t stream, bool *handled) {
    *handled = true;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(x) |
                           reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(gamma) |
                           reinterpret_cast<uintptr_t>(gv);
    if (bits % 16 != 0) {
        *handled = false;
        return cudaSuccess;
    }
    switch (cols) {
        case 128: return launch_bwd<1>(gy, x, res, gamma, gv, ggamma, gbeta, rows, eps, stream);
        case 256: return launch_bwd<2>(gy, x, res, gamma, gv, ggamma, gbeta, rows, eps, stream);
        default: *handled = false; return cudaSuccess;
    }
}

}  // namespace msda

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

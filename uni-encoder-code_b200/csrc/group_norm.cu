// group_norm.cu -- GroupNorm on NCHW fp32 maps with the pixel decoder's epilogues fused in.
//
// around the encoder, MSDeformAttnPixelDecoder.forward_features normalises five
// feature maps with GroupNorm(32, conv_dim): the three input projections (msdeformattn.py:233-248,
// :341-345) and, at stride 4, the lateral and the output convolution of the FPN level
// (msdeformattn.py:286-300, :369-379), where it also adds the bilinearly up-sampled coarser map to
// the lateral (`y = cur + F.interpolate(out[-1], size=cur.shape[-2:], mode="bilinear",
// align_corners=False)`) and applies ReLU after the output convolution.  In torch eager that is a
// moments kernel + an elementwise kernel per GroupNorm, an up-sampling kernel and an add: at
// BASELINE configs[2] (two 1 GB maps at stride 4) 4.0 + 2.1 + 0.4 ms of a 36 ms decoder forward,
// all of it HBM-bound work running at a fraction of the roofline.
//
// Here: two launches per map.
//   1. group_stats_kernel   per (image, group) chunk -- contiguous in NCHW -- kSplit CTAs each sum
//                           (x - K) and (x - K)^2 over their part (K = the chunk's first element:
//                           the shifted sums keep E[x^2] - E[x]^2 from cancelling), in fp32 per
//                           thread over a few thousand elements, combined in fp64;
//   2. group_apply_kernel   y = (x - mean) * rstd * gamma[c] + beta[c]  [ReLU]  [+ up-sampled src],
//                           one pass, 128-bit accesses; the kSplit partials of the group are
//                           combined by every CTA itself (no atomics: deterministic).
// Algorithmic bytes: read x twice + write y once = 12 B per element (+ 1/4 of a read for the
// up-sampled map); torch's sequence moves 12 (GroupNorm) + 4 + 12 (up-sample, add) = 28.
//
// Inference only (no backward); the reference semantics are torch.nn.functional.group_norm and
// F.interpolate(mode="bilinear", align_corners=False).
#include <cuda_runtime.h>
#include <cstdint>

#include "msda_common.cuh"

namespace msda {

namespace {

constexpr int kGnSplit = 16;       // CTAs per (image, group) chunk in the statistics pass
constexpr int kGnThreads = 512;

// CBIAS: a per-channel constant (the bias of the convolution that produced x, left out of that
// convolution's epilogue) is added to x on the fly: group_norm(x + cbias[c])
template <bool CBIAS>
__global__ void __launch_bounds__(kGnThreads)
group_stats_kernel(const float *__restrict__ x, double *__restrict__ partial, long long chunk /* elements */,
                   const float *__restrict__ cbias, int groups, int cpg, unsigned hw4 /* float4 per channel plane */) {
    // blockIdx.x = (image * groups + group), blockIdx.y = part
    const float *base = x + (long long)blockIdx.x * chunk;
    const long long chunk4 = chunk >> 2;
    const long long per = (chunk4 + gridDim.y - 1) / gridDim.y;
    const long long beg = (long long)blockIdx.y * per, end = min(beg + per, chunk4);
    const float *cb = CBIAS ? cbias + (blockIdx.x % (unsigned)groups) * cpg : nullptr;   // the group's first channel
    const float K = __ldg(base) + (CBIAS ? __ldg(cb) : 0.f);
    float s = 0.f, q = 0.f;
    const float4 *b4 = reinterpret_cast<const float4 *>(base);
    for (long long i = beg + threadIdx.x; i < end; i += kGnThreads) {
        const float4 v = ldg_stream_f4(b4 + i);
        const float Kc = CBIAS ? K - __ldg(cb + (unsigned)i / hw4) : K;      // (x + b_c) - K
        const float a = v.x - Kc, b = v.y - Kc, c = v.z - Kc, d = v.w - Kc;
        s += (a + b) + (c + d);
        q = fmaf(a, a, fmaf(b, b, fmaf(c, c, fmaf(d, d, q))));
    }
    double ds = (double)s, dq = (double)q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ds += __shfl_xor_sync(kFullMask, ds, o);
        dq += __shfl_xor_sync(kFullMask, dq, o);
    }
    __shared__ double ws[kGnThreads / 32], wq[kGnThreads / 32];
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = ds; wq[threadIdx.x >> 5] = dq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < kGnThreads / 32; ++w) { a += ws[w]; b += wq[w]; }
        double *out = partial + ((long long)blockIdx.x * gridDim.y + blockIdx.y) * 2;
        out[0] = a;
        out[1] = b;
    }
}

// source index of torch's bilinear up-sampling with align_corners=False
// (area_pixel_compute_source_index: scale * (dst + 0.5) - 0.5, clamped at 0)
__device__ __forceinline__ void up_index(int dst, float scale, int in_size, int &i0, int &i1, float &l1) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = min((int)src, in_size - 1);
    i1 = min(i0 + 1, in_size - 1);
    l1 = src - (float)i0;
}

template <bool RELU, bool UP>
__global__ void __launch_bounds__(256)
group_apply_kernel(const float *__restrict__ x, const double *__restrict__ partial, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float *__restrict__ y, int C, int H, int W, int cpg, float eps,
                   const float *__restrict__ up, int uh, int uw, float sh, float sw,
                   const float *__restrict__ cbias) {
    // blockIdx.x = image * C + channel (one plane), blockIdx.y = part of the plane
    const int plane = blockIdx.x, c = plane % C, ng = plane / cpg;      // ng = image * groups + group
    __shared__ float s_scale, s_shift;
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        const double *p = partial + (long long)ng * kGnSplit * 2;
        for (int k = 0; k < kGnSplit; ++k) { a += p[2 * k]; b += p[2 * k + 1]; }
        const double n = (double)cpg * H * W;
        const int g0 = (ng % (C / cpg)) * cpg;                           // the group's first channel
        const double K = (double)(__ldg(x + (long long)ng * cpg * H * W) + (cbias ? __ldg(cbias + g0) : 0.f));
        const double mean_k = a / n;                                     // mean of (x + cbias - K)
        double var = b / n - mean_k * mean_k;
        var = var < 0.0 ? 0.0 : var;
        const float mean = (float)(mean_k + K);
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float sc = rstd * gamma[c];
        s_scale = sc;
        s_shift = beta[c] + ((cbias ? __ldg(cbias + c) : 0.f) - mean) * sc;
    }
    __syncthreads();
    const float sc = s_scale, sf = s_shift;
    const long long hw4 = ((long long)H * W) >> 2;
    const long long per = (hw4 + gridDim.y - 1) / gridDim.y;
    const long long beg = (long long)blockIdx.y * per, end = min(beg + per, hw4);
    const float4 *x4 = reinterpret_cast<const float4 *>(x) + (long long)plane * hw4;
    float4 *y4 = reinterpret_cast<float4 *>(y) + (long long)plane * hw4;
    const float *uplane = UP ? up + (long long)plane * uh * uw : nullptr;
    const int w4 = W >> 2;
    for (long long i = beg + threadIdx.x; i < end; i += 256) {
        const float4 v = ldg_stream_f4(x4 + i);
        float4 o = make_float4(fmaf(v.x, sc, sf), fmaf(v.y, sc, sf), fmaf(v.z, sc, sf), fmaf(v.w, sc, sf));
        if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        if (UP) {
            const int row = (int)(i / w4), col = (int)(i - (long long)row * w4) * 4;
            int y0, y1;
            float ly;
            up_index(row, sh, uh, y0, y1, ly);
            const float *r0 = uplane + (long long)y0 * uw, *r1 = uplane + (long long)y1 * uw;
            float add[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int x0, x1;
                float lx;
                up_index(col + k, sw, uw, x0, x1, lx);
                const float top = fmaf(lx, __ldg(r0 + x1), (1.f - lx) * __ldg(r0 + x0));
                const float bot = fmaf(lx, __ldg(r1 + x1), (1.f - lx) * __ldg(r1 + x0));
                add[k] = fmaf(ly, bot, (1.f - ly) * top);
            }
            o.x += add[0]; o.y += add[1]; o.z += add[2]; o.w += add[3];
        }
        y4[i] = o;
    }
}

// The same normalisation written as ROWS: y_rows[n][pixel][c] (row = the C channels of one pixel, `row_stride`
// floats apart, images `image_stride` floats apart) -- the layout the deformable encoder consumes
// (`src.flatten(2).transpose(1, 2)`, msdeformattn.py:72-79), so that the per-level maps land directly in
// their slice of the concatenated [N, S, C] tensor and torch's transposing `cat` (one more read and write of
// every map) disappears.  One CTA = 32 channels x a range of pixels; a warp turns 32 x 32 tiles through shared
// memory: 128-byte reads along the pixels of one channel, 128-bit stores along the channels of one pixel.
constexpr int kRowsWarps = 8;
constexpr int kRowsTilesPerWarp = 4;      // 32-pixel tiles per warp and CTA

__global__ void __launch_bounds__(kRowsWarps * 32)
group_apply_rows_kernel(const float *__restrict__ x, const double *__restrict__ partial, const float *__restrict__ gamma,
                        const float *__restrict__ beta, float *__restrict__ y, int C, long long hw, int cpg, float eps,
                        const float *__restrict__ cbias, long long row_stride, long long image_stride, int relu) {
    // blockIdx.x = image * (C / 32) + channel block, blockIdx.y = pixel range
    __shared__ float s_scale[32], s_shift[32];
    __shared__ float s_tile[kRowsWarps][32 * 33];
    const int cblocks = C / 32, n = blockIdx.x / cblocks, c0 = (blockIdx.x % cblocks) * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        const int c = c0 + threadIdx.x, groups = C / cpg, g = c / cpg;
        const long long ng = (long long)n * groups + g;
        double a = 0.0, b = 0.0;
        const double *p = partial + ng * kGnSplit * 2;
        for (int k = 0; k < kGnSplit; ++k) { a += p[2 * k]; b += p[2 * k + 1]; }
        const double cnt = (double)cpg * (double)hw;
        const double K = (double)(__ldg(x + ng * cpg * hw) + (cbias ? __ldg(cbias + g * cpg) : 0.f));
        const double mean_k = a / cnt;
        double var = b / cnt - mean_k * mean_k;
        var = var < 0.0 ? 0.0 : var;
        const float mean = (float)(mean_k + K);
        const float sc = (float)(1.0 / sqrt(var + (double)eps)) * gamma[c];
        s_scale[threadIdx.x] = sc;
        s_shift[threadIdx.x] = beta[c] + ((cbias ? __ldg(cbias + c) : 0.f) - mean) * sc;
    }
    __syncthreads();
    float *tile = s_tile[warp];
    const float *xin = x + ((long long)n * C + c0) * hw;               // channel c0 of image n
    float *yout = y + (long long)n * image_stride + c0;
    const int sub = lane >> 3, ch4 = (lane & 7) * 4;
    const long long first = ((long long)blockIdx.y * kRowsWarps + warp) * kRowsTilesPerWarp;
    for (int t = 0; t < kRowsTilesPerWarp; ++t) {
        const long long p0 = (first + t) * 32;
        if (p0 >= hw) break;
        const long long p = p0 + lane;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const float v = p < hw ? ldg_stream_f1(xin + (long long)j * hw + p) : 0.f;
            float o = fmaf(v, s_scale[j], s_shift[j]);
            if (relu) o = fmaxf(o, 0.f);
            tile[j * 33 + lane] = o;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int pix = i * 4 + sub;
            const float4 o = make_float4(tile[(ch4 + 0) * 33 + pix], tile[(ch4 + 1) * 33 + pix],
                                         tile[(ch4 + 2) * 33 + pix], tile[(ch4 + 3) * 33 + pix]);
            if (p0 + pix < hw) *reinterpret_cast<float4 *>(yout + (p0 + pix) * row_stride + ch4) = o;
        }
        __syncwarp();
    }
}

}  // namespace

// workspace: N * groups * kGnSplit * 2 doubles
cudaError_t launch_group_norm(const float *x, const float *cbias, const float *gamma, const float *beta, float *y,
                              int N, int C, int H, int W, int groups, float eps, int relu, const float *up, int uh,
                              int uw, double *workspace, cudaStream_t stream, bool *handled) {
    *handled = false;
    const long long hw = (long long)H * W;
    // planes are walked as float4 (H*W % 4 == 0); only the up-sampled add needs 4 outputs of one row (W % 4 == 0)
    if (groups <= 0 || C % groups != 0 || (hw & 3) || (up != nullptr && (W & 3)) || hw > 0x7fffffffLL ||
        (long long)N * C > 0x7fffffffLL || (long long)N * groups > 0x7fffffffLL)
        return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) return cudaSuccess;
    if (up != nullptr && (uh <= 0 || uw <= 0)) return cudaSuccess;
    const int cpg = C / groups;
    const long long chunk = (long long)cpg * hw;
    if (cbias != nullptr && chunk > 0x7fffffffLL) return cudaSuccess;    // 32-bit channel lookup in the statistics pass
    *handled = true;
    const dim3 sgrid((unsigned)(N * groups), kGnSplit);
    if (cbias != nullptr)
        group_stats_kernel<true><<<sgrid, kGnThreads, 0, stream>>>(x, workspace, chunk, cbias, groups, cpg, (unsigned)(hw >> 2));
    else
        group_stats_kernel<false><<<sgrid, kGnThreads, 0, stream>>>(x, workspace, chunk, nullptr, groups, cpg, 1u);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // parts per plane: enough CTAs for the whole GPU, each with a few thousand float4
    int parts = (int)((hw / 4 + 8191) / 8192);
    if (parts < 1) parts = 1;
    if (parts > 64) parts = 64;
    const dim3 grid((unsigned)(N * C), (unsigned)parts);
    const float sh = up ? (float)uh / (float)H : 0.f, sw = up ? (float)uw / (float)W : 0.f;
#define GN_APPLY(R, U) group_apply_kernel<R, U><<<grid, 256, 0, stream>>>(x, workspace, gamma, beta, y, C, H, W, cpg, eps, up, uh, uw, sh, sw, cbias)
    if (up != nullptr) { if (relu) GN_APPLY(true, true); else GN_APPLY(false, true); }
    else { if (relu) GN_APPLY(true, false); else GN_APPLY(false, false); }
#undef GN_APPLY
    note_launch();
    return cudaGetLastError();
}

// GroupNorm of x[N, C, H*W] written as rows: y[n * image_stride + pixel * row_stride + c]
cudaError_t launch_group_norm_rows(const float *x, const float *cbias, const float *gamma, const float *beta, float *y,
                                   int N, int C, long long hw, int groups, float eps, int relu, long long row_stride,
                                   long long image_stride, double *workspace, cudaStream_t stream, bool *handled) {
    *handled = false;
    if (groups <= 0 || C % groups != 0 || C % 32 != 0 || (hw & 3) || hw > 0x7fffffffLL || row_stride < C ||
        (row_stride & 3) || (image_stride & 3) || (long long)N * (C / 32) > 0x7fffffffLL ||
        (long long)N * groups > 0x7fffffffLL)
        return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) return cudaSuccess;
    const int cpg = C / groups;
    const long long chunk = (long long)cpg * hw;
    if (cbias != nullptr && chunk > 0x7fffffffLL) return cudaSuccess;
    const long long per_cta = 32LL * kRowsWarps * kRowsTilesPerWarp;
    const long long parts = (hw + per_cta - 1) / per_cta;
    if (parts > 65535) return cudaSuccess;
    *handled = true;
    const dim3 sgrid((unsigned)(N * groups), kGnSplit);
    if (cbias != nullptr)
        group_stats_kernel<true><<<sgrid, kGnThreads, 0, stream>>>(x, workspace, chunk, cbias, groups, cpg, (unsigned)(hw >> 2));
    else
        group_stats_kernel<false><<<sgrid, kGnThreads, 0, stream>>>(x, workspace, chunk, nullptr, groups, cpg, 1u);
    note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    group_apply_rows_kernel<<<dim3((unsigned)(N * (C / 32)), (unsigned)parts), kRowsWarps * 32, 0, stream>>>(
        x, workspace, gamma, beta, y, C, hw, cpg, eps, cbias, row_stride, image_stride, relu);
    note_launch();
    return cudaGetLastError();
}

int group_norm_workspace_doubles(int N, int groups) { return N * groups * kGnSplit * 2; }

// x[n, c, :, :] += bias[c] in place (the bias of a convolution run without one), 128-bit accesses
namespace {
__global__ void __launch_bounds__(256)
channel_bias_kernel(float4 *__restrict__ x, const float *__restrict__ bias, int C, long long hw4) {
    const int plane = blockIdx.x;
    const float b = __ldg(bias + plane % C);
    const long long per = (hw4 + gridDim.y - 1) / gridDim.y;
    const long long beg = (long long)blockIdx.y * per, end = min(beg + per, hw4);
    float4 *p = x + (long long)plane * hw4;
    for (long long i = beg + threadIdx.x; i < end; i += 256) {
        float4 v = ldg_stream_f4(p + i);
        v.x += b; v.y += b; v.z += b; v.w += b;
        p[i] = v;
    }
}
}  // namespace

cudaError_t launch_channel_bias(float *x, const float *bias, int N, int C, long long hw, cudaStream_t stream,
                                bool *handled) {
    *handled = false;
    if ((hw & 3) || (long long)N * C > 0x7fffffffLL || (reinterpret_cast<uintptr_t>(x) & 15u)) return cudaSuccess;
    *handled = true;
    int parts = (int)((hw / 4 + 8191) / 8192);
    parts = parts < 1 ? 1 : parts > 64 ? 64 : parts;
    channel_bias_kernel<<<dim3((unsigned)(N * C), (unsigned)parts), 256, 0, stream>>>(
        reinterpret_cast<float4 *>(x), bias, C, hw >> 2);
    note_launch();
    return cudaGetLastError();
}

}  // namespace msda

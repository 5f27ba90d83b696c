// msda_generic.cu -- shape-generic MSDA kernels (any channels-per-head, level and
// point count; fp32 and fp64) and the integer known-answer kernel.
//
// The 32-channel fp32 kernels in msda_fwd.cu / msda_bwd.cu are the hot path; these
// cover the rest of the reference op's domain (AT_DISPATCH_FLOATING_TYPES at
// ms_deform_attn_cuda.cu:69,139 instantiates float and double, and the col2im
// launcher accepts any channel count, ms_deform_im2col_cuda.cuh:980-1323).
//
// One warp per (image, query, head) row; lanes stride over channels, so a corner
// row is read as one coalesced segment.  grad_sampling_loc / grad_attn_weight are
// warp-shuffle sums (no shared memory, no barrier); grad_value uses scalar atomics.
#include "msda_common.cuh"

namespace msda {

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const T *__restrict__ value, const int64_t *__restrict__ shapes,
                        const int64_t *__restrict__ lstart, const T *__restrict__ loc,
                        const T *__restrict__ attw, const Dims d, T *__restrict__ out) {
    __shared__ LevelTable lt;
    fill_level_table(lt, shapes, lstart, d.L, d.P, d.S, d.Lq, 8, 1, 8, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long rows = (long long)d.N * d.Lq * d.M;
    const long long pix = (long long)d.M * d.D;
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += warps) {
        const int m = (int)(row % d.M);
        const long long n = row / ((long long)d.Lq * d.M);
        const T *vb = value + n * (long long)d.S * pix + (long long)m * d.D;
        for (int c0 = 0; c0 < d.D; c0 += 32) {
            const int c = c0 + lane;
            T acc = 0;
            for (int l = 0; l < d.L; ++l) {
                const int H = lt.H[l], W = lt.W[l];
                const T *lb = vb + (long long)lt.start[l] * pix;
                for (int p = 0; p < d.P; ++p) {
                    const long long s = (row * d.L + l) * d.P + p;
                    const Geom<T> g = decompose(loc[2 * s], loc[2 * s + 1], H, W);
                    if (g.valid && c < d.D) {
                        const T hh = 1 - g.lh, hw = 1 - g.lw;
                        const T *p0 = lb + ((long long)g.h_low * W + g.w_low) * pix + c;
#ifdef MSDA_CHECK_BOUNDS
                        {
                            const T *lo = value + n * (long long)d.S * pix, *hi = lo + (long long)d.S * pix;
                            assert(!(g.cmask & 1) || (p0 >= lo && p0 < hi));
                            assert(!(g.cmask & 8) || (p0 + (long long)W * pix + pix >= lo && p0 + (long long)W * pix + pix < hi));
                        }
#endif
                        const T v1 = (g.cmask & 1) ? p0[0] : (T)0;
                        const T v2 = (g.cmask & 2) ? p0[pix] : (T)0;
                        const T v3 = (g.cmask & 4) ? p0[(long long)W * pix] : (T)0;
                        const T v4 = (g.cmask & 8) ? p0[(long long)W * pix + pix] : (T)0;
                        const T val = (hh * hw) * v1 + (hh * g.lw) * v2 + (g.lh * hw) * v3 + (g.lh * g.lw) * v4;
                        acc += val * attw[s];
                    }
                }
            }
            if (c < d.D) out[row * d.D + c] = acc;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
msda_bwd_generic_kernel(const T *__restrict__ grad_out, const T *__restrict__ value,
                        const int64_t *__restrict__ shapes, const int64_t *__restrict__ lstart,
                        const T *__restrict__ loc, const T *__restrict__ attw, const Dims d,
                        T *__restrict__ grad_value, T *__restrict__ grad_loc,
                        T *__restrict__ grad_attw) {
    __shared__ LevelTable lt;
    fill_level_table(lt, shapes, lstart, d.L, d.P, d.S, d.Lq, 8, 1, 8, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long rows = (long long)d.N * d.Lq * d.M;
    const long long pix = (long long)d.M * d.D;
    for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows;
         row += warps) {
        const int m = (int)(row % d.M);
        const long long n = row / ((long long)d.Lq * d.M);
        const long long img = n * (long long)d.S * pix + (long long)m * d.D;
        for (int l = 0; l < d.L; ++l) {
            const int H = lt.H[l], W = lt.W[l];
            const long long lvl = img + (long long)lt.start[l] * pix;
            for (int p = 0; p < d.P; ++p) {
                const long long s = (row * d.L + l) * d.P + p;
                const Geom<T> g = decompose(loc[2 * s], loc[2 * s + 1], H, W);
                T gx = 0, gy = 0, ga = 0;
                if (g.valid) {
                    const T a = attw[s];
                    const T hh = 1 - g.lh, hw = 1 - g.lw;
                    const T cw[4] = {hh * hw, hh * g.lw, g.lh * hw, g.lh * g.lw};
                    const T dh[4] = {-hw, -g.lw, hw, g.lw};     // d val / d h per corner
                    const T dw[4] = {-hh, hh, -g.lh, g.lh};     // d val / d w per corner
                    const long long p0 = lvl + ((long long)g.h_low * W + g.w_low) * pix;
                    const long long co[4] = {0, pix, (long long)W * pix, (long long)W * pix + pix};
                    for (int c = lane; c < d.D; c += 32) {
                        const T tg = grad_out[row * d.D + c];
                        const T tgv = tg * a;
                        T val = 0, sh = 0, sw = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (g.cmask >> k & 1) {
                                const long long o = p0 + co[k] + c;
#ifdef MSDA_CHECK_BOUNDS
                                assert(o >= n * (long long)d.S * pix && o < (n + 1) * (long long)d.S * pix);
#endif
                                const T v = value[o];
                                sh += dh[k] * v;
                                sw += dw[k] * v;
                                val += cw[k] * v;
                                atomicAdd(grad_value + o, cw[k] * tgv);
                            }
                        }
                        ga += tg * val;
                        gx += (T)W * sw * tgv;
                        gy += (T)H * sh * tgv;
                    }
                }
                gx = warp_sum(gx);
                gy = warp_sum(gy);
                ga = warp_sum(ga);
                if (lane == 0) {
                    grad_loc[2 * s] = gx;
                    grad_loc[2 * s + 1] = gy;
                    grad_attw[s] = ga;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256)
msda_debug_indices_kernel(const int64_t *__restrict__ shapes, const int64_t *__restrict__ lstart,
                          const float *__restrict__ loc, const Dims d, int32_t *__restrict__ idx,
                          int64_t *__restrict__ off) {
    __shared__ LevelTable lt;
    fill_level_table(lt, shapes, lstart, d.L, d.P, d.S, d.Lq, 8, 1, 8, 0);
    __syncthreads();
    const long long total = (long long)d.N * d.Lq * d.M * d.L * d.P;
    const long long pix = (long long)d.M * d.D;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total;
         s += (long long)gridDim.x * blockDim.x) {
        const int l = (int)((s / d.P) % d.L);
        const long long row = s / ((long long)d.L * d.P);
        const int m = (int)(row % d.M);
        const long long n = row / ((long long)d.Lq * d.M);
        const int H = lt.H[l], W = lt.W[l];
        const Geom<float> g = decompose(loc[2 * s], loc[2 * s + 1], H, W);
        idx[4 * s + 0] = g.valid;
        idx[4 * s + 1] = g.h_low;
        idx[4 * s + 2] = g.w_low;
        idx[4 * s + 3] = g.cmask;
        const long long p0 = n * (long long)d.S * pix + (long long)m * d.D +
                             ((long long)lt.start[l] + (long long)g.h_low * W + g.w_low) * pix;
        off[4 * s + 0] = (g.cmask & 1) ? p0 : -1;
        off[4 * s + 1] = (g.cmask & 2) ? p0 + pix : -1;
        off[4 * s + 2] = (g.cmask & 4) ? p0 + (long long)W * pix : -1;
        off[4 * s + 3] = (g.cmask & 8) ? p0 + (long long)W * pix + pix : -1;
    }
}

static unsigned grid_for(long long work_items, int per_block) {
    long long b = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)sm_count() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

template <typename T>
cudaError_t launch_fwd_generic(const T *value, const int64_t *shapes, const int64_t *lstart,
                               const T *loc, const T *attw, const Dims &d, T *out,
                               cudaStream_t stream) {
    const long long rows = (long long)d.N * d.Lq * d.M;
    msda_fwd_generic_kernel<T><<<grid_for(rows, 8), 256, 0, stream>>>(value, shapes, lstart, loc,
                                                                       attw, d, out);
    note_launch();
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_bwd_generic(const T *grad_out, const T *value, const int64_t *shapes,
                               const int64_t *lstart, const T *loc, const T *attw, const Dims &d,
                               T *gv, T *gl, T *gw, cudaStream_t stream) {
    const long long rows = (long long)d.N * d.Lq * d.M;
    msda_bwd_generic_kernel<T><<<grid_for(rows, 8), 256, 0, stream>>>(grad_out, value, shapes, lstart,
                                                                       loc, attw, d, gv, gl, gw);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_debug_indices(const int64_t *shapes, const int64_t *lstart, const float *loc,
                                 const Dims &d, int32_t *idx, int64_t *off, cudaStream_t stream) {
    const long long total = (long long)d.N * d.Lq * d.M * d.L * d.P;
    msda_debug_indices_kernel<<<grid_for(total, 256), 256, 0, stream>>>(shapes, lstart, loc, d, idx, off);
    note_launch();
    return cudaGetLastError();
}

template cudaError_t launch_fwd_generic<float>(const float *, const int64_t *, const int64_t *,
                                               const float *, const float *, const Dims &, float *,
                                               cudaStream_t);
template cudaError_t launch_fwd_generic<double>(const double *, const int64_t *, const int64_t *,
                                                const double *, const double *, const Dims &,
                                                double *, cudaStream_t);
template cudaError_t launch_bwd_generic<float>(const float *, const float *, const int64_t *,
                                               const int64_t *, const float *, const float *,
                                               const Dims &, float *, float *, float *, cudaStream_t);
template cudaError_t launch_bwd_generic<double>(const double *, const double *, const int64_t *,
                                                const int64_t *, const double *, const double *,
                                                const Dims &, double *, double *, double *,
                                                cudaStream_t);

}  // namespace msda

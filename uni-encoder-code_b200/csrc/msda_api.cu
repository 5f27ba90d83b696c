// msda_api.cu -- the extern "C" boundary declared in include/msda_b200.h.
//
// Host-side counterpart of the reference's ms_deform_attn_cuda_forward/backward
// (ms_deform_attn_cuda.cu:25-85, 88-158): argument validation and kernel choice.
// Tensor-level checks (contiguity, device, dtype, im2col_step divisibility) live in
// the Python shim because they need tensor metadata; this layer sees raw pointers.
#include <atomic>
#include <cstdint>
#include <cstring>

#include "../../include/msda_b200.h"
#include "msda_common.cuh"

namespace msda {

// implemented in the kernel translation units
cudaError_t launch_fwd_d32(const float *, const int64_t *, const int64_t *, const float *,
                           const float *, const Dims &, float *, cudaStream_t, bool *handled);
cudaError_t launch_bwd_d32(const float *, const float *, const int64_t *, const int64_t *,
                           const float *, const float *, const Dims &, float *, float *, float *,
                           cudaStream_t, bool *handled, int gate);
cudaError_t launch_bwd_sorted(const float *, const float *, const int64_t *, const int64_t *,
                              const float *, const float *, const Dims &, float *, float *, float *,
                              cudaStream_t, bool *handled, int gate);
bool bwd_sorted_applies(const float *value, const float *grad_value, const Dims &);
cudaError_t launch_fwd_d32_fused(const float *, const int64_t *, const int64_t *, const float *,
                                 long long, const float *, const float *, const Dims &, float *,
                                 cudaStream_t, bool *handled, int off_qstride = 0, int logit_qstride = 0,
                                 const float *off_table = nullptr, const float *logit_table = nullptr);
cudaError_t launch_bwd_d32_fused(const float *, const float *, const int64_t *, const int64_t *,
                                 const float *, long long, const float *, const float *,
                                 const Dims &, float *, float *, float *, cudaStream_t, bool *handled, int gate);
cudaError_t launch_bwd_sorted_fused(const float *, const float *, const int64_t *, const int64_t *,
                                    const float *, long long, const float *, const float *,
                                    const Dims &, float *, float *, float *, cudaStream_t, bool *handled, int gate);
template <typename T>
cudaError_t launch_fwd_generic(const T *, const int64_t *, const int64_t *, const T *, const T *,
                               const Dims &, T *, cudaStream_t);
template <typename T>
cudaError_t launch_bwd_generic(const T *, const T *, const int64_t *, const int64_t *, const T *,
                               const T *, const Dims &, T *, T *, T *, cudaStream_t);
cudaError_t launch_linear_tf32x3(const float *, const float *, const float *, float *, int, int, int, int,
                                 float *, cudaStream_t, bool *handled);
cudaError_t launch_split_weight(const float *, float *, int, int, cudaStream_t, bool *handled);
cudaError_t launch_add_layernorm(const float *, const float *, const float *, const float *, float *, long long,
                                 int, float, cudaStream_t, bool *handled);
cudaError_t launch_add_layernorm_bwd(const float *, const float *, const float *, const float *, float *, float *,
                                     float *, long long, int, float, cudaStream_t, bool *handled);
cudaError_t launch_linear_wgrad(const float *, const float *, float *, float *, long long, int, int, cudaStream_t,
                                bool *handled);
cudaError_t launch_transpose(const float *, float *, long long, int, cudaStream_t);
cudaError_t launch_group_norm(const float *, const float *, const float *, const float *, float *, int, int, int, int,
                              int, float, int, const float *, int, int, double *, cudaStream_t, bool *handled);
cudaError_t launch_channel_bias(float *, const float *, int, int, long long, cudaStream_t, bool *handled);
cudaError_t launch_group_norm_rows(const float *, const float *, const float *, const float *, float *, int, int,
                                   long long, int, float, int, long long, long long, double *, cudaStream_t,
                                   bool *handled);
int group_norm_workspace_doubles(int N, int groups);
cudaError_t launch_debug_indices(const int64_t *, const int64_t *, const float *, const Dims &,
                                 int32_t *, int64_t *, cudaStream_t);

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_options[OPT_COUNT];   // zero-initialised: 0 == auto

void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int option_value(int which) { return g_options[which].load(std::memory_order_relaxed); }

int sm_count() {
    static std::atomic<int> cached[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;   // launches refuse such a device
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

static int check_dims(const Dims &d) {
    if (d.N <= 0 || d.S <= 0 || d.M <= 0 || d.D <= 0 || d.L <= 0 || d.Lq <= 0 || d.P <= 0)
        return MSDA_ERR_BAD_SHAPE;
    if (d.L > MSDA_MAX_LEVELS) return MSDA_ERR_UNSUPPORTED;
    return MSDA_OK;
}

// the 32-channel kernels use 128-bit accesses on value / output rows and 64-bit ones on the
// locations; a contiguous tensor view with an odd storage offset goes to the generic kernels,
// which accept any alignment like the reference's scalar kernels do
static bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool aligned8(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

static int option_index(const char *name) {
    if (!name) return -1;
    if (!strcmp(name, "fwd_variant")) return OPT_FWD_VARIANT;
    if (!strcmp(name, "bwd_variant")) return OPT_BWD_VARIANT;
    if (!strcmp(name, "tile_order")) return OPT_TILE_ORDER;
    if (!strcmp(name, "ctas_per_sm")) return OPT_CTAS_PER_SM;
    if (!strcmp(name, "linear_variant")) return OPT_LINEAR_VARIANT;
    if (!strcmp(name, "wgrad_chunk")) return OPT_WGRAD_CHUNK;
#ifdef MSDA_PROFILE_KNOBS      // `make profile`: knobs that skip work on purpose (WRONG results), timing experiments only
    if (!strcmp(name, "whatif_drop_reds")) return OPT_WHATIF_DROP_REDS;
    if (!strcmp(name, "whatif_linear")) return OPT_WHATIF_LINEAR;
#endif
    return -1;
}

}  // namespace msda

using namespace msda;

extern "C" {

int msda_b200_abi_version(void) { return MSDA_B200_ABI_VERSION; }

const char *msda_b200_error_string(int code) {
    switch (code) {
        case MSDA_OK: return "success";
        case MSDA_ERR_NULL_POINTER: return "msda_b200: a required pointer is NULL";
        case MSDA_ERR_BAD_SHAPE: return "msda_b200: a tensor size is zero or negative";
        case MSDA_ERR_UNSUPPORTED: return "msda_b200: unsupported configuration (num_levels > 16, or a shape the fused path does not cover)";
        case MSDA_ERR_BAD_OPTION: return "msda_b200: unknown option name or value out of range";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "msda_b200: unknown error code";
}

long long msda_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int msda_b200_set_option(const char *name, int value) {
    const int i = option_index(name);
    if (i < 0 || value < 0 || value > 65535) return MSDA_ERR_BAD_OPTION;
    g_options[i].store(value, std::memory_order_relaxed);
    return MSDA_OK;
}

int msda_b200_get_option(const char *name, int *value) {
    const int i = option_index(name);
    if (i < 0 || !value) return MSDA_ERR_BAD_OPTION;
    *value = g_options[i].load(std::memory_order_relaxed);
    return MSDA_OK;
}

int msda_b200_forward_f32(const float *value, const int64_t *spatial_shapes,
                          const int64_t *level_start, const float *sampling_loc,
                          const float *attn_weight, int batch, int spatial_size, int num_heads,
                          int channels, int num_levels, int num_query, int num_point, float *output,
                          void *stream) {
    if (!value || !spatial_shapes || !level_start || !sampling_loc || !attn_weight || !output)
        return MSDA_ERR_NULL_POINTER;
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (int rc = check_dims(d)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    bool handled = false;
    cudaError_t e = cudaSuccess;
    const bool fast_ok = aligned16(value) && aligned16(output) && aligned8(sampling_loc);
    if (option_value(OPT_FWD_VARIANT) != 63 && fast_ok)   // 63 forces the generic kernel (tests)
        e = launch_fwd_d32(value, spatial_shapes, level_start, sampling_loc, attn_weight, d, output, st,
                           &handled);
    if (e == cudaSuccess && !handled)
        e = launch_fwd_generic<float>(value, spatial_shapes, level_start, sampling_loc, attn_weight, d,
                                      output, st);
    return (int)e;
}

int msda_b200_forward_f64(const double *value, const int64_t *spatial_shapes,
                          const int64_t *level_start, const double *sampling_loc,
                          const double *attn_weight, int batch, int spatial_size, int num_heads,
                          int channels, int num_levels, int num_query, int num_point,
                          double *output, void *stream) {
    if (!value || !spatial_shapes || !level_start || !sampling_loc || !attn_weight || !output)
        return MSDA_ERR_NULL_POINTER;
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (int rc = check_dims(d)) return rc;
    return (int)launch_fwd_generic<double>(value, spatial_shapes, level_start, sampling_loc,
                                           attn_weight, d, output, (cudaStream_t)stream);
}

int msda_b200_backward_f32(const float *grad_output, const float *value,
                           const int64_t *spatial_shapes, const int64_t *level_start,
                           const float *sampling_loc, const float *attn_weight, int batch,
                           int spatial_size, int num_heads, int channels, int num_levels,
                           int num_query, int num_point, float *grad_value,
                           float *grad_sampling_loc, float *grad_attn_weight, void *stream) {
    if (!grad_output || !value || !spatial_shapes || !level_start || !sampling_loc || !attn_weight ||
        !grad_value || !grad_sampling_loc || !grad_attn_weight)
        return MSDA_ERR_NULL_POINTER;
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (int rc = check_dims(d)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    bool handled = false;
    cudaError_t e = cudaSuccess;
    const int variant = option_value(OPT_BWD_VARIANT);
    const bool fast_ok = aligned16(value) && aligned16(grad_output) && aligned16(grad_value) &&
                         aligned8(sampling_loc) && aligned8(grad_sampling_loc);
    // 0 (default): where the queries are the value pixels (encoder self-attention) the in-SM merging
    //     kernel (points near their query) and the per-row reduction kernel (points anywhere) are both
    //     launched; every CTA of both runs the same probe of the sampling locations first and the
    //     kernel the verdict goes against returns at once (msda_common.cuh);
    // 20..39: merging kernel forced (and its tuning variants); 1..8: CTA shapes of the per-row
    //     reduction kernel; 63: generic kernel (tests)
    int gate = GATE_NONE;
    if (variant == 0 && fast_ok && bwd_sorted_applies(value, grad_value, d)) {
        e = launch_bwd_sorted(grad_output, value, spatial_shapes, level_start, sampling_loc, attn_weight,
                              d, grad_value, grad_sampling_loc, grad_attn_weight, st, &handled, GATE_RUN_IF_LOCAL);
        handled = false;                     // the per-row kernel follows, gated the other way
        gate = GATE_RUN_IF_SPREAD;
    } else if (variant >= 20 && variant < 40 && fast_ok) {
        e = launch_bwd_sorted(grad_output, value, spatial_shapes, level_start, sampling_loc, attn_weight,
                              d, grad_value, grad_sampling_loc, grad_attn_weight, st, &handled, GATE_NONE);
    }
    if (e == cudaSuccess && !handled && variant != 63 && fast_ok)
        e = launch_bwd_d32(grad_output, value, spatial_shapes, level_start, sampling_loc, attn_weight,
                           d, grad_value, grad_sampling_loc, grad_attn_weight, st, &handled, gate);
    if (e == cudaSuccess && !handled)
        e = launch_bwd_generic<float>(grad_output, value, spatial_shapes, level_start, sampling_loc,
                                      attn_weight, d, grad_value, grad_sampling_loc,
                                      grad_attn_weight, st);
    return (int)e;
}

int msda_b200_backward_f64(const double *grad_output, const double *value,
                           const int64_t *spatial_shapes, const int64_t *level_start,
                           const double *sampling_loc, const double *attn_weight, int batch,
                           int spatial_size, int num_heads, int channels, int num_levels,
                           int num_query, int num_point, double *grad_value,
                           double *grad_sampling_loc, double *grad_attn_weight, void *stream) {
    if (!grad_output || !value || !spatial_shapes || !level_start || !sampling_loc || !attn_weight ||
        !grad_value || !grad_sampling_loc || !grad_attn_weight)
        return MSDA_ERR_NULL_POINTER;
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (int rc = check_dims(d)) return rc;
    return (int)launch_bwd_generic<double>(grad_output, value, spatial_shapes, level_start,
                                           sampling_loc, attn_weight, d, grad_value,
                                           grad_sampling_loc, grad_attn_weight, (cudaStream_t)stream);
}

int msda_b200_fused_forward_strided_f32(const float *value, const int64_t *spatial_shapes,
                                        const int64_t *level_start, const float *reference_points,
                                        long long ref_batch_stride, const float *sampling_offsets,
                                        int offsets_row_stride, const float *attn_logits,
                                        int logits_row_stride, const float *offsets_table,
                                        const float *logits_table, int batch, int spatial_size, int num_heads,
                                        int channels, int num_levels, int num_query, int num_point,
                                        float *output, void *stream) {
    if (!value || !spatial_shapes || !level_start || !reference_points || !sampling_offsets ||
        !attn_logits || !output)
        return MSDA_ERR_NULL_POINTER;
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (int rc = check_dims(d)) return rc;
    if (ref_batch_stride != 0 && ref_batch_stride != (long long)num_query * num_levels * 2)
        return MSDA_ERR_BAD_SHAPE;
    const long long packed = (long long)num_heads * num_levels * num_point;
    if (offsets_row_stride < 2 * packed || logits_row_stride < packed || (offsets_row_stride & 1))
        return MSDA_ERR_BAD_SHAPE;
    if ((offsets_table != nullptr) != (logits_table != nullptr)) return MSDA_ERR_NULL_POINTER;
    if (!aligned16(value) || !aligned16(output) || !aligned8(sampling_offsets) || !aligned8(reference_points) ||
        !aligned8(offsets_table))
        return MSDA_ERR_UNSUPPORTED;
    bool handled = false;
    cudaError_t e = launch_fwd_d32_fused(value, spatial_shapes, level_start, reference_points,
                                         ref_batch_stride, sampling_offsets, attn_logits, d, output,
                                         (cudaStream_t)stream, &handled, offsets_row_stride, logits_row_stride,
                                         offsets_table, logits_table);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_fused_forward_f32(const float *value, const int64_t *spatial_shapes,
                                const int64_t *level_start, const float *reference_points,
                                long long ref_batch_stride, const float *sampling_offsets,
                                const float *attn_logits, int batch, int spatial_size, int num_heads,
                                int channels, int num_levels, int num_query, int num_point,
                                float *output, void *stream) {
    return msda_b200_fused_forward_strided_f32(value, spatial_shapes, level_start, reference_points,
                                               ref_batch_stride, sampling_offsets,
                                               2 * num_heads * num_levels * num_point, attn_logits,
                                               num_heads * num_levels * num_point, nullptr, nullptr, batch, spatial_size,
                                               num_heads, channels, num_levels, num_query, num_point, output,
                                               stream);
}

int msda_b200_fused_backward_f32(const float *grad_output, const float *value,
                                 const int64_t *spatial_shapes, const int64_t *level_start,
                                 const float *reference_points, long long ref_batch_stride,
                                 const float *sampling_offsets, const float *attn_logits, int batch,
                                 int spatial_size, int num_heads, int channels, int num_levels,
                                 int num_query, int num_point, float *grad_value,
                                 float *grad_sampling_offsets, float *grad_attn_logits, void *stream) {
    if (!grad_output || !value || !spatial_shapes || !level_start || !reference_points ||
        !sampling_offsets || !attn_logits || !grad_value || !grad_sampling_offsets || !grad_attn_logits)
        return MSDA_ERR_NULL_POINTER;
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (int rc = check_dims(d)) return rc;
    if (ref_batch_stride != 0 && ref_batch_stride != (long long)num_query * num_levels * 2)
        return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = cudaSuccess;
    const int variant = option_value(OPT_BWD_VARIANT);
    const bool aligned = aligned16(value) && aligned16(grad_output) && aligned16(grad_value) &&
                         aligned8(sampling_offsets) && aligned8(grad_sampling_offsets) && aligned8(reference_points);
    if (!aligned) return MSDA_ERR_UNSUPPORTED;
    // as in msda_b200_backward_f32: merging kernel and per-row kernel behind the same location probe
    // (here on the raw offsets), or one of them forced by "bwd_variant" (20..39 / 1..8)
    int gate = GATE_NONE;
    if ((variant == 0 || (variant >= 20 && variant < 40)) && bwd_sorted_applies(value, grad_value, d)) {
        e = launch_bwd_sorted_fused(grad_output, value, spatial_shapes, level_start, reference_points,
                                    ref_batch_stride, sampling_offsets, attn_logits, d, grad_value,
                                    grad_sampling_offsets, grad_attn_logits, (cudaStream_t)stream, &handled,
                                    variant == 0 ? GATE_RUN_IF_LOCAL : GATE_NONE);
        if (variant == 0 && handled) { handled = false; gate = GATE_RUN_IF_SPREAD; }
    }
    if (e == cudaSuccess && !handled)
        e = launch_bwd_d32_fused(grad_output, value, spatial_shapes, level_start,
                                 reference_points, ref_batch_stride, sampling_offsets,
                                 attn_logits, d, grad_value, grad_sampling_offsets,
                                 grad_attn_logits, (cudaStream_t)stream, &handled, gate);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_linear_f32(const float *x, const float *weight, const float *bias, float *y, int rows,
                         int out_features, int in_features, int relu, float *workspace, void *stream) {
    if (!x || !weight || !y) return MSDA_ERR_NULL_POINTER;
    if (rows <= 0 || out_features <= 0 || in_features <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_linear_tf32x3(x, weight, bias, y, rows, out_features, in_features, relu,
                                         workspace, (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_split_weight_f32(const float *weight, float *split_weight, int out_features, int in_features,
                               void *stream) {
    if (!weight || !split_weight) return MSDA_ERR_NULL_POINTER;
    if (out_features <= 0 || in_features <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_split_weight(weight, split_weight, out_features, in_features, (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_linear_presplit_f32(const float *x, const float *split_weight, const float *bias, float *y, int rows,
                                  int out_features, int in_features, int relu, void *stream) {
    if (!x || !split_weight || !y) return MSDA_ERR_NULL_POINTER;
    if (rows <= 0 || out_features <= 0 || in_features <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    // the kernel only reads the split weight
    cudaError_t e = launch_linear_tf32x3(x, nullptr, bias, y, rows, out_features, in_features, relu,
                                         const_cast<float *>(split_weight), (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_add_layernorm_f32(const float *x, const float *residual, const float *gamma, const float *beta,
                                float *y, long long rows, int cols, float eps, void *stream) {
    if (!x || !gamma || !beta || !y) return MSDA_ERR_NULL_POINTER;
    if (rows <= 0 || cols <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_add_layernorm(x, residual, gamma, beta, y, rows, cols, eps, (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_add_layernorm_backward_f32(const float *grad_y, const float *x, const float *residual,
                                         const float *gamma, float *grad_v, float *grad_gamma, float *grad_beta,
                                         long long rows, int cols, float eps, void *stream) {
    if (!grad_y || !x || !gamma || !grad_v || !grad_gamma || !grad_beta) return MSDA_ERR_NULL_POINTER;
    if (rows <= 0 || cols <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_add_layernorm_bwd(grad_y, x, residual, gamma, grad_v, grad_gamma, grad_beta, rows, cols,
                                             eps, (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_linear_wgrad_f32(const float *grad_y, const float *x, float *grad_weight, float *grad_bias,
                               long long rows, int out_features, int in_features, void *stream) {
    if (!grad_y || !x || !grad_weight) return MSDA_ERR_NULL_POINTER;
    if (rows <= 0 || out_features <= 0 || in_features <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_linear_wgrad(grad_y, x, grad_weight, grad_bias, rows, out_features, in_features,
                                        (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_group_norm_nchw_f32(const float *x, const float *channel_bias, const float *gamma, const float *beta,
                                  float *y, int batch, int channels, int height, int width, int groups, float eps,
                                  int relu, const float *up, int up_h, int up_w, void *workspace, void *stream) {
    if (!x || !gamma || !beta || !y || !workspace) return MSDA_ERR_NULL_POINTER;
    if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || groups <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_group_norm(x, channel_bias, gamma, beta, y, batch, channels, height, width, groups, eps,
                                      relu, up, up_h, up_w, static_cast<double *>(workspace), (cudaStream_t)stream,
                                      &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_group_norm_nchw_to_rows_f32(const float *x, const float *channel_bias, const float *gamma,
                                          const float *beta, float *y_rows, long long row_stride,
                                          long long image_stride, int batch, int channels, long long plane,
                                          int groups, float eps, int relu, void *workspace, void *stream) {
    if (!x || !gamma || !beta || !y_rows || !workspace) return MSDA_ERR_NULL_POINTER;
    if (batch <= 0 || channels <= 0 || plane <= 0 || groups <= 0 || row_stride < channels || image_stride < 0)
        return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_group_norm_rows(x, channel_bias, gamma, beta, y_rows, batch, channels, plane, groups, eps,
                                           relu, row_stride, image_stride, static_cast<double *>(workspace),
                                           (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

int msda_b200_add_channel_bias_nchw_f32(float *x, const float *bias, int batch, int channels, long long plane,
                                        void *stream) {
    if (!x || !bias) return MSDA_ERR_NULL_POINTER;
    if (batch <= 0 || channels <= 0 || plane <= 0) return MSDA_ERR_BAD_SHAPE;
    bool handled = false;
    cudaError_t e = launch_channel_bias(x, bias, batch, channels, plane, (cudaStream_t)stream, &handled);
    if (e == cudaSuccess && !handled) return MSDA_ERR_UNSUPPORTED;
    return (int)e;
}

long long msda_b200_group_norm_workspace_bytes(int batch, int groups) {
    if (batch <= 0 || groups <= 0) return 0;
    return (long long)group_norm_workspace_doubles(batch, groups) * (long long)sizeof(double);
}

int msda_b200_transpose_f32(const float *x, float *y, long long rows, int cols, void *stream) {
    if (!x || !y) return MSDA_ERR_NULL_POINTER;
    if (rows <= 0 || cols <= 0 || (cols + 31) / 32 > 65535) return MSDA_ERR_BAD_SHAPE;
    return (int)launch_transpose(x, y, rows, cols, (cudaStream_t)stream);
}

int msda_b200_debug_indices_f32(const int64_t *spatial_shapes, const int64_t *level_start,
                                const float *sampling_loc, int batch, int spatial_size,
                                int num_heads, int channels, int num_levels, int num_query,
                                int num_point, int32_t *idx, int64_t *off, void *stream) {
    if (!spatial_shapes || !level_start || !sampling_loc || !idx || !off) return MSDA_ERR_NULL_POINTER;
    const Dims d{batch, spatial_size, num_heads, channels, num_levels, num_query, num_point};
    if (int rc = check_dims(d)) return rc;
    return (int)launch_debug_indices(spatial_shapes, level_start, sampling_loc, d, idx, off,
                                     (cudaStream_t)stream);
}

}  // extern "C"

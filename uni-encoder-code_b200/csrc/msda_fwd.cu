// msda_fwd.cu -- forward MSDA kernel for sm_100a, specialised for 32 fp32
// channels per head (the pixel-decoder shape: d_model 256 / 8 heads).
//
// Replaces the reference's ms_deformable_im2col_gpu_kernel
// (ms_deform_im2col_cuda.cuh:242-304), which runs one thread per output element
// and recomputes every sampling point's geometry in all 32 lanes.
//
// Design:
//   * persistent CTAs walk work items = (image n, head m, group of 8*WARPS queries);
//     when the queries are laid out like the value pixels (encoder self-attention)
//     a group is a 2-D tile of one level, so neighbouring queries' bilinear
//     footprints overlap in L1;
//   * each warp owns 8 consecutive queries of the item.  Phase 1: the 8*L*P
//     sampling points are spread over the lanes (one point per lane per round):
//     load (x, y, weight), run decompose() once per point and write the four
//     corner records {offset, bilinear*attention weight} to shared memory, one
//     plane per corner (conflict-free 8-byte stores);
//   * phase 2, per query: lane = corner*8 + chunk.  One 128-bit shared load gives a
//     lane its corner's records of TWO points; per point it fetches 16 bytes of that
//     corner's 128-byte value row (one LDG.128; a warp instruction covers the four
//     corner rows of a point) and accumulates 4 channels;
//   * the four corner groups are summed with a 3-shuffle reduce-scatter that leaves
//     one output channel per lane: the store is a single coalesced 128-byte line.
//
// The kernel is bound by the L1TEX data pipe: one wavefront per 128-byte value row
// (48 per (query, head)) plus the shared-memory wavefronts of the records
// (profiles/); the record layout exists to keep the latter small.
#include "msda_common.cuh"

namespace msda {

// QPW = queries per warp per item: their records are resident in shared memory together, so
// a smaller QPW leaves more of the SM's 228 KB to L1 (the value rows)
template <int LP, int WARPS, int TILE_W, int QPW>
struct FwdCfg {
    static_assert(LP % 2 == 0, "points are consumed in pairs");
    static constexpr int kQPW = QPW;
    static constexpr int kGroup = WARPS * kQPW;           // queries per item
    static constexpr int kTileH = kGroup / TILE_W;
    static constexpr int kRounds = (kQPW * LP + 31) / 32; // phase-1 rounds
    static constexpr int kRecPerWarp = kQPW * LP;         // records per corner plane
    // plane stride in records: +2 (16 bytes) so the four corner planes start 4 banks apart and the
    // 4-address LDS.128 of phase 2 is conflict-free (an unpadded stride is a multiple of 128 bytes)
    static constexpr int kPlane = kRecPerWarp + 2;
    static constexpr size_t kSmem = (size_t)WARPS * 4 * kPlane * sizeof(uint2);
    static_assert(TILE_W % QPW == 0 && kGroup % TILE_W == 0, "tile shape");
};

template <int LP, int WARPS, int TILE_W, int MIN_CTAS, int QPW, bool FUSED>
__global__ void __launch_bounds__(WARPS * 32, MIN_CTAS)
msda_fwd_d32_kernel(const float *__restrict__ value, const int64_t *__restrict__ shapes,
                    const int64_t *__restrict__ lstart, const float *__restrict__ loc,
                    const float *__restrict__ attw, const Producers pr, const Dims d,
                    const int want_spatial, float *__restrict__ out) {
    using Cfg = FwdCfg<LP, WARPS, TILE_W, QPW>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ LevelTable lt;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int corner = lane >> 3, chunk = lane & 7;
    // corner plane k of this warp: rec + k * kPlane, indexed by (query slot * LP + point)
    uint2 *rec = reinterpret_cast<uint2 *>(smem_raw) + (size_t)warp * 4 * Cfg::kPlane;

    fill_level_table(lt, shapes, lstart, d.L, d.P, d.S, d.Lq, Cfg::kGroup, Cfg::kTileH, TILE_W,
                     want_spatial);
    __syncthreads();

    const int M = d.M;
    const long long items = (long long)d.N * M * lt.groups;
    const uint32_t pix_stride = (uint32_t)M * 8u;          // float4 units per pixel

    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int m = (int)(item % M);
        const long long rest = item / M;
        const int g = (int)(rest % lt.groups);
        const long long n = rest / lt.groups;
        int q0, cnt;
        warp_queries(lt, d.L, g, warp, Cfg::kGroup, Cfg::kTileH, TILE_W, d.Lq, Cfg::kQPW, q0, cnt);

        // ---- phase 1: one sampling point per lane per round -> records ----
        phase1_records<FUSED, LP, Cfg::kQPW, Cfg::kPlane, false>(lt, rec, nullptr, loc, attw, pr, n, q0, cnt, m,
                                                                 M, d.Lq, d.L, pix_stride, (uint32_t)d.S * pix_stride, lane);
        __syncwarp();

        // ---- phase 2: gather + weighted reduction, one query at a time ----
        const float4 *vb = reinterpret_cast<const float4 *>(value) + n * (long long)d.S * M * 8 + chunk;
        const uint4 *plane = reinterpret_cast<const uint4 *>(rec + corner * Cfg::kPlane);
        for (int qi = 0; qi < cnt; ++qi) {
            const uint4 *rq = plane + qi * (LP / 2);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int sp2 = 0; sp2 < LP / 2; ++sp2) {
                const uint4 e = lds_u4(rq + sp2);           // this corner's records of 2 points
                if (e.x != kNoCorner) {
                    const float4 v = ldg_keep_f4(at_off16(vb, e.x));
                    const float w = __uint_as_float(e.y);
                    acc.x = fmaf(w, v.x, acc.x);
                    acc.y = fmaf(w, v.y, acc.y);
                    acc.z = fmaf(w, v.z, acc.z);
                    acc.w = fmaf(w, v.w, acc.w);
                }
                if (e.z != kNoCorner) {
                    const float4 v = ldg_keep_f4(at_off16(vb, e.z));
                    const float w = __uint_as_float(e.w);
                    acc.x = fmaf(w, v.x, acc.x);
                    acc.y = fmaf(w, v.y, acc.y);
                    acc.z = fmaf(w, v.z, acc.z);
                    acc.w = fmaf(w, v.w, acc.w);
                }
            }
            // sum the 4 corner groups (lane bits 3 and 4); lane ends with channel 4*chunk+corner
            const bool up16 = lane & 16, up8 = lane & 8;
            float a0 = up16 ? acc.z : acc.x, a1 = up16 ? acc.w : acc.y;
            const float s0 = up16 ? acc.x : acc.z, s1 = up16 ? acc.y : acc.w;
            a0 += __shfl_xor_sync(kFullMask, s0, 16);
            a1 += __shfl_xor_sync(kFullMask, s1, 16);
            const float keep = up8 ? a1 : a0, send = up8 ? a0 : a1;
            const float res = keep + __shfl_xor_sync(kFullMask, send, 8);
            float *o = out + ((n * d.Lq + q0 + qi) * M + m) * 32LL + chunk * 4 + corner;
            stg_stream_f1(o, res);
        }
        __syncwarp();   // records are overwritten by the next item
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
template <int LP, int WARPS, int TILE_W, int MIN_CTAS, int QPW, bool FUSED = false>
static cudaError_t launch_fwd_cfg(const float *value, const int64_t *shapes, const int64_t *lstart,
                                  const float *loc, const float *attw, const Dims &d,
                                  float *out, cudaStream_t stream, Producers pr = Producers{nullptr, 0, 0, 0, nullptr, nullptr}) {
    using Cfg = FwdCfg<LP, WARPS, TILE_W, QPW>;
    auto kern = msda_fwd_d32_kernel<LP, WARPS, TILE_W, MIN_CTAS, QPW, FUSED>;
    // the opt-in shared-memory size is a per-device function attribute: set it (and query the
    // occupancy) once per device; the values are immutable afterwards
    static std::atomic<int> ctas_per_sm_of[kMaxDevices];
    cudaError_t e = cudaSuccess;
    const int dev = device_slot(&e);
    if (dev < 0) return e;
    int ctas_per_sm = ctas_per_sm_of[dev].load(std::memory_order_acquire);
    if (ctas_per_sm == 0) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
        if (e != cudaSuccess) return e;
        int nb = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, WARPS * 32, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        ctas_per_sm = nb > 0 ? nb : 1;
        ctas_per_sm_of[dev].store(ctas_per_sm, std::memory_order_release);
    }
    int per_sm = ctas_per_sm;
    const int cap = option_value(OPT_CTAS_PER_SM);
    if (cap > 0 && cap < per_sm) per_sm = cap;
    // every group holds at least one query, so items <= N*M*Lq
    long long blocks = (long long)sm_count() * per_sm;
    const long long items_ub = (long long)d.N * d.M * d.Lq;
    if (blocks > items_ub) blocks = items_ub;
    if (blocks < 1) blocks = 1;
    const int want_spatial = option_value(OPT_TILE_ORDER) != 1;
    kern<<<(unsigned)blocks, WARPS * 32, Cfg::kSmem, stream>>>(value, shapes, lstart, loc, attw, pr, d,
                                                              want_spatial, out);
    note_launch();
    return cudaGetLastError();
}

template <int LP>
static cudaError_t launch_fwd_lp(const float *value, const int64_t *shapes, const int64_t *lstart,
                                 const float *loc, const float *attw, const Dims &d, float *out,
                                 cudaStream_t stream) {
    // variant = (warps, query tile width, min CTAs per SM -> register budget, queries per warp)
#define MSDA_FWD(W, TW, C, Q) launch_fwd_cfg<LP, W, TW, C, Q>(value, shapes, lstart, loc, attw, d, out, stream)
    switch (option_value(OPT_FWD_VARIANT)) {
        case 1: return MSDA_FWD(8, 8, 4, 8);     //  8 warps, tile  8x8
        case 3: return MSDA_FWD(32, 16, 1, 8);   // 32 warps, tile 16x16, one CTA per SM
        case 4: return MSDA_FWD(16, 8, 2, 8);    // 16 warps, tile 16x8
        case 5: return MSDA_FWD(16, 16, 2, 4);   // 16 warps, tile  4x16, 4 queries per warp (50 KB smem per SM)
        case 6: return MSDA_FWD(16, 8, 2, 4);    // 16 warps, tile  8x8,  4 queries per warp
        case 7: return MSDA_FWD(32, 16, 1, 4);   // 32 warps, tile  8x16, 4 queries per warp, one CTA per SM
        case 8: return MSDA_FWD(16, 16, 3, 4);   // 16 warps, tile  4x16, 48 warps per SM (40 regs)
        case 2:
        default: return MSDA_FWD(16, 16, 2, 8);  // 16 warps, tile  8x16, 8 queries per warp (98 KB smem per SM)
    }
#undef MSDA_FWD
}

// `handled` tells the caller whether this specialised path took the problem
// (otherwise the generic kernel is used).
cudaError_t launch_fwd_d32(const float *value, const int64_t *shapes, const int64_t *lstart,
                           const float *loc, const float *attw, const Dims &d, float *out,
                           cudaStream_t stream, bool *handled) {
    *handled = true;
    const int LP = d.L * d.P;
    if (d.D != 32 || (long long)d.S * d.M * 8 >= 0x7fffffffLL) {
        *handled = false;
        return cudaSuccess;
    }
    switch (LP) {
        case 4: return launch_fwd_lp<4>(value, shapes, lstart, loc, attw, d, out, stream);
        case 8: return launch_fwd_lp<8>(value, shapes, lstart, loc, attw, d, out, stream);
        case 12: return launch_fwd_lp<12>(value, shapes, lstart, loc, attw, d, out, stream);
        case 16: return launch_fwd_lp<16>(value, shapes, lstart, loc, attw, d, out, stream);
        default: *handled = false; return cudaSuccess;
    }
}

// Fused producers: `off` = raw sampling offsets, `logits` = raw attention logits,
// `ref` = reference points [N or 1, Lq, L, 2].  Only the default CTA shape is instantiated.
cudaError_t launch_fwd_d32_fused(const float *value, const int64_t *shapes, const int64_t *lstart,
                                 const float *ref, long long ref_bstride, const float *off,
                                 const float *logits, const Dims &d, float *out, cudaStream_t stream,
                                 bool *handled, int off_qstride, int logit_qstride, const float *off_table,
                                 const float *logit_table) {
    *handled = true;
    const int lp = d.L * d.P;
    if (off_qstride == d.M * lp * 2 && logit_qstride == d.M * lp) off_qstride = logit_qstride = 0;   // packed
    const Producers pr{ref, ref_bstride, off_qstride, logit_qstride, off_table, logit_table};
    if ((off_qstride != 0) != (logit_qstride != 0) || (off_qstride & 1) || off_qstride < 0 || logit_qstride < 0 ||
        (off_table != nullptr) != (logit_table != nullptr)) {
        *handled = false;
        return cudaSuccess;
    }
    if (d.D != 32 || (long long)d.S * d.M * 8 >= 0x7fffffffLL) {
        *handled = false;
        return cudaSuccess;
    }
    switch (d.L * d.P) {
        case 4: return launch_fwd_cfg<4, 16, 16, 2, 8, true>(value, shapes, lstart, off, logits, d, out, stream, pr);
        case 8: return launch_fwd_cfg<8, 16, 16, 2, 8, true>(value, shapes, lstart, off, logits, d, out, stream, pr);
        case 12: return launch_fwd_cfg<12, 16, 16, 2, 8, true>(value, shapes, lstart, off, logits, d, out, stream, pr);
        case 16: return launch_fwd_cfg<16, 16, 16, 2, 8, true>(value, shapes, lstart, off, logits, d, out, stream, pr);
        default: *handled = false; return cudaSuccess;
    }
}

}  // namespace msda

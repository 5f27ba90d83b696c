// linear_tf32x3.cu -- Y = X * W^T + bias (optionally ReLU) in fp32 on the sm_100a tensor cores.
//
// once the MSDA op is fast, the pixel decoder's time is in its fp32 nn.Linear layers
// (value_proj / sampling_offsets / attention_weights / output_proj / FFN, msdeformattn.py:126-142,
// ops/modules/ms_deform_attn.py:62-65), which the reference runs as fp32 SIMT GEMMs (autocast is
// disabled, msdeformattn.py:338; TF32 matmuls are off by default in torch).  A plain TF32 GEMM would
// change the results (10-bit mantissa), so this kernel uses the error-compensated split
//     x = x_hi + x_lo,   x_hi = tf32(x),   x_lo = tf32(x - x_hi)
//     X * W^T  ~=  Xhi*Whi^T + Xlo*Whi^T + Xhi*Wlo^T          (the dropped Xlo*Wlo^T term is ~2^-22)
// with fp32 accumulation in tensor memory: fp32-class accuracy at tensor-core speed.
//
// The tensor core adds each MMA's result into the fp32 accumulator with truncation (measured: the error
// of a single accumulator grows linearly with the number of MMAs), so the two small products go to a
// SECOND accumulator (2^-11 of the magnitude, hence 2^-11 of the truncation error) that is added in the
// epilogue: one third of the accumulations into the main one.
//
// Structure (one CTA per 128 x BN output tile, 10 warps):
//   pre-pass  W is split once per call into Whi / Wlo (workspace, 2 x N x K floats);
//   warp 0   one thread: TMA loads of the fp32 X tile (128 x 32) and the Whi / Wlo tiles (BN x 32) of
//            each k-block into a STAGES-deep ring (128-byte swizzle: a tile row is one swizzle span);
//   warps 2-9 split the landed X tile element-wise into hi (in place) and lo (second buffer, same
//            offsets -- the swizzle is irrelevant to an element-wise pass), fence to the async proxy
//            and hand the stage to
//   warp 1   one thread: 3 x 4 tcgen05.mma.kind::tf32 (128 x BN x 8) per k-block into the two TMEM
//            accumulators, tcgen05.commit to release the stage;
//   warps 2-9 epilogue: tcgen05.ld both accumulators (lane = row), add them and the bias, ReLU,
//            transpose 32 x 32 blocks through shared memory and store whole 128-byte row segments.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "msda_common.cuh"

namespace msda {

namespace {

constexpr int kBM = 128;          // rows of X per CTA (= TMEM lanes)
constexpr int kBK = 32;           // fp32 elements per k-block: 128 bytes = one swizzle span
constexpr int kUmmaK = 8;         // tf32 MMA depth
constexpr int kSplitWarps = 8;
constexpr int kThreads = 64 + 32 * kSplitWarps;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// shared-memory matrix descriptor: K-major tile whose rows are one 128-byte swizzle span, 8-row
// groups 1024 bytes apart (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 | SBO << 32 |
// version 1 << 46 | SWIZZLE_128B (2) << 61)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, tf32 x tf32, both K-major
__device__ __forceinline__ constexpr uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// round to tf32 (10 explicit mantissa bits), nearest with ties away from zero like cvt.rna.tf32.f32, in
// two full-rate integer instructions (the cvt runs on the quarter-rate conversion pipe)
__device__ __forceinline__ uint32_t to_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

template <int BN, int STAGES>
struct LinCfg {
    static constexpr int kXBytes = kBM * kBK * 4, kWBytes = BN * kBK * 4;
    static constexpr int kStageBytes = 2 * kXBytes + 2 * kWBytes;  // X hi, X lo, W hi, W lo
    static constexpr int kTxBytes = kXBytes + 2 * kWBytes;         // what TMA delivers per stage
    static constexpr int kTileBytes = STAGES * kStageBytes;
    static constexpr int kSmem = kTileBytes + 1024 /* alignment slack */ + 256 /* barriers */;
    static constexpr int kTmemCols = 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
    static_assert(kTileBytes >= kSplitWarps * 32 * 33 * 4, "epilogue scratch lives in the ring");
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
linear_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_wh,
                     const __grid_constant__ CUtensorMap map_wl, const float *__restrict__ bias, float *__restrict__ y, int M, int N, int K, int relu) {
    using Cfg = LinCfg<BN, STAGES>;
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;         // 128-byte swizzle: 1024-byte aligned tiles
    uint8_t *base_ptr = smem_dyn + (base - smem_u32(smem_dyn));
    const uint32_t bars = base + Cfg::kTileBytes;
    // barriers: full[s] (TMA landed), ready[s] (split done), empty[s] (MMAs done), acc (accumulator complete)
    auto full = [&](int s) { return bars + 8u * s; };
    auto ready = [&](int s) { return bars + 8u * (STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    const uint32_t acc_bar = bars + 8u * 3 * STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(base_ptr + Cfg::kTileBytes + 8 * (3 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * BN;
    const int kblocks = K / kBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full(s), 1);
            mbar_init(ready(s), 32 * kSplitWarps);
            mbar_init(empty(s), 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(empty(s), ((kb / STAGES) & 1) ^ 1);
                const uint32_t st = base + s * Cfg::kStageBytes;
                mbar_arrive_expect_tx(full(s), Cfg::kTxBytes);
                tma_load_2d(st, &map_x, full(s), kb * kBK, m0);
                tma_load_2d(st + 2 * Cfg::kXBytes, &map_wh, full(s), kb * kBK, n0);
                tma_load_2d(st + 2 * Cfg::kXBytes + Cfg::kWBytes, &map_wl, full(s), kb * kBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(kBM, BN);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(ready(s), (kb / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t xh = base + s * Cfg::kStageBytes, xl = xh + Cfg::kXBytes;
                const uint32_t wh = xl + Cfg::kXBytes, wl = wh + Cfg::kWBytes;
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k) {
                    const uint32_t ko = k * kUmmaK * 4;                    // 32 bytes inside the swizzle span
                    umma_tf32(tmem_acc + BN, umma_desc(xl + ko), umma_desc(wh + ko), idesc, (kb | k) != 0);
                    umma_tf32(tmem_acc + BN, umma_desc(xh + ko), umma_desc(wl + ko), idesc, 1);
                    umma_tf32(tmem_acc, umma_desc(xh + ko), umma_desc(wh + ko), idesc, (kb | k) != 0);
                }
                umma_commit(empty(s));                                    // stage reusable once these MMAs are done
            }
            umma_commit(acc_bar);
        }
    } else {
        // ---- split warps: x -> (tf32(x), tf32(x - tf32(x))) ----
        const int t = threadIdx.x - 64;                                   // 0..255
        for (int kb = 0; kb < kblocks; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(full(s), (kb / STAGES) & 1);
            float4 *hi = reinterpret_cast<float4 *>(base_ptr + s * Cfg::kStageBytes);
            float4 *lo = reinterpret_cast<float4 *>(base_ptr + s * Cfg::kStageBytes + Cfg::kXBytes);
#pragma unroll
            for (int c = t; c < Cfg::kXBytes / 16; c += 32 * kSplitWarps) {
                const float4 v = hi[c];
                uint4 h, l;
                h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
                l.x = to_tf32(v.x - __uint_as_float(h.x));
                l.y = to_tf32(v.y - __uint_as_float(h.y));
                l.z = to_tf32(v.z - __uint_as_float(h.z));
                l.w = to_tf32(v.w - __uint_as_float(h.w));
                *reinterpret_cast<uint4 *>(hi + c) = h;
                *reinterpret_cast<uint4 *>(lo + c) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor core reads
            mbar_arrive(ready(s));
        }
        // ---- epilogue: TMEM lane = row; warp w may touch lanes 32*(w%4) .. +31 ----
        mbar_wait(acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                                           // TMEM lane quarter of this warp
        const int part = (warp - 2) >> 2;                                 // which of the 32-column blocks: c % 2 == part
        float *tile = reinterpret_cast<float *>(base_ptr) + (warp - 2) * 32 * 33;   // the ring is idle now
        for (int c = part; c < BN / 32; c += kSplitWarps / 4) {
            uint32_t v[32], u[32];
            tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c * 32), u);
            const int col = n0 + c * 32;
            const float b = (bias != nullptr && col + lane < N) ? bias[col + lane] : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
            __syncwarp();
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const int row = m0 + q * 32 + r;
                float o = tile[r * 33 + lane] + b;
                if (relu) o = fmaxf(o, 0.f);
                if (row < M && col + lane < N) y[(long long)row * N + col + lane] = o;
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(Cfg::kTmemCols) : "memory");
}

// W -> (tf32(W), tf32(W - tf32(W))), once per call
__global__ void split_weight_kernel(const float4 *__restrict__ w, float4 *__restrict__ hi, float4 *__restrict__ lo, long long n4) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = w[i];
        uint4 h, l;
        h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
        l.x = to_tf32(v.x - __uint_as_float(h.x));
        l.y = to_tf32(v.y - __uint_as_float(h.y));
        l.z = to_tf32(v.z - __uint_as_float(h.z));
        l.w = to_tf32(v.w - __uint_as_float(h.w));
        reinterpret_cast<uint4 *>(hi)[i] = h;
        reinterpret_cast<uint4 *>(lo)[i] = l;
    }
}

// ---- host side ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess &&
            st == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp32 matrix [rows x cols], box = box_rows x 32 columns, 128-byte swizzle, zero fill outside
bool make_map(CUtensorMap *map, const float *ptr, int rows, int cols, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int STAGES>
cudaError_t launch_linear(const float *x, const float *w, const float *bias, float *y, int M, int N, int K,
                          int relu, float *workspace, cudaStream_t stream) {
    using Cfg = LinCfg<BN, STAGES>;
    auto kern = linear_tf32x3_kernel<BN, STAGES>;
    static bool attr_set[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    float *whi = workspace, *wlo = workspace + (size_t)N * K;
    CUtensorMap mx, mwh, mwl;
    if (!make_map(&mx, x, M, K, kBM) || !make_map(&mwh, whi, N, K, BN) || !make_map(&mwl, wlo, N, K, BN))
        return cudaErrorInvalidValue;
    const long long n4 = (long long)N * K / 4;
    split_weight_kernel<<<(unsigned)((n4 + 255) / 256 < 592 ? (n4 + 255) / 256 : 592), 256, 0, stream>>>(
        reinterpret_cast<const float4 *>(w), reinterpret_cast<float4 *>(whi), reinterpret_cast<float4 *>(wlo), n4);
    note_launch();
    dim3 grid((M + kBM - 1) / kBM, (N + BN - 1) / BN);
    kern<<<grid, kThreads, Cfg::kSmem, stream>>>(mx, mwh, mwl, bias, y, M, N, K, relu);
    note_launch();
    return cudaGetLastError();
}

}  // namespace

// y[M, N] = x[M, K] * w[N, K]^T + bias[N]; K % 32 == 0, 16-byte aligned rows; workspace: 2*N*K floats
cudaError_t launch_linear_tf32x3(const float *x, const float *w, const float *bias, float *y, int M, int N, int K,
                                 int relu, float *workspace, cudaStream_t stream, bool *handled) {
    *handled = true;
    if (M <= 0 || N <= 0 || K <= 0 || K % kBK != 0 || N % 4 != 0 ||
        (reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(workspace)) % 16 != 0) {
        *handled = false;
        return cudaSuccess;
    }
    if (N % 256 == 0 || N > 192) return launch_linear<256, 2>(x, w, bias, y, M, N, K, relu, workspace, stream);
    if (N > 128) return launch_linear<192, 2>(x, w, bias, y, M, N, K, relu, workspace, stream);
    if (N > 96) return launch_linear<128, 3>(x, w, bias, y, M, N, K, relu, workspace, stream);
    return launch_linear<96, 4>(x, w, bias, y, M, N, K, relu, workspace, stream);
}

}  // namespace msda

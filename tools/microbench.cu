// microbench.cu -- per-SM cost (cycles per warp instruction) of the memory-pipe operations
// the MSDA kernels are built from, measured on the B200 they run on.  Results are
// summarised in profiles/ and drive the kernel design in DESIGN.md.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
//   ./tools/microbench > gpurun_out/microbench.txt
//
// Every test runs 148*2 CTAs of 512 threads; each warp executes ITERS x UNROLL copies of the
// operation; the figure of merit is SM cycles per warp-level instruction with all 32 warps of
// an SM competing (i.e. reciprocal throughput of the shared LSU / L1TEX / XBAR path).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int THREADS = 512;
constexpr int ITERS = 256;
constexpr int UNROLL = 8;
constexpr int ROW_BYTES = 128;
constexpr int NS_SLOT = 148 * 16;   // cycles[NS_SLOT..+1]: wall nanoseconds and cycles of CTA 0 (actual SM clock)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

enum Test { LDS32_BCAST, LDS32_DISTINCT, LDS64_4ADDR, LDS64_DISTINCT, LDS128_4ADDR, LDS128_BCAST,
            LDS128_DISTINCT, LDS128_2SAMPLES, SHFL, STS32_DISTINCT, LDS_STS_RMW,
            LDG128_4ROWS_L1, LDG32_1ROW_L1, LDG64_2ROWS_L1, LDG128_4ROWS_L2, LDG32_1ROW_L2,
            RED128_4ROWS, RED32_1ROW, RED64_2ROWS, TMA_RED_ROW, STG128_4ROWS,
            SHFL2, ATOMS_INT_SPREAD, ATOMS_F32_ROW, ATOMS_F32_ROW_SMALLWIN, ATOMS_INT_RET, MIX_RED_TMA, NUM_TESTS };
const char *kNames[] = {"lds32 broadcast (1 addr)", "lds32 32 distinct banks", "lds64 4 addrs (corner groups)",
    "lds64 32 distinct", "lds128 4 addrs (corner groups)", "lds128 broadcast (1 addr)", "lds128 32 distinct",
    "lds128 4 addrs x 16B contiguous 64B", "shfl.bfly", "sts32 32 distinct banks", "lds32+fadd+sts32 row RMW",
    "ldg128 4 rows/instr, L1-resident", "ldg32 1 row/instr, L1-resident", "ldg64 2 rows/instr, L1-resident",
    "ldg128 4 rows/instr, L2-resident random", "ldg32 1 row/instr, L2-resident random",
    "red.v4.f32 4 rows/instr, random rows", "red.f32 1 row/instr, random rows", "red.v2.f32 2 rows/instr, random rows",
    "TMA cp.reduce.async.bulk 128B row (sts row + 1 bulk op)", "stg128 4 rows/instr, random rows",
    "shfl.bfly (dependent chain, 32 warps)", "smem red.add.s32, 32 random words of 8K", "smem atomicAdd(float) one row of 256 (32 lanes = 32 banks)",
    "smem atomicAdd(float) one row of 16 (cross-warp conflicts)", "smem atomicAdd(int) with return, 32 random words of 8K",
    "MIX even warps red.v4.f32 (4 rows), odd warps TMA bulk reduce (1 row)"};

template <int TEST>
__global__ void __launch_bounds__(THREADS) bench(float *gbuf, uint32_t nrows_mask, float *sink, long long *cycles,
                                                  uint32_t active_sms) {
    __shared__ __align__(128) float sm[8192];   // 32 KB
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // CTAs that land on an SM beyond `active_sms` leave at once (fewer SMs against the same L2)
        uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (smid >= active_sms) {
            if (threadIdx.x < 2) cycles[blockIdx.x * 2 + threadIdx.x] = 0;
            return;
        }
    }
    for (int i = threadIdx.x; i < 8192; i += THREADS) sm[i] = (float)i;
    __syncthreads();
    float acc = 0.f;
    // address generation is kept to ~2 integer instructions per operation so that it does not
    // hide a 1-cycle-per-wavefront pipe: per-lane base + (iteration * stride) & mask
    uint32_t rows[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) rows[u] = hash32(blockIdx.x * 977u + warp * 131u + u * 17u + 7u);
    uint32_t lane_base = 0;   // float index into sm
    if (TEST == LDS32_DISTINCT || TEST == STS32_DISTINCT || TEST == LDS_STS_RMW) lane_base = lane;
    if (TEST == LDS64_4ADDR) lane_base = (lane >> 3) * 2;
    if (TEST == LDS64_DISTINCT) lane_base = lane * 2;
    if (TEST == LDS128_4ADDR) lane_base = (lane >> 3) * 8;        // {off,w} of a 32-byte record, stride 8 floats? (4 x 16B slots 32B apart)
    if (TEST == LDS128_DISTINCT) lane_base = lane * 4;
    if (TEST == LDS128_2SAMPLES) lane_base = (lane >> 3) * 4;     // 4 contiguous 16-byte slots
    const uint32_t sbase = smem_u32(sm + lane_base);
    unsigned long long ns0, ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t k = (uint32_t)(it * UNROLL + u);
            if (TEST == LDS32_BCAST || TEST == LDS32_DISTINCT) {
                float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbase + ((k * 128u) & 32767u))); acc += v;
            } else if (TEST == LDS64_4ADDR || TEST == LDS64_DISTINCT) {
                float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sbase + ((k * 256u) & 32767u))); acc += v.x + v.y;
            } else if (TEST == LDS128_4ADDR || TEST == LDS128_BCAST || TEST == LDS128_DISTINCT || TEST == LDS128_2SAMPLES) {
                float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sbase + ((k * 512u) & 32767u))); acc += v.x + v.w;
            } else if (TEST == SHFL) {
                acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            } else if (TEST == SHFL2) {
                acc = __shfl_xor_sync(0xffffffffu, acc + (float)lane, 8) * 1.0001f;
            } else if (TEST == ATOMS_INT_SPREAD) {
                const uint32_t w = (rows[u] + k * 2654435761u + lane * 40503u) & 8191u;
                asm volatile("red.shared.add.s32 [%0], %1;" :: "r"(smem_u32(sm + w)), "r"(1) : "memory");
            } else if (TEST == ATOMS_INT_RET) {
                const uint32_t w = (rows[u] + k * 2654435761u + lane * 40503u) & 8191u;
                int old; asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(sm + w)), "r"(1) : "memory");
                acc += (float)old;
            } else if (TEST == ATOMS_F32_ROW) {
                const uint32_t row = (rows[u] + k * 2654435761u) >> 24;           // 0..255
                atomicAdd(sm + row * 32 + lane, 1.0f);
            } else if (TEST == ATOMS_F32_ROW_SMALLWIN) {
                const uint32_t row = (rows[u] + k * 2654435761u) >> 28;           // 0..15
                atomicAdd(sm + row * 32 + lane, 1.0f);
            } else if (TEST == STS32_DISTINCT) {
                asm volatile("st.shared.f32 [%0], %1;" :: "r"(sbase + ((k * 128u) & 32767u)), "f"(acc) : "memory");
            } else if (TEST == LDS_STS_RMW) {
                const uint32_t a = sbase + ((k * 128u) & 32767u);
                float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
                asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v + 1.0f) : "memory");
            } else {
                // global-memory tests: a row index per (warp, u), advanced every iteration
                const bool l1 = (TEST == LDG128_4ROWS_L1 || TEST == LDG32_1ROW_L1 || TEST == LDG64_2ROWS_L1);
                const uint32_t step = rows[u] + (uint32_t)it * 40503u;
                if (TEST == LDG128_4ROWS_L1 || TEST == LDG128_4ROWS_L2 || TEST == RED128_4ROWS || TEST == STG128_4ROWS) {
                    uint32_t row = step + (uint32_t)(lane >> 3) * 7919u;
                    row = l1 ? (row & 255u) + blockIdx.x * 256u : (row & nrows_mask);
                    float *p = gbuf + (size_t)row * 32 + (lane & 7) * 4;
                    if (TEST == RED128_4ROWS) {
                        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
                    } else if (TEST == STG128_4ROWS) {
                        asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
                    } else {
                        float4 v; asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); acc += v.x + v.w;
                    }
                } else if (TEST == LDG64_2ROWS_L1 || TEST == RED64_2ROWS) {
                    uint32_t row = step + (uint32_t)(lane >> 4) * 7919u;
                    row = l1 ? (row & 255u) + blockIdx.x * 256u : (row & nrows_mask);
                    float *p = gbuf + (size_t)row * 32 + (lane & 15) * 2;
                    if (TEST == RED64_2ROWS) {
                        asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
                    } else {
                        float2 v; asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p)); acc += v.x + v.y;
                    }
                } else if (TEST == LDG32_1ROW_L1 || TEST == LDG32_1ROW_L2 || TEST == RED32_1ROW) {
                    uint32_t row = l1 ? (step & 255u) + blockIdx.x * 256u : (step & nrows_mask);
                    float *p = gbuf + (size_t)row * 32 + lane;
                    if (TEST == RED32_1ROW) {
                        asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(1.0f) : "memory");
                    } else {
                        float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); acc += v;
                    }
                } else if (TEST == MIX_RED_TMA && !(warp & 1)) {
                    uint32_t row = (step + (uint32_t)(lane >> 3) * 7919u) & nrows_mask;
                    float *p = gbuf + (size_t)row * 32 + (lane & 7) * 4;
                    asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
                } else if (TEST == TMA_RED_ROW || TEST == MIX_RED_TMA) {
                    // each warp owns UNROLL 128-byte staging rows in shared memory
                    float *srow = sm + (warp * UNROLL + u) * 32;
                    asm volatile("st.shared.f32 [%0], %1;" :: "r"(smem_u32(srow + lane)), "f"(1.0f) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        float *p = gbuf + (size_t)(step & nrows_mask) * 32;
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                     :: "l"(p), "r"(smem_u32(srow)), "n"(ROW_BYTES) : "memory");
                    }
                }
            }
        }
        if (TEST == TMA_RED_ROW || (TEST == MIX_RED_TMA && (warp & 1))) {
            if (lane == 0) {
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
        }
    }
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    if (acc == 123.456f) sink[0] = acc;
    // the CTA's time is that of its LAST warp (the LSU serves older warps first: warp 0 alone finishes
    // its loop well before the CTA's reductions are all issued); even / odd warps separately for MIX
    __shared__ unsigned long long last[2];
    if (threadIdx.x < 2) last[threadIdx.x] = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&last[TEST == MIX_RED_TMA ? (warp & 1) : 0], (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x < 2) cycles[blockIdx.x * 2 + threadIdx.x] = (long long)last[threadIdx.x];
    if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[NS_SLOT] = (long long)(ns1 - ns0); cycles[NS_SLOT + 1] = t1 - t0; }
}


// Do LSU reductions (red.global.add.v4.f32) and TMA bulk reductions (cp.reduce.async.bulk) leave the
// SM through one path or two?  Every warp issues, per iteration, R4 four-row vector reductions and T
// one-row bulk reductions from a double-buffered staging ring (the previous iteration's bulk group
// may still be reading while this one is filled), so both engines are kept busy together.
template <int R4, int T>
__global__ void __launch_bounds__(THREADS) mix2(float *gbuf, uint32_t nrows_mask, long long *cycles) {
    constexpr int TT = T > 0 ? T : 1;
    __shared__ __align__(128) float stage[THREADS / 32][2][TT][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t h = hash32(blockIdx.x * 977u + warp * 131u + 7u);
    __syncthreads();
    unsigned long long ns0, ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
    const long long t0 = clock64();
    for (int it = 0; it < ITERS * 2; ++it) {
#pragma unroll
        for (int u = 0; u < R4; ++u) {
            h = h * 1664525u + 1013904223u;
            const uint32_t row = ((h >> 8) + (uint32_t)(lane >> 3) * 7919u) & nrows_mask;
            float *p = gbuf + (size_t)row * 32 + (lane & 7) * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" :: "l"(p), "f"(1.0f) : "memory");
        }
        if (T > 0) {
            const int b = it & 1;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // buffer b is free again
            __syncwarp();
#pragma unroll
            for (int u = 0; u < TT; ++u)
                asm volatile("st.shared.f32 [%0], %1;" :: "r"(smem_u32(&stage[warp][b][u][lane])), "f"(1.0f) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane < TT) {
                h = h * 1664525u + 1013904223u;
                const uint32_t row = ((h >> 8) + (uint32_t)lane * 104729u) & nrows_mask;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             :: "l"(gbuf + (size_t)row * 32), "r"(smem_u32(&stage[warp][b][lane][0])), "n"(ROW_BYTES) : "memory");
            }
            if (lane == 0) asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (T > 0) { if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); __syncwarp(); }
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    __shared__ unsigned long long last;
    if (threadIdx.x == 0) last = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&last, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (long long)last;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cycles[NS_SLOT] = (long long)(ns1 - ns0); cycles[NS_SLOT + 1] = t1 - t0; }
}

template <int R4, int T>
void run_mix2(float *gbuf, uint32_t nrows_mask, long long *cycles_d, int sms) {
    const int grid = sms * 2;
    const int pad = 98 * 1024 - (THREADS / 32) * 2 * (T > 0 ? T : 1) * 128;   // exactly 2 CTAs per SM
    CHECK(cudaFuncSetAttribute(mix2<R4, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
    for (int w = 0; w < 20; ++w) mix2<R4, T><<<grid, THREADS, pad>>>(gbuf, nrows_mask, cycles_d);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    CHECK(cudaEventRecord(e0));
    mix2<R4, T><<<grid, THREADS, pad>>>(gbuf, nrows_mask, cycles_d);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    long long *h = (long long *)malloc(sizeof(long long) * grid);
    CHECK(cudaMemcpy(h, cycles_d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
    free(h);
    const double rows_per_sm = 2.0 * (THREADS / 32) * (ITERS * 2) * (4.0 * R4 + T);
    long long nsclk[2];
    CHECK(cudaMemcpy(nsclk, cycles_d + NS_SLOT, sizeof(nsclk), cudaMemcpyDeviceToHost));
    printf("[%4.0f MHz] ", nsclk[0] > 0 ? 1e3 * (double)nsclk[1] / (double)nsclk[0] : 0.0);
    printf("mix2: per warp-iteration %d x red.v4 (4 rows) + %d x TMA bulk reduce (1 row): %8.3f ms  %9.0f clk/CTA  %5.2f SM-cycles per row  (%6.1f GB/s)\n",
           R4, T, ms, mean, mean / rows_per_sm, rows_per_sm * sms * 128.0 / (ms * 1e-3) / 1e9);
}

template <int TEST>
void run(float *gbuf, uint32_t nrows_mask, float *sink, long long *cycles_d, int sms, int ctas_per_sm = 2,
         int active_sms = 1 << 20) {
    const int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    // dynamic shared memory that nothing uses: exactly `ctas_per_sm` CTAs fit on an SM, so the grid is
    // spread evenly (without it up to 4 CTAs fit and the block scheduler loads the SMs unevenly)
    const int pad = (200 * 1024) / ctas_per_sm - 33 * 1024;
    CHECK(cudaFuncSetAttribute(bench<TEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad));
    for (int w = 0; w < 20; ++w)   // warm-up: the clocks of an idle GPU take a while to reach the boost state
        bench<TEST><<<grid, THREADS, pad>>>(gbuf, nrows_mask, sink, cycles_d, (uint32_t)active_sms);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    bench<TEST><<<grid, THREADS, pad>>>(gbuf, nrows_mask, sink, cycles_d, (uint32_t)active_sms);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0; CHECK(cudaEventElapsedTime(&ms, e0, e1));
    long long *h = (long long *)malloc(sizeof(long long) * grid * 2);
    CHECK(cudaMemcpy(h, cycles_d, sizeof(long long) * grid * 2, cudaMemcpyDeviceToHost));
    double mean = 0, mean1 = 0; int live = 0;
    for (int i = 0; i < grid; ++i) if (h[2 * i]) { mean += (double)h[2 * i]; mean1 += (double)h[2 * i + 1]; ++live; }
    mean /= live; mean1 /= live;
    double cmin = 1e30, cmax = 0;
    for (int i = 0; i < grid; ++i) if (h[2 * i]) { cmin = h[2 * i] < cmin ? h[2 * i] : cmin; cmax = h[2 * i] > cmax ? h[2 * i] : cmax; }
    if (TEST == RED128_4ROWS || TEST == STG128_4ROWS || TEST == LDG128_4ROWS_L2)
        printf("[CTA clk min %.0f max %.0f] ", cmin, cmax);
    if (TEST == RED128_4ROWS && ctas_per_sm == 2 && live == grid && nrows_mask == (64u << 20) / ROW_BYTES - 1) {
        // distribution of CTA times in 10 buckets between min and max
        int hist[10] = {0};
        for (int i = 0; i < grid; ++i) { int b = (int)((h[2 * i] - cmin) / (cmax - cmin + 1) * 10); hist[b]++; }
        printf("\n    CTA-time histogram (min..max, 10 buckets):");
        for (int b = 0; b < 10; ++b) printf(" %d", hist[b]);
        printf("\n    ");
    }
    free(h);
    long long nsclk[2];
    CHECK(cudaMemcpy(nsclk, cycles_d + NS_SLOT, sizeof(nsclk), cudaMemcpyDeviceToHost));
    printf("[%4.0f MHz] ", nsclk[0] > 0 ? 1e3 * (double)nsclk[1] / (double)nsclk[0] : 0.0);
    if (live != grid) printf("[%3d of %d CTAs live, ~%d SMs] ", live, grid, live / ctas_per_sm);
    if (nrows_mask != (64u << 20) / ROW_BYTES - 1) printf("[%u KB of rows] ", (nrows_mask + 1) / 8);
    if (TEST == MIX_RED_TMA) {
        const double n = (double)ctas_per_sm * (THREADS / 64) * ITERS * UNROLL;   // ops per group per SM
        printf("%s: red warps %.0f clk (%.2f cyc per 4-row instr alone-equivalent), tma warps %.0f clk (%.2f cyc per row op); %.3f ms\n",
               kNames[TEST], mean, mean / n, mean1, mean1 / n, ms);
        return;
    }
    // warp instructions per SM = ctas_per_sm * warps * ITERS * UNROLL
    const double instr_per_sm = (double)ctas_per_sm * (THREADS / 32) * ITERS * UNROLL;
    if (ctas_per_sm != 2) printf("[%d CTAs = %2d warps/SM] ", ctas_per_sm, ctas_per_sm * THREADS / 32);
    printf("%-58s  %8.3f ms  %9.0f clk/CTA  %7.2f SM-cycles per warp-instr  (%6.1f GB/s row payload @128B/row-instr)\n",
           kNames[TEST], ms, mean, mean / instr_per_sm,
           instr_per_sm * sms * 128.0 / (ms * 1e-3) / 1e9);
}

int main() {
    int dev = 0, sms = 0, clk = 0;
    CHECK(cudaGetDevice(&dev));
    CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CHECK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    printf("SMs %d, max clock %d kHz; %d threads/CTA, 2 CTAs/SM, %d x %d ops per warp\n", sms, clk, THREADS, ITERS, UNROLL);
    const uint32_t nrows = (64u << 20) / ROW_BYTES;   // 64 MB of 128-byte rows (L2-resident, like grad_value at configs[1])
    float *gbuf, *sink; long long *cycles;
    CHECK(cudaMalloc(&gbuf, (size_t)nrows * ROW_BYTES));
    CHECK(cudaMemset(gbuf, 0, (size_t)nrows * ROW_BYTES));
    CHECK(cudaMalloc(&sink, 16));
    CHECK(cudaMalloc(&cycles, sizeof(long long) * (NS_SLOT + 2)));
    for (int w = 0; w < 2000; ++w) bench<LDG128_4ROWS_L2><<<sms * 2, THREADS>>>(gbuf, nrows - 1, sink, cycles, 1u << 20);   // ~0.6 s
    CHECK(cudaDeviceSynchronize());
    run<LDS32_BCAST>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS32_DISTINCT>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS64_4ADDR>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS64_DISTINCT>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS128_4ADDR>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS128_BCAST>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS128_DISTINCT>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS128_2SAMPLES>(gbuf, nrows - 1, sink, cycles, sms);
    run<SHFL>(gbuf, nrows - 1, sink, cycles, sms);
    run<STS32_DISTINCT>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDS_STS_RMW>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDG128_4ROWS_L1>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDG32_1ROW_L1>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDG64_2ROWS_L1>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDG128_4ROWS_L2>(gbuf, nrows - 1, sink, cycles, sms);
    run<LDG32_1ROW_L2>(gbuf, nrows - 1, sink, cycles, sms);
    run<RED128_4ROWS>(gbuf, nrows - 1, sink, cycles, sms);
    run<RED32_1ROW>(gbuf, nrows - 1, sink, cycles, sms);
    run<RED64_2ROWS>(gbuf, nrows - 1, sink, cycles, sms);
    run<TMA_RED_ROW>(gbuf, nrows - 1, sink, cycles, sms);
    run<STG128_4ROWS>(gbuf, nrows - 1, sink, cycles, sms);
    run<SHFL2>(gbuf, nrows - 1, sink, cycles, sms);
    run<ATOMS_INT_SPREAD>(gbuf, nrows - 1, sink, cycles, sms);
    run<ATOMS_INT_RET>(gbuf, nrows - 1, sink, cycles, sms);
    run<ATOMS_F32_ROW>(gbuf, nrows - 1, sink, cycles, sms);
    run<ATOMS_F32_ROW_SMALLWIN>(gbuf, nrows - 1, sink, cycles, sms);
    // does the reduction / load rate depend on how many warps are resident? (1..4 CTAs of 16 warps)
    for (int c = 1; c <= 4; ++c) run<RED128_4ROWS>(gbuf, nrows - 1, sink, cycles, sms, c);
    for (int c = 1; c <= 4; ++c) run<RED32_1ROW>(gbuf, nrows - 1, sink, cycles, sms, c);
    for (int c = 1; c <= 4; ++c) run<LDG128_4ROWS_L1>(gbuf, nrows - 1, sink, cycles, sms, c);
    for (int c = 1; c <= 4; ++c) run<LDG128_4ROWS_L2>(gbuf, nrows - 1, sink, cycles, sms, c);
    for (int c = 1; c <= 4; ++c) run<STG128_4ROWS>(gbuf, nrows - 1, sink, cycles, sms, c);
    // is the reduction rate an SM-side or an L2-side limit?  same kernel on a quarter / half / 3/4 of the SMs
    for (int a : {37, 74, 111, 148}) run<RED128_4ROWS>(gbuf, nrows - 1, sink, cycles, sms, 2, a);
    for (int a : {37, 74, 111, 148}) run<STG128_4ROWS>(gbuf, nrows - 1, sink, cycles, sms, 2, a);
    for (int a : {37, 74, 148}) run<LDG128_4ROWS_L2>(gbuf, nrows - 1, sink, cycles, sms, 2, a);
    // does it depend on how many distinct rows are hit (same-row collisions in the L2 atomic units)?
    for (uint32_t kb : {65536u, 4096u, 256u, 16u}) run<RED128_4ROWS>(gbuf, kb * 8 - 1, sink, cycles, sms);
    // do LSU reductions and TMA bulk reductions share one path?
    run<MIX_RED_TMA>(gbuf, nrows - 1, sink, cycles, sms);
    run<MIX_RED_TMA>(gbuf, nrows - 1, sink, cycles, sms, 2, 74);
    run_mix2<2, 0>(gbuf, nrows - 1, cycles, sms);
    run_mix2<0, 8>(gbuf, nrows - 1, cycles, sms);
    run_mix2<2, 2>(gbuf, nrows - 1, cycles, sms);
    run_mix2<2, 4>(gbuf, nrows - 1, cycles, sms);
    run_mix2<2, 8>(gbuf, nrows - 1, cycles, sms);
    run_mix2<1, 8>(gbuf, nrows - 1, cycles, sms);
    printf("done\n");
    return 0;
}

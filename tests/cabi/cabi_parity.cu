// cabi_parity.cu -- a torch-free consumer of the C ABI (include/msda_b200.h): plain cudaMalloc'd
// buffers in, plain buffers out, checked against the C oracle (oracle/msda_oracle.c, linked as
// test infrastructure).  Built and run by tests/test_cabi_gpu.py on the GPU box:
//   nvcc -o cabi_parity cabi_parity.cu -I include -L <lib> -lmsda_b200 -L oracle/_build -lmsda_oracle
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "msda_b200.h"

extern "C" {
void msda_oracle_forward_f64g32(const double *, const int64_t *, const int64_t *, const double *,
                                const double *, int, int, int, int, int, int, int, double *);
void msda_oracle_backward_f64g32(const double *, const double *, const int64_t *, const int64_t *,
                                 const double *, const double *, int, int, int, int, int, int, int,
                                 double *, double *, double *);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (rng_state >> 11) * (1.0 / 9007199254740992.0); }

template <typename T> static T *to_dev(const std::vector<T> &h) {
    T *d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

int main() {
    const int N = 2, M = 8, D = 32, L = 3, P = 4;
    const int64_t shapes_h[6] = {5, 9, 10, 18, 20, 36};
    int64_t lsi_h[3]; int S = 0;
    for (int l = 0; l < L; ++l) { lsi_h[l] = S; S += (int)(shapes_h[2 * l] * shapes_h[2 * l + 1]); }
    const int Lq = S;
    std::vector<float> value((size_t)N * S * M * D), loc((size_t)N * Lq * M * L * P * 2), w((size_t)N * Lq * M * L * P), go((size_t)N * Lq * M * D);
    for (auto &v : value) v = (float)(urand() * 2 - 1);
    for (auto &v : go) v = (float)(urand() * 2 - 1);
    for (auto &v : loc) v = (float)(urand() * 1.2 - 0.1);          // ~25 % of the points out of range
    for (size_t r = 0; r < w.size(); r += L * P) {                  // normalised weights per (query, head)
        double s = 0; for (int k = 0; k < L * P; ++k) { w[r + k] = (float)(urand() + 0.05); s += w[r + k]; }
        for (int k = 0; k < L * P; ++k) w[r + k] = (float)(w[r + k] / s);
    }
    std::vector<int64_t> shapes_v(shapes_h, shapes_h + 6), lsi_v(lsi_h, lsi_h + 3);
    float *d_value = to_dev(value), *d_loc = to_dev(loc), *d_w = to_dev(w), *d_go = to_dev(go);
    int64_t *d_shapes = to_dev(shapes_v), *d_lsi = to_dev(lsi_v);
    float *d_out, *d_gv, *d_gl, *d_gw;
    CK(cudaMalloc(&d_out, go.size() * 4)); CK(cudaMalloc(&d_gv, value.size() * 4));
    CK(cudaMalloc(&d_gl, loc.size() * 4)); CK(cudaMalloc(&d_gw, w.size() * 4));
    CK(cudaMemset(d_gv, 0, value.size() * 4));                      // contract: grad_value zero on entry
    cudaStream_t st; CK(cudaStreamCreate(&st));

    if (msda_b200_abi_version() != MSDA_B200_ABI_VERSION) { printf("ABI mismatch\n"); return 1; }
    if (msda_b200_forward_f32(nullptr, d_shapes, d_lsi, d_loc, d_w, N, S, M, D, L, Lq, P, d_out, st) != MSDA_ERR_NULL_POINTER) { printf("NULL not rejected\n"); return 1; }
    const long long l0 = msda_b200_launch_count();
    int rc = msda_b200_forward_f32(d_value, d_shapes, d_lsi, d_loc, d_w, N, S, M, D, L, Lq, P, d_out, st);
    if (rc) { printf("forward: %s\n", msda_b200_error_string(rc)); return 1; }
    rc = msda_b200_backward_f32(d_go, d_value, d_shapes, d_lsi, d_loc, d_w, N, S, M, D, L, Lq, P, d_gv, d_gl, d_gw, st);
    if (rc) { printf("backward: %s\n", msda_b200_error_string(rc)); return 1; }
    CK(cudaStreamSynchronize(st));
    // forward + backward; where the queries are the value pixels the backward launches the merging and
    // the per-row kernel, one of which returns at once (include/msda_b200.h)
    const long long nl = msda_b200_launch_count() - l0;
    if (nl != 2 && nl != 3) { printf("launch count\n"); return 1; }

    std::vector<float> out(go.size()), gv(value.size()), gl(loc.size()), gw(w.size());
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gv.data(), d_gv, gv.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gl.data(), d_gl, gl.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gw.data(), d_gw, gw.size() * 4, cudaMemcpyDeviceToHost));

    std::vector<double> v64(value.begin(), value.end()), l64(loc.begin(), loc.end()), w64(w.begin(), w.end()), g64(go.begin(), go.end());
    std::vector<double> r_out(out.size()), r_gv(gv.size(), 0.0), r_gl(gl.size()), r_gw(gw.size());
    msda_oracle_forward_f64g32(v64.data(), shapes_h, lsi_h, l64.data(), w64.data(), N, S, M, D, L, Lq, P, r_out.data());
    msda_oracle_backward_f64g32(g64.data(), v64.data(), shapes_h, lsi_h, l64.data(), w64.data(), N, S, M, D, L, Lq, P, r_gv.data(), r_gl.data(), r_gw.data());
    auto cmp = [](const std::vector<float> &a, const std::vector<double> &b, bool relative) {
        double num = 0, den = 0;
        for (size_t i = 0; i < a.size(); ++i) { num = fmax(num, fabs(a[i] - b[i])); den = fmax(den, fabs(b[i])); }
        return relative ? num / den : num;
    };
    const double e_out = cmp(out, r_out, false), e_gv = cmp(gv, r_gv, true), e_gl = cmp(gl, r_gl, true), e_gw = cmp(gw, r_gw, true);
    printf("forward max abs err %.3e; grad_value %.3e grad_loc %.3e grad_weight %.3e (relative)\n", e_out, e_gv, e_gl, e_gw);
    if (!(e_out <= 1e-5 && e_gv <= 1e-4 && e_gl <= 1e-4 && e_gw <= 1e-4)) { printf("FAIL\n"); return 1; }
    printf("CABI PARITY OK\n");
    return 0;
}

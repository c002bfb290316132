// Micro-benchmark: HBM bandwidth when a row-major [n, N] f32 matrix is read as column panels.
// Each warp reads `WB` contiguous bytes of every row of its panel (rows 4*N bytes apart), U rows in flight.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <random>

template <int VEC, int U>
__global__ void probe(const float* __restrict__ X, long long ld, int n_rows, const int* __restrict__ perm, int rows_per_cta,
                      int warps_per_row_strip, float* sink) {
    // blockIdx.x: panel of (warps per CTA * 32 * VEC) floats; blockIdx.y: row chunk
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long col = ((long long)blockIdx.x * (blockDim.x >> 5) + w) * 32 * VEC + lane * VEC;
    if (col + VEC > ld) return;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(n_rows, r0 + rows_per_cta);
    float acc = 0.f;
    for (int r = r0; r < r1; r += U) {
        float v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int rr = min(r + u, r1 - 1);
            int row = perm ? perm[rr] : rr;
            const float* p = X + (long long)row * ld + col;
            if (VEC == 4) { float4 t = __ldcs((const float4*)p); v[u][0] = t.x; v[u][1 % VEC] = t.y; v[u][2 % VEC] = t.z; v[u][3 % VEC] = t.w; }
            else if (VEC == 2) { float2 t = __ldcs((const float2*)p); v[u][0] = t.x; v[u][1 % VEC] = t.y; }
            else v[u][0] = __ldcs(p);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc += v[u][k];
    }
    if (acc == 123.456f) *sink = acc;
}

template <int VEC, int U>
void run(const char* name, const float* X, long long ld, int n, const int* perm, int warps, int rows_per_cta, float* sink) {
    dim3 grid((unsigned)((ld + warps * 32 * VEC - 1) / (warps * 32 * VEC)), (n + rows_per_cta - 1) / rows_per_cta);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) probe<VEC, U><<<grid, warps * 32>>>(X, ld, n, perm, rows_per_cta, 0, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 3; ++i) probe<VEC, U><<<grid, warps * 32>>>(X, ld, n, perm, rows_per_cta, 0, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
    printf("%-44s warps/CTA %d rows/CTA %4d : %7.3f ms  %7.1f GB/s  (%s)\n", name, warps, rows_per_cta, ms,
           (double)n * ld * 4 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int n = 300000; const long long ld = 8000;
    float* X; cudaMalloc(&X, (size_t)n * ld * 4); cudaMemset(X, 0, (size_t)n * ld * 4);
    float* sink; cudaMalloc(&sink, 4);
    std::vector<int> h(n); for (int i = 0; i < n; ++i) h[i] = i;
    std::mt19937 g(1); std::shuffle(h.begin(), h.end(), g);
    int* perm; cudaMalloc(&perm, n * 4); cudaMemcpy(perm, h.data(), n * 4, cudaMemcpyHostToDevice);
    for (int pass = 0; pass < 2; ++pass) {
        const int* P = pass ? perm : nullptr;
        printf("---- rows %s\n", pass ? "RANDOM (permuted)" : "sequential");
        run<1, 8>("128 B per warp-row, 8 in flight", X, ld, n, P, 8, 2048, sink);
        run<1, 16>("128 B per warp-row, 16 in flight", X, ld, n, P, 8, 2048, sink);
        run<2, 8>("256 B per warp-row, 8 in flight", X, ld, n, P, 8, 2048, sink);
        run<4, 8>("512 B per warp-row, 8 in flight", X, ld, n, P, 8, 2048, sink);
        run<4, 8>("512 B per warp-row, 8 in flight, 4 warps", X, ld, n, P, 4, 2048, sink);
        run<4, 16>("512 B per warp-row, 16 in flight", X, ld, n, P, 8, 2048, sink);
        run<4, 8>("512 B per warp-row, 8 in flight, 144 rows", X, ld, n, P, 8, 144, sink);
        run<4, 4>("512 B per warp-row, 4 in flight", X, ld, n, P, 8, 2048, sink);
        run<1, 8>("128 B per warp-row, 8 in flight, 1 warp/CTA", X, ld, n, P, 1, 2048, sink);
    }
    return 0;
}

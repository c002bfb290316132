// common.cuh -- shared device helpers for the illico_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/illico_b200.h"

namespace illico {

constexpr unsigned FULL = 0xffffffffu;

// ---- host-side error plumbing (api.cu owns the storage) -----------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Per-kernel timing for bench.py's roofline (ILLICO_PROFILE=1): every launch made through ILLICO_LAUNCH is bracketed
// by two CUDA events on the launching stream; illico_profile_report() adds them up per kernel name.  Off by default
// (one getenv per launch).
bool profiling_on();
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);
struct ProfScope {
    cudaStream_t s;
    bool on;
    ProfScope(const char* name, cudaStream_t stream) : s(stream), on(profiling_on()) { if (on) prof_begin(name, s); }
    ~ProfScope() { if (on) prof_end(s); }
};
#define ILLICO_LAUNCH(name, stream, ...)              \
    do {                                              \
        ::illico::ProfScope _ps(name, stream);        \
        __VA_ARGS__;                                  \
        ::illico::count_launch();                     \
    } while (0)

#define ILLICO_CUDA_OK(expr)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::illico::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

// ---- order-preserving key of a float --------------------------------------------------------------
// Ascending unsigned order of the key == ascending order of the value; -0.0 is never keyed because
// zeros are filtered out at staging time (x != 0.0f is false for both zeros).
__device__ __forceinline__ uint32_t f2key(float v) {
    uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
constexpr uint32_t KEY_ZERO = 0x80000000u;  // f2key(+0.0f): keys above it are positive values

// ---- small reductions / scans -----------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Block-wide sum of a u64 / f64; `red` is shared scratch of >= 32 elements; all threads get the result.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    T r = (lane < nw) ? red[lane] : T(0);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
    return r;
}

// lower_bound / upper_bound over a sorted key array (generic address space: shared or global)
__device__ __forceinline__ int lower_bound_u32(const uint32_t* a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int upper_bound_u32(const uint32_t* a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// exact integer t^3 - t
__device__ __forceinline__ long long cube_minus(long long t) { return t * t * t - t; }

constexpr double TWO53 = 9007199254740992.0;

}  // namespace illico

// recode.cu -- values float32 cannot hold (float64) -> order-preserving float32 codes + float64 group sums (sm_100a).
//
// The reference ranks whatever dtype the matrix has (numba specialises dense_ovr_mwu_kernel & co. per dtype,
// illico/ovr/dense_ovr.py:46-53).  The kernels here rank float32 keys.  Ranks, U statistics and tie sums only depend on the
// ORDER of a gene's values and on which of them are equal, so a gene batch of float64 values is first recoded, per gene, to
// small integers with the same order and the same ties (zero -> 0, negatives < 0 < positives; exact as float32 up to 2^24
// distinct values per gene), and the fold change -- the one output that needs the values themselves -- takes float64
// group sums computed here from the original values (illico_flags_t::group_sums).  Three steps, all on the caller's stream:
//   collect   every non-zero of the batch becomes an order-preserving 64-bit key in its gene's list; f(x) is added to
//             the (group, gene) sum (f = expm1 for log1p data, illico/utils/math.py:212);
//   sort      one CTA per gene: LSD radix sort of the 64-bit keys (sort.cuh), then the distinct keys, in order, and how
//             many of them are negative;
//   recode    every non-zero looks its key up in its gene's distinct list (binary search) and becomes its signed rank.
// The codes keep the input's layout (dense matrix / the `data` array of the CSR or CSC structure), so the ordinary
// dispatchers run on them unchanged.
#include "common.cuh"
#include "sort.cuh"

#include <cuda_fp16.h>

namespace illico {

namespace {

constexpr unsigned long long KEY_ZERO64 = 1ull << 63;   // key of +0.0 (zeros are never keyed: x != 0 is false for both)
__device__ __forceinline__ unsigned long long d2key(double x) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | KEY_ZERO64);
}
__device__ __forceinline__ double fval64(double x, int is_log1p) { return is_log1p ? expm1(x) : x; }

struct Lists {
    unsigned long long* keys;     // per-gene lists, gene j at keys + off(j)
    unsigned long long* tmp;      // ping-pong partner of the sort, same layout
    unsigned int* count;          // [b] keys in gene j's list
    unsigned int* ndist;          // [b] distinct keys
    unsigned int* nneg;           // [b] distinct negative keys
    unsigned char* in_tmp;        // [b] 1 = the distinct list of gene j is in `tmp`
    long long cap;                // list capacity per gene when offsets == nullptr (dense, CSR: n_cells)
    const long long* offsets;     // CSC: gene j's list starts at offsets[j] - offsets[0] (= its column's extent)
    __device__ __forceinline__ long long off(int j) const { return offsets ? offsets[j] - offsets[0] : (long long)j * cap; }
};

// signed rank of x among its gene's distinct values: ..., -2, -1 (negatives), 0 (zero), 1, 2, ... (positives)
__device__ __forceinline__ float code_of(const Lists& L, int j, double x) {
    if (!(x != 0.0)) return 0.0f;
    const unsigned long long key = d2key(x);
    const unsigned long long* d = (L.in_tmp[j] ? L.tmp : L.keys) + L.off(j);
    int lo = 0, hi = (int)L.ndist[j];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (d[mid] < key) lo = mid + 1; else hi = mid;
    }
    const int neg = (int)L.nneg[j];
    return (float)(lo < neg ? lo - neg : lo - neg + 1);
}

__device__ __forceinline__ void put(const Lists& L, int j, double x, int group, int b, int is_log1p, double* __restrict__ sums) {
    const unsigned int pos = atomicAdd(&L.count[j], 1u);
    L.keys[L.off(j) + pos] = d2key(x);
    atomicAdd(&sums[(long long)group * b + j], fval64(x, is_log1p));
}

// ---- dense: thread per element, genes fastest (coalesced reads of the row-major matrix)
__global__ void __launch_bounds__(256) wide_collect_dense_kernel(const double* __restrict__ X, long long ld, int gene_lb, int b, long long n,
                                                                 const int32_t* __restrict__ cell_group, int is_log1p, Lists L,
                                                                 double* __restrict__ sums) {
    const long long total = n * b;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / b;
        const int j = (int)(i - r * b);
        const double x = X[r * ld + gene_lb + j];
        if (x != 0.0) put(L, j, x, cell_group[r], b, is_log1p, sums);
    }
}
__global__ void __launch_bounds__(256) wide_recode_dense_kernel(const double* __restrict__ X, long long ld, int gene_lb, int b, long long n,
                                                                Lists L, float* __restrict__ codes) {
    const long long total = n * b;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / b;
        const int j = (int)(i - r * b);
        codes[i] = code_of(L, j, X[r * ld + gene_lb + j]);
    }
}

// ---- CSC: CTA per gene column (its stored values are contiguous)
template <bool RECODE>
__global__ void __launch_bounds__(256) wide_csc_kernel(const double* __restrict__ data, const int32_t* __restrict__ indices,
                                                       const long long* __restrict__ indptr, int gene_lb, int b,
                                                       const int32_t* __restrict__ cell_group, int is_log1p, Lists L,
                                                       double* __restrict__ sums, float* __restrict__ codes) {
    for (int j = blockIdx.x; j < b; j += gridDim.x) {
        const long long e0 = indptr[gene_lb + j], e1 = indptr[gene_lb + j + 1];
        for (long long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
            const double x = data[e];
            if (RECODE) codes[e] = code_of(L, j, x);
            else if (x != 0.0) put(L, j, x, cell_group[indices[e]], b, is_log1p, sums);
        }
    }
}

// ---- CSR: warp per row; the batch's columns form one run of the row (indices ascend)
template <bool RECODE>
__global__ void __launch_bounds__(256) wide_csr_kernel(const double* __restrict__ data, const int32_t* __restrict__ indices,
                                                       const long long* __restrict__ indptr, long long n, int gene_lb, int b,
                                                       const int32_t* __restrict__ cell_group, int is_log1p, Lists L,
                                                       double* __restrict__ sums, float* __restrict__ codes) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < n; r += nwarps) {
        long long lo = indptr[r], hi = indptr[r + 1];
        const long long e1 = hi;
        while (lo < hi) { const long long mid = (lo + hi) >> 1; if (indices[mid] < gene_lb) lo = mid + 1; else hi = mid; }
        const int group = RECODE ? 0 : cell_group[r];
        for (long long e = lo + lane; e < e1; e += 32) {
            const int c = indices[e];
            if (c >= gene_lb + b) break;
            const double x = data[e];
            if (RECODE) codes[e] = code_of(L, c - gene_lb, x);
            else if (x != 0.0) put(L, c - gene_lb, x, group, b, is_log1p, sums);
        }
    }
}

// ---- sort + distinct: CTA per gene
__global__ void __launch_bounds__(256) wide_sort_kernel(int b, Lists L) {
    __shared__ uint32_t hist[8 * 256];
    __shared__ uint32_t aux[RADIX_AUX_WORDS];
    __shared__ int wsum[8];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int j = blockIdx.x; j < b; j += gridDim.x) {
        const int m = (int)L.count[j];
        unsigned long long* a = L.keys + L.off(j);
        unsigned long long* t = L.tmp + L.off(j);
        unsigned long long* s = block_radix_sort_t<unsigned long long>(a, t, m, hist, aux);
        unsigned long long* d = (s == a) ? t : a;                 // the distinct keys go to the other buffer
        __syncthreads();
        int base = 0;
        for (int i0 = 0; i0 < m; i0 += 256) {
            const int i = i0 + tid;
            const bool head = i < m && (i == 0 || s[i - 1] != s[i]);
            const unsigned bal = __ballot_sync(FULL, head);
            if (lane == 0) wsum[w] = __popc(bal);
            __syncthreads();
            int before = 0, total = 0;
            for (int ww = 0; ww < 8; ++ww) { const int c = wsum[ww]; if (ww < w) before += c; total += c; }
            if (head) d[base + before + __popc(bal & ((1u << lane) - 1u))] = s[i];
            base += total;
            __syncthreads();
        }
        if (tid == 0) {
            int lo = 0, hi = base;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (d[mid] < KEY_ZERO64) lo = mid + 1; else hi = mid; }
            L.ndist[j] = (unsigned)base;
            L.nneg[j] = (unsigned)lo;
            L.in_tmp[j] = (d == t) ? 1 : 0;
        }
        __syncthreads();
    }
}

// ---- dtype conversion on upload: any real dtype -> float32 when float32 holds every value exactly, else -> float64
template <typename T>
__device__ __forceinline__ double load_as_double(const void* p, long long i) { return (double)reinterpret_cast<const T*>(p)[i]; }
__device__ __forceinline__ double load_value(const void* p, int dtype, long long i) {
    switch (dtype) {
        case ILLICO_DTYPE_F32: return load_as_double<float>(p, i);
        case ILLICO_DTYPE_F64: return load_as_double<double>(p, i);
        case ILLICO_DTYPE_F16: return (double)__half2float(reinterpret_cast<const __half*>(p)[i]);
        case ILLICO_DTYPE_I8: return load_as_double<signed char>(p, i);
        case ILLICO_DTYPE_U8: return load_as_double<unsigned char>(p, i);
        case ILLICO_DTYPE_I16: return load_as_double<short>(p, i);
        case ILLICO_DTYPE_I32: return load_as_double<int>(p, i);
        default: return (double)reinterpret_cast<const long long*>(p)[i];   // ILLICO_DTYPE_I64 (|x| > 2^53 rounds, as in the reference's float arithmetic)
    }
}
// dst32 (optional): the values as float32, *inexact |= some value changed by that; dst64 (optional): as float64
__global__ void __launch_bounds__(256) convert_values_kernel(const void* __restrict__ src, int dtype, long long count,
                                                             float* __restrict__ dst32, double* __restrict__ dst64, int* __restrict__ inexact) {
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        const double x = load_value(src, dtype, i);
        if (dst64) dst64[i] = x;
        if (dst32) {
            const float f = (float)x;
            dst32[i] = f;
            bad |= !((double)f == x) && (x == x);      // (NaN stays NaN: not a loss)
        }
    }
    if (__any_sync(FULL, bad) && (threadIdx.x & 31) == 0) atomicOr(inexact, 1);
}

size_t lists_header_bytes(int b) { return (((size_t)b * 13 + 255) & ~(size_t)255); }

}  // namespace

// workspace: [count | ndist | nneg | in_tmp] + two key buffers of `total_keys` 64-bit keys each
size_t recode_workspace_bytes(long long total_keys, int b) { return lists_header_bytes(b) + 2 * (size_t)total_keys * 8 + 512; }

static int carve(Lists& L, void* ws, size_t ws_bytes, long long total_keys, int b, long long cap, const long long* offsets, cudaStream_t stream) {
    if (ws_bytes < recode_workspace_bytes(total_keys, b)) { set_error("recode workspace too small: %zu bytes", ws_bytes); return 1; }
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    L.count = reinterpret_cast<unsigned int*>(p);
    L.ndist = L.count + b;
    L.nneg = L.ndist + b;
    L.in_tmp = reinterpret_cast<unsigned char*>(L.nneg + b);
    p += lists_header_bytes(b);
    L.keys = reinterpret_cast<unsigned long long*>(p);
    L.tmp = L.keys + total_keys;
    L.cap = cap;
    L.offsets = offsets;
    ILLICO_CUDA_OK(cudaMemsetAsync(L.count, 0, (size_t)b * sizeof(unsigned int), stream));
    return 0;
}

static int blocks_for(long long work, int per_block) {
    long long blocks = (work + per_block - 1) / per_block;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

int launch_convert_values(const void* src, int dtype, long long count, float* dst32, double* dst64, int* inexact, cudaStream_t stream) {
    if (dtype < ILLICO_DTYPE_F32 || dtype > ILLICO_DTYPE_I64) { set_error("illico_convert_values: unknown dtype %d", dtype); return 1; }
    if (dst32 && inexact) ILLICO_CUDA_OK(cudaMemsetAsync(inexact, 0, sizeof(int), stream));
    if (count <= 0) return 0;
    ILLICO_LAUNCH("convert_values_kernel", stream,
                  convert_values_kernel<<<blocks_for(count, 256 * 8), 256, 0, stream>>>(src, dtype, count, dst32, dst64, inexact));
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_recode_dense(const double* X, long long ld, int gene_lb, int b, long long n, const int32_t* cell_group, int n_groups,
                        int is_log1p, float* codes, double* sums, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (b <= 0 || n <= 0) return 0;
    Lists L;
    if (carve(L, ws, ws_bytes, n * b, b, n, nullptr, stream)) return 1;
    ILLICO_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)n_groups * b * sizeof(double), stream));
    const int blocks = blocks_for(n * b, 256 * 8);
    ILLICO_LAUNCH("wide_collect_dense_kernel", stream, wide_collect_dense_kernel<<<blocks, 256, 0, stream>>>(X, ld, gene_lb, b, n, cell_group, is_log1p, L, sums));
    ILLICO_LAUNCH("wide_sort_kernel", stream, wide_sort_kernel<<<b < 148 * 8 ? b : 148 * 8, 256, 0, stream>>>(b, L));
    ILLICO_LAUNCH("wide_recode_dense_kernel", stream, wide_recode_dense_kernel<<<blocks, 256, 0, stream>>>(X, ld, gene_lb, b, n, L, codes));
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_recode_csc(const double* data, const int32_t* indices, const long long* indptr, int gene_lb, int b, const int32_t* cell_group,
                      int n_groups, long long batch_nnz, int is_log1p, float* codes, double* sums, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
    if (b <= 0) return 0;
    Lists L;
    if (carve(L, ws, ws_bytes, batch_nnz, b, 0, indptr + gene_lb, stream)) return 1;
    ILLICO_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)n_groups * b * sizeof(double), stream));
    const int blocks = b < 148 * 8 ? b : 148 * 8;
    ILLICO_LAUNCH("wide_collect_csc_kernel", stream, wide_csc_kernel<false><<<blocks, 256, 0, stream>>>(data, indices, indptr, gene_lb, b, cell_group, is_log1p, L, sums, codes));
    ILLICO_LAUNCH("wide_sort_kernel", stream, wide_sort_kernel<<<blocks, 256, 0, stream>>>(b, L));
    ILLICO_LAUNCH("wide_recode_csc_kernel", stream, wide_csc_kernel<true><<<blocks, 256, 0, stream>>>(data, indices, indptr, gene_lb, b, cell_group, is_log1p, L, sums, codes));
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_recode_csr(const double* data, const int32_t* indices, const long long* indptr, long long n, int gene_lb, int b,
                      const int32_t* cell_group, int n_groups, int is_log1p, float* codes, double* sums, void* ws, size_t ws_bytes,
                      cudaStream_t stream) {
    if (b <= 0 || n <= 0) return 0;
    Lists L;
    if (carve(L, ws, ws_bytes, n * b, b, n, nullptr, stream)) return 1;
    ILLICO_CUDA_OK(cudaMemsetAsync(sums, 0, (size_t)n_groups * b * sizeof(double), stream));
    const int blocks = blocks_for(n, 8);
    ILLICO_LAUNCH("wide_collect_csr_kernel", stream, wide_csr_kernel<false><<<blocks, 256, 0, stream>>>(data, indices, indptr, n, gene_lb, b, cell_group, is_log1p, L, sums, codes));
    ILLICO_LAUNCH("wide_sort_kernel", stream, wide_sort_kernel<<<b < 148 * 8 ? b : 148 * 8, 256, 0, stream>>>(b, L));
    ILLICO_LAUNCH("wide_recode_csr_kernel", stream, wide_csr_kernel<true><<<blocks, 256, 0, stream>>>(data, indices, indptr, n, gene_lb, b, cell_group, is_log1p, L, sums, codes));
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace illico

// stage.cu -- input formats -> group-segmented, gene-major lists of non-zero values (sm_100a).
//
// Replaces the reference's batch re-layout steps: chunk_and_fortranize (illico/utils/math.py:247-278),
// the CSR/CSC slicers (illico/utils/sparse/csr.py:103-257, csc.py:99-183).  The output ("IR") is the same
// for every input format, so one pair of rank kernels serves all six reference dispatchers:
//
//   ir_vals[gene, seg_base[s] + k]  k-th non-zero value of the cells of segment s (a slice of one group)
//   ir_cnt [gene, s]                how many there are
//
// Zeros are dropped here -- for dense input too -- because the rank kernels treat them as one analytic
// tie block.  This is HBM-bound byte shuffling: the dense kernel reads every
// element exactly once with warp-coalesced row segments and writes only the ~10 % that are non-zero.
#include "common.cuh"

namespace illico {

constexpr int STAGE_WARPS = 8;

// One warp = 32 adjacent genes (one coalesced 128-byte row segment per cell); the CTA's 8 warps cover 256
// adjacent genes, i.e. 1 KB of every row they touch.  Each lane owns one (gene, segment) slot at a time and
// appends the non-zeros of the segment's cells in order (deterministic layout, no atomics).
__global__ void __launch_bounds__(STAGE_WARPS * 32) stage_dense_kernel(const float* __restrict__ X, long long ld,
                                                                       int gene_lb, int b, const illico_plan_t pl,
                                                                       float* __restrict__ ir_vals,
                                                                       uint32_t* __restrict__ ir_cnt,
                                                                       int segs_per_cta) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int jb = (blockIdx.x * STAGE_WARPS + w) * 32 + lane;
    const bool active = jb < b;
    const float* col = X + gene_lb + (active ? jb : 0);
    const int S = pl.n_segments;
    const int s_begin = blockIdx.y * segs_per_cta, s_end = min(S, s_begin + segs_per_cta);
    for (int s = s_begin; s < s_end; ++s) {
        const int p0 = pl.seg_pos[s], p1 = pl.seg_pos[s + 1];
        float* out = ir_vals + (long long)(active ? jb : 0) * pl.slot_cap + pl.seg_base[s];
        uint32_t cnt = 0;
        for (int p = p0; p < p1; p += 32) {
            const int myrow = (p + lane < p1) ? pl.perm[p + lane] : 0;
            const int nrows = min(32, p1 - p);
            for (int k0 = 0; k0 < nrows; k0 += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = k0 + u;
                    const int row = __shfl_sync(FULL, myrow, k & 31);
                    v[u] = (active && k < nrows) ? __ldcs(col + (long long)row * ld) : 0.0f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (v[u] != 0.0f) out[cnt++] = v[u];
            }
        }
        if (active) ir_cnt[(long long)jb * S + s] = cnt;
    }
}

// Vectorised variant: one lane = 4 adjacent genes (one 128-bit load per cell), one warp = 128 genes (512 B of
// the row), one CTA = 4 warps = 512 genes (2 KB of every row it touches).  Non-zeros are buffered four at a
// time per gene stream and written with 128-bit stores into the 16-byte aligned slots; the per-(gene, segment)
// counts go through a shared-memory tile so that each gene's counts leave as one contiguous run.
constexpr int SD_WARPS = 4;
constexpr int SD_GENES = SD_WARPS * 128;
constexpr int SD_MAX_SEGS = 16;

#define ILLICO_APPEND(K, XV)                                                              \
    if ((XV) != 0.0f) {                                                                   \
        const uint32_t m_ = c[K] & 3u;                                                    \
        if (m_ == 0) pend[K].x = (XV);                                                    \
        else if (m_ == 1) pend[K].y = (XV);                                               \
        else if (m_ == 2) pend[K].z = (XV);                                               \
        else { pend[K].w = (XV); *reinterpret_cast<float4*>(out[K] + (c[K] & ~3u)) = pend[K]; } \
        ++c[K];                                                                           \
    }

__global__ void __launch_bounds__(SD_WARPS * 32) stage_dense_v4_kernel(const float* __restrict__ X, long long ld,
                                                                        int gene_lb, int b, const illico_plan_t pl,
                                                                        float* __restrict__ ir_vals,
                                                                        uint32_t* __restrict__ ir_cnt,
                                                                        int segs_per_cta) {
    __shared__ uint32_t cnt_tile[SD_MAX_SEGS][SD_GENES];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int jl = w * 128 + lane * 4;                 // first of this lane's 4 genes inside the CTA tile
    const int jb = blockIdx.x * SD_GENES + jl;         // ... inside the batch (b is a multiple of 4)
    const bool active = jb < b;
    const float4* col = reinterpret_cast<const float4*>(X + gene_lb + (active ? jb : 0));
    const long long ld4 = ld >> 2;
    const int S = pl.n_segments;
    const int s_begin = blockIdx.y * segs_per_cta, s_end = min(S, s_begin + segs_per_cta);
    for (int s = s_begin; s < s_end; ++s) {
        const int p0 = pl.seg_pos[s], p1 = pl.seg_pos[s + 1];
        float* out[4];
        float4 pend[4];
        uint32_t c[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            out[k] = ir_vals + (long long)((active ? jb : 0) + k) * pl.slot_cap + pl.seg_base[s];
            pend[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = p0; p < p1; p += 32) {
            const int myrow = (p + lane < p1) ? pl.perm[p + lane] : 0;
            const int nrows = min(32, p1 - p);
            for (int k0 = 0; k0 < nrows; k0 += 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = k0 + u;
                    const int row = __shfl_sync(FULL, myrow, k & 31);
                    v[u] = (active && k < nrows) ? __ldcs(col + (long long)row * ld4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    ILLICO_APPEND(0, v[u].x)
                    ILLICO_APPEND(1, v[u].y)
                    ILLICO_APPEND(2, v[u].z)
                    ILLICO_APPEND(3, v[u].w)
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // the slot is padded to a multiple of 4 floats: the tail leaves as one (partly unused) 128-bit store
            if (c[k] & 3u) *reinterpret_cast<float4*>(out[k] + (c[k] & ~3u)) = pend[k];
            cnt_tile[s - s_begin][jl + k] = c[k];
        }
    }
    __syncthreads();
    // counts: thread t owns genes t, t+128, ... of the tile and writes their consecutive segments
    const int nseg = s_end - s_begin;
    for (int g = threadIdx.x; g < SD_GENES; g += SD_WARPS * 32) {
        const int j = blockIdx.x * SD_GENES + g;
        if (j >= b) break;
        uint32_t* dst = ir_cnt + (long long)j * S + s_begin;
        for (int ls = 0; ls < nseg; ++ls) dst[ls] = cnt_tile[ls][g];
    }
}

__device__ __forceinline__ long long lower_bound_i32(const int32_t* a, long long n, int key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// CSR: one warp per cell (row).  The row's sorted gene indices are narrowed to the batch by binary search
// (as illico/utils/sparse/csr.py:171,226 does) and every stored value claims the next free place of its
// (gene, segment) slot with one atomic.  The order inside a slot is therefore arbitrary; ranks, U and tie
// sums do not depend on it.
__global__ void __launch_bounds__(256) stage_csr_kernel(const float* __restrict__ data, const int32_t* __restrict__ indices,
                                                        const long long* __restrict__ indptr, int gene_lb, int b,
                                                        const illico_plan_t pl, float* __restrict__ ir_vals,
                                                        uint32_t* __restrict__ ir_cnt) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int S = pl.n_segments;
    // Cells are visited in group order (perm) so that the values of one (gene, segment) slot are written
    // close together in time and merge into full sectors in L2 instead of one DRAM read-modify-write each.
    for (long long p = warp; p < pl.n_cells; p += nwarps) {
        const long long r = pl.perm[p];
        const long long start = indptr[r], end = indptr[r + 1];
        if (end <= start) continue;
        const int s = pl.cell_seg[r];
        const int base = pl.seg_base[s], cap = pl.seg_base[s + 1] - base;
        const long long lo = start + lower_bound_i32(indices + start, end - start, gene_lb);
        const long long hi = start + lower_bound_i32(indices + start, end - start, gene_lb + b);
        for (long long k = lo + lane; k < hi; k += 32) {
            const float v = __ldcs(data + k);
            if (v == 0.0f) continue;  // explicitly stored zeros are zeros
            const int j = indices[k] - gene_lb;
            const uint32_t slot = atomicAdd(&ir_cnt[(long long)j * S + s], 1u);
            if (slot < (uint32_t)cap) ir_vals[(long long)j * pl.slot_cap + base + slot] = v;
        }
    }
}

// CSC: one CTA per gene column of the batch.
__global__ void __launch_bounds__(256) stage_csc_kernel(const float* __restrict__ data, const int32_t* __restrict__ indices,
                                                        const long long* __restrict__ indptr, int gene_lb, int b,
                                                        const illico_plan_t pl, float* __restrict__ ir_vals,
                                                        uint32_t* __restrict__ ir_cnt) {
    const int S = pl.n_segments;
    for (int j = blockIdx.x; j < b; j += gridDim.x) {
        const long long start = indptr[gene_lb + j], end = indptr[gene_lb + j + 1];
        for (long long k = start + threadIdx.x; k < end; k += blockDim.x) {
            const float v = __ldcs(data + k);
            if (v == 0.0f) continue;
            const int s = pl.cell_seg[indices[k]];
            const int base = pl.seg_base[s], cap = pl.seg_base[s + 1] - base;
            const uint32_t slot = atomicAdd(&ir_cnt[(long long)j * S + s], 1u);
            if (slot < (uint32_t)cap) ir_vals[(long long)j * pl.slot_cap + base + slot] = v;
        }
    }
}

// illico/utils/ranking.py:245-273: every row's indices must ascend
__global__ void check_csr_sorted_kernel(const int32_t* __restrict__ indices, const long long* __restrict__ indptr,
                                        long long n_rows, int* flag) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < n_rows; r += nwarps) {
        const long long start = indptr[r], end = indptr[r + 1];
        for (long long k = start + 1 + lane; k < end; k += 32)
            if (indices[k] < indices[k - 1]) *flag = 0;
    }
}

// ------------------------------------------------------------------------------------------------------
int launch_stage_dense(const float* X, long long ld, int gene_lb, int b, const illico_plan_t* plan, float* ir_vals,
                       uint32_t* ir_cnt, cudaStream_t stream) {
    if (b <= 0 || plan->n_segments <= 0) return 0;
    const int S = plan->n_segments;
    long long avg = plan->n_cells / S;
    if (avg < 1) avg = 1;
    const bool vec = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && (ld % 4 == 0) && (gene_lb % 4 == 0) && (b % 4 == 0);
    int segs_per_cta = (int)(1024 / avg);
    if (segs_per_cta < 1) segs_per_cta = 1;
    const int cap = vec ? SD_MAX_SEGS : 64;
    if (segs_per_cta > cap) segs_per_cta = cap;
    long long gy = (S + segs_per_cta - 1) / segs_per_cta;
    if (gy > 65535) {
        if (vec) { set_error("too many segments (%d) for one staging launch", S); return 1; }
        while (gy > 65535) { segs_per_cta *= 2; gy = (S + segs_per_cta - 1) / segs_per_cta; }
    }
    if (vec) {
        dim3 grid((b + SD_GENES - 1) / SD_GENES, (unsigned)gy);
        stage_dense_v4_kernel<<<grid, SD_WARPS * 32, 0, stream>>>(X, ld, gene_lb, b, *plan, ir_vals, ir_cnt, segs_per_cta);
    } else {
        dim3 grid((b + STAGE_WARPS * 32 - 1) / (STAGE_WARPS * 32), (unsigned)gy);
        stage_dense_kernel<<<grid, STAGE_WARPS * 32, 0, stream>>>(X, ld, gene_lb, b, *plan, ir_vals, ir_cnt, segs_per_cta);
    }
    count_launch();
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_stage_csr(const float* data, const int32_t* indices, const long long* indptr, int gene_lb, int b,
                     const illico_plan_t* plan, float* ir_vals, uint32_t* ir_cnt, cudaStream_t stream) {
    if (b <= 0 || plan->n_cells <= 0) return 0;
    long long warps = plan->n_cells;
    long long blocks = (warps + 7) / 8;
    if (blocks > 148 * 64) blocks = 148 * 64;
    stage_csr_kernel<<<(unsigned)blocks, 256, 0, stream>>>(data, indices, indptr, gene_lb, b, *plan, ir_vals, ir_cnt);
    count_launch();
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_stage_csc(const float* data, const int32_t* indices, const long long* indptr, int gene_lb, int b,
                     const illico_plan_t* plan, float* ir_vals, uint32_t* ir_cnt, cudaStream_t stream) {
    if (b <= 0) return 0;
    int blocks = b < 148 * 32 ? b : 148 * 32;
    stage_csc_kernel<<<blocks, 256, 0, stream>>>(data, indices, indptr, gene_lb, b, *plan, ir_vals, ir_cnt);
    count_launch();
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_check_csr_sorted(const int32_t* indices, const long long* indptr, long long n_rows, int* d_flag,
                            int* h_sorted, cudaStream_t stream) {
    int one = 1;
    ILLICO_CUDA_OK(cudaMemcpyAsync(d_flag, &one, sizeof(int), cudaMemcpyHostToDevice, stream));
    if (n_rows > 0) {
        long long blocks = (n_rows + 7) / 8;
        if (blocks > 148 * 32) blocks = 148 * 32;
        check_csr_sorted_kernel<<<(unsigned)blocks, 256, 0, stream>>>(indices, indptr, n_rows, d_flag);
        count_launch();
        ILLICO_CUDA_OK(cudaGetLastError());
    }
    int h = -1;
    ILLICO_CUDA_OK(cudaMemcpyAsync(&h, d_flag, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ILLICO_CUDA_OK(cudaStreamSynchronize(stream));
    *h_sorted = h;
    return 0;
}

}  // namespace illico

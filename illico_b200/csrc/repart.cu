// repart.cu -- rows -> genes repartition of a CSR matrix across GPUs (sm_100a).
//
// One-versus-reference on a 2M-cell x 20k-gene CSR matrix (BASELINE configs[4]) shards by GENES, but a CSR matrix is laid
// out by ROWS: every row holds entries of every shard.  The reference cuts gene batches out of the rows on the CPU
// (csr_get_contig_cols_into_csr, illico/utils/sparse/csr.py:144-196).  Here every GPU receives a contiguous block of ROWS
// (one plain copy of its slice of data / indices, no host work, each byte crosses PCIe once), finds where its rows cross
// the shard boundaries (column indices ascend inside a row: one binary search per row and boundary), and then writes every
// row piece STRAIGHT INTO THE OWNING GPU'S shard arrays through peer memory (NVLink): the exchange is the store
// instruction of the kernel that cuts the rows, there is no send buffer and no separate collective.
#include "common.cuh"

namespace illico {

namespace {

// cnt[j * n_rows + r] = entries of row r with bounds[j] <= column < bounds[j + 1]; totals[j] += the block's sum.
// Thread per (row, shard boundary) would search 2 (S + 1) times per row; a thread per row walks the S - 1 inner
// boundaries left to right instead (each search starts where the last one ended).
__global__ void __launch_bounds__(256) csr_shard_count_kernel(const int32_t* __restrict__ indices, const long long* __restrict__ indptr,
                                                              long long n_rows, const int32_t* __restrict__ bounds, int n_shards,
                                                              int32_t* __restrict__ cnt, unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long tot[32];
    if (threadIdx.x < 32) tot[threadIdx.x] = 0ull;
    __syncthreads();
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
        const long long e0 = indptr[r], e1 = indptr[r + 1];
        long long lo = e0;
        {   // entries below the first boundary belong to nobody (the caller passes bounds[0] = 0)
            long long hi = e1;
            const int32_t b0 = bounds[0];
            while (lo < hi) { const long long mid = (lo + hi) >> 1; if (indices[mid] < b0) lo = mid + 1; else hi = mid; }
        }
        for (int j = 0; j < n_shards; ++j) {
            const int32_t bj = bounds[j + 1];
            long long a = lo, hi = e1;
            while (a < hi) { const long long mid = (a + hi) >> 1; if (indices[mid] < bj) a = mid + 1; else hi = mid; }
            const int c = (int)(a - lo);
            cnt[(long long)j * n_rows + r] = c;
            if (c) atomicAdd(&tot[j & 31], (unsigned long long)c);   // (n_shards <= 32)
            lo = a;
        }
    }
    __syncthreads();
    if (threadIdx.x < n_shards && tot[threadIdx.x]) atomicAdd(totals + threadIdx.x, tot[threadIdx.x]);
}

// Warp per row: the row's piece for shard j (its entries are contiguous in the row) goes to the shard's arrays at
// out_pos[j][r] -- arrays that may live on another GPU (peer-mapped pointers) -- with the column index rebased to the
// shard; the row's count goes to the shard's row-count array at its global row number.
struct ShardOut {
    float* data[32];
    int32_t* indices[32];
    int32_t* row_cnt[32];
};
__global__ void __launch_bounds__(256) csr_shard_scatter_kernel(const float* __restrict__ data, const int32_t* __restrict__ indices,
                                                                const long long* __restrict__ indptr, long long n_rows,
                                                                long long row0, const int32_t* __restrict__ bounds, int n_shards,
                                                                const int32_t* __restrict__ cnt,
                                                                const long long* __restrict__ out_pos, ShardOut out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < n_rows; r += nwarps) {
        long long e = indptr[r];
        {   // skip what lies below the first boundary
            long long hi = indptr[r + 1];
            const int32_t b0 = bounds[0];
            while (e < hi) { const long long mid = (e + hi) >> 1; if (indices[mid] < b0) e = mid + 1; else hi = mid; }
        }
        for (int j = 0; j < n_shards; ++j) {
            const int c = cnt[(long long)j * n_rows + r];
            if (lane == 0) out.row_cnt[j][row0 + r] = c;
            if (c) {
                const long long o = out_pos[(long long)j * n_rows + r];
                const int32_t base = bounds[j];
                for (int i = lane; i < c; i += 32) {
                    out.data[j][o + i] = data[e + i];
                    out.indices[j][o + i] = indices[e + i] - base;
                }
                e += c;
            }
        }
    }
}

}  // namespace

int launch_csr_shard_count(const int32_t* indices, const long long* indptr, long long n_rows, const int32_t* bounds, int n_shards,
                           int32_t* cnt, unsigned long long* totals, cudaStream_t stream) {
    if (n_shards < 1 || n_shards > 32) { set_error("illico_csr_shard_count: 1 .. 32 shards"); return 1; }
    if (n_rows <= 0) return 0;
    long long blocks = (n_rows + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ILLICO_LAUNCH("csr_shard_count_kernel", stream,
                  csr_shard_count_kernel<<<(unsigned)blocks, 256, 0, stream>>>(indices, indptr, n_rows, bounds, n_shards, cnt, totals));
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_csr_shard_scatter(const float* data, const int32_t* indices, const long long* indptr, long long n_rows, long long row0,
                             const int32_t* bounds, int n_shards, const int32_t* cnt, const long long* out_pos,
                             float* const* out_data, int32_t* const* out_indices, int32_t* const* out_row_cnt, cudaStream_t stream) {
    if (n_shards < 1 || n_shards > 32) { set_error("illico_csr_shard_scatter: 1 .. 32 shards"); return 1; }
    if (n_rows <= 0) return 0;
    ShardOut out;
    for (int j = 0; j < 32; ++j) {
        out.data[j] = j < n_shards ? out_data[j] : nullptr;
        out.indices[j] = j < n_shards ? out_indices[j] : nullptr;
        out.row_cnt[j] = j < n_shards ? out_row_cnt[j] : nullptr;
    }
    long long blocks = (n_rows + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    ILLICO_LAUNCH("csr_shard_scatter_kernel", stream,
                  csr_shard_scatter_kernel<<<(unsigned)blocks, 256, 0, stream>>>(data, indices, indptr, n_rows, row0, bounds, n_shards, cnt,
                                                                                out_pos, out));
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace illico

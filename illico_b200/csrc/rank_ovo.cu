// rank_ovo.cu -- one-versus-reference Mann-Whitney U on group-segmented non-zero lists (sm_100a).
//
// Replaces illico/ovo/dense_ovo.py:15-137, illico/ovo/sparse_ovo.py:22-158 and the merge
// illico/utils/ranking.py:52-158.  The reference re-walks the whole sorted control column for every
// perturbation (G x n_ref steps per gene).  Here the control's non-zero values are sorted ONCE per gene
// into shared memory and every perturbation value is ranked against it by binary search:
//
//     2U_g = sum_{v in g} ( 2 #{ref > v} + #{ref == v} )                      
//     T_g  = T_ref + sum_{distinct v in g} [ (a+b)^3 - (a+b) - (a^3 - a) ] + (Z^3 - Z)
//
// with a / b the multiplicity of v in the control / the perturbation and Z the number of zeros of the
// pair (zeros are never stored: they are one analytic tie block, for dense input too).  Everything is
// exact integer arithmetic; the f64 epilogue follows illico/utils/math.py:95-118 operation by operation.
//
// One CTA per gene (persistent, grid-strided).  Perturbations are processed in three tiers by their
// number of non-zeros m:  m <= small_cap  one THREAD per group (insertion sort in a private shared
// column), m <= WARP_CAP one WARP per group (bitonic sort in shared memory), larger groups one at a time
// by the whole CTA (radix sort in the CTA's global slab).
#include "common.cuh"
#include "epilogue.cuh"
#include "sort.cuh"

#include <stdlib.h>

namespace illico {

constexpr int WARP_CAP = 1024;   // keys per warp buffer in the warp tier
constexpr int GROUP_CHUNK = 2048;  // groups handled per sweep (bounds the deferred lists)
constexpr int DT_CAP = 22;         // distinct control values for the table fast path (<= small_cap)
constexpr int DT_HASH = 64;        // slots of the key -> table-index hash (load factor <= 1/3)
constexpr int ST_CAP = 1024;      // distinct control values kept in the shared-memory search table
constexpr int FAST_MAX = 96;       // largest group (non-zeros) a single thread streams through the table path

struct OvoParams {
    const float* ir_vals;
    const uint32_t* ir_cnt;
    int n_genes;
    illico_plan_t plan;
    illico_flags_t flags;
    double* results;
    long long gstride;
    uint32_t* slab;          // global scratch, slab_words per CTA
    long long slab_words;
    int ref_cap;             // capacity (keys) of each of the two control buffers
    int small_cap;           // thread-tier capacity (keys per thread)
    int scratch_words;       // shared scratch (control ping-pong partner, later the tier buffers)
    int use_search_table;    // ILLICO_OVO_SEARCH_TABLE
    long long* dbg_u2;
    double* dbg_tie;
    long long* dbg_tie_exact;
};

struct RefInfo {
    const uint32_t* keys;  // sorted non-zero control keys
    int nnz;               // how many
    int npos;              // control values > 0
    long long zeros;       // control zeros
    long long n_ref;
    unsigned long long tie;  // sum over control non-zero runs of a^3 - a
    double sum;              // sum of f(x) over the control
    // search table (optional, st_n >= 0): the control's distinct keys ascending and the position of each one's first
    // occurrence in `keys` (st_lo[st_n] = nnz).  A rank then costs a binary search over <= 1024 shared-memory entries
    // instead of one over the whole sorted control, which lives in the CTA's global slab when it is large: measured
    // 25 ms per dense high-count gene without it (every probe a dependent L2 load).
    const uint32_t* st_key;
    const int* st_lo;
    int st_n;
};

// contribution of one distinct perturbation value (key, multiplicity b)
__device__ __forceinline__ void rank_value(const RefInfo& R, uint32_t key, long long b, unsigned long long& u2,
                                           unsigned long long& tie) {
    int lo, hi;
    if (R.st_n >= 0) {
        const int t = lower_bound_u32(R.st_key, R.st_n, key);
        lo = R.st_lo[t];
        hi = (t < R.st_n && R.st_key[t] == key) ? R.st_lo[t + 1] : lo;
    } else {
        lo = lower_bound_u32(R.keys, R.nnz, key);
        hi = lo;
        if (lo < R.nnz && R.keys[lo] == key) hi = upper_bound_u32(R.keys, R.nnz, key);
    }
    long long a = hi - lo;
    long long gt = (long long)(R.nnz - hi) + ((key < KEY_ZERO) ? R.zeros : 0);
    u2 += (unsigned long long)(b * (2 * gt + a));
    long long t = a + b;
    tie += (unsigned long long)(cube_minus(t) - cube_minus(a));
}

// The reference's sequential f64 tie accumulation for ONE pair (illico/utils/ranking.py:52-158 for the dense
// kernels: merged runs ascending, only t > 1 added, the zero block at its sorted position;
// illico/ovo/sparse_ovo.py:85 for the sparse ones: non-zero runs, then Z^3 - Z last).  Only needed when the
// exact sum reaches 2^53 (pairs of more than 208 063 cells); one thread walks both sorted key lists.
__device__ __noinline__ double ordered_pair_tie(const uint32_t* A, int nA, const uint32_t* B, int strideB, int nB,
                                                long long Z, bool sparse_order) {
    int i = 0, jx = 0;
    double acc = 0.0;
    bool zero_done = sparse_order || Z == 0;
    while (i < nA || jx < nB) {
        const uint32_t ka = (i < nA) ? A[i] : 0xffffffffu, kb = (jx < nB) ? B[(long long)jx * strideB] : 0xffffffffu;
        const uint32_t k = min(ka, kb);
        if (!zero_done && k > KEY_ZERO) {
            if (Z > 1) acc += (double)cube_minus(Z);
            zero_done = true;
        }
        long long t = 0;
        while (i < nA && A[i] == k) { ++i; ++t; }
        while (jx < nB && B[(long long)jx * strideB] == k) { ++jx; ++t; }
        if (t > 1) acc += (double)cube_minus(t);
    }
    if (!zero_done && Z > 1) acc += (double)cube_minus(Z);
    if (sparse_order) acc += (double)cube_minus(Z);
    return acc;
}

__device__ __forceinline__ void finalize_group(const OvoParams& P, const RefInfo& R, int j, int g, long long m,
                                               unsigned long long u2, unsigned long long tie_nz, double sum,
                                               const uint32_t* gkeys = nullptr, int gstride = 1) {
    const long long n_t = P.plan.group_size[g];
    const long long z_t = n_t - m;
    if (P.flags.group_sums) sum = P.flags.group_sums[(long long)g * P.n_genes + j];
    const long long Z = R.zeros + z_t;
    u2 += (unsigned long long)(z_t * (2ll * R.npos + R.zeros));
    const unsigned long long tie_exact = R.tie + tie_nz + (unsigned long long)cube_minus(Z);
    // Every partial sum of the reference's sequential f64 accumulation is an exact integer while the
    // total stays below 2^53 (pairs of up to 208 063 cells), so the exact sum converts without rounding;
    // above it the accumulation is replayed in the reference's order.
    double tie = (double)tie_exact;
    if (tie >= TWO53 && gkeys != nullptr)
        tie = ordered_pair_tie(R.keys, R.nnz, gkeys, gstride, (int)m, Z, P.flags.tie_order == ILLICO_TIES_SPARSE);
    const double U = (double)u2 / 2.0;
    const double mu = (double)(R.n_ref * n_t) / 2.0;
    const double cc = P.flags.use_continuity ? 0.5 : 0.0;
    const double p = compute_pval(R.n_ref, n_t, R.n_ref + n_t, P.flags.tie_correct ? tie : 0.0, U, mu, cc,
                                  P.flags.alternative);
    const double mean_t = sum / (double)n_t;
    const double mean_r = R.sum / (double)R.n_ref;
    const double fc = (mean_r == 0.0) ? INFINITY : mean_t / mean_r;
    double* o = P.results + (long long)g * P.gstride + (long long)j * 3;
    o[0] = p; o[1] = U; o[2] = fc;
    const long long di = (long long)g * P.n_genes + j;
    if (P.dbg_u2) P.dbg_u2[di] = (long long)u2;
    if (P.dbg_tie) P.dbg_tie[di] = tie;
    if (P.dbg_tie_exact) P.dbg_tie_exact[di] = (long long)tie_exact;
}

template <int OVO_THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(OVO_THREADS, MIN_CTAS) ovo_kernel(const OvoParams P) {
    constexpr int OVO_NW = OVO_THREADS / 32;
    extern __shared__ __align__(16) uint32_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const illico_plan_t& pl = P.plan;
    const int S = pl.n_segments, G = pl.n_groups, ref = pl.ref_group;

    // ---- shared carve-up
    uint32_t* refA = smem;                                  // [ref_cap]
    uint32_t* scratch = refA + P.ref_cap;                   // [scratch_words] (>= ref_cap)
    uint32_t* hist = scratch + P.scratch_words;             // [OVO_NW * 256]
    uint32_t* aux = hist + OVO_NW * 256;                    // [RADIX_AUX_WORDS]
    uint16_t* mlist = (uint16_t*)(aux + RADIX_AUX_WORDS);   // [GROUP_CHUNK] group index inside the chunk
    uint16_t* blist = mlist + GROUP_CHUNK;                  // [GROUP_CHUNK]
    int* counters = (int*)(blist + GROUP_CHUNK);            // [8] three rotating {medium, big} list counters + table counter
    double* redd = (double*)(counters + 8);                 // [32]
    unsigned long long* redu = (unsigned long long*)(redd + 32);  // [32]
    double* dval = (double*)(redu + 32);                    // [DT_CAP]   f(x) of each distinct control value
    unsigned long long* dw = (unsigned long long*)(dval + DT_CAP);  // [DT_CAP] 2 #{ref > v} + a
    unsigned long long* dA3 = dw + DT_CAP;                  // [DT_CAP]   3 a^2 - 1
    uint2* hkv = (uint2*)(dA3 + DT_CAP);                    // [DT_HASH]  open-addressed {float bits, byte offset of the bin}
    uint32_t* dkey = (uint32_t*)(hkv + DT_HASH);            // [DT_CAP]   distinct control keys, ascending
    int* dlo = (int*)(dkey + DT_CAP);                       // [DT_CAP+1] first position of each in the sorted control
    uint32_t* st_key = (uint32_t*)(dlo + DT_CAP + 1);       // [ST_CAP]   search table: distinct control keys, ascending
    int* st_lo = (int*)(st_key + ST_CAP);                   // [ST_CAP+1] first position of each

    uint32_t* slab = P.slab + (long long)blockIdx.x * P.slab_words;
    const int maxg = pl.max_group_size;
    uint32_t* gA = slab;            // block-tier group buffers (global)
    uint32_t* gB = slab + maxg;
    __shared__ int mmax3[3];

    const int ref_s0 = pl.group_seg[ref], ref_s1 = pl.group_seg[ref + 1];
    if (tid < 8) counters[tid] = 0;
    if (tid < 3) mmax3[tid] = 0;
    int cc = 0;  // chunk counter: chunk c uses counter set c % 3 and clears set (c + 1) % 3 for the next chunk
    __syncthreads();

    for (int j = blockIdx.x; j < P.n_genes; j += gridDim.x) {
        const uint32_t* cnt = P.ir_cnt + (long long)j * S;
        const float* vals = P.ir_vals + (long long)j * pl.slot_cap;

        // ================= phase 1: control keys, sorted once per gene =================
        // offsets of the control's segments (few): serial prefix by thread 0 into hist[] (free now)
        if (tid == 0) {
            uint32_t acc = 0;
            for (int s = ref_s0; s < ref_s1; ++s) { hist[s - ref_s0] = acc; acc += cnt[s]; }
            hist[ref_s1 - ref_s0] = acc;
        }
        __syncthreads();
        const int nref_nz = (int)hist[ref_s1 - ref_s0];
        // control keys live in shared memory when they fit, else in this CTA's global slab
        const bool ref_smem = nref_nz <= P.ref_cap;
        uint32_t* rA = ref_smem ? refA : slab + 2ll * maxg;
        uint32_t* rB = ref_smem ? scratch : slab + 3ll * maxg;
        double rsum = 0.0;
        for (int s = ref_s0 + w; s < ref_s1; s += OVO_NW) {
            const uint32_t off = hist[s - ref_s0];
            const int c = (int)cnt[s];
            const float* src = vals + pl.seg_base[s];
            for (int i = lane; i < c; i += 32) {
                float v = src[i];
                rA[off + i] = f2key(v);
                rsum += fc_value(v, P.flags.is_log1p);
            }
        }
        __syncthreads();
        rsum = block_sum<double>(rsum, redd);
        const uint32_t* rk = block_radix_sort(rA, rB, nref_nz, hist, aux);
        if (rk != rA) {  // keep the sorted keys in rA: rB is recycled as tier scratch
            for (int i = tid; i < nref_nz; i += OVO_THREADS) rA[i] = rB[i];
            __syncthreads();
        }
        RefInfo R;
        R.keys = rA;
        R.nnz = nref_nz;
        R.n_ref = pl.group_size[ref];
        R.zeros = R.n_ref - nref_nz;
        R.npos = nref_nz - upper_bound_u32(rA, nref_nz, KEY_ZERO);
        R.sum = P.flags.group_sums ? P.flags.group_sums[(long long)ref * P.n_genes + j] : rsum;
        {
            unsigned long long t = 0;
            for (int i = tid; i < nref_nz; i += OVO_THREADS) {
                uint32_t k = rA[i];
                if (i == 0 || rA[i - 1] != k) {
                    long long a = upper_bound_u32(rA, nref_nz, k) - i;
                    t += (unsigned long long)cube_minus(a);
                }
            }
            R.tie = block_sum<unsigned long long>(t, redu);
        }
        // ---- distinct-value table of the control (raw counts have a handful of distinct values): groups whose
        // values all occur in it are ranked by a private histogram over the table, without sorting or searching
        if (tid == 0) counters[6] = 0;
        __syncthreads();
        for (int i = tid; i < nref_nz; i += OVO_THREADS) {
            const uint32_t k = rA[i];
            if (i == 0 || rA[i - 1] != k) {
                const int slot = atomicAdd(&counters[6], 1);
                if (slot < DT_CAP) { dkey[slot] = k; dlo[slot] = i; }
            }
        }
        __syncthreads();
        const int D = counters[6];
        const bool table = D <= DT_CAP;
        // ---- search table for controls with more distinct values than the histogram path takes (D counts them all)
        R.st_key = st_key; R.st_lo = st_lo; R.st_n = -1;
        if (!table && D <= ST_CAP && P.use_search_table) {
            // ordered compaction of the run starts of rA: chunks of OVO_THREADS elements, ballot + warp totals
            int base = 0;
            for (int i0 = 0; i0 < nref_nz; i0 += OVO_THREADS) {
                const int i = i0 + tid;
                const bool head = i < nref_nz && (i == 0 || rA[i - 1] != rA[i]);
                const unsigned bal = __ballot_sync(FULL, head);
                if (lane == 0) hist[w] = (uint32_t)__popc(bal);
                __syncthreads();
                int before = 0, total = 0;
                for (int ww = 0; ww < OVO_NW; ++ww) { const int c = (int)hist[ww]; if (ww < w) before += c; total += c; }
                if (head) {
                    const int slot = base + before + __popc(bal & ((1u << lane) - 1u));
                    st_key[slot] = rA[i];
                    st_lo[slot] = i;
                }
                base += total;
                __syncthreads();
            }
            if (tid == 0) st_lo[D] = nref_nz;
            R.st_n = D;
            __syncthreads();
        }
        if (table && tid == 0) {
            for (int a = 1; a < D; ++a) {  // order the <= 22 entries by position (= by key)
                const uint32_t k = dkey[a];
                const int l = dlo[a];
                int q = a - 1;
                while (q >= 0 && dlo[q] > l) { dlo[q + 1] = dlo[q]; dkey[q + 1] = dkey[q]; --q; }
                dlo[q + 1] = l; dkey[q + 1] = k;
            }
            dlo[D] = nref_nz;
            for (int a = 0; a < DT_HASH; ++a) hkv[a] = make_uint2(0u, 0u);  // +0.0f is never staged
            for (int a = 0; a < D; ++a) {
                const float v = key2f(dkey[a]);
                dval[a] = fc_value(v, P.flags.is_log1p);
                const unsigned long long mult = (unsigned long long)(dlo[a + 1] - dlo[a]);
                const unsigned long long gt = (unsigned long long)(nref_nz - dlo[a + 1]) +
                                              ((dkey[a] < KEY_ZERO) ? (unsigned long long)R.zeros : 0ull);
                dw[a] = 2ull * gt + mult;
                dA3[a] = 3ull * mult * mult - 1ull;
                const uint32_t bits = __float_as_uint(v);
                uint32_t h = (bits * 2654435761u) >> 26;
                while (hkv[h].x != 0u) h = (h + 1) & (DT_HASH - 1);
                hkv[h] = make_uint2(bits, (uint32_t)(a * OVO_THREADS * 4));
            }
        }
        __syncthreads();
        const uint32_t hkv_s = (uint32_t)__cvta_generic_to_shared(hkv);

        // ================= phase 2: perturbations, in chunks of GROUP_CHUNK groups =================
        for (int g0 = 0; g0 < G; g0 += GROUP_CHUNK) {
            const int g1 = min(G, g0 + GROUP_CHUNK);
            int* cnt_m = counters + 2 * (cc % 3);      // [0] medium list length, [1] big list length
            int* m_max = mmax3 + (cc % 3);             // largest non-zero count among the chunk's warp-tier groups
            if (tid == 0) { const int nx = (cc + 1) % 3; counters[2 * nx] = 0; counters[2 * nx + 1] = 0; mmax3[nx] = 0; }
            ++cc;
            // ---- thread tier
            for (int g = g0 + tid; g < g1; g += OVO_THREADS) {
                if (g == ref) {
                    // control row: the sparse kernels' convention (ovo/sparse_ovo.py:140-143); fold change of the
                    // control against itself (utils/math.py:191-192)
                    double* o = P.results + (long long)g * P.gstride + (long long)j * 3;
                    double mean_r = R.sum / (double)R.n_ref;
                    o[0] = 1.0; o[1] = -1.0; o[2] = (mean_r == 0.0) ? INFINITY : mean_r / mean_r;
                    const long long di = (long long)g * P.n_genes + j;
                    if (P.dbg_u2) P.dbg_u2[di] = -2;
                    if (P.dbg_tie) P.dbg_tie[di] = 0.0;
                    if (P.dbg_tie_exact) P.dbg_tie_exact[di] = 0;
                    continue;
                }
                const int s0 = pl.group_seg[g], s1 = pl.group_seg[g + 1];
                int m = 0;
                for (int s = s0; s < s1; ++s) m += (int)cnt[s];
                uint32_t* col = scratch + tid;  // private column: element k at col[k * OVO_THREADS]
                const bool big_pair = R.n_ref + (long long)pl.group_size[g] > 208063;  // tie sum may pass 2^53
                if (table && m <= FAST_MAX && !big_pair) {
                    // ---- table path: private histogram over the control's distinct values
                    for (int t = 0; t < D; ++t) col[t * OVO_THREADS] = 0;
                    bool ok = true;
                    int ne = 0;                                   // values the control does not have: (bits, count)
                    const int ne_cap = (P.small_cap - D) >> 1;    // pairs kept after the D histogram bins
                    const uint32_t col_s = (uint32_t)__cvta_generic_to_shared(col);
                    // one stored value: hash probe (one 64-bit shared load on a hit) + bump of the private bin
                    auto bump = [&](float v) {
                        const uint32_t bits = __float_as_uint(v);
                        uint32_t h = (bits * 2654435761u) >> 26;
                        uint32_t k, off;
                        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(k), "=r"(off) : "r"(hkv_s + h * 8u));
                        if (k != bits) {  // collision chain, or a value the control does not have (rare)
                            while (k != bits && k != 0u) {
                                h = (h + 1) & (DT_HASH - 1);
                                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(k), "=r"(off) : "r"(hkv_s + h * 8u));
                            }
                            if (k == 0u) {
                                int x = 0;
                                while (x < ne && col[(D + 2 * x) * OVO_THREADS] != bits) ++x;
                                if (x < ne) col[(D + 2 * x + 1) * OVO_THREADS] += 1;
                                else if (ne < ne_cap) {
                                    col[(D + 2 * ne) * OVO_THREADS] = bits;
                                    col[(D + 2 * ne + 1) * OVO_THREADS] = 1;
                                    ++ne;
                                } else ok = false;
                                return;
                            }
                        }
                        uint32_t r;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(col_s + off) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" :: "r"(col_s + off), "r"(r + 1u) : "memory");
                    };
                    for (int s = s0; s < s1; ++s) {
                        const int c = (int)cnt[s];
                        const float4* src4 = reinterpret_cast<const float4*>(vals + pl.seg_base[s]);  // 32-byte aligned slot
                        const int nfull = c >> 2;
                        float4 nxt = (c > 0) ? src4[0] : make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int i4 = 0; i4 < nfull; ++i4) {
                            const float4 q4 = nxt;
                            if (4 * i4 + 4 < c) nxt = src4[i4 + 1];  // next 16 bytes are in flight while these are ranked
                            bump(q4.x); bump(q4.y); bump(q4.z); bump(q4.w);
                        }
                        const int rem = c & 3;
                        if (rem > 0) bump(nxt.x);
                        if (rem > 1) bump(nxt.y);
                        if (rem > 2) bump(nxt.z);
                    }
                    if (ok) {
                        unsigned long long u2 = 0, tie = 0;
                        double sum = 0.0;
                        for (int t = 0; t < D; ++t) {
                            const uint32_t bq = col[t * OVO_THREADS];
                            if (bq) {
                                // b (2 gt + a)  and  (a+b)^3 - (a+b) - (a^3 - a) = b (3a^2 - 1 + b (3a + b))
                                const uint32_t a3 = 3u * (uint32_t)(dlo[t + 1] - dlo[t]);
                                u2 += (unsigned long long)bq * dw[t];
                                tie += (unsigned long long)bq * (dA3[t] + (unsigned long long)bq * (unsigned long long)(a3 + bq));
                                sum += (double)bq * dval[t];
                            }
                        }
                        for (int x = 0; x < ne; ++x) {  // absent from the control: a = 0, position by binary search
                            const uint32_t key = f2key(__uint_as_float(col[(D + 2 * x) * OVO_THREADS]));
                            const long long bq = col[(D + 2 * x + 1) * OVO_THREADS];
                            const long long gt = (long long)(nref_nz - lower_bound_u32(rA, nref_nz, key)) +
                                                 ((key < KEY_ZERO) ? R.zeros : 0);
                            u2 += (unsigned long long)(bq * 2 * gt);
                            tie += (unsigned long long)cube_minus(bq);
                            sum += (double)bq * fc_value(key2f(key), P.flags.is_log1p);
                        }
                        finalize_group(P, R, j, g, m, u2, tie, sum);
                        continue;
                    }
                    // too many values the control does not have: a whole warp ranks this group (general path)
                    mlist[atomicAdd(&cnt_m[0], 1)] = (uint16_t)(g - g0);
                    atomicMax(m_max, m);
                    continue;
                }
                if (m > P.small_cap) {
                    mlist[atomicAdd(&cnt_m[0], 1)] = (uint16_t)(g - g0);
                    atomicMax(m_max, m);
                    continue;
                }
                double sum = 0.0;
                int k = 0;
                for (int s = s0; s < s1; ++s) {
                    const int c = (int)cnt[s];
                    const float* src = vals + pl.seg_base[s];
                    for (int i = 0; i < c; ++i) {
                        float v = src[i];
                        uint32_t key = f2key(v);
                        // insertion sort, ascending
                        int q = k - 1;
                        while (q >= 0 && col[q * OVO_THREADS] > key) { col[(q + 1) * OVO_THREADS] = col[q * OVO_THREADS]; --q; }
                        col[(q + 1) * OVO_THREADS] = key;
                        ++k;
                    }
                }
                unsigned long long u2 = 0, tie = 0;
                int i = 0;
                while (i < m) {
                    uint32_t key = col[i * OVO_THREADS];
                    int r = i + 1;
                    while (r < m && col[r * OVO_THREADS] == key) ++r;
                    rank_value(R, key, r - i, u2, tie);
                    sum += (double)(r - i) * fc_value(key2f(key), P.flags.is_log1p);
                    i = r;
                }
                finalize_group(P, R, j, g, m, u2, tie, sum, col, OVO_THREADS);
            }
            __syncthreads();
            // ---- warp tier (usually empty: then this barrier is the only one of the chunk)
            const int nm = cnt_m[0];
            if (nm == 0) continue;
            // warp-tier buffers: 512 keys each when every group of the chunk fits (all 8 warps then own one; with 1024-key
            // buffers only 5 do -- a dense high-count gene has all its 2000 groups here, and its CTA is the kernel's tail)
            const int wcap = (*m_max <= WARP_CAP / 2) ? WARP_CAP / 2 : WARP_CAP;
            const int nwb = min(OVO_NW, P.scratch_words / wcap);
            for (int e = w; e < nm && w < nwb; e += nwb) {
                const int g = g0 + mlist[e];
                const int s0 = pl.group_seg[g], s1 = pl.group_seg[g + 1];
                int m = 0;
                for (int s = s0; s < s1; ++s) m += (int)cnt[s];
                if (m > WARP_CAP) {
                    if (lane == 0) blist[atomicAdd(&cnt_m[1], 1)] = (uint16_t)(g - g0);
                    continue;
                }
                uint32_t* buf = scratch + w * wcap;
                int k = 0;
                for (int s = s0; s < s1; ++s) {
                    const int c = (int)cnt[s];
                    const float* src = vals + pl.seg_base[s];
                    for (int i = lane; i < c; i += 32) buf[k + i] = f2key(src[i]);
                    k += c;
                }
                const int Pw = next_pow2(m);
                for (int i = m + lane; i < Pw; i += 32) buf[i] = 0xffffffffu;
                __syncwarp();
                warp_bitonic_sort(buf, Pw, lane);
                unsigned long long u2 = 0, tie = 0;
                double sum = 0.0;
                for (int i = lane; i < m; i += 32) {
                    uint32_t key = buf[i];
                    sum += fc_value(key2f(key), P.flags.is_log1p);
                    if (i == 0 || buf[i - 1] != key) {
                        long long b = upper_bound_u32(buf, m, key) - i;
                        rank_value(R, key, b, u2, tie);
                    }
                }
                u2 = warp_sum_u64(u2);
                tie = warp_sum_u64(tie);
                sum = warp_sum_f64(sum);
                if (lane == 0) finalize_group(P, R, j, g, m, u2, tie, sum, buf, 1);
                __syncwarp();
            }
            __syncthreads();
            // ---- block tier: one group at a time, sorted in the CTA's global slab
            const int nb = cnt_m[1];
            for (int e = 0; e < nb; ++e) {
                const int g = g0 + blist[e];
                const int s0 = pl.group_seg[g], s1 = pl.group_seg[g + 1];
                int m = 0;
                for (int s = s0; s < s1; ++s) {
                    const int c = (int)cnt[s];
                    const float* src = vals + pl.seg_base[s];
                    for (int i = tid; i < c; i += OVO_THREADS) gA[m + i] = f2key(src[i]);
                    m += c;
                }
                __syncthreads();
                const uint32_t* gk = block_radix_sort(gA, gB, m, hist, aux);
                unsigned long long u2 = 0, tie = 0;
                double sum = 0.0;
                for (int i = tid; i < m; i += OVO_THREADS) {
                    uint32_t key = gk[i];
                    sum += fc_value(key2f(key), P.flags.is_log1p);
                    if (i == 0 || gk[i - 1] != key) {
                        long long b = upper_bound_u32(gk, m, key) - i;
                        rank_value(R, key, b, u2, tie);
                    }
                }
                u2 = block_sum<unsigned long long>(u2, redu);
                tie = block_sum<unsigned long long>(tie, redu);
                sum = block_sum<double>(sum, redd);
                if (tid == 0) finalize_group(P, R, j, g, m, u2, tie, sum, gk, 1);
                __syncthreads();
            }
            __syncthreads();
        }
    }
}

// max over the batch's genes of the control group's non-zero count: sizes the shared control buffer
__global__ void max_ref_nnz_kernel(const uint32_t* __restrict__ ir_cnt, int n_genes, int S, int s0, int s1, int* out) {
    int m = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_genes; j += gridDim.x * blockDim.x) {
        int c = 0;
        for (int s = s0; s < s1; ++s) c += (int)ir_cnt[(long long)j * S + s];
        m = max(m, c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

// ------------------------------------------------------------------------------------------------------
size_t ovo_workspace_bytes(const illico_plan_t* plan, int n_ctas) {
    return (size_t)n_ctas * 4 * (size_t)plan->max_group_size * sizeof(uint32_t);
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <int NT, int MIN_CTAS>
static int launch_ovo_t(OvoParams& P, const illico_plan_t* plan, void* workspace, size_t workspace_bytes, int sms,
                        int max_smem, cudaStream_t stream) {
    constexpr int NW = NT / 32;
    // shared memory: fixed part + control buffer + scratch.  Genes whose control has more non-zeros than ref_cap
    // keep the control in the CTA's global slab.
    const size_t fixed = (size_t)(NW * 256 + RADIX_AUX_WORDS + GROUP_CHUNK + 8) * 4 + 32 * 8 * 2 + 3 * DT_CAP * 8 +
                         (2 * DT_CAP + 1) * 4 + DT_HASH * 8 + 64 + (2 * ST_CAP + 1) * 4;
    const int small_cap = DT_CAP;
    int scratch_words = small_cap * NT;               // thread tier: small_cap keys per thread
    if (scratch_words < 2 * WARP_CAP) scratch_words = 2 * WARP_CAP;  // at least two warp-tier buffers
    // The control buffer is sized to what this batch needs (P.ref_cap holds the measured maximum): shared memory
    // that is not carved out stays L1 cache for the scattered reads of the group slots.
    int ref_cap = (P.ref_cap + 3) & ~3;
    if (ref_cap < 4) ref_cap = 4;
    const int ref_cap_max = env_int("ILLICO_OVO_REF_CAP", scratch_words);
    if (ref_cap > ref_cap_max) ref_cap = ref_cap_max;
    if (ref_cap > scratch_words) ref_cap = scratch_words;  // the ping-pong partner is the scratch area
    const size_t need = fixed + (size_t)(ref_cap + scratch_words) * 4;
    if (need > (size_t)max_smem) { set_error("ovo_kernel needs %zu bytes of shared memory", need); return 1; }
    P.ref_cap = ref_cap; P.small_cap = small_cap; P.scratch_words = scratch_words;
    P.use_search_table = env_int("ILLICO_OVO_SEARCH_TABLE", 1);
    auto kern = ovo_kernel<NT, MIN_CTAS>;
    ILLICO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    int occ = 0;
    ILLICO_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, need));
    if (occ < 1) { set_error("ovo_kernel does not fit: %zu bytes of shared memory", need); return 1; }
    int grid = sms * occ;
    if (grid > P.n_genes) grid = P.n_genes;
    const size_t slab_words = 4 * (size_t)plan->max_group_size;
    if ((size_t)grid * slab_words * 4 > workspace_bytes) {
        grid = (int)(workspace_bytes / (slab_words * 4));
        if (grid < 1) { set_error("rank workspace too small: %zu bytes", workspace_bytes); return 1; }
    }
    P.slab = (uint32_t*)workspace; P.slab_words = (long long)slab_words;
    ILLICO_LAUNCH("ovo_kernel", stream, kern<<<grid, NT, need, stream>>>(P));
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_ovo(const float* ir_vals, const uint32_t* ir_cnt, int n_genes, const illico_plan_t* plan,
               const illico_flags_t* flags, double* results, long long gstride, void* workspace,
               size_t workspace_bytes, const illico_debug_t* dbg, cudaStream_t stream) {
    if (n_genes <= 0) return 0;
    int dev = 0, sms = 0, max_smem = 0;
    ILLICO_CUDA_OK(cudaGetDevice(&dev));
    ILLICO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ILLICO_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    OvoParams P;
    P.ir_vals = ir_vals; P.ir_cnt = ir_cnt; P.n_genes = n_genes; P.plan = *plan; P.flags = *flags;
    P.results = results; P.gstride = gstride;
    P.dbg_u2 = dbg ? (long long*)dbg->u2 : nullptr; P.dbg_tie = dbg ? dbg->tie_sum : nullptr;
    P.dbg_tie_exact = dbg ? (long long*)dbg->tie_exact : nullptr;
    // measured control size (one tiny kernel + a 4-byte read-back; the staging kernel is already in flight)
    {
        int* d_max = reinterpret_cast<int*>(workspace);
        int h_max = 0;
        ILLICO_CUDA_OK(cudaMemsetAsync(d_max, 0, sizeof(int), stream));
        ILLICO_LAUNCH("max_ref_nnz_kernel", stream, max_ref_nnz_kernel<<<(n_genes + 255) / 256, 256, 0, stream>>>(ir_cnt, n_genes, plan->n_segments, plan->ref_seg_begin,
                                                                      plan->ref_seg_end, d_max));
        ILLICO_CUDA_OK(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, stream));
        ILLICO_CUDA_OK(cudaStreamSynchronize(stream));
        P.ref_cap = h_max;
        workspace = reinterpret_cast<char*>(workspace) + 256;
        workspace_bytes -= 256;
    }
    if (env_int("ILLICO_OVO_THREADS", 256) == 256)
        return launch_ovo_t<256, 4>(P, plan, workspace, workspace_bytes, sms, max_smem, stream);
    return launch_ovo_t<512, 2>(P, plan, workspace, workspace_bytes, sms, max_smem, stream);
}

}  // namespace illico

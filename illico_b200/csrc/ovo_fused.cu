// ovo_fused.cu -- dense one-versus-reference in ONE pass over the matrix, without staged lists (sm_100a).
//
// Replaces, for genes whose control has few distinct values (raw or log1p counts), the pair stage_dense + ovo_kernel,
// i.e. illico/ovo/dense_ovo.py:15-137 with illico/utils/ranking.py:52-158 and illico/utils/math.py:64-118,168-221.
//
// Why: writing the non-zero lists (1.2 GB of scattered 32-byte sectors at the K562 shape) costs the staging kernel
// a third of its time although it is a ninth of its bytes, and the rank kernel then reads them back.  When the
// control of a gene has at most DCAP distinct values, everything the test needs about a perturbation is a histogram
// over those values:
//
//     2U_g = sum_t b_t (2 #{ref > v_t} + a_t) + z_g (2 npos + Z_ref)      a_t / b_t = multiplicity of v_t in control / g
//     T_g  = T_ref + sum_t [(a_t + b_t)^3 - (a_t + b_t) - (a_t^3 - a_t)] + (Z^3 - Z)
//
// so the pass that reads the matrix can finish the test itself:
//   1. the control's segments are staged alone (3.6 % of the rows) and `ovo_ctab_kernel` turns each gene's control
//      into a sorted table of (value, multiplicity);
//   2. `ovo_fused_kernel` streams every other row through the TMA ring of stage_dense_tma.cu; lane = gene; non-zeros
//      are compacted into a lane-private shared column (no divergence per element) and, at the end of each group,
//      looked up in the lane's table (values the control lacks are inserted with multiplicity 0) and folded into
//      exact integers (2U without the zero block, the tie term, the expression sum), 24 bytes per (gene, group)
//      written where the result will be;
//   3. `ovo_fused_epilogue_kernel` turns those 24 bytes into (p, U, fold change) in place, in the reference's f64
//      operation order (epilogue.cuh).
// Genes that do not qualify (more than DCAP distinct values, negative values) are flagged and go through the general
// stage + rank path afterwards, in merged runs.
#include "common.cuh"
#include "epilogue.cuh"
#include "tma.cuh"

#include <stdlib.h>

#include <vector>

namespace illico {

// general path (stage.cu, rank_ovo.cu) for the genes the fused path hands back
int launch_stage_dense(const float*, long long, int, int, const illico_plan_t*, float*, uint32_t*, cudaStream_t);
int launch_ovo(const float*, const uint32_t*, int, const illico_plan_t*, const illico_flags_t*, double*, long long, void*,
               size_t, const illico_debug_t*, cudaStream_t);
bool stage_dense_tma_ok(const float* X, long long ld, int gene_lb, int b, const illico_plan_t* plan);
int launch_stage_dense_tma(const float* X, long long ld, int gene_lb, int b, const illico_plan_t* plan, float* ir_vals,
                           uint32_t* ir_cnt, int seg_lo, int seg_hi, cudaStream_t stream, int segs_per_cta = 0);

namespace {

constexpr int FUSED_WARPS = 8;
constexpr int FUSED_LANES = FUSED_WARPS * 32;      // genes per CTA (lane = gene)
constexpr int FUSED_THREADS = FUSED_LANES + 32;    // + one producer warp
constexpr int DCAP_MAX = 16;
constexpr long long PAIR_MAX = 208063;             // above it a pair's tie sum can pass 2^53 (ordered replay needed)
constexpr unsigned long long M_SHIFT = 48;         // record word 0 = 2U (without zeros) | m << 48

// Per-gene control tables, structure-of-arrays over the batch's genes (lane = gene reads are coalesced).
struct Ctab {
    int* D;                    // [b]        distinct control values; -1 = gene takes the general path
    float* key;                // [dcap][b]  ascending
    uint32_t* mult;            // [dcap][b]  multiplicity in the control
    uint32_t* nnz;             // [b]        control non-zeros
    unsigned long long* tie;   // [b]        sum over control runs of a^3 - a
    double* sum;               // [b]        sum of f(x) over the control
    unsigned char* bad;        // [b]        1 = general path (set by the table kernel or by the fused kernel)
    int* n_bad;                // [1]        genes flagged by the table kernel
    // one-versus-rest only (the table comes from a sample of the cells; `mult` then holds the whole gene's histogram)
    uint32_t* r2;              // [dcap][b]  doubled mid-rank of each table value among all cells
    double* fval;              // [dcap][b]  f(value) for the fold change
    uint32_t* r2zero;          // [b]        doubled mid-rank of the zero block
};

size_t ctab_bytes(int b, int dcap) {
    const size_t bb = (size_t)((b + 63) & ~63);
    return bb * (4 + (size_t)dcap * 20 + 4 + 8 + 8 + 4 + 1) + 1024;
}
Ctab ctab_carve(void* ws, int b, int dcap) {
    const size_t bb = (size_t)((b + 63) & ~63);
    char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    Ctab c;
    c.tie = reinterpret_cast<unsigned long long*>(p); p += bb * 8;
    c.sum = reinterpret_cast<double*>(p); p += bb * 8;
    c.fval = reinterpret_cast<double*>(p); p += bb * 8 * dcap;
    c.key = reinterpret_cast<float*>(p); p += bb * 4 * dcap;
    c.r2 = reinterpret_cast<uint32_t*>(p); p += bb * 4 * dcap;
    c.r2zero = reinterpret_cast<uint32_t*>(p); p += bb * 4;
    c.mult = reinterpret_cast<uint32_t*>(p); p += bb * 4 * dcap;
    c.D = reinterpret_cast<int*>(p); p += bb * 4;
    c.nnz = reinterpret_cast<uint32_t*>(p); p += bb * 4;
    c.n_bad = reinterpret_cast<int*>(p); p += 64;
    c.bad = reinterpret_cast<unsigned char*>(p);
    return c;
}

// ---- 1. control tables --------------------------------------------------------------------------------------
// One warp per gene; lane t holds table entry t in registers (dcap <= 32).  Equal values of a 32-value load are
// merged with match.any, their leaders are inserted one after the other.
__global__ void __launch_bounds__(256) ovo_ctab_kernel(const float* __restrict__ ir_vals, const uint32_t* __restrict__ ir_cnt,
                                                       int b, const illico_plan_t pl, int seg_lo, int seg_hi, int is_log1p, int dcap,
                                                       Ctab ct, int bstride) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int S = pl.n_segments;
    for (int j = warp; j < b; j += nwarps) {
        float mykey = 0.0f;
        uint32_t mycnt = 0;
        int D = 0;
        bool bad = false;
        for (int s = seg_lo; s < seg_hi && !bad; ++s) {
            const int c = (int)ir_cnt[(long long)j * S + s];
            const float* src = ir_vals + (long long)j * pl.slot_cap + pl.seg_base[s];
            for (int i0 = 0; i0 < c && !bad; i0 += 32) {
                const bool valid = i0 + lane < c;
                const float v = valid ? src[i0 + lane] : 0.0f;
                if (__any_sync(FULL, valid && !(v > 0.0f))) { bad = true; break; }   // negative or NaN: general path
                const unsigned same = __match_any_sync(FULL, __float_as_uint(v));
                const bool leader = valid && (__ffs(same) - 1 == lane);
                const uint32_t n = (uint32_t)__popc(same);
                unsigned leaders = __ballot_sync(FULL, leader);
                while (leaders) {
                    const int l = __ffs(leaders) - 1;
                    leaders &= leaders - 1;
                    const float vv = __shfl_sync(FULL, v, l);
                    const uint32_t nn = __shfl_sync(FULL, n, l);
                    const unsigned hit = __ballot_sync(FULL, lane < D && mykey == vv);
                    if (hit) {
                        if (lane == __ffs(hit) - 1) mycnt += nn;
                    } else if (D < dcap) {
                        if (lane == D) { mykey = vv; mycnt = nn; }
                        ++D;
                    } else {
                        bad = true;
                        break;
                    }
                }
            }
        }
        if (bad) {
            if (lane == 0) { ct.D[j] = -1; ct.bad[j] = 1; atomicAdd(ct.n_bad, 1); }
            continue;
        }
        // order the entries by value: position = number of smaller keys
        int pos = 0;
        for (int t = 0; t < D; ++t) pos += (__shfl_sync(FULL, mykey, t) < mykey) ? 1 : 0;
        const bool mine = lane < D;
        if (mine) { ct.key[(long long)pos * bstride + j] = mykey; ct.mult[(long long)pos * bstride + j] = mycnt; }
        else if (lane < dcap) { ct.key[(long long)lane * bstride + j] = 0.0f; ct.mult[(long long)lane * bstride + j] = 0u; }  // empty slots
        const uint32_t nnz = warp_sum<uint32_t>(mine ? mycnt : 0u);
        const unsigned long long tie = warp_sum_u64(mine ? (unsigned long long)cube_minus((long long)mycnt) : 0ull);
        const double sum = warp_sum_f64(mine ? (double)mycnt * fc_value(mykey, is_log1p) : 0.0);
        if (lane == 0) { ct.D[j] = D; ct.nnz[j] = nnz; ct.tie[j] = tie; ct.sum[j] = sum; ct.bad[j] = 0; }
    }
}

// ---- 2. the pass over the matrix ----------------------------------------------------------------------------
template <int ROWS, int STAGES, int DCAP, int BUF>
struct FusedLayout {
    static constexpr int ROW_BYTES = FUSED_LANES * 4;
    static constexpr int STAGE_BYTES = ROWS * ROW_BYTES;
    static constexpr int RING_OFF = 0;
    static constexpr int NZ_OFF = STAGES * STAGE_BYTES;            // float [BUF][256]   compacted non-zeros of the group
    static constexpr int KEY_OFF = NZ_OFF + BUF * ROW_BYTES;       // float [DCAP][256]  table values
    static constexpr int HIST_OFF = KEY_OFF + DCAP * ROW_BYTES;    // u16   [DCAP][256]  multiplicity in the current group
    static constexpr int BAR_OFF = HIST_OFF + DCAP * FUSED_LANES * 2;
    static constexpr int BYTES = BAR_OFF + 2 * STAGES * 8;
};

template <int ROWS, int STAGES, int DCAP, int BUF, int MINB, bool OVR>
__global__ void __launch_bounds__(FUSED_THREADS, MINB) ovo_fused_kernel(const float* __restrict__ X, long long ld, int gene_lb, int b,
                                                                  const illico_plan_t pl, int groups_per_cta, int is_log1p,
                                                                  Ctab ct, int bstride, unsigned long long* __restrict__ rec,
                                                                  long long gstride) {
    using L = FusedLayout<ROWS, STAGES, DCAP, BUF>;
    static_assert(32 % ROWS == 0 && BUF > ROWS && DCAP <= DCAP_MAX && (!OVR || DCAP == 12), "layout");
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t smem_a = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bars = smem_a + L::BAR_OFF;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int G = pl.n_groups, ref = OVR ? -1 : pl.ref_group;       // one-versus-rest: every group is streamed
    const int gy0 = blockIdx.y * groups_per_cta, gy1 = min(G, gy0 + groups_per_cta);
    const int p_begin = pl.seg_pos[pl.group_seg[gy0]], p_end = pl.seg_pos[pl.group_seg[gy1]];
    const bool ref_in = ref >= gy0 && ref < gy1;
    const int ref_p0 = ref_in ? pl.seg_pos[pl.group_seg[ref]] : 0;
    const int ref_len = ref_in ? pl.seg_pos[pl.group_seg[ref + 1]] - ref_p0 : 0;
    const int nv = p_end - p_begin - ref_len;                   // rows this CTA streams (the control's are skipped)
    const int g0 = blockIdx.x * FUSED_LANES;
    const uint32_t row_bytes = (uint32_t)min(FUSED_LANES, (b - g0 + 3) & ~3) * 4u;
    if (nv <= 0) return;

    if (t == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(bars + 8 * i, 1);
            mbar_init(bars + 8 * (STAGES + i), FUSED_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (w == FUSED_WARPS) {
        // ---------------- producer warp (as in stage_dense_tma.cu); virtual row i -> position in perm, control skipped
        uint64_t policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        const char* base = reinterpret_cast<const char*>(X + gene_lb + g0);
        const unsigned long long ldb = (unsigned long long)ld * 4ull;
        auto row_of = [&](int i) {
            int p = p_begin + i;
            if (ref_in && p >= ref_p0) p += ref_len;
            return pl.perm[p];
        };
        // (reading the permutation two groups ahead, or the group boundaries one group ahead, measured 5-10 % slower)
        int myrow = (lane < nv) ? row_of(lane) : 0;
        int k = 0;
        for (int i0 = 0; i0 < nv; i0 += 32) {
            const int nxt = (i0 + 32 + lane < nv) ? row_of(i0 + 32 + lane) : 0;
            const int nrows = min(32, nv - i0);
#pragma unroll
            for (int q = 0; q < 32 / ROWS; ++q, ++k) {
                if (q * ROWS >= nrows) break;
                const int slot = k % STAGES;
                const uint32_t full = bars + 8 * slot, empty = bars + 8 * (STAGES + slot);
                mbar_wait(empty, ((k / STAGES) & 1) ^ 1);
                const int rows_here = min(ROWS, nrows - q * ROWS);
                if (lane == 0) mbar_expect_tx(full, (uint32_t)rows_here * row_bytes);
                __syncwarp();
                const int u = lane - q * ROWS;
                if (u >= 0 && u < rows_here)
                    bulk_g2s(smem_a + L::RING_OFF + slot * L::STAGE_BYTES + u * L::ROW_BYTES,
                             base + (unsigned long long)(uint32_t)myrow * ldb, row_bytes, full, policy);
            }
            myrow = nxt;
        }
        return;
    }

    // ---------------- consumer warps: lane = gene.  All shared-memory traffic below uses explicit 32-bit shared
    // addresses (entry q of the lane's column of an array lives at array + q * 1024 + 4 * t).
    const int j = g0 + t;
    const bool in_batch = j < b;
    const uint32_t keys_a = smem_a + L::KEY_OFF + t * 4;     // float [DCAP][256]: control values ascending, then extras
    const uint32_t hist_a = smem_a + L::HIST_OFF + t * 2;    // u16   [DCAP][256]: multiplicity in the current group
    const uint32_t nz_a = smem_a + L::NZ_OFF + t * 4;        // float [BUF][256]:  compacted non-zeros of the group
    auto lds_f = [](uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; };
    auto sts_f = [](uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); };
    auto lds_h = [](uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return (uint32_t)v; };
    auto sts_h = [](uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((uint16_t)v) : "memory"); };
    int Dc = in_batch ? ct.D[j] : -1;                         // control entries (immutable); extras follow in arrival order
    bool bad = Dc < 0;
    if (bad) Dc = 0;
    int D = Dc;
    const float* gkey = ct.key + j;
    const uint32_t* gmult = ct.mult + j;
    // One-versus-rest: the table is global per gene (the records are histograms over its slots).  Slots [Dc, DCAP) are
    // claimed with atomicCAS by whichever CTA first meets a value the sample lacked; every slot that is already
    // taken is cached here.
#pragma unroll
    for (int q = 0; q < DCAP; ++q) {
        float kq = 0.0f;
        if (OVR) {
            if (!bad) kq = __ldcg(gkey + (long long)q * bstride);
            if (kq != 0.0f && q >= D) D = q + 1;
        } else if (q < Dc) {
            kq = gkey[(long long)q * bstride];
        }
        sts_f(keys_a + q * L::ROW_BYTES, kq);
        sts_h(hist_a + q * (FUSED_LANES * 2), 0u);
    }
    uint32_t Hacc[OVR ? DCAP : 1];                            // OVR: this CTA's share of the gene's whole histogram
#pragma unroll
    for (int q = 0; q < (OVR ? DCAP : 1); ++q) Hacc[q] = 0u;
    uint32_t wr = nz_a;                                       // shared address of the lane's next free nz entry

    // looks the buffered non-zeros up in the lane's table and bumps the group's histogram
    auto drain = [&]() {
        const uint32_t mywr = bad ? nz_a : wr;
        const uint32_t maxwr = __reduce_max_sync(FULL, mywr - nz_a);
        for (uint32_t o = 0; o < maxwr; o += L::ROW_BYTES) {
            if (nz_a + o < mywr) {
                const float v = lds_f(nz_a + o);
                // counts: value c sits at entry c - 1 when the control has 1 .. c; otherwise scan for equality
                int q = (int)v - 1;
                if (!(q >= 0 && q < D && lds_f(keys_a + q * L::ROW_BYTES) == v)) {
                    q = 0;
                    while (q < D && lds_f(keys_a + q * L::ROW_BYTES) != v) ++q;
                    if (q == D) {
                        if (OVR) {
                            // not cached: find or claim the value's slot in the gene's global table
                            int slot = -1;
                            if (v > 0.0f) {
                                for (int r = D; r < DCAP && slot < 0; ++r) {
                                    const unsigned old = atomicCAS(reinterpret_cast<unsigned*>(ct.key + (long long)r * bstride + j), 0u,
                                                                   __float_as_uint(v));
                                    const float kv = old ? __uint_as_float(old) : v;
                                    sts_f(keys_a + r * L::ROW_BYTES, kv);
                                    D = r + 1;
                                    if (kv == v) slot = r;
                                }
                            }
                            if (slot < 0) { bad = true; q = 0; }   // a 13th distinct value, a negative one or a NaN: general path
                            else q = slot;
                        } else if (D < DCAP && v > 0.0f) {
                            // a value the control does not have: appended with multiplicity 0
                            sts_f(keys_a + q * L::ROW_BYTES, v);
                            ++D;
                        } else {
                            bad = true;                       // table full, or a negative / NaN value: general path
                            q = 0;
                        }
                    }
                }
                const uint32_t ha = hist_a + q * (FUSED_LANES * 2);
                sts_h(ha, lds_h(ha) + 1u);
            }
        }
        wr = nz_a;
    };
    // folds the group's histogram into exact integers and writes the 24-byte record
    auto close_group = [&](int g) {
        drain();
        if (OVR) {
            // record = the group's histogram over the gene's table, 12 x u16 = 24 bytes
            unsigned long long wds[3] = {0ull, 0ull, 0ull};
#pragma unroll
            for (int q = 0; q < DCAP; ++q) {
                const uint32_t bq = lds_h(hist_a + q * (FUSED_LANES * 2));
                wds[q >> 2] |= (unsigned long long)bq << (16 * (q & 3));
                Hacc[OVR ? q : 0] += bq;
                sts_h(hist_a + q * (FUSED_LANES * 2), 0u);
            }
            if (in_batch && !bad) {
                unsigned long long* o = rec + (long long)g * gstride + (long long)j * 3;
                o[0] = wds[0]; o[1] = wds[1]; o[2] = wds[2];
            }
            return;
        }
        uint32_t a[DCAP];                                     // control multiplicities (L2 resident, coalesced over lanes)
#pragma unroll
        for (int q = 0; q < DCAP; ++q) a[q] = (q < Dc) ? __ldg(gmult + (long long)q * bstride) : 0u;
        unsigned long long u2 = 0, tie = 0, gt = 0;
        double sum = 0.0;
        uint32_t m = 0;
#pragma unroll
        for (int q = DCAP - 1; q >= 0; --q) {
            if (q < Dc) {
                const unsigned long long bq = lds_h(hist_a + q * (FUSED_LANES * 2)), aq = a[q];
                if (bq) {
                    u2 += bq * (2ull * gt + aq);
                    tie += bq * (3ull * aq * aq - 1ull + bq * (3ull * aq + bq));   // (a+b)^3 - (a+b) - (a^3 - a)
                    sum += (double)bq * fc_value(lds_f(keys_a + q * L::ROW_BYTES), is_log1p);
                    m += (uint32_t)bq;
                    sts_h(hist_a + q * (FUSED_LANES * 2), 0u);
                }
                gt += aq;
            }
        }
        for (int q = Dc; q < D; ++q) {                        // extras (rare): a = 0, position among the control's values
            const unsigned long long bq = lds_h(hist_a + q * (FUSED_LANES * 2));
            if (bq) {
                const float v = lds_f(keys_a + q * L::ROW_BYTES);
                unsigned long long gtv = 0;
#pragma unroll
                for (int r = 0; r < DCAP; ++r)
                    if (r < Dc && lds_f(keys_a + r * L::ROW_BYTES) > v) gtv += a[r];
                u2 += bq * 2ull * gtv;
                tie += bq * bq * bq - bq;
                sum += (double)bq * fc_value(v, is_log1p);
                m += (uint32_t)bq;
                sts_h(hist_a + q * (FUSED_LANES * 2), 0u);
            }
        }
        if (in_batch && !bad) {
            unsigned long long* o = rec + (long long)g * gstride + (long long)j * 3;
            o[0] = u2 | ((unsigned long long)m << M_SHIFT);
            o[1] = tie;
            o[2] = (unsigned long long)__double_as_longlong(sum);
        }
    };
    auto group_end_v = [&](int gg) { return pl.seg_pos[pl.group_seg[gg + 1]] - p_begin - ((ref_in && gg > ref) ? ref_len : 0); };
    int g = (gy0 == ref) ? gy0 + 1 : gy0;
    int gend = group_end_v(g);

    auto append = [&](float v) {
        asm volatile("{ .reg .pred p; setp.neu.f32 p, %1, 0f00000000; @p st.shared.f32 [%0], %1; @p add.u32 %0, %0, %2; }"
                     : "+r"(wr)
                     : "f"(v), "n"(L::ROW_BYTES)
                     : "memory");
    };
    const uint32_t ring_a = smem_a + L::RING_OFF + t * 4;
    const uint32_t wr_limit = nz_a + (uint32_t)(BUF - ROWS) * L::ROW_BYTES;
    int k = 0, slot = 0;
    uint32_t parity = 0;
    for (int i = 0; i < nv; i += ROWS, ++k) {
        mbar_wait(bars + 8 * slot, parity);
        const uint32_t src = ring_a + slot * L::STAGE_BYTES;
        if (i + ROWS <= gend) {
            float v[ROWS];
#pragma unroll
            for (int u = 0; u < ROWS; ++u) v[u] = lds_f(src + u * L::ROW_BYTES);
#pragma unroll
            for (int u = 0; u < ROWS; ++u) append(v[u]);
        } else {
            const int nr = min(ROWS, nv - i);
#pragma unroll 1
            for (int u = 0; u < nr; ++u) {
                if (i + u == gend) {                                     // CTA-uniform: the next group starts here
                    close_group(g);
                    ++g;
                    if (g == ref) ++g;
                    gend = group_end_v(g);
                }
                append(lds_f(src + u * L::ROW_BYTES));
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (STAGES + slot));
        if (++slot == STAGES) { slot = 0; parity ^= 1u; }
        if (__any_sync(FULL, wr > wr_limit)) drain();                     // keep room for one more stage
    }
    close_group(g);
    if (bad && in_batch) ct.bad[j] = 1;
    if (OVR && in_batch && !bad) {
#pragma unroll
        for (int q = 0; q < (OVR ? DCAP : 1); ++q)
            if (Hacc[q]) atomicAdd(ct.mult + (long long)q * bstride + j, Hacc[q]);
    }
}

// ---- 3. epilogue: 24-byte integer record -> (p, U, fold change), in place ---------------------------------------
__global__ void __launch_bounds__(256) ovo_fused_epilogue_kernel(int b, const illico_plan_t pl, const illico_flags_t fl, Ctab ct,
                                                                 double* __restrict__ results, long long gstride,
                                                                 long long* dbg_u2, double* dbg_tie, long long* dbg_tie_exact) {
    const int G = pl.n_groups, ref = pl.ref_group;
    const long long total = (long long)G * b;
    const long long n_ref = pl.group_size[ref];
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(idx / b), j = (int)(idx - (long long)g * b);
        if (ct.bad[j]) continue;
        double* o = results + (long long)g * gstride + (long long)j * 3;
        const long long di = (long long)g * b + j;
        const double rsum = fl.group_sums ? fl.group_sums[(long long)ref * b + j] : ct.sum[j];
        const double mean_r = rsum / (double)n_ref;
        if (g == ref) {
            // control row: (1, -1, fold change of the control against itself), as ovo_kernel writes it
            o[0] = 1.0; o[1] = -1.0; o[2] = (mean_r == 0.0) ? INFINITY : mean_r / mean_r;
            if (dbg_u2) dbg_u2[di] = -2;
            if (dbg_tie) dbg_tie[di] = 0.0;
            if (dbg_tie_exact) dbg_tie_exact[di] = 0;
            continue;
        }
        const unsigned long long* r = reinterpret_cast<const unsigned long long*>(o);
        const unsigned long long w0 = r[0], tie_nz = r[1];
        double sum = __longlong_as_double((long long)r[2]);
        const long long m = (long long)(w0 >> M_SHIFT);
        unsigned long long u2 = w0 & ((1ull << M_SHIFT) - 1ull);
        const long long n_t = pl.group_size[g];
        const long long z_t = n_t - m;
        const long long nnz_r = ct.nnz[j], zeros_r = n_ref - nnz_r;      // every control value is positive
        if (fl.group_sums) sum = fl.group_sums[(long long)g * b + j];
        const long long Z = zeros_r + z_t;
        u2 += (unsigned long long)(z_t * (2ll * nnz_r + zeros_r));
        const unsigned long long tie_exact = ct.tie[j] + tie_nz + (unsigned long long)cube_minus(Z);
        const double tie = (double)tie_exact;                            // < 2^53: pairs of at most 208 063 cells
        const double U = (double)u2 / 2.0;
        const double mu = (double)(n_ref * n_t) / 2.0;
        const double cc = fl.use_continuity ? 0.5 : 0.0;
        const double p = compute_pval(n_ref, n_t, n_ref + n_t, fl.tie_correct ? tie : 0.0, U, mu, cc, fl.alternative);
        const double mean_t = sum / (double)n_t;
        o[0] = p; o[1] = U; o[2] = (mean_r == 0.0) ? INFINITY : mean_t / mean_r;
        if (dbg_u2) dbg_u2[di] = (long long)u2;
        if (dbg_tie) dbg_tie[di] = tie;
        if (dbg_tie_exact) dbg_tie_exact[di] = (long long)tie_exact;
    }
}

// ---- one-versus-rest: per-gene ranks from the whole gene's histogram, then the per-group epilogue -----------------
// Thread per gene.  `mult` holds the gene's histogram over its table (summed by the fused pass).  Doubled mid-ranks as
// in ovr_table_kernel (rank_ovr.cu): r2 = 2 lo + c + 1 with the zero block below every (positive) value; tie sum in
// the dense kernels' order (illico/utils/ranking.py:31-47): zero block first, then the runs ascending, sequential f64
// once the exact total reaches 2^53.
constexpr int OVR_DCAP = 12;
__global__ void __launch_bounds__(128) ovr_fused_gene_kernel(int b, const illico_plan_t pl, illico_flags_t fl, Ctab ct, int bstride,
                                                             double* dbg_tie, long long* dbg_tie_exact) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= b || ct.bad[j]) return;
    // the gene's table: every claimed slot (values are positive, 0 = empty), visited in ascending value order
    float key[OVR_DCAP];
    int order[OVR_DCAP];
    int D = 0;
    for (int q = 0; q < OVR_DCAP; ++q) {
        const float kq = ct.key[(long long)q * bstride + j];
        if (kq == 0.0f) break;
        int a = D - 1;
        while (a >= 0 && key[order[a]] > kq) { order[a + 1] = order[a]; --a; }
        key[q] = kq;
        order[a + 1] = q;
        ++D;
    }
    const long long n = pl.n_cells;
    unsigned long long nnz = 0;
    for (int q = 0; q < D; ++q) nnz += ct.mult[(long long)q * bstride + j];
    const long long n0 = n - (long long)nnz;
    unsigned long long lo = (unsigned long long)n0, t_exact = 0;
    double total = 0.0;
    const unsigned long long zterm = (unsigned long long)cube_minus(n0);
    double walk = (double)(long long)zterm;                                  // sequential accumulation, zero block first
    for (int a = 0; a < D; ++a) {
        const int q = order[a];
        const unsigned long long c = ct.mult[(long long)q * bstride + j];
        ct.r2[(long long)q * bstride + j] = (uint32_t)(2ull * lo + c + 1ull);
        const double f = fc_value(key[q], fl.is_log1p);
        ct.fval[(long long)q * bstride + j] = f;
        total += (double)c * f;
        const unsigned long long t3 = (unsigned long long)cube_minus((long long)c);
        t_exact += t3;
        walk += (double)(long long)t3;
        lo += c;
    }
    for (int q = D; q < OVR_DCAP; ++q) { ct.r2[(long long)q * bstride + j] = 0u; ct.fval[(long long)q * bstride + j] = 0.0; }
    const double tie = ((double)t_exact + (double)zterm >= TWO53) ? walk : (double)(t_exact + zterm);
    ct.r2zero[j] = (uint32_t)(n0 + 1);
    ct.sum[j] = total;
    ct.tie[j] = (unsigned long long)__double_as_longlong(tie);
    if (dbg_tie) dbg_tie[j] = tie;
    if (dbg_tie_exact) dbg_tie_exact[j] = (long long)(t_exact + zterm);
}

// Thread per (group, gene): 24-byte histogram record -> (p, U, fold change) in place (illico/ovr/dense_ovr.py:57-78).
__global__ void __launch_bounds__(256) ovr_fused_epilogue_kernel(int b, const illico_plan_t pl, const illico_flags_t fl, Ctab ct,
                                                                 int bstride, double* __restrict__ results, long long gstride,
                                                                 long long* dbg_u2) {
    const int g = blockIdx.y;
    const long long n = pl.n_cells, n_t = pl.group_size[g], n_r = n - n_t;
    const double cc = fl.use_continuity ? 0.5 : 0.0;
    const double mu = (double)(n_r * n_t) / 2.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < b; j += gridDim.x * blockDim.x) {
        if (ct.bad[j]) continue;
        double* o = results + (long long)g * gstride + (long long)j * 3;
        const unsigned long long* r = reinterpret_cast<const unsigned long long*>(o);
        const unsigned long long wds[3] = {r[0], r[1], r[2]};
        unsigned long long R2 = 0, nnz_g = 0;
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < OVR_DCAP; ++q) {
            const unsigned long long bq = (wds[q >> 2] >> (16 * (q & 3))) & 0xffffull;
            if (bq) {
                R2 += bq * ct.r2[(long long)q * bstride + j];
                sum += (double)bq * ct.fval[(long long)q * bstride + j];
                nnz_g += bq;
            }
        }
        R2 += (unsigned long long)(n_t - (long long)nnz_g) * ct.r2zero[j];
        const long long u2 = 2 * n_r * n_t + n_t * (n_t + 1) - (long long)R2;
        const double U = (double)u2 / 2.0;
        const double tie = __longlong_as_double((long long)ct.tie[j]);
        const double p = compute_pval(n_r, n_t, n, fl.tie_correct ? tie : 0.0, U, mu, cc, fl.alternative);
        const double total = ct.sum[j];
        const double mu_t = sum / (double)n_t;
        const double mu_r = (total - sum) / (double)(n - n_t);
        o[0] = p; o[1] = U; o[2] = (mu_r == 0.0) ? INFINITY : mu_t / mu_r;
        if (dbg_u2) dbg_u2[(long long)g * b + j] = u2;
    }
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

thread_local float g_last_fused_ms = -1.0f;

template <int ROWS, int STAGES, int DCAP, int BUF, int MINB, bool OVR>
int launch_fused_t(const float* X, long long ld, int gene_lb, int b, const illico_plan_t* plan, int gpc, int is_log1p, Ctab ct,
                   int bstride, double* results, long long gstride, cudaStream_t stream) {
    using L = FusedLayout<ROWS, STAGES, DCAP, BUF>;
    auto kern = ovo_fused_kernel<ROWS, STAGES, DCAP, BUF, MINB, OVR>;
    ILLICO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
    const dim3 grid((unsigned)((b + FUSED_LANES - 1) / FUSED_LANES), (unsigned)((plan->n_groups + gpc - 1) / gpc));
    kern<<<grid, FUSED_THREADS, L::BYTES, stream>>>(X, ld, gene_lb, b, *plan, gpc, is_log1p, ct, bstride,
                                                    reinterpret_cast<unsigned long long*>(results), gstride);
    count_launch();
    ILLICO_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

float ovo_fused_last_ms() { return g_last_fused_ms; }

size_t ovo_fused_workspace_bytes(int b) { return ctab_bytes(b, DCAP_MAX); }

// 0 = done, 1 = error, -1 = not applicable (the caller runs the general path on the whole batch)
int launch_ovo_dense_fused(const float* X, long long ld, int gene_lb, int b, const illico_plan_t* plan,
                           const illico_flags_t* flags, const illico_batch_buffers_t* buf, double* results, long long gstride,
                           const illico_debug_t* dbg, cudaStream_t stream) {
    if (env_int("ILLICO_OVO_FUSED", 1) == 0 || b <= 0) return -1;
    if (!stage_dense_tma_ok(X, ld, gene_lb, b, plan)) return -1;
    if (plan->max_group_size >= 65536) return -1;                                     // 16-bit histograms and m
    if ((long long)plan->ref_group_size + plan->max_group_size > PAIR_MAX) return -1; // tie sums stay below 2^53
    if (plan->n_groups < 2 || plan->n_groups > 65535 * 16) return -1;
    const int cfg = env_int("ILLICO_OVO_FUSED_CFG", 0);
    const int dcap = 10;
    if (buf->workspace_bytes < ctab_bytes(b, dcap)) return -1;
    const int bstride = (b + 63) & ~63;
    Ctab ct = ctab_carve(buf->workspace, b, dcap);

    // 1. the control group alone -> its per-gene tables
    ILLICO_CUDA_OK(cudaMemsetAsync(ct.n_bad, 0, sizeof(int), stream));
    {
        const int rc = launch_stage_dense_tma(X, ld, gene_lb, b, plan, buf->ir_vals, buf->ir_cnt, plan->ref_seg_begin,
                                              plan->ref_seg_end, stream, /*segs_per_cta=*/1);
        if (rc != 0) return rc;
    }
    {
        int blocks = (b + 7) / 8;
        if (blocks > 148 * 8) blocks = 148 * 8;
        ovo_ctab_kernel<<<blocks, 256, 0, stream>>>(buf->ir_vals, buf->ir_cnt, b, *plan, plan->ref_seg_begin, plan->ref_seg_end,
                                                    flags->is_log1p, dcap, ct, bstride);
        count_launch();
        ILLICO_CUDA_OK(cudaGetLastError());
    }
    int n_bad = 0;
    ILLICO_CUDA_OK(cudaMemcpyAsync(&n_bad, ct.n_bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ILLICO_CUDA_OK(cudaStreamSynchronize(stream));
    if (2 * n_bad > b) return -1;   // mostly continuous data: the general path is the right one for this batch

    // 2. the pass over the matrix, 3. the epilogue
    long long avg = plan->n_cells / plan->n_groups;
    if (avg < 1) avg = 1;
    int gpc = (int)(env_int("ILLICO_OVO_FUSED_ROWS", 1536) / avg);
    if (gpc < 1) gpc = 1;
    if ((plan->n_groups + gpc - 1) / gpc > 65535) gpc = (plan->n_groups + 65534) / 65535;
    const bool timed = env_int("ILLICO_PROFILE", 0) != 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        ILLICO_CUDA_OK(cudaEventCreate(&e0));
        ILLICO_CUDA_OK(cudaEventCreate(&e1));
        ILLICO_CUDA_OK(cudaEventRecord(e0, stream));
    }
    int rc;
#define FUSED_ARGS X, ld, gene_lb, b, plan, gpc, flags->is_log1p, ct, bstride, results, gstride, stream
    // ring / buffer shapes measured at the K562 shape (profiles/README.md): 5 stages of 8 rows, a 16-entry group
    // buffer and 3 CTAs per SM is the fastest; 4 CTAs per SM with a 3-stage ring is within 3 % of it
    switch (cfg) {
        case 1: rc = launch_fused_t<8, 3, 10, 32, 3, false>(FUSED_ARGS); break;
        case 2: rc = launch_fused_t<8, 3, 10, 16, 4, false>(FUSED_ARGS); break;
        default: rc = launch_fused_t<8, 5, 10, 16, 3, false>(FUSED_ARGS); break;
    }
#undef FUSED_ARGS
    if (rc) return rc;
    if (timed) ILLICO_CUDA_OK(cudaEventRecord(e1, stream));
    {
        long long total = (long long)plan->n_groups * b;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        ovo_fused_epilogue_kernel<<<(unsigned)blocks, 256, 0, stream>>>(b, *plan, *flags, ct, results, gstride,
                                                                        dbg ? (long long*)dbg->u2 : nullptr,
                                                                        dbg ? dbg->tie_sum : nullptr,
                                                                        dbg ? (long long*)dbg->tie_exact : nullptr);
        count_launch();
        ILLICO_CUDA_OK(cudaGetLastError());
    }
    // 4. genes handed back: general path, in merged runs
    std::vector<unsigned char> bad((size_t)b);
    ILLICO_CUDA_OK(cudaMemcpyAsync(bad.data(), ct.bad, (size_t)b, cudaMemcpyDeviceToHost, stream));
    ILLICO_CUDA_OK(cudaStreamSynchronize(stream));
    if (timed) {
        ILLICO_CUDA_OK(cudaEventElapsedTime(&g_last_fused_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    int first = -1, last = -1, count = 0;
    for (int j = 0; j < b; ++j)
        if (bad[j]) { if (first < 0) first = j; last = j; ++count; }
    if (count == 0) return 0;
    if (dbg || flags->group_sums) return -1;   // their [G, b] side arrays are indexed by the whole batch: redo it all
    const int gap = env_int("ILLICO_OVO_FUSED_GAP", 128);   // a launch costs about as much as this many genes
    int lb = first;
    while (lb <= last) {
        int ub = lb + 1, j = lb + 1;
        while (j <= last) {
            if (bad[j]) { ub = j + 1; ++j; }
            else if (j - ub < gap) ++j;
            else break;
        }
        // genes [lb, ub) of the batch (good genes inside a merged run are simply recomputed)
        if (launch_stage_dense(X, ld, gene_lb + lb, ub - lb, plan, buf->ir_vals, buf->ir_cnt, stream)) return 1;
        if (launch_ovo(buf->ir_vals, buf->ir_cnt, ub - lb, plan, flags, results + (long long)lb * 3, gstride, buf->workspace,
                       buf->workspace_bytes, nullptr, stream)) return 1;
        lb = ub;
        while (lb <= last && !bad[lb]) ++lb;
    }
    return 0;
}


int launch_ovr(const float*, const uint32_t*, int, const illico_plan_t*, const illico_flags_t*, double*, long long, void*,
               size_t, const illico_debug_t*, cudaStream_t);

// One-versus-rest through the same pass: the per-gene table comes from a sample of the cells (the first segments),
// count tables are extended to 1 .. 12, the pass writes each group's histogram over the table and sums the gene's
// whole histogram, `ovr_fused_gene_kernel` turns that into mid-ranks and the tie sum, the epilogue finishes the tests.
// 0 = done, 1 = error, -1 = not applicable.
int launch_ovr_dense_fused(const float* X, long long ld, int gene_lb, int b, const illico_plan_t* plan,
                           const illico_flags_t* flags, const illico_batch_buffers_t* buf, double* results, long long gstride,
                           const illico_debug_t* dbg, cudaStream_t stream) {
    if (env_int("ILLICO_OVR_FUSED", 1) == 0 || b <= 0) return -1;
    if (!stage_dense_tma_ok(X, ld, gene_lb, b, plan)) return -1;
    if (plan->max_group_size >= 65536 || plan->n_groups < 2 || plan->n_groups > 65535) return -1;
    if (flags->group_sums) return -1;
    if (buf->workspace_bytes < ctab_bytes(b, OVR_DCAP)) return -1;
    const int bstride = (b + 63) & ~63;
    Ctab ct = ctab_carve(buf->workspace, b, OVR_DCAP);

    // 1. tables from a sample: the first segments, about 16k cells
    long long avg = plan->n_cells / plan->n_segments;
    if (avg < 1) avg = 1;
    int seg_hi = (int)((env_int("ILLICO_OVR_FUSED_SAMPLE", 16384) + avg - 1) / avg);
    if (seg_hi < 1) seg_hi = 1;
    if (seg_hi > plan->n_segments) seg_hi = plan->n_segments;
    ILLICO_CUDA_OK(cudaMemsetAsync(ct.n_bad, 0, sizeof(int), stream));
    {
        const int rc = launch_stage_dense_tma(X, ld, gene_lb, b, plan, buf->ir_vals, buf->ir_cnt, 0, seg_hi, stream, 1);
        if (rc != 0) return rc;
    }
    {
        int blocks = (b + 7) / 8;
        if (blocks > 148 * 8) blocks = 148 * 8;
        ovo_ctab_kernel<<<blocks, 256, 0, stream>>>(buf->ir_vals, buf->ir_cnt, b, *plan, 0, seg_hi, flags->is_log1p, OVR_DCAP, ct,
                                                    bstride);
        count_launch();
        ILLICO_CUDA_OK(cudaGetLastError());
    }
    ILLICO_CUDA_OK(cudaMemsetAsync(ct.mult, 0, (size_t)bstride * OVR_DCAP * sizeof(uint32_t), stream));  // whole-gene histogram
    int n_bad = 0;
    ILLICO_CUDA_OK(cudaMemcpyAsync(&n_bad, ct.n_bad, sizeof(int), cudaMemcpyDeviceToHost, stream));
    ILLICO_CUDA_OK(cudaStreamSynchronize(stream));
    if (2 * n_bad > b) return -1;

    // 2. the pass, 3. per-gene ranks, 4. the epilogue
    long long avg_g = plan->n_cells / plan->n_groups;
    if (avg_g < 1) avg_g = 1;
    int gpc = (int)(env_int("ILLICO_OVO_FUSED_ROWS", 1536) / avg_g);
    if (gpc < 1) gpc = 1;
    const bool timed = env_int("ILLICO_PROFILE", 0) != 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timed) {
        ILLICO_CUDA_OK(cudaEventCreate(&e0));
        ILLICO_CUDA_OK(cudaEventCreate(&e1));
        ILLICO_CUDA_OK(cudaEventRecord(e0, stream));
    }
    if (launch_fused_t<8, 4, 12, 16, 3, true>(X, ld, gene_lb, b, plan, gpc, flags->is_log1p, ct, bstride, results, gstride, stream))
        return 1;
    if (timed) ILLICO_CUDA_OK(cudaEventRecord(e1, stream));
    ovr_fused_gene_kernel<<<(b + 127) / 128, 128, 0, stream>>>(b, *plan, *flags, ct, bstride, dbg ? dbg->tie_sum : nullptr,
                                                               dbg ? (long long*)dbg->tie_exact : nullptr);
    count_launch();
    ILLICO_CUDA_OK(cudaGetLastError());
    {
        int gx = (b + 255) / 256;
        if (gx > 64) gx = 64;
        ovr_fused_epilogue_kernel<<<dim3((unsigned)gx, (unsigned)plan->n_groups), 256, 0, stream>>>(
            b, *plan, *flags, ct, bstride, results, gstride, dbg ? (long long*)dbg->u2 : nullptr);
        count_launch();
        ILLICO_CUDA_OK(cudaGetLastError());
    }
    std::vector<unsigned char> bad((size_t)b);
    ILLICO_CUDA_OK(cudaMemcpyAsync(bad.data(), ct.bad, (size_t)b, cudaMemcpyDeviceToHost, stream));
    ILLICO_CUDA_OK(cudaStreamSynchronize(stream));
    if (timed) {
        ILLICO_CUDA_OK(cudaEventElapsedTime(&g_last_fused_ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    int first = -1, last = -1, count = 0;
    for (int j = 0; j < b; ++j)
        if (bad[j]) { if (first < 0) first = j; last = j; ++count; }
    if (count == 0) return 0;
    if (dbg) return -1;   // the debug arrays are indexed by the whole batch: redo it all through the general path
    const int gap = env_int("ILLICO_OVO_FUSED_GAP", 128);
    int lb = first;
    while (lb <= last) {
        int ub = lb + 1, j = lb + 1;
        while (j <= last) {
            if (bad[j]) { ub = j + 1; ++j; }
            else if (j - ub < gap) ++j;
            else break;
        }
        if (launch_stage_dense(X, ld, gene_lb + lb, ub - lb, plan, buf->ir_vals, buf->ir_cnt, stream)) return 1;
        if (launch_ovr(buf->ir_vals, buf->ir_cnt, ub - lb, plan, flags, results + (long long)lb * 3, gstride, buf->workspace,
                       buf->workspace_bytes, nullptr, stream)) return 1;
        lb = ub;
        while (lb <= last && !bad[lb]) ++lb;
    }
    return 0;
}

}  // namespace illico

// Single-query exact search, HBM-bound: an FP32 SIMT scan streams the sample matrix once
// (4*N*ld bytes), the last CTA to finish picks the candidates whose FP32 score is within a
// rigorous rounding bound of the k-th best, and a second small kernel re-ranks those few rows
// with the canonical FP64 sums and orders them under the reference rule.  Results are identical
// to the FP64 scan (morna_angular_distances + morna_select_topk); ties wider than the candidate
// list fall back to it.  Replaces the per-query loop of exact_search_nn (morna.py:697-712).
#include <math.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace morna {

constexpr int kS1Threads = 512, kS1Warps = kS1Threads / 32;
constexpr int kS1List = 2048;          // survivors of the pivot kept in shared memory
constexpr int kS1Cand = 1024;          // candidates handed to the FP64 re-rank

struct SingleWs {          // device-side control block at the start of the workspace
    unsigned int scan_ticket, rerank_ticket;
    int cand_count, fallback;
    float margin, pad;
};

__device__ __forceinline__ uint32_t fkey(float v) {
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ uint32_t warp_kth_largest32(const uint32_t (&key)[32], int kk) {
    uint32_t best = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t trial = best | (1u << bit);
        int c = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) c += key[j] >= trial;
        c = __reduce_add_sync(kFull, c);
        if (c >= kk) best = trial;
    }
    return best;
}

// Candidate selection over all N scores, run by ONE CTA (the last to finish the scan).
// Keeps every row whose score is >= (k-th largest score) - margin.
__device__ void select_candidates(const float *__restrict__ score, int64_t n, int k, float margin,
                                  int32_t *__restrict__ cand, SingleWs *ctl, uint32_t *s_key, int *s_idx,
                                  int *s_small) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int &s_count = s_small[0];
    int &s_out = s_small[1];
    uint32_t &s_pivot = reinterpret_cast<uint32_t &>(s_small[2]);
    uint32_t &s_best = reinterpret_cast<uint32_t &>(s_small[3]);
    int *s_warp = s_small + 4;         // [kS1Warps]
    const int kk = (int)min((int64_t)k, n);
    if (tid == 0) { s_count = 0; s_out = 0; s_pivot = 0; s_best = 0; }
    __syncthreads();
    if (kk == 0) { if (tid == 0) { ctl->cand_count = 0; ctl->fallback = 0; } return; }
    // pivot: r-th largest of 1024 strided samples, r sized so that >= k rows survive with margin
    if (n > kS1List) {
        if (warp == 0) {
            uint32_t sk[32];
            const int64_t stride = n >> 10;
#pragma unroll
            for (int j = 0; j < 32; ++j) sk[j] = fkey(__ldcg(score + (int64_t)(j * 32 + lane) * stride));
            const float want = (float)kk * 1024.0f / (float)n;
            const int r = (int)(want + 4.0f * sqrtf(want) + 4.0f);
            const uint32_t pv = r <= 1024 ? warp_kth_largest32(sk, r) : 0u;
            if (lane == 0) s_pivot = pv;
        }
        __syncthreads();
    }
    const uint32_t pivot = s_pivot;
    // stream every score once, eight independent loads in flight per thread
    for (int64_t i0 = 0; i0 < n; i0 += 8 * kS1Threads) {
        uint32_t key[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t i = i0 + u * kS1Threads + tid;
            key[u] = i < n ? fkey(__ldcg(score + i)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t i = i0 + u * kS1Threads + tid;
            const bool keep = i < n && key[u] >= pivot;
            const unsigned mask = __ballot_sync(kFull, keep);
            if (mask) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_count, __popc(mask));
                base = __shfl_sync(kFull, base, 0);
                const int at = base + __popc(mask & ((1u << lane) - 1u));
                if (keep && at < kS1List) { s_key[at] = key[u]; s_idx[at] = (int)i; }
            }
        }
    }
    __syncthreads();
    const int m = s_count;
    if (m < kk || m > kS1List) {       // unlucky pivot or heavy ties: the FP64 scan answers
        if (tid == 0) { ctl->cand_count = 0; ctl->fallback = 1; }
        return;
    }
    // exact kk-th largest of the m survivors
    uint32_t mine[kS1List / kS1Threads];
#pragma unroll
    for (int j = 0; j < kS1List / kS1Threads; ++j) {
        const int at = j * kS1Threads + tid;
        mine[j] = at < m ? s_key[at] : 0u;
    }
    uint32_t best = 0;
    if (m <= 1024) {                   // one warp, no block barriers
        if (warp == 0) {
            uint32_t sk[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { const int at = j * 32 + lane; sk[j] = at < m ? s_key[at] : 0u; }
            const uint32_t b = warp_kth_largest32(sk, kk);
            if (lane == 0) s_best = b;
        }
        __syncthreads();
        best = s_best;
    } else {
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t trial = best | (1u << bit);
            int c = 0;
#pragma unroll
            for (int j = 0; j < kS1List / kS1Threads; ++j) c += mine[j] >= trial;
            c = __reduce_add_sync(kFull, c);
            if (lane == 0) s_warp[warp] = c;
            __syncthreads();
            int tot = 0;
#pragma unroll
            for (int w = 0; w < kS1Warps; ++w) tot += s_warp[w];
            __syncthreads();
            if (tot >= kk) best = trial;
        }
    }
    const float cut = kk >= k ? __fsub_rd(fkey_inv(best), margin) : -INFINITY;
    const uint32_t cut_key = cut == -INFINITY ? 0u : fkey(cut);
    if (cut_key < pivot) {             // the cut fell below the pivot: survivors may miss candidates
        if (tid == 0) { ctl->cand_count = 0; ctl->fallback = 1; }
        return;
    }
#pragma unroll
    for (int j = 0; j < kS1List / kS1Threads; ++j) {
        const int at = j * kS1Threads + tid;
        const bool keep = at < m && mine[j] >= cut_key;
        const unsigned mask = __ballot_sync(kFull, keep);
        if (mask) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_out, __popc(mask));
            base = __shfl_sync(kFull, base, 0);
            const int o = base + __popc(mask & ((1u << lane) - 1u));
            if (keep && o < kS1Cand) cand[o] = s_idx[at];
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int kept = s_out;
        ctl->cand_count = kept <= kS1Cand ? kept : 0;
        ctl->fallback = kept <= kS1Cand ? 0 : 1;
    }
}

// FP32 scan: one warp per row, four 16-byte loads in flight per lane, the query as floats in
// shared memory.  score[i] = (sum_j v[i][j] * q32[j]) / sqrt(pp[i]); zero rows score 0.
__global__ void __launch_bounds__(kS1Threads)
scan32_select_kernel(const float *__restrict__ vectors, const double *__restrict__ pp, int64_t n, int64_t ld,
                     int32_t dim, const double *__restrict__ query, int32_t k, float eps_rel,
                     float *__restrict__ score, int32_t *__restrict__ cand, SingleWs *ctl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *qs = reinterpret_cast<float *>(smem_raw);                   // [ld] floats
    uint32_t *s_key = reinterpret_cast<uint32_t *>(smem_raw);          // reused by the last CTA
    int *s_idx = reinterpret_cast<int *>(s_key + kS1List);
    __shared__ int s_small[4 + kS1Warps];
    __shared__ double s_norm[kS1Warps];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunks = (int)(ld >> 2);
    double part = 0.0;
    for (int c = tid; c < ld; c += kS1Threads) {
        const double v = c < dim ? query[c] : 0.0;
        qs[c] = (float)v;
        part = fma(v, v, part);
    }
    part = warp_sum(part);
    if (lane == 0) s_norm[warp] = part;
    __syncthreads();
    const int64_t warps_total = (int64_t)gridDim.x * kS1Warps;
    for (int64_t row = (int64_t)blockIdx.x * kS1Warps + warp; row < n; row += warps_total) {
        const float4 *src = reinterpret_cast<const float4 *>(vectors + row * ld);
        float acc = 0.f;
        int c = lane;
        for (; c + 96 < chunks; c += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_stream_f4(src + c + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 qv = *reinterpret_cast<const float4 *>(qs + 4 * (c + 32 * u));
                acc = fmaf(v[u].x, qv.x, acc); acc = fmaf(v[u].y, qv.y, acc);
                acc = fmaf(v[u].z, qv.z, acc); acc = fmaf(v[u].w, qv.w, acc);
            }
        }
        for (; c < chunks; c += 32) {
            const float4 v = ldg_stream_f4(src + c);
            const float4 qv = *reinterpret_cast<const float4 *>(qs + 4 * c);
            acc = fmaf(v.x, qv.x, acc); acc = fmaf(v.y, qv.y, acc);
            acc = fmaf(v.z, qv.z, acc); acc = fmaf(v.w, qv.w, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
        if (lane == 0) {
            const float p = (float)pp[row];
            score[row] = p > 0.f ? acc * (1.0f / sqrtf(p)) : 0.f;
        }
    }
    // last CTA to arrive selects the candidates from all scores
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&ctl->scan_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double qq = 0.0;
#pragma unroll
    for (int w = 0; w < kS1Warps; ++w) qq += s_norm[w];
    // |fp32 score - cos * |q|| <= eps_rel * |q|  (rounding of q to fp32, the fp32 sums, the scaling)
    const float margin = __fmul_ru(2.0f * eps_rel, __double2float_ru(sqrt(qq)));
    if (tid == 0) { ctl->margin = margin; ctl->scan_ticket = 0; }
    __syncthreads();                     // qs is dead from here on: its storage holds the survivor list
    select_candidates(score, n, k, margin, cand, ctl, s_key, s_idx, s_small);
}

// ---------------------------------------------------------------- bulk-copy staged scan
// Same arithmetic as scan32_select_kernel, but the rows reach the SM as large contiguous
// cp.async.bulk copies (kTmaRows rows = up to 96 KB per copy, double buffered) issued by one
// producer thread, so DRAM sees a few hundred long sequential streams instead of one 512-byte
// stream per warp; eight consumer warps read the staged rows from shared memory.
constexpr int kTmaStageBytes = 96 * 1024, kTmaStages = 2, kTmaConsumers = 8;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(kS1Threads, 1)
scan32_tma_kernel(const float *__restrict__ vectors, const double *__restrict__ pp, int64_t n, int64_t ld,
                  int32_t dim, int32_t rows_per_stage, const double *__restrict__ query, int32_t k, float eps_rel,
                  float *__restrict__ score, int32_t *__restrict__ cand, SingleWs *ctl) {
    using namespace tc;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t stage_bytes = (size_t)rows_per_stage * ld * sizeof(float);
    auto stage_at = [&](int s) { return reinterpret_cast<float *>(smem_raw + (size_t)s * stage_bytes); };
    float *qs = reinterpret_cast<float *>(smem_raw + kTmaStages * stage_bytes);            // [ld]
    uint64_t *bars = reinterpret_cast<uint64_t *>(qs + ld);                                  // full[2], empty[2]
    uint32_t *s_key = reinterpret_cast<uint32_t *>(smem_raw);                                // reused by the last CTA
    int *s_idx = reinterpret_cast<int *>(s_key + kS1List);
    __shared__ int s_small[4 + kS1Warps];
    __shared__ double s_norm[kS1Warps];
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunks = (int)(ld >> 2);
    auto full_bar = [&](int s) { return smem_u32(bars + s); };
    auto empty_bar = [&](int s) { return smem_u32(bars + kTmaStages + s); };
    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kTmaConsumers); }
        fence_barrier_init();
    }
    double part = 0.0;
    for (int c = tid; c < ld; c += kS1Threads) {
        const double v = c < dim ? query[c] : 0.0;
        qs[c] = (float)v;
        part = fma(v, v, part);
    }
    part = warp_sum(part);
    if (lane == 0) s_norm[warp] = part;
    __syncthreads();

    const int64_t n_blocks = (n + rows_per_stage - 1) / rows_per_stage;
    if (warp == 0) {
        if (lane == 0) {                 // producer
            int stage = 0; uint32_t phase = 0;
            for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
                const int64_t row0 = b * rows_per_stage;
                const int64_t rows = min((int64_t)rows_per_stage, n - row0);
                const uint32_t bytes = (uint32_t)(rows * ld * sizeof(float));
                mbar_wait(empty_bar(stage), phase ^ 1);
                mbar_expect_tx(full_bar(stage), bytes);
                bulk_load(smem_u32(stage_at(stage)), vectors + row0 * ld, bytes, full_bar(stage));
                if (++stage == kTmaStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp <= kTmaConsumers) {
        const int cw = warp - 1;
        int stage = 0; uint32_t phase = 0;
        for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
            const int64_t row0 = b * rows_per_stage;
            const int rows = (int)min((int64_t)rows_per_stage, n - row0);
            mbar_wait(full_bar(stage), phase);
            for (int r = cw; r < rows; r += kTmaConsumers) {
                const float4 *src = reinterpret_cast<const float4 *>(stage_at(stage) + (int64_t)r * ld);
                float acc = 0.f;
#pragma unroll 4
                for (int c = lane; c < chunks; c += 32) {      // same lane/chunk order as the LDG scan
                    const float4 v = src[c];
                    const float4 qv = *reinterpret_cast<const float4 *>(qs + 4 * c);
                    acc = fmaf(v.x, qv.x, acc); acc = fmaf(v.y, qv.y, acc);
                    acc = fmaf(v.z, qv.z, acc); acc = fmaf(v.w, qv.w, acc);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
                if (lane == 0) {
                    const float p = (float)pp[row0 + r];
                    score[row0 + r] = p > 0.f ? acc * (1.0f / sqrtf(p)) : 0.f;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar(stage));
            if (++stage == kTmaStages) { stage = 0; phase ^= 1; }
        }
    }
    // last CTA to arrive selects the candidates from all scores
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&ctl->scan_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double qq = 0.0;
#pragma unroll
    for (int w = 0; w < kS1Warps; ++w) qq += s_norm[w];
    const float margin = __fmul_ru(2.0f * eps_rel, __double2float_ru(sqrt(qq)));
    if (tid == 0) { ctl->margin = margin; ctl->scan_ticket = 0; }
    __syncthreads();                     // the staging buffers are dead: they now hold the survivor list
    select_candidates(score, n, k, margin, cand, ctl, s_key, s_idx, s_small);
}

// FP64 re-rank of the candidate rows (canonical sums), one warp per candidate; the last CTA sorts
// under the reference order and writes the first k.
__global__ void __launch_bounds__(kS1Threads)
rerank_single_kernel(const float *__restrict__ vectors, const double *__restrict__ pp, int64_t ld, int32_t dim,
                     int32_t id_base, const double *__restrict__ query, const int32_t *__restrict__ cand,
                     double *__restrict__ cand_dist, SingleWs *ctl, int32_t k, int32_t *__restrict__ out_ids,
                     double *__restrict__ out_dist, int32_t *__restrict__ fallback_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *qs = reinterpret_cast<double *>(smem_raw);                 // [ld]
    double *sd = qs;                                                   // the sort reuses the storage
    int *si = reinterpret_cast<int *>(sd + kS1Cand);
    __shared__ double s_qq;
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int count = ctl->cand_count;
    const int chunks = (int)(ld >> 2);
    if (ctl->fallback) {
        if (blockIdx.x == 0 && tid == 0) { *fallback_out = 1; ctl->rerank_ticket = 0; }
        return;
    }
    for (int c = tid; c < ld; c += kS1Threads) qs[c] = c < dim ? query[c] : 0.0;
    __syncthreads();
    if (warp == 0) {                     // qq with the canonical tree
        double acc = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            const double2 a = *reinterpret_cast<const double2 *>(qs + 4 * c);
            const double2 b = *reinterpret_cast<const double2 *>(qs + 4 * c + 2);
            acc = fma(a.x, a.x, acc); acc = fma(a.y, a.y, acc);
            acc = fma(b.x, b.x, acc); acc = fma(b.y, b.y, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) s_qq = acc;
    }
    __syncthreads();
    const double qqv = s_qq;
    for (int i = blockIdx.x * kS1Warps + warp; i < count; i += gridDim.x * kS1Warps) {
        const int64_t row = cand[i];
        const float4 *src = reinterpret_cast<const float4 *>(vectors + row * ld);
        double acc = 0.0;
        int c = lane;
        for (; c + 96 < chunks; c += 128) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(src + c + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double *qp = qs + 4 * (c + 32 * u);
                const double2 qa = *reinterpret_cast<const double2 *>(qp);
                const double2 qb = *reinterpret_cast<const double2 *>(qp + 2);
                acc = fma((double)v[u].x, qa.x, acc); acc = fma((double)v[u].y, qa.y, acc);
                acc = fma((double)v[u].z, qb.x, acc); acc = fma((double)v[u].w, qb.y, acc);
            }
        }
        for (; c < chunks; c += 32) {
            const float4 v = __ldg(src + c);
            const double2 qa = *reinterpret_cast<const double2 *>(qs + 4 * c);
            const double2 qb = *reinterpret_cast<const double2 *>(qs + 4 * c + 2);
            acc = fma((double)v.x, qa.x, acc); acc = fma((double)v.y, qa.y, acc);
            acc = fma((double)v.z, qb.x, acc); acc = fma((double)v.w, qb.y, acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) cand_dist[i] = angular_from_sums(pp[row], qqv, acc);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&ctl->rerank_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) { ctl->rerank_ticket = 0; *fallback_out = 0; }
    int P = 32;
    while (P < count) P <<= 1;
    for (int i = tid; i < P; i += kS1Threads) {
        sd[i] = i < count ? __ldcg(cand_dist + i) : INFINITY;
        si[i] = i < count ? id_base + cand[i] : -1;
    }
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int i = tid; i < (P >> 1); i += kS1Threads) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const double dl = sd[lo], dh = sd[hi];
                const int il = si[lo], ih = si[hi];
                const bool swap = asc ? before(dh, ih, dl, il) : before(dl, il, dh, ih);
                if (swap) { sd[lo] = dh; sd[hi] = dl; si[lo] = ih; si[hi] = il; }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < k; i += kS1Threads) {
        const bool ok = i < count && si[i] >= 0;
        out_ids[i] = ok ? si[i] : -1;
        out_dist[i] = ok ? sd[i] : INFINITY;
    }
}

static int g_single_tma = 1;      // morna_debug_set_tuning key 3: bulk-copy staged scan (1) or per-warp loads (0)
void set_single_tma(int v) { g_single_tma = v ? 1 : 0; }

static int sm_count_s() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

struct SingleLayout { size_t ctl, score, cand, cand_dist, total; };
static SingleLayout single_layout(int64_t n) {
    SingleLayout w{};
    size_t off = 0;
    auto take = [&](size_t b) { size_t at = off; off += align_up(b, 256); return at; };
    w.ctl = take(sizeof(SingleWs));
    w.score = take((size_t)n * sizeof(float));
    w.cand = take((size_t)kS1Cand * sizeof(int32_t));
    w.cand_dist = take((size_t)kS1Cand * sizeof(double));
    w.total = off + 256;
    return w;
}

}  // namespace morna

using namespace morna;

extern "C" size_t morna_knn_single_workspace_bytes(int64_t n) { return single_layout(n > 0 ? n : 1).total; }

extern "C" int morna_knn_single(const float *vectors, const double *pp, int64_t n, int32_t dim, int64_t ld,
                                int32_t id_base, const double *query, int32_t k, int32_t *out_ids, double *out_dist,
                                int32_t *fallback, void *workspace, size_t workspace_bytes, void *stream) {
    if (!vectors || !pp || !query || !out_ids || !out_dist || !fallback || n <= 0 || n > 0x7fffffff || dim <= 0 ||
        ld < dim || (ld & 3) || k <= 0 || k > kS1Cand / 2)
        return MORNA_ERR_INVALID_ARGUMENT;
    SingleLayout w = single_layout(n);
    if (!workspace || workspace_bytes < w.total) return MORNA_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = (unsigned char *)workspace;
    SingleWs *ctl = (SingleWs *)(ws + w.ctl);
    float *score = (float *)(ws + w.score);
    int32_t *cand = (int32_t *)(ws + w.cand);
    double *cand_dist = (double *)(ws + w.cand_dist);
    MORNA_CUDA_TRY(cudaMemsetAsync(ctl, 0, sizeof(SingleWs), s));
    // fp32 rounding bound in units of |q|: per-lane chain of 4*ceil(ld/128) fused adds, a 5-level
    // shuffle tree, the fp32 rounding of q, 1/sqrt(pp) and the final scaling
    const int terms = 4 * (int)((ld + 127) / 128) + 5 + 8;
    const float eps_rel = (float)((double)terms * 5.9604644775390625e-08 * 1.05);
    size_t smem1 = (size_t)ld * sizeof(float);
    const size_t list_bytes = (size_t)kS1List * (sizeof(uint32_t) + sizeof(int));
    if (smem1 < list_bytes) smem1 = list_bytes;
    size_t smem2 = (size_t)ld * sizeof(double);
    const size_t sort_bytes = (size_t)kS1Cand * (sizeof(double) + sizeof(int));
    if (smem2 < sort_bytes) smem2 = sort_bytes;
    if (smem1 > 200 * 1024 || smem2 > 200 * 1024) return MORNA_ERR_INVALID_ARGUMENT;
    // shared-memory opt-in is sticky per function: only raise it when a call needs more than before
    static size_t attr1 = 48 * 1024, attr2 = 48 * 1024, attr_t = 48 * 1024;
    if (smem1 > attr1) {
        MORNA_CUDA_TRY(cudaFuncSetAttribute(scan32_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        attr1 = smem1;
    }
    if (smem2 > attr2) {
        MORNA_CUDA_TRY(cudaFuncSetAttribute(rerank_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        attr2 = smem2;
    }
    const size_t row_bytes = (size_t)ld * sizeof(float);
    int rows_per_stage = (int)(kTmaStageBytes / row_bytes);
    if (rows_per_stage > 8) rows_per_stage = 8;
    if (g_single_tma && rows_per_stage >= 1) {         // rows arrive as bulk copies staged in shared memory
        size_t smem_t = kTmaStages * rows_per_stage * row_bytes + row_bytes + 64;
        if (smem_t < list_bytes) smem_t = list_bytes;          // the last CTA keeps its survivor list there
        if (smem_t > attr_t) {
            MORNA_CUDA_TRY(cudaFuncSetAttribute(scan32_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
            attr_t = smem_t;
        }
        int64_t blocks = (n + rows_per_stage - 1) / rows_per_stage;
        if (blocks > sm_count_s()) blocks = sm_count_s();
        scan32_tma_kernel<<<(unsigned)blocks, kS1Threads, smem_t, s>>>(vectors, pp, n, ld, dim, rows_per_stage, query, k,
                                                                     eps_rel, score, cand, ctl);
        MORNA_LAUNCH_CHECK();
    } else {
        int64_t blocks = (n + kS1Warps - 1) / kS1Warps;
        const int64_t cap = (int64_t)sm_count_s() * 3;
        if (blocks > cap) blocks = cap;
        scan32_select_kernel<<<(unsigned)blocks, kS1Threads, smem1, s>>>(vectors, pp, n, ld, dim, query, k, eps_rel,
                                                                       score, cand, ctl);
        MORNA_LAUNCH_CHECK();
    }
    rerank_single_kernel<<<32, kS1Threads, smem2, s>>>(vectors, pp, ld, dim, id_base, query, cand, cand_dist, ctl, k,
                                                      out_ids, out_dist, fallback);
    MORNA_LAUNCH_CHECK();
    return MORNA_OK;
}

// Single-query exact search, HBM-bound, ONE kernel: every warp streams its contiguous share of the
// sample matrix once (4*N*ld bytes) and computes the exact angular distance of each row with the
// canonical FP64 sums (B200's FP64 pipe needs ~4 us for 21,504 x 3000 FMAs -- the float->double
// widening, done with integer instructions, and the loads are what has to be scheduled), writing
// the distance and a monotone float key per row.  The last CTA to finish finds the k-th smallest
// key (sampled pivot -> survivors in shared memory -> exact k-th), takes every row at or below it
// -- a superset of the true top-k, ties included, because rounding down is monotone -- reads those
// rows' double distances back, sorts them under the reference rule and writes the answer.  Results
// are identical to the FP64 scan (morna_angular_distances + morna_select_topk); ties wider than
// the candidate list fall back to it.  Replaces the per-query loop of exact_search_nn
// (morna.py:697-712).
#include <math.h>

#include "common.cuh"

namespace morna {

constexpr int kS1Threads = 512, kS1Warps = kS1Threads / 32;
constexpr int kS1List = 2048;          // survivors of the pivot kept in shared memory
constexpr int kS1Cand = 1024;          // candidates ordered by their double distances
constexpr int kBins = 8192;            // distance histogram: bin = floor(d * kBins / 2), d in [0, 2]
constexpr int kBinCap = 64;            // rows remembered per bin

struct SingleWs {          // device-side control block at the start of the workspace; the ticket is zero between calls
    unsigned int scan_ticket, pad[3];
    unsigned long long stamp[8];   // %globaltimer (ns): CTA 0 start, last CTA at the ticket, after the k-th key, at the end;
                                   // 4..7: inside the selection (pivot, filter, k-th of survivors, emission)
};
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ uint32_t fkey(float v) {
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// kk-th largest of m keys in shared memory, by all 16 warps of the CTA: four bits per round, warp w
// counts the keys >= best | (w << shift), one barrier per round (counts double-buffered).  Eight
// rounds of ~m/32 shared loads per lane replace 32 dependent rounds of a single warp.
__device__ uint32_t block_kth_largest(const uint32_t *s_keys, int m, int kk, int *s_cnt /* [2][16] */) {
    static_assert(kS1Warps == 16, "one warp per non-zero hex digit");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t best = 0;
    int buf = 0;
    for (int shift = 28; shift >= 0; shift -= 4, buf ^= 16) {
        if (warp > 0) {
            const uint32_t trial = best | ((uint32_t)warp << shift);
            int c = 0;
            for (int j = lane; j < m; j += 32) c += s_keys[j] >= trial;
            c = __reduce_add_sync(kFull, c);
            if (lane == 0) s_cnt[buf + warp] = c;
        }
        __syncthreads();
        int digit = 0;
#pragma unroll
        for (int w = 1; w < 16; ++w) digit += s_cnt[buf + w] >= kk;      // counts are non-increasing in w
        best |= (uint32_t)digit << shift;
    }
    return best;
}

// appends this thread's `mine` flagged items to a shared list: one shared atomic per warp, order free
__device__ __forceinline__ int warp_reserve(int mine, int *s_counter, int lane) {
    const int total = __reduce_add_sync(kFull, mine);
    if (total == 0) return 0;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    int base = 0;
    if (lane == 0) base = atomicAdd(s_counter, total);
    return __shfl_sync(kFull, base, 0) + incl - mine;
}

constexpr int kSelVec = 8;             // 16-byte key loads per thread per round of the filter (32 keys)

// Candidate selection over all N scores, run by ONE CTA (the last to finish the scan): keeps every
// row whose score is >= the k-th largest score.  A pivot from 512 strided samples cuts the keys to
// a few hundred survivors held in shared memory; the exact k-th largest of the survivors follows.
// Keys are read 32 per thread per round, all loads of a round in flight together.
// out: cand[0..s_small[5]) rows; s_small[4] = 1 if the generic scan must answer instead (the pivot
// missed, or ties wider than the lists).  `score` must be 16-byte aligned and readable up to the
// next multiple of four floats.
__device__ void select_candidates(const float *__restrict__ score, int64_t n, int k, int32_t *cand,
                                  uint32_t *s_key, int *s_idx, int *s_small, unsigned long long *stamp) {
    const int tid = threadIdx.x, lane = tid & 31;
    int &s_count = s_small[0];
    int &s_out = s_small[1];
    int &s_fallback = s_small[4];      // out: 1 = the generic scan must answer
    int &s_kept = s_small[5];          // out: candidates written to cand[]
    int *s_cnt = s_small + 6;          // [2][16]
    const int kk = (int)min((int64_t)k, n);
    const int n32 = (int)n;            // n <= INT_MAX (checked by the host entry)
    if (tid == 0) { s_count = 0; s_out = 0; s_fallback = 0; s_kept = 0; }
    uint32_t pivot = 1u;               // real keys are >= 1; 0 marks padding
    if (kk > 0 && n32 > kS1List) {     // pivot: r-th largest of 512 strided samples, r sized so that >= k rows survive
        s_key[tid] = fkey(__ldcg(score + (int64_t)tid * (n32 >> 9)));
        __syncthreads();
        const float want = (float)kk * 512.0f / (float)n32;
        const int r = (int)(want + 4.0f * sqrtf(want) + 4.0f);
        if (r <= 512) pivot = max(block_kth_largest(s_key, 512, r, s_cnt), 1u);
    }
    __syncthreads();
    if (kk == 0) return;
    const int groups = (n32 + 3) >> 2;
    const uint4 *src = reinterpret_cast<const uint4 *>(score);
    for (int g0 = 0; g0 < groups; g0 += kSelVec * kS1Threads) {
        uint4 raw[kSelVec];
#pragma unroll
        for (int u = 0; u < kSelVec; ++u) {
            const int g = g0 + u * kS1Threads + tid;
            raw[u] = g < groups ? __ldcg(src + g) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t key[kSelVec][4];
        int mine = 0;
#pragma unroll
        for (int u = 0; u < kSelVec; ++u) {
            const int i = 4 * (g0 + u * kS1Threads + tid);
            const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                key[u][t] = i + t < n32 ? fkey(__uint_as_float(w[t])) : 0u;
                mine += key[u][t] >= pivot;
            }
        }
        int at = warp_reserve(mine, &s_count, lane);
        if (mine) {
#pragma unroll
            for (int u = 0; u < kSelVec; ++u)
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (key[u][t] >= pivot) {
                        if (at < kS1List) { s_key[at] = key[u][t]; s_idx[at] = 4 * (g0 + u * kS1Threads + tid) + t; }
                        ++at;
                    }
        }
    }
    __syncthreads();
    const int m = s_count;
    if (m < kk || m > kS1List) {       // unlucky pivot or heavy ties: the generic scan answers
        if (tid == 0) s_fallback = 1;
        __syncthreads();
        return;
    }
    // exact kk-th largest of the m survivors; with fewer than k rows in all, everything is a candidate
    const uint32_t best = block_kth_largest(s_key, m, kk, s_cnt);
    const uint32_t cut_key = kk >= k ? best : 1u;
    if (cut_key < pivot) {             // the cut fell below the pivot: survivors may miss candidates
        if (tid == 0) s_fallback = 1;
        __syncthreads();
        return;
    }
    for (int j0 = 0; j0 < m; j0 += kS1Threads) {
        const int at = j0 + tid;
        const bool keep = at < m && s_key[at] >= cut_key;
        const int o = warp_reserve(keep ? 1 : 0, &s_out, lane);
        if (keep && o < kS1Cand) cand[o] = s_idx[at];
    }
    __syncthreads();
    if (tid == 0) {
        const int kept = s_out;
        s_kept = kept <= kS1Cand ? kept : 0;
        s_fallback = kept <= kS1Cand ? 0 : 1;
    }
    __syncthreads();
}

// A chain of queries on one stream (morna_knn_single_stream): the kernels after the first are launched with
// programmatic stream serialisation, so query j+1's kernel may begin once every CTA of query j's kernel has called
// this -- after its share of the scan -- and its scan then runs under query j's selection tail (one CTA, ~7 us, 147 SMs
// idle) and the launch gap.  Workspaces alternate between two halves: the wait BEFORE the trigger makes "kernel j+1
// let kernel j+2 start" imply "kernel j has completed", so the half kernel j+2 reuses is free and zeroed again.
__device__ __forceinline__ void chain_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void chain_handoff() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// rows per warp pass R, chunk steps in flight U (R*U 16-byte loads per lane)
template <int R, int U>
__global__ void __launch_bounds__(kS1Threads, 2)
scan64_select_kernel(const float *__restrict__ vectors, const double *__restrict__ pp, int64_t n, int64_t ld,
                     int32_t dim, int32_t id_base, const double *__restrict__ query, int32_t k, int64_t rows_per_warp,
                     int32_t prefetch_bytes, int32_t prefetch_rows, int32_t chain, double *__restrict__ dist, float *__restrict__ sel, unsigned int *__restrict__ hist,
                     int32_t *__restrict__ lists, SingleWs *ctl,
                     int32_t *__restrict__ out_ids, double *__restrict__ out_dist, int32_t *__restrict__ fallback_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *qs = reinterpret_cast<double *>(smem_raw);                 // [ld] query * 2^896
    // the last CTA reuses the storage: survivor list, candidate list, sort buffers
    unsigned int *hist_s = reinterpret_cast<unsigned int *>(smem_raw); // [kBins] copy of the histogram
    uint32_t *s_key = reinterpret_cast<uint32_t *>(smem_raw);          // [kS1List]  (select_candidates, after hist_s is dead)
    int *s_idx = reinterpret_cast<int *>(s_key + kS1List);             // [kS1List]
    int *s_cand = reinterpret_cast<int *>(hist_s + kBins);             // [kS1Cand]
    double *sd = reinterpret_cast<double *>(s_cand + kS1Cand);         // [kS1Cand]
    int *si = reinterpret_cast<int *>(sd + kS1Cand);                   // [kS1Cand]
    __shared__ int s_small[6 + 2 * kS1Warps];
    __shared__ int s_huge;
    __shared__ bool s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunks = (int)(ld >> 2);
    if (tid == 0) s_huge = 0;
    // hand-over 3: the next query's kernel may begin as soon as this one is resident; each kernel then waits for its
    // predecessor just before its first write into the workspace (see chain_wait below)
    if (chain == 3) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    bool must_wait = chain == 3;
    if (blockIdx.x == 0 && tid == 0 && chain != 3) ctl->stamp[0] = global_ns();     // (debug stamp; a workspace write)
    if (prefetch_bytes > 0) {
        // The query staging and the qq sum below take ~3 us in which no row is read.  Pull the head of this warp's
        // first-pass rows into L2 meanwhile (no registers held): HBM starts streaming at once.  Measured at 21,504 x 3000
        // (scripts/single_prefetch_sweep.py, scripts/single_bench_ab.py): 6 KB per row -- a quarter of the matrix -- gives
        // 59.1 -> 55.8 us one query at a time and 39.2 -> 37.0 us with two in flight; whole rows (half the matrix = the
        // L2's size) give 55.5 us but 42.5 us with two in flight, and more rows or prefetching from inside the loop
        // overflow L2 (57-78 us).
        const int64_t gw0 = (int64_t)blockIdx.x * kS1Warps + warp;
        const int64_t first = gw0 * rows_per_warp;
        const int lines = min(prefetch_bytes, (int)(ld * 4)) >> 7;            // 128-byte lines per row
        const int rows_ahead = prefetch_rows > 0 ? prefetch_rows : R;
        for (int u = 0; u < rows_ahead && u < rows_per_warp; ++u) {
            if (first + u >= n) break;
            const char *row = reinterpret_cast<const char *>(vectors + (first + u) * ld);
            for (int l = lane; l < lines; l += 32) asm volatile("prefetch.global.L2 [%0];" :: "l"(row + 128 * l));
        }
    }
    __syncthreads();
    {
        bool huge = false;
        for (int c = tid; c < ld; c += kS1Threads) {
            const double v = c < dim ? query[c] : 0.0;
            huge |= !(fabs(v) < kScaledQueryMax);
            qs[c] = v * kTwo896;
        }
        if (huge) s_huge = 1;
    }
    __syncthreads();
    if (s_huge) {                        // q * 2^896 is not finite: the generic FP64 scan answers
        if (blockIdx.x == 0 && tid == 0) *fallback_out = 1;
        if (chain) chain_handoff();      // (a kernel of a chain never ends before its predecessor: the kernel after it reuses that one's workspace)
        return;
    }
    // qq with the canonical tree (every warp computes it for itself; 2^-896 undoes the staging exactly)
    double qq = 0.0;
    for (int c = lane; c < chunks; c += 32) {
        const double2 a = *reinterpret_cast<const double2 *>(qs + 4 * c);
        const double2 b = *reinterpret_cast<const double2 *>(qs + 4 * c + 2);
        const double x0 = a.x * 0x1p-896, x1 = a.y * 0x1p-896, x2 = b.x * 0x1p-896, x3 = b.y * 0x1p-896;
        qq = fma(x0, x0, qq); qq = fma(x1, x1, qq); qq = fma(x2, x2, qq); qq = fma(x3, x3, qq);
    }
    qq = warp_sum(qq);

    const int64_t gw = (int64_t)blockIdx.x * kS1Warps + warp;
    const int64_t row_begin = gw * rows_per_warp;
    const int64_t row_end = min(n, row_begin + rows_per_warp);
    int64_t r = row_begin;
    for (; r + R <= row_end; r += R) {                                 // R whole rows, regular stride
        const float4 *src = reinterpret_cast<const float4 *>(vectors + r * ld);
        const double my_pp = lane < R ? pp[r + lane] : 0.0;            // in flight under the row loads
        double acc[R];
#pragma unroll
        for (int u = 0; u < R; ++u) acc[u] = 0.0;
        int c = lane;
        for (; c + 32 * (U - 1) < chunks; c += 32 * U) {
            float4 v[U][R];
#pragma unroll
            for (int h = 0; h < U; ++h)
#pragma unroll
                for (int u = 0; u < R; ++u) v[h][u] = ldg_stream_f4(src + (int64_t)u * chunks + c + 32 * h);
#pragma unroll
            for (int h = 0; h < U; ++h) {
                const double2 qa = *reinterpret_cast<const double2 *>(qs + 4 * (c + 32 * h));
                const double2 qb = *reinterpret_cast<const double2 *>(qs + 4 * (c + 32 * h) + 2);
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    acc[u] = fma(f32_scaled_f64(v[h][u].x), qa.x, acc[u]);
                    acc[u] = fma(f32_scaled_f64(v[h][u].y), qa.y, acc[u]);
                    acc[u] = fma(f32_scaled_f64(v[h][u].z), qb.x, acc[u]);
                    acc[u] = fma(f32_scaled_f64(v[h][u].w), qb.y, acc[u]);
                }
            }
        }
        for (; c < chunks; c += 32) {
            const double2 qa = *reinterpret_cast<const double2 *>(qs + 4 * c);
            const double2 qb = *reinterpret_cast<const double2 *>(qs + 4 * c + 2);
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const float4 v = ldg_stream_f4(src + (int64_t)u * chunks + c);
                acc[u] = fma(f32_scaled_f64(v.x), qa.x, acc[u]);
                acc[u] = fma(f32_scaled_f64(v.y), qa.y, acc[u]);
                acc[u] = fma(f32_scaled_f64(v.z), qb.x, acc[u]);
                acc[u] = fma(f32_scaled_f64(v.w), qb.y, acc[u]);
            }
        }
        double mine = 0.0;
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const double s = warp_sum(acc[u]);
            if (lane == u) mine = s;
        }
        if (must_wait) { chain_wait(); must_wait = false; }           // (hand-over 3) the predecessor has completed: the workspace is ours
        if (lane < R) {
            const double d = angular_from_sums(my_pp, qq, mine);
            dist[r + lane] = d;
            sel[r + lane] = -__double2float_rd(d);                     // larger = nearer; monotone in d
            const int bin = min(kBins - 1, (int)(d * (kBins / 2)));
            const unsigned int slot = atomicAdd(hist + bin, 1u);
            if (slot < kBinCap) lists[bin * kBinCap + slot] = (int)(r + lane);
        }
        if (chain == 2 && r == row_begin) chain_handoff();
    }
    for (; r < row_end; ++r) {                                         // leftover rows of this warp, one at a time
        const float4 *src = reinterpret_cast<const float4 *>(vectors + r * ld);
        double acc = 0.0;
        for (int c = lane; c < chunks; c += 32) {
            const float4 v = ldg_stream_f4(src + c);
            const double2 qa = *reinterpret_cast<const double2 *>(qs + 4 * c);
            const double2 qb = *reinterpret_cast<const double2 *>(qs + 4 * c + 2);
            acc = fma(f32_scaled_f64(v.x), qa.x, acc); acc = fma(f32_scaled_f64(v.y), qa.y, acc);
            acc = fma(f32_scaled_f64(v.z), qb.x, acc); acc = fma(f32_scaled_f64(v.w), qb.y, acc);
        }
        acc = warp_sum(acc);
        if (must_wait) { chain_wait(); must_wait = false; }
        if (lane == 0) {
            const double d = angular_from_sums(pp[r], qq, acc);
            dist[r] = d;
            sel[r] = -__double2float_rd(d);
            const int bin = min(kBins - 1, (int)(d * (kBins / 2)));
            const unsigned int slot = atomicAdd(hist + bin, 1u);
            if (slot < kBinCap) lists[bin * kBinCap + slot] = (int)r;
        }
    }
    if (chain == 1 || (chain == 2 && row_end - row_begin <= R)) chain_handoff();
    if (must_wait) { chain_wait(); must_wait = false; }               // warps without rows; the ticket below is a workspace write
    // last CTA to arrive answers the query from all keys
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&ctl->scan_ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) { ctl->scan_ticket = 0; ctl->stamp[1] = global_ns(); }   // ticket left zeroed for the next call
    // The distance histogram all CTAs filled (and each bin's first rows) gives the candidates directly:
    // the bin holding the k-th nearest row, and every row in a bin at or below it.
    const int kk = (int)min((int64_t)k, n);
    int count = -1;                                                    // -1: not answered by the histogram
    {
        static_assert(kBins == 16 * kS1Threads, "sixteen bins per thread");
        uint4 h4[4];
        uint4 *hsrc = reinterpret_cast<uint4 *>(hist) + 4 * tid;
#pragma unroll
        for (int u = 0; u < 4; ++u) h4[u] = __ldcg(hsrc + u);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            hsrc[u] = make_uint4(0u, 0u, 0u, 0u);                      // left zeroed for the next call
            reinterpret_cast<uint4 *>(hist_s)[4 * tid + u] = h4[u];
        }
        const unsigned int h[16] = {h4[0].x, h4[0].y, h4[0].z, h4[0].w, h4[1].x, h4[1].y, h4[1].z, h4[1].w,
                                    h4[2].x, h4[2].y, h4[2].z, h4[2].w, h4[3].x, h4[3].y, h4[3].z, h4[3].w};
        int local = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) local += (int)h[j];
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        int *s_wsum = s_small + 6;                                     // [16] warp totals
        if (tid == 0) { s_small[0] = -1; s_small[1] = 0; s_small[2] = 0; s_small[3] = 0; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int before_me = incl - local;
        for (int w = 0; w < warp; ++w) before_me += s_wsum[w];
        if (kk > 0 && before_me < kk && kk <= before_me + local) {     // the k-th nearest row is in one of my bins
            int run = before_me, j = 0;
            for (; j < 16; ++j) { run += (int)h[j]; if (run >= kk) break; }
            s_small[0] = 16 * tid + j;                                 // last bin taken
            s_small[1] = run;                                          // rows in the bins up to it
        }
        __syncthreads();
        const int last_bin = s_small[0], m = s_small[1];
        if (kk == 0) count = 0;
        else if (last_bin >= 0 && m <= kS1Cand) {
            for (int b = tid; b <= last_bin; b += kS1Threads) {
                const int c = (int)hist_s[b];
                if (c == 0) continue;
                if (c > kBinCap) { s_small[3] = 1; continue; }         // a bin overflowed its row list
                const int o = atomicAdd(&s_small[2], c);
                for (int t = 0; t < c; ++t) s_cand[o + t] = __ldcg(lists + b * kBinCap + t);
            }
            __syncthreads();
            if (!s_small[3]) count = m;
        }
        __syncthreads();
    }
    if (tid == 0) ctl->stamp[2] = global_ns();
    if (count < 0) {                   // crowded bins (heavy ties): select from the float keys instead
        select_candidates(sel, n, k, s_cand, s_key, s_idx, s_small, ctl->stamp);
        if (s_small[4]) {
            if (tid == 0) *fallback_out = 1;
            return;
        }
        count = s_small[5];
    }
    if (tid == 0) *fallback_out = 0;
    if (count <= 256) {
        // short list: every thread ranks one candidate against all others (no barriers in the loop) and
        // writes it straight to its place
        double my_d = INFINITY;
        int my_id = -1;
        if (tid < count) { const int row = s_cand[tid]; my_d = __ldcg(dist + row); my_id = id_base + row; }
        if (tid < 256) { sd[tid] = my_d; si[tid] = my_id; }
        __syncthreads();
        if (tid < count) {
            int rank = 0;
            for (int j = 0; j < count; ++j) rank += before(sd[j], si[j], my_d, my_id);
            if (rank < k) { out_ids[rank] = my_id; out_dist[rank] = my_d; }
        }
        for (int i = count + tid; i < k; i += kS1Threads) { out_ids[i] = -1; out_dist[i] = INFINITY; }
    } else {
        int P = 512;
        while (P < count) P <<= 1;
        for (int i = tid; i < P; i += kS1Threads) {
            const int row = i < count ? s_cand[i] : 0;
            sd[i] = i < count ? __ldcg(dist + row) : INFINITY;
            si[i] = i < count ? id_base + row : -1;
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
    __syncthreads();
    if (tid == 0) ctl->stamp[3] = global_ns();
}

static int g_single_prefetch = 6144;   // morna_debug_set_tuning key 27: bytes (clamped to the row) of each first-pass row pulled into L2 before the query is staged
void set_single_prefetch(int v) { g_single_prefetch = v >= 0 ? v : 0; }
static int g_single_prefetch_rows = 0;   // key 28: rows of the warp's share covered by that prefetch (0 = the first pass)
void set_single_prefetch_rows(int v) { g_single_prefetch_rows = v >= 0 ? v : 0; }
static int g_single_rows = 0;     // morna_debug_set_tuning key 3: rows per warp pass (0 = automatic)
void set_single_tma(int v) { g_single_rows = v; }

static int sm_count_s() { return sm_count_current(); }

struct SingleLayout { size_t ctl, hist, lists, dist, sel, total; };
static SingleLayout single_layout(int64_t n) {
    SingleLayout w{};
    size_t off = 0;
    auto take = [&](size_t b) { size_t at = off; off += align_up(b, 256); return at; };
    w.ctl = take(sizeof(SingleWs));
    w.hist = take((size_t)kBins * sizeof(unsigned int));              // ctl + hist: zero between calls
    w.lists = take((size_t)kBins * kBinCap * sizeof(int32_t));
    w.dist = take((size_t)n * sizeof(double));
    w.sel = take((size_t)n * sizeof(float));
    w.total = off + 256;
    return w;
}

template <int R, int U>
static int launch_scan64(unsigned grid, size_t smem, cudaStream_t s, const float *vectors, const double *pp, int64_t n,
                         int64_t ld, int32_t dim, int32_t id_base, const double *query, int32_t k, int64_t rows_per_warp,
                         double *dist, float *sel, unsigned int *hist, int32_t *lists, SingleWs *ctl, int32_t *out_ids, double *out_dist, int32_t *fallback,
                         int chain, bool early) {
    { int rca = ensure_dynamic_smem((const void *)scan64_select_kernel<R, U>, smem); if (rca != MORNA_OK) return rca; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kS1Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = early ? 1 : 0;          // early: may begin before the preceding kernel of the chain has drained
    MORNA_CUDA_TRY(cudaLaunchKernelEx(&cfg, scan64_select_kernel<R, U>, vectors, pp, n, ld, dim, id_base, query, k, rows_per_warp,
                                      (int32_t)g_single_prefetch, (int32_t)g_single_prefetch_rows, (int32_t)chain, dist, sel, hist, lists,
                                      ctl, out_ids, out_dist, fallback));
    MORNA_LAUNCH_CHECK();
    return MORNA_OK;
}

}  // namespace morna

using namespace morna;

extern "C" size_t morna_knn_single_workspace_bytes(int64_t n) { return single_layout(n > 0 ? n : 1).total; }

extern "C" int morna_knn_single_workspace_init(void *workspace, size_t workspace_bytes, void *stream) {
    if (!workspace) return MORNA_ERR_WORKSPACE_TOO_SMALL;
    const SingleLayout w = single_layout(1);
    if (workspace_bytes < w.lists) return MORNA_ERR_WORKSPACE_TOO_SMALL;
    MORNA_CUDA_TRY(cudaMemsetAsync(workspace, 0, w.lists, (cudaStream_t)stream));
    return MORNA_OK;
}

static int knn_single_launch(const float *vectors, const double *pp, int64_t n, int32_t dim, int64_t ld,
                             int32_t id_base, const double *query, int32_t k, int32_t *out_ids, double *out_dist,
                             int32_t *fallback, void *workspace, size_t workspace_bytes, void *stream, int chain, bool early) {
    if (!vectors || !pp || !query || !out_ids || !out_dist || !fallback || n <= 0 || n > 0x7fffffff || dim <= 0 ||
        ld < dim || (ld & 3) || k <= 0 || k > kS1Cand / 2)
        return MORNA_ERR_INVALID_ARGUMENT;
    SingleLayout w = single_layout(n);
    if (!workspace || workspace_bytes < w.total) return MORNA_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = (unsigned char *)workspace;
    SingleWs *ctl = (SingleWs *)(ws + w.ctl);
    double *dist = (double *)(ws + w.dist);
    float *sel = (float *)(ws + w.sel);
    unsigned int *hist = (unsigned int *)(ws + w.hist);
    int32_t *lists = (int32_t *)(ws + w.lists);
    size_t smem = (size_t)ld * sizeof(double);
    const size_t tail_bytes = (size_t)kBins * 4 + (size_t)kS1Cand * (4 + 8 + 4);
    static_assert(kBins * 4 >= kS1List * 8, "the survivor lists overlay the histogram copy");
    if (smem < tail_bytes) smem = tail_bytes;
    if (smem > 100 * 1024) {
        // very wide --features (D > 12,800): the staged query would not leave room for two CTAs per SM; the generic
        // scan (morna_knn_exact, which reads a query this wide from global memory) answers -- same results
        static const int32_t one = 1;
        MORNA_CUDA_TRY(cudaMemcpyAsync(fallback, &one, sizeof(one), cudaMemcpyHostToDevice, s));
        return MORNA_OK;
    }
    // rows per warp, in passes of R rows.  Measured on B200 (scripts/single_rows_sweep.py): passes of three
    // rows with two chunk steps in flight beat one pass of the warp's whole share -- the first pass's
    // histogram atomics and distance stores overlap the second pass's loads instead of all landing at
    // the end -- and beat four or five rows per pass, whose coarser shares leave warps idle.
    const int64_t max_warps = (int64_t)sm_count_s() * 2 * kS1Warps;
    int64_t rpw = (n + max_warps - 1) / max_warps;
    int R = rpw < 3 ? (int)rpw : 3;
    if (g_single_rows > 0 && g_single_rows <= 5) R = g_single_rows;
    rpw = (rpw + R - 1) / R * R;
    const int64_t warps = (n + rpw - 1) / rpw;
    const unsigned grid = (unsigned)((warps + kS1Warps - 1) / kS1Warps);
#define MORNA_SCAN64(RR, UU)                                                                                          \
    case RR: return launch_scan64<RR, UU>(grid, smem, s, vectors, pp, n, ld, dim, id_base, query, k, rpw, dist, sel, \
                                          hist, lists, ctl, out_ids, out_dist, fallback, chain, early)
    switch (R) {
        MORNA_SCAN64(1, 8); MORNA_SCAN64(2, 4); MORNA_SCAN64(3, 2); MORNA_SCAN64(4, 2); MORNA_SCAN64(5, 1);
    }
#undef MORNA_SCAN64
    return MORNA_ERR_INVALID_ARGUMENT;
}

extern "C" int morna_knn_single(const float *vectors, const double *pp, int64_t n, int32_t dim, int64_t ld,
                                int32_t id_base, const double *query, int32_t k, int32_t *out_ids, double *out_dist,
                                int32_t *fallback, void *workspace, size_t workspace_bytes, void *stream) {
    return knn_single_launch(vectors, pp, n, dim, ld, id_base, query, k, out_ids, out_dist, fallback, workspace, workspace_bytes,
                             stream, 0, false);
}

static int g_single_chain = 2;      // morna_debug_set_tuning key 29: 0 = plain launches, 1 = hand over after the scan, 2 = after the first pass (measured
                                    // at 21,504 x 3000, scripts/single_chain_probe.py: 58.3 / 49.0 / 44.5 us per query)
namespace morna { void set_single_chain(int v) { g_single_chain = v < 0 ? 0 : (v > 3 ? 3 : v); } }

extern "C" int morna_knn_single_stream(const float *vectors, const double *pp, int64_t n, int32_t dim, int64_t ld,
                                       int32_t id_base, const double *queries, int64_t query_stride, int32_t n_queries,
                                       int32_t k, int32_t *out_ids, double *out_dist, int32_t *fallback, void *workspace,
                                       size_t workspace_bytes, void *stream) {
    if (n_queries < 0 || (n_queries > 0 && (!queries || query_stride < dim))) return MORNA_ERR_INVALID_ARGUMENT;
    const size_t half = single_layout(n > 0 ? n : 1).total;
    if (!workspace || workspace_bytes < 2 * half) return MORNA_ERR_WORKSPACE_TOO_SMALL;
    const int chain = g_single_chain;
    for (int32_t j = 0; j < n_queries; ++j) {
        // the first kernel is an ordinary launch: it starts after everything the stream already holds, inputs included
        const int rc = knn_single_launch(vectors, pp, n, dim, ld, id_base, queries + (int64_t)j * query_stride, k,
                                         out_ids + (int64_t)j * k, out_dist + (int64_t)j * k, fallback + j,
                                         (unsigned char *)workspace + (size_t)(j & 1) * half, half, stream, chain,
                                         chain != 0 && j > 0);
        if (rc != MORNA_OK) return rc;
    }
    return MORNA_OK;
}

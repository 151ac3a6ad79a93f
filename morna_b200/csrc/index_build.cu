// Index build: feature hashing, first-seen internal ids, order-faithful scatter-add,
// float32 round/store.  Replaces morna.py:369-388 and the add_item cast (:405-407).
#include <math.h>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace morna {

// ------------------------------------------------------------------ K1 hashing
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// MurmurHash3_x86_32, seed 0 (what mmh3.hash computes at morna.py:369)
__device__ __forceinline__ uint32_t murmur3_32(const uint8_t *__restrict__ key, int len) {
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = 0;
    const int nblocks = len >> 2;
    for (int i = 0; i < nblocks; ++i) {
        const uint8_t *p = key + 4 * i;          // keys are byte-packed: no alignment assumed
        uint32_t k = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        k *= c1; k = rotl32(k, 15); k *= c2;
        h ^= k; h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
    }
    const uint8_t *tail = key + 4 * nblocks;
    uint32_t k = 0;
    const int rem = len & 3;
    if (rem == 3) k ^= (uint32_t)tail[2] << 16;
    if (rem >= 2) k ^= (uint32_t)tail[1] << 8;
    if (rem >= 1) {
        k ^= tail[0];
        k *= c1; k = rotl32(k, 15); k *= c2; h ^= k;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256)
hash_junctions_kernel(const uint8_t *__restrict__ keys, const int32_t *__restrict__ key_off, int64_t n_rows,
                      int32_t dim, int32_t *__restrict__ raw, int32_t *__restrict__ bucket,
                      int8_t *__restrict__ sign) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_rows;
         j += (int64_t)gridDim.x * blockDim.x) {
        int32_t b = key_off[j], e = key_off[j + 1];
        int32_t h = (int32_t)murmur3_32(keys + b, e - b);
        int32_t m = h % dim;                 // C remainder has the dividend's sign ...
        if (m < 0) m += dim;                 // ... Python's floor-mod does not (morna.py:371)
        raw[j] = h; bucket[j] = m; sign[j] = h < 0 ? (int8_t)-1 : (int8_t)1;
    }
}

// ------------------------------------------------------------------ K2 internal ids
__global__ void __launch_bounds__(256)
ids_init_kernel(unsigned long long *first_pos, int32_t *vals, int64_t m, unsigned long long sentinel) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        first_pos[i] = sentinel;
        vals[i] = (int32_t)i;
    }
}

// one warp per row: earliest pair position of every sample among passing rows
__global__ void __launch_bounds__(256)
first_position_kernel(const int64_t *__restrict__ row_off, const uint8_t *__restrict__ pass, int64_t n_rows,
                      const int32_t *__restrict__ sample, int32_t max_sample_id,
                      unsigned long long *__restrict__ first_pos) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n_rows; j += warps_total) {
        if (!pass[j]) continue;
        const int64_t b = row_off[j], e = row_off[j + 1];
        for (int64_t p = b + lane; p < e; p += 32) {
            int32_t s = sample[p];
            if ((uint32_t)s <= (uint32_t)max_sample_id) atomicMin(&first_pos[s], (unsigned long long)p);
        }
    }
}

__global__ void __launch_bounds__(256)
ids_assign_kernel(const unsigned long long *__restrict__ sorted_pos, const int32_t *__restrict__ sorted_sample,
                  int64_t m, unsigned long long sentinel, int32_t *__restrict__ id_of_sample,
                  int32_t *__restrict__ n_kept) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        bool seen = sorted_pos[i] < sentinel;
        id_of_sample[sorted_sample[i]] = seen ? (int32_t)i : -1;
        bool next_seen = (i + 1 < m) && sorted_pos[i + 1] < sentinel;
        if (seen && !next_seen) *n_kept = (int32_t)(i + 1);
        if (i == 0 && !seen) *n_kept = 0;
    }
}

// ------------------------------------------------------------------ K3 scatter-add
__global__ void __launch_bounds__(256)
bucket_keys_kernel(const uint8_t *__restrict__ pass, const int32_t *__restrict__ bucket, int64_t n_rows,
                   int32_t dim, int32_t *__restrict__ keys, int32_t *__restrict__ vals) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_rows; j += (int64_t)gridDim.x * blockDim.x) {
        keys[j] = pass[j] ? bucket[j] : dim;   // rows under the threshold sort past the last bucket
        vals[j] = (int32_t)j;
    }
}

// begin[b] = first position in the bucket-sorted row list whose key >= b, b in [0, dim+1]
__global__ void __launch_bounds__(256)
bucket_begin_kernel(const int32_t *__restrict__ sorted_keys, int64_t n_rows, int32_t dim,
                    int32_t *__restrict__ begin) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_rows; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t prev = i > 0 ? sorted_keys[i - 1] : -1;
        int32_t cur = i < n_rows ? sorted_keys[i] : dim + 1;
        for (int32_t b = prev + 1; b <= cur; ++b) begin[b] = (int32_t)i;
    }
}

constexpr int kAccThreads = 512;
constexpr int kAccTile = 24576;   // doubles of one bucket column held in shared memory (192 KB)

// CTA (b, t) owns bucket b for internal ids [lo, hi).  Rows of the bucket are applied
// in file order with a barrier between rows, so each cell sees its addends in the
// reference's order (morna.py:376-388) and the double sums are bit-identical.
__global__ void __launch_bounds__(kAccThreads)
index_accumulate_kernel(const int64_t *__restrict__ row_off, const int8_t *__restrict__ sign,
                        const double *__restrict__ idf, const int32_t *__restrict__ sample,
                        const int32_t *__restrict__ cov, const int32_t *__restrict__ id_of_sample,
                        const int32_t *__restrict__ rows_by_bucket, const int32_t *__restrict__ bucket_begin,
                        int32_t id_lo, int32_t id_hi, int32_t tile, double *__restrict__ acc, int64_t acc_ld) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *col = reinterpret_cast<double *>(smem_raw);
    const int b = blockIdx.x;
    const int32_t lo = id_lo + (int32_t)blockIdx.y * tile;
    const int32_t hi = min(id_hi, lo + tile);
    const int32_t width = hi - lo;
    const int32_t r_begin = bucket_begin[b], r_end = bucket_begin[b + 1];
    double *out = acc + (int64_t)b * acc_ld + (lo - id_lo);
    if (r_begin == r_end) return;            // acc was zero-filled by the caller
    for (int i = threadIdx.x; i < width; i += kAccThreads) col[i] = 0.0;
    __syncthreads();
    for (int32_t r = r_begin; r < r_end; ++r) {
        const int32_t j = rows_by_bucket[r];
        const double w = idf[j];
        const double mult = (double)sign[j];
        const int64_t pb = row_off[j], pe = row_off[j + 1];
        for (int64_t p = pb + threadIdx.x; p < pe; p += kAccThreads) {
            int32_t id = id_of_sample[sample[p]];
            if (id >= lo && id < hi) {
                double tfidf = (double)cov[p] * w;
                atomicAdd(&col[id - lo], mult * tfidf);
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < width; i += kAccThreads) out[i] = col[i];
}

// ------------------------------------------------------------------ round + transpose
__global__ void __launch_bounds__(256)
round_store_kernel(const double *__restrict__ acc, int64_t acc_ld, int32_t n_ids, int32_t dim,
                   float *__restrict__ vectors, int64_t ld) {
    __shared__ double tile[32][33];
    const int id0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {        // r: bucket within tile, x: id within tile
        int b = b0 + r, id = id0 + threadIdx.x;
        tile[r][threadIdx.x] = (b < dim && id < n_ids) ? acc[(int64_t)b * acc_ld + id] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {        // r: id within tile, x: bucket within tile
        int id = id0 + r, b = b0 + threadIdx.x;
        if (id < n_ids && b < ld) vectors[(int64_t)id * ld + b] = (float)tile[threadIdx.x][r];
    }
}

static int bits_for(uint64_t v) {
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}

struct IdsWs { size_t first_pos, sorted_pos, vals_in, vals_out, cub, cub_bytes, total; };
static IdsWs ids_ws_layout(int64_t m) {
    IdsWs w{};
    size_t off = 0;
    w.first_pos = off; off += align_up((size_t)m * 8, 256);
    w.sorted_pos = off; off += align_up((size_t)m * 8, 256);
    w.vals_in = off; off += align_up((size_t)m * 4, 256);
    w.vals_out = off; off += align_up((size_t)m * 4, 256);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr,
                                    (unsigned long long *)nullptr, (const int32_t *)nullptr,
                                    (int32_t *)nullptr, (int)m, 0, 64);
    w.cub = off; w.cub_bytes = cub_bytes; off += align_up(cub_bytes, 256);
    w.total = off + 256;
    return w;
}

struct AccWs { size_t keys_in, keys_out, vals_in, vals_out, begin, cub, cub_bytes, total; };
static AccWs acc_ws_layout(int64_t n_rows, int32_t dim) {
    AccWs w{};
    size_t off = 0;
    w.keys_in = off; off += align_up((size_t)n_rows * 4, 256);
    w.keys_out = off; off += align_up((size_t)n_rows * 4, 256);
    w.vals_in = off; off += align_up((size_t)n_rows * 4, 256);
    w.vals_out = off; off += align_up((size_t)n_rows * 4, 256);
    w.begin = off; off += align_up((size_t)(dim + 2) * 4, 256);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const int32_t *)nullptr, (int32_t *)nullptr,
                                    (const int32_t *)nullptr, (int32_t *)nullptr, (int)n_rows, 0, 32);
    w.cub = off; w.cub_bytes = cub_bytes; off += align_up(cub_bytes, 256);
    w.total = off + 256;
    return w;
}

static unsigned grid_for(int64_t work, int threads) {
    int64_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > 148 * 32) g = 148 * 32;
    return (unsigned)g;
}

}  // namespace morna

using namespace morna;

extern "C" int morna_hash_junctions(const uint8_t *keys, const int32_t *key_off, int64_t n_rows, int32_t dim,
                                    int32_t *raw, int32_t *bucket, int8_t *sign, void *stream) {
    if (n_rows < 0 || dim <= 0 || !key_off || !raw || !bucket || !sign) return MORNA_ERR_INVALID_ARGUMENT;
    if (n_rows == 0) return MORNA_OK;
    if (!keys) return MORNA_ERR_INVALID_ARGUMENT;
    hash_junctions_kernel<<<grid_for(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(keys, key_off, n_rows, dim,
                                                                                  raw, bucket, sign);
    MORNA_LAUNCH_CHECK();
    return MORNA_OK;
}

extern "C" int morna_idf_host(const int64_t *running_freq, const uint8_t *pass, int64_t n_rows,
                              int64_t sample_count, double *idf) {
    if (n_rows < 0 || (n_rows > 0 && (!running_freq || !pass || !idf))) return MORNA_ERR_INVALID_ARGUMENT;
    for (int64_t j = 0; j < n_rows; ++j) {
        if (pass[j] && running_freq[j] > 0)
            idf[j] = log((double)sample_count / (double)running_freq[j]);   // morna.py:372-374
        else
            idf[j] = 0.0;
    }
    return MORNA_OK;
}

extern "C" size_t morna_assign_internal_ids_workspace_bytes(int64_t n_rows, int64_t nnz, int32_t max_sample_id) {
    (void)n_rows; (void)nnz;
    if (max_sample_id < 0) return 256;
    return ids_ws_layout((int64_t)max_sample_id + 1).total;
}

extern "C" int morna_assign_internal_ids(const int64_t *row_off, const uint8_t *pass, int64_t n_rows,
                                         const int32_t *sample, int64_t nnz, int32_t max_sample_id,
                                         int32_t *id_of_sample, int32_t *n_kept, void *workspace,
                                         size_t workspace_bytes, void *stream) {
    if (!row_off || !pass || !id_of_sample || !n_kept || n_rows < 0 || nnz < 0 || max_sample_id < 0)
        return MORNA_ERR_INVALID_ARGUMENT;
    if (nnz > 0 && !sample) return MORNA_ERR_INVALID_ARGUMENT;
    const int64_t m = (int64_t)max_sample_id + 1;
    IdsWs w = ids_ws_layout(m);
    if (!workspace || workspace_bytes < w.total) return MORNA_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = (unsigned char *)workspace;
    auto *first_pos = (unsigned long long *)(ws + w.first_pos);
    auto *sorted_pos = (unsigned long long *)(ws + w.sorted_pos);
    auto *vals_in = (int32_t *)(ws + w.vals_in);
    auto *vals_out = (int32_t *)(ws + w.vals_out);
    const unsigned long long sentinel = (unsigned long long)nnz;   // one past the last pair position
    ids_init_kernel<<<grid_for(m, 256), 256, 0, s>>>(first_pos, vals_in, m, sentinel);
    MORNA_LAUNCH_CHECK();
    if (n_rows > 0 && nnz > 0) {
        first_position_kernel<<<grid_for(n_rows * 32, 256), 256, 0, s>>>(row_off, pass, n_rows, sample,
                                                                        max_sample_id, first_pos);
        MORNA_LAUNCH_CHECK();
    }
    size_t cub_bytes = w.cub_bytes;
    MORNA_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws + w.cub, cub_bytes, first_pos, sorted_pos, vals_in, vals_out,
                                                   (int)m, 0, bits_for(sentinel), s));
    count_launch(3);
    ids_assign_kernel<<<grid_for(m, 256), 256, 0, s>>>(sorted_pos, vals_out, m, sentinel, id_of_sample, n_kept);
    MORNA_LAUNCH_CHECK();
    return MORNA_OK;
}

extern "C" size_t morna_index_accumulate_workspace_bytes(int64_t n_rows, int32_t dim) {
    if (n_rows < 0 || dim <= 0) return 256;
    return acc_ws_layout(n_rows, dim).total;
}

extern "C" int morna_index_accumulate(const int64_t *row_off, const uint8_t *pass, const int32_t *bucket,
                                      const int8_t *sign, const double *idf, int64_t n_rows,
                                      const int32_t *sample, const int32_t *cov, int64_t nnz,
                                      const int32_t *id_of_sample, int32_t id_lo, int32_t id_hi, int32_t dim,
                                      double *acc, int64_t acc_ld, void *workspace, size_t workspace_bytes,
                                      void *stream) {
    if (!row_off || !pass || !bucket || !sign || !idf || !id_of_sample || !acc || n_rows < 0 || nnz < 0 ||
        dim <= 0 || id_lo < 0 || id_hi < id_lo || acc_ld < (int64_t)(id_hi - id_lo) || n_rows > 0x7fffffff)
        return MORNA_ERR_INVALID_ARGUMENT;
    if (nnz > 0 && (!sample || !cov)) return MORNA_ERR_INVALID_ARGUMENT;
    AccWs w = acc_ws_layout(n_rows, dim);
    if (!workspace || workspace_bytes < w.total) return MORNA_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t s = (cudaStream_t)stream;
    MORNA_CUDA_TRY(cudaMemsetAsync(acc, 0, (size_t)dim * (size_t)acc_ld * sizeof(double), s));
    if (n_rows == 0 || id_hi == id_lo) return MORNA_OK;
    unsigned char *ws = (unsigned char *)workspace;
    auto *keys_in = (int32_t *)(ws + w.keys_in);
    auto *keys_out = (int32_t *)(ws + w.keys_out);
    auto *vals_in = (int32_t *)(ws + w.vals_in);
    auto *vals_out = (int32_t *)(ws + w.vals_out);
    auto *begin = (int32_t *)(ws + w.begin);
    bucket_keys_kernel<<<grid_for(n_rows, 256), 256, 0, s>>>(pass, bucket, n_rows, dim, keys_in, vals_in);
    MORNA_LAUNCH_CHECK();
    size_t cub_bytes = w.cub_bytes;
    // LSD radix sort is stable: rows of one bucket stay in file order
    MORNA_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws + w.cub, cub_bytes, keys_in, keys_out, vals_in, vals_out,
                                                   (int)n_rows, 0, bits_for((uint64_t)dim), s));
    count_launch(3);
    bucket_begin_kernel<<<grid_for(n_rows + 1, 256), 256, 0, s>>>(keys_out, n_rows, dim, begin);
    MORNA_LAUNCH_CHECK();
    const int32_t range = id_hi - id_lo;
    const int32_t tile = range < kAccTile ? range : kAccTile;
    const int32_t tiles = (range + tile - 1) / tile;
    const size_t smem = (size_t)tile * sizeof(double);
    if (smem > 48 * 1024)
        MORNA_CUDA_TRY(cudaFuncSetAttribute(index_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(kAccTile * sizeof(double))));
    dim3 grid((unsigned)dim, (unsigned)tiles);
    index_accumulate_kernel<<<grid, kAccThreads, smem, s>>>(row_off, sign, idf, sample, cov, id_of_sample, vals_out,
                                                          begin, id_lo, id_hi, tile, acc, acc_ld);
    MORNA_LAUNCH_CHECK();
    return MORNA_OK;
}

extern "C" int morna_round_store(const double *acc, int64_t acc_ld, int32_t n_ids, int32_t dim, float *vectors,
                                 int64_t ld, void *stream) {
    if (!acc || !vectors || n_ids < 0 || dim <= 0 || ld < dim || (ld & 3) || acc_ld < n_ids)
        return MORNA_ERR_INVALID_ARGUMENT;
    if (n_ids == 0) return MORNA_OK;
    dim3 grid((unsigned)((n_ids + 31) / 32), (unsigned)((ld + 31) / 32));
    if (grid.y > 65535) return MORNA_ERR_INVALID_ARGUMENT;
    round_store_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(acc, acc_ld, n_ids, dim, vectors, ld);
    MORNA_LAUNCH_CHECK();
    return MORNA_OK;
}

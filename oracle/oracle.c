/* TEST INFRASTRUCTURE ONLY -- C restatement of morna's hot path (CPU oracle).
 *
 * Same arithmetic as oracle/morna_oracle.py, which follows
 * /root/reference/morna.py line by line (Python floats are IEEE doubles, sums run
 * left to right).  Build with -ffp-contract=off and without -ffast-math so no
 * sum is reassociated or fused: outputs are bit-identical to the Python oracle.
 * Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Never linked into the product library.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

/* mmh3.hash(key) == MurmurHash3_x86_32(key, seed 0) as signed int32; morna.py:369 */
int32_t oracle_murmur3_32(const uint8_t *key, int32_t len, uint32_t seed) {
    const uint32_t c1 = 0xcc9e2d51u, c2 = 0x1b873593u;
    uint32_t h = seed;
    int32_t nblocks = len / 4;
    for (int32_t i = 0; i < nblocks; ++i) {
        uint32_t k;
        memcpy(&k, key + 4 * i, 4);          /* little-endian host */
        k *= c1; k = rotl32(k, 15); k *= c2;
        h ^= k; h = rotl32(h, 13); h = h * 5u + 0xe6546b64u;
    }
    const uint8_t *tail = key + 4 * nblocks;
    uint32_t k = 0;
    switch (len & 3) {
        case 3: k ^= (uint32_t)tail[2] << 16; /* fallthrough */
        case 2: k ^= (uint32_t)tail[1] << 8;  /* fallthrough */
        case 1: k ^= tail[0];
                k *= c1; k = rotl32(k, 15); k *= c2; h ^= k;
    }
    h ^= (uint32_t)len;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return (int32_t)h;
}

/* morna.py:369-371 for J packed keys: raw hash, Python floor-mod bucket, sign */
void oracle_hash_rows(const uint8_t *keys, const int32_t *key_off, int64_t n_rows,
                      int32_t dim, int32_t *raw, int32_t *bucket, int8_t *sign) {
    for (int64_t j = 0; j < n_rows; ++j) {
        int32_t h = oracle_murmur3_32(keys + key_off[j], key_off[j + 1] - key_off[j], 0);
        int32_t m = h % dim;
        if (m < 0) m += dim;
        raw[j] = h; bucket[j] = m; sign[j] = h < 0 ? -1 : 1;
    }
}

/* morna.py:101-114.  v: float32-valued stored row, q: query doubles. */
static double cosine_distance_qq(const float *v, const double *q, int32_t d,
                                 double qq, int clamp) {
    double pp = 0.0, pq = 0.0;
    for (int32_t i = 0; i < d; ++i) {
        double a = (double)v[i];
        pp += a * a;
        pq += a * q[i];
    }
    double ppqq = pp * qq, dist;
    if (ppqq > 0.0) dist = 2.0 - 2.0 * pq / sqrt(ppqq);
    else dist = 2.0;
    if (clamp && dist < 0.0) dist = 0.0;
    return sqrt(dist);   /* NaN where the reference raises ValueError (no clamp) */
}

static double sum_sq(const double *q, int32_t d) {
    double qq = 0.0;
    for (int32_t i = 0; i < d; ++i) qq += q[i] * q[i];
    return qq;
}

void oracle_distances(const float *S, int64_t n, int32_t d, int64_t ld,
                      const double *q, int clamp, double *out) {
    double qq = sum_sq(q, d);   /* same value the reference recomputes per row */
    for (int64_t i = 0; i < n; ++i) out[i] = cosine_distance_qq(S + i * ld, q, d, qq, clamp);
}

/* morna.py:697-712: scan in id order, bisect_left insert, truncate to k. */
int32_t oracle_exact_search(const float *S, int64_t n, int32_t d, int64_t ld,
                            const double *q, int32_t k, int clamp,
                            int32_t *ids, double *dists) {
    double qq = sum_sq(q, d);
    int32_t len = 0;
    for (int64_t i = 0; i < n; ++i) {
        double cur = cosine_distance_qq(S + i * ld, q, d, qq, clamp);
        int32_t lo = 0, hi = len;          /* bisect_left */
        while (lo < hi) {
            int32_t mid = (lo + hi) / 2;
            if (dists[mid] < cur) lo = mid + 1; else hi = mid;
        }
        if (lo < k) {
            int32_t last = len < k ? len : k - 1;   /* drop the k-th on overflow */
            for (int32_t t = last; t > lo; --t) { dists[t] = dists[t - 1]; ids[t] = ids[t - 1]; }
            dists[lo] = cur; ids[lo] = (int32_t)i;
            if (len < k) ++len;
        }
    }
    return len;
}

typedef struct {
    const float *S; int64_t n; int32_t d; int64_t ld;
    const double *Q; int64_t q_lo, q_hi; int32_t k; int clamp;
    int32_t *ids; double *dists;
} batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    for (int64_t qi = j->q_lo; qi < j->q_hi; ++qi)
        oracle_exact_search(j->S, j->n, j->d, j->ld, j->Q + qi * j->d, j->k, j->clamp,
                            j->ids + qi * j->k, j->dists + qi * j->k);
    return NULL;
}

/* The reference answers one query per process; a batch is nq independent runs.
 * Threads split the queries (the only parallelism the reference admits). */
void oracle_exact_search_batch(const float *S, int64_t n, int32_t d, int64_t ld,
                               const double *Q, int64_t nq, int32_t k, int clamp,
                               int32_t n_threads, int32_t *ids, double *dists) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > nq) n_threads = (int32_t)(nq > 0 ? nq : 1);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * n_threads);
    for (int32_t t = 0; t < n_threads; ++t) {
        batch_job jb = { S, n, d, ld, Q, nq * t / n_threads, nq * (t + 1) / n_threads,
                         k, clamp, ids, dists };
        jobs[t] = jb;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int32_t t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th); free(jobs);
}

/* morna.py:376-388 over pre-tokenised rows (CSR).  `pass[j]`, bucket/sign/idf per
 * row come from the caller (threshold :361, hash :369-371, idf :372-374).
 * id_of_sample: dense map sample id -> internal id, -1 = unseen; first-seen order.
 * acc: [capacity x dim] doubles, zero-initialised.  Returns number of ids assigned. */
int32_t oracle_index_accumulate(const int64_t *row_off, const uint8_t *pass,
                                const int32_t *bucket, const int8_t *sign,
                                const double *idf, int64_t n_rows,
                                const int32_t *sample, const int32_t *cov,
                                int32_t dim, int32_t *id_of_sample, int32_t next_id,
                                double *acc) {
    for (int64_t j = 0; j < n_rows; ++j) {
        if (!pass[j]) continue;
        double w = idf[j];
        int32_t b = bucket[j];
        double mult = (double)sign[j];
        for (int64_t p = row_off[j]; p < row_off[j + 1]; ++p) {
            int32_t s = sample[p];
            int32_t id = id_of_sample[s];
            if (id < 0) { id = next_id++; id_of_sample[s] = id; }
            double tfidf = (double)cov[p] * w;
            acc[(int64_t)id * dim + b] += mult * tfidf;
        }
    }
    return next_id;
}

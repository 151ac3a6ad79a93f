// Host tokenizer for intropolis rows: text -> the binary CSR the index kernels stream
// (packed junction keys + offsets, pair offsets, int32 sample ids / coverages).  Replaces the per-row
// Python of go_index's loop (morna.py:848-853) and the field scan of count_samples (:809-822) for
// well-formed rows; any row that is not plainly canonical (see flags) is left to the Python tokenizer,
// which has the reference's exact semantics (str.strip, str.split, int()).
//
//   tokens = line.strip().split("\t"); key = " ".join(tokens[:3])
//   samples = tokens[-2].split(","); coverages = tokens[-1].split(",")      (zip() -> min of the lengths)
//
// Multithreaded: the buffer is cut at line boundaries into one slab per thread; a counting pass sizes
// the outputs, a prefix over the slabs places them, a filling pass writes them.
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/morna_b200.h"

namespace {

struct SlabCount { int64_t rows = 0, key_bytes = 0, pairs = 0; };

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// One line [p, e) without its '\n'.  Returns false if the row must go to the Python tokenizer.
// On success: key fields [k0,k1) [k2,k3) [k4,k5), sample list [s0,s1), coverage list [c0,c1).
struct LineView { const char *k[6]; const char *s0, *s1, *c0, *c1; };

inline bool split_line(const char *p, const char *e, LineView &v) {
    if (e > p && e[-1] == '\r') --e;                         // strip() removes it
    if (p == e) return false;
    // leading / trailing whitespace changes what strip() + split("\t") sees: leave those rows to Python
    const char first = *p, last = e[-1];
    if (first == ' ' || first == '\t' || first == '\v' || first == '\f' || last == ' ' || last == '\t' || last == '\v' ||
        last == '\f' || last == '\r')
        return false;
    const char *tabs[4] = {nullptr, nullptr, nullptr, nullptr};    // first three tabs, and we track the last two
    const char *last_tab = nullptr, *prev_tab = nullptr;
    int n_tabs = 0;
    for (const char *q = p; q < e; ++q) {
        if (*q == '\t') {
            if (n_tabs < 3) tabs[n_tabs] = q;
            prev_tab = last_tab; last_tab = q;
            ++n_tabs;
        }
    }
    if (n_tabs < 4) return false;                            // fewer than five fields: let Python raise what it raises
    v.k[0] = p; v.k[1] = tabs[0]; v.k[2] = tabs[0] + 1; v.k[3] = tabs[1]; v.k[4] = tabs[1] + 1; v.k[5] = tabs[2];
    v.s0 = prev_tab + 1; v.s1 = last_tab; v.c0 = last_tab + 1; v.c1 = e;
    return true;
}

// comma-separated canonical non-negative decimal integers (no sign, no spaces, no leading zeros, < 2^31);
// returns the count, or -1 if anything else appears.  out may be null (counting pass).
inline int64_t parse_list(const char *p, const char *e, int32_t *out) {
    int64_t n = 0;
    while (true) {
        if (p >= e || !is_digit(*p)) return -1;
        if (*p == '0' && p + 1 < e && is_digit(p[1])) return -1;      // "007" != "7" as a count_samples string
        int64_t val = 0;
        int digits = 0;
        while (p < e && is_digit(*p)) { val = val * 10 + (*p - '0'); ++p; if (++digits > 10) return -1; }
        if (val > 0x7fffffffLL) return -1;
        if (out) out[n] = (int32_t)val;
        ++n;
        if (p == e) return n;
        if (*p != ',') return -1;
        ++p;
    }
}

template <typename F>
void for_each_line(const char *p, const char *e, F &&f) {
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = nl ? nl : e;
        f(p, le);
        p = nl ? nl + 1 : e;
    }
}

std::vector<const char *> slab_bounds(const char *text, size_t nbytes, int n_threads) {
    std::vector<const char *> b;
    const char *end = text + nbytes;
    b.push_back(text);
    for (int t = 1; t < n_threads; ++t) {
        const char *guess = text + nbytes * (size_t)t / (size_t)n_threads;
        if (guess <= b.back()) guess = b.back();
        const char *nl = guess < end ? (const char *)memchr(guess, '\n', (size_t)(end - guess)) : nullptr;
        b.push_back(nl ? nl + 1 : end);
    }
    b.push_back(end);
    return b;
}

}  // namespace

extern "C" int morna_tokenize_count(const char *text, size_t nbytes, int32_t n_threads, int64_t *n_rows,
                                    int64_t *key_bytes, int64_t *n_pairs) {
    if (!text || !n_rows || !key_bytes || !n_pairs || n_threads <= 0) return MORNA_ERR_INVALID_ARGUMENT;
    if (n_threads > 256) n_threads = 256;
    const auto bounds = slab_bounds(text, nbytes, n_threads);
    std::vector<SlabCount> counts((size_t)n_threads);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&, t]() {
            SlabCount c;
            for_each_line(bounds[(size_t)t], bounds[(size_t)t + 1], [&](const char *p, const char *e) {
                ++c.rows;
                LineView v;
                if (!split_line(p, e, v)) return;
                const int64_t ns = parse_list(v.s0, v.s1, nullptr), nc = parse_list(v.c0, v.c1, nullptr);
                if (ns < 0 || ns != nc) return;                  // unequal lists: zip() and count_samples differ -> Python
                c.key_bytes += (v.k[1] - v.k[0]) + (v.k[3] - v.k[2]) + (v.k[5] - v.k[4]) + 2;
                c.pairs += ns;
            });
            counts[(size_t)t] = c;
        });
    for (auto &th : pool) th.join();
    *n_rows = *key_bytes = *n_pairs = 0;
    for (const auto &c : counts) { *n_rows += c.rows; *key_bytes += c.key_bytes; *n_pairs += c.pairs; }
    return MORNA_OK;
}

extern "C" int morna_tokenize_fill(const char *text, size_t nbytes, int32_t n_threads, uint8_t *keys, int32_t *key_off,
                                   int64_t *row_off, int32_t *sample, int32_t *cov, int64_t *line_off,
                                   uint8_t *needs_python) {
    if (!text || !key_off || !row_off || !line_off || !needs_python || n_threads <= 0) return MORNA_ERR_INVALID_ARGUMENT;
    if (n_threads > 256) n_threads = 256;
    const auto bounds = slab_bounds(text, nbytes, n_threads);
    // pass A: per-slab sizes (same rules as morna_tokenize_count), then exclusive prefixes
    std::vector<SlabCount> counts((size_t)n_threads);
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t)
            pool.emplace_back([&, t]() {
                SlabCount c;
                for_each_line(bounds[(size_t)t], bounds[(size_t)t + 1], [&](const char *p, const char *e) {
                    ++c.rows;
                    LineView v;
                    if (!split_line(p, e, v)) return;
                    const int64_t ns = parse_list(v.s0, v.s1, nullptr), nc = parse_list(v.c0, v.c1, nullptr);
                    if (ns < 0 || ns != nc) return;
                    c.key_bytes += (v.k[1] - v.k[0]) + (v.k[3] - v.k[2]) + (v.k[5] - v.k[4]) + 2;
                    c.pairs += ns;
                });
                counts[(size_t)t] = c;
            });
        for (auto &th : pool) th.join();
    }
    std::vector<SlabCount> base((size_t)n_threads + 1);
    for (int t = 0; t < n_threads; ++t) {
        base[(size_t)t + 1].rows = base[(size_t)t].rows + counts[(size_t)t].rows;
        base[(size_t)t + 1].key_bytes = base[(size_t)t].key_bytes + counts[(size_t)t].key_bytes;
        base[(size_t)t + 1].pairs = base[(size_t)t].pairs + counts[(size_t)t].pairs;
    }
    if (base[(size_t)n_threads].key_bytes > 0x7fffffffLL) return MORNA_ERR_INVALID_ARGUMENT;      // key offsets are int32
    // pass B: fill
    std::vector<std::thread> pool;
    std::vector<int32_t> scratch_fail((size_t)n_threads, 0);
    for (int t = 0; t < n_threads; ++t)
        pool.emplace_back([&, t]() {
            int64_t r = base[(size_t)t].rows, kb = base[(size_t)t].key_bytes, pr = base[(size_t)t].pairs;
            for_each_line(bounds[(size_t)t], bounds[(size_t)t + 1], [&](const char *p, const char *e) {
                line_off[r] = p - text;
                key_off[r] = (int32_t)kb;
                row_off[r] = pr;
                needs_python[r] = 1;
                LineView v;
                if (split_line(p, e, v)) {
                    // parse into place; a malformed or unequal pair of lists leaves the row empty and flagged
                    const int64_t ns = parse_list(v.s0, v.s1, nullptr), nc = parse_list(v.c0, v.c1, nullptr);
                    if (ns >= 0 && ns == nc) {
                        const int64_t n = ns;
                        parse_list(v.s0, v.s1, sample + pr);
                        parse_list(v.c0, v.c1, cov + pr);
                        for (int f = 0; f < 3; ++f) {
                            const size_t len = (size_t)(v.k[2 * f + 1] - v.k[2 * f]);
                            memcpy(keys + kb, v.k[2 * f], len);
                            kb += (int64_t)len;
                            if (f < 2) keys[kb++] = ' ';
                        }
                        pr += n;
                        needs_python[r] = 0;
                    }
                }
                ++r;
            });
            if (r != base[(size_t)t + 1].rows || kb != base[(size_t)t + 1].key_bytes || pr != base[(size_t)t + 1].pairs)
                scratch_fail[(size_t)t] = 1;
        });
    for (auto &th : pool) th.join();
    for (int32_t f : scratch_fail) if (f) return MORNA_ERR_INVALID_ARGUMENT;
    const int64_t rows = base[(size_t)n_threads].rows;
    key_off[rows] = (int32_t)base[(size_t)n_threads].key_bytes;
    row_off[rows] = base[(size_t)n_threads].pairs;
    line_off[rows] = (int64_t)nbytes;
    return MORNA_OK;
}

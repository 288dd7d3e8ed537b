/*
 * hkcsa_oracle.c -- CPU restatement of the reference's index-build and
 * backward-search path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this library.  The product path
 * (high-order-entropy-compressed-suffix-array_b200/) never links, imports or
 * executes anything in oracle/.
 *
 * Parity pin: the reference (ajaynair710/High-Order-Entropy-Compressed-Suffix-Array)
 * is pure Python and holds no asserting tests or golden vectors of its own
 * (SURVEY.md section 4).  This restatement is therefore pinned against outputs
 * of the reference ITSELF, executed in the authoring container by
 * tests/golden/make_golden.py and frozen under tests/golden/ (the script is
 * committed beside the fixtures); tests/test_oracle_golden.py re-checks every
 * function below against those fixtures on every CPU test run.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference repository root).  Symbols are bytes (the reference decodes
 * its corpora as latin-1, utils/data_loader.py:4, so one code point == one
 * byte) and all ordering is unsigned byte order == Python str code-point order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HKO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Workload synthesis (no reference counterpart: Pizza&Chili is unreachable    */
/* offline, tests/dataset_benchmark.py:10-16).  Integer-only, counter-based so */
/* that the CUDA generator (csrc/textgen.cu) reproduces the same bytes.        */
/* ------------------------------------------------------------------------- */

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

#define HKO_CHUNK 65536ULL

/* ENG96 alphabet: 0x20..0x7E without '$' (94 symbols), then '\n', '\t'. */
static void eng96_alphabet(uint8_t a[96])
{
    int k = 0;
    for (int c = 0x20; c <= 0x7E; ++c)
        if (c != 0x24) a[k++] = (uint8_t)c;
    a[k++] = 0x0A;
    a[k++] = 0x09;
}

static const uint8_t PERM4[24][4] = {
    {0,1,2,3},{0,1,3,2},{0,2,1,3},{0,2,3,1},{0,3,1,2},{0,3,2,1},
    {1,0,2,3},{1,0,3,2},{1,2,0,3},{1,2,3,0},{1,3,0,2},{1,3,2,0},
    {2,0,1,3},{2,0,3,1},{2,1,0,3},{2,1,3,0},{2,3,0,1},{2,3,1,0},
    {3,0,1,2},{3,0,2,1},{3,1,0,2},{3,1,2,0},{3,2,0,1},{3,2,1,0}};

/* kind 0 = ENG96 (order-3 Markov, 6 Zipf(1) successors per context),
 * kind 1 = DNA4  (order-5 Markov over ACGT, weights (8,4,2,2)/16 permuted per
 * context).  Context resets at every 64 KiB chunk so chunks are independent. */
HKO_API int hko_gen_text(int kind, uint64_t seed, uint64_t n, uint8_t *out)
{
    const uint64_t s1 = splitmix64(seed);
    const uint64_t s2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ULL);
    uint8_t alpha[96];
    eng96_alphabet(alpha);
    const uint64_t nchunks = (n + HKO_CHUNK - 1) / HKO_CHUNK;
    if (kind != 0 && kind != 1) return -1;
#pragma omp parallel for schedule(dynamic, 4)
    for (uint64_t q = 0; q < nchunks; ++q) {
        uint64_t lo = q * HKO_CHUNK, hi = lo + HKO_CHUNK;
        if (hi > n) hi = n;
        if (kind == 0) {
            uint32_t c1 = 0, c2 = 0, c3 = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                uint64_t r = splitmix64(s1 + i);
                uint64_t ctx = (uint64_t)c1 + 96u * c2 + 9216u * c3;
                uint64_t h = splitmix64(s2 ^ ctx);
                uint32_t u = (uint32_t)(r % 147u);
                int k = (u < 60) ? 0 : (u < 90) ? 1 : (u < 110) ? 2 : (u < 125) ? 3 : (u < 137) ? 4 : 5;
                uint32_t id = (uint32_t)((h >> (10 * k)) & 1023u) % 96u;
                out[i] = alpha[id];
                c3 = c2; c2 = c1; c1 = id;
            }
        } else {
            uint32_t ctx = 0;
            for (uint64_t i = lo; i < hi; ++i) {
                uint64_t r = splitmix64(s1 + i);
                uint32_t pidx = (uint32_t)(splitmix64(s2 ^ (uint64_t)ctx) % 24u);
                uint32_t r4 = (uint32_t)(r & 15u);
                int slot = (r4 < 8) ? 0 : (r4 < 12) ? 1 : (r4 < 14) ? 2 : 3;
                uint32_t id = PERM4[pidx][slot];
                out[i] = (uint8_t)"ACGT"[id];
                ctx = ((ctx << 2) | id) & 1023u;
            }
        }
    }
    return 0;
}

/* Pattern workload (semantics of tests/test_patterns.py:3-9: substrings of the
 * text at random offsets), seeded.  Length uniform in [min_len,max_len]; odd
 * r3 => one symbol substituted by another alphabet symbol (early-miss path).
 * `alpha`/`sigma` = sorted distinct symbols of the text.  Two-phase: lengths
 * first (caller prefix-sums them into offsets[P+1]), then bytes. */
HKO_API void hko_pattern_lengths(uint64_t seed, uint64_t P, uint32_t min_len, uint32_t max_len,
                                 uint64_t n, uint32_t *len_out)
{
    const uint64_t s1 = splitmix64(seed);
    for (uint64_t p = 0; p < P; ++p) {
        uint64_t r1 = splitmix64(s1 + 3 * p);
        uint32_t len = min_len + (uint32_t)(r1 % (uint64_t)(max_len - min_len + 1));
        if (len > n) len = (uint32_t)n;
        len_out[p] = len;
    }
}

HKO_API void hko_pattern_fill(uint64_t seed, uint64_t P, const uint8_t *text, uint64_t n,
                              const uint8_t *alpha, uint32_t sigma, const int64_t *offsets,
                              uint8_t *out)
{
    const uint64_t s1 = splitmix64(seed);
#pragma omp parallel for schedule(static)
    for (uint64_t p = 0; p < P; ++p) {
        uint64_t r2 = splitmix64(s1 + 3 * p + 1);
        uint64_t r3 = splitmix64(s1 + 3 * p + 2);
        uint32_t len = (uint32_t)(offsets[p + 1] - offsets[p]);
        uint64_t start = r2 % (n - len + 1);
        uint8_t *dst = out + offsets[p];
        memcpy(dst, text + start, len);
        if ((r3 & 1u) && len > 0 && sigma > 1) {
            uint32_t at = (uint32_t)((r3 >> 1) % len);
            uint32_t pick = (uint32_t)((r3 >> 32) % sigma);
            if (alpha[pick] == dst[at]) pick = (pick + 1) % sigma;
            dst[at] = alpha[pick];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* a1: build_suffix_array -- csa/suffix_array.py:131-134                       */
/*   suffixes = [(text[i:], i) ...]; suffixes.sort(); return indices.          */
/* Python compares the suffix strings by code point; a proper prefix sorts     */
/* first; ties are impossible (distinct suffixes have distinct lengths).       */
/* Restated as a comparison sort of suffix start positions with exactly that   */
/* comparator -- the same algorithm minus the O(n^2) copies of :132.           */
/* ------------------------------------------------------------------------- */

static const uint8_t *g_text;
static uint64_t g_n;

static inline int suffix_less_eq_cmp(uint32_t a, uint32_t b)
{
    uint64_t la = g_n - a, lb = g_n - b;
    uint64_t m = la < lb ? la : lb;
    int c = memcmp(g_text + a, g_text + b, m);
    if (c) return c;
    return (la < lb) ? -1 : (la > lb);
}

static int suffix_cmp_q(const void *pa, const void *pb)
{
    return suffix_less_eq_cmp(*(const uint32_t *)pa, *(const uint32_t *)pb);
}

static void merge_runs(const uint32_t *src, uint32_t *dst, uint64_t lo, uint64_t mid, uint64_t hi)
{
    uint64_t i = lo, j = mid, k = lo;
    while (i < mid && j < hi)
        dst[k++] = (suffix_less_eq_cmp(src[j], src[i]) < 0) ? src[j++] : src[i++];
    while (i < mid) dst[k++] = src[i++];
    while (j < hi) dst[k++] = src[j++];
}

/* threads <= 0: use every core OpenMP reports.  Returns threads used. */
HKO_API int hko_sa_build(const uint8_t *text, uint64_t n, uint32_t *sa, int threads)
{
    g_text = text;
    g_n = n;
    for (uint64_t i = 0; i < n; ++i) sa[i] = (uint32_t)i;
    int T = 1;
#ifdef _OPENMP
    T = threads > 0 ? threads : omp_get_max_threads();
#endif
    /* number of initial runs: power of two >= T, but never more than n */
    uint64_t runs = 1;
    while (runs < (uint64_t)T) runs <<= 1;
    while (runs > 1 && n / runs < 1024) runs >>= 1;
    if (n < 2) return 1;
    if (runs == 1) {
        qsort(sa, n, sizeof(uint32_t), suffix_cmp_q);
        return 1;
    }
    uint64_t *bound = (uint64_t *)malloc((runs + 1) * sizeof(uint64_t));
    for (uint64_t r = 0; r <= runs; ++r) bound[r] = n * r / runs;
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
    for (uint64_t r = 0; r < runs; ++r)
        qsort(sa + bound[r], bound[r + 1] - bound[r], sizeof(uint32_t), suffix_cmp_q);
    uint32_t *tmp = (uint32_t *)malloc(n * sizeof(uint32_t));
    uint32_t *src = sa, *dst = tmp;
    for (uint64_t width = 1; width < runs; width <<= 1) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
        for (uint64_t r = 0; r < runs; r += 2 * width)
            merge_runs(src, dst, bound[r], bound[r + width], bound[r + 2 * width]);
        uint32_t *t = src; src = dst; dst = t;
    }
    if (src != sa) memcpy(sa, src, n * sizeof(uint32_t));
    free(tmp);
    free(bound);
    return T;
}

/* ------------------------------------------------------------------------- */
/* a3: bwt_transform -- csa/bwt.py:3-13                                        */
/*   pos = SA[i]-1; if pos < 0: pos = n-1; bwt[i] = text[pos]                  */
/* ------------------------------------------------------------------------- */
HKO_API void hko_bwt(const uint8_t *text, const uint32_t *sa, uint64_t n, uint8_t *bwt)
{
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < n; ++i) {
        int64_t pos = (int64_t)sa[i] - 1;
        if (pos < 0) pos = (int64_t)n - 1;
        bwt[i] = text[pos];
    }
}

/* ------------------------------------------------------------------------- */
/* a4: build_count -- utils/utils.py:16-24                                     */
/*   C[c] = number of symbols with code point < c, for c present in the text.  */
/*   Output: cnt[256] raw histogram, C[256] exclusive prefix over byte order   */
/*   (entries of absent bytes are still filled; the dict view keeps present).  */
/* ------------------------------------------------------------------------- */
HKO_API void hko_count_table(const uint8_t *text, uint64_t n, uint64_t cnt[256], uint64_t C[256])
{
    memset(cnt, 0, 256 * sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) cnt[text[i]]++;
    uint64_t total = 0;
    for (int c = 0; c < 256; ++c) {
        C[c] = total;
        total += cnt[c];
    }
}

/* ------------------------------------------------------------------------- */
/* a5: build_occ -- utils/utils.py:26-32                                       */
/*   occ[c][i] = #c in bwt[0:i], i in [0,n].  Dense form for one symbol        */
/*   (out has n+1 entries) and a checkpointed form for backward search at      */
/*   sizes where sigma*(n+1) integers cannot exist.                            */
/* ------------------------------------------------------------------------- */
HKO_API void hko_occ_dense(const uint8_t *bwt, uint64_t n, uint8_t c, uint32_t *out)
{
    uint32_t acc = 0;
    out[0] = 0;
    for (uint64_t i = 0; i < n; ++i) {
        acc += (bwt[i] == c);
        out[i + 1] = acc;
    }
}

#define HKO_OCC_STEP 128ULL

typedef struct {
    const uint8_t *bwt;
    uint64_t n;
    uint64_t C[256];
    uint64_t cnt[256];
    int slot[256];   /* byte -> dense slot or -1 (absent: "char not in occ") */
    int sigma;
    uint32_t *ckpt;  /* [(n/STEP)+1][sigma] */
} hko_fm;

HKO_API hko_fm *hko_fm_new(const uint8_t *bwt, uint64_t n)
{
    hko_fm *f = (hko_fm *)calloc(1, sizeof(hko_fm));
    f->bwt = bwt;
    f->n = n;
    hko_count_table(bwt, n, f->cnt, f->C);
    f->sigma = 0;
    for (int c = 0; c < 256; ++c) f->slot[c] = f->cnt[c] ? f->sigma++ : -1;
    uint64_t nck = n / HKO_OCC_STEP + 1;
    f->ckpt = (uint32_t *)malloc(nck * (uint64_t)(f->sigma ? f->sigma : 1) * sizeof(uint32_t));
    uint32_t run[256];
    memset(run, 0, sizeof(run));
    for (uint64_t i = 0; i <= n; ++i) {
        if (i % HKO_OCC_STEP == 0)
            for (int c = 0; c < 256; ++c)
                if (f->slot[c] >= 0) f->ckpt[(i / HKO_OCC_STEP) * f->sigma + f->slot[c]] = run[c];
        if (i < n) run[bwt[i]]++;
    }
    return f;
}

HKO_API void hko_fm_free(hko_fm *f)
{
    if (!f) return;
    free(f->ckpt);
    free(f);
}

/* EnhancedFMIndex.rank -- csa/enhanced_fm_index.py:34-40 (occ[c][index],
 * 0 for a symbol that never occurs, index clamped to n). */
static inline uint64_t fm_rank(const hko_fm *f, uint8_t c, uint64_t i)
{
    if (f->slot[c] < 0) return 0;
    if (i > f->n) i = f->n;
    uint64_t b = i / HKO_OCC_STEP;
    uint64_t r = f->ckpt[b * f->sigma + f->slot[c]];
    for (uint64_t j = b * HKO_OCC_STEP; j < i; ++j) r += (f->bwt[j] == c);
    return r;
}

HKO_API uint64_t hko_fm_rank(const hko_fm *f, uint8_t c, uint64_t i) { return fm_rank(f, c, i); }

/* ------------------------------------------------------------------------- */
/* a12: EnhancedFMIndex.find_range -- csa/enhanced_fm_index.py:21-32           */
/*   l, r = 0, n-1; for char in reversed(query):                               */
/*     new_l = rank(char, l) + C.get(char, 0)                                  */
/*     new_r = rank(char, r+1) + C.get(char, 0) - 1                            */
/*     if new_l > new_r: return (-1,-1)                                        */
/* ------------------------------------------------------------------------- */
HKO_API void hko_find_range(const hko_fm *f, const uint8_t *pat, uint64_t m, int64_t *lo, int64_t *hi)
{
    int64_t l = 0, r = (int64_t)f->n - 1;
    for (uint64_t k = m; k-- > 0;) {
        uint8_t c = pat[k];
        int64_t Cc = (f->slot[c] >= 0) ? (int64_t)f->C[c] : 0;
        int64_t nl = (int64_t)fm_rank(f, c, (uint64_t)l) + Cc;
        int64_t nr = (int64_t)fm_rank(f, c, (uint64_t)(r + 1)) + Cc - 1;
        if (nl > nr) { *lo = -1; *hi = -1; return; }
        l = nl; r = nr;
    }
    *lo = l; *hi = r;
}

HKO_API void hko_find_range_batch(const hko_fm *f, const uint8_t *pats, const int64_t *off, uint64_t P,
                                  int64_t *lo, int64_t *hi, int threads)
{
    int T = 1;
#ifdef _OPENMP
    T = threads > 0 ? threads : omp_get_max_threads();
#endif
    (void)T;
#pragma omp parallel for schedule(dynamic, 256) num_threads(T)
    for (uint64_t p = 0; p < P; ++p)
        hko_find_range(f, pats + off[p], (uint64_t)(off[p + 1] - off[p]), &lo[p], &hi[p]);
}

/* ------------------------------------------------------------------------- */
/* a7: WaveletTree.build_tree -- csa/wavelet_tree.py:72-100 (left spine only)  */
/*   alphabet = sorted(set(text)); while len(alphabet) > 1:                    */
/*     mid = len//2; bitmap[j] = current[j] in alphabet[mid:];                 */
/*     next = [c for c in current if c in alphabet[:mid]]; alphabet = left     */
/* Writes every level's bitmap (one byte per bit) back to back into `bits`     */
/* (caller sizes it levels*n worst case; in fact sum of level lengths) and     */
/* level_len[l].  Returns the number of levels.                                */
/* ------------------------------------------------------------------------- */
HKO_API int hko_wt_spine(const uint8_t *text, uint64_t n, uint8_t *bits, uint64_t *level_len,
                         uint8_t *alpha_out, int *sigma_out)
{
    uint64_t cnt[256] = {0};
    for (uint64_t i = 0; i < n; ++i) cnt[text[i]]++;
    uint8_t alpha[256];
    int sigma = 0;
    for (int c = 0; c < 256; ++c)
        if (cnt[c]) alpha[sigma++] = (uint8_t)c;
    if (alpha_out) memcpy(alpha_out, alpha, (size_t)sigma);
    if (sigma_out) *sigma_out = sigma;
    uint8_t *cur = (uint8_t *)malloc(n ? n : 1);
    memcpy(cur, text, n);
    uint64_t cur_n = n, w = 0;
    int len = sigma, levels = 0;
    while (len > 1) {
        int mid = len / 2;
        uint8_t is_right[256] = {0};
        for (int k = mid; k < len; ++k) is_right[alpha[k]] = 1;
        uint64_t nn = 0;
        for (uint64_t j = 0; j < cur_n; ++j) {
            uint8_t b = is_right[cur[j]];
            bits[w++] = b;
            if (!b) cur[nn++] = cur[j];
        }
        level_len[levels++] = cur_n;
        cur_n = nn;
        len = mid;
    }
    free(cur);
    return levels;
}

/* a6: SuccinctRankSelect -- csa/wavelet_tree.py:5-25.
 *   rank_support[i] = #1 in bits[0:i] (uint32, n+1 entries);
 *   select(k): binary search for the smallest p in [0,n] with rank(p) >= k. */
HKO_API void hko_rank_support(const uint8_t *bits, uint64_t n, uint32_t *rs)
{
    rs[0] = 0;
    for (uint64_t i = 1; i <= n; ++i) rs[i] = rs[i - 1] + bits[i - 1];
}

HKO_API uint64_t hko_select(const uint32_t *rs, uint64_t n, uint64_t k)
{
    uint64_t low = 0, high = n;
    while (low < high) {
        uint64_t mid = (low + high) / 2;
        if (rs[mid] < k) low = mid + 1; else high = mid;
    }
    return low;
}

/* a8: GolombRiceEncoder -- csa/wavelet_tree.py:27-63.
 *   m = max(1, int(log2(total/ones))), 1 when ones == 0  (:33-38).  Integer
 *   restatement: the largest k with ones * 2^k <= total, floored at 1.
 *   encode (:40-63): every maximal run of ones of length v, closed by a zero
 *   or by end of input, emits v//m zeros, a one, then v%m as exactly m binary
 *   digits MSB first.  Runs of zeros emit nothing.
 *   Returns the number of code bits written (one byte per bit); pass out=NULL
 *   to size the buffer. */
HKO_API uint32_t hko_golomb_m(uint64_t ones, uint64_t total)
{
    if (ones == 0) return 1;
    uint32_t k = 0;
    while (k < 62 && (ones << (k + 1)) <= total) ++k;
    return k < 1 ? 1 : k;
}

HKO_API uint64_t hko_golomb_encode(const uint8_t *bits, uint64_t n, uint32_t m, uint8_t *out)
{
    uint64_t w = 0, q = 0;
    for (uint64_t i = 0; i <= n; ++i) {
        int bit = (i < n) ? bits[i] : 0;
        if (bit) { q++; continue; }
        if (q > 0) {
            uint64_t quo = q / m, rem = q % m;
            if (out) { memset(out + w, 0, quo); out[w + quo] = 1; }
            w += quo + 1;
            for (uint32_t b = 0; b < m; ++b) {
                uint32_t shift = m - 1 - b;
                uint8_t d = (shift < 64) ? (uint8_t)((rem >> shift) & 1u) : 0;
                if (out) out[w] = d;
                ++w;
            }
        }
        q = 0;
    }
    return w;
}

/* a14: FMIndex.precompute_rank -- csa/csa.py:13-19 == main.py:13-19:
 *   rank[c] = ascending positions of c in the BWT.  Concatenated over byte
 *   order this is a stable counting sort of positions by symbol. */
HKO_API void hko_symbol_positions(const uint8_t *bwt, uint64_t n, uint32_t *pos_out, uint64_t start[257])
{
    uint64_t cnt[256] = {0};
    for (uint64_t i = 0; i < n; ++i) cnt[bwt[i]]++;
    uint64_t t = 0;
    for (int c = 0; c < 256; ++c) { start[c] = t; t += cnt[c]; }
    start[256] = t;
    uint64_t fill[256];
    memcpy(fill, start, sizeof(fill));
    for (uint64_t i = 0; i < n; ++i) pos_out[fill[bwt[i]]++] = (uint32_t)i;
}

/* oracle.c -- CPU restatement of the reference algorithms on Trinity's k-mer hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product (libtrinity_gpu + the three executables)
 * never links or calls anything under oracle/.
 *
 * Parity status
 *   S (fastaToKmerCoverageStats) and R (ReadsToTranscripts): PINNED -- tests/test_oracle_golden.py checks this
 *     file against outputs of the unmodified reference binaries (oracle/_ref, built by oracle/Makefile.ref
 *     straight from /root/reference) committed under tests/golden/, and against the survey's md5 vectors.
 *   J (jellyfish count/dump/histo): "parity unpinned" against real jellyfish -- gmarcais/Jellyfish 2.3.0
 *     (Docker/Dockerfile:179) is a third-party binary that is neither vendored in the reference tree nor
 *     installed here.  orc_jf_* restates its published semantics and is anchored on the
 *     reference's own equivalent counter: orc_jf_count must agree with the Inchworm KmerCounter restatement
 *     (and, through `fastaToKmerCoverageStats --kmers <dump>` vs `--kmers_from_reads`, with the real
 *     reference binary) -- see tests/test_oracle_golden.py::test_dump_feeds_reference_stats.  The one real jellyfish
 *     dump in the reference tree (trinity_ext_sample_data/test_Inchworm/jellyfish.kmers.fa.gz, head committed under
 *     tests/golden/) additionally pins the dump format, the printed representative (lexicographic minimum of k-mer
 *     and reverse complement) and the --kmers loader: ::test_real_jellyfish_dump_fixture_pins_*.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -march, like the reference's -O2 build, so fp32
 * expressions are evaluated exactly as in Inchworm/Chrysalis: no FMA, one rounding per operation).
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------ */
/* small open-addressing map u64 -> u32 (stands in for __gnu_cxx::hash_map<kmer_int_type_t,unsigned int>,   */
/* Inchworm/src/KmerCounter.hpp:55, and for NonRedKmerTable's sorted vector + BinSearch)                   */
/* ------------------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t* keys;
    uint32_t* vals;
    uint8_t* used;
    uint64_t cap, n;
} orc_map;

static uint64_t mixh(uint64_t x) {
    x ^= x >> 31; x *= 0x7fb5d329728ea185ULL; x ^= x >> 27; x *= 0x81dadef4bc2dd44dULL; x ^= x >> 33;
    return x;
}
static void map_init(orc_map* m, uint64_t cap) {
    m->cap = cap; m->n = 0;
    m->keys = (uint64_t*)malloc(cap * 8);
    m->vals = (uint32_t*)calloc(cap, 4);
    m->used = (uint8_t*)calloc(cap, 1);
}
static void map_free(orc_map* m) { free(m->keys); free(m->vals); free(m->used); }
static uint32_t* map_slot(orc_map* m, uint64_t key, int insert);
static void map_grow(orc_map* m) {
    orc_map o = *m;
    map_init(m, o.cap * 2);
    for (uint64_t i = 0; i < o.cap; i++)
        if (o.used[i]) *map_slot(m, o.keys[i], 1) = o.vals[i];
    map_free(&o);
}
static uint32_t* map_slot(orc_map* m, uint64_t key, int insert) {
    if (insert && (m->n + 1) * 10 > m->cap * 6) map_grow(m);
    uint64_t i = mixh(key) & (m->cap - 1);
    while (m->used[i]) {
        if (m->keys[i] == key) return &m->vals[i];
        i = (i + 1) & (m->cap - 1);
    }
    if (!insert) return NULL;
    m->used[i] = 1; m->keys[i] = key; m->vals[i] = 0; m->n++;
    return &m->vals[i];
}

/* ------------------------------------------------------------------------------------------------------ */
/* Inchworm sequenceUtil (Inchworm/src/sequenceUtil.cpp)                                                   */
/* ------------------------------------------------------------------------------------------------------ */
/* _base_to_int: G=0 A=1 T=2 C=3, both cases, everything else > 3   (sequenceUtil.cpp:10-25) */
static int iw_base(unsigned char c) {
    switch (c) {
        case 'G': case 'g': return 0;
        case 'A': case 'a': return 1;
        case 'T': case 't': return 2;
        case 'C': case 'c': return 3;
        default: return 255;
    }
}
/* contains_non_gatc (sequenceUtil.cpp:30-50) */
static int iw_contains_non_gatc(const char* s, int k) {
    for (int i = 0; i < k; i++) if (iw_base((unsigned char)s[i]) > 3) return 1;
    return 0;
}
/* kmer_to_intval (sequenceUtil.cpp:258-296): first base most significant */
static uint64_t iw_kmer_to_intval(const char* s, int k) {
    uint64_t v = 0;
    for (int i = 0; i < k; i++) { v <<= 2; v |= (uint64_t)iw_base((unsigned char)s[i]); }
    return v;
}
/* revcomp_val (sequenceUtil.cpp:181-195): complement = ~, then reverse the 2-bit fields */
static uint64_t iw_revcomp_val(uint64_t kmer, int k) {
    uint64_t rev = 0;
    kmer = ~kmer;
    for (int i = 0; i < k; i++) { rev = (rev << 2) + (kmer & 3); kmer >>= 2; }
    return rev;
}
/* get_DS_kmer_val (sequenceUtil.cpp:376-385): the LARGER of kmer / revcomp under G<A<T<C */
static uint64_t iw_ds_val(uint64_t v, int k) {
    uint64_t r = iw_revcomp_val(v, k);
    return r > v ? r : v;
}

/* ------------------------------------------------------------------------------------------------------ */
/* KmerCounter (Inchworm/src/KmerCounter.cpp)                                                              */
/* ------------------------------------------------------------------------------------------------------ */
typedef struct { orc_map m; int k; int ds; } orc_kc;

orc_kc* orc_kc_new(int k, int ds) {
    orc_kc* kc = (orc_kc*)malloc(sizeof *kc);
    map_init(&kc->m, 1u << 16);
    kc->k = k; kc->ds = ds;
    return kc;
}
void orc_kc_free(orc_kc* kc) { if (kc) { map_free(&kc->m); free(kc); } }
uint64_t orc_kc_size(orc_kc* kc) { return kc->m.n; }

/* add_kmer(kmer_int_type_t, count)  KmerCounter.cpp:476-489: DS -> canonical, then map[k] += count */
static void kc_add_val(orc_kc* kc, uint64_t v, uint32_t count) {
    if (kc->ds) v = iw_ds_val(v, kc->k);
    *map_slot(&kc->m, v, 1) += count;
}
/* add_kmer(string, count)  KmerCounter.cpp:493-505: k-mers with a non-GATC character are not stored */
void orc_kc_add_kmer_str(orc_kc* kc, const char* kmer, uint32_t count) {
    if (iw_contains_non_gatc(kmer, kc->k)) return;
    kc_add_val(kc, iw_kmer_to_intval(kmer, kc->k), count);
}
/* add_sequence  KmerCounter.cpp:34-44: every window */
void orc_kc_add_sequence(orc_kc* kc, const char* seq, int64_t len) {
    for (int64_t i = 0; i + kc->k <= len; i++) orc_kc_add_kmer_str(kc, seq + i, 1);
}
/* populate_kmer_counter_from_reads  fastaToKmerCoverageStats.cpp:230-293: reads shorter than k+1 are skipped
 * (so a read of exactly k bases contributes nothing).  Records are '\n'-terminated, offs has n+1 entries. */
void orc_kc_add_records(orc_kc* kc, const char* recs, const uint64_t* offs, uint64_t nreads) {
    for (uint64_t r = 0; r < nreads; r++) {
        int64_t len = (int64_t)(offs[r + 1] - offs[r]) - 1;
        if (len < kc->k + 1) continue;
        orc_kc_add_sequence(kc, recs + offs[r], len);
    }
}
/* get_kmer_count  KmerCounter.cpp:439-457 (+ find_kmer :367-375) */
static uint32_t kc_get(orc_kc* kc, uint64_t v) {
    if (kc->ds) v = iw_ds_val(v, kc->k);
    uint32_t* p = map_slot(&kc->m, v, 0);
    return p ? *p : 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* fastaToKmerCoverageStats (Inchworm/src/fastaToKmerCoverageStats.cpp)                                    */
/* ------------------------------------------------------------------------------------------------------ */
static int cmp_u32(const void* a, const void* b) {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : x > y;
}
/* median_coverage :337-347 */
static uint32_t median_coverage(uint32_t* sorted_copy, int n) {
    if (n == 0) return 0;
    qsort(sorted_copy, (size_t)n, 4, cmp_u32);
    if (n % 2 == 1) return sorted_copy[n / 2];
    return (sorted_copy[(n - 1) / 2] + sorted_copy[n / 2]) / 2;     /* unsigned int arithmetic */
}
/* sum :371-378, mean :380-387 */
static float mean_cov(const uint32_t* v, size_t n) {
    if (n == 0) return 0;
    long s = 0;
    for (int i = 0; i < (int)n; i++) s += v[i];
    float avg = (float)s / n;
    return avg;
}
/* stDev :389-402 */
static float stdev_cov(const uint32_t* v, size_t n) {
    float avg = mean_cov(v, n);
    int num_vals = (int)n;
    float sum_avg_diffs_sqr = 0;
    for (int i = 0; i < num_vals; i++) {
        float delta = v[i] - avg;
        sum_avg_diffs_sqr += (delta * delta);
    }
    float stdev = sqrtf(sum_avg_diffs_sqr / (num_vals - 1));
    return stdev;
}

/* per read: compute_kmer_coverage :300-335 then the three statistics.  Records are upper-cased by the
 * reference's reader (Fasta_reader.cpp:117); iw_base is case-insensitive, which is equivalent.
 * per_kmer may be NULL; otherwise per_kmer[offs[r] + j] = coverage of window j. */
void orc_cov_stats(orc_kc* kc, const char* recs, const uint64_t* offs, uint64_t nreads, uint32_t* median,
                   float* mean, float* stdev, uint32_t* per_kmer) {
    uint32_t* cov = NULL; uint32_t* tmp = NULL; size_t capn = 0;
    for (uint64_t r = 0; r < nreads; r++) {
        const char* seq = recs + offs[r];
        int64_t len = (int64_t)(offs[r + 1] - offs[r]) - 1;
        size_t n = len >= kc->k ? (size_t)(len - kc->k + 1) : 0;     /* shorter than k -> empty vector */
        if (n > capn) { capn = n * 2; cov = (uint32_t*)realloc(cov, capn * 4); tmp = (uint32_t*)realloc(tmp, capn * 4); }
        for (size_t i = 0; i < n; i++) {
            uint32_t c = 0;
            if (!iw_contains_non_gatc(seq + i, kc->k)) c = kc_get(kc, iw_kmer_to_intval(seq + i, kc->k));
            if (c < 1) c = 1;                                        /* :328-330 */
            cov[i] = c;
            if (per_kmer) per_kmer[offs[r] + i] = c;
        }
        if (n) memcpy(tmp, cov, n * 4);
        median[r] = median_coverage(tmp, (int)n);
        mean[r] = mean_cov(cov, n);
        stdev[r] = stdev_cov(cov, n);
    }
    free(cov); free(tmp);
}

/* ------------------------------------------------------------------------------------------------------ */
/* jellyfish count / dump / histo -- restated semantics (third-party, parity unpinned)    */
/* ------------------------------------------------------------------------------------------------------ */
static int jf_base(unsigned char c) {   /* A<C<G<T, case-insensitive */
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return -1;
    }
}
static uint64_t jf_revcomp(uint64_t v, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) { r = (r << 2) | (3 - (v & 3)); v >>= 2; }
    return r;
}
static void radix_sort_u64(uint64_t* a, uint64_t n) {
    uint64_t* b = (uint64_t*)malloc(n * 8);
    for (int pass = 0; pass < 8; pass++) {
        uint64_t cnt[257] = {0};
        for (uint64_t i = 0; i < n; i++) cnt[((a[i] >> (8 * pass)) & 255) + 1]++;
        for (int i = 0; i < 256; i++) cnt[i + 1] += cnt[i];
        for (uint64_t i = 0; i < n; i++) b[cnt[(a[i] >> (8 * pass)) & 255]++] = a[i];
        uint64_t* t = a; a = b; b = t;
    }
    free(b);   /* 8 passes: data is back in the caller's array */
}
/* J1 + J2: every window of k consecutive ACGTacgt characters of every record is counted (a record of
 * exactly k bases included); canonical folds onto the lexicographically smaller strand.  Returns the
 * distinct k-mers with count >= min_count in ascending packed order; caller frees with orc_free. */
uint64_t orc_jf_count(const char* recs, uint64_t nbytes, int k, int canonical, uint32_t min_count,
                      uint64_t** keys_out, uint32_t** counts_out) {
    uint64_t* all = (uint64_t*)malloc((nbytes + 1) * 8);
    uint64_t n = 0, v = 0;
    const uint64_t mask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1);
    int run = 0;
    for (uint64_t i = 0; i < nbytes; i++) {
        int c = jf_base((unsigned char)recs[i]);
        if (c < 0) { run = 0; v = 0; continue; }       /* non-base (incl. the '\n' terminator) restarts the window */
        v = ((v << 2) | (uint64_t)c) & mask;
        if (++run >= k) {
            uint64_t key = v;
            if (canonical) { uint64_t rc = jf_revcomp(v, k); if (rc < key) key = rc; }
            all[n++] = key;
        }
    }
    radix_sort_u64(all, n);
    uint64_t* keys = (uint64_t*)malloc((n + 1) * 8);
    uint32_t* counts = (uint32_t*)malloc((n + 1) * 4);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n;) {
        uint64_t j = i;
        while (j < n && all[j] == all[i]) j++;
        if (j - i >= min_count) { keys[m] = all[i]; counts[m] = (uint32_t)(j - i); m++; }
        i = j;
    }
    free(all);
    *keys_out = keys; *counts_out = counts;
    return m;
}
/* J3: bins[c] for c in 1..10000, bins[10001] = everything larger */
void orc_jf_histo(const uint32_t* counts, uint64_t n, uint64_t* bins /* 10002 */) {
    memset(bins, 0, 10002 * 8);
    for (uint64_t i = 0; i < n; i++) bins[counts[i] > 10000 ? 10001 : counts[i]]++;
}
void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------------------ */
/* ReadsToTranscripts (Chrysalis/analysis/ReadsToTranscripts.cc, NonRedKmerTable.cc, sequenceUtil.cc)      */
/* ------------------------------------------------------------------------------------------------------ */
typedef struct { orc_map m; int k; } orc_rt;

orc_rt* orc_rt_new(int k) {
    orc_rt* t = (orc_rt*)malloc(sizeof *t);
    map_init(&t->m, 1u << 16);
    t->k = k;
    return t;
}
void orc_rt_free(orc_rt* t) { if (t) { map_free(&t->m); free(t); } }
uint64_t orc_rt_size(orc_rt* t) { return t->m.n; }

/* Regular() NonRedKmerTable.cc:3-8 applied to the upper-cased bundle (vecDNAVector::Read allUpper) */
static int rt_pack(const char* s, int k, uint64_t* out) {
    uint64_t v = 0;
    for (int i = 0; i < k; i++) {
        int c;
        switch (s[i]) { case 'A': c = 0; break; case 'C': c = 1; break; case 'G': c = 2; break; case 'T': c = 3; break; default: return 0; }
        v = (v << 2) | (uint64_t)c;
    }
    *out = v;
    return 1;
}
/* SetUp(dna,true) + SetAllCounts(-1) + the SetCount loop (ReadsToTranscripts.cc:144-169) in bundle order:
 * the last (= highest-index) bundle containing a k-mer owns it.  Stored value = index + 1. */
void orc_rt_label(orc_rt* t, const char* recs, const uint64_t* offs, uint64_t nbundles, uint32_t first_index) {
    char* up = NULL; size_t upcap = 0;
    for (uint64_t b = 0; b < nbundles; b++) {
        const char* seq = recs + offs[b];
        int64_t len = (int64_t)(offs[b + 1] - offs[b]) - 1;
        if ((size_t)len + 1 > upcap) { upcap = (size_t)len * 2 + 16; up = (char*)realloc(up, upcap); }
        for (int64_t i = 0; i < len; i++) up[i] = (char)((seq[i] >= 'a' && seq[i] <= 'z') ? seq[i] - 32 : seq[i]);
        for (int64_t j = 0; j + t->k <= len; j++) {
            uint64_t key;
            if (!rt_pack(up + j, t->k, &key)) continue;
            *map_slot(&t->m, key, 1) = first_index + (uint32_t)b + 1;
        }
    }
    free(up);
}
/* compute_entropy(string&)  Chrysalis/analysis/sequenceUtil.cc:326-355: counts of 'G','A','T','C' (any other
 * character only adds to the length), fp32, `log` on a float is the float overload under <math.h> + g++. */
static float rt_entropy(const char* s, int k) {
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < k; i++) {
        switch (s[i]) { case 'G': cnt[0]++; break; case 'A': cnt[1]++; break; case 'T': cnt[2]++; break; case 'C': cnt[3]++; break; default: break; }
    }
    float entropy = 0;
    for (int i = 0; i < 4; i++) {
        float prob = (float)cnt[i] / k;
        if (prob > 0) {
            float val = prob * logf(1 / prob) / logf(2.0f);
            entropy += val;
        }
    }
    return entropy;
}
/* DNACodec reverse complement (DNAVector.cc:14-59, 261-281): IUPAC pairs, anything unknown -> '\0' */
static char rt_rc_char(char c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'K': return 'M'; case 'M': return 'K'; case 'R': return 'Y'; case 'Y': return 'R';
        case 'S': return 'S'; case 'W': return 'W'; case 'B': return 'V'; case 'V': return 'B';
        case 'H': return 'D'; case 'D': return 'H'; case '-': return '-'; case 'N': return 'N'; case 'X': return 'X';
        default: return 0;
    }
}
/* exported for tests: the reference's entropy of one window */
float orc_entropy(const char* s, int k) { return rt_entropy(s, k); }

static int cmp_i32(const void* a, const void* b) {
    int x = *(const int*)a, y = *(const int*)b;
    return x < y ? -1 : x > y;
}
/* the per-read loop, ReadsToTranscripts.cc:216-274 */
void orc_rt_assign(orc_rt* t, const char* recs, const uint64_t* offs, uint64_t nreads, int strand,
                   float min_kmer_entropy, int32_t* best_out, int32_t* pct_out, int32_t* score_out) {
    const int k = t->k;
    char* d = NULL; char* dd = NULL; int* comp = NULL; size_t capn = 0;
    for (uint64_t r = 0; r < nreads; r++) {
        const char* seq = recs + offs[r];
        int64_t len = (int64_t)(offs[r + 1] - offs[r]) - 1;
        if ((size_t)len + 1 > capn) {
            capn = (size_t)len * 2 + 16;
            d = (char*)realloc(d, capn); dd = (char*)realloc(dd, capn); comp = (int*)realloc(comp, 2 * capn * sizeof(int));
        }
        for (int64_t i = 0; i < len; i++) d[i] = (char)((seq[i] >= 'a' && seq[i] <= 'z') ? seq[i] - 32 : seq[i]);
        int ncomp = 0;
        int num_kmer_pos = (int)len - k + 1;
        for (int j = 0; j <= (int)len - k; j++) {
            float entropy = rt_entropy(d + j, k);
            if (entropy < min_kmer_entropy) continue;
            uint64_t key;
            if (!rt_pack(d + j, k, &key)) continue;           /* a k-mer with a non-ACGT char is not in the table */
            uint32_t* p = map_slot(&t->m, key, 0);
            if (p) comp[ncomp++] = (int)*p - 1;
        }
        if (!strand) {
            for (int64_t i = 0; i < len; i++) dd[i] = rt_rc_char(d[len - 1 - i]);
            for (int j = 0; j <= (int)len - k; j++) {
                float entropy = rt_entropy(dd + j, k);
                if (entropy < min_kmer_entropy) continue;
                uint64_t key;
                if (!rt_pack(dd + j, k, &key)) continue;
                uint32_t* p = map_slot(&t->m, key, 0);
                if (p) comp[ncomp++] = (int)*p - 1;
            }
        }
        qsort(comp, (size_t)ncomp, sizeof(int), cmp_i32);
        int best = -1, max = 0, run = 0;
        for (int j = 1; j < ncomp; j++) {
            if (comp[j] != comp[j - 1] || j + 1 == ncomp) {
                if (run > max) { max = run; best = comp[j - 1]; }
                run = 0;
            } else {
                run++;
            }
        }
        int pct_read_mapped = num_kmer_pos > 0 ? (int)((float)max / num_kmer_pos * 100 + 0.5) : 0;
        best_out[r] = best;
        pct_out[r] = pct_read_mapped;
        if (score_out) score_out[r] = max;
    }
    free(d); free(dd); free(comp);
}

/* =========================================================================================================
 * GraphFromFasta weldmer counting
 *   NonRedKmerTable::SetUp(templ)  Chrysalis/analysis/NonRedKmerTable.cc:12-96   the candidates, sorted and made unique
 *   NonRedKmerTable::AddData(DNAStringStreamFast&)  :162-200   every read upper-cased; every window of m_k characters is
 *       looked up by binary search over the sorted STRINGS (exact match: forward strand, no canonicalisation) and bumps
 *       the counter of the candidate it equals
 *   NonRedKmerTable::GetCount  NonRedKmerTable.h:50-55          the counter, 0 for a string that is not a candidate
 * counts[i] = counter of candidate i after all records (duplicate candidates share one counter).
 * ========================================================================================================= */
static int g_weld_kk;
static const char* g_weld_base;
static int weld_cmp_idx(const void* a, const void* b) {
    return memcmp(g_weld_base + *(const uint64_t*)a * (uint64_t)g_weld_kk, g_weld_base + *(const uint64_t*)b * (uint64_t)g_weld_kk,
                  (size_t)g_weld_kk);
}
void orc_weld_count(const char* weldmers, uint64_t n, int kk, const char* recs, const uint64_t* offs, uint64_t nreads,
                    int32_t* counts) {
    uint64_t* order = (uint64_t*)malloc((n ? n : 1) * sizeof *order);
    int32_t* cnt = (int32_t*)calloc(n ? n : 1, sizeof *cnt);             /* counter of the candidate at sorted position i */
    for (uint64_t i = 0; i < n; i++) order[i] = i;
    g_weld_kk = kk; g_weld_base = weldmers;
    qsort(order, n, sizeof *order, weld_cmp_idx);                          /* UniqueSort: equal strings end up adjacent */
    char* win = (char*)malloc((size_t)kk + 1);
    for (uint64_t r = 0; r < nreads; r++) {
        const char* s = recs + offs[r];
        const int64_t len = (int64_t)(offs[r + 1] - offs[r]) - 1;
        for (int64_t j = 0; j <= len - kk; j++) {
            for (int x = 0; x < kk; x++) win[x] = (char)toupper((unsigned char)s[j + x]);
            /* BinSearch: first sorted position whose string is >= the window */
            uint64_t lo = 0, hi = n;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) / 2;
                if (memcmp(weldmers + order[mid] * (uint64_t)kk, win, (size_t)kk) < 0) lo = mid + 1; else hi = mid;
            }
            if (lo < n && memcmp(weldmers + order[lo] * (uint64_t)kk, win, (size_t)kk) == 0) cnt[lo]++;
        }
    }
    /* every input candidate reports the counter of the first sorted position holding its string */
    for (uint64_t i = 0; i < n; i++) {
        uint64_t lo = 0, hi = n;
        const char* w = weldmers + i * (uint64_t)kk;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) / 2;
            if (memcmp(weldmers + order[mid] * (uint64_t)kk, w, (size_t)kk) < 0) lo = mid + 1; else hi = mid;
        }
        counts[i] = cnt[lo];
    }
    free(win); free(cnt); free(order);
}

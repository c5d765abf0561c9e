/*
 * indexlr_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle / CPU baseline).
 *
 * A CPU restatement, in plain C, of what btllib's `indexlr --long --pos --strand [--len]`
 * computes (ntHash canonical rolling hash, second hash, windowed minimizers).
 *
 * The reference (bcgsc/ntLink v1.3.11) does NOT vendor this algorithm: it shells out to the
 * external, un-vendored dependency btllib (<= 1.6.2; README.md:133, requirements.txt:3), at
 *   ntLink:198-199   indexlr --long --pos --strand -k K -w W -t T target > target.kK.wW.tsv
 *   ntLink:221-225   gzip -cd reads | indexlr --long --pos --strand --len -k K -w W -t T - | ntlink_pair.py ...
 * so this file restates btllib's published algorithm (ntHash: Mohamadi et al. 2016 / Kazemi et al. 2022;
 * btllib Indexlr::minimize / calc_minimizer) following SURVEY.md section 8a S0-S4, and parity is
 * PINNED on the reference's own golden vectors
 *   tests/expected_outputs/scaffolds_{1,2,3,4}.fa.k*.w*.tsv   (byte-for-byte, see tests/test_oracle_sketch.py)
 * and transitively on pairs.tsv / scaffold.dot / verbose_mapping.tsv / the PAF lines of
 * tests/ntlink_pytest.py:189-194.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may call
 * this code. The product path (ntlink_b200/) never does.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp; also as a shared library for ctypes).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <pthread.h>

/* minimal parallel-for over [0,n) with dynamic scheduling (pthreads; no OpenMP dependency) */
typedef void (*pf_body_t)(long i, void *ctx);
typedef struct { pf_body_t body; void *ctx; long n; long next; long chunk; } pf_job_t;
static void *pf_worker(void *arg) {
    pf_job_t *j = (pf_job_t *)arg;
    for (;;) {
        long b = __atomic_fetch_add(&j->next, j->chunk, __ATOMIC_RELAXED);
        if (b >= j->n) break;
        long e = b + j->chunk < j->n ? b + j->chunk : j->n;
        for (long i = b; i < e; i++) j->body(i, j->ctx);
    }
    return NULL;
}
static void parallel_for(long n, int threads, long chunk, pf_body_t body, void *ctx) {
    pf_job_t job = { body, ctx, n, 0, chunk > 0 ? chunk : 1 };
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads == 1 || n <= 1) { pf_worker(&job); return; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < threads - 1; t++) if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) started++;
    pf_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------ S1: ntHash ---- */
/* seeds (btllib nthash_consts) */
#define SEED_A 0x3c8bfbb395c60474ULL
#define SEED_C 0x3193c18562a02b4cULL
#define SEED_G 0x20323ed082572324ULL
#define SEED_T 0x295549f54be24456ULL

/* base -> code 0..3 (A,C,G,T; case-insensitive), 4 = invalid (N, IUPAC, U, anything else) */
static unsigned char CODE[256];
static const uint64_t SEED[5] = { SEED_A, SEED_C, SEED_G, SEED_T, 0 };
static const uint64_t SEED_RC[5] = { SEED_T, SEED_G, SEED_C, SEED_A, 0 }; /* seed of the complement */

/* thread-safe: the first sketches of a process may run on several worker threads at once, and a second thread
 * re-initialising the table (memset to "invalid") under a running sketch would make valid bases look like N */
static void init_tables_once(void) {
    memset(CODE, 4, sizeof CODE);
    CODE['A'] = CODE['a'] = 0; CODE['C'] = CODE['c'] = 1;
    CODE['G'] = CODE['g'] = 2; CODE['T'] = CODE['t'] = 3;
}
static void init_tables(void) {
    static pthread_once_t once = PTHREAD_ONCE_INIT;
    pthread_once(&once, init_tables_once);
}

/* split rotate left by 1: low 33 bits and high 31 bits rotate separately (bit32->bit0, bit63->bit33) */
static inline uint64_t srol1(uint64_t x) {
    uint64_t m = ((x & 0x8000000000000000ULL) >> 30) | ((x & 0x100000000ULL) >> 32);
    return ((x << 1) & 0xFFFFFFFDFFFFFFFEULL) | m;
}
/* split rotate right by 1 (inverse of srol1: bit0->bit32, bit33->bit63) */
static inline uint64_t sror1(uint64_t x) {
    uint64_t m = ((x & 0x200000000ULL) << 30) | ((x & 1ULL) << 32);
    return ((x >> 1) & 0xFFFFFFFEFFFFFFFFULL) | m;
}
/* split rotate left by d (any d): 33-bit part by d%33, 31-bit part by d%31 */
static inline uint64_t sroln(uint64_t x, unsigned d) {
    const uint64_t M33 = (1ULL << 33) - 1, M31 = (1ULL << 31) - 1;
    uint64_t lo = x & M33, hi = x >> 33;
    unsigned a = d % 33, b = d % 31;
    if (a) lo = ((lo << a) | (lo >> (33 - a))) & M33;
    if (b) hi = ((hi << b) | (hi >> (31 - b))) & M31;
    return lo | (hi << 33);
}

/* S2: second hash printed by indexlr (ntHash extra hash i=1) */
static inline uint64_t second_hash(uint64_t h0, unsigned k) {
    uint64_t t = h0 * (1ULL ^ ((uint64_t)k * 0x90b45d39fb6da1faULL));
    return t ^ (t >> 27);
}

/* ------------------------------------------------------------------ S3: minimizers ---- */
typedef struct { uint64_t h0, h1; uint32_t pos; uint8_t fwd; } hk_t;

typedef struct { uint64_t *hash; uint32_t *pos; uint8_t *strand; size_t n, cap; } mxlist_t;

static void mx_push(mxlist_t *m, uint64_t h, uint32_t p, uint8_t s) {
    if (m->n == m->cap) {
        m->cap = m->cap ? m->cap * 2 : 64;
        m->hash = (uint64_t *)realloc(m->hash, m->cap * sizeof(uint64_t));
        m->pos = (uint32_t *)realloc(m->pos, m->cap * sizeof(uint32_t));
        m->strand = (uint8_t *)realloc(m->strand, m->cap);
    }
    m->hash[m->n] = h; m->pos[m->n] = p; m->strand[m->n] = s; m->n++;
}

/*
 * Sketch one sequence. Follows btllib Indexlr::minimize + calc_minimizer:
 *   - k-mers containing an invalid base are skipped; valid k-mers get consecutive indices idx=0,1,2...
 *   - a window is w consecutive VALID k-mers in idx space (it spans N gaps);
 *   - window minimizer = rightmost argmin of the canonical hash h0 (<= on rescan and on slide);
 *   - emitted when its position is greater than the last emitted position (and h0 != UINT64_MAX);
 *   - sequences with k > L or w > L-k+1 yield nothing.
 * The hash printed is the second hash h1; strand '+' iff forward hash <= reverse hash.
 */
static void sketch_sequence(const char *seq, size_t L, unsigned k, unsigned w, mxlist_t *out) {
    init_tables();
    out->n = 0;
    if (k == 0 || w == 0 || (size_t)k > L || (size_t)w > L - k + 1) return;
    hk_t *ring = (hk_t *)malloc((size_t)w * sizeof(hk_t));
    const uint64_t *srolk_cache = NULL; (void)srolk_cache;
    uint64_t out_f[5], out_r[5];           /* srol^k(seed[c]) for leaving base (fwd), and rc seed entering */
    for (int c = 0; c < 5; c++) { out_f[c] = sroln(SEED[c], k); out_r[c] = sroln(SEED_RC[c], k); }

    size_t idx = 0;                        /* number of valid k-mers seen so far */
    long min_pos_prev = -1;
    const hk_t *cur = NULL;                /* current window minimum (points into ring) */
    uint64_t fh = 0, rh = 0;
    size_t i = 0;                          /* k-mer start */
    int have = 0;                          /* is (fh,rh) the hash of k-mer i? */
    const size_t nk = L - k + 1;
    while (i < nk) {
        if (!have) {
            /* (re)initialise at i: find invalid base in [i, i+k) */
            size_t bad = (size_t)-1;
            for (size_t j = i + k; j-- > i;) if (CODE[(unsigned char)seq[j]] == 4) { bad = j; break; }
            if (bad != (size_t)-1) { i = bad + 1; continue; }     /* jump past the last invalid base */
            fh = 0; rh = 0;
            for (unsigned j = 0; j < k; j++) {
                unsigned c = CODE[(unsigned char)seq[i + j]];
                fh = srol1(fh) ^ SEED[c];
                rh ^= sroln(SEED_RC[c], j);
            }
            have = 1;
        }
        /* record valid k-mer i */
        {
            hk_t *slot = &ring[idx % w];
            /* the ring slot being overwritten left the window already unless it is `cur`, which is
               checked below through its position */
            hk_t nk_;
            nk_.h0 = fh + rh; nk_.h1 = 0; nk_.pos = (uint32_t)i; nk_.fwd = (fh <= rh);
            /* btllib keeps w+? slots; we keep exactly w and handle `cur` leaving explicitly */
            int cur_overwritten = (cur == slot);
            hk_t saved; if (cur_overwritten) saved = *cur;
            *slot = nk_;
            if (idx + 1 >= w) {
                size_t left_idx = idx + 1 - w;
                uint32_t left_pos = ring[left_idx % w].pos;
                if (cur == NULL || cur_overwritten || cur->pos < left_pos) {
                    (void)saved;
                    cur = &ring[left_idx % w];
                    for (size_t q = left_idx; q <= idx; q++) {
                        const hk_t *c = &ring[q % w];
                        if (c->h0 <= cur->h0) cur = c;
                    }
                } else if (slot->h0 <= cur->h0) {
                    cur = slot;
                }
                if ((long)cur->pos > min_pos_prev && cur->h0 != UINT64_MAX) {
                    min_pos_prev = (long)cur->pos;
                    mx_push(out, second_hash(cur->h0, k), cur->pos, cur->fwd);
                }
            }
            idx++;
        }
        /* roll to i+1 */
        if (i + 1 < nk) {
            unsigned cin = CODE[(unsigned char)seq[i + k]];
            if (cin == 4) { have = 0; i = i + k + 1; continue; }  /* skip every k-mer containing it */
            unsigned cout = CODE[(unsigned char)seq[i]];
            fh = srol1(fh) ^ SEED[cin] ^ out_f[cout];
            rh = sror1(rh ^ out_r[cin] ^ SEED_RC[cout]);
        }
        i++;
    }
    free(ring);
}

/* ---------------------------------------------------------------- library entry points ---- */
/* Sketch one sequence into caller-provided arrays (cap entries); returns the number of minimizers
 * (which may exceed cap; only the first cap are written). */
size_t ntl_oracle_sketch(const char *seq, size_t L, unsigned k, unsigned w,
                         uint64_t *hash, uint32_t *pos, uint8_t *strand, size_t cap) {
    mxlist_t m = {0};
    sketch_sequence(seq, L, k, w, &m);
    size_t n = m.n < cap ? m.n : cap;
    if (n) { memcpy(hash, m.hash, n * 8); memcpy(pos, m.pos, n * 4); memcpy(strand, m.strand, n); }
    size_t total = m.n;
    free(m.hash); free(m.pos); free(m.strand);
    return total;
}

typedef struct { const char *seq; const uint64_t *offsets; unsigned k, w; mxlist_t *lists; } batch_ctx_t;
static void batch_body(long i, void *ctx) {
    batch_ctx_t *b = (batch_ctx_t *)ctx;
    sketch_sequence(b->seq + b->offsets[i], (size_t)(b->offsets[i + 1] - b->offsets[i]), b->k, b->w, &b->lists[i]);
}

/* Batch form: nseq sequences concatenated in `seq` with offsets[nseq+1]; results appended in order.
 * out_off[nseq+1] receives per-sequence offsets. Multi-threaded over sequences (OpenMP).
 * Returns total minimizers, or (size_t)-1 if cap is too small. */
size_t ntl_oracle_sketch_batch(const char *seq, const uint64_t *offsets, uint32_t nseq, unsigned k, unsigned w,
                               int threads, uint64_t *hash, uint32_t *pos, uint8_t *strand, size_t cap,
                               uint64_t *out_off) {
    mxlist_t *lists = (mxlist_t *)calloc(nseq ? nseq : 1, sizeof(mxlist_t));
    batch_ctx_t bc = { seq, offsets, k, w, lists };
    parallel_for((long)nseq, threads, 16, batch_body, &bc);
    size_t total = 0;
    for (uint32_t i = 0; i < nseq; i++) { out_off[i] = total; total += lists[i].n; }
    out_off[nseq] = total;
    if (total <= cap) {
        for (uint32_t i = 0; i < nseq; i++) {
            size_t o = out_off[i], n = lists[i].n;
            if (n) { memcpy(hash + o, lists[i].hash, n * 8); memcpy(pos + o, lists[i].pos, n * 4);
                     memcpy(strand + o, lists[i].strand, n); }
        }
    }
    for (uint32_t i = 0; i < nseq; i++) { free(lists[i].hash); free(lists[i].pos); free(lists[i].strand); }
    free(lists);
    return total <= cap ? total : (size_t)-1;
}

/* raw hash known-answer access for unit tests: hashes of k-mer at seq[0..k) */
void ntl_oracle_kmer_hashes(const char *kmer, unsigned k, uint64_t *fh_out, uint64_t *rh_out,
                            uint64_t *h0_out, uint64_t *h1_out) {
    init_tables();
    uint64_t fh = 0, rh = 0;
    for (unsigned j = 0; j < k; j++) {
        unsigned c = CODE[(unsigned char)kmer[j]];
        fh = srol1(fh) ^ SEED[c];
        rh ^= sroln(SEED_RC[c], j);
    }
    *fh_out = fh; *rh_out = rh; *h0_out = fh + rh; *h1_out = second_hash(fh + rh, k);
}

uint64_t ntl_oracle_srol(uint64_t x, unsigned d) { return sroln(x, d); }

#ifndef NTL_ORACLE_NO_MAIN
/* ------------------------------------------------------------------ S0: reader + S4: TSV ---- */
typedef struct { char *name; char *seq; size_t len; } rec_t;

typedef struct { gzFile f; char *buf; size_t cap; int eof; char *pend; } reader_t;

static char *rd_line(reader_t *r, size_t *len_out) {
    /* returns a line without the trailing newline (and without '\r'); NULL at EOF */
    size_t n = 0;
    for (;;) {
        if (r->cap - n < 2) { r->cap = r->cap ? r->cap * 2 : (1 << 16); r->buf = (char *)realloc(r->buf, r->cap); }
        if (!gzgets(r->f, r->buf + n, (int)(r->cap - n > 0x7fffffff ? 0x7fffffff : r->cap - n))) {
            if (n == 0) return NULL;
            break;
        }
        n += strlen(r->buf + n);
        if (n && r->buf[n - 1] == '\n') { n--; break; }
    }
    if (n && r->buf[n - 1] == '\r') n--;
    r->buf[n] = 0; *len_out = n;
    return r->buf;
}

/* readfq-style parser: FASTA or FASTQ, multi-line, id = header up to first whitespace */
static int next_record(reader_t *r, rec_t *rec, char **hdr_carry) {
    size_t n; char *ln;
    char *hdr = *hdr_carry; *hdr_carry = NULL;
    while (!hdr) {
        ln = rd_line(r, &n);
        if (!ln) return 0;
        if (ln[0] == '>' || ln[0] == '@') hdr = strdup(ln);
    }
    size_t e = 1; while (hdr[e] && hdr[e] != ' ' && hdr[e] != '\t') e++;
    rec->name = (char *)malloc(e); memcpy(rec->name, hdr + 1, e - 1); rec->name[e - 1] = 0;
    free(hdr);
    size_t cap = 1 << 12, len = 0; char *s = (char *)malloc(cap);
    int saw_plus = 0;
    while ((ln = rd_line(r, &n))) {
        /* readfq semantics (bin/read_fasta.py:27-30): any line starting with '@', '+' or '>' ends the sequence */
        if (ln[0] == '>' || ln[0] == '@') { *hdr_carry = strdup(ln); break; }
        if (ln[0] == '+') { saw_plus = 1; break; }
        if (len + n + 1 > cap) { while (len + n + 1 > cap) cap *= 2; s = (char *)realloc(s, cap); }
        memcpy(s + len, ln, n); len += n;
    }
    s[len] = 0;
    if (saw_plus) {                         /* consume quality: as many characters as the sequence */
        size_t q = 0;
        while (q < len && (ln = rd_line(r, &n))) q += n;
    }
    rec->seq = s; rec->len = len;
    return 1;
}

static size_t fmt_u64(char *p, uint64_t v) {
    char t[24]; int n = 0;
    do { t[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = 0; i < n; i++) p[i] = t[n - 1 - i];
    return (size_t)n;
}

typedef struct { rec_t *recs; char **outs; size_t *outn; unsigned k, w; int f_pos, f_strand, f_len; } cli_ctx_t;
static void cli_body(long i, void *ctx) {
    cli_ctx_t *c = (cli_ctx_t *)ctx;
    rec_t *recs = c->recs; char **outs = c->outs; size_t *outn = c->outn;
    unsigned k = c->k, w = c->w; int f_pos = c->f_pos, f_strand = c->f_strand, f_len = c->f_len;

            mxlist_t m = {0};
            sketch_sequence(recs[i].seq, recs[i].len, k, w, &m);
            size_t cap = strlen(recs[i].name) + 32 + m.n * 34 + 2;
            char *o = (char *)malloc(cap), *p = o;
            size_t nl = strlen(recs[i].name); memcpy(p, recs[i].name, nl); p += nl;
            if (f_len) { *p++ = '\t'; p += fmt_u64(p, recs[i].len); }
            *p++ = '\t';
            for (size_t j = 0; j < m.n; j++) {
                if (j) *p++ = ' ';
                p += fmt_u64(p, m.hash[j]);
                if (f_pos) { *p++ = ':'; p += fmt_u64(p, m.pos[j]); }
                if (f_strand) { *p++ = ':'; *p++ = m.strand[j] ? '+' : '-'; }
            }
            *p++ = '\n';
            outs[i] = o; outn[i] = (size_t)(p - o);
            free(m.hash); free(m.pos); free(m.strand);
}

int main(int argc, char **argv) {
    unsigned k = 0, w = 0; int threads = 1, f_pos = 0, f_strand = 0, f_len = 0;
    const char *path = NULL;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--long")) continue;
        else if (!strcmp(argv[i], "--pos")) f_pos = 1;
        else if (!strcmp(argv[i], "--strand")) f_strand = 1;
        else if (!strcmp(argv[i], "--len")) f_len = 1;
        else if (!strcmp(argv[i], "-k") && i + 1 < argc) k = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-w") && i + 1 < argc) w = (unsigned)atoi(argv[++i]);
        else if (!strcmp(argv[i], "-t") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (!strncmp(argv[i], "-k", 2) && argv[i][2]) k = (unsigned)atoi(argv[i] + 2);
        else if (!strncmp(argv[i], "-w", 2) && argv[i][2]) w = (unsigned)atoi(argv[i] + 2);
        else if (!strncmp(argv[i], "-t", 2) && argv[i][2]) threads = atoi(argv[i] + 2);
        else if (argv[i][0] != '-' || !strcmp(argv[i], "-")) path = argv[i];
        else { fprintf(stderr, "indexlr_oracle: unknown option %s\n", argv[i]); return 2; }
    }
    if (!k || !w || !path) {
        fprintf(stderr, "usage: indexlr_oracle --long [--pos] [--strand] [--len] -k K -w W [-t T] FILE|-\n");
        return 2;
    }
    reader_t rd; memset(&rd, 0, sizeof rd);
    rd.f = strcmp(path, "-") ? gzopen(path, "rb") : gzdopen(0, "rb");
    if (!rd.f) { fprintf(stderr, "indexlr_oracle: cannot open %s\n", path); return 1; }
    gzbuffer(rd.f, 1 << 20);
    enum { BATCH = 2048 };
    rec_t *recs = (rec_t *)calloc(BATCH, sizeof(rec_t));
    char **outs = (char **)calloc(BATCH, sizeof(char *));
    size_t *outn = (size_t *)calloc(BATCH, sizeof(size_t));
    char *carry = NULL;
    for (;;) {
        int nb = 0; size_t bases = 0;
        while (nb < BATCH && bases < (256u << 20) && next_record(&rd, &recs[nb], &carry)) { bases += recs[nb].len; nb++; }
        if (!nb) break;
        cli_ctx_t cc = { recs, outs, outn, k, w, f_pos, f_strand, f_len };
        parallel_for(nb, threads, 1, cli_body, &cc);
        for (int i = 0; i < nb; i++) {
            fwrite(outs[i], 1, outn[i], stdout);
            free(outs[i]); free(recs[i].name); free(recs[i].seq);
        }
    }
    gzclose(rd.f);
    return 0;
}
#endif

/* ORACLE (test infrastructure, NOT product code) -- plain-C restatement of the vafator
 * per-variant pileup path, evaluated directly per (variant, read) pair.
 *
 * PARITY STATUS: the vafator layer (vafator/pileups.py:162-362, pileup_utils.py:64-119) is
 * pinned through tests/golden/pileup_*.json, which were produced by the UNMODIFIED reference
 * functions (oracle/gen_golden.py).  The pysam-0.21.0 / htslib-1.17 layer underneath (which
 * reads reach a column, qpos / indel / is_del, the mate-overlap quality tweak) is restated
 * from the published algorithm: **parity unpinned** for that layer --
 * the reference's known answers need a BAM that is absent (.MISSING_LARGE_BLOBS).
 *
 * It is deliberately a different formulation from oracle/htslib_pileup.py (a streaming
 * push/next state machine): here every pair is resolved from scratch, lists are sorted and
 * ranked by binary search.  Agreement of the two on the golden scenarios is checked in
 * tests/test_oracle_pins.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/liboracle.so oracle/pileup_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { OP_M = 0, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X };
enum { F_PAIRED = 1, F_PROPER = 2, F_UNMAP = 4, F_MUNMAP = 8 };

typedef struct {
    int64_t n;
    const int32_t *tid, *pos;
    const uint16_t *flag;
    const uint8_t *mapq;
    const int32_t *l_qseq, *n_cigar;
    const int64_t *cigar_off, *seq_off, *qual_off;
    const uint32_t *cigar;
    const uint8_t *seq, *qual; /* seq: BAM 4-bit packed, high nibble first, each read byte aligned */
    const int32_t *mtid, *mpos, *isize;
    const uint64_t *qname_id;
    const uint32_t *qname_hash; /* khash X31 of the read name */
} orc_reads;

typedef struct {
    int32_t min_mapq, min_baseq;
    uint32_t flag_filter; /* pysam default 0x704 */
    int32_t ignore_orphans, ignore_overlaps;
} orc_params;

typedef struct {
    int64_t n;
    const int32_t *tid, *pos1;            /* 1-based POS */
    const int32_t *stop;                  /* pysam fetch stop of the chromosome group = its last POS
                                             (vafator/pileups.py:137-138); reads with pos >= stop are never read */
    const int64_t *ref_off, *alt_off, *alt0_off; /* offsets into text, n+1 entries each */
    const char *ref_txt, *alt_txt, *alt0_txt;
} orc_variants;

typedef struct {
    int32_t supported; /* 0: MNV/complex -> vafator returns None */
    int32_t dp, n_amb, ac_ref, ac_alt;
    int32_t n_val[3][2]; /* [mq,bq,pos][ref,alt] number of values in the distribution */
    int32_t med2[3][2];  /* twice the median, -1 when empty */
    uint64_t s2[3];      /* twice the sum of the ALT mid-ranks in the pooled ranking */
    int32_t n_col;       /* reads in the htslib column before the base-quality filter (0: no column) */
    int32_t pad;
} orc_counts;

static inline int is_ref_op(int op) { return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X; }
static inline int is_aln_op(int op) { return op == OP_M || op == OP_EQ || op == OP_X; }

static inline int32_t ref_span(const orc_reads *r, int64_t i) {
    int32_t s = 0;
    const uint32_t *c = r->cigar + r->cigar_off[i];
    for (int k = 0; k < r->n_cigar[i]; ++k)
        if (is_ref_op(c[k] & 15)) s += (int32_t)(c[k] >> 4);
    return s;
}
static inline int base_at(const orc_reads *r, int64_t i, int32_t q) {
    uint8_t b = r->seq[r->seq_off[i] + (q >> 1)];
    return (q & 1) ? (b & 15) : (b >> 4);
}

/* ---- htslib khash.h Wang hash; SURVEY.md A.2 ---- */
static inline uint32_t wang(uint32_t k) {
    k += ~(k << 15); k ^= (k >> 10); k += (k << 3); k ^= (k >> 6); k += ~(k << 11); k ^= (k >> 16);
    return k;
}

/* ---- which reads enter the pileup: pysam __advance_samtools + bam_plp_push ---- */
static inline int read_passes(const orc_reads *r, int64_t i, const orc_params *p) {
    uint16_t f = r->flag[i];
    if (r->tid[i] < 0 || (f & F_UNMAP)) return 0;
    if (f & p->flag_filter) return 0;
    if ((int)r->mapq[i] < p->min_mapq) return 0;
    if (p->ignore_orphans && (f & F_PAIRED) && !(f & F_PROPER)) return 0;
    return 1;
}

/* ---- mate linking: emulation of htslib overlap_push / overlap_remove (sam.c) ----
 * mate[i] = partner index or -1; first[i] = 1 when read i is the earlier ("a") record. */
typedef struct { uint64_t id; uint32_t h; int64_t idx; int used; } slot_t;

int orc_link_mates(const orc_reads *r, const orc_params *p, int64_t *mate) {
    int64_t n = r->n;
    for (int64_t i = 0; i < n; ++i) mate[i] = -1;
    if (!p->ignore_overlaps || n == 0) return 0;
    uint64_t cap = 16;
    while (cap < (uint64_t)n * 2) cap <<= 1;
    slot_t *tab = (slot_t *)calloc(cap, sizeof(slot_t));
    if (!tab) return -1;
    int32_t cur_tid = -1, prev_beg = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (!read_passes(r, i, p)) continue;
        if (r->tid[i] != cur_tid) { /* new contig: every node of the old one is gone */
            memset(tab, 0, cap * sizeof(slot_t));
            cur_tid = r->tid[i];
            prev_beg = r->pos[i];
        }
        int32_t beg = r->pos[i], end = beg + ref_span(r, i);
        int pushed_to_list = 1; /* bam_plp_push appends iff end > iter->pos (== prev_beg) */
        if (!(end > prev_beg)) pushed_to_list = 0;
        int32_t pb = prev_beg;
        prev_beg = beg;
        if (!pushed_to_list) continue;
        uint16_t f = r->flag[i];
        if ((f & F_MUNMAP) || !(f & F_PROPER)) continue;
        if ((r->mtid[i] >= 0 && r->tid[i] != r->mtid[i]) ||
            (llabs((long long)r->isize[i]) >= 2LL * r->l_qseq[i] && r->mpos[i] >= end))
            continue;
        uint64_t key = r->qname_id[i] * 0x9E3779B97F4A7C15ull ^ r->qname_hash[i];
        uint64_t s = key & (cap - 1);
        int64_t found = -1;
        uint64_t fs = 0;
        while (tab[s].used) {
            if (tab[s].used == 1 && tab[s].id == r->qname_id[i] && tab[s].h == r->qname_hash[i]) { found = tab[s].idx; fs = s; break; }
            s = (s + 1) & (cap - 1);
        }
        if (found >= 0) {
            /* entry whose node left the buffer before this push: a column >= its end was visited,
             * i.e. end < beg of the previously pushed read */
            int32_t aend = r->pos[found] + ref_span(r, found);
            if (aend < pb) { tab[fs].used = 2; found = -1; }
        }
        if (found < 0) {
            if (r->mpos[i] >= r->pos[i] || ((f & F_PAIRED) && r->mpos[i] == -1)) {
                s = key & (cap - 1);
                while (tab[s].used == 1) s = (s + 1) & (cap - 1);
                tab[s].used = 1; tab[s].id = r->qname_id[i]; tab[s].h = r->qname_hash[i]; tab[s].idx = i;
            }
        } else {
            mate[found] = i;
            mate[i] = found;
            tab[fs].used = 2; /* tombstone: kh_del */
        }
    }
    free(tab);
    return 0;
}

/* ---- resolve one read at one reference position (htslib resolve_cigar2 semantics) ---- */
typedef struct { int covered, qpos, indel, is_del, is_refskip; } plp_t;

static plp_t resolve_at(const orc_reads *r, int64_t i, int32_t p) {
    plp_t o = {0, 0, 0, 0, 0};
    const uint32_t *c = r->cigar + r->cigar_off[i];
    int n = r->n_cigar[i];
    int32_t x = r->pos[i], y = 0;
    if (p < x) return o;
    for (int k = 0; k < n; ++k) {
        int op = c[k] & 15;
        int32_t l = (int32_t)(c[k] >> 4);
        if (is_ref_op(op)) {
            if (p < x + l) {
                o.covered = 1;
                if (x + l - 1 == p && k + 1 < n) {
                    int op2 = c[k + 1] & 15;
                    int32_t l2 = (int32_t)(c[k + 1] >> 4);
                    if (op2 == OP_D && op != OP_D) {
                        o.indel = -l2;
                        for (int j = k + 2; j < n && (c[j] & 15) == OP_D; ++j) o.indel -= (int32_t)(c[j] >> 4);
                    } else if (op2 == OP_I) {
                        o.indel = l2;
                        for (int j = k + 2; j < n; ++j) {
                            int oj = c[j] & 15;
                            if (oj == OP_I) o.indel += (int32_t)(c[j] >> 4);
                            else if (oj != OP_P) break;
                        }
                    } else if (op2 == OP_P && k + 2 < n) {
                        int32_t l3 = 0;
                        for (int j = k + 2; j < n; ++j) {
                            int oj = c[j] & 15;
                            if (oj == OP_I) l3 += (int32_t)(c[j] >> 4);
                            else if (is_ref_op(oj)) break;
                        }
                        if (l3 > 0) o.indel = l3;
                    }
                }
                if (is_aln_op(op)) o.qpos = y + (p - x);
                else { o.qpos = y; o.is_del = 1; o.is_refskip = (op == OP_N); }
                return o;
            }
            x += l;
            if (is_aln_op(op)) y += l;
        } else if (op == OP_I || op == OP_S) {
            y += l;
        }
    }
    return o;
}

/* reference position aligned to query index q, or -1 when q is not inside an M/=/X op */
static int32_t refpos_of_query(const orc_reads *r, int64_t i, int32_t q) {
    const uint32_t *c = r->cigar + r->cigar_off[i];
    int32_t x = r->pos[i], y = 0;
    for (int k = 0; k < r->n_cigar[i]; ++k) {
        int op = c[k] & 15;
        int32_t l = (int32_t)(c[k] >> 4);
        if (is_aln_op(op)) {
            if (q < y + l) return x + (q - y);
            x += l; y += l;
        } else if (op == OP_D || op == OP_N) x += l;
        else if (op == OP_I || op == OP_S) { if (q < y + l) return -1; y += l; }
    }
    return -1;
}

/* quality of query base q of read i as seen when column `col` is emitted, i.e. after htslib
 * tweak_overlap_quality has run for every pair whose second mate has been
 * pushed by then.  bam_plp_auto pushes records until one starts beyond `col`, so that is: every
 * record with pos <= col plus the first pushed record after it (`nxt`, -1 if the fetch ended). */
static int tweaked_qual(const orc_reads *r, const int64_t *mate, int64_t i, int32_t q, int32_t col, int64_t nxt) {
    if (q >= r->l_qseq[i]) return 0;
    int qi = r->qual[r->qual_off[i] + q];
    int64_t m = mate[i];
    if (m < 0) return qi;
    if (m > i && r->pos[m] > col && m != nxt) return qi; /* second mate not pushed yet */
    int32_t rp = refpos_of_query(r, i, q);
    if (rp < 0) return qi;
    plp_t pm = resolve_at(r, m, rp);
    if (!pm.covered || pm.is_del) return qi;
    if (pm.qpos >= r->l_qseq[m]) return qi;
    int qm = r->qual[r->qual_off[m] + pm.qpos];
    int64_t a = i < m ? i : m; /* the record that arrived first holds the hash slot */
    int amul = (int)(wang(r->qname_hash[a]) & 1u);
    int mymul = (i == a) ? amul : 1 - amul;
    if (base_at(r, i, q) == base_at(r, m, pm.qpos)) {
        int s = qi + qm;
        if (s > 200) s = 200;
        return mymul * s;
    }
    if (qi > qm) return (int)(uint8_t)(0.8 * qi);
    if (qi < qm) return 0;
    return (int)(uint8_t)(mymul * 0.8 * qi);
}

static const char SEQ_CHARS[] = "=ACMGRSVTWYHKDBN";
static inline int is_ambiguous_code(int c) { /* vafator/constants.py:5 */
    char ch = SEQ_CHARS[c];
    return !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == '=');
}

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}
static int64_t lower_b(const int32_t *a, int64_t n, int32_t v) { int64_t lo = 0, hi = n; while (lo < hi) { int64_t m = (lo + hi) / 2; if (a[m] < v) lo = m + 1; else hi = m; } return lo; }
static int64_t upper_b(const int32_t *a, int64_t n, int32_t v) { int64_t lo = 0, hi = n; while (lo < hi) { int64_t m = (lo + hi) / 2; if (a[m] <= v) lo = m + 1; else hi = m; } return lo; }

typedef struct { int32_t *v; int64_t n, cap; } vec_t;
static void vpush(vec_t *a, int32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (int32_t *)realloc(a->v, a->cap * sizeof(int32_t)); }
    a->v[a->n++] = x;
}

/* vafator/pileups.py:267-290: how many times the read supports the insertion */
static int insertion_hits(const orc_reads *r, int64_t i, int32_t pos1, const char *ins, int32_t L) {
    const uint32_t *c = r->cigar + r->cigar_off[i];
    int n = r->n_cigar[i];
    /* pysam `query`: soft clips trimmed (getQueryStart / getQueryEnd) */
    int32_t qs = 0, qe = r->l_qseq[i];
    for (int k = 0; k < n; ++k) { int op = c[k] & 15; if (op == OP_H) continue; if (op == OP_S) qs += (int32_t)(c[k] >> 4); else break; }
    for (int k = n - 1; k >= 1; --k) { int op = c[k] & 15; if (op == OP_H) continue; if (op == OP_S) qe -= (int32_t)(c[k] >> 4); else break; }
    int32_t qlen = qe - qs;
    int32_t index = r->pos[i], rel = 0;
    int hits = 0;
    for (int k = 0; k < n; ++k) {
        int op = c[k] & 15;
        int32_t l = (int32_t)(c[k] >> 4);
        if (is_ref_op(op)) { index += l; if (index > pos1) break; }
        if (op == OP_M || op == OP_I || op == OP_S || op == OP_EQ || op == OP_X) rel += l;
        if (op == OP_I) {
            if (index == pos1 && l == L) {
                int ok = (rel + L <= qlen) && rel >= 0;
                for (int t = 0; ok && t < L; ++t)
                    if (SEQ_CHARS[base_at(r, i, qs + rel + t)] != ins[t]) ok = 0;
                hits += ok;
            }
        }
    }
    return hits;
}

/* vafator/pileups.py:333-347 */
static int deletion_hits(const orc_reads *r, int64_t i, int32_t pos1, int32_t L) {
    const uint32_t *c = r->cigar + r->cigar_off[i];
    int32_t start = r->pos[i];
    for (int k = 0; k < r->n_cigar[i]; ++k) {
        int op = c[k] & 15;
        int32_t l = (int32_t)(c[k] >> 4);
        if (op == OP_M || op == OP_N || op == OP_EQ || op == OP_X) start += l;
        else if (op == OP_D) {
            if (start == pos1 && l == L) return 1;
            start += l;
        }
        if (start > pos1) break;
    }
    return 0;
}

static void finish(vec_t lists[3][2], orc_counts *o) {
    for (int m = 0; m < 3; ++m) {
        for (int s = 0; s < 2; ++s) {
            vec_t *a = &lists[m][s];
            if (a->n) qsort(a->v, a->n, sizeof(int32_t), cmp_i32);
            o->n_val[m][s] = (int32_t)a->n;
            o->med2[m][s] = a->n ? a->v[(a->n - 1) / 2] + a->v[a->n / 2] : -1;
        }
        vec_t *rf = &lists[m][0], *al = &lists[m][1];
        uint64_t s2 = 0;
        for (int64_t j = 0; j < al->n; ++j) {
            int32_t v = al->v[j];
            int64_t below = lower_b(rf->v, rf->n, v) + lower_b(al->v, al->n, v);
            int64_t ties = (upper_b(rf->v, rf->n, v) - lower_b(rf->v, rf->n, v)) + (upper_b(al->v, al->n, v) - lower_b(al->v, al->n, v));
            s2 += (uint64_t)(2 * below + ties + 1);
        }
        o->s2[m] = s2;
    }
}

int orc_pileup(const orc_reads *r, const orc_variants *v, const orc_params *p, const int64_t *mate, orc_counts *out) {
    int64_t n = r->n;
    /* the largest reference span bounds how far back overlapping reads can start */
    int32_t max_span = 1;
    int32_t *span = (int32_t *)malloc((n ? n : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < n; ++i) { span[i] = ref_span(r, i); if (span[i] > max_span) max_span = span[i]; }
    int err = 0;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t u = 0; u < v->n; ++u) {
        orc_counts *o = &out[u];
        memset(o, 0, sizeof(*o));
        for (int m = 0; m < 3; ++m) o->med2[m][0] = o->med2[m][1] = -1;
        const char *ref = v->ref_txt + v->ref_off[u];
        int32_t lref = (int32_t)(v->ref_off[u + 1] - v->ref_off[u]);
        const char *alt = v->alt_txt + v->alt_off[u];
        int32_t lalt = (int32_t)(v->alt_off[u + 1] - v->alt_off[u]);
        const char *alt0 = v->alt0_txt + v->alt0_off[u];
        int32_t lalt0 = (int32_t)(v->alt0_off[u + 1] - v->alt0_off[u]);
        int kind; /* vafator/pileup_utils.py:95-119, decided on ALT[0] */
        if (lref == 1 && lalt0 == 1) kind = 0;
        else if (lref == 1 && lalt0 > 1) kind = 1;
        else if (lalt0 == 1 && lref > 1) kind = 2;
        else continue;
        o->supported = 1;
        int32_t pos1 = v->pos1[u], p0 = pos1 - 1, tid = v->tid[u];
        /* first read of (tid, pos > p0) */
        int64_t lo = 0, hi = n;
        while (lo < hi) { int64_t m = (lo + hi) / 2; if (r->tid[m] < tid || (r->tid[m] == tid && r->pos[m] <= p0)) lo = m + 1; else hi = m; }
        int64_t last = lo;
        int64_t nxt = -1;
        for (int64_t i = last; i < n && r->tid[i] == tid && r->pos[i] < v->stop[u]; ++i)
            if (read_passes(r, i, p)) { nxt = i; break; }
        vec_t L[3][2];
        memset(L, 0, sizeof(L));
        /* reads enter the column in file order */
        int64_t first = last;
        while (first > 0 && r->tid[first - 1] == tid && r->pos[first - 1] > p0 - max_span) --first;
        for (int64_t i = first; i < last; ++i) {
            if (p0 >= r->pos[i] + span[i]) continue;
            if (!read_passes(r, i, p)) continue;
            plp_t pl = resolve_at(r, i, p0);
            if (!pl.covered) continue;
            o->n_col++;
            int q = tweaked_qual(r, mate, i, pl.qpos, p0, nxt);
            if (q < p->min_baseq) continue; /* pysam pileup_base_qual_skip */
            int mq = r->mapq[i];
            if (kind == 0) {
                if (pl.is_refskip) continue;
                o->dp++;
                if (pl.is_del) continue;
                int code = base_at(r, i, pl.qpos);
                char ch = SEQ_CHARS[code];
                if (is_ambiguous_code(code)) o->n_amb++;
                if (lref == 1 && ch == ref[0]) { o->ac_ref++; vpush(&L[0][0], mq); vpush(&L[1][0], q); vpush(&L[2][0], pl.qpos); }
                if (lalt == 1 && ch == alt[0]) { o->ac_alt++; vpush(&L[0][1], mq); vpush(&L[1][1], q); vpush(&L[2][1], pl.qpos); }
            } else {
                o->dp++;
                int hits = 0;
                if (pl.indel == 0) { o->ac_ref++; vpush(&L[0][0], mq); vpush(&L[2][0], pl.qpos); continue; }
                if (kind == 1 && pl.indel > 0) hits = insertion_hits(r, i, pos1, alt0 + 1, lalt0 - 1);
                else if (kind == 2 && pl.indel < 0) hits = deletion_hits(r, i, pos1, lref - lalt0);
                for (int h = 0; h < hits; ++h) { o->ac_alt++; vpush(&L[0][1], mq); vpush(&L[2][1], pl.qpos); }
            }
        }
        finish(L, o);
        for (int m = 0; m < 3; ++m) for (int s = 0; s < 2; ++s) free(L[m][s].v);
    }
    free(span);
    return err;
}

int orc_sizeof_counts(void) { return (int)sizeof(orc_counts); }

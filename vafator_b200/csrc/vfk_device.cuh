// vfk_device.cuh -- device-side building blocks shared by the kernels (sm_100a).
//
// The kernels read the CANONICAL columnar read batch (include/vfk.h: the fields of the decoded BAM records,
// array by array) plus what prep_kernel (vfk_prep.cu) derives from it on the device in every call: the reference end
// and the CIGAR shape word of every read (one 8-byte pair) and -- per chunk / block of reads -- the running maximum of
// the ends, which is the position index.  Nothing is prepared on the host.
//
// What is evaluated per (variant column, read) pair follows pysam 0.21.0 / htslib 1.17 as
// vafator drives them (vafator/pileups.py:71-80, :184-362):
//   read filter          pysam __advance_samtools (flag filter, MAPQ, orphans) + bam_plp_push
//   column resolution    htslib resolve_cigar2 (qpos, is_del, is_refskip, indel)
//   mate overlap         htslib overlap_push / tweak_overlap_quality, applied on the fly
//   base-quality filter  pysam pileup_base_qual_skip
//   allele support       vafator/pileups.py:199-212 (SNV), :263-294 (insertion), :329-351 (deletion)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "vfk.h"

namespace vfk {

constexpr unsigned FULL = 0xFFFFFFFFu;
enum : int { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8 };
enum : unsigned { F_PAIRED = 1, F_PROPER = 2, F_UNMAP = 4, F_MUNMAP = 8 };

// CIGAR shapes (the `ev` word prep_kernel writes per read):
//   SHAPE_SIMPLE     ev == EV_SIMPLE: the CIGAR is one M/=/X op covering the whole read (qpos = column - pos)
//   SHAPE_GENERAL    ev == EV_GENERAL: any CIGAR, walked op by op
//   SHAPE_ONE_EVENT  [H][S] M [ I|D|N  M ] [S][H] -- clips and at most one event between two aligned blocks,
//                    resolved without a loop from ev = leading soft clip | first block length << 12 |
//                    EVK_* << 24 | event length << 26 (lengths < 4096 / < 64, else the read is GENERAL)
constexpr uint32_t SHAPE_SIMPLE = 0u, SHAPE_GENERAL = 1u, SHAPE_ONE_EVENT = 2u;
constexpr uint32_t EV_SIMPLE = 0u, EV_GENERAL = 0xFFFFFFFFu;
constexpr int EVK_NONE = 0, EVK_INS = 1, EVK_DEL = 2, EVK_SKIP = 3;
constexpr int PREP_BLOCK = 1024;  // reads per entry of the block-level position index (a chunk = 32 reads: the level below)
constexpr int PREP_THREADS = 256;

struct ReadsDev {
    // canonical arrays (device memory, or page-locked host memory read in place)
    const int32_t *__restrict__ pos;  // device copy when the caller's array lives on the host (the locate pass searches it)
    const uint16_t *__restrict__ flag;
    const uint8_t *__restrict__ mapq;
    const int32_t *__restrict__ l_qseq;
    const int32_t *__restrict__ n_cigar;
    const int64_t *__restrict__ cigar_off;
    const int64_t *__restrict__ seq_off;
    const int64_t *__restrict__ qual_off;
    const uint32_t *__restrict__ cigar;
    const uint8_t *__restrict__ seq;
    const uint8_t *__restrict__ qual;
    const int32_t *__restrict__ mtid;
    const int32_t *__restrict__ mpos;
    const int32_t *__restrict__ isize;
    const uint64_t *__restrict__ qname_id;
    const uint32_t *__restrict__ qname_hash;
    const int64_t *__restrict__ contig_off;
    int64_t n_reads;
    int64_t n_cigar_words, n_seq_bytes, n_qual_bytes;
    int32_t n_contigs;
    // derived per call by prep_kernel / scan_kernel (device memory)
    const uint2 *__restrict__ ee;          // {end = pos + reference bases consumed, CIGAR shape word}
    const int64_t *__restrict__ chunk_max; // [n_reads / 32]: max over the 32 reads of a chunk of tid << 32 | end
    const int64_t *__restrict__ blk_incl;  // [n_blocks + 1]: entry b + 1 = max over the reads of blocks 0..b of the same key
    // overlap partners of the reads inside the windows of listed units (link kernels): partner | tie_break << 31
    const uint32_t *__restrict__ mate;
};

struct UnitsDev {
    int64_t n;
    const int32_t *__restrict__ tid;
    const int32_t *__restrict__ pos;
    const uint32_t *__restrict__ meta;
    const int32_t *__restrict__ length;
    const uint32_t *__restrict__ seq_off;
    const uint8_t *__restrict__ ins_seq;
    const int32_t *__restrict__ stop;  // may be nullptr
};

// A unit after the locate pass: everything the pileup kernels need, one 32-byte record.
struct __align__(16) ResolvedUnit {
    int32_t col;       // 0-based column
    uint32_t meta;     // kind | ref_code << 8 | alt_code << 16
    int32_t len;       // inserted / deleted length
    uint32_t ins_off;  // offset of ALT0[1:] in ins_seq
    uint32_t lo, hi;   // candidate reads [lo, hi): hi = first read starting beyond col, lo = back[hi-1]
    int32_t stop;      // reads starting at or beyond it were never fetched by the reference
    int32_t tid;       // contig of the column (its read range ends at contig_off[tid + 1]); the candidate range is empty for an unknown contig
};

struct Unit {
    int32_t col;
    int32_t stop;
    uint32_t kind;
    uint32_t ref_code, alt_code;
    int32_t len;
    const uint8_t *ins;
    int32_t tid;
    int64_t lo, hi;
};

// One read's contribution to a column, as two words (what the shallow kernel reduces directly):
//   pk : mq | bq << 8 | qpos << 16 | PK_REF | PK_ALT   -- non-zero iff the read's values enter a reported list
//   fl : one counter per byte, ready to be summed over reads: covers | in_col << 8 | in_dp << 16 | ambiguous << 24
//   alt: how many times the values enter the ALT list (an insertion can be supported more than once)
constexpr uint32_t PK_REF = 1u << 28, PK_ALT = 1u << 29;
constexpr uint32_t FL_COVERS = 1u, FL_IN_COL = 1u << 8, FL_IN_DP = 1u << 16, FL_AMBIGUOUS = 1u << 24;
struct Eval {
    uint32_t pk, fl, alt;
    uint32_t hard;  // the read sits in a deletion / skip at the column: its base-quality filter looks at the NEXT read base, whose
                    // overlap adjustment depends on the push timing of its mate -- only the list path (mate[] + next_pushed) decides
};

// decoded view used by the histogram kernels
struct Contribution {
    bool covers;   // the read spans the column (before any filter)
    bool in_col;   // ... and is part of the htslib column (read filter passed, CIGAR covers it)
    bool in_dp;
    bool ambiguous;
    bool ref;      // value goes to the REF list
    uint32_t alt;  // how many times the value goes to the ALT list
    int mq, bq, qpos;
};
__device__ __forceinline__ Contribution decode(const Eval &e) {
    Contribution c;
    c.covers = e.fl & FL_COVERS; c.in_col = e.fl & FL_IN_COL; c.in_dp = e.fl & FL_IN_DP; c.ambiguous = e.fl & FL_AMBIGUOUS;
    c.ref = e.pk & PK_REF;
    c.alt = e.alt;
    c.mq = (int)(e.pk & 0xFFu); c.bq = (int)((e.pk >> 8) & 0xFFu); c.qpos = (int)((e.pk >> 16) & 0xFFFu);
    return c;
}

__device__ __forceinline__ bool is_ref_op(int op) { return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X; }
__device__ __forceinline__ bool is_aln_op(int op) { return op == OP_M || op == OP_EQ || op == OP_X; }

// ---- base / quality lookup: word = quality | BAM 4-bit base code << 8, straight from the canonical qual / seq pools.
// The load and the extraction are separate so that a caller can have several lookups in flight before
// it consumes the first one (own base + mate base: one round trip instead of two).
struct PayLoad {
    uint32_t q, s;  // the quality byte, the sequence byte
    uint32_t odd;
};
__device__ __forceinline__ PayLoad pay_issue(const ReadsDev &R, uint32_t i, uint32_t q) {
    PayLoad l;
    l.q = (uint32_t)__ldg(R.qual + __ldg(R.qual_off + i) + q);
    l.s = (uint32_t)__ldg(R.seq + __ldg(R.seq_off + i) + (q >> 1));
    l.odd = q & 1u;
    return l;
}
// the same pair of bytes as the host gathered them for a fetch list (vfk_pileup.cu request_kernel): quality | sequence byte << 8
__device__ __forceinline__ PayLoad pay_fetched(const uint16_t *__restrict__ fetched, uint32_t pair, uint32_t q) {
    const uint32_t g = (uint32_t)__ldg(fetched + pair);
    PayLoad l;
    l.q = g & 0xFFu;
    l.s = g >> 8;
    l.odd = q & 1u;
    return l;
}
__device__ __forceinline__ uint32_t pay_word(const PayLoad &l) { return l.q | ((l.odd ? (l.s & 15u) : (l.s >> 4)) << 8); }
__device__ __forceinline__ uint32_t pay_lookup(const ReadsDev &R, uint32_t i, uint32_t q) { return pay_word(pay_issue(R, i, q)); }
__device__ __forceinline__ int word_qual(uint32_t w) { return (int)(w & 0xFFu); }
__device__ __forceinline__ int word_base(uint32_t w) { return (int)(w >> 8); }

// ---- per-read records assembled from the arrays.  hot = {pos, span, flag | mapq << 16 | shape << 24, read index},
// cold = {cigar_off, n_cigar, l_qseq, one-event word}: the shapes the evaluation code below is written against.
__device__ __forceinline__ uint32_t shape_of_ev(uint32_t ev) { return ev == EV_SIMPLE ? SHAPE_SIMPLE : ev == EV_GENERAL ? SHAPE_GENERAL : SHAPE_ONE_EVENT; }
__device__ __forceinline__ int32_t ld_end(const ReadsDev &R, int64_t i) { return (int32_t)__ldg(&R.ee[i].x); }
__device__ __forceinline__ uint32_t ld_ev(const ReadsDev &R, int64_t i) { return __ldg(&R.ee[i].y); }
__device__ __forceinline__ uint4 load_hot(const ReadsDev &R, int64_t i) {
    const int32_t p = __ldg(R.pos + i);
    const uint32_t fm = (uint32_t)__ldg(R.flag + i) | ((uint32_t)__ldg(R.mapq + i) << 16);
    const uint2 e = __ldg(R.ee + i);
    return make_uint4((uint32_t)p, (uint32_t)((int32_t)e.x - p), fm | (shape_of_ev(e.y) << 24), (uint32_t)i);
}
__device__ __forceinline__ uint4 load_cold(const ReadsDev &R, int64_t i) {
    const uint32_t ev = ld_ev(R, i);
    return make_uint4((uint32_t)__ldg(R.cigar_off + i), (uint32_t)__ldg(R.n_cigar + i), (uint32_t)__ldg(R.l_qseq + i),
                      ev == EV_GENERAL ? 0u : ev);
}

// BAM 4-bit codes "=ACMGRSVTWYHKDBN": everything but = A C G T is IUPAC-ambiguous (vafator/constants.py:5)
__device__ __forceinline__ bool code_is_ambiguous(int c) { return ((0x0117u >> c) & 1u) == 0u; }

__device__ __forceinline__ bool read_passes(uint32_t fmk, const vfk_params &P) {
    uint32_t flag = fmk & 0xFFFFu;
    int mapq = (fmk >> 16) & 0xFF;
    if (flag & (P.flag_filter | F_UNMAP)) return false;
    if (mapq < P.min_mapq) return false;
    if (P.ignore_orphans && (flag & F_PAIRED) && !(flag & F_PROPER)) return false;
    return true;
}

// ---- column resolution ----
struct Resolved {
    bool covered;
    bool is_del, is_refskip;
    int qpos;
    int indel;
    int l_qseq;
};

// general CIGAR walk (shape 1; also correct for shape 2, which keeps its CIGAR)
__device__ __forceinline__ Resolved resolve_general(const ReadsDev &R, int32_t rpos, const uint4 &cold, int32_t col) {
    Resolved o;
    o.covered = false; o.is_del = false; o.is_refskip = false; o.qpos = 0; o.indel = 0; o.l_qseq = (int)cold.z;
    const uint32_t *__restrict__ c = R.cigar + cold.x;
    const int n = (int)cold.y;
    int32_t x = rpos, y = 0;
    for (int k = 0; k < n; ++k) {
        uint32_t w = __ldg(c + k);
        int op = w & 15;
        int32_t l = (int32_t)(w >> 4);
        if (is_ref_op(op)) {
            if (col < x + l) {
                o.covered = true;
                if (x + l - 1 == col && k + 1 < n) {  // peek: does an indel start right after this column?
                    uint32_t w2 = __ldg(c + k + 1);
                    int op2 = w2 & 15;
                    int32_t l2 = (int32_t)(w2 >> 4);
                    if (op2 == OP_D && op != OP_D) {
                        o.indel = -l2;
                        for (int j = k + 2; j < n; ++j) {
                            uint32_t wj = __ldg(c + j);
                            if ((wj & 15) != OP_D) break;
                            o.indel -= (int32_t)(wj >> 4);
                        }
                    } else if (op2 == OP_I) {
                        o.indel = l2;
                        for (int j = k + 2; j < n; ++j) {
                            uint32_t wj = __ldg(c + j);
                            int oj = wj & 15;
                            if (oj == OP_I) o.indel += (int32_t)(wj >> 4);
                            else if (oj != OP_P) break;
                        }
                    } else if (op2 == OP_P && k + 2 < n) {
                        int32_t l3 = 0;
                        for (int j = k + 2; j < n; ++j) {
                            uint32_t wj = __ldg(c + j);
                            int oj = wj & 15;
                            if (oj == OP_I) l3 += (int32_t)(wj >> 4);
                            else if (is_ref_op(oj)) break;
                        }
                        if (l3 > 0) o.indel = l3;
                    }
                }
                if (is_aln_op(op)) {
                    o.qpos = y + (col - x);
                } else {
                    o.qpos = y;
                    o.is_del = true;
                    o.is_refskip = (op == OP_N);
                }
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

// shape-2 read: [H][S lead] M m1 [ (I|D|N) e  M m2 ] [S][H] -- at most one event between two aligned blocks.
// cold.w = lead | m1 << 12 | event << 24 | e << 26 (vfk.h); same answers as resolve_general, no loop.
__device__ __forceinline__ Resolved resolve_one_event(int32_t rpos, const uint4 &cold, int32_t col) {
    Resolved o;
    const int lead = (int)(cold.w & 0xFFFu), m1 = (int)((cold.w >> 12) & 0xFFFu), ev = (int)((cold.w >> 24) & 3u), e = (int)(cold.w >> 26);
    const int d = col - rpos;                   // 0 <= d < span (the caller checked)
    const int gap = ev >= EVK_DEL ? e : 0;   // D / N consume reference
    o.covered = true; o.l_qseq = (int)cold.z; o.is_del = false; o.is_refskip = false; o.indel = 0;
    if (d < m1) {
        o.qpos = lead + d;
        if (d == m1 - 1) o.indel = ev == EVK_INS ? e : ev == EVK_DEL ? -e : 0;
    } else if (d < m1 + gap) {
        o.qpos = lead + m1;
        o.is_del = true;
        o.is_refskip = ev == EVK_SKIP;
    } else {
        o.qpos = lead + m1 + (ev == EVK_INS ? e : 0) + (d - m1 - gap);
    }
    return o;
}

// reference position aligned to query index q of a general read, -1 if q is not in an M/=/X op
__device__ __forceinline__ int32_t refpos_of_query(const ReadsDev &R, int32_t rpos, const uint4 &cold, int q) {
    const uint32_t *__restrict__ c = R.cigar + cold.x;
    int32_t x = rpos, y = 0;
    for (int k = 0; k < (int)cold.y; ++k) {
        uint32_t w = __ldg(c + k);
        int op = w & 15;
        int32_t l = (int32_t)(w >> 4);
        if (is_aln_op(op)) {
            if (q < y + l) return x + (q - y);
            x += l; y += l;
        } else if (op == OP_D || op == OP_N) {
            x += l;
        } else if (op == OP_I || op == OP_S) {
            if (q < y + l) return -1;
            y += l;
        }
    }
    return -1;
}

// end of the read range of contig `tid` (only for units with candidate reads: their contig is known)
__device__ __forceinline__ int64_t contig_end(const ReadsDev &R, int32_t tid) { return __ldg(R.contig_off + tid + 1); }

// first read at or after `from` (same contig, start < stop) that the reference would have pushed
__device__ __forceinline__ int64_t next_pushed(const ReadsDev &R, const vfk_params &P, int64_t from, int64_t c1, int32_t stop) {
    for (int64_t j = from; j < c1; ++j) {
        if (__ldg(R.pos + j) >= stop) break;
        if (read_passes((uint32_t)__ldg(R.flag + j) | ((uint32_t)__ldg(R.mapq + j) << 16), P)) return j;
    }
    return -1;
}

// ---- htslib overlap_push (sam.c): which records reach the name hash, and which store themselves ----
__device__ __forceinline__ uint32_t wang_hash(uint32_t k) {
    k += ~(k << 15); k ^= (k >> 10); k += (k << 3); k ^= (k >> 6); k += ~(k << 11); k ^= (k >> 16);
    return k;
}
__device__ __forceinline__ uint32_t fm_of(const ReadsDev &R, int64_t i) { return (uint32_t)__ldg(R.flag + i) | ((uint32_t)__ldg(R.mapq + i) << 16); }

// the column the engine stands on while record i is pushed: the start of the record pushed just before it
// (bam_plp_push appends iff end > that column); the first pushed record of a contig sees its own start
__device__ __forceinline__ int32_t push_column(const ReadsDev &R, const vfk_params &P, int64_t i, int64_t c0) {
    for (int64_t j = i - 1; j >= c0; --j)
        if (read_passes(fm_of(R, j), P)) return __ldg(R.pos + j);
    return __ldg(R.pos + i);
}
// does record i reach the name hash (overlap_push past its early exits)?
__device__ __forceinline__ bool reaches_hash(const ReadsDev &R, const vfk_params &P, int64_t i, int32_t tid, int32_t column) {
    const uint32_t fm = fm_of(R, i);
    if (!read_passes(fm, P)) return false;
    const int32_t end = ld_end(R, i);
    if (end <= column) return false;  // never appended to the buffer
    if ((fm & F_MUNMAP) || !(fm & F_PROPER)) return false;
    const int32_t mt = __ldg(R.mtid + i);
    if (mt >= 0 && mt != tid) return false;
    const long long isz = (long long)__ldg(R.isize + i);
    if ((isz < 0 ? -isz : isz) >= 2LL * __ldg(R.l_qseq + i) && __ldg(R.mpos + i) >= end) return false;
    return true;
}
__device__ __forceinline__ bool stores_itself(const ReadsDev &R, int64_t i) {
    const int32_t mp = __ldg(R.mpos + i);
    return mp >= __ldg(R.pos + i) || ((__ldg(R.flag + i) & F_PAIRED) && mp == -1);
}


// Quality of query base q of read i as the reference sees it when column `col` is emitted:
// htslib tweak_overlap_quality has run for the pair iff the second mate has been pushed, i.e. it
// starts at or before `col` or is the first pushed record after it (bam_plp_auto reads one ahead).
// `own` is the read's own payload lookup, already issued; it is consumed only after the mate's lookup
// has been issued too.  Returns the quality; `my_word` receives the read's own payload word.
__device__ __forceinline__ int tweaked_quality(const ReadsDev &R, const vfk_params &P, const Unit &U, int64_t i,
                                               const uint4 &hot, uint32_t mw, const uint4 &cold_if_general, bool general,
                                               int q, int l_qseq, bool q_is_column_base, const PayLoad &own, uint32_t &my_word) {
    my_word = 0u;
    if (q >= l_qseq) return 0;
    const uint32_t flag = hot.z & 0xFFFFu;
    bool paired = P.ignore_overlaps && (flag & F_PROPER) && !(flag & F_MUNMAP) && mw != VFK_NO_MATE;
    PayLoad mate;
    mate.q = mate.s = mate.odd = 0u;
    int64_t m = 0;
    if (paired) {
        m = (int64_t)(mw & 0x7FFFFFFFu);
        const uint4 mh = load_hot(R, m);
        if (m > i && (int32_t)mh.x > U.col)  // the mate starts beyond the column: it can only matter to a base further on
            paired = !q_is_column_base && next_pushed(R, P, U.hi, contig_end(R, U.tid), U.stop) == m;
        int32_t rp = -1;
        if (paired) rp = q_is_column_base ? U.col : (general ? refpos_of_query(R, (int32_t)hot.x, cold_if_general, q)
                                                             : (int32_t)hot.x + q);
        paired = paired && rp >= (int32_t)mh.x && rp < (int32_t)mh.x + (int32_t)mh.y;  // rp < 0: not an aligned base
        if (paired) {
            int mq_index, ml;
            if ((mh.z >> 24) == SHAPE_SIMPLE) {
                mq_index = rp - (int32_t)mh.x;
                ml = (int)mh.y;
            } else {
                const uint4 mc = load_cold(R, m);
                const Resolved mr = (mh.z >> 24) == SHAPE_ONE_EVENT ? resolve_one_event((int32_t)mh.x, mc, rp)
                                                                        : resolve_general(R, (int32_t)mh.x, mc, rp);
                paired = mr.covered && !mr.is_del;
                mq_index = mr.qpos;
                ml = mr.l_qseq;
            }
            paired = paired && mq_index < ml;
            if (paired) mate = pay_issue(R, mh.w, (uint32_t)mq_index);
        }
    }
    my_word = pay_word(own);
    const int qi = word_qual(my_word);
    if (!paired) return qi;
    const uint32_t mate_word = pay_word(mate);
    const int qm = word_qual(mate_word);
    const bool i_first = i < m;  // the earlier record holds the hash slot ("a" in htslib)
    const uint32_t sel = mw >> 31;
    const uint32_t mymul = i_first ? sel : (1u - sel);
    if (word_base(my_word) == word_base(mate_word)) {
        int s = qi + qm;
        if (s > 200) s = 200;
        return mymul ? s : 0;
    }
    // (uint8_t)(0.8 * q) in IEEE double == floor(4q/5) for every q in 0..255 (tests/test_host.py)
    if (qi > qm) return (qi * 4) / 5;
    if (qi < qm) return 0;
    return mymul ? (qi * 4) / 5 : 0;
}

// vafator/pileups.py:267-290 -- how many times the read supports the insertion
__device__ __forceinline__ uint32_t insertion_hits(const ReadsDev &R, const uint4 &hot, const uint4 &cold, const Unit &U) {
    const uint32_t *__restrict__ c = R.cigar + cold.x;
    const int n = (int)cold.y;
    const int32_t pos1 = U.col + 1;
    int32_t qs = 0, qe = (int32_t)cold.z;  // pysam `query`: soft clips trimmed at both ends
    for (int k = 0; k < n; ++k) {
        uint32_t w = __ldg(c + k);
        int op = w & 15;
        if (op == OP_H) continue;
        if (op == OP_S) qs += (int32_t)(w >> 4); else break;
    }
    for (int k = n - 1; k >= 1; --k) {
        uint32_t w = __ldg(c + k);
        int op = w & 15;
        if (op == OP_H) continue;
        if (op == OP_S) qe -= (int32_t)(w >> 4); else break;
    }
    const int32_t qlen = qe - qs;
    int32_t index = (int32_t)hot.x, rel = 0;
    uint32_t hits = 0;
    for (int k = 0; k < n; ++k) {
        uint32_t w = __ldg(c + k);
        int op = w & 15;
        int32_t l = (int32_t)(w >> 4);
        if (is_ref_op(op)) {
            index += l;
            if (index > pos1) break;
        }
        if (op == OP_M || op == OP_I || op == OP_S || op == OP_EQ || op == OP_X) rel += l;
        if (op == OP_I && index == pos1 && l == U.len) {
            // the reference compares the L bases FOLLOWING the insertion (rel already includes it,
            // and any leading soft clip) against ALT[0][1:]  -- SURVEY.md A.5
            bool ok = rel + U.len <= qlen;
            for (int t = 0; ok && t < U.len; ++t)
                ok = word_base(pay_lookup(R, hot.w, (uint32_t)(qs + rel + t))) == (int)U.ins[t];
            hits += ok ? 1u : 0u;
        }
    }
    return hits;
}

// vafator/pileups.py:333-347
__device__ __forceinline__ uint32_t deletion_hits(const ReadsDev &R, const uint4 &hot, const uint4 &cold, const Unit &U) {
    const uint32_t *__restrict__ c = R.cigar + cold.x;
    const int32_t pos1 = U.col + 1;
    int32_t start = (int32_t)hot.x;
    for (int k = 0; k < (int)cold.y; ++k) {
        uint32_t w = __ldg(c + k);
        int op = w & 15;
        int32_t l = (int32_t)(w >> 4);
        if (op == OP_M || op == OP_N || op == OP_EQ || op == OP_X) {
            start += l;
        } else if (op == OP_D) {
            if (start == pos1 && l == U.len) return 1u;
            start += l;
        }
        if (start > pos1) break;
    }
    return 0u;
}

// Everything one read contributes to one unit.  `hot` / `mw` are the read's header and mate word
// (the caller loads them, possibly one unit ahead of use).
__device__ __forceinline__ Eval evaluate_packed(const ReadsDev &R, const vfk_params &P, const Unit &U, int64_t i, const uint4 &hot,
                                                uint32_t mw) {
    Eval o;
    o.pk = 0u; o.fl = 0u; o.alt = 0u; o.hard = 0u;
    const int32_t rpos = (int32_t)hot.x;
    if ((uint32_t)(U.col - rpos) >= hot.y) return o;  // col < pos wraps to a huge unsigned
    o.fl = FL_COVERS;
    if (!read_passes(hot.z, P)) return o;
    const uint32_t shape = hot.z >> 24;
    const bool general = shape != SHAPE_SIMPLE;
    uint4 cold = make_uint4(0, 0, 0, 0);
    Resolved r;
    if (!general) {
        r.covered = true; r.is_del = false; r.is_refskip = false; r.indel = 0;
        r.qpos = U.col - rpos;
        r.l_qseq = (int)hot.y;
    } else {
        cold = load_cold(R, i);
        if (shape == SHAPE_ONE_EVENT) {
            r = resolve_one_event(rpos, cold, U.col);
        } else {
            r = resolve_general(R, rpos, cold, U.col);
            if (!r.covered) return o;
        }
    }
    o.fl |= FL_IN_COL;
    o.hard = (r.is_del && r.qpos < r.l_qseq) ? 1u : 0u;
    PayLoad own;
    own.q = own.s = own.odd = 0u;
    if (r.qpos < r.l_qseq) own = pay_issue(R, hot.w, (uint32_t)r.qpos);
    uint32_t my_word;
    const int q = tweaked_quality(R, P, U, i, hot, mw, cold, general, r.qpos, r.l_qseq, !r.is_del, own, my_word);
    if (q < P.min_baseq) return o;  // pysam pileup_base_qual_skip
    const uint32_t mq_pos = ((hot.z >> 16) & 0xFFu) | ((uint32_t)r.qpos << 16);
    if (U.kind == VFK_KIND_SNV) {
        if (r.is_refskip) return o;
        o.fl |= FL_IN_DP;
        if (r.is_del) return o;  // counted in DP as the "" allele, belongs to no reported list
        const uint32_t code = (uint32_t)word_base(my_word);
        if (code_is_ambiguous((int)code)) o.fl |= FL_AMBIGUOUS;
        const uint32_t lists = (code == U.ref_code ? PK_REF : 0u) | (code == U.alt_code ? PK_ALT : 0u);
        if (lists) o.pk = mq_pos | ((uint32_t)q << 8) | lists;
        o.alt = code == U.alt_code ? 1u : 0u;
    } else {
        o.fl |= FL_IN_DP;
        if (r.indel == 0) {
            o.pk = mq_pos | PK_REF;  // "considers all reads without indels to be the reference" (pileups.py:291-294)
        } else if (general) {
            if (U.kind == VFK_KIND_INS && r.indel > 0) o.alt = insertion_hits(R, hot, cold, U);
            else if (U.kind == VFK_KIND_DEL && r.indel < 0) o.alt = deletion_hits(R, hot, cold, U);
            if (o.alt) o.pk = mq_pos | PK_ALT;
        }
    }
    return o;
}

// the rare, long pieces of the register kernel's per-read code, out of line: its hot loop should fit the instruction cache
static __device__ __noinline__ uint32_t indel_hits(const ReadsDev &R, int64_t i, uint32_t kind, int32_t col, int32_t len, const uint8_t *ins) {
    Unit U;
    U.col = col; U.len = len; U.ins = ins; U.kind = kind; U.stop = 0; U.ref_code = U.alt_code = 0u; U.tid = 0; U.lo = U.hi = 0;
    const uint4 hot = load_hot(R, i), cold = load_cold(R, i);
    return kind == VFK_KIND_INS ? insertion_hits(R, hot, cold, U) : deletion_hits(R, hot, cold, U);
}
static __device__ __noinline__ Resolved resolve_general_call(const ReadsDev &R, int64_t i, int32_t rpos, int32_t col) {
    return resolve_general(R, rpos, load_cold(R, i), col);
}

// ---- the same evaluation in three steps, for the register kernel: all candidate reads of a column sit in the lanes of one
// warp, so a read's overlap partner is another lane and its base / quality at the column arrive by a shuffle instead of a
// second gather (resolve_column -> payload loads in flight -> exchange -> finish_column).
struct ColRead {
    uint32_t fl;   // FL_COVERS | FL_IN_COL
    bool aligned;  // in the column at an aligned base inside the read: its base / quality at the column exist
    bool hard;     // in the column inside a deletion / skip, with a next base: the base-quality filter looks at that base (general protocol)
    bool is_del, is_refskip;
    bool pay;      // a base / quality lookup is due (aligned or hard)
    int qpos, indel;
};
// `passes`: the read is valid and passes the read filter (read_passes)
__device__ __forceinline__ ColRead resolve_column(const ReadsDev &R, int32_t col, int64_t i, const uint4 &hot, bool valid, bool passes) {
    ColRead c;
    c.fl = 0u; c.aligned = false; c.hard = false; c.is_del = false; c.is_refskip = false; c.pay = false; c.qpos = 0; c.indel = 0;
    const int32_t rpos = (int32_t)hot.x;
    if (!valid || (uint32_t)(col - rpos) >= hot.y) return c;  // col < pos wraps to a huge unsigned
    c.fl = FL_COVERS;
    if (!passes) return c;
    const uint32_t shape = hot.z >> 24;
    int l_qseq = (int)hot.y;
    if (shape == SHAPE_SIMPLE) {
        c.qpos = col - rpos;
    } else {
        const Resolved r = shape == SHAPE_ONE_EVENT ? resolve_one_event(rpos, load_cold(R, i), col) : resolve_general_call(R, i, rpos, col);
        if (!r.covered) return c;
        c.is_del = r.is_del; c.is_refskip = r.is_refskip; c.qpos = r.qpos; c.indel = r.indel;
        l_qseq = r.l_qseq;
    }
    c.fl |= FL_IN_COL;
    c.pay = c.qpos < l_qseq;
    c.aligned = !c.is_del && c.pay;
    c.hard = c.is_del && c.pay;
    return c;
}
// htslib tweak_overlap_quality for a read and its partner that both show an aligned base at the column.
// partner: quality | base << 8 | 1 << 31 when the partner has such a base, else 0; sel = the name-hash tie-break bit.
__device__ __forceinline__ int overlap_quality(uint32_t my_word, uint32_t partner, bool i_first, uint32_t sel) {
    const int qi = word_qual(my_word);
    if (!(partner >> 31)) return qi;
    const int qm = (int)(partner & 0xFFu);
    const uint32_t mymul = i_first ? sel : (1u - sel);
    if (((my_word ^ partner) & 0xF00u) == 0u) {
        int s = qi + qm;
        if (s > 200) s = 200;
        return mymul ? s : 0;
    }
    if (qi > qm) return (qi * 4) / 5;
    if (qi < qm) return 0;
    return mymul ? (qi * 4) / 5 : 0;
}
// quality `q` (after the overlap adjustment) and payload word known: filter + allele classification
__device__ __forceinline__ Eval finish_column(const ReadsDev &R, const vfk_params &P, const Unit &U, int64_t i, const uint4 &hot, const ColRead &c,
                                              int q, uint32_t my_word) {
    Eval o;
    o.pk = 0u; o.fl = c.fl; o.alt = 0u; o.hard = 0u;
    if (!(c.fl & FL_IN_COL)) return o;
    if (q < P.min_baseq) return o;  // pysam pileup_base_qual_skip
    const uint32_t mq_pos = ((hot.z >> 16) & 0xFFu) | ((uint32_t)c.qpos << 16);
    if (U.kind == VFK_KIND_SNV) {
        if (c.is_refskip) return o;
        o.fl |= FL_IN_DP;
        if (c.is_del) return o;  // counted in DP as the "" allele, belongs to no reported list
        const uint32_t code = (uint32_t)word_base(my_word);
        if (code_is_ambiguous((int)code)) o.fl |= FL_AMBIGUOUS;
        const uint32_t lists = (code == U.ref_code ? PK_REF : 0u) | (code == U.alt_code ? PK_ALT : 0u);
        if (lists) o.pk = mq_pos | ((uint32_t)q << 8) | lists;
        o.alt = code == U.alt_code ? 1u : 0u;
    } else {
        o.fl |= FL_IN_DP;
        if (c.indel == 0) {
            o.pk = mq_pos | PK_REF;  // "considers all reads without indels to be the reference" (pileups.py:291-294)
        } else if ((hot.z >> 24) != SHAPE_SIMPLE) {
            if ((U.kind == VFK_KIND_INS && c.indel > 0) || (U.kind == VFK_KIND_DEL && c.indel < 0)) o.alt = indel_hits(R, i, U.kind, U.col, U.len, U.ins);
            if (o.alt) o.pk = mq_pos | PK_ALT;
        }
    }
    return o;
}

__device__ __forceinline__ Contribution evaluate(const ReadsDev &R, const vfk_params &P, const Unit &U, int64_t i, const uint4 &hot,
                                                 uint32_t mw) {
    return decode(evaluate_packed(R, P, U, i, hot, mw));
}
__device__ __forceinline__ Contribution evaluate(const ReadsDev &R, const vfk_params &P, const Unit &U, int64_t i) {
    return evaluate(R, P, U, i, load_hot(R, i), __ldcg(R.mate + i));  // written by the link passes, possibly earlier in the same launch: not the read-only path
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk) completed on an mbarrier ----
// Used by the deep kernel to stage the next 32 read headers of a column in shared memory while the
// current 32 are evaluated.  SASS: UBLKCP / SYNCS.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void unit_from(const ResolvedUnit &r, const uint8_t *__restrict__ ins_seq, Unit &U) {
    U.col = r.col;
    U.stop = r.stop;
    U.kind = r.meta & 3u;
    U.ref_code = (r.meta >> 8) & 0xFFu;
    U.alt_code = (r.meta >> 16) & 0xFFu;
    U.len = r.len;
    U.ins = ins_seq + r.ins_off;
    U.tid = r.tid;
    U.lo = (int64_t)r.lo;
    U.hi = (int64_t)r.hi;
}

}  // namespace vfk

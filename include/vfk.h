/* vfk.h -- C ABI of the B200 pileup-annotation library (libvfk.so).
 *
 * The reference (TRON-Bioinformatics/vafator v3.0.0) is pure Python and has no FFI; the seam
 * this library fills is three Python call sites inside vafator/annotator.py:
 *
 *   collect_metrics_for_chrom(...)      vafator/pileups.py:106-159   (called at annotator.py:225-232, :56-63)
 *     -> vfk_pileup()                   per-(variant, BAM) allele counts, depth, medians and
 *                                       rank-sum sufficient statistics, all integers
 *   get_rank_sum_tests(...)             vafator/rank_sum_test.py:15-26 (called at annotator.py:350, :417)
 *   PowerCalculator.calculate_power / calculate_absolute_power
 *                                       vafator/power.py:39-50, :116-131 (annotator.py:325-333, :394-400)
 *     -> vfk_epilogue()                 FP64 z / p / pu / pw / k from those integers
 *   vfk_annotate_host() runs both with HOST buffers (pinned or pageable), copies included.
 *
 * Conventions: every function returns 0 (VFK_OK) or a negative VFK_E* code and never throws or
 * aborts; vfk_last_error() gives the thread-local message.  All buffers are caller owned; the
 * library keeps no reference to them after the stream work it enqueued has completed.  Calls
 * are asynchronous with respect to the host until vfk_sync() (except vfk_annotate_host, which
 * returns after its results are in host memory).  `stream` is a cudaStream_t passed as void*
 * (NULL = the CUDA default stream; the handle's own streams are used by vfk_annotate_host only).
 * A handle belongs to one device and owns per-launch scratch (located units, deferred-unit lists):
 * drive it from one host thread at a time and keep the vfk_pileup calls of one handle on one stream
 * (stream order is what protects the scratch); use one handle per stream / per host thread for
 * concurrency.  Different handles are independent.
 */
#ifndef VFK_H
#define VFK_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFK_VERSION 100 /* 0.1.0 */

enum {
    VFK_OK = 0,
    VFK_EINVAL = -1,       /* bad argument */
    VFK_ECUDA = -2,        /* CUDA runtime error (message in vfk_last_error) */
    VFK_ENOMEM = -3,
    VFK_EUNSUPPORTED = -4, /* e.g. reads longer than VFK_MAX_READ_LEN */
    VFK_ENODEVICE = -5
};

#define VFK_MAX_READ_LEN 4095 /* longest l_qseq the read-position histograms cover */
#define VFK_NO_MATE 0xFFFFFFFFu
#define VFK_SHAPE_SIMPLE 0u
#define VFK_SHAPE_GENERAL 1u
#define VFK_SHAPE_ONE_EVENT 2u
#define VFK_EV_NONE 0
#define VFK_EV_INS 1
#define VFK_EV_DEL 2
#define VFK_EV_SKIP 3
#define VFK_CODE_NONE 0xFFu

/* ---- read batch, HBM layout (DESIGN.md "Data layout") -------------------------------------
 * Reads are coordinate sorted (BAM order); unplaced reads are not part of the batch.          */
typedef struct {
    int32_t pos;   /* 0-based leftmost reference position                                      */
    uint32_t span; /* reference bases consumed by the CIGAR (M,D,N,=,X); end = pos + span       */
    uint32_t fmk;  /* flag (bits 0-15) | mapq << 16 | shape << 24;
                      shape 0 (VFK_SHAPE_SIMPLE): the CIGAR is one M/=/X op (qpos = column - pos,
                              l_qseq = span); the cold record is never read,
                      shape 1 (VFK_SHAPE_GENERAL): any CIGAR, walked op by op,
                      shape 2 (VFK_SHAPE_ONE_EVENT): [H][S] M [ I|D|N  M ] [S][H] -- clips and at most
                              one event between two aligned blocks; resolved without a loop from
                              cold.event (the CIGAR words are kept for the indel-support tests)  */
    uint32_t pay;  /* payload offset, in 32-byte sectors                                       */
} vfk_read_hot;

typedef struct {
    uint32_t cigar_off; /* index into cigar[] (BAM encoding len<<4|op)                         */
    uint32_t n_cigar;
    uint32_t l_qseq;
    uint32_t event; /* shape 2: leading soft clip | first block length << 12 | VFK_EV_* << 24 |
                       event length << 26 (lengths < 4096 / < 64, else the read is shape 1); 0 otherwise */
} vfk_read_cold;

/* payload: read i owns ceil(l_qseq/20) sectors of 32 bytes from hot.pay.  A sector is four 8-byte
 * little-endian granules; granule g of the read holds query positions 5g .. 5g+4, 12 bits each:
 * bits [12r, 12r+8) base quality, bits [12r+8, 12r+12) BAM 4-bit base code.  One (read, column)
 * lookup = one 64-bit load = one 32-byte DRAM sector (BAM's split seq / qual arrays cost two).      */
/* ---- column store (optional; built by vfkio_columns_*, include/vfk_io.h) --------------------------------
 * The reference cut into windows of VFK_COL_WIDTH positions; every read leaves one 32-byte sector in every
 * window it overlaps, the sectors of a window stored together in read order:
 *   col_wbase[t]            first window of contig t; window of (t, pos) = col_wbase[t] + pos / VFK_COL_WIDTH
 *                           (contig t has col_wbase[t+1] - col_wbase[t] windows: up to its last covered base)
 *   col_off[w], col_off[w+1] the sectors of window w
 *   sector                  12 entries of 16 bits (position p of the window in entry p): base quality |
 *                           BAM base code << 8 | state << 12, state 0 = not covered, 1 = aligned base,
 *                           2 = deletion, 3 = reference skip; then two header words:
 *     h0  query position at the first covered position of the window (12 bits) | MAPQ << 12 |
 *         signed sector offset of the overlap partner inside the window << 20 (0 = none) |
 *         read passes the read filter << 28 | tie-break bit << 29 | read precedes its partner << 30 |
 *         sector cannot be used (two insertions in the window, partner out of reach) << 31
 *     h1  aligned-position mask (12 bits) | position the insertion follows << 12 | its length << 16 |
 *         has insertion << 28;  query position of entry p = h0.qpos + popcount(mask below p) + (insertion
 *         before p ? length : 0)
 * The read filter verdict and the partner links depend on vfk_params: build the store with the parameters
 * the kernels are called with.  With a column store present SNV units are evaluated from it (one contiguous
 * read per column); units it cannot serve go to the read-major path.                                      */
#define VFK_COL_WIDTH 12

typedef struct {
    int64_t n_reads;
    int32_t n_contigs;
    int32_t max_l_qseq;
    const int64_t *contig_off;    /* [n_contigs+1] reads of contig t are [off[t], off[t+1])      */
    const vfk_read_hot *hot;      /* [n_reads]                                                  */
    const vfk_read_cold *cold;    /* [n_reads]                                                  */
    const uint32_t *mate;         /* [n_reads] overlap partner | tie_break << 31, VFK_NO_MATE   */
    const int32_t *sb;            /* [n_reads][2] {start, back}: the position index;
                                     back[i] = first read of the contig whose running-maximum end
                                     exceeds start[i], i.e. no earlier read can overlap it        */
    const uint32_t *cigar;        /* [n_cigar_words]                                            */
    const uint8_t *payload;       /* [n_sectors * 32], 32-byte aligned                          */
    int64_t n_cigar_words;
    int64_t n_sectors;
    /* column store, all NULL / 0 when absent */
    const int64_t *col_wbase;     /* [n_contigs+1]                                              */
    const uint32_t *col_off;      /* [n_col_windows+1]                                          */
    const uint8_t *col_sectors;   /* [n_col_sectors * 32], 32-byte aligned                      */
    int64_t n_col_windows;
    int64_t n_col_sectors;
} vfk_read_batch;

/* ---- variant units: one per (VCF record, ALT allele) that needs a column evaluation --------- */
#define VFK_KIND_SNV 0u
#define VFK_KIND_INS 1u
#define VFK_KIND_DEL 2u
typedef struct {
    int64_t n_units;
    const int32_t *tid;      /* contig index in the read batch                                  */
    const int32_t *pos;      /* 0-based column (POS - 1)                                        */
    const uint32_t *meta;    /* kind | ref_code << 8 | alt_code << 16 (BAM 4-bit code or VFK_CODE_NONE) */
    const int32_t *length;   /* inserted / deleted length (len(ALT0)-len(REF) / len(REF)-len(ALT0)) */
    const uint32_t *seq_off; /* insertions: offset of ALT0[1:] in ins_seq                       */
    const uint8_t *ins_seq;  /* one 4-bit code per byte, VFK_CODE_NONE = can never match        */
    const int32_t *stop;     /* pysam fetch stop of the unit's chromosome group = its last POS
                                (vafator/pileups.py:137-138); may be NULL (= no bound)          */
    int64_t n_ins_seq;
} vfk_variant_batch;

/* ---- parameters (vafator/pileups.py:71-80 keyword arguments + pysam 0.21.0 defaults) -------- */
typedef struct {
    int32_t min_mapq;          /* --mapping-quality   (CLI default 1, Annotator default 0)       */
    int32_t min_baseq;         /* --base-call-quality (CLI default 30, Annotator default 29)     */
    uint32_t flag_filter;      /* 0x704: UNMAP|SECONDARY|QCFAIL|DUP                              */
    int32_t ignore_orphans;    /* 1 */
    int32_t ignore_overlaps;   /* 1 */
    int32_t include_ambiguous; /* 1 unless --exclude-ambiguous-bases                             */
} vfk_params;

/* ---- pileup output: integers only ---------------------------------------------------------- */
#define VFK_ST_KIND_MASK 0x3u
#define VFK_ST_DEFERRED 0x10u  /* column deeper than the warp path handles; done by the block path */
#define VFK_ST_READLEN 0x20u   /* a read position fell outside the histogram: result invalid      */
typedef struct {
    int32_t dp;         /* depth as vafator reports it (pileups.py:224-227, :264, :330)          */
    int32_t n_amb;      /* SNV: IUPAC-ambiguous bases in the column (annotator.py:386)           */
    int32_t ac_ref;     /* SNV: reads showing REF; indel: reads with no indel here (REF support)  */
    int32_t ac_alt;     /* reads supporting this unit's ALT                                       */
    int32_t med2[3][2]; /* [mq,bq,pos][ref,alt]: twice the median, -1 when the list is empty      */
    uint32_t status;
    uint32_t n_cand;    /* reads inspected for the unit (diagnostic): candidate reads of the position index on
                           the read-major path, sectors of the column's window on the column-store path      */
    uint64_t s2[3];     /* [mq,bq,pos]: twice the sum of the ALT mid-ranks in the pooled ranking   */
    uint32_t n_cover;   /* reads overlapping the column before any filter (D_raw of SURVEY.md 8d)   */
    uint32_t n_col;     /* reads in the htslib column (read filter passed) BEFORE the base-quality filter:
                           0 means the reference sees no column at all (-> EMPTY_METRICS)             */
} vfk_counts; /* 80 bytes */

/* ---- epilogue output: FP64, unrounded ------------------------------------------------------ */
typedef struct {
    double z[3];  /* [mq,bq,pos] rank-sum statistic as scipy.stats.ranksums(alt, ref); NaN if a side is empty */
    double p[3];  /* two-sided p-value */
    double pu;    /* binom.cdf(ac_alt, dp, eaf)          vafator/power.py:39-50   */
    double pw;    /* Carter-2012 power                   vafator/power.py:116-131 */
    int32_t k;    /* minimum supporting reads            vafator/power.py:106-114 */
    int32_t pad;
} vfk_stats; /* 72 bytes */

typedef struct vfk_handle vfk_handle;

int vfk_version(void);
const char *vfk_last_error(void);
int vfk_create(int device, vfk_handle **out);
int vfk_destroy(vfk_handle *h);
int vfk_sync(vfk_handle *h);

/* all pointers inside the batches and `out` are DEVICE pointers */
int vfk_pileup(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
               const vfk_params *params, vfk_counts *out, void *stream);

/* counts / eaf / out are DEVICE pointers; eaf[u] = expected VAF of unit u (vafator/power.py:52-82) */
int vfk_epilogue(vfk_handle *h, int64_t n_units, const vfk_counts *counts, const double *eaf,
                 double fpr, double error_rate, vfk_stats *out, void *stream);

/* Same path with HOST buffers: copies the batches to the device (staging memory owned by the
 * handle, grown on demand), runs both kernels and copies counts/stats back before returning.   */
int vfk_annotate_host(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                      const vfk_params *params, const double *eaf, double fpr, double error_rate,
                      vfk_counts *counts_out, vfk_stats *stats_out);

/* Double-buffered variant: returns once the batch is enqueued on one of the handle's two staging
 * slots; *ticket (0 or 1) names the slot.  Submitting a third batch waits for the slot it reuses.
 * The host buffers (inputs and outputs) must stay valid until vfk_wait(h, ticket) returns.
 * How a batch crosses the bus is chosen per batch:
 *  - sparse variant set (n_units * 64 <= n_reads): the units are located on the host from the
 *    caller's position index, 32 bytes per unit are copied and the index is not; otherwise the
 *    index is copied and the locate kernel runs            (VFK_HOST_LOCATE=0/1 forces the choice);
 *  - page-locked payload / cold / cigar buffers and few candidate reads: the kernel gathers the
 *    sectors it needs straight from host memory instead of copying those arrays
 *    (VFK_ZERO_COPY_PAYLOAD=0/1), and so it does for the hot / mate arrays if they are page-locked
 *    too (VFK_ZERO_COPY_HOT=0 keeps copying them).
 * Results are identical whichever way is taken.                                                   */
int vfk_annotate_host_async(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                            const vfk_params *params, const double *eaf, double fpr, double error_rate,
                            vfk_counts *counts_out, vfk_stats *stats_out, int *ticket);
int vfk_wait(vfk_handle *h, int ticket);
/* how many batches of this handle took the zero-copy route */
int64_t vfk_zero_copy_batches(vfk_handle *h);
/* transfer accounting of the host entry point since vfk_create: out[0] bytes copied host -> device by
 * bulk copies, out[1] batches whose payload was gathered (zero copy), out[2] batches whose hot / mate
 * headers were gathered too, out[3] batches located on the host                                    */
int vfk_transfer_stats(vfk_handle *h, int64_t out[4]);
/* how many batches of the host entry point evaluated their SNV units from the column store (page-locked
 * buffers, units located on the host; VFK_HOST_COLUMNS=0 keeps to the read-major arrays)               */
int64_t vfk_column_batches(vfk_handle *h);

/* ---- list-shaped auxiliaries (DEVICE pointers) ------------------------------------------------
 * vfk_pileup_values: per unit, the per-read values the reference keeps in all_mqs / all_bqs /
 * all_positions (vafator/pileups.py:214-216), one packed word each:
 *     mq | bq << 8 | read_position << 16 | in_REF_list << 28 | in_ALT_list << 29
 * values[u * max_per_unit + k], n_values[u] = number of values (may exceed max_per_unit: call
 * again with a larger buffer).  Used for the replicate merge of vafator/annotator.py:285-287.
 * vfk_ranksum_values: s2[p][mq,bq,pos] = twice the ALT mid-rank sum of the pooled lists of pair p
 * (alt[alt_off[p]:alt_off[p+1]] vs ref[ref_off[p]:ref_off[p+1]]), packed words as above -- the
 * integer core of scipy.stats.ranksums(x=alt, y=ref) behind vafator/rank_sum_test.py:6-26.     */
int vfk_pileup_values(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                      const vfk_params *params, int32_t max_per_unit, uint32_t *values, int32_t *n_values, void *stream);
int vfk_ranksum_values(vfk_handle *h, int64_t n_pairs, const uint32_t *alt, const int64_t *alt_off, const uint32_t *ref,
                       const int64_t *ref_off, uint64_t *s2, void *stream);

/* Number of kernel launches issued through this handle since it was created. */
int64_t vfk_launch_count(vfk_handle *h);

/* Host-side helper of vfk_annotate_host (VFK_HOST_GATHER=1, staged): column units {first sector of the window's run,
 * sectors, position, meta} [n][4] + the column store (sector form, or window-transposed when `transposed`) -> per unit a
 * compact run (header pairs, then the entries of the unit's own position; 10 bytes per read, padded to 32) in `out`, and
 * units pointing at those runs.  out == NULL: only *out_bytes is computed.  No CUDA call. */
int vfk_host_gather_columns(const uint32_t *units_in, int64_t n_units, const uint8_t *sectors, int transposed,
                            uint32_t *units_out, uint8_t *out, int64_t out_cap, int64_t *out_bytes);

#ifdef __cplusplus
}
#endif
#endif

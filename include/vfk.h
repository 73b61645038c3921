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
 * What crosses the boundary is the CANONICAL columnar read batch -- the fields of the decoded BAM records
 * the reference hands to htslib at vafator/pileups.py:71-80 (start, flag, MAPQ, CIGAR, 4-bit sequence,
 * qualities, mate position, template length, read name id / hash), untouched.  Everything derived from them
 * -- reference spans, the position index, CIGAR shapes, the read filter, htslib's mate-overlap pairing --
 * is computed by the library's kernels on the device, inside every call.
 *
 * Conventions: every function returns 0 (VFK_OK) or a negative VFK_E* code and never throws or
 * aborts; vfk_last_error() gives the thread-local message.  All buffers are caller owned; the
 * library keeps no reference to them after the stream work it enqueued has completed.  Calls
 * are asynchronous with respect to the host until vfk_sync() (except vfk_annotate_host, which
 * returns after its results are in host memory).  `stream` is a cudaStream_t passed as void*
 * (NULL = the CUDA default stream; the handle's own streams are used by vfk_annotate_host only).
 * A handle belongs to one device and owns per-launch scratch (derived per-read words, located units,
 * deferred-unit lists): drive it from one host thread at a time and keep the calls of one handle on one
 * stream (stream order is what protects the scratch); use one handle per stream / per host thread for
 * concurrency.  Different handles are independent.
 */
#ifndef VFK_H
#define VFK_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFK_VERSION 200 /* 0.2.0 */

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
#define VFK_CODE_NONE 0xFFu

/* ---- canonical columnar read batch --------------------------------------------------------------------
 * The placed reads (tid >= 0) of one BAM, or of one interval shard of it, coordinate sorted (BAM order).
 * Array by array what a BAM record holds (SAMv1 4.2); `contig_off`, `max_l_qseq` and the pool sizes are
 * batch metadata every decoder has at hand when it finishes a batch.
 *   pos        0-based leftmost reference position (>= 0)
 *   cigar      BAM encoding len << 4 | op; read i owns cigar[cigar_off[i] .. + n_cigar[i])
 *   seq        BAM 4-bit codes "=ACMGRSVTWYHKDBN", high nibble first, each read byte aligned:
 *              base q of read i = nibble q of seq[seq_off[i] ...]
 *   qual       one byte per base at qual[qual_off[i] + q]
 *   mtid/mpos/isize   mate contig, mate start, template length (RNEXT / PNEXT / TLEN)
 *   cigar_off / seq_off / qual_off   where read i's CIGAR ops / packed bases / qualities start in their pools.  All three may be
 *              NULL: the pools are then contiguous in read order (what a decoder that appends record by record produces) and
 *              the offsets -- prefix sums of n_cigar, (l_qseq + 1) / 2 and l_qseq -- are derived on the device (the host entry
 *              point then moves 8 bytes per read less and reads no offset array over the bus).
 *   qname_id   any 64-bit id that is equal iff the read names are equal
 *   qname_hash khash X31 of the read name (htslib's overlap tie-break hashes it)
 * vfk_pileup: every pointer is a DEVICE pointer.  vfk_annotate_host: HOST pointers.                    */
typedef struct {
    int64_t n_reads;
    int32_t n_contigs;
    int32_t max_l_qseq;
    const int64_t *contig_off;    /* [n_contigs+1] reads of contig t are [off[t], off[t+1])      */
    const int32_t *pos;           /* [n_reads]                                                   */
    const uint16_t *flag;
    const uint8_t *mapq;
    const int32_t *l_qseq, *n_cigar;
    const int64_t *cigar_off, *seq_off, *qual_off;
    const uint32_t *cigar;        /* [n_cigar_words]                                             */
    const uint8_t *seq, *qual;    /* [n_seq_bytes], [n_qual_bytes]                               */
    const int32_t *mtid, *mpos, *isize;
    const uint64_t *qname_id;
    const uint32_t *qname_hash;
    int64_t n_cigar_words, n_seq_bytes, n_qual_bytes;
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
#define VFK_ST_DEFERRED 0x10u  /* never set in a finished record: the unit is waiting on a device-side list    */
#define VFK_ST_READLEN 0x20u   /* a read position fell outside the histogram: result invalid      */
#define VFK_ST_BADBATCH 0x40u  /* the read batch is malformed (unsorted, negative start, CIGAR outside its pool):
                                  every record of the call carries the bit, results are invalid    */
typedef struct {
    int32_t dp;         /* depth as vafator reports it (pileups.py:224-227, :264, :330)          */
    int32_t n_amb;      /* SNV: IUPAC-ambiguous bases in the column (annotator.py:386)           */
    int32_t ac_ref;     /* SNV: reads showing REF; indel: reads with no indel here (REF support)  */
    int32_t ac_alt;     /* reads supporting this unit's ALT                                       */
    int32_t med2[3][2]; /* [mq,bq,pos][ref,alt]: twice the median, -1 when the list is empty      */
    uint32_t status;
    uint32_t n_cand;    /* reads inspected for the unit (diagnostic): candidate reads of the position index */
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

/* all pointers inside the batches and `out` are DEVICE pointers.  One call = read preparation (reference
 * spans, CIGAR shapes, position index) + locate + pileup kernels, all on `stream`. */
int vfk_pileup(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
               const vfk_params *params, vfk_counts *out, void *stream);

/* counts / eaf / out are DEVICE pointers; eaf[u] = expected VAF of unit u (vafator/power.py:52-82) */
int vfk_epilogue(vfk_handle *h, int64_t n_units, const vfk_counts *counts, const double *eaf,
                 double fpr, double error_rate, vfk_stats *out, void *stream);

/* Same path with HOST buffers, results in host memory before it returns.
 * Page-locked read arrays (cudaHostAlloc / cudaHostRegister, as vafator_b200.bamio and .synth allocate them):
 * what the position index is built from -- pos, n_cigar, cigar_off (unless NULL), l_qseq, the CIGAR pool: 17 to 25 bytes
 * per read -- is copied to the device; the one quality byte and the one sequence byte a column needs of a read reach it
 * through a fetch list (vfk_fetch_stats below); flags, MAPQ, mate fields and name ids are read IN PLACE by the kernels
 * over PCIe, and only for the reads that overlap a variant column.  A deep or dense variant set (sampled on the host), or
 * pageable arrays: everything is staged by asynchronous copies first.  Unit arrays and result buffers may be pageable
 * either way (page-locked bounce buffers inside).  Results are identical.                                            */
int vfk_annotate_host(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                      const vfk_params *params, const double *eaf, double fpr, double error_rate,
                      vfk_counts *counts_out, vfk_stats *stats_out);

/* Pipelined variant: returns once the batch is enqueued on one of the handle's VFK_SLOTS staging slots (each with its own
 * stream and staging memory: the copies of one batch, the host gather of another and the kernels of a third overlap);
 * *ticket (0 .. VFK_SLOTS-1) names the slot.  Submitting one batch more waits for the slot it reuses.
 * The host buffers (inputs and outputs) must stay valid until vfk_wait(h, ticket) returns; vfk_wait
 * returns VFK_EINVAL when the batch turned out to be malformed (VFK_ST_BADBATCH).                    */
int vfk_annotate_host_async(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                            const vfk_params *params, const double *eaf, double fpr, double error_rate,
                            vfk_counts *counts_out, vfk_stats *stats_out, int *ticket);
int vfk_wait(vfk_handle *h, int ticket);
#define VFK_SLOTS 2
int vfk_slot_count(void); /* VFK_SLOTS of the library at hand */
/* transfer accounting of the host entry point since vfk_create: out[0] bytes copied host -> device by
 * bulk copies, out[1] batches whose arrays were read in place (zero copy), out[2] batches staged in full,
 * out[3] bytes copied device -> host                                                               */
int vfk_transfer_stats(vfk_handle *h, int64_t out[4]);
/* fetch lists of the host entry point since vfk_create (batches read in place: the one quality byte and the one sequence
 * byte a column needs of each read are named by the device, gathered by host threads between two kernels of the call and
 * copied in bulk, instead of being read over the bus one request each): out[0] (unit, read) pairs gathered, out[1] batches
 * that carried a list, out[2] nanoseconds the host spent gathering, out[3] host threads of a gather.
 * Environment, read by vfk_create: VFK_HOST_THREADS=n sets the host threads of the gather (default: the OpenMP maximum, at
 * most 16); the lists are used when there are at least six of them (several GPUs sharing few host cores: reading in place
 * is faster), VFK_HOST_FETCH=1 / 0 decides outright.                                                                  */
int vfk_fetch_stats(vfk_handle *h, int64_t out[4]);
/* The host's half of a fetch list on its own (no device involved; what the stream callback of vfk_annotate_host_async runs):
 * req holds two 64-bit words per pair as request_kernel writes them -- word 0 = ~0: nothing wanted (out = 0); with
 * qual_off / seq_off given: word 0 = read index, word 1 = read position; with both NULL: word 0 = offset of the quality byte
 * in `qual`, word 1 = offset of the sequence byte in `seq`.  out[p] = quality | sequence byte << 8; an entry that points
 * outside the arrays yields 0.                                                                                         */
int vfk_fetch_gather(const uint64_t *req, int64_t n_pairs, const uint8_t *qual, int64_t n_qual_bytes, const uint8_t *seq,
                     int64_t n_seq_bytes, const int64_t *qual_off, const int64_t *seq_off, int64_t n_reads, int threads,
                     uint16_t *out);

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

/* The words the read-preparation kernel derives per read, for tests and tools (DEVICE pointers): end[i] = pos + reference bases
 * consumed (htslib bam_endpos); shape[i] = 0 when the CIGAR is one M/=/X op covering the read, 0xFFFFFFFF for a general CIGAR,
 * else leading soft clip | first block length << 12 | event (1 I, 2 D, 3 N) << 24 | event length << 26 for
 * [H*][S] M [ (I|D|N) M ] [S][H*]; *status = VFK_ST_BADBATCH or 0. */
int vfk_prep_words(vfk_handle *h, const vfk_read_batch *reads, int32_t *end_out, uint32_t *shape_out, uint32_t *status_out, void *stream);

/* Per-phase device times of the handle's calls, measured with CUDA events on the launching stream (what bench.py's roofline
 * is computed from).  vfk_set_timing(h, 1), then after a call (and, for the host entry point, its vfk_wait):
 * vfk_last_timing -> ms[0] read preparation (prep + scan kernels), ms[1] locate kernel, ms[2] the register pileup kernel
 * (pileup_warp_kernel), ms[3] list path (link + histogram kernels), ms[4] epilogue kernel (host entry point only, else 0). */
int vfk_set_timing(vfk_handle *h, int on);
int vfk_last_timing(vfk_handle *h, double ms[5]);

/* Number of kernel launches issued through this handle since it was created. */
int64_t vfk_launch_count(vfk_handle *h);

#ifdef __cplusplus
}
#endif
#endif

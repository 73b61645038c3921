/* vfk_io.h -- host-side companion library (libvfkio.so, no CUDA dependency).
 *
 * Produces the canonical columnar read batch include/vfk.h consumes from BGZF/BAM files, and prints INFO text (the
 * synthetic benchmark batches of SURVEY.md 8(d) live in their own library: include/vfk_synth.h).  It stands where
 * pysam.AlignmentFile stands in the reference
 * (vafator/annotator.py:161-164, vafator/pileups.py:71-80).  Nothing here looks inside a CIGAR, filters a read or
 * pairs mates: that is the kernels' work.
 */
#ifndef VFK_IO_H
#define VFK_IO_H
#include <stddef.h>
#include <stdint.h>
#include "vfk.h"

#ifdef __cplusplus
extern "C" {
#endif

/* canonical columnar batch (host memory), coordinate sorted, unplaced reads (tid < 0) last */
typedef struct {
    int64_t n;
    int32_t n_contigs;
    int32_t pad;
    const int32_t *tid, *pos;
    const uint16_t *flag;
    const uint8_t *mapq;
    const int32_t *l_qseq, *n_cigar;
    const int64_t *cigar_off, *seq_off, *qual_off;
    const uint32_t *cigar;
    const uint8_t *seq, *qual;
    const int32_t *mtid, *mpos, *isize;
    const uint64_t *qname_id;
    const uint32_t *qname_hash;
} vfkio_reads;

const char *vfkio_last_error(void);

/* Raw DEFLATE (RFC 1951) of one whole stream into a buffer of exactly the inflated size -- the shape of a BGZF
 * member (SAMv1 4.1), which is what the BAM decoder spends its time on.  Returns 0 when the stream ends after
 * exactly out_n bytes; non-zero on anything else (malformed, truncated, or merely unusual input) without writing
 * outside [out, out + out_n) -- the BAM decoder then asks zlib for the authoritative answer. */
int vfkio_inflate_raw(const uint8_t *in, size_t in_n, uint8_t *out, size_t out_n);

/* ---- BGZF / BAM decoding (coordinate-sorted BAM; CRAM is reported as unsupported) -----------
 * vfkio_bam_open inflates the file (blocks in parallel) and indexes its placed records;
 * vfkio_bam_fill writes them, in file order, into caller-allocated canonical arrays sized by
 * vfkio_bam_plan.  Unplaced reads (tid < 0) are skipped: they can never reach a pileup column. */
typedef struct vfkio_bam vfkio_bam;
typedef struct {
    int64_t n_reads, n_cigar, n_seq_bytes, n_qual_bytes;
} vfkio_synth_plan; /* also the size plan of the synthetic generator (include/vfk_synth.h) */
int vfkio_bam_open(const char *path, vfkio_bam **out);
/* Indexed access (path.bai): only the records overlapping the regions [beg, end) (0-based, half open; sorted
 * by contig order of the BAM and start, disjoint) are decoded -- what the reference fetches per chromosome
 * group through pysam (vafator/pileups.py:71-80, :137-138).  n_regions == 0 reads the header only.           */
int vfkio_bam_open_regions(const char *path, const char *index_path, int32_t n_regions, const char *const *contig,
                           const int32_t *beg, const int32_t *end, vfkio_bam **out);
void vfkio_bam_close(vfkio_bam *b);
int32_t vfkio_bam_n_contigs(const vfkio_bam *b);
const char *vfkio_bam_contig_name(const vfkio_bam *b, int32_t i);
int64_t vfkio_bam_contig_length(const vfkio_bam *b, int32_t i);
const char *vfkio_bam_header_text(const vfkio_bam *b);
int32_t vfkio_bam_is_sorted(const vfkio_bam *b);
int vfkio_bam_plan(const vfkio_bam *b, vfkio_synth_plan *plan);
int vfkio_bam_fill(const vfkio_bam *b, vfkio_reads *out);

/* ---- INFO text of one single-BAM sample (vfkio_format.cpp) --------------------------------------------------
 * The strings vafator/annotator.py:304-420 builds per record ("{sample}_af=..;{sample}_ac=..;...;{sample}_pos=..", then the
 * rank-sum pairs that exist), from the arrays vafator_b200/engine.py holds: per record i (rows first[i] .. first[i+1], one per
 * ALT): kind (0 SNV, 1 INS, 2 DEL, -1 none), has (a pileup column exists), dp, n, ac_ref, med2_ref[i][mq,bq,pos] (twice the
 * median), eaf; per row: alt_idx, present (ALT text is upper case), ac, med2_alt, pu / pw / z[3] / p[3] (already rounded), k.
 * Floats are printed as Python's repr(float) prints them, af = round(ac / dp, 5) as Python rounds.  *text_out (malloc'ed:
 * vfkio_free) holds the records back to back, record i = [offsets_out[i], offsets_out[i+1]) (n_rec + 1 entries). */
typedef struct {
    const char *sample;
    int64_t n_rec;
    const int64_t *first;
    const int64_t *kind_rec;
    const uint8_t *has_rec;
    const int64_t *dp, *n, *ac_ref, *med2_ref;
    const double *eaf;
    const int64_t *alt_idx;
    const uint8_t *present_row;
    const int64_t *ac, *med2_alt;
    const double *pu, *pw;
    const int64_t *k;
    const double *z, *p;
} vfkio_format_in;
int vfkio_format_info(const vfkio_format_in *in, char **text_out, int64_t *offsets_out);
void vfkio_free(void *p);
/* repr(float) as CPython prints it, into out32 (>= 32 bytes, not NUL terminated); returns the length */
int vfkio_py_repr(double x, char *out32);

#ifdef __cplusplus
}
#endif
#endif

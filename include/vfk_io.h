/* vfk_io.h -- host-side companion library (libvfkio.so, no CUDA dependency).
 *
 * Produces the canonical columnar read batch include/vfk.h consumes: decodes BGZF/BAM files and generates the
 * synthetic benchmark batches of SURVEY.md 8(d).  It stands where pysam.AlignmentFile stands in the reference
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
} vfkio_synth_plan; /* also the size plan of the synthetic generator below */
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

/* ---- synthetic read batches: seeded, reproducible, generated tile by tile ----
 * A contig is covered by `tiles_per_contig` tiles starting every `tile_stride` bases from
 * `tile_offset`; the first mate of each of the `pairs_per_tile` pairs of a tile starts inside
 * [tile, tile+tile_len).  tile_stride == tile_len gives contiguous (WGS-like) coverage, a larger
 * stride gives target-like (WES / amplicon) islands.  The reference genome, the planted variant
 * sites and every read are pure functions of (seed, coordinates / pair index).                   */
typedef struct {
    uint64_t seed;
    int32_t n_contigs;
    int32_t read_len;
    int32_t tile_offset, tile_len, tile_stride, tiles_per_contig, pairs_per_tile;
    int32_t var_lo, var_hi;     /* planted sites live in [tile+var_lo, tile+var_hi) */
    int32_t snv_period;         /* one planted SNV site per ~snv_period bases of that window, 0 = none */
    int32_t indel_period;       /* same for planted insertions / deletions */
    int32_t multiallelic_pct;   /* % of SNV sites that get a second ALT */
    int32_t normal_vaf;         /* 1: VAF in {0.5, 1.0} (normal sample), 0: tumour-like spectrum */
    float frag_mean, frag_sd;
    float p_softclip, p_del, p_ins, p_skip;            /* private CIGAR noise per read */
    float p_dup, p_supp, p_secondary, p_qcfail, p_orphan;
    float p_mapq60, base_error, p_n;
    uint64_t read_salt;         /* mixed into the seed of every read pair, not into the genome / planted sites: samples that
                                   share their variant sites but not their reads (0 = the plain seed) */
    int64_t g_first, g_count;   /* generate only the tiles [g_first, g_first + g_count) of the genome, global tile index =
                                   contig * tiles_per_contig + tile (g_count 0 = all): an interval shard of the same input */
} vfkio_synth_cfg;

/* tile_sizes: [4 * generated tiles] scratch filled by the plan pass and handed to the fill pass */
int vfkio_synth_plan_reads(const vfkio_synth_cfg *cfg, vfkio_synth_plan *plan, int64_t *tile_sizes);
/* fills caller-allocated canonical arrays (sizes from the plan); `out` points at writable arrays */
int vfkio_synth_fill_reads(const vfkio_synth_cfg *cfg, const vfkio_synth_plan *plan, const int64_t *tile_sizes,
                           vfkio_reads *out);

/* planted variant sites in genome order.  kind 0 SNV / 1 INS / 2 DEL; `bases`: per site 24 bytes of
 * text: [0] REF base, [1] ALT base (SNV), [2] second ALT or 0, [3..] inserted bases (INS, NUL padded);
 * `length`: inserted / deleted length.  Returns the number of sites (<= cap) or a negative error. */
int64_t vfkio_synth_sites(const vfkio_synth_cfg *cfg, int64_t cap, int32_t *tid, int32_t *pos1, int32_t *kind,
                          int32_t *length, char *bases);
/* reference bases (ASCII) of [start, start+len) */
int vfkio_synth_ref(const vfkio_synth_cfg *cfg, int32_t tid, int32_t start, int32_t len, char *out);

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

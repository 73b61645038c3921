/* vfk_io.h -- host-side companion library (libvfkio.so, no CUDA dependency).
 *
 * Produces what include/vfk.h consumes: turns decoded BAM records (the canonical columnar
 * batch, vafator_b200/batch.py) into the HBM layout, emulates htslib's mate-overlap pairing,
 * decodes BGZF/BAM files and generates the synthetic benchmark batches of SURVEY.md 8(d).
 * It stands where pysam.AlignmentFile stands in the reference (vafator/annotator.py:161-164,
 * vafator/pileups.py:71-80).
 */
#ifndef VFK_IO_H
#define VFK_IO_H
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

typedef struct {
    int64_t n_placed;      /* reads with tid >= 0 (the packed batch holds exactly these) */
    int64_t n_cigar_words; /* CIGAR words of the shape-1 reads */
    int64_t n_sectors;     /* 32-byte payload sectors */
    int32_t max_l_qseq;
    int32_t pad;
} vfkio_plan;

const char *vfkio_last_error(void);

/* sizes of the packed arrays */
int vfkio_pack_plan(const vfkio_reads *r, vfkio_plan *plan);

/* fill caller-allocated arrays sized by the plan (host memory, pinned or not) */
int vfkio_pack(const vfkio_reads *r, const vfkio_plan *plan, int64_t *contig_off /*[n_contigs+1]*/,
               vfk_read_hot *hot, vfk_read_cold *cold, int32_t *starts, int32_t *pmax_end, uint32_t *cigar,
               uint8_t *payload);

/* htslib overlap_push / overlap_remove emulation (sam.c): mate[i] = partner |
 * tie_break << 31 or VFK_NO_MATE, for the reads the reference would push under `params`. */
int vfkio_link_mates(const vfkio_reads *r, const vfk_params *params, uint32_t *mate /*[n_placed]*/);

#ifdef __cplusplus
}
#endif
#endif

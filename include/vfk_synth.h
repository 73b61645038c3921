/* vfk_synth.h -- the synthetic benchmark / test data of SURVEY.md 8(d) (libvfksynth.so, vafator_b200/csrc/vfkio_synth.cpp).
 * Not part of the product path: bench.py (both arms), the tests and the tools generate their inputs with it, in its own
 * library, so that a process that only generates data and runs the CPU oracle (bench.py --impl reference) loads no product
 * code.  Fills the same canonical arrays a BAM decoder fills (include/vfk_io.h vfkio_reads).                            */
#ifndef VFK_SYNTH_H
#define VFK_SYNTH_H
#include "vfk_io.h"
#ifdef __cplusplus
extern "C" {
#endif

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

#ifdef __cplusplus
}
#endif
#endif

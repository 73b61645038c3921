// vfkio_synth.cpp -- seeded synthetic read batches and planted variant sites.
//
// Everything is a pure function of (seed, coordinates or pair index), so tiles are generated
// independently and in parallel, the plan pass and the fill pass see the same reads, and the
// same data can be regenerated on any box.  Reads carry their own bases (the reference path
// uses no FASTA: vafator/pileups.py:71-80 passes no `fastafile`).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "vfk_synth.h"

int vfkio_fail(int code, const char *msg);

namespace {

inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint64_t hash3(uint64_t a, uint64_t b, uint64_t c) { return mix64(mix64(mix64(a) ^ b) ^ c); }

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { return mix64(s += 0x9E3779B97F4A7C15ull); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int)(next() % (uint64_t)n); }
    double gauss() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

const char ACGT[] = "ACGT";
const uint8_t CODE_OF_ACGT[4] = {1, 2, 4, 8};  // BAM 4-bit codes of A C G T
const uint8_t CODE_N = 15;
const uint8_t IUPAC_RARE[4] = {5, 10, 12, 3};  // R Y K M

// reference genome: 25 % of the 8-base blocks are homopolymers so that tandem-duplication
// insertions exist (the reference only ever matches those: SURVEY.md A.5)
inline int ref_base(const vfkio_synth_cfg *c, int32_t tid, int32_t pos) {
    const uint64_t hb = hash3(c->seed ^ 0x5EEDBA5Eull, (uint64_t)tid, (uint64_t)(pos >> 3));
    if ((hb & 3) == 0) return (int)((hb >> 8) & 3);
    return (int)(hash3(c->seed ^ 0xBA5Eull, (uint64_t)tid, (uint64_t)pos) & 3);
}

struct Site {
    int kind;  // 0 SNV, 1 INS, 2 DEL
    int len;
    float vaf;
    int alt, alt2;  // SNV: base indexes, alt2 = -1 if none
    uint8_t ins[8];  // base indexes
};

inline bool in_window(const vfkio_synth_cfg *c, int32_t pos) {
    const int64_t rel = (int64_t)pos - c->tile_offset;
    if (rel < 0) return false;
    const int64_t k = rel / c->tile_stride;
    const int64_t off = rel - k * c->tile_stride;
    if (k < c->tiles_per_contig && off >= c->var_lo && off < c->var_hi) return true;
    // a window may reach into the next stride
    if (k >= 1 && k - 1 < c->tiles_per_contig && off + c->tile_stride >= c->var_lo && off + c->tile_stride < c->var_hi) return true;
    return false;
}

inline bool site_at(const vfkio_synth_cfg *c, int32_t tid, int32_t pos, Site &s) {
    if (pos < 1 || !in_window(c, pos)) return false;
    const uint64_t h = hash3(c->seed ^ 0x517Eull, (uint64_t)tid, (uint64_t)pos);
    const bool snv = c->snv_period > 0 && (h % (uint64_t)c->snv_period) == 0;
    const bool indel = !snv && c->indel_period > 0 && ((h >> 24) % (uint64_t)c->indel_period) == 0;
    if (!snv && !indel) return false;
    const uint64_t g = mix64(h);
    static const float TUMOR[8] = {0.03f, 0.05f, 0.1f, 0.15f, 0.25f, 0.35f, 0.5f, 1.0f};
    s.vaf = c->normal_vaf ? ((g & 1) ? 0.5f : 1.0f) : TUMOR[g & 7];
    const int rb = ref_base(c, tid, pos);
    if (snv) {
        s.kind = 0;
        s.len = 0;
        s.alt = (rb + 1 + (int)((g >> 8) % 3)) & 3;
        s.alt2 = -1;
        if ((int)((g >> 16) % 100) < c->multiallelic_pct) {
            for (int d = 1; d <= 3; ++d) {
                const int cand = (rb + d) & 3;
                if (cand != s.alt) { s.alt2 = cand; break; }
            }
        }
        return true;
    }
    if ((g >> 3) & 1) {
        s.kind = 1;
        s.len = 1 + (int)((g >> 8) % 8);
        const bool tandem = ((g >> 4) & 1) != 0;  // copy of the bases that follow: matches the reference's shifted compare
        for (int t = 0; t < s.len; ++t)
            s.ins[t] = (uint8_t)(tandem ? ref_base(c, tid, pos + 1 + t) : (int)((g >> (20 + 2 * t)) & 3));
    } else {
        s.kind = 2;
        s.len = 1 + (int)((g >> 8) % 10);
    }
    s.alt = s.alt2 = -1;
    return true;
}

// Per-tile cache of the reference bases and planted sites the tile's reads can touch, so that the
// per-base work of make_read is a table lookup instead of three hashes.
struct TileCtx {
    int32_t lo = 0, hi = 0;
    std::vector<uint8_t> ref;
    std::vector<int16_t> site_idx;
    std::vector<Site> sites;
    void build(const vfkio_synth_cfg *c, int32_t tid, int32_t from, int32_t to) {
        lo = from; hi = to;
        const int n = to - from;
        ref.resize(n);
        site_idx.assign(n, -1);
        sites.clear();
        for (int i = 0; i < n; ++i) {
            ref[i] = (uint8_t)ref_base(c, tid, from + i);
            Site s;
            if (sites.size() < 32000 && site_at(c, tid, from + i, s)) { site_idx[i] = (int16_t)sites.size(); sites.push_back(s); }
        }
    }
    inline int base(const vfkio_synth_cfg *c, int32_t tid, int32_t pos) const {
        return (pos >= lo && pos < hi) ? ref[pos - lo] : ref_base(c, tid, pos);
    }
    inline bool site(const vfkio_synth_cfg *c, int32_t tid, int32_t pos, Site &s) const {
        if (pos >= lo && pos < hi) {
            const int16_t k = site_idx[pos - lo];
            if (k < 0) return false;
            s = sites[k];
            return true;
        }
        return site_at(c, tid, pos, s);
    }
};

struct ReadOut {
    int32_t pos, mpos, isize;
    uint16_t flag;
    uint8_t mapq;
    uint64_t qid;
    uint32_t tie;  // sort tie-break
    std::vector<uint32_t> cigar;
    std::vector<uint8_t> codes, qual;
};

inline void push_op(std::vector<uint32_t> &cg, uint32_t op, uint32_t len) {
    if (len == 0) return;
    if (!cg.empty() && (cg.back() & 15u) == op) cg.back() += len << 4;
    else cg.push_back((len << 4) | op);
}

void make_read(const vfkio_synth_cfg *c, const TileCtx &ctx, Rng &rng, int32_t tid, int32_t start, ReadOut &o) {
    const int L = c->read_len;
    o.cigar.clear();
    o.codes.clear();
    o.qual.clear();
    o.pos = start;
    int lead = 0, trail = 0;
    int priv_kind = -1, priv_at = -1, priv_len = 0;  // private CIGAR noise: 1 I, 2 D, 3 N
    double u = rng.uni();
    if (u < c->p_softclip) {
        const int n = 1 + rng.below(30);
        const int where = rng.below(3);
        if (where != 1) lead = n;
        if (where != 0) trail = std::min(n, 20);
    } else if ((u -= c->p_softclip) < c->p_del) {
        priv_kind = 2; priv_len = 1 + rng.below(10);
    } else if ((u -= c->p_del) < c->p_ins) {
        priv_kind = 1; priv_len = 1 + rng.below(8);
    } else if ((u -= c->p_ins) < c->p_skip) {
        priv_kind = 3; priv_len = 100 + rng.below(4900);
    }
    const int body = L - lead - trail;
    if (priv_kind > 0) priv_at = 10 + rng.below(std::max(1, body - 20));
    for (int t = 0; t < lead; ++t) o.codes.push_back(CODE_OF_ACGT[rng.below(4)]);
    push_op(o.cigar, 4, (uint32_t)lead);
    int32_t r = start;
    int used = 0;
    while (used < body) {
        Site s;
        const bool has = ctx.site(c, tid, r, s);
        const bool carry = has && rng.uni() < s.vaf;
        int b = ctx.base(c, tid, r);
        if (carry && s.kind == 0) b = (s.alt2 >= 0 && rng.uni() < 0.3) ? s.alt2 : s.alt;
        uint8_t code = CODE_OF_ACGT[b];
        const double e = rng.uni();
        if (e < c->base_error) code = CODE_OF_ACGT[rng.below(4)];
        else if (e < c->base_error + c->p_n) code = (rng.below(8) == 0) ? IUPAC_RARE[rng.below(4)] : CODE_N;
        o.codes.push_back(code);
        push_op(o.cigar, 0, 1);
        ++used;
        ++r;
        if (used >= body) break;
        const int left = body - used;
        if (carry && s.kind == 1 && left > s.len + 4 && used > 4) {
            for (int t = 0; t < s.len; ++t) o.codes.push_back(CODE_OF_ACGT[s.ins[t]]);
            push_op(o.cigar, 1, (uint32_t)s.len);
            used += s.len;
        } else if (carry && s.kind == 2 && left > 4 && used > 4) {
            push_op(o.cigar, 2, (uint32_t)s.len);
            r += s.len;
        } else if (used == priv_at && left > priv_len + 4) {
            if (priv_kind == 1) {
                const bool tandem = rng.below(2) == 0;
                for (int t = 0; t < priv_len; ++t) o.codes.push_back(CODE_OF_ACGT[tandem ? ctx.base(c, tid, r + t) : rng.below(4)]);
                push_op(o.cigar, 1, (uint32_t)priv_len);
                used += priv_len;
            } else {
                push_op(o.cigar, priv_kind == 2 ? 2u : 3u, (uint32_t)priv_len);
                r += priv_len;
            }
        }
    }
    for (int t = 0; t < trail; ++t) o.codes.push_back(CODE_OF_ACGT[rng.below(4)]);
    push_op(o.cigar, 4, (uint32_t)trail);
    // binned Illumina-like qualities with 3' decay
    const int n = (int)o.codes.size();
    o.qual.resize(n);
    for (int t = 0; t < n; ++t) {
        const double frac = (double)t / (double)n;
        const double v = rng.uni();
        uint8_t q;
        if (v < 0.55 - 0.25 * frac) q = 41;
        else if (v < 0.80 - 0.30 * frac) q = 37;
        else if (v < 0.92 - 0.20 * frac) q = 30;
        else if (v < 0.96 - 0.10 * frac) q = 23;
        else if (v < 0.985) q = 12;
        else q = 2;
        o.qual[t] = q;
    }
}

inline int32_t tile_start(const vfkio_synth_cfg *c, int64_t k) { return (int32_t)(c->tile_offset + k * c->tile_stride); }

// all pairs whose first mate starts in tile k; appends the reads whose own start lies in [own_lo, own_hi)
void gen_pairs(const vfkio_synth_cfg *c, int32_t tid, int64_t k, int32_t own_lo, int32_t own_hi, std::vector<ReadOut> &out,
               TileCtx &ctx) {
    static const uint8_t LOWMQ[6] = {0, 1, 10, 20, 29, 40};
    const int32_t T = tile_start(c, k);
    ctx.build(c, tid, T, T + c->tile_len + 700 + c->read_len + 64);
    for (int j = 0; j < c->pairs_per_tile; ++j) {
        const uint64_t pair_id = ((uint64_t)tid * (uint64_t)c->tiles_per_contig + (uint64_t)k) * (uint64_t)c->pairs_per_tile + (uint64_t)j + 1;
        Rng rng(hash3(c->seed ^ c->read_salt, pair_id, 0x9A12ull));
        int frag = (int)std::lround(c->frag_mean + c->frag_sd * rng.gauss());
        frag = std::max(frag, c->read_len);
        frag = std::min(frag, c->read_len + 700);
        const int32_t s1 = T + rng.below(c->tile_len);
        const int32_t s2 = s1 + frag - c->read_len;
        const bool fwd_first = rng.below(2) == 0;
        uint16_t f1 = fwd_first ? 99 : 163, f2 = fwd_first ? 147 : 83;
        bool emit2 = true, mate_unmapped = false;
        double u = rng.uni();
        if (u < c->p_dup) { f1 |= 0x400; f2 |= 0x400; }
        else if ((u -= c->p_dup) < c->p_supp) { if (rng.below(2)) f1 |= 0x800; else f2 |= 0x800; }
        else if ((u -= c->p_supp) < c->p_secondary) { if (rng.below(2)) f1 |= 0x100; else f2 |= 0x100; }
        else if ((u -= c->p_secondary) < c->p_qcfail) { if (rng.below(2)) f1 |= 0x200; else f2 |= 0x200; }
        else if ((u -= c->p_qcfail) < c->p_orphan) {
            if (rng.below(2)) { f1 = 65 | (fwd_first ? 32 : 16); f2 = 129 | (fwd_first ? 16 : 32); }
            else { f1 = 73; mate_unmapped = true; }
        }
        const uint8_t mq1 = rng.uni() < c->p_mapq60 ? 60 : LOWMQ[rng.below(6)];
        const uint8_t mq2 = rng.uni() < c->p_mapq60 ? 60 : LOWMQ[rng.below(6)];
        Rng r1(hash3(c->seed ^ c->read_salt, pair_id, 1)), r2(hash3(c->seed ^ c->read_salt, pair_id, 2));
        if (s1 >= own_lo && s1 < own_hi) {
            out.emplace_back();
            ReadOut &o = out.back();
            make_read(c, ctx, r1, tid, s1, o);
            o.flag = f1; o.mapq = mq1; o.qid = pair_id; o.tie = (uint32_t)(2 * (uint64_t)j);
            o.mpos = mate_unmapped ? s1 : s2;
            o.isize = mate_unmapped ? 0 : frag;
        }
        if (mate_unmapped) {
            // the unmapped mate is stored at its partner's coordinate (SAM convention), no CIGAR
            if (s1 >= own_lo && s1 < own_hi) {
                out.emplace_back();
                ReadOut &o = out.back();
                o.pos = s1; o.mpos = s1; o.isize = 0; o.flag = 133; o.mapq = 0; o.qid = pair_id; o.tie = (uint32_t)(2 * (uint64_t)j + 1);
                o.codes.assign((size_t)c->read_len, CODE_N);
                o.qual.assign((size_t)c->read_len, 2);
            }
            emit2 = false;
        }
        if (emit2 && s2 >= own_lo && s2 < own_hi) {
            out.emplace_back();
            ReadOut &o = out.back();
            make_read(c, ctx, r2, tid, s2, o);
            o.flag = f2; o.mapq = mq2; o.qid = pair_id; o.tie = (uint32_t)(2 * (uint64_t)j + 1);
            o.mpos = s1;
            o.isize = -frag;
        }
    }
}

// reads owned by tile k: start in [tile_start(k), tile_start(k+1)) -- second mates may come from tile k-1
void gen_tile(const vfkio_synth_cfg *c, int32_t tid, int64_t k, std::vector<ReadOut> &out, TileCtx &ctx) {
    out.clear();
    const int32_t lo = tile_start(c, k);
    const int32_t hi = (k + 1 < c->tiles_per_contig) ? tile_start(c, k + 1) : INT32_MAX;
    gen_pairs(c, tid, k, lo, hi, out, ctx);
    // second mates start at most 700 bases after the first: only then can tile k-1 spill into k
    if (k > 0 && c->tile_stride - c->tile_len < 704) gen_pairs(c, tid, k - 1, lo, hi, out, ctx);
    std::sort(out.begin(), out.end(), [](const ReadOut &a, const ReadOut &b) {
        if (a.pos != b.pos) return a.pos < b.pos;
        if (a.qid != b.qid) return a.qid < b.qid;
        return a.tie < b.tie;
    });
}

int check_cfg(const vfkio_synth_cfg *c) {
    if (!c) return vfkio_fail(VFK_EINVAL, "synth: NULL config");
    if (c->n_contigs < 1 || c->read_len < 30 || c->read_len > 4000 || c->tile_len < 1 || c->tile_stride < c->tile_len ||
        c->tiles_per_contig < 1 || c->pairs_per_tile < 0)
        return vfkio_fail(VFK_EINVAL, "synth: bad geometry");
    // a second mate may start at most read_len+700-read_len = 700 bases after the first: it must
    // fall in the same or the next stride, never further
    if (c->tile_stride < 1024) return vfkio_fail(VFK_EINVAL, "synth: tile_stride must be >= 1024");
    if ((int64_t)c->tile_offset + (int64_t)c->tile_stride * c->tiles_per_contig > (int64_t)INT32_MAX - 8192)
        return vfkio_fail(VFK_EINVAL, "synth: contig exceeds int32 coordinates");
    if (c->g_first < 0 || c->g_count < 0 || c->g_first + c->g_count > (int64_t)c->n_contigs * c->tiles_per_contig)
        return vfkio_fail(VFK_EINVAL, "synth: generated tile range outside the genome");
    return VFK_OK;
}
// the tiles this call generates: [first, first + count) of the genome's global tile index
inline int64_t gen_first(const vfkio_synth_cfg *c) { return c->g_count > 0 ? c->g_first : 0; }
inline int64_t gen_count(const vfkio_synth_cfg *c) { return c->g_count > 0 ? c->g_count : (int64_t)c->n_contigs * c->tiles_per_contig; }

}  // namespace

extern "C" int vfkio_synth_plan_reads(const vfkio_synth_cfg *c, vfkio_synth_plan *plan, int64_t *tile_sizes) {
    int rc = check_cfg(c);
    if (rc) return rc;
    if (!plan) return vfkio_fail(VFK_EINVAL, "synth: NULL plan");
    const int64_t tiles = gen_count(c), g0 = gen_first(c);
    int64_t nr = 0, nc = 0, ns = 0, nq = 0;
#pragma omp parallel
    {
        std::vector<ReadOut> buf;
        TileCtx ctx;
#pragma omp for schedule(dynamic, 64) reduction(+ : nr, nc, ns, nq)
        for (int64_t g = 0; g < tiles; ++g) {
            gen_tile(c, (int32_t)((g0 + g) / c->tiles_per_contig), (g0 + g) % c->tiles_per_contig, buf, ctx);
            int64_t tcg = 0, tsq = 0, tql = 0;
            for (const ReadOut &o : buf) {
                tcg += (int64_t)o.cigar.size();
                tsq += ((int64_t)o.codes.size() + 1) / 2;
                tql += (int64_t)o.qual.size();
            }
            nr += (int64_t)buf.size(); nc += tcg; ns += tsq; nq += tql;
            if (tile_sizes) { tile_sizes[4 * g] = (int64_t)buf.size(); tile_sizes[4 * g + 1] = tcg; tile_sizes[4 * g + 2] = tsq; tile_sizes[4 * g + 3] = tql; }
        }
    }
    plan->n_reads = nr;
    plan->n_cigar = nc;
    plan->n_seq_bytes = ns;
    plan->n_qual_bytes = nq;
    return VFK_OK;
}

extern "C" int vfkio_synth_fill_reads(const vfkio_synth_cfg *c, const vfkio_synth_plan *plan, const int64_t *tile_sizes,
                                      vfkio_reads *out) {
    int rc = check_cfg(c);
    if (rc) return rc;
    if (!plan || !out || !tile_sizes) return vfkio_fail(VFK_EINVAL, "synth: NULL argument");
    const int64_t tiles = gen_count(c), g0 = gen_first(c);
    // per-tile sizes (from the plan pass) -> offsets
    std::vector<int64_t> tr(tiles + 1, 0), tc(tiles + 1, 0), ts(tiles + 1, 0), tq(tiles + 1, 0);
    for (int64_t g = 0; g < tiles; ++g) {
        tr[g + 1] = tr[g] + tile_sizes[4 * g]; tc[g + 1] = tc[g] + tile_sizes[4 * g + 1];
        ts[g + 1] = ts[g] + tile_sizes[4 * g + 2]; tq[g + 1] = tq[g] + tile_sizes[4 * g + 3];
    }
    if (tr[tiles] != plan->n_reads || tc[tiles] != plan->n_cigar || ts[tiles] != plan->n_seq_bytes || tq[tiles] != plan->n_qual_bytes)
        return vfkio_fail(VFK_EINVAL, "synth: plan does not match the config");
    int32_t *tid = (int32_t *)out->tid, *pos = (int32_t *)out->pos, *lq = (int32_t *)out->l_qseq, *ncg = (int32_t *)out->n_cigar;
    int32_t *mtid = (int32_t *)out->mtid, *mpos = (int32_t *)out->mpos, *isz = (int32_t *)out->isize;
    uint16_t *flag = (uint16_t *)out->flag;
    uint8_t *mapq = (uint8_t *)out->mapq, *seq = (uint8_t *)out->seq, *qual = (uint8_t *)out->qual;
    int64_t *co = (int64_t *)out->cigar_off, *so = (int64_t *)out->seq_off, *qo = (int64_t *)out->qual_off;
    uint32_t *cig = (uint32_t *)out->cigar, *qh = (uint32_t *)out->qname_hash;
    uint64_t *qid = (uint64_t *)out->qname_id;
#pragma omp parallel
    {
        std::vector<ReadOut> buf;
        TileCtx ctx;
#pragma omp for schedule(dynamic, 64)
        for (int64_t g = 0; g < tiles; ++g) {
            const int32_t t = (int32_t)((g0 + g) / c->tiles_per_contig);
            gen_tile(c, t, (g0 + g) % c->tiles_per_contig, buf, ctx);
            if ((int64_t)buf.size() != tile_sizes[4 * g]) continue;  // inconsistent plan: caught by the caller's checks
            int64_t i = tr[g], ic = tc[g], is = ts[g], iq = tq[g];
            for (const ReadOut &o : buf) {
                tid[i] = t; pos[i] = o.pos; flag[i] = o.flag; mapq[i] = o.mapq;
                lq[i] = (int32_t)o.codes.size(); ncg[i] = (int32_t)o.cigar.size();
                co[i] = ic; so[i] = is; qo[i] = iq;
                mtid[i] = t; mpos[i] = o.mpos; isz[i] = o.isize;
                qid[i] = o.qid; qh[i] = (uint32_t)mix64(o.qid ^ 0xC0FFEEull);
                if (!o.cigar.empty()) memcpy(cig + ic, o.cigar.data(), o.cigar.size() * 4);
                ic += (int64_t)o.cigar.size();
                const size_t n = o.codes.size();
                for (size_t b = 0; b < n; b += 2) seq[is++] = (uint8_t)((o.codes[b] << 4) | (b + 1 < n ? o.codes[b + 1] : 0));
                if (n) memcpy(qual + iq, o.qual.data(), n);
                iq += (int64_t)n;
                ++i;
            }
        }
    }
    out->n = plan->n_reads;
    out->n_contigs = c->n_contigs;
    return VFK_OK;
}

extern "C" int64_t vfkio_synth_sites(const vfkio_synth_cfg *c, int64_t cap, int32_t *tid, int32_t *pos1, int32_t *kind,
                                     int32_t *length, char *bases) {
    int rc = check_cfg(c);
    if (rc) return rc;
    int64_t n = 0;
    for (int32_t t = 0; t < c->n_contigs; ++t) {
        int32_t prev_end = -1;
        for (int64_t k = 0; k < c->tiles_per_contig; ++k) {
            const int32_t T = tile_start(c, k);
            int32_t a = std::max<int32_t>(T + c->var_lo, prev_end), b = T + c->var_hi;
            for (int32_t p = a; p < b; ++p) {
                Site s;
                if (!site_at(c, t, p, s)) continue;
                if (n < cap) {
                    tid[n] = t; pos1[n] = p + 1; kind[n] = s.kind; length[n] = s.len;
                    char *o = bases + n * 24;
                    memset(o, 0, 24);
                    o[0] = ACGT[ref_base(c, t, p)];
                    if (s.kind == 0) { o[1] = ACGT[s.alt]; o[2] = s.alt2 >= 0 ? ACGT[s.alt2] : 0; }
                    if (s.kind == 1) for (int j = 0; j < s.len; ++j) o[3 + j] = ACGT[s.ins[j]];
                }
                ++n;
            }
            prev_end = std::max(prev_end, b);
        }
    }
    return n;
}

extern "C" int vfkio_synth_ref(const vfkio_synth_cfg *c, int32_t tid, int32_t start, int32_t len, char *out) {
    if (!c || !out || len < 0) return vfkio_fail(VFK_EINVAL, "synth_ref: bad argument");
    for (int32_t i = 0; i < len; ++i) out[i] = ACGT[ref_base(c, tid, start + i)];
    return VFK_OK;
}

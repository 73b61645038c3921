// vfkio_pack.cpp -- canonical columnar batch -> HBM layout, and mate linking (host side).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>
#include "vfk_io.h"

static thread_local char g_err[512] = "";
extern "C" const char *vfkio_last_error(void) { return g_err; }
int vfkio_fail(int code, const char *msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

namespace {
inline bool consumes_ref(unsigned op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
inline bool is_match(unsigned op) { return op == 0 || op == 7 || op == 8; }

struct Shape {
    uint32_t span;
    bool simple;     // one M/=/X op spanning the whole read
    uint32_t event;  // non-zero flag word of a one-event read (vfk.h VFK_SHAPE_ONE_EVENT), else 0
    bool one_event;
};
// [H*][S] M [ (I|D|N) M ] [S][H*] with every length in range -> the packed event word
inline bool one_event_word(const uint32_t *c, int n, int32_t l_qseq, uint32_t *word) {
    int k = 0, e = n;
    while (k < e && (c[k] & 15u) == 5u) ++k;      // leading H
    while (e > k && (c[e - 1] & 15u) == 5u) --e;  // trailing H
    uint32_t lead = 0;
    if (k < e && (c[k] & 15u) == 4u) lead = c[k++] >> 4;
    if (e > k && (c[e - 1] & 15u) == 4u) --e;  // trailing S only lengthens the query
    const int ops = e - k;
    if (ops != 1 && ops != 3) return false;
    if (!is_match(c[k] & 15u)) return false;
    const uint32_t m1 = c[k] >> 4;
    uint32_t ev = VFK_EV_NONE, len = 0;
    if (ops == 3) {
        const uint32_t op = c[k + 1] & 15u;
        ev = op == 1u ? VFK_EV_INS : op == 2u ? VFK_EV_DEL : op == 3u ? VFK_EV_SKIP : VFK_EV_NONE;
        len = c[k + 1] >> 4;
        if (ev == VFK_EV_NONE || !is_match(c[k + 2] & 15u) || (c[k + 2] >> 4) == 0u || len == 0u) return false;
    }
    if (m1 == 0u || m1 >= 4096u || lead >= 4096u || len >= 64u || l_qseq > VFK_MAX_READ_LEN) return false;
    *word = lead | (m1 << 12) | (ev << 24) | (len << 26);
    return true;
}
inline Shape shape_of(const vfkio_reads *r, int64_t i) {
    const uint32_t *c = r->cigar + r->cigar_off[i];
    const int n = r->n_cigar[i];
    uint32_t span = 0;
    for (int k = 0; k < n; ++k)
        if (consumes_ref(c[k] & 15u)) span += c[k] >> 4;
    Shape sh;
    sh.span = span;
    sh.simple = n == 1 && is_match(c[0] & 15u) && (int64_t)(c[0] >> 4) == (int64_t)r->l_qseq[i];
    sh.event = 0;
    sh.one_event = !sh.simple && one_event_word(c, n, r->l_qseq[i], &sh.event);
    return sh;
}
inline int64_t sectors_of(int32_t l_qseq) { return ((int64_t)l_qseq + 19) / 20; }  // 4 granules x 5 positions
}  // namespace

extern "C" int vfkio_pack_plan(const vfkio_reads *r, vfkio_plan *plan) {
    if (!r || !plan) return vfkio_fail(VFK_EINVAL, "vfkio_pack_plan: NULL argument");
    int64_t placed = 0;
    while (placed < r->n && r->tid[placed] >= 0) ++placed;
    for (int64_t i = placed; i < r->n; ++i)
        if (r->tid[i] >= 0) return vfkio_fail(VFK_EINVAL, "unplaced reads must come last (coordinate-sorted input)");
    int64_t words = 0, sectors = 0;
    int32_t maxl = 0;
    for (int64_t i = 0; i < placed; ++i) {
        if (i > 0 && (r->tid[i] < r->tid[i - 1] || (r->tid[i] == r->tid[i - 1] && r->pos[i] < r->pos[i - 1])))
            return vfkio_fail(VFK_EINVAL, "read batch is not coordinate sorted");
        if (r->tid[i] >= r->n_contigs) return vfkio_fail(VFK_EINVAL, "read tid outside the contig table");
        if (!shape_of(r, i).simple) words += r->n_cigar[i];
        sectors += sectors_of(r->l_qseq[i]);
        maxl = std::max(maxl, r->l_qseq[i]);
    }
    if (sectors >= (int64_t)1 << 32) return vfkio_fail(VFK_EUNSUPPORTED, "payload exceeds 2^32 sectors: split the batch");
    if (words >= (int64_t)1 << 32) return vfkio_fail(VFK_EUNSUPPORTED, "CIGAR pool exceeds 2^32 words: split the batch");
    plan->n_placed = placed;
    plan->n_cigar_words = words;
    plan->n_sectors = sectors;
    plan->max_l_qseq = maxl;
    plan->pad = 0;
    return VFK_OK;
}

extern "C" int vfkio_pack(const vfkio_reads *r, const vfkio_plan *plan, int64_t *contig_off, vfk_read_hot *hot,
                          vfk_read_cold *cold, int32_t *sb, uint32_t *cigar, uint8_t *payload) {
    if (!r || !plan || !contig_off) return vfkio_fail(VFK_EINVAL, "vfkio_pack: NULL argument");
    const int64_t n = plan->n_placed;
    // per-read offsets (serial prefix sums), then a parallel fill
    std::vector<uint32_t> pay(n + 1), cig(n + 1);
    uint64_t s = 0, w = 0;
    for (int64_t i = 0; i < n; ++i) {
        pay[i] = (uint32_t)s;
        cig[i] = (uint32_t)w;
        s += (uint64_t)sectors_of(r->l_qseq[i]);
        if (!shape_of(r, i).simple) w += (uint64_t)r->n_cigar[i];
    }
    if ((int64_t)s != plan->n_sectors || (int64_t)w != plan->n_cigar_words)
        return vfkio_fail(VFK_EINVAL, "vfkio_pack: plan does not match the batch");
    // contig table
    {
        int64_t i = 0;
        for (int32_t t = 0; t <= r->n_contigs; ++t) {
            while (i < n && r->tid[i] < t) ++i;
            contig_off[t] = i;
        }
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const Shape sh = shape_of(r, i);
        hot[i].pos = r->pos[i];
        hot[i].span = sh.span;
        const uint32_t shape = sh.simple ? VFK_SHAPE_SIMPLE : sh.one_event ? VFK_SHAPE_ONE_EVENT : VFK_SHAPE_GENERAL;
        hot[i].fmk = (uint32_t)r->flag[i] | ((uint32_t)r->mapq[i] << 16) | (shape << 24);
        hot[i].pay = pay[i];
        cold[i].cigar_off = sh.simple ? 0u : cig[i];
        cold[i].n_cigar = (uint32_t)r->n_cigar[i];
        cold[i].l_qseq = (uint32_t)r->l_qseq[i];
        cold[i].event = sh.one_event ? sh.event : 0u;
        sb[2 * i] = r->pos[i];
        if (!sh.simple) memcpy(cigar + cig[i], r->cigar + r->cigar_off[i], (size_t)r->n_cigar[i] * 4);
        // payload: 8-byte granules of 5 positions x 12 bits (quality | base << 8), 4 granules per sector
        const int32_t l = r->l_qseq[i];
        const uint8_t *q = r->qual + r->qual_off[i];
        const uint8_t *sq = r->seq + r->seq_off[i];
        uint64_t *dst = reinterpret_cast<uint64_t *>(payload + ((size_t)pay[i] << 5));
        const int64_t n_gran = sectors_of(l) * 4;
        for (int64_t g = 0; g < n_gran; ++g) {
            uint64_t w = 0;
            for (int t = 0; t < 5; ++t) {
                const int64_t j = g * 5 + t;
                if (j >= l) break;
                const uint64_t code = (j & 1) ? (sq[j >> 1] & 15) : (sq[j >> 1] >> 4);
                w |= ((uint64_t)q[j] | (code << 8)) << (12 * t);
            }
            dst[g] = w;
        }
    }
    // back pointers: with pmax[j] the running maximum of the reference end inside the contig,
    // back[i] = first j with pmax[j] > start[i] (two monotone cursors)
    for (int32_t t = 0; t < r->n_contigs; ++t) {
        const int64_t a = contig_off[t], b = contig_off[t + 1];
        int64_t j = a, mx = INT64_MIN;
        std::vector<int64_t> pmax;
        pmax.reserve((size_t)(b - a));
        for (int64_t i = a; i < b; ++i) {
            mx = std::max<int64_t>(mx, (int64_t)hot[i].pos + hot[i].span);
            pmax.push_back(mx);
        }
        for (int64_t i = a; i < b; ++i) {
            while (j < i && pmax[(size_t)(j - a)] <= (int64_t)hot[i].pos) ++j;
            sb[2 * i + 1] = (int32_t)j;
        }
    }
    return VFK_OK;
}

// ---------------------------------------------------------------------------------------------
// Mate linking.  htslib keeps, per pileup iterator, a hash read-name -> first-seen mate
// (sam.c overlap_push).  A record that passes the pysam read filter and is proper-paired with a
// mapped mate looks its name up: found -> the two are tweaked against each other and the entry
// is deleted; not found -> it is stored if its mate is still to come (mpos >= pos).  Entries
// vanish when their record leaves the pileup buffer (overlap_remove), which happens once a column
// at or beyond its end has been emitted -- i.e. before a push iff end < start of the record
// pushed just before.  The tie-break bit is Wang(X31(name)) & 1 of the first record.
// ---------------------------------------------------------------------------------------------
namespace {
struct NameKey {
    uint64_t id;
    uint32_t h;
    bool operator==(const NameKey &o) const { return id == o.id && h == o.h; }
};
struct NameHash {
    size_t operator()(const NameKey &k) const { return (size_t)(k.id ^ ((uint64_t)k.h * 0x9E3779B97F4A7C15ull)); }
};
inline uint32_t wang(uint32_t k) {
    k += ~(k << 15); k ^= (k >> 10); k += (k << 3); k ^= (k >> 6); k += ~(k << 11); k ^= (k >> 16);
    return k;
}
}  // namespace

extern "C" int vfkio_link_mates(const vfkio_reads *r, const vfk_params *p, uint32_t *mate) {
    if (!r || !p || !mate) return vfkio_fail(VFK_EINVAL, "vfkio_link_mates: NULL argument");
    int64_t n = 0;
    while (n < r->n && r->tid[n] >= 0) ++n;
    std::fill(mate, mate + n, VFK_NO_MATE);
    if (!p->ignore_overlaps) return VFK_OK;
    std::unordered_map<NameKey, int64_t, NameHash> waiting;
    int32_t contig = -1;
    int32_t last_start = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t flag = r->flag[i];
        if (flag & (p->flag_filter | 0x4u)) continue;
        if ((int32_t)r->mapq[i] < p->min_mapq) continue;
        if (p->ignore_orphans && (flag & 0x1u) && !(flag & 0x2u)) continue;
        if (r->tid[i] != contig) {
            waiting.clear();
            contig = r->tid[i];
            last_start = r->pos[i];
        }
        const int32_t start = r->pos[i];
        const int64_t end = (int64_t)start + shape_of(r, i).span;
        const int32_t column = last_start;  // the column the engine stands on while this record is pushed
        last_start = start;
        if (end <= column) continue;  // never appended to the buffer
        if ((flag & 0x8u) || !(flag & 0x2u)) continue;
        if (r->mtid[i] >= 0 && r->mtid[i] != r->tid[i]) continue;
        if (std::llabs((long long)r->isize[i]) >= 2LL * r->l_qseq[i] && (int64_t)r->mpos[i] >= end) continue;
        const NameKey key{r->qname_id[i], r->qname_hash[i]};
        auto it = waiting.find(key);
        if (it != waiting.end()) {
            const int64_t a = it->second;
            const int64_t a_end = (int64_t)r->pos[a] + shape_of(r, a).span;
            if (a_end < column) {  // its record already left the buffer
                waiting.erase(it);
                it = waiting.end();
            }
        }
        if (it == waiting.end()) {
            if (r->mpos[i] >= r->pos[i] || ((flag & 0x1u) && r->mpos[i] == -1)) waiting[key] = i;
        } else {
            const int64_t a = it->second;
            const uint32_t sel = wang(r->qname_hash[a]) & 1u;
            mate[a] = (uint32_t)i | (sel << 31);
            mate[i] = (uint32_t)a | (sel << 31);
            waiting.erase(it);
        }
    }
    return VFK_OK;
}

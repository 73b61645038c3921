// vfkio_columns.cpp -- the reference-anchored column store (host side).
//
// The read-major HBM layout (vfkio_pack.cpp) makes a pileup column a gather: one header per candidate
// read, one 32-byte payload sector per read somewhere else, the overlap partner's header and sector
// somewhere else again.  This builder transposes the batch once: the reference is cut into windows of
// VFK_COL_WIDTH (12) positions and every read leaves one 32-byte sector in every window it overlaps,
// the sectors of a window stored together in read order.  A sector holds everything an SNV column
// needs from that read at those 12 positions (include/vfk.h "column store"):
//   12 entries of 16 bits   base quality | BAM base code << 8 | state << 12
//                           (state 0 the read does not cover the position, 1 aligned base, 2 deletion,
//                            3 reference skip -- htslib resolve_cigar2's is_del / is_refskip)
//   header word 0           query position at the window start, MAPQ, offset of the overlap partner's
//                           sector inside the same window (htslib overlap_push pairing, vfkio_link_mates),
//                           read filter verdict (pysam __advance_samtools), tie-break and order bits
//   header word 1           which positions hold aligned bases, and the one insertion inside the window
// so that a column is ONE contiguous read of (depth x 32) bytes, the mate-overlap adjustment a register
// exchange between two lanes, and no CIGAR is walked on the device.
#include <algorithm>
#include <climits>
#include <cstring>
#include <vector>
#include "vfk_io.h"

int vfkio_fail(int code, const char *msg);

namespace {
constexpr int W = VFK_COL_WIDTH;
inline bool consumes_ref(unsigned op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
inline bool is_match(unsigned op) { return op == 0 || op == 7 || op == 8; }

inline int64_t ref_span(const vfkio_reads *r, int64_t i) {
    const uint32_t *c = r->cigar + r->cigar_off[i];
    int64_t span = 0;
    for (int k = 0; k < r->n_cigar[i]; ++k)
        if (consumes_ref(c[k] & 15u)) span += c[k] >> 4;
    return span;
}
inline bool passes(const vfkio_reads *r, int64_t i, const vfk_params *p) {
    const uint32_t flag = r->flag[i];
    if (flag & (p->flag_filter | 4u)) return false;
    if ((int)r->mapq[i] < p->min_mapq) return false;
    if (p->ignore_orphans && (flag & 1u) && !(flag & 2u)) return false;
    return true;
}
int64_t placed_reads(const vfkio_reads *r) {
    int64_t n = 0;
    while (n < r->n && r->tid[n] >= 0) ++n;
    return n;
}

struct Sector {
    uint16_t e[W];
    uint32_t h0, h1;
};
static_assert(sizeof(Sector) == 32, "a column-store sector is 32 bytes");
}  // namespace

extern "C" int vfkio_columns_plan(const vfkio_reads *r, int64_t *wbase /*[n_contigs+1]*/, vfkio_col_plan *plan) {
    if (!r || !wbase || !plan) return vfkio_fail(VFK_EINVAL, "vfkio_columns_plan: NULL argument");
    const int64_t n = placed_reads(r);
    std::vector<int64_t> extent((size_t)r->n_contigs, 0);
    int64_t sectors = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t span = ref_span(r, i);
        if (span <= 0) continue;
        const int64_t beg = r->pos[i], end = beg + span;
        extent[(size_t)r->tid[i]] = std::max(extent[(size_t)r->tid[i]], end);
        sectors += (end - 1) / W - beg / W + 1;
    }
    wbase[0] = 0;
    for (int32_t t = 0; t < r->n_contigs; ++t) wbase[t + 1] = wbase[t] + (extent[(size_t)t] + W - 1) / W;
    if (sectors >= ((int64_t)1 << 32)) return vfkio_fail(VFK_EUNSUPPORTED, "column store exceeds 2^32 sectors: split the batch");
    plan->n_windows = wbase[r->n_contigs];
    plan->n_sectors = sectors;
    return VFK_OK;
}

extern "C" int vfkio_columns_fill(const vfkio_reads *r, const vfk_params *params, const uint32_t *mate, const int64_t *wbase,
                                  const vfkio_col_plan *plan, uint32_t *woff /*[n_windows+1]*/, uint8_t *sectors_out) {
    if (!r || !params || !mate || !wbase || !plan || !woff || (!sectors_out && plan->n_sectors))
        return vfkio_fail(VFK_EINVAL, "vfkio_columns_fill: NULL argument");
    const int64_t n = placed_reads(r);
    const int64_t nw = plan->n_windows;
    Sector *sec = reinterpret_cast<Sector *>(sectors_out);
    // ---- read chunks.  The global index of a read's first window (wbase[tid] + pos / W) never decreases along the
    // sorted batch, so a contiguous chunk of reads touches one contiguous range of windows [lo, hi]; chunks are
    // processed in parallel and meet only in the few windows their ranges share.
    const int64_t n_chunks = std::max<int64_t>(1, std::min<int64_t>(512, n / 2048));
    struct Chunk {
        int64_t r0, r1;   // reads
        int64_t lo, hi;   // windows touched (hi < lo: none)
        std::vector<uint32_t> cnt;  // sectors this chunk leaves in window lo + k
    };
    std::vector<Chunk> chunks((size_t)n_chunks);
    std::vector<int32_t> span_of((size_t)n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < n_chunks; ++t) {
        Chunk &ck = chunks[(size_t)t];
        ck.r0 = n * t / n_chunks;
        ck.r1 = n * (t + 1) / n_chunks;
        ck.lo = 0;
        ck.hi = -1;
        bool any = false;
        for (int64_t i = ck.r0; i < ck.r1; ++i) {
            const int64_t span = ref_span(r, i);
            span_of[(size_t)i] = (int32_t)std::min<int64_t>(span, INT32_MAX);
            if (span <= 0) continue;
            const int64_t w0 = wbase[r->tid[i]] + r->pos[i] / W, w1 = wbase[r->tid[i]] + (r->pos[i] + span - 1) / W;
            if (!any) { ck.lo = w0; any = true; }
            ck.hi = std::max(ck.hi, w1);
        }
        if (!any) continue;
        ck.cnt.assign((size_t)(ck.hi - ck.lo + 1), 0u);
        for (int64_t i = ck.r0; i < ck.r1; ++i) {
            const int64_t span = span_of[(size_t)i];
            if (span <= 0) continue;
            const int64_t w0 = wbase[r->tid[i]] + r->pos[i] / W, w1 = wbase[r->tid[i]] + (r->pos[i] + span - 1) / W;
            for (int64_t w = w0; w <= w1; ++w) ++ck.cnt[(size_t)(w - ck.lo)];
        }
    }
    // ---- sectors per window, offsets
    {
        std::vector<uint32_t> count((size_t)nw + 1, 0);
        for (const Chunk &ck : chunks)
            for (size_t k = 0; k < ck.cnt.size(); ++k) count[(size_t)ck.lo + k] += ck.cnt[k];
        uint64_t acc = 0;
        for (int64_t w = 0; w < nw; ++w) {
            woff[w] = (uint32_t)acc;
            acc += count[(size_t)w];
        }
        woff[nw] = (uint32_t)acc;
        if ((int64_t)acc != plan->n_sectors) return vfkio_fail(VFK_EINVAL, "vfkio_columns_fill: plan does not match the batch");
    }
    // ---- fill: reads in index order inside a chunk, a chunk's first slot in a window = what the chunks before it
    // left there, so that the sectors of a window are sorted by read
    std::vector<uint32_t> sec_read((size_t)plan->n_sectors);  // owner of every sector, for the partner offsets
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < n_chunks; ++t) {
        const Chunk &ck = chunks[(size_t)t];
        if (ck.hi < ck.lo) continue;
        std::vector<uint32_t> cursor(ck.cnt.size(), 0u);
        for (int64_t s2 = 0; s2 < t; ++s2) {
            const Chunk &e = chunks[(size_t)s2];
            const int64_t a = std::max(e.lo, ck.lo), b = std::min(e.hi, ck.hi);
            for (int64_t w = a; w <= b; ++w) cursor[(size_t)(w - ck.lo)] += e.cnt[(size_t)(w - e.lo)];
        }
        for (int64_t i = ck.r0; i < ck.r1; ++i) {
            if (span_of[(size_t)i] <= 0) continue;
            const int64_t wb = wbase[r->tid[i]];
            const uint32_t *c = r->cigar + r->cigar_off[i];
            const int nc = r->n_cigar[i];
            const uint8_t *q = r->qual + r->qual_off[i], *sq = r->seq + r->seq_off[i];
            const bool pass = passes(r, i, params);
            const uint32_t flag = r->flag[i];
            const bool linked = params->ignore_overlaps && (flag & 2u) && !(flag & 8u) && mate[i] != VFK_NO_MATE;
            const uint32_t m = mate[i] & 0x7FFFFFFFu;
            const uint32_t h0_fixed = ((uint32_t)r->mapq[i] << 12) | (pass ? 1u << 28 : 0u) |
                                      (linked && (mate[i] >> 31) ? 1u << 29 : 0u) | (linked && (uint32_t)i < m ? 1u << 30 : 0u);
            int64_t y = 0;
            int64_t cur_w = r->pos[i] / W;       // window (of the contig) holding the next reference position
            int p = (int)(r->pos[i] - cur_w * W);  // ... and the slot of that position inside it
            Sector cur;
            bool open = false;  // `cur` holds entries of window cur_w
            uint32_t qpos0 = 0, mask = 0, ins_at = 0, ins_len = 0;
            bool has_ins = false, complex = false;
            auto flush = [&]() {
                if (!open) return;
                if (qpos0 >= 4096u || ins_len >= 4096u) complex = true;
                cur.h0 = h0_fixed | (qpos0 & 0xFFFu) | (complex ? 1u << 31 : 0u);
                cur.h1 = mask | (ins_at << 12) | ((ins_len & 0xFFFu) << 16) | (has_ins ? 1u << 28 : 0u);
                const int64_t w = wb + cur_w;
                const uint32_t at = woff[w] + cursor[(size_t)(w - ck.lo)]++;
                sec[at] = cur;
                sec_read[at] = (uint32_t)i;
                open = false;
            };
            auto enter = [&]() {  // the first entry of window cur_w is about to be written
                memset(&cur, 0, sizeof(cur));
                qpos0 = (uint32_t)y;
                mask = 0; ins_at = 0; ins_len = 0;
                has_ins = false; complex = false;
                open = true;
            };
            for (int k = 0; k < nc; ++k) {
                const uint32_t op = c[k] & 15u;
                int64_t l = c[k] >> 4;
                if (consumes_ref(op)) {
                    const bool match = is_match(op);
                    const uint16_t gap = (uint16_t)((op == 3 ? 3u : 2u) << 12);
                    while (l > 0) {
                        if (!open) enter();
                        const int run = (int)std::min<int64_t>(l, W - p);
                        if (match) {
                            for (int j = 0; j < run; ++j, ++y) {
                                const uint32_t code = (y & 1) ? (sq[y >> 1] & 15u) : (sq[y >> 1] >> 4);
                                cur.e[p + j] = (uint16_t)(q[y] | (code << 8) | (1u << 12));
                            }
                            mask |= ((1u << run) - 1u) << p;
                        } else {
                            for (int j = 0; j < run; ++j) cur.e[p + j] = gap;
                        }
                        p += run;
                        l -= run;
                        if (p == W) {
                            flush();
                            p = 0;
                            ++cur_w;
                        }
                    }
                } else if (op == 1) {  // insertion: follows the reference position before (cur_w, p)
                    if (open && p != 0) {  // that position lies in the open window
                        if (has_ins) complex = true;
                        has_ins = true;
                        ins_at = (uint32_t)(p - 1);
                        ins_len = (uint32_t)std::min<int64_t>(l, 4096);
                    }
                    y += l;
                } else if (op == 4) {
                    y += l;
                }
            }
            flush();
        }
    }
    // ---- partner offsets: the partner's sector inside the same window, found by read index (runs are sorted)
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t w = 0; w < nw; ++w) {
        const uint32_t a = woff[w], b = woff[w + 1];
        for (uint32_t s = a; s < b; ++s) {
            const int64_t i = sec_read[s];
            const uint32_t flag = r->flag[i];
            if (!(params->ignore_overlaps && (flag & 2u) && !(flag & 8u) && mate[i] != VFK_NO_MATE)) continue;
            const uint32_t m = mate[i] & 0x7FFFFFFFu;
            const uint32_t *lo = std::lower_bound(sec_read.data() + a, sec_read.data() + b, m);
            if (lo == sec_read.data() + b || *lo != m) continue;  // the partner does not reach this window
            const int64_t off = (int64_t)(lo - sec_read.data()) - (int64_t)s;
            if (off < -127 || off > 127) sec[s].h0 |= 1u << 31;  // out of reach for one byte: leave the unit to the read-major path
            else sec[s].h0 |= ((uint32_t)off & 0xFFu) << 20;
        }
    }
    return VFK_OK;
}

// ---------------------------------------------------------------------------------------------
// Window-transposed form of the same store (prepared for the next kernel revision; the shipped kernels read the
// sector form above).  A column needs 10 of the 32 bytes of each sector -- one entry and the two header words --
// and the sector form makes the device (or, on the host entry point, PCIe) move all 32.  Transposed, the run of a
// window with n sectors keeps its place and size (32 n bytes from byte 32 * woff[w]) but is laid out as
//     n header pairs {h0, h1}            8 n bytes
//     12 rows of n entries (u16)         row p = position p of the window: 2 n bytes each
// so that a column reads one contiguous 8 n-byte block and one contiguous 2 n-byte row: 310 bytes instead of
// 992 at n = 31.  Sector k of the run <-> header pair k and entry k of every row.
// ---------------------------------------------------------------------------------------------
extern "C" int vfkio_columns_transpose(const uint32_t *woff, int64_t n_windows, const uint8_t *sectors, uint8_t *out) {
    if (!woff || n_windows < 0 || ((!sectors || !out) && n_windows && woff[n_windows])) return vfkio_fail(VFK_EINVAL, "vfkio_columns_transpose: NULL argument");
    if (sectors == out) return vfkio_fail(VFK_EINVAL, "vfkio_columns_transpose: cannot work in place");
#pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t w = 0; w < n_windows; ++w) {
        const uint32_t a = woff[w], n = woff[w + 1] - a;
        if (!n) continue;
        const Sector *src = reinterpret_cast<const Sector *>(sectors) + a;
        uint8_t *base = out + (size_t)a * 32u;
        uint32_t *hdr = reinterpret_cast<uint32_t *>(base);
        uint16_t *rows = reinterpret_cast<uint16_t *>(base + (size_t)n * 8u);
        for (uint32_t k = 0; k < n; ++k) {
            hdr[2 * k] = src[k].h0;
            hdr[2 * k + 1] = src[k].h1;
            for (int p = 0; p < W; ++p) rows[(size_t)p * n + k] = src[k].e[p];
        }
    }
    return VFK_OK;
}


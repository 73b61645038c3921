// vfkio_bam.cpp -- BGZF / BAM decoder on zlib: file -> canonical columnar batch.
//
// Stands where pysam.AlignmentFile / htslib stand in the reference (vafator/annotator.py:161-164);
// htslib is not available on either box, so the container format is decoded here
// from the SAM/BAM specification (SAMv1 section 4): BGZF = concatenated gzip members carrying a
// 'BC' extra field with the block size; BAM = magic, header text, reference table, records.
// Blocks are inflated in parallel (OpenMP), records are indexed in one sequential hop and then
// converted to columns in parallel.  CRAM is not supported (the reference accepts it through
// htslib; out of scope here and reported as an error).
#include <fcntl.h>
#include <omp.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "vfk_io.h"

int vfkio_fail(int code, const char *msg);

// Arena of inflated bytes.  Large ones are 2 MB aligned and marked for transparent huge pages: a fresh 4 KB-page mapping costs
// a page fault per 4 KB on first touch (a third of a decode went there), a huge-page one per 2 MB.
struct ArenaFree {
    void operator()(uint8_t *p) const { free(p); }
};
using Arena = std::unique_ptr<uint8_t[], ArenaFree>;
inline Arena arena_alloc(size_t bytes) {
    if (bytes < (size_t)(4u << 20)) return Arena((uint8_t *)malloc(bytes ? bytes : 1));
    void *p = nullptr;
    const size_t rounded = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    if (posix_memalign(&p, (size_t)2 << 20, rounded) != 0) return Arena((uint8_t *)malloc(bytes));
    madvise(p, rounded, MADV_HUGEPAGE);  // a hint: harmless where THP is off
    return Arena((uint8_t *)p);
}

struct vfkio_bam {
    // inflated BAM bytes: the whole stream (vfkio_bam_open) or one arena per batch of blocks of a region scan
    // (vfkio_bam_open_regions).  Allocated uninitialised: the inflating threads touch the pages first, in parallel.
    std::vector<Arena> arenas;
    std::vector<std::string> names;
    std::vector<int64_t> lengths;
    std::string text;
    std::vector<const uint8_t *> rec;  // every placed record (its block_size field), in file order, inside an arena
    int64_t n_unplaced = 0;
    int64_t n_cigar = 0, n_seq = 0, n_qual = 0;
    int32_t sorted = 1;
};

namespace {
inline uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t le16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// one inflater per thread: the own decoder first, a zlib stream (reset between blocks) when that one declines
struct Inflater {
    z_stream zs;
    bool ready = false;
    ~Inflater() { if (ready) inflateEnd(&zs); }
    // inflate one BGZF payload to exactly `cap` bytes and verify the member's CRC32 (the four bytes after the payload), as htslib does
    int block(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
        if (inflate_only(src, n, dst, cap)) return -1;
        return (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, (uInt)cap) == le32(src + n) ? 0 : -1;
    }
    int inflate_only(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
        static const bool zlib_only = getenv("VFKIO_ZLIB") != nullptr;
        if (!zlib_only && vfkio_inflate_raw(src, n, dst, cap) == 0) return 0;  // vfkio_inflate.cpp; zlib decides the rest
        if (!ready) {  // the zlib state (32 KB window and all) is only set up by a thread that needs it
            memset(&zs, 0, sizeof(zs));
            if (inflateInit2(&zs, -15) != Z_OK) return -1;
            ready = true;
        } else if (inflateReset(&zs) != Z_OK) {
            return -1;
        }
        zs.next_in = const_cast<uint8_t *>(src);
        zs.avail_in = (uInt)n;
        zs.next_out = dst;
        zs.avail_out = (uInt)cap;
        const int rc = inflate(&zs, Z_FINISH);
        return (rc == Z_STREAM_END && zs.avail_out == 0) ? 0 : -1;
    }
};
// the compressed file, mapped read-only
struct Mapped {
    const uint8_t *p = nullptr;
    size_t n = 0;
    ~Mapped() { if (p && n) munmap(const_cast<uint8_t *>(p), n); }
    bool open(const char *path) {
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) { ::close(fd); return false; }
        n = (size_t)st.st_size;
        if (n) {
            void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m == MAP_FAILED) { ::close(fd); n = 0; return false; }
            madvise(m, n, MADV_SEQUENTIAL);
            p = (const uint8_t *)m;
        }
        ::close(fd);
        return true;
    }
    const uint8_t *data() const { return p; }
    size_t size() const { return n; }
    uint8_t operator[](size_t i) const { return p[i]; }
};

// fixed fields + name + CIGAR + sequence + qualities must lie inside the record
inline bool record_fits(const uint8_t *r, uint32_t block_size) {
    const uint64_t l_name = r[8], n_cig = le16(r + 12), l_seq = le32(r + 16);
    return 32 + l_name + 4 * n_cig + (l_seq + 1) / 2 + l_seq <= block_size;
}
struct Blk { size_t src, n, dst, isize; };  // deflate payload [src, src+n) of the file -> [dst, dst+isize) of the arena
// BGZF member at file offset p (SAMv1 4.1): 1 = parsed (payload -> blk.src / blk.n / blk.isize, member length -> blen),
// 0 = clean end of file, -1 = not BGZF / truncated
int bgzf_member(const Mapped &raw, size_t p, Blk &blk, size_t &blen) {
    if (p == raw.size()) return 0;
    if (p + 18 > raw.size()) return -1;
    const uint8_t *h = raw.data() + p;
    if (!(h[0] == 31 && h[1] == 139 && h[2] == 8 && (h[3] & 4))) return -1;
    const uint16_t xlen = le16(h + 10);
    size_t x = p + 12;
    const size_t xe = x + xlen;
    if (xe > raw.size()) return -1;
    int bsize = -1;
    while (x + 4 <= xe) {
        const uint16_t slen = le16(raw.data() + x + 2);
        if (raw[x] == 'B' && raw[x + 1] == 'C' && slen == 2 && x + 6 <= xe) bsize = le16(raw.data() + x + 4);
        x += 4 + slen;
    }
    if (bsize < 0) return -1;
    blen = (size_t)bsize + 1;
    if (p + blen > raw.size() || blen < (xe - p) + 8) return -1;
    blk.src = xe;
    blk.n = blen - (xe - p) - 8;
    blk.isize = le32(raw.data() + p + blen - 4);
    return 1;
}
// inflate a table of blocks into `dst` (all cores); false on a corrupt block
bool inflate_blocks(const Mapped &raw, const std::vector<Blk> &blocks, uint8_t *dst) {
    int bad = 0;
#pragma omp parallel reduction(| : bad)
    {
        Inflater inf;
#pragma omp for schedule(dynamic, 1)
        for (int64_t i = 0; i < (int64_t)blocks.size(); ++i) {
            const Blk &k = blocks[(size_t)i];
            if (k.isize && inf.block(raw.data() + k.src, k.n, dst + k.dst, k.isize)) bad |= 1;
        }
    }
    return !bad;
}
}  // namespace

extern "C" int vfkio_bam_open(const char *path, vfkio_bam **out) {
    if (!path || !out) return vfkio_fail(VFK_EINVAL, "vfkio_bam_open: NULL argument");
    Mapped raw;
    if (!raw.open(path)) return vfkio_fail(VFK_EINVAL, "cannot open alignment file");
    if (raw.size() >= 4 && memcmp(raw.data(), "CRAM", 4) == 0)
        return vfkio_fail(VFK_EUNSUPPORTED, "CRAM input is not supported by this build: convert to BAM");
    // ---- BGZF block table
    std::vector<Blk> blocks;
    size_t p = 0, total = 0;
    for (;;) {
        Blk k;
        size_t blen = 0;
        const int rc = bgzf_member(raw, p, k, blen);
        if (rc == 0) break;
        if (rc < 0) return vfkio_fail(VFK_EINVAL, blocks.empty() ? "not a BGZF/BAM file" : "corrupt BGZF block header");
        k.dst = total;
        blocks.push_back(k);
        total += k.isize;
        p += blen;
    }
    vfkio_bam *b = new vfkio_bam();
    b->arenas.push_back(arena_alloc(std::max<size_t>(total, 1)));
    uint8_t *const d = b->arenas[0].get();
    // ---- BAM header, then the record index: a sequential hop from record to record (placed reads only) over the
    // inflated prefix of the stream.  It runs on one thread while the others inflate the next batch of blocks.
    struct {
        size_t q = 0;
        bool header = false;
        uint32_t n_ref = 0;
        int32_t last_tid = -1, last_pos = -1;
        bool seen_unplaced = false;
    } hs;
    auto hop = [&](size_t avail, bool final) -> const char * {
        if (!hs.header) {
            if (avail < 12) return final ? "not a BAM file (bad magic)" : nullptr;
            if (memcmp(d, "BAM\1", 4) != 0) return "not a BAM file (bad magic)";
            size_t q = 4;
            const uint32_t l_text = le32(d + q); q += 4;
            if (q + l_text + 4 > avail) return final ? "truncated BAM header" : nullptr;
            const size_t text_at = q;
            q += l_text;
            const uint32_t n_ref = le32(d + q); q += 4;
            std::vector<std::string> names;
            std::vector<int64_t> lengths;
            for (uint32_t r = 0; r < n_ref; ++r) {
                if (q + 4 > avail) return final ? "truncated BAM reference table" : nullptr;
                const uint32_t l_name = le32(d + q); q += 4;
                if (q + l_name + 4 > avail) return final ? "truncated BAM reference table" : nullptr;
                names.emplace_back((const char *)d + q, l_name ? l_name - 1 : 0);
                q += l_name;
                lengths.push_back((int64_t)le32(d + q));
                q += 4;
            }
            b->text.assign((const char *)d + text_at, l_text);
            b->names.swap(names);
            b->lengths.swap(lengths);
            hs.n_ref = n_ref;
            hs.q = q;
            hs.header = true;
        }
        while (hs.q + 4 <= avail) {
            const size_t q = hs.q;
            const uint32_t bs = le32(d + q);
            if (bs < 32) return "truncated BAM record";
            if (q + 4 + bs > avail) {  // the record ends in a block that is not inflated yet
                if (final) return "truncated BAM record";
                break;
            }
            const uint8_t *r = d + q + 4;
            const int32_t tid = (int32_t)le32(r), pos = (int32_t)le32(r + 4);
            if (!record_fits(r, bs)) return "corrupt BAM record";
            if (tid >= 0 && tid < (int32_t)hs.n_ref) {
                if (hs.seen_unplaced || tid < hs.last_tid || (tid == hs.last_tid && pos < hs.last_pos)) b->sorted = 0;
                hs.last_tid = tid;
                hs.last_pos = pos;
                b->rec.push_back(d + q);
                b->n_cigar += le16(r + 12);
                const int64_t l_seq = (int64_t)le32(r + 16);
                b->n_seq += (l_seq + 1) / 2;
                b->n_qual += l_seq;
            } else {
                hs.seen_unplaced = true;
                ++b->n_unplaced;
            }
            hs.q = q + 4 + bs;
        }
        return nullptr;
    };
    const char *err = nullptr;
    int bad = 0;
    constexpr size_t BATCH = 512, PER_TASK = 8;
#pragma omp parallel
#pragma omp single
    {
        size_t inflated = 0;  // length of the prefix every block of which has been inflated
        for (size_t k0 = 0; k0 < blocks.size(); k0 += BATCH) {
            const size_t k1 = std::min(blocks.size(), k0 + BATCH);
            for (size_t i = k0; i < k1; i += PER_TASK) {
#pragma omp task firstprivate(i) shared(bad, blocks, raw)
                {
                    Inflater inf;
                    for (size_t j = i; j < std::min(k1, i + PER_TASK); ++j) {
                        const Blk &k = blocks[j];
                        if (k.isize && inf.block(raw.data() + k.src, k.n, d + k.dst, k.isize)) {
#pragma omp atomic write
                            bad = 1;
                        }
                    }
                }
            }
            if (!err) err = hop(inflated, false);  // the batch before this one, while this one inflates
#pragma omp taskwait
            inflated = blocks[k1 - 1].dst + blocks[k1 - 1].isize;
        }
        int failed;
#pragma omp atomic read
        failed = bad;
        if (!err && !failed) err = hop(total, true);
    }
    if (bad) { delete b; return vfkio_fail(VFK_EINVAL, "BGZF block failed to inflate"); }
    if (err) { delete b; return vfkio_fail(VFK_EINVAL, err); }
    *out = b;
    return VFK_OK;
}


// ---------------------------------------------------------------------------------------------
// Indexed access: decode only the records overlapping a list of regions, the way the reference
// fetches [first.POS-1, last.POS) of every chromosome group through pysam (vafator/pileups.py:71-80,
// :137-138).  The BAI linear index (SAMv1 5.1.3/5.2: one virtual offset per 16 kbp window = the
// first record overlapping that window) gives a lower bound; records are then read sequentially until
// one starts at or beyond the region end.  The file is mapped; BGZF blocks are inflated in growing batches (OpenMP)
// into arenas the kept records then live in (no per-record copy).
// ---------------------------------------------------------------------------------------------
namespace {
inline bool consumes_ref_op(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }

// header text + reference table from the start of an inflated stream; returns bytes consumed, 0 = need more data
size_t parse_header(const std::vector<uint8_t> &d, vfkio_bam *b, bool *bad) {
    *bad = false;
    const size_t n = d.size();
    if (n < 12) return 0;
    if (memcmp(d.data(), "BAM\1", 4) != 0) { *bad = true; return 0; }
    size_t q = 4;
    const uint32_t l_text = le32(d.data() + q); q += 4;
    if (q + l_text + 4 > n) return 0;
    const size_t text_at = q;
    q += l_text;
    const uint32_t n_ref = le32(d.data() + q); q += 4;
    std::vector<std::string> names;
    std::vector<int64_t> lengths;
    for (uint32_t r = 0; r < n_ref; ++r) {
        if (q + 4 > n) return 0;
        const uint32_t l_name = le32(d.data() + q); q += 4;
        if (q + l_name + 4 > n) return 0;
        names.emplace_back((const char *)d.data() + q, l_name ? l_name - 1 : 0);
        q += l_name;
        lengths.push_back((int64_t)le32(d.data() + q));
        q += 4;
    }
    b->text.assign((const char *)d.data() + text_at, l_text);
    b->names.swap(names);
    b->lengths.swap(lengths);
    return q;
}
}  // namespace

extern "C" int vfkio_bam_open_regions(const char *path, const char *index_path, int32_t n_regions, const char *const *contig,
                                      const int32_t *beg, const int32_t *end, vfkio_bam **out) {
    if (!path || !index_path || !out || n_regions < 0 || (n_regions > 0 && (!contig || !beg || !end)))
        return vfkio_fail(VFK_EINVAL, "vfkio_bam_open_regions: bad argument");
    Mapped raw;
    if (!raw.open(path)) return vfkio_fail(VFK_EINVAL, "cannot open alignment file");
    if (raw.size() >= 4 && memcmp(raw.data(), "CRAM", 4) == 0)
        return vfkio_fail(VFK_EUNSUPPORTED, "CRAM input is not supported by this build: convert to BAM");
    vfkio_bam *b = new vfkio_bam();
    // ---- header: inflate leading blocks until the text and the reference table are complete
    {
        std::vector<uint8_t> head;
        size_t cp = 0, used = 0;
        while (used == 0) {
            std::vector<Blk> blocks;
            size_t total = head.size();
            for (int k = 0; k < 16; ++k) {
                Blk blk;
                size_t blen = 0;
                const int rc = bgzf_member(raw, cp, blk, blen);
                if (rc < 0) { delete b; return vfkio_fail(VFK_EINVAL, "not a BGZF/BAM file"); }
                if (rc == 0) break;
                blk.dst = total;
                blocks.push_back(blk);
                total += blk.isize;
                cp += blen;
            }
            head.resize(total);
            if (!inflate_blocks(raw, blocks, head.data())) { delete b; return vfkio_fail(VFK_EINVAL, "not a BGZF/BAM file"); }
            bool bad = false;
            used = parse_header(head, b, &bad);
            if (bad) { delete b; return vfkio_fail(VFK_EINVAL, "not a BAM file (bad magic)"); }
            if (used == 0 && blocks.empty()) { delete b; return vfkio_fail(VFK_EINVAL, "truncated BAM header"); }
        }
    }
    if (n_regions == 0) { *out = b; return VFK_OK; }
    // ---- the BAI: per contig the binning index (bin -> chunks of virtual offsets) and the linear index
    struct Chunk { uint64_t beg, end; };
    std::vector<std::vector<std::pair<uint32_t, Chunk>>> bins(b->names.size());
    std::vector<std::vector<uint64_t>> linear(b->names.size());
    auto le64 = [](const uint8_t *p8) { return (uint64_t)le32(p8) | ((uint64_t)le32(p8 + 4) << 32); };
    {
        FILE *fi = fopen(index_path, "rb");
        if (!fi) { delete b; return vfkio_fail(VFK_EINVAL, "cannot open the BAM index"); }
        fseek(fi, 0, SEEK_END);
        const long isz = ftell(fi);
        fseek(fi, 0, SEEK_SET);
        std::vector<uint8_t> ix((size_t)std::max<long>(isz, 0));
        const bool ok = isz > 0 && fread(ix.data(), 1, (size_t)isz, fi) == (size_t)isz;
        fclose(fi);
        if (!ok || ix.size() < 8 || memcmp(ix.data(), "BAI\1", 4) != 0) { delete b; return vfkio_fail(VFK_EINVAL, "not a BAI index"); }
        size_t q = 4;
        const uint32_t n_ref = le32(ix.data() + q); q += 4;
        if (n_ref != b->names.size()) { delete b; return vfkio_fail(VFK_EINVAL, "BAM index does not match the BAM header"); }
        for (uint32_t r = 0; r < n_ref; ++r) {
            if (q + 4 > ix.size()) { delete b; return vfkio_fail(VFK_EINVAL, "truncated BAM index"); }
            const uint32_t n_bin = le32(ix.data() + q); q += 4;
            for (uint32_t k = 0; k < n_bin; ++k) {
                if (q + 8 > ix.size()) { delete b; return vfkio_fail(VFK_EINVAL, "truncated BAM index"); }
                const uint32_t bin = le32(ix.data() + q), n_chunk = le32(ix.data() + q + 4);
                q += 8;
                if (q + (size_t)n_chunk * 16 > ix.size()) { delete b; return vfkio_fail(VFK_EINVAL, "truncated BAM index"); }
                for (uint32_t c = 0; c < n_chunk; ++c, q += 16)
                    if (bin < 37450u) bins[r].push_back({bin, {le64(ix.data() + q), le64(ix.data() + q + 8)}});  // 37450: metadata
            }
            if (q + 4 > ix.size()) { delete b; return vfkio_fail(VFK_EINVAL, "truncated BAM index"); }
            const uint32_t n_intv = le32(ix.data() + q); q += 4;
            if (q + (size_t)n_intv * 8 > ix.size()) { delete b; return vfkio_fail(VFK_EINVAL, "truncated BAM index"); }
            std::stable_sort(bins[r].begin(), bins[r].end(),
                             [](const std::pair<uint32_t, Chunk> &x, const std::pair<uint32_t, Chunk> &y) { return x.first < y.first; });
            linear[r].resize(n_intv);
            for (uint32_t k = 0; k < n_intv; ++k) linear[r][k] = le64(ix.data() + q + (size_t)k * 8);
            q += (size_t)n_intv * 8;
        }
    }
    // ---- regions, in file order: validate, find each one's scan start, then scan.  Many regions (windows around
    // sparse variants) are scanned in parallel, each with a serial inflate; a few long ones one after the other,
    // each inflating its block batches on all cores.
    struct Job {
        int32_t tid, r_beg, r_end, keep_from;
        uint64_t voff;
        std::vector<Arena> arenas;
        std::vector<const uint8_t *> rec;
        int64_t n_cigar = 0, n_seq = 0, n_qual = 0;
        const char *error = nullptr;
    };
    std::vector<Job> jobs;
    {
        int32_t prev_tid = -1, prev_end = -1;        // region order check
        int32_t scanned_tid = -1, scanned_end = 0;   // the region scanned before this one
        for (int32_t g = 0; g < n_regions; ++g) {
            int32_t tid = -1;
            for (size_t t = 0; t < b->names.size(); ++t)
                if (b->names[t] == contig[g]) tid = (int32_t)t;
            if (tid < 0) continue;  // a contig the BAM does not know has no reads
            if (tid < prev_tid || (tid == prev_tid && beg[g] < prev_end)) {
                delete b;
                return vfkio_fail(VFK_EINVAL, "vfkio_bam_open_regions: regions must be sorted and disjoint");
            }
            prev_tid = tid;
            prev_end = end[g];
            const std::vector<uint64_t> &lin = linear[(size_t)tid];
            const int32_t r_beg = std::max(beg[g], 0), r_end = end[g];
            if (lin.empty() || r_end <= r_beg) continue;
            size_t w = std::min<size_t>((size_t)(r_beg >> 14), lin.size() - 1);
            uint64_t min_off = lin[w];
            // an empty window carries 0 (or the value of the window before it): every read overlapping the region
            // then overlaps a later window, whose entry is a valid lower bound
            while (min_off == 0 && w + 1 < lin.size()) min_off = lin[++w];
            if (min_off == 0) continue;
            // Start of the scan: the earliest chunk of a bin overlapping the region that is not entirely before the
            // linear-index bound (the linear index only knows mapped reads; a placed unmapped mate in front of the
            // first mapped read of a window is reachable through its bin only).  From there the file is read in order.
            uint64_t voff = min_off;
            {
                const int64_t e1 = (int64_t)r_end - 1;
                const auto &bv = bins[(size_t)tid];  // sorted by bin number
                auto consider = [&](uint32_t first_bin, uint32_t last_bin) {
                    auto it = std::lower_bound(bv.begin(), bv.end(), first_bin,
                                               [](const std::pair<uint32_t, Chunk> &x, uint32_t v) { return x.first < v; });
                    for (; it != bv.end() && it->first <= last_bin; ++it)
                        if (it->second.end > min_off && it->second.beg < voff) voff = it->second.beg;
                };
                consider(0u, 0u);
                for (int lvl = 1, shift = 26; lvl <= 5; ++lvl, shift -= 3) {
                    const uint32_t base = ((1u << (3 * lvl)) - 1) / 7;
                    consider(base + (uint32_t)(r_beg >> shift), base + (uint32_t)(e1 >> shift));
                }
            }
            jobs.emplace_back();
            Job &j = jobs.back();
            j.tid = tid; j.r_beg = r_beg; j.r_end = r_end; j.voff = voff;
            // a record starting before the end of the previous region of this contig and reaching into this one was
            // kept there (it overlaps both): every record is delivered once, in file order
            j.keep_from = tid == scanned_tid ? scanned_end : INT32_MIN;
            scanned_tid = tid;
            scanned_end = r_end;
        }
    }
    auto scan = [&raw](Job &j, bool parallel_inflate) {
        const int32_t tid = j.tid, r_beg = j.r_beg, r_end = j.r_end;
        size_t cp = (size_t)(j.voff >> 16);      // file offset of the next BGZF member
        size_t q = (size_t)(j.voff & 0xFFFF);    // scan position inside the current arena
        std::vector<uint8_t> carry;              // head of a record that continues in the next batch
        // blocks per batch: one per core to begin with (a window around one variant needs the few blocks between the
        // 16 kbp index point and its last read), doubling, so that a chromosome-long region soon runs in large batches
        size_t batch_blocks = parallel_inflate ? (size_t)std::max(4, omp_get_max_threads()) : 2;
        Inflater serial;
        bool done = false;
        while (!done) {
            std::vector<Blk> blocks;
            size_t total = carry.size();
            while (blocks.size() < batch_blocks) {
                Blk blk;
                size_t blen = 0;
                const int rc = bgzf_member(raw, cp, blk, blen);
                if (rc < 0) { j.error = "BGZF block failed to inflate"; return; }
                if (rc == 0) break;
                blk.dst = total;
                blocks.push_back(blk);
                total += blk.isize;
                cp += blen;
            }
            Arena arena = arena_alloc(std::max<size_t>(total, 1));
            if (!carry.empty()) memcpy(arena.get(), carry.data(), carry.size());
            if (parallel_inflate) {
                if (!inflate_blocks(raw, blocks, arena.get())) { j.error = "BGZF block failed to inflate"; return; }
            } else {
                for (const Blk &k : blocks)
                    if (k.isize && serial.block(raw.data() + k.src, k.n, arena.get() + k.dst, k.isize)) {
                        j.error = "BGZF block failed to inflate";
                        return;
                    }
            }
            const uint8_t *buf = arena.get();
            if (q > total) { j.error = "BAM index points beyond a BGZF block"; return; }
            bool kept = false;
            while (q + 4 <= total) {
                const uint32_t bs = le32(buf + q);
                if (bs < 32) { j.error = "corrupt BAM record"; return; }
                if (q + 4 + bs > total) break;  // the record continues in the next batch
                const uint8_t *r = buf + q + 4;
                const int32_t rt = (int32_t)le32(r), rp = (int32_t)le32(r + 4);
                if (rt != tid || rp >= r_end) {
                    if (rt > tid || rt < 0 || (rt == tid && rp >= r_end)) { done = true; break; }
                    q += 4 + bs;  // still on an earlier contig (lower bound of the index): skip
                    continue;
                }
                if (rp < j.keep_from) { q += 4 + bs; continue; }
                const uint32_t l_name = r[8], n_cig = le16(r + 12);
                if (!record_fits(r, bs)) { j.error = "corrupt BAM record"; return; }
                const uint8_t *pc = r + 32 + l_name;
                int64_t span = 0;
                for (uint32_t k = 0; k < n_cig; ++k) {
                    const uint32_t cw = le32(pc + 4 * k);
                    if (consumes_ref_op(cw & 15u)) span += cw >> 4;
                }
                if ((int64_t)rp + std::max<int64_t>(span, 1) > (int64_t)r_beg) {  // overlaps the region (bam_endpos semantics)
                    j.rec.push_back(buf + q);
                    kept = true;
                    j.n_cigar += n_cig;
                    const int64_t l_seq = (int64_t)le32(r + 16);
                    j.n_seq += (l_seq + 1) / 2;
                    j.n_qual += l_seq;
                }
                q += 4 + bs;
            }
            if (!done) {
                if (blocks.empty()) done = true;  // end of file
                else carry.assign(buf + q, buf + total);  // keep the partial record
                q = 0;
            }
            if (kept) j.arenas.push_back(std::move(arena));
            batch_blocks = std::min<size_t>(batch_blocks * 2, 4096);
        }
    };
    const int64_t n_jobs = (int64_t)jobs.size();
    if (n_jobs >= 2 * (int64_t)omp_get_max_threads()) {
#pragma omp parallel for schedule(dynamic, 1)
        for (int64_t k = 0; k < n_jobs; ++k) scan(jobs[(size_t)k], false);
    } else {
        for (int64_t k = 0; k < n_jobs; ++k) scan(jobs[(size_t)k], true);
    }
    // ---- stitch the regions together in order
    int32_t last_tid = -1, last_pos = -1;
    for (Job &j : jobs) {
        if (j.error) {
            const char *msg = j.error;
            delete b;
            return vfkio_fail(VFK_EINVAL, msg);
        }
        for (const uint8_t *r : j.rec) {
            const int32_t rt = (int32_t)le32(r + 4), rp = (int32_t)le32(r + 8);
            if (rt < last_tid || (rt == last_tid && rp < last_pos)) b->sorted = 0;
            last_tid = rt;
            last_pos = rp;
        }
        b->rec.insert(b->rec.end(), j.rec.begin(), j.rec.end());
        for (auto &a2 : j.arenas) b->arenas.push_back(std::move(a2));
        b->n_cigar += j.n_cigar;
        b->n_seq += j.n_seq;
        b->n_qual += j.n_qual;
    }
    *out = b;
    return VFK_OK;
}

extern "C" void vfkio_bam_close(vfkio_bam *b) { delete b; }
extern "C" int32_t vfkio_bam_n_contigs(const vfkio_bam *b) { return b ? (int32_t)b->names.size() : 0; }
extern "C" const char *vfkio_bam_contig_name(const vfkio_bam *b, int32_t i) {
    return (b && i >= 0 && i < (int32_t)b->names.size()) ? b->names[(size_t)i].c_str() : "";
}
extern "C" int64_t vfkio_bam_contig_length(const vfkio_bam *b, int32_t i) {
    return (b && i >= 0 && i < (int32_t)b->lengths.size()) ? b->lengths[(size_t)i] : 0;
}
extern "C" const char *vfkio_bam_header_text(const vfkio_bam *b) { return b ? b->text.c_str() : ""; }
extern "C" int32_t vfkio_bam_is_sorted(const vfkio_bam *b) { return b ? b->sorted : 0; }

extern "C" int vfkio_bam_plan(const vfkio_bam *b, vfkio_synth_plan *plan) {
    if (!b || !plan) return vfkio_fail(VFK_EINVAL, "vfkio_bam_plan: NULL argument");
    plan->n_reads = (int64_t)b->rec.size();
    plan->n_cigar = b->n_cigar;
    plan->n_seq_bytes = b->n_seq;
    plan->n_qual_bytes = b->n_qual;
    return VFK_OK;
}

// khash X31 of the NUL-terminated name and a 64-bit FNV-1a used to group equal names
static inline void name_hashes(const uint8_t *s, uint32_t l, uint32_t &x31, uint64_t &fnv) {
    x31 = l ? s[0] : 0;
    fnv = 0xCBF29CE484222325ull;
    for (uint32_t i = 0; i < l; ++i) {
        if (i) x31 = (x31 << 5) - x31 + s[i];
        fnv = (fnv ^ s[i]) * 0x100000001B3ull;
    }
}

extern "C" int vfkio_bam_fill(const vfkio_bam *b, vfkio_reads *out) {
    if (!b || !out) return vfkio_fail(VFK_EINVAL, "vfkio_bam_fill: NULL argument");
    if (!b->sorted) return vfkio_fail(VFK_EINVAL, "BAM file is not coordinate sorted");
    const int64_t n = (int64_t)b->rec.size();
    int32_t *tid = (int32_t *)out->tid, *pos = (int32_t *)out->pos, *lq = (int32_t *)out->l_qseq, *ncg = (int32_t *)out->n_cigar;
    int32_t *mtid = (int32_t *)out->mtid, *mpos = (int32_t *)out->mpos, *isz = (int32_t *)out->isize;
    uint16_t *flag = (uint16_t *)out->flag;
    uint8_t *mapq = (uint8_t *)out->mapq, *seq = (uint8_t *)out->seq, *qual = (uint8_t *)out->qual;
    int64_t *co = (int64_t *)out->cigar_off, *so = (int64_t *)out->seq_off, *qo = (int64_t *)out->qual_off;
    uint32_t *cig = (uint32_t *)out->cigar, *qh = (uint32_t *)out->qname_hash;
    uint64_t *qid = (uint64_t *)out->qname_id;
    // offsets: serial prefix sums over the record index
    int64_t c = 0, s = 0, q = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *r = b->rec[(size_t)i] + 4;
        co[i] = c; so[i] = s; qo[i] = q;
        const int64_t l_seq = (int64_t)le32(r + 16);
        c += le16(r + 12);
        s += (l_seq + 1) / 2;
        q += l_seq;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const uint8_t *r = b->rec[(size_t)i] + 4;
        tid[i] = (int32_t)le32(r);
        pos[i] = (int32_t)le32(r + 4);
        const uint32_t l_name = r[8];
        mapq[i] = r[9];
        ncg[i] = le16(r + 12);
        flag[i] = le16(r + 14);
        const int64_t l_seq = (int64_t)le32(r + 16);
        lq[i] = (int32_t)l_seq;
        mtid[i] = (int32_t)le32(r + 20);
        mpos[i] = (int32_t)le32(r + 24);
        isz[i] = (int32_t)le32(r + 28);
        const uint8_t *name = r + 32;
        name_hashes(name, l_name ? l_name - 1 : 0, qh[i], qid[i]);
        const uint8_t *pc = name + l_name;
        memcpy(cig + co[i], pc, (size_t)ncg[i] * 4);
        const uint8_t *ps = pc + (size_t)ncg[i] * 4;
        memcpy(seq + so[i], ps, (size_t)((l_seq + 1) / 2));
        memcpy(qual + qo[i], ps + (l_seq + 1) / 2, (size_t)l_seq);
    }
    out->n = n;
    out->n_contigs = (int32_t)b->names.size();
    return VFK_OK;
}

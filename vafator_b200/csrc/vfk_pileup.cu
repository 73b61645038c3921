// vfk_pileup.cu -- the pileup kernels (sm_100a).
//
// Replaces vafator/pileups.py:106-362 + vafator/pileup_utils.py:64-92: for every unit (variant
// column x ALT allele) find the overlapping reads, classify them, and reduce the per-allele
// MQ / BQ / read-position distributions to integers: counts, twice-the-median, and twice the
// ALT mid-rank sum that scipy.stats.ranksums would compute (vafator/rank_sum_test.py:11).
// Every kernel reads the canonical read arrays (include/vfk.h) and the per-read words prep_kernel derived
// from them in the same call (vfk_prep.cu); nothing has been prepared on the host.
//
// locate_kernel       : one thread per unit: 8-ary search of the sorted starts for the end of the candidate range,
//                       binary search of the block-level running maximum of the read ends + one chunk pass for its start.
// pileup_warp_kernel  : one warp per unit, lanes = candidate reads (two slots of 32).  Persistent grid, units handed out in
//                       contiguous grabs from a global counter (shrinking towards the end).  Every column with <= 64
//                       candidates -- all of 30x, three quarters of 60x: mate pairing in registers, a read's overlap
//                       partner answers by shuffle, <= 32 contributing reads are ranked by a bit-sliced (radix) comparison
//                       built on warp ballots, more go through per-warp shared-memory histograms in the same kernel.
// medium_kernel       : one warp per listed unit with 65 .. 192 candidates: slots of 32, per-read state and a name-hash table
//                       in shared memory.
// tile_kernel         : deep panels: a CTA per group of neighbouring deep units, read-major, htslib's name hash replayed in
//                       shared memory, a pair of histograms per column in shared memory.
// tail_kernel         : what is left on the lists (deeper than a tile): global mate linking of the listed windows, one warp
//                       per unit with u16 histograms, one CTA per unit with u32 counters beyond 32767 candidates; one
//                       cooperative launch.
// The three list-path kernels leave at their first line when nothing was listed for them (every 30x call).
#include <algorithm>
#include <cstdlib>
#include <cooperative_groups.h>
#include "vfk_link.cuh"

namespace cg = cooperative_groups;

namespace vfk {

#ifndef VFK_MIN_CTAS
#define VFK_MIN_CTAS 4
#endif
constexpr int WARPS_PER_CTA = 8;
#ifndef VFK_GRAB
#define VFK_GRAB 8
#endif
// Units a warp takes per visit to the global counter (<= 32).  A launch hands every warp only ~20 units, so
// the grabs shrink towards the end of the unit list (guided self-scheduling): GRAB while more than 16 units
// per warp are left, 2 below that, 1 for the last 4 per warp -- the tail is then a unit, not a grab, long.
constexpr int GRAB = VFK_GRAB;
constexpr int WARP_PATH_MAX_CAND = 32767;  // u16 bins, weight <= 2 per read

template <int POSB>
struct Layout {
    static constexpr int MQ_OFF = 0;
    static constexpr int BQ_OFF = 256;
    static constexpr int POS_OFF = 512;
    static constexpr int BINS = 512 + POSB;  // per slot
    static constexpr int POS_BITS = POSB == 256 ? 8 : POSB == 512 ? 9 : POSB == 1024 ? 10 : 12;
};

// ------------------------------------------------------------------------------------------
// locate pass: one THREAD per unit.  A search is a chain of dependent loads; with one thread per
// unit every unit's chain is in flight at once, which a warp-cooperative search cannot offer.
// Emits the 32-byte ResolvedUnit records the pileup kernels consume.
//   hi = first read of the contig starting beyond the column
//   lo = back[hi-1] = first read whose running-maximum end exceeds start[hi-1] (<= col):
//        no earlier read can overlap the column
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ ResolvedUnit locate_unit(const ReadsDev &R, const UnitsDev &V, int64_t u) {
    const int32_t tid = __ldg(V.tid + u);
    const int32_t col = __ldg(V.pos + u);
    const bool known = tid >= 0 && tid < R.n_contigs;  // an unknown contig gets an empty read range
    const int64_t c0 = known ? __ldg(R.contig_off + tid) : 0, c1 = known ? __ldg(R.contig_off + tid + 1) : 0;
    // first read of the contig starting beyond the column, in [a, b]: an 8-ary search -- seven independent probes
    // per step, so the chain of dependent loads is 8 long for a 14 M-read batch instead of 24
    int64_t a = c0, b = c1;
    while (b - a > 8) {
        const int64_t step = (b - a) >> 3;
        int32_t v[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) v[j] = __ldg(R.pos + a + (j + 1) * step);
        int c = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) c += v[j] <= col ? 1 : 0;  // sorted: the first c probes start at or before col
        const int64_t a0 = a;
        if (c < 7) b = a0 + (c + 1) * step;
        if (c > 0) a = a0 + c * step + 1;
    }
    {
        int c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (a + j < b) c += __ldg(R.pos + a + j) <= col ? 1 : 0;
        a += c;
    }
    // lo = first read of the contig whose end exceeds the column (no earlier read can overlap it).  blk_incl[b] is the
    // maximum of tid << 32 | end over the reads of blocks 0 .. b-1: a binary search names the first block that holds
    // such a read, the chunk maxima of that block name the chunk (32 reads), one pass over the chunk finds the read.  The
    // chunk of read a-1 may also hold reads starting beyond the column (their ends exceed it trivially), so it is scanned
    // directly instead of being judged by a maximum.
    int64_t lo = a;
    if (a > c0) {
        const long long key = ((long long)tid << 32) | (long long)(uint32_t)col;
        const int64_t b_first = c0 / PREP_BLOCK, b_last = (a - 1) / PREP_BLOCK;
        int64_t ch = b_last * (PREP_BLOCK / 32);  // first chunk to look at
        if (b_last > b_first && __ldg(R.blk_incl + b_last) > key) {
            int64_t ba = b_first, bb = b_last - 1;  // smallest block with an inclusive maximum beyond the key
            while (ba < bb) {
                const int64_t m = (ba + bb) >> 1;
                if (__ldg(R.blk_incl + m + 1) > key) bb = m; else ba = m + 1;
            }
            ch = ba * (PREP_BLOCK / 32);
        }
        ch = max(ch, c0 >> 5);
        const int64_t ch_last = (a - 1) >> 5;
        // a chunk before the last one qualifies by its maximum (all its reads precede `a`; reads of an earlier contig carry a
        // smaller key); the last one is scanned whatever its maximum says
        while (ch < ch_last && __ldg(R.chunk_max + ch) <= key) ++ch;
        // first read of [from, to) whose end exceeds the column: the chunk's {end, shape} pairs are fetched eight reads at a time
        // (four independent 128-bit loads; a chunk starts on a 256-byte boundary of `ee`) instead of one dependent load per read
        const int64_t from = max(c0, ch << 5), to = min(a, (ch + 1) << 5);
        lo = a;
        for (int64_t g = (ch << 5); g < to && lo == a; g += 8) {
            const uint4 *q = reinterpret_cast<const uint4 *>(R.ee + g);
            const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
            const int32_t e[8] = {(int32_t)v0.x, (int32_t)v0.z, (int32_t)v1.x, (int32_t)v1.z, (int32_t)v2.x, (int32_t)v2.z, (int32_t)v3.x, (int32_t)v3.z};
#pragma unroll
            for (int j = 7; j >= 0; --j)
                if (g + j >= from && g + j < to && e[j] > col) lo = g + j;
        }
    }
    ResolvedUnit r;
    r.col = col;
    r.meta = __ldg(V.meta + u);
    r.len = __ldg(V.length + u);
    r.ins_off = __ldg(V.seq_off + u);
    r.hi = (uint32_t)a;
    r.lo = (uint32_t)lo;
    r.stop = V.stop ? __ldg(V.stop + u) : 0x7FFFFFFF;
    r.tid = tid;
    return r;
}
__device__ __forceinline__ void store_resolved(ResolvedUnit *dst, const ResolvedUnit &r) {
    uint4 *o = reinterpret_cast<uint4 *>(dst);
    const uint4 *q = reinterpret_cast<const uint4 *>(&r);
    o[0] = q[0];
    o[1] = q[1];
}

__global__ void __launch_bounds__(256) locate_kernel(ReadsDev R, UnitsDev V, ResolvedUnit *__restrict__ res) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= V.n) return;
    store_resolved(res + u, locate_unit(R, V, u));
}

// ------------------------------------------------------------------------------------------
// fetch lists (host entry point, read arrays left in place in page-locked host memory).  A column needs ONE quality byte and
// ONE sequence byte of each read that shows a base at it; read in place each costs the link a request (~3.4 ns, as much as
// 190 bytes of bulk copy) and the two make up three quarters of a call's requests.  Everything that decides WHICH bytes --
// starts, CIGARs, the position index -- is on the device already, so the call asks for them instead:
//   request_kernel (a warp per unit with <= 64 candidates -- <= 192 where the medium kernel exists --, lane = candidate): the
//   column is resolved exactly as the register / medium kernel will resolve it, and the pool offsets of the read's quality / sequence byte at the column are written -- 16 bytes per
//   (unit, candidate) pair, coalesced, straight into page-locked host memory -- at pair index base[u] + k, base[u] handed out
//   by an atomic counter (a unit that no longer fits the list keeps reading in place: base[u] = NO_FETCH);
//   the host (cudaLaunchHostFunc between the two kernels, vfk_api.cu) gathers the bytes named, two per pair;
//   one bulk copy brings them back and pileup_warp_kernel<.., FETCHED> / medium_kernel take a read's own base / quality from pair
//   base[u] + k.
// The read filter is not consulted here (flags and MAPQ stay on the host): a request is made for every covering read, a superset.
// ------------------------------------------------------------------------------------------
constexpr uint32_t NO_FETCH = 0xFFFFFFFFu;
constexpr unsigned long long NO_REQUEST = ~0ull;
__global__ void __launch_bounds__(256) request_kernel(const __grid_constant__ ReadsDev R, const ResolvedUnit *__restrict__ res, int64_t n_units,
                                                      uint32_t *__restrict__ pair_base, ulonglong2 *__restrict__ req, uint32_t *__restrict__ pair_counter,
                                                      uint32_t cap, uint32_t max_cand, int by_index, const uint32_t *__restrict__ batch_status) {
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= n_units) return;
    const uint4 rb = __ldg(reinterpret_cast<const uint4 *>(res + u) + 1);  // lo, hi, stop, tid
    const int32_t col = __ldg(&res[u].col);
    const uint32_t n_cand = rb.y - rb.x;
    if (*batch_status || n_cand == 0u || n_cand > max_cand) {
        if (lane == 0) pair_base[u] = NO_FETCH;
        return;
    }
    uint32_t b = 0u;
    if (lane == 0) b = atomicAdd(pair_counter, n_cand);
    b = __shfl_sync(0xFFFFFFFFu, b, 0);
    const bool fits = (uint64_t)b + n_cand <= (uint64_t)cap;
    if (lane == 0) pair_base[u] = fits ? b : NO_FETCH;
    for (uint32_t k = (uint32_t)lane; k < n_cand; k += 32u) {
        const uint64_t p = (uint64_t)b + k;
        if (p >= (uint64_t)cap) break;  // the host walks [0, min(total, cap)): every entry below the cap is written, request or not
        ulonglong2 rq = make_ulonglong2(NO_REQUEST, 0ull);
        if (fits) {
            const int64_t i = (int64_t)rb.x + k;
            const int32_t rp = __ldg(R.pos + i);
            const uint2 e = __ldg(R.ee + i);
            const uint4 hot = make_uint4((uint32_t)rp, (uint32_t)((int32_t)e.x - rp), shape_of_ev(e.y) << 24, (uint32_t)i);
            const ColRead c = resolve_column(R, col, i, hot, true, true);
            // by_index: the pool offsets are the caller's own arrays on the host -- the host looks them up (read index, read position);
            // else they were derived on the device (a batch with implicit offsets): the request names the two bytes outright
            if (c.pay)
                rq = by_index ? make_ulonglong2((unsigned long long)i, (unsigned long long)c.qpos)
                              : make_ulonglong2((unsigned long long)(__ldg(R.qual_off + i) + c.qpos), (unsigned long long)(__ldg(R.seq_off + i) + (c.qpos >> 1)));
        }
        req[p] = rq;
    }
}

// ------------------------------------------------------------------------------------------
// histogram access.  CT = uint16_t: two bins per 32-bit word; CT = uint32_t: one bin per word.
// ------------------------------------------------------------------------------------------
template <typename CT>
__device__ __forceinline__ void hist_add(uint32_t *slot, int bin, uint32_t w) {
    if (sizeof(CT) == 2) atomicAdd(slot + (bin >> 1), w << ((bin & 1) * 16));
    else atomicAdd(slot + bin, w);
}
template <typename CT>
__device__ __forceinline__ uint32_t hist_get(const uint32_t *slot, int bin) {
    if (sizeof(CT) == 2) return (slot[bin >> 1] >> ((bin & 1) * 16)) & 0xFFFFu;
    return slot[bin];
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// One metric (B bins starting at bin `off` of each slot): medians of both lists and the ALT
// mid-rank sum over the pooled ranking.  Executed by one full warp.
//   2*rank_sum(ALT) = sum_v alt[v] * (2*below(v) + ties(v) + 1)
template <typename CT, int B>
__device__ __forceinline__ void reduce_metric(const uint32_t *ref_slot, const uint32_t *alt_slot, int off, int lane,
                                           int &med2_ref, int &med2_alt, uint64_t &s2) {
    constexpr int BPL = B / 32;  // consecutive bins per lane
    const int b0 = off + lane * BPL;
    uint32_t tr = 0, ta = 0;
#pragma unroll 4
    for (int j = 0; j < BPL; ++j) {
        tr += hist_get<CT>(ref_slot, b0 + j);
        ta += hist_get<CT>(alt_slot, b0 + j);
    }
    const uint32_t ir = warp_incl_scan(tr, lane), ia = warp_incl_scan(ta, lane);
    const uint32_t nr = __shfl_sync(FULL, ir, 31), na = __shfl_sync(FULL, ia, 31);
    uint32_t cr = ir - tr, ca = ia - ta;  // exclusive prefixes
    // 0-based order statistics that make the median
    const uint32_t r1 = nr ? (nr - 1) >> 1 : 0, r2 = nr >> 1, a1 = na ? (na - 1) >> 1 : 0, a2 = na >> 1;
    int vr1 = -1, vr2 = -1, va1 = -1, va2 = -1;
    uint64_t acc = 0;
    if (tr | ta) {
        for (int j = 0; j < BPL; ++j) {
            const uint32_t hr = hist_get<CT>(ref_slot, b0 + j), ha = hist_get<CT>(alt_slot, b0 + j);
            const int v = lane * BPL + j;
            if (ha) acc += (uint64_t)ha * (uint64_t)(2u * (cr + ca) + hr + ha + 1u);
            if (hr) {
                if (r1 >= cr && r1 < cr + hr) vr1 = v;
                if (r2 >= cr && r2 < cr + hr) vr2 = v;
            }
            if (ha) {
                if (a1 >= ca && a1 < ca + ha) va1 = v;
                if (a2 >= ca && a2 < ca + ha) va2 = v;
            }
            cr += hr;
            ca += ha;
        }
    }
    s2 = na ? warp_sum_u64(acc) : 0;
    if (nr) {
        vr1 = __reduce_max_sync(FULL, vr1);
        vr2 = __reduce_max_sync(FULL, vr2);
        med2_ref = vr1 + vr2;
    } else {
        med2_ref = -1;
    }
    if (na) {
        va1 = __reduce_max_sync(FULL, va1);
        va2 = __reduce_max_sync(FULL, va2);
        med2_alt = va1 + va2;
    } else {
        med2_alt = -1;
    }
}

__device__ __forceinline__ void store_counts(vfk_counts *out, int dp, int n_amb, int ac_ref, int ac_alt,
                                             const int med2[3][2], const uint64_t s2[3], uint32_t status, uint32_t n_cand,
                                             uint32_t n_cover, uint32_t n_col = 0u) {
    uint4 *o = reinterpret_cast<uint4 *>(out);
    o[0] = make_uint4((uint32_t)dp, (uint32_t)n_amb, (uint32_t)ac_ref, (uint32_t)ac_alt);
    o[1] = make_uint4((uint32_t)med2[0][0], (uint32_t)med2[0][1], (uint32_t)med2[1][0], (uint32_t)med2[1][1]);
    o[2] = make_uint4((uint32_t)med2[2][0], (uint32_t)med2[2][1], status, n_cand);
    o[3] = make_uint4((uint32_t)s2[0], (uint32_t)(s2[0] >> 32), (uint32_t)s2[1], (uint32_t)(s2[1] >> 32));
    o[4] = make_uint4((uint32_t)s2[2], (uint32_t)(s2[2] >> 32), n_cover, n_col);
}

template <typename CT, int POSB>
__device__ __forceinline__ void accumulate(uint32_t *ref_slot, uint32_t *alt_slot, const Contribution &c, bool with_bq,
                                           uint32_t &status) {
    using L = Layout<POSB>;
    if (!(c.ref || c.alt)) return;
    const bool pos_ok = c.qpos < POSB;
    if (!pos_ok) status |= VFK_ST_READLEN;
    if (c.ref) {
        hist_add<CT>(ref_slot, L::MQ_OFF + c.mq, 1u);
        if (with_bq) hist_add<CT>(ref_slot, L::BQ_OFF + c.bq, 1u);
        if (pos_ok) hist_add<CT>(ref_slot, L::POS_OFF + c.qpos, 1u);
    }
    if (c.alt) {
        hist_add<CT>(alt_slot, L::MQ_OFF + c.mq, c.alt);
        if (with_bq) hist_add<CT>(alt_slot, L::BQ_OFF + c.bq, c.alt);
        if (pos_ok) hist_add<CT>(alt_slot, L::POS_OFF + c.qpos, c.alt);
    }
}

// ------------------------------------------------------------------------------------------
// shallow path: one value per lane, ranks from bit-sliced comparisons
// ------------------------------------------------------------------------------------------
// For this lane's value v (NBITS wide): which contributing lanes hold a smaller value (lt) and
// which an equal one (eq, itself included).  Walks the bits from the top; lanes that do not contribute
// hold v == 0, so every ballot is already restricted to the contributing set.  One step is a
// predicate, a ballot and three predicated logic ops (ptxas fuses the bit tests of a value into one R2P and
// lowers a step to six instructions; the compiler's own lowering of the same C++ spends eight to ten, and
// a least-significant-bit-first recurrence with two LOP3 per step costs registers the kernel does not have).
template <int BIT>
__device__ __forceinline__ void radix_step(uint32_t v, uint32_t &lt, uint32_t &eq) {
    asm("{\n"
        ".reg .pred p;\n"
        ".reg .b32 b, t;\n"
        "and.b32 t, %2, %3;\n"
        "setp.ne.u32 p, t, 0;\n"
        "vote.sync.ballot.b32 b, p, 0xffffffff;\n"
        "@p lop3.b32 %0, %0, %1, b, 0xF4;\n"  // mine set:   lt |= eq & ~b   (equal so far, their bit clear)
        "@p and.b32 %1, %1, b;\n"             //             eq &= b
        "@!p lop3.b32 %1, %1, b, 0, 0x30;\n"  // mine clear: eq &= ~b
        "}\n"
        : "+r"(lt), "+r"(eq)
        : "r"(v), "r"(1u << BIT));
}
template <int NBITS, int BIT = NBITS - 1>
__device__ __forceinline__ void radix_steps(uint32_t v, uint32_t &lt, uint32_t &eq) {
    radix_step<BIT>(v, lt, eq);
    if constexpr (BIT > 0) radix_steps<NBITS, BIT - 1>(v, lt, eq);
}
template <int NBITS>
__device__ __forceinline__ void radix_compare(uint32_t v, uint32_t contributing, uint32_t &lt, uint32_t &eq) {
    lt = 0u;
    eq = contributing;
    radix_steps<NBITS>(v, lt, eq);
}

// Medians of the REF / ALT lists and this lane's share of the ALT mid-rank sum, from the lt / eq masks.
//
// Median: ties are broken by lane id, so every member of a list owns exactly one 0-based position of
// the sorted list, pos = #smaller + #equal-in-lower-lanes.  The member(s) at the middle position(s)
// contribute value * weight (weight 2 when the count is odd, 1 + 1 when it is even) and ONE additive
// reduction returns twice the median of both lists: REF in the low half-word, ALT in the high one.
//
// Rank sum: scipy ranks the pooled sample; with b(a) / t(a) the REF values below / equal to an ALT value a,
//   2 * rank_sum(ALT) = sum_a (2 * below(a) + ties(a) + 1) = na^2 + na + sum_a (2 * b(a) + t(a))
// because the ALT-vs-ALT part of the sum is na^2 whatever the values are.  The lane returns
// 2 * b + t (<= 2 * nr; summed over na lanes <= 2 * 16 * 16 = 512 while nr + na <= 32): the three metrics'
// terms share one reduction, 10 bits each.
struct SlotShape {  // metric independent
    uint32_t R, A;          // list membership (lanes)
    uint32_t R_low, A_low;  // ... restricted to the lanes below this one
    int r1, a1;             // lower middle position
    uint32_t dr, da;        // upper middle position - lower (0: odd count, 1: even)
    uint32_t sh_r, sh_a;    // weight shift (+16 for ALT)
    bool is_ref, is_alt;
};
__device__ __forceinline__ uint32_t radix_finish(uint32_t v, uint32_t lt, uint32_t eq, const SlotShape &sh, uint32_t &med2_sum) {
    const uint32_t br = __popc(lt & sh.R), tr = __popc(eq & sh.R);
    const uint32_t pos_r = br + __popc(eq & sh.R_low), pos_a = __popc(lt & sh.A) + __popc(eq & sh.A_low);
    const bool mid_r = sh.is_ref && (uint32_t)(pos_r - (uint32_t)sh.r1) <= sh.dr;
    const bool mid_a = sh.is_alt && (uint32_t)(pos_a - (uint32_t)sh.a1) <= sh.da;
    med2_sum = __reduce_add_sync(FULL, (mid_r ? v << sh.sh_r : 0u) + (mid_a ? v << sh.sh_a : 0u));
    return sh.is_alt ? 2u * br + tr : 0u;
}

// ------------------------------------------------------------------------------------------
// deep path of the warp kernel: shared-memory histograms
// ------------------------------------------------------------------------------------------
template <int POSB>
__device__ __forceinline__ void unit_by_histogram(const ReadsDev &R, const vfk_params &P, const Unit &U, uint32_t *ref_slot,
                                               uint32_t *alt_slot, int lane, vfk_counts *out, uint32_t bad) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS / 2;
    {
        uint4 *z = reinterpret_cast<uint4 *>(ref_slot);
        for (int w = lane; w < (2 * WORDS) / 4; w += 32) z[w] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    int dp = 0, n_amb = 0, ac_ref = 0, ac_alt = 0, n_cover = 0, n_col = 0;
    uint32_t status = U.kind | bad;
    const bool with_bq = U.kind == VFK_KIND_SNV;
    {
        for (int64_t b = U.lo; b < U.hi; b += 32) {
            const int64_t i = b + lane;
            if (i < U.hi) {
                const Contribution c = evaluate(R, P, U, i);
                n_cover += c.covers ? 1 : 0;
                n_col += c.in_col ? 1 : 0;
                dp += c.in_dp ? 1 : 0;
                n_amb += c.ambiguous ? 1 : 0;
                ac_ref += c.ref ? 1 : 0;
                ac_alt += (int)c.alt;
                accumulate<uint16_t, POSB>(ref_slot, alt_slot, c, with_bq, status);
            }
        }
    }
    __syncwarp();
    dp = __reduce_add_sync(FULL, dp);
    n_amb = __reduce_add_sync(FULL, n_amb);
    ac_ref = __reduce_add_sync(FULL, ac_ref);
    ac_alt = __reduce_add_sync(FULL, ac_alt);
    n_cover = __reduce_add_sync(FULL, n_cover);
    n_col = __reduce_add_sync(FULL, n_col);
    status = __reduce_or_sync(FULL, status);
    int med2[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};
    uint64_t s2[3] = {0, 0, 0};
    if (ac_ref | ac_alt) {
        reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::MQ_OFF, lane, med2[0][0], med2[0][1], s2[0]);
        if (with_bq) reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::BQ_OFF, lane, med2[1][0], med2[1][1], s2[1]);
        reduce_metric<uint16_t, POSB>(ref_slot, alt_slot, L::POS_OFF, lane, med2[2][0], med2[2][1], s2[2]);
    }
    if (U.kind == VFK_KIND_SNV && !P.include_ambiguous) dp -= n_amb;  // pileups.py:224-227
    if (lane == 0) store_counts(out, dp, n_amb, ac_ref, ac_alt, med2, s2, status, (uint32_t)(U.hi - U.lo), (uint32_t)n_cover,
                                (uint32_t)n_col);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// warp per unit
// ------------------------------------------------------------------------------------------
// One candidate read as a lane holds it: the assembled header (vfk_device.cuh load_hot) and, after the pairing step,
// its overlap partner word.
struct Cand {
    uint4 hot;
    uint32_t mw;
    uint32_t st;  // pairing state: ST_STORES = the record reached the name hash and stored itself (its mate still to come)
};
constexpr uint32_t ST_STORES = 1u;

// htslib's mate-overlap pairing (sam.c overlap_push) for the candidate reads of ONE column, in registers.
// Lane k holds candidate k (chunk A) and candidate 32 + k (chunk B).  A record reaches the name hash when it passes the
// read filter, is a proper pair with a mapped mate on the same contig and is not "obviously non-overlapping"; two such
// records with the same name pair up iff the earlier one stored itself (its mate still to come).  Names are matched with
// MATCH.ANY inside a chunk and by a short loop across the chunks.  Returns false when the column needs the general
// protocol (three records of one name, a zero-span record): the unit then goes to the list path, whose link kernels
// replay the hash in full.
// passA / passB: the slot holds a read that passes the read filter.
__device__ __forceinline__ bool pair_in_registers(const ReadsDev &R, const vfk_params &P, int32_t tid, int lane, uint32_t lt_mask,
                                                  bool passA, bool passB, int64_t iA, int64_t iB, Cand &A, Cand &B) {
    A.mw = VFK_NO_MATE;
    B.mw = VFK_NO_MATE;
    A.st = B.st = 0u;
    if (!P.ignore_overlaps) return true;
    const bool preA = passA && (A.hot.z & (F_PROPER | F_MUNMAP)) == F_PROPER, preB = passB && (B.hot.z & (F_PROPER | F_MUNMAP)) == F_PROPER;
    unsigned long long idA = 0ull, idB = 0ull;
    int32_t mposA = 0, mposB = 0;
    bool eligA = false, eligB = false, zero = false;
    auto reaches = [&](int64_t i, const uint4 &h, unsigned long long &id, int32_t &mp) {
        id = __ldg(R.qname_id + i);
        mp = __ldg(R.mpos + i);
        const int32_t mt = __ldg(R.mtid + i);
        if (mt >= 0 && mt != tid) return false;
        const int32_t end = (int32_t)h.x + (int32_t)h.y;
        if (mp >= end) {
            const long long isz = (long long)__ldg(R.isize + i);
            const int lq = (h.z >> 24) == SHAPE_SIMPLE ? (int)h.y : __ldg(R.l_qseq + i);
            if ((isz < 0 ? -isz : isz) >= 2LL * lq) return false;
        }
        return true;
    };
    if (preA) { eligA = reaches(iA, A.hot, idA, mposA); zero = eligA && A.hot.y == 0u; }
    if (preB) { eligB = reaches(iB, B.hot, idB, mposB); zero = zero || (eligB && B.hot.y == 0u); }
    const uint32_t mA = __ballot_sync(FULL, eligA), mB = __ballot_sync(FULL, eligB);
    if (__any_sync(FULL, zero)) return false;  // bam_plp_push may not append it: push order decides
    uint32_t gA = 0u, gB = 0u, xA = 0u, xB = 0u;  // same-name lanes: own chunk (self included) / other chunk
    // (all lanes vote: a lane that is out holds a private key and is masked off the result)
    gA = __match_any_sync(FULL, eligA ? idA : (unsigned long long)lane) & mA;
    if (!eligA) gA = 0u;
    if (mB) {
        gB = __match_any_sync(FULL, eligB ? idB : (unsigned long long)lane) & mB;
        if (!eligB) gB = 0u;
        // every chunk-B record is looked up in chunk A (a few per column at 30x): its partner, or a third record of its name
        uint32_t todo = mB;
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1u;
            const unsigned long long idj = __shfl_sync(FULL, idB, j);
            const uint32_t m = __ballot_sync(FULL, eligA && idA == idj);
            if (lane == j) xB = m;
            if (eligA && idA == idj) xA |= 1u << j;
        }
    }
    const int sizeA = __popc(gA) + __popc(xA), sizeB = __popc(gB) + __popc(xB);
    if (__any_sync(FULL, sizeA > 2 || sizeB > 2)) return false;
    // partner of a size-2 group; the earlier record (chunk A before chunk B, lower lane first) must have stored itself:
    // mpos >= pos, or a paired record whose mate position is unset
    const bool storesA = eligA && (mposA >= (int32_t)A.hot.x || ((A.hot.z & F_PAIRED) && mposA == -1));
    const bool storesB = eligB && (mposB >= (int32_t)B.hot.x || ((B.hot.z & F_PAIRED) && mposB == -1));
    const uint32_t sA = __ballot_sync(FULL, storesA), sB = __ballot_sync(FULL, storesB);
    const uint32_t me = 1u << lane;
    A.st = storesA ? ST_STORES : 0u;
    B.st = storesB ? ST_STORES : 0u;
    // name hashes of the two records are equal, so either one gives htslib's tie-break bit
    if (sizeA == 2) {
        if (xA) {  // partner in chunk B: this record is the earlier one
            const int pl = __ffs(xA) - 1;
            if (storesA) A.mw = ((uint32_t)(iA - lane) + 32u + (uint32_t)pl) | ((wang_hash(__ldg(R.qname_hash + iA)) & 1u) << 31);
        } else {
            const uint32_t other = gA & ~me;
            const int pl = __ffs(other) - 1;
            const bool first_stores = pl > lane ? storesA : ((sA >> pl) & 1u);
            if (first_stores) A.mw = ((uint32_t)(iA - lane) + (uint32_t)pl) | ((wang_hash(__ldg(R.qname_hash + iA)) & 1u) << 31);
        }
    }
    if (sizeB == 2) {
        if (xB) {  // partner in chunk A: that record is the earlier one
            const int pl = __ffs(xB) - 1;
            if ((sA >> pl) & 1u) B.mw = ((uint32_t)(iB - 32 - lane) + (uint32_t)pl) | ((wang_hash(__ldg(R.qname_hash + iB)) & 1u) << 31);
        } else {
            const uint32_t other = gB & ~me;
            const int pl = __ffs(other) - 1;
            const bool first_stores = pl > lane ? storesB : ((sB >> pl) & 1u);
            if (first_stores) B.mw = ((uint32_t)(iB - lane) + (uint32_t)pl) | ((wang_hash(__ldg(R.qname_hash + iB)) & 1u) << 31);
        }
    }
    return true;
}

// A unit whose contributions do not fit the one-value-per-lane ranking (more than 32 contributing reads -- every 60x column --,
// an insertion one read supports more than once, REF == ALT): the per-read results are already in registers, only the
// reduction changes -- per-warp shared-memory histograms, no second evaluation, no second kernel.
template <int POSB>
__device__ __noinline__ void reduce_by_histogram(Eval e0, Eval e1, uint32_t kind, int include_ambiguous, uint32_t *ref_slot, uint32_t *alt_slot, int lane,
                                              vfk_counts *out, uint32_t n_cand) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS / 2;
    {
        uint4 *z = reinterpret_cast<uint4 *>(ref_slot);
        for (int w = lane; w < (2 * WORDS) / 4; w += 32) z[w] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    uint32_t status = kind;
    const bool with_bq = kind == VFK_KIND_SNV;
    const Contribution c0 = decode(e0), c1 = decode(e1);
    accumulate<uint16_t, POSB>(ref_slot, alt_slot, c0, with_bq, status);
    accumulate<uint16_t, POSB>(ref_slot, alt_slot, c1, with_bq, status);
    __syncwarp();
    const uint32_t tally = __reduce_add_sync(FULL, e0.fl + e1.fl);  // four counters (each <= 64) in the byte lanes of one word
    const int n_cover = (int)(tally & 0xFFu), n_col = (int)((tally >> 8) & 0xFFu), n_amb = (int)(tally >> 24);
    int dp = (int)((tally >> 16) & 0xFFu);
    const int ac_ref = __reduce_add_sync(FULL, (c0.ref ? 1 : 0) + (c1.ref ? 1 : 0));
    const int ac_alt = __reduce_add_sync(FULL, (int)(c0.alt + c1.alt));
    status = __reduce_or_sync(FULL, status);
    int med2[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};
    uint64_t s2[3] = {0, 0, 0};
    if (ac_ref | ac_alt) {
        reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::MQ_OFF, lane, med2[0][0], med2[0][1], s2[0]);
        if (with_bq) reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::BQ_OFF, lane, med2[1][0], med2[1][1], s2[1]);
        reduce_metric<uint16_t, POSB>(ref_slot, alt_slot, L::POS_OFF, lane, med2[2][0], med2[2][1], s2[2]);
    }
    if (kind == VFK_KIND_SNV && !include_ambiguous) dp -= n_amb;  // pileups.py:224-227
    if (lane == 0) store_counts(out, dp, n_amb, ac_ref, ac_alt, med2, s2, status, n_cand, (uint32_t)n_cover, (uint32_t)n_col);
    __syncwarp();
}

// The base-quality filter of a read that sits in a deletion / reference skip at the column looks at its NEXT base; whether that
// base has been overlap-adjusted depends on the push timing of the mate.  Rare (a few units in a hundred have such
// a read): kept out of line so that it costs the common path no registers.
static __device__ __noinline__ int hard_quality(const ReadsDev &R, const vfk_params &P, int32_t col, int32_t stop, int32_t tid, int64_t hi, int64_t i,
                                         uint32_t mw, int qpos, uint32_t own_q, uint32_t own_s) {
    Unit U;
    U.col = col; U.stop = stop; U.tid = tid; U.hi = hi; U.lo = 0; U.kind = 0u; U.ref_code = U.alt_code = 0u; U.len = 0; U.ins = nullptr;
    PayLoad own;
    own.q = own_q; own.s = own_s; own.odd = (uint32_t)qpos & 1u;
    const uint4 hot = load_hot(R, i), cold = load_cold(R, i);
    uint32_t my_word;
    return tweaked_quality(R, P, U, i, hot, mw, cold, true, qpos, (int)cold.z, false, own, my_word);
}

// ------------------------------------------------------------------------------------------
// medium columns: 65 .. MED_CAP candidate reads (every fourth 60x column).  Same protocol as the two register slots, for as
// many slots of 32 as the column has, with the per-read state parked in shared memory between the passes:
//   pass 1  read filter, column resolution, the read's own base / quality, overlap_push eligibility -> key / cp / xw
//   pairing the eligible reads meet in a small hash table keyed by the X31 name hash (names confirmed by their 64-bit ids): the
//           first record of a name claims a slot, the second finds it and the two name each other, a third makes the column
//           one for the general protocol (list), exactly as in pair_in_registers
//   pass 2  overlap adjustment from the partner's parked word, base-quality filter, classification, histograms
// ------------------------------------------------------------------------------------------
constexpr int MED_CAP = 192, MED_TAB = 512;
constexpr uint32_t MED_NONE = 0xFFFFu;
struct MedSmem {
    uint32_t key[MED_CAP];   // X31 name hash of an eligible read
    uint32_t cp[MED_CAP];    // packed per-read state (bits below)
    uint16_t xw[MED_CAP];    // quality | base << 8 of the read at the column (or at its next base), bit 15: aligned at the column
    uint16_t prt[MED_CAP];   // partner (candidate index) or MED_NONE
    uint16_t tab[MED_TAB];   // name hash table: candidate index or MED_NONE
    uint32_t general;        // the column needs the general protocol
    uint32_t pad[3];
};
static_assert(sizeof(MedSmem) % 16 == 0, "per-warp blocks stay 16-byte aligned");
constexpr uint32_t CP_QPOS = 0xFFFu, CP_COVERS = 1u << 20, CP_IN_COL = 1u << 21, CP_DEL = 1u << 22, CP_SKIP = 1u << 23, CP_ALIGNED = 1u << 24,
                   CP_HARD = 1u << 25, CP_INS = 1u << 26, CP_DELN = 1u << 27, CP_NONSIMPLE = 1u << 28, CP_ELIG = 1u << 29, CP_STORES = 1u << 30;

// compare-and-swap on a 16-bit cell of shared memory, through its 32-bit word
__device__ __forceinline__ uint32_t cas16(uint16_t *cell, uint32_t expect, uint32_t val) {
    uint32_t *word = reinterpret_cast<uint32_t *>(reinterpret_cast<uintptr_t>(cell) & ~(uintptr_t)3);
    const uint32_t sh = (reinterpret_cast<uintptr_t>(cell) & 2) ? 16u : 0u;
    uint32_t seen = *reinterpret_cast<volatile uint32_t *>(word);
    for (;;) {
        const uint32_t cur = (seen >> sh) & 0xFFFFu;
        if (cur != expect) return cur;
        const uint32_t old = atomicCAS(word, seen, (seen & ~(0xFFFFu << sh)) | (val << sh));
        if (old == seen) return cur;
        seen = old;
    }
}

template <int POSB>
static __device__ __forceinline__ bool medium_unit(const ReadsDev &R, const vfk_params &P, int32_t col, uint32_t meta, int32_t len, const uint8_t *ins,
                                                uint32_t lo, uint32_t hi, int32_t stop, int32_t tid, MedSmem *M, uint32_t *ref_slot, uint32_t *alt_slot,
                                                vfk_counts *out, uint32_t fb, const uint16_t *__restrict__ fetched) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS / 2;
    const int lane = threadIdx.x & 31;
    const int n = (int)(hi - lo);
    Unit U;
    U.col = col; U.stop = stop; U.kind = meta & 3u; U.ref_code = (meta >> 8) & 0xFFu; U.alt_code = (meta >> 16) & 0xFFu; U.len = len; U.ins = ins;
    U.tid = tid; U.lo = lo; U.hi = hi;
    for (int k = lane; k < MED_TAB / 2; k += 32) reinterpret_cast<uint32_t *>(M->tab)[k] = 0xFFFFFFFFu;
    if (lane == 0) M->general = 0u;
    {
        uint4 *z = reinterpret_cast<uint4 *>(ref_slot);
        for (int w = lane; w < (2 * WORDS) / 4; w += 32) z[w] = make_uint4(0, 0, 0, 0);
    }
    // ---- pass 1
    bool zero = false;
    for (int k = lane; k < n; k += 32) {
        const int64_t i = (int64_t)lo + k;
        const uint4 hot = load_hot(R, i);
        const bool pass = read_passes(hot.z, P);
        const ColRead c = resolve_column(R, col, i, hot, true, pass);
        PayLoad own;
        own.q = own.s = own.odd = 0u;
        if (c.pay) own = fb != NO_FETCH ? pay_fetched(fetched, fb + (uint32_t)k, (uint32_t)c.qpos) : pay_issue(R, hot.w, (uint32_t)c.qpos);  // (fetch list of the host entry point)
        bool elig = P.ignore_overlaps && pass && (hot.z & (F_PROPER | F_MUNMAP)) == F_PROPER, stores = false;
        uint32_t key = 0u;
        if (elig) {  // htslib overlap_push past its early exits (cf. pair_in_registers)
            const int32_t mp = __ldg(R.mpos + i), mt = __ldg(R.mtid + i);
            key = __ldg(R.qname_hash + i);
            if (mt >= 0 && mt != tid) elig = false;
            const int32_t end = (int32_t)hot.x + (int32_t)hot.y;
            if (elig && mp >= end) {
                const long long isz = (long long)__ldg(R.isize + i);
                const int lq = (hot.z >> 24) == SHAPE_SIMPLE ? (int)hot.y : __ldg(R.l_qseq + i);
                if ((isz < 0 ? -isz : isz) >= 2LL * lq) elig = false;
            }
            zero = zero || (elig && hot.y == 0u);
            stores = elig && (mp >= (int32_t)hot.x || ((hot.z & F_PAIRED) && mp == -1));
        }
        const uint32_t word = pay_word(own);
        M->key[k] = key;
        M->prt[k] = (uint16_t)MED_NONE;
        M->xw[k] = (uint16_t)((c.pay ? (word & 0xFFFu) : 0u) | (c.aligned ? 0x8000u : 0u));
        M->cp[k] = ((uint32_t)c.qpos & CP_QPOS) | (((hot.z >> 16) & 0xFFu) << 12) | ((c.fl & FL_COVERS) ? CP_COVERS : 0u) | ((c.fl & FL_IN_COL) ? CP_IN_COL : 0u) |
                   (c.is_del ? CP_DEL : 0u) | (c.is_refskip ? CP_SKIP : 0u) | (c.aligned ? CP_ALIGNED : 0u) | (c.hard ? CP_HARD : 0u) |
                   (c.indel > 0 ? CP_INS : 0u) | (c.indel < 0 ? CP_DELN : 0u) | ((hot.z >> 24) != SHAPE_SIMPLE ? CP_NONSIMPLE : 0u) |
                   (elig ? CP_ELIG : 0u) | (stores ? CP_STORES : 0u);
    }
    if (__any_sync(FULL, zero)) return false;  // bam_plp_push may not append a zero-span record: push order decides
    __syncwarp();
    // ---- pairing
    for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        const bool elig = k < n && (M->cp[k < n ? k : 0] & CP_ELIG);
        const uint32_t key = M->key[k < n ? k : 0];
        uint32_t slot = (key * 2654435761u) >> 23;  // 9 bits
        bool open = elig;
        for (int probe = 0; probe < MED_TAB && __any_sync(FULL, open); ++probe) {
            if (open) {
                const uint32_t old = cas16(&M->tab[slot], MED_NONE, (uint32_t)k);
                if (old == MED_NONE) {
                    open = false;  // first record of its name
                } else if (M->key[old] == key && __ldg(R.qname_id + lo + old) == __ldg(R.qname_id + lo + k)) {
                    if (cas16(&M->prt[old], MED_NONE, (uint32_t)k) != MED_NONE) M->general = 1u;  // a third record of the name
                    else M->prt[k] = (uint16_t)old;
                    open = false;
                } else {
                    slot = (slot + 1u) & (MED_TAB - 1);
                }
            }
        }
    }
    __syncwarp();
    const bool general = __any_sync(FULL, M->general != 0u);  // a third record of some name: the column goes to the list
    // ---- pass 2
    int dp = 0, n_amb = 0, ac_ref = 0, ac_alt = 0, n_cover = 0, n_col = 0;
    uint32_t status = U.kind;
    const bool with_bq = U.kind == VFK_KIND_SNV;
    long long nx = -2;  // the first record pushed after the column, looked up once if a read needs it
    for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        const bool valid = k < n;
        const int64_t i = (int64_t)lo + k;
        const uint32_t cp = valid ? M->cp[k] : 0u;
        const uint32_t pj = valid ? (uint32_t)M->prt[k] : MED_NONE;
        // two records of one name pair up iff the earlier one stored itself
        const bool paired = pj != MED_NONE && (M->cp[min((uint32_t)k, pj)] & CP_STORES);  // (pj != MED_NONE implies a valid k)
        const uint32_t xw = valid ? (uint32_t)M->xw[k] : 0u;
        const uint32_t word = xw & 0xFFFu;
        const uint32_t sel = wang_hash(valid ? M->key[k] : 0u) & 1u;
        uint32_t mw = paired ? (((uint32_t)lo + pj) | (sel << 31)) : VFK_NO_MATE;
        int q = 0;
        uint32_t my_word = 0u;
        if (cp & CP_ALIGNED) {
            const uint32_t px = paired ? (uint32_t)M->xw[pj] : 0u;
            my_word = word;
            q = overlap_quality(word, (px & 0x8000u) ? ((px & 0xFFFu) | 0x80000000u) : 0u, (uint32_t)k < pj, sel);
        }
        const bool hard = cp & CP_HARD;
        if (__any_sync(FULL, hard)) {
            // a stored read without a partner among the candidates: the record after the column may be its mate (bam_plp_auto reads one ahead)
            const bool t = hard && !paired && (cp & CP_STORES);
            if (__any_sync(FULL, t)) {
                if (nx == -2) {
                    long long v = -1;
                    if (lane == 0) {
                        v = next_pushed(R, P, U.hi, contig_end(R, U.tid), U.stop);
                        if (v >= 0 && !(ld_end(R, v) > __ldg(R.pos + v) && reaches_hash(R, P, v, U.tid, INT32_MIN))) v = -1;
                    }
                    nx = __shfl_sync(FULL, v, 0);
                }
                if (t && nx >= 0 && __ldg(R.qname_id + i) == __ldg(R.qname_id + nx) && __ldg(R.qname_hash + i) == __ldg(R.qname_hash + nx))
                    mw = (uint32_t)nx | ((wang_hash(__ldg(R.qname_hash + nx)) & 1u) << 31);
            }
            if (hard) {
                my_word = word;
                q = hard_quality(R, P, U.col, U.stop, U.tid, U.hi, i, mw, (int)(cp & CP_QPOS), word & 0xFFu, (word >> 8) * 0x11u);  // the base in both nibbles
            }
        }
        ColRead c;
        c.fl = ((cp & CP_COVERS) ? FL_COVERS : 0u) | ((cp & CP_IN_COL) ? FL_IN_COL : 0u);
        c.aligned = cp & CP_ALIGNED; c.hard = hard; c.is_del = cp & CP_DEL; c.is_refskip = cp & CP_SKIP; c.pay = false;
        c.qpos = (int)(cp & CP_QPOS); c.indel = (cp & CP_INS) ? 1 : (cp & CP_DELN) ? -1 : 0;
        uint4 hot = make_uint4(0u, 0u, ((cp >> 12) & 0xFFu) << 16 | ((cp & CP_NONSIMPLE) ? (SHAPE_GENERAL << 24) : 0u), 0u);  // what finish_column reads of it
        const Eval e = finish_column(R, P, U, i, hot, c, q, my_word);
        const Contribution cb = decode(e);
        n_cover += cb.covers ? 1 : 0;
        n_col += cb.in_col ? 1 : 0;
        dp += cb.in_dp ? 1 : 0;
        n_amb += cb.ambiguous ? 1 : 0;
        ac_ref += cb.ref ? 1 : 0;
        ac_alt += (int)cb.alt;
        accumulate<uint16_t, POSB>(ref_slot, alt_slot, cb, with_bq, status);
    }
    __syncwarp();
    dp = __reduce_add_sync(FULL, dp);
    n_amb = __reduce_add_sync(FULL, n_amb);
    ac_ref = __reduce_add_sync(FULL, ac_ref);
    ac_alt = __reduce_add_sync(FULL, ac_alt);
    n_cover = __reduce_add_sync(FULL, n_cover);
    n_col = __reduce_add_sync(FULL, n_col);
    status = __reduce_or_sync(FULL, status);
    int med2[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};
    uint64_t s2[3] = {0, 0, 0};
    if (ac_ref | ac_alt) {
        reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::MQ_OFF, lane, med2[0][0], med2[0][1], s2[0]);
        if (with_bq) reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::BQ_OFF, lane, med2[1][0], med2[1][1], s2[1]);
        reduce_metric<uint16_t, POSB>(ref_slot, alt_slot, L::POS_OFF, lane, med2[2][0], med2[2][1], s2[2]);
    }
    if (U.kind == VFK_KIND_SNV && !P.include_ambiguous) dp -= n_amb;  // pileups.py:224-227
    if (lane == 0 && !general) store_counts(out, dp, n_amb, ac_ref, ac_alt, med2, s2, status, (uint32_t)n, (uint32_t)n_cover, (uint32_t)n_col);
    __syncwarp();
    return !general;
}

// one warp per listed medium unit; a unit that turns out to need the general protocol moves on to the tail kernel's list
template <int POSB>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
medium_kernel(const __grid_constant__ ReadsDev R, const ResolvedUnit *__restrict__ res, const uint8_t *__restrict__ ins_seq,
              const __grid_constant__ vfk_params P, vfk_counts *__restrict__ out, const uint32_t *__restrict__ medium_list,
              const uint32_t *__restrict__ n_medium, uint32_t *__restrict__ warp_list, uint32_t *__restrict__ n_warp,
              const uint32_t *__restrict__ fetch_base, const uint16_t *__restrict__ fetched) {
    using L = Layout<POSB>;
    extern __shared__ __align__(16) uint32_t dyn_smem[];  // per warp: a pair of u16 histograms, then MedSmem
    constexpr int PER_WARP = L::BINS + (int)(sizeof(MedSmem) / 4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *hist = dyn_smem + (size_t)warp * PER_WARP;
    MedSmem *med = reinterpret_cast<MedSmem *>(hist + L::BINS);
    const uint32_t n = *n_medium;  // written by the register kernel (an earlier launch)
    for (uint32_t k = blockIdx.x * WARPS_PER_CTA + warp; k < n; k += gridDim.x * WARPS_PER_CTA) {
        const uint32_t u = medium_list[k];
        const ResolvedUnit ru = res[u];
        const uint32_t fb = fetch_base ? __ldg(fetch_base + u) : NO_FETCH;
        const bool ok = medium_unit<POSB>(R, P, ru.col, ru.meta, ru.len, ins_seq + ru.ins_off, ru.lo, ru.hi, ru.stop, ru.tid, med, hist, hist + L::BINS / 2, out + u,
                                          fb, fetched);
        if (!ok && lane == 0) warp_list[atomicAdd(n_warp, 1u)] = u;  // its record keeps VFK_ST_DEFERRED
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// deep panels: neighbouring deep columns share their reads (1000x amplicon: a read covers ~6 variant columns, the ten columns of
// an amplicon see the same ~1 800 reads).  A CTA takes a TILE -- up to TILE_UNITS consecutive deep units whose windows fit
// TILE_READS reads together -- and evaluates read-major: a read's header and pairing fields are fetched ONCE for all the
// columns it covers, its base / quality bytes come from the same one or two cache lines, the reads are paired ONCE per tile in
// a shared-memory name hash, and every column of the tile has its pair of histograms in shared memory.
//   A  per read: header -> smem, read filter, overlap_push eligibility, name hash
//   B  pairing: htslib's name hash replayed over the tile's reads in shared memory (the general protocol: a deep column nearly
//      always holds a supplementary record, i.e. three records of one name)
//   C  per (read, column): column resolution, the read's own base / quality -> xw[column][read]
//   D  per (read, column): overlap adjustment from the partner's xw, base-quality filter, classification, histograms
//   E  per column: medians / rank sums from the histograms, the unit's record
// ------------------------------------------------------------------------------------------
constexpr int TILE_READS = 2048, TILE_UNITS = 16, TILE_TAB = 4096, TILE_THREADS = 1024;  // one CTA per SM (shared memory): 32 warps of it
constexpr uint32_t RS_PASS = 1u, RS_ELIG = 2u, RS_STORES = 4u;
template <int POSB>
struct TileSmem {
    uint4 hot[TILE_READS];
    uint32_t key[TILE_READS];
    uint16_t prt[TILE_READS];
    uint16_t nxt[TILE_READS];   // chain of the records that share a name-hash bucket
    int32_t pcol[TILE_READS];   // push column of a chained record (bam_plp_push: the start of the record pushed just before it)
    uint16_t tab[TILE_TAB];     // bucket -> last chained record
    uint8_t rs[TILE_READS];
    uint16_t xw[TILE_UNITS][TILE_READS];
    uint32_t hist[TILE_UNITS][Layout<POSB>::BINS];  // per column: REF slot then ALT slot of BINS / 2 words (u16 bins)
    int cnt[TILE_UNITS][8];                          // n_cover, n_col, dp, n_amb, ac_ref, ac_alt, status
    int med2[TILE_UNITS][3][2];
    unsigned long long s2[TILE_UNITS][3];
    ResolvedUnit ru[TILE_UNITS];
    uint32_t unit[TILE_UNITS];                       // unit index of column j of the tile
    uint32_t deep[TILE_UNITS];                       // group scan: unit k of the group is a deep, deferred unit
    uint32_t n_cols, t_lo, t_hi, next_k;
};

template <int POSB>
__global__ void __launch_bounds__(TILE_THREADS, 1)
tile_kernel(const __grid_constant__ ReadsDev R, const ResolvedUnit *__restrict__ res, int64_t n_units, const uint8_t *__restrict__ ins_seq,
            const __grid_constant__ vfk_params P, vfk_counts *out, const uint32_t *__restrict__ n_deep, uint32_t *__restrict__ n_tiled) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS / 2;
    if (*n_deep == 0u) return;  // written by the register kernel (an earlier launch): a 30x / 60x call leaves here
    extern __shared__ __align__(16) uint8_t tile_raw[];
    TileSmem<POSB> &S = *reinterpret_cast<TileSmem<POSB> *>(tile_raw);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t n_groups = (n_units + TILE_UNITS - 1) / TILE_UNITS;
    for (int64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int64_t u0 = g * TILE_UNITS;
        __syncthreads();
        if (t < TILE_UNITS) {
            const int64_t u = u0 + t;
            bool deep = false;
            if (u < n_units) {
                const uint4 w = reinterpret_cast<const uint4 *>(out + u)[2];  // {med2[2][0], med2[2][1], status, n_cand}
                deep = (w.z & VFK_ST_DEFERRED) && w.w > 64u && w.w <= (uint32_t)TILE_READS;
                if (deep) S.ru[t] = res[u];
            }
            S.deep[t] = deep ? 1u : 0u;
        }
        if (t == 0) S.next_k = 0u;
        __syncthreads();
        for (;;) {  // the group's deep units, one tile after the other
            if (t == 0) {  // next tile: consecutive deep units of one contig while their windows fit TILE_READS reads together
                uint32_t k = S.next_k, n_cols = 0u, lo = 0u, hi = 0u;
                int32_t tid = 0;
                while (k < (uint32_t)TILE_UNITS && !S.deep[k]) ++k;
                for (; k < (uint32_t)TILE_UNITS; ++k) {
                    if (!S.deep[k]) continue;
                    const ResolvedUnit &r = S.ru[k];
                    if (n_cols == 0u) { lo = r.lo; hi = r.hi; tid = r.tid; }
                    else if (r.tid != tid || max(hi, r.hi) - min(lo, r.lo) > (uint32_t)TILE_READS) break;
                    lo = min(lo, r.lo); hi = max(hi, r.hi);
                    S.unit[n_cols++] = k;
                }
                S.next_k = k; S.n_cols = n_cols; S.t_lo = lo; S.t_hi = hi;
            }
            __syncthreads();
            const int J = (int)S.n_cols;
            if (J == 0) break;
            const uint32_t t_lo = S.t_lo, T = S.t_hi - S.t_lo;
            const int32_t tid = S.ru[S.unit[0]].tid;
            const int64_t c0 = __ldg(R.contig_off + tid);
            for (int k = t; k < TILE_TAB / 2; k += TILE_THREADS) reinterpret_cast<uint32_t *>(S.tab)[k] = 0xFFFFFFFFu;
            for (int k = t; k < J * L::BINS; k += TILE_THREADS) (&S.hist[0][0])[k] = 0u;
            if (t < J * 8) (&S.cnt[0][0])[t] = (t & 7) == 6 ? (int)(S.ru[S.unit[t >> 3]].meta & 3u) : 0;
            // ---- A
            for (uint32_t r = t; r < T; r += TILE_THREADS) {
                const int64_t i = (int64_t)t_lo + r;
                const uint4 hot = load_hot(R, i);
                const bool pass = read_passes(hot.z, P);
                bool elig = P.ignore_overlaps && pass && (hot.z & (F_PROPER | F_MUNMAP)) == F_PROPER, stores = false;
                uint32_t key = 0u;
                if (elig) {  // htslib overlap_push past its early exits (cf. pair_in_registers)
                    const int32_t mp = __ldg(R.mpos + i), mt = __ldg(R.mtid + i);
                    key = __ldg(R.qname_hash + i);
                    if (mt >= 0 && mt != tid) elig = false;
                    const int32_t end = (int32_t)hot.x + (int32_t)hot.y;
                    if (elig && mp >= end) {
                        const long long isz = (long long)__ldg(R.isize + i);
                        const int lq = (hot.z >> 24) == SHAPE_SIMPLE ? (int)hot.y : __ldg(R.l_qseq + i);
                        if ((isz < 0 ? -isz : isz) >= 2LL * lq) elig = false;
                    }
                    int32_t column = 0;
                    if (elig) {
                        column = push_column(R, P, i, c0);
                        if ((int32_t)hot.x + (int32_t)hot.y <= column) elig = false;  // never appended to the pileup buffer (a zero-span record)
                    }
                    S.pcol[r] = column;
                    stores = elig && (mp >= (int32_t)hot.x || ((hot.z & F_PAIRED) && mp == -1));
                }
                S.hot[r] = hot;
                S.key[r] = key;
                S.prt[r] = (uint16_t)MED_NONE;
                S.nxt[r] = (uint16_t)MED_NONE;
                S.rs[r] = (uint8_t)((pass ? RS_PASS : 0u) | (elig ? RS_ELIG : 0u) | (stores ? RS_STORES : 0u));
            }
            __syncthreads();
            // ---- B: htslib's name hash replayed for the tile (cf. link_one_read): the records that reach overlap_push are chained per
            // bucket, then every record walks the records of its name in file order -- found: pair and delete; not found: stored
            // if the mate is still to come; an entry whose record has left the pileup buffer is dropped
            for (uint32_t r = t; r < T; r += TILE_THREADS) {
                if (!(S.rs[r] & RS_ELIG)) continue;
                const uint32_t slot = (S.key[r] * 2654435761u) >> 20;  // 12 bits
                uint32_t head = S.tab[slot];
                for (;;) {  // push onto the bucket's chain
                    S.nxt[r] = (uint16_t)head;
                    const uint32_t seen = cas16(&S.tab[slot], head, r);
                    if (seen == head) break;
                    head = seen;
                }
            }
            __syncthreads();
            for (uint32_t r = t; r < T; r += TILE_THREADS) {
                if (!(S.rs[r] & RS_ELIG)) continue;
                const uint32_t key = S.key[r];
                const unsigned long long id = __ldg(R.qname_id + t_lo + r);
                const uint32_t head = S.tab[(key * 2654435761u) >> 20];
                int waiting = -1, cur = -1;
                uint32_t result = MED_NONE;
                for (;;) {
                    int m = -1;  // smallest member of the name beyond cur
                    for (uint32_t j = head; j != MED_NONE; j = S.nxt[j])
                        if ((int)j > cur && (m < 0 || (int)j < m) && S.key[j] == key && __ldg(R.qname_id + t_lo + j) == id) m = (int)j;
                    if (m < 0) break;
                    if (waiting >= 0 && (int32_t)S.hot[waiting].x + (int32_t)S.hot[waiting].y < S.pcol[m]) waiting = -1;  // its record already left the buffer
                    if (waiting < 0) {
                        if (S.rs[m] & RS_STORES) waiting = m;
                    } else {
                        if (waiting == (int)r || m == (int)r) {
                            result = (uint32_t)(waiting == (int)r ? m : waiting);
                            break;
                        }
                        waiting = -1;
                    }
                    cur = m;
                    if (m >= (int)r && waiting != (int)r) break;  // record r has been through without a partner
                }
                S.prt[r] = (uint16_t)result;
            }
            __syncthreads();
            // ---- C
            for (uint32_t r = t; r < T; r += TILE_THREADS) {
                const int64_t i = (int64_t)t_lo + r;
                const uint4 hot = S.hot[r];
                const bool pass = S.rs[r] & RS_PASS;
                int64_t qo = 0, so = 0;
                if (pass) { qo = __ldg(R.qual_off + i); so = __ldg(R.seq_off + i); }
                for (int j = 0; j < J; ++j) {
                    const ResolvedUnit &ru = S.ru[S.unit[j]];
                    uint32_t x = 0u;
                    if (i >= (int64_t)ru.lo && i < (int64_t)ru.hi) {
                        const ColRead c = resolve_column(R, ru.col, i, hot, true, pass);
                        if (c.pay) {
                            const uint32_t q = (uint32_t)__ldg(R.qual + qo + c.qpos), b = (uint32_t)__ldg(R.seq + so + (c.qpos >> 1));
                            x = q | (((c.qpos & 1) ? (b & 15u) : (b >> 4)) << 8) | (c.aligned ? 0x8000u : 0u);
                        }
                    }
                    S.xw[j][r] = (uint16_t)x;
                }
            }
            __syncthreads();
            // ---- D: a warp works on one column at a time (its counters are reduced across the warp), warps start at different columns
            for (uint32_t r0 = warp * 32; r0 < T; r0 += TILE_THREADS) {
                const uint32_t r = r0 + lane;
                const bool in = r < T;
                const int64_t i = (int64_t)t_lo + r;
                const uint4 hot = in ? S.hot[r] : make_uint4(0, 0, 0, 0);
                const uint32_t rs = in ? (uint32_t)S.rs[r] : 0u;
                const uint32_t pj = in ? (uint32_t)S.prt[r] : MED_NONE;
                const bool paired = pj != MED_NONE;  // (the replay has applied the protocol: stored / found / deleted)
                const uint32_t sel = wang_hash(in ? S.key[r] : 0u) & 1u;
                for (int jj = 0; jj < J; ++jj) {
                    int j = jj + (warp & (TILE_UNITS - 1));  // warps start at different columns
                    j = j >= J ? j - J : j;
                    j = j >= J ? j % J : j;
                    const ResolvedUnit &ru = S.ru[S.unit[j]];
                    Unit U;
                    unit_from(ru, ins_seq, U);
                    const bool cand = in && i >= (int64_t)ru.lo && i < (int64_t)ru.hi;
                    const ColRead c = resolve_column(R, ru.col, i, hot, cand, cand && (rs & RS_PASS));
                    const uint32_t xw = cand ? (uint32_t)S.xw[j][r] : 0u;
                    const uint32_t word = xw & 0xFFFu;
                    int q = 0;
                    uint32_t my_word = 0u;
                    if (c.aligned) {
                        const uint32_t px = paired ? (uint32_t)S.xw[j][pj] : 0u;
                        my_word = word;
                        q = overlap_quality(word, (px & 0x8000u) ? ((px & 0xFFFu) | 0x80000000u) : 0u, r < pj, sel);
                    } else if (c.hard) {
                        uint32_t mw = paired ? ((t_lo + pj) | (sel << 31)) : VFK_NO_MATE;
                        if (!paired && (rs & RS_STORES)) {  // the record after the column may be its mate (bam_plp_auto reads one ahead)
                            const int64_t nx = next_pushed(R, P, U.hi, contig_end(R, U.tid), U.stop);
                            if (nx >= 0 && ld_end(R, nx) > __ldg(R.pos + nx) && reaches_hash(R, P, nx, U.tid, INT32_MIN) &&
                                __ldg(R.qname_id + i) == __ldg(R.qname_id + nx) && __ldg(R.qname_hash + i) == __ldg(R.qname_hash + nx))
                                mw = (uint32_t)nx | ((wang_hash(__ldg(R.qname_hash + nx)) & 1u) << 31);
                        }
                        my_word = word;
                        q = hard_quality(R, P, U.col, U.stop, U.tid, U.hi, i, mw, c.qpos, word & 0xFFu, (word >> 8) * 0x11u);
                    }
                    const Eval e = finish_column(R, P, U, i, hot, c, q, my_word);
                    const Contribution cb = decode(e);
                    uint32_t status = 0u;
                    accumulate<uint16_t, POSB>(S.hist[j], S.hist[j] + WORDS, cb, U.kind == VFK_KIND_SNV, status);
                    const int n_cover = __reduce_add_sync(FULL, cb.covers ? 1 : 0), n_col = __reduce_add_sync(FULL, cb.in_col ? 1 : 0);
                    const int dp = __reduce_add_sync(FULL, cb.in_dp ? 1 : 0), n_amb = __reduce_add_sync(FULL, cb.ambiguous ? 1 : 0);
                    const int ac_ref = __reduce_add_sync(FULL, cb.ref ? 1 : 0), ac_alt = __reduce_add_sync(FULL, (int)cb.alt);
                    status = __reduce_or_sync(FULL, status);
                    if (lane == 0 && n_cover) {
                        atomicAdd(&S.cnt[j][0], n_cover); atomicAdd(&S.cnt[j][1], n_col); atomicAdd(&S.cnt[j][2], dp);
                        atomicAdd(&S.cnt[j][3], n_amb); atomicAdd(&S.cnt[j][4], ac_ref); atomicAdd(&S.cnt[j][5], ac_alt);
                        if (status) atomicOr(&S.cnt[j][6], (int)status);
                    }
                }
            }
            __syncthreads();
            // ---- E: one warp per (column, metric); then the records
            for (int task = warp; task < 3 * J; task += TILE_THREADS / 32) {
                const int j = task / 3, m = task - 3 * j;
                const ResolvedUnit &ru = S.ru[S.unit[j]];
                const int *cn = S.cnt[j];
                int m_ref = -1, m_alt = -1;
                uint64_t s2 = 0;
                const uint32_t *ref_slot = S.hist[j], *alt_slot = S.hist[j] + WORDS;
                if ((cn[4] | cn[5]) && (m != 1 || (ru.meta & 3u) == VFK_KIND_SNV)) {
                    if (m == 0) reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::MQ_OFF, lane, m_ref, m_alt, s2);
                    else if (m == 1) reduce_metric<uint16_t, 256>(ref_slot, alt_slot, L::BQ_OFF, lane, m_ref, m_alt, s2);
                    else reduce_metric<uint16_t, POSB>(ref_slot, alt_slot, L::POS_OFF, lane, m_ref, m_alt, s2);
                }
                if (lane == 0) { S.med2[j][m][0] = m_ref; S.med2[j][m][1] = m_alt; S.s2[j][m] = s2; }
            }
            __syncthreads();
            if (t < J) {
                const int j = t;
                const ResolvedUnit &ru = S.ru[S.unit[j]];
                const uint32_t kind = ru.meta & 3u;
                const int *cn = S.cnt[j];
                int med2[3][2];
                uint64_t s2[3];
                for (int m = 0; m < 3; ++m) { med2[m][0] = S.med2[j][m][0]; med2[m][1] = S.med2[j][m][1]; s2[m] = S.s2[j][m]; }
                int dp = cn[2];
                if (kind == VFK_KIND_SNV && !P.include_ambiguous) dp -= cn[3];  // pileups.py:224-227
                store_counts(out + u0 + S.unit[j], dp, cn[3], cn[4], cn[5], med2, s2, (uint32_t)cn[6], ru.hi - ru.lo, (uint32_t)cn[0], (uint32_t)cn[1]);
            }
            if (t == 0) atomicAdd(n_tiled, (uint32_t)J);
            __syncthreads();
        }
    }
}

// dynamic shared memory of the register kernel, per warp
template <int POSB>
struct WarpSmem {
    static constexpr bool HIST_HERE = POSB <= 512;  // per-warp histograms (and with them the medium path) fit next to four CTAs per SM
    static constexpr int HIST_WORDS = HIST_HERE ? Layout<POSB>::BINS : 0;  // two slots of BINS / 2 words
    static constexpr int WORDS_PER_WARP = HIST_WORDS;
    static constexpr size_t BYTES = (size_t)WARPS_PER_CTA * WORDS_PER_WARP * 4;
};

// INWARP = true: overlap partners are found in registers (pair_in_registers); false (debug hook, VFK_DEBUG_LINK_ALL=1 at
// vfk_create): the link kernels have run over every unit and the partners come from R.mate, as on the list path.
// FETCHED = true (host entry point with the read arrays in place): a lane's own base / quality come from the bytes the host
// gathered for the call's fetch list (request_kernel), at pair fetch_base[u] + candidate index.
template <int POSB, bool INWARP, bool FETCHED>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, VFK_MIN_CTAS)
pileup_warp_kernel(const __grid_constant__ ReadsDev R, const ResolvedUnit *__restrict__ res, int64_t n_units, const uint8_t *__restrict__ ins_seq,
                   const __grid_constant__ vfk_params P, vfk_counts *__restrict__ out, unsigned long long *__restrict__ work_counter,
                   uint32_t *__restrict__ block_list, uint32_t *__restrict__ n_block, uint32_t *__restrict__ warp_list,
                   uint32_t *__restrict__ n_warp, uint32_t *__restrict__ medium_list, uint32_t *__restrict__ n_medium,
                   const uint32_t *__restrict__ batch_status, const uint32_t *__restrict__ fetch_base, const uint16_t *__restrict__ fetched) {
    using L = Layout<POSB>;
    constexpr bool HIST_HERE = WarpSmem<POSB>::HIST_HERE;
    __shared__ uint32_t stage_all[WARPS_PER_CTA][32];
    extern __shared__ __align__(16) uint32_t dyn_smem[];  // per warp: a pair of u16 histograms (WarpSmem)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t *stage = stage_all[warp];
    uint32_t *hist = dyn_smem + (size_t)warp * WarpSmem<POSB>::WORDS_PER_WARP;
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_CTA;
    int64_t last_base = 0;
    const uint32_t bad = *batch_status;  // prep_kernel has finished: VFK_ST_BADBATCH or 0

    for (;;) {
        unsigned long long base = 0;
        const int64_t left = n_units - last_base;  // as of this warp's previous visit
        const int grab = left > 16 * n_warps ? GRAB : left > 4 * n_warps ? 2 : 1;
        if (lane == 0) base = atomicAdd(work_counter, (unsigned long long)grab);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= n_units) break;
        last_base = (int64_t)base;
        const int n_here = (int)min((int64_t)grab, n_units - (int64_t)base);
        // lane j keeps the record of unit base+j; one coalesced read for the whole grab
        uint4 ra = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
        if (lane < n_here) {
            const uint4 *q = reinterpret_cast<const uint4 *>(res + base + lane);
            ra = q[0];  // written by locate_kernel (an earlier launch of this call)
            rb = q[1];
        }
        uint32_t fbv = NO_FETCH;
        if constexpr (FETCHED) {
            if (lane < n_here) fbv = __ldg(fetch_base + base + lane);
        }
        // software pipeline: the header words of the NEXT unit's first 32 candidates are requested before the current
        // unit is processed
        uint4 hot_n = make_uint4(0, 0, 0, 0);
        {
            const uint32_t lo0 = __shfl_sync(FULL, rb.x, 0), hi0 = __shfl_sync(FULL, rb.y, 0);
            if (!bad && lo0 + lane < hi0) hot_n = load_hot(R, (int64_t)lo0 + lane);
        }
        for (int j = 0; j < n_here; ++j) {
            const int64_t u = (int64_t)base + j;
            ResolvedUnit ru;
            ru.col = (int32_t)__shfl_sync(FULL, ra.x, j); ru.meta = __shfl_sync(FULL, ra.y, j);
            ru.len = (int32_t)__shfl_sync(FULL, ra.z, j); ru.ins_off = __shfl_sync(FULL, ra.w, j);
            ru.lo = __shfl_sync(FULL, rb.x, j); ru.hi = __shfl_sync(FULL, rb.y, j);
            ru.stop = (int32_t)__shfl_sync(FULL, rb.z, j); ru.tid = (int32_t)__shfl_sync(FULL, rb.w, j);
            Unit U;
            unit_from(ru, ins_seq, U);
            uint32_t fb = NO_FETCH;
            if constexpr (FETCHED) fb = __shfl_sync(FULL, fbv, j);
            Cand A;
            A.hot = hot_n;
            A.mw = VFK_NO_MATE;
            A.st = 0u;
            if (j + 1 < n_here) {
                const uint32_t lo1 = __shfl_sync(FULL, rb.x, j + 1), hi1 = __shfl_sync(FULL, rb.y, j + 1);
                if (!bad && lo1 + lane < hi1) hot_n = load_hot(R, (int64_t)lo1 + lane);
            }
            const int64_t n_cand = U.hi - U.lo;
            int med2[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};
            uint64_t s2[3] = {0, 0, 0};
            if (n_cand == 0 || bad) {
                if (lane == 0) store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind | bad, 0u, 0u);
                continue;
            }
            if constexpr (WarpSmem<POSB>::HIST_HERE && INWARP) {
                if (n_cand > 64 && n_cand <= MED_CAP) {  // medium: its own kernel (slots of 32, per-read state parked in shared memory)
                    if (lane == 0) {
                        medium_list[atomicAdd(n_medium, 1u)] = (uint32_t)u;
                        store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind | VFK_ST_DEFERRED, (uint32_t)n_cand, 0u);
                    }
                    continue;
                }
            }
            if (n_cand > 64) {  // deep: handed to the histogram kernels through a device-side list
                if (lane == 0) {
                    if (n_cand > WARP_PATH_MAX_CAND) block_list[atomicAdd(n_block, 1u)] = (uint32_t)u;
                    else warp_list[atomicAdd(n_warp, 1u)] = (uint32_t)u;
                    atomicAdd(n_medium + 1, 1u);  // counters @28: deep units (the tile kernel looks for them only if there are any)
                    store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind | VFK_ST_DEFERRED, (uint32_t)n_cand, 0u);
                }
                continue;
            }
            // ---- shallow: lane k holds candidate k (slot A) and, when more than 32 reads are candidates, candidate 32 + k (slot B)
            const bool two = n_cand > 32;
            const bool validA = lane < n_cand, validB = lane + 32 < n_cand;
            const int64_t i_a = U.lo + lane, i_b = U.lo + 32 + lane;
            Cand B;
            B.hot = make_uint4(0, 0, 0, 0);
            B.mw = VFK_NO_MATE;
            B.st = 0u;
            if (validB) B.hot = load_hot(R, i_b);
            // column resolution and the reads' own base / quality lookups, all in flight before the pairing consumes anything
            const bool passA = validA && read_passes(A.hot.z, P), passB = validB && read_passes(B.hot.z, P);
            const ColRead cA = resolve_column(R, U.col, i_a, A.hot, validA, passA);
            PayLoad ownA, ownB;
            ownA.q = ownA.s = ownA.odd = 0u;
            ownB = ownA;
            if (cA.pay) ownA = (FETCHED && fb != NO_FETCH) ? pay_fetched(fetched, fb + (uint32_t)lane, (uint32_t)cA.qpos) : pay_issue(R, A.hot.w, (uint32_t)cA.qpos);
            ColRead cB = resolve_column(R, U.col, i_b, B.hot, false, false);
            if (two) {
                cB = resolve_column(R, U.col, i_b, B.hot, validB, passB);
                if (cB.pay) ownB = (FETCHED && fb != NO_FETCH) ? pay_fetched(fetched, fb + 32u + (uint32_t)lane, (uint32_t)cB.qpos) : pay_issue(R, B.hot.w, (uint32_t)cB.qpos);
            }
            bool general_protocol = false;
            if constexpr (INWARP) {
                general_protocol = !pair_in_registers(R, P, U.tid, lane, lt_mask, passA, passB, i_a, i_b, A, B);
            } else {
                if (validA) A.mw = __ldg(R.mate + i_a);
                if (validB) B.mw = __ldg(R.mate + i_b);
            }
            if (general_protocol) {  // three records of one name, a zero-span record: the link kernels replay the hash in full
                if (lane == 0) {
                    warp_list[atomicAdd(n_warp, 1u)] = (uint32_t)u;
                    store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind | VFK_ST_DEFERRED, (uint32_t)n_cand, 0u);
                }
                continue;
            }
            if constexpr (INWARP) {
                // A read inside a deletion / skip at the column is filtered on its NEXT base, and the overlap adjustment of that
                // base exists once the mate has been pushed -- which also holds for the first record pushed after the column
                // (bam_plp_auto reads one ahead).  That record lies outside the candidate range: when such a read stored itself
                // and found no partner among the candidates, the record after the column is tried.
                const bool t0 = cA.hard && A.mw == VFK_NO_MATE && (A.st & ST_STORES), t1 = cB.hard && B.mw == VFK_NO_MATE && (B.st & ST_STORES);
                if (__any_sync(FULL, t0 || t1)) {
                    long long nx = -1;
                    if (lane == 0) nx = next_pushed(R, P, U.hi, contig_end(R, U.tid), U.stop);
                    nx = __shfl_sync(FULL, nx, 0);
                    if (nx >= 0 && ld_end(R, nx) > __ldg(R.pos + nx) && reaches_hash(R, P, nx, U.tid, INT32_MIN)) {
                        const unsigned long long idn = __ldg(R.qname_id + nx);
                        const uint32_t hn = __ldg(R.qname_hash + nx);
                        if (t0 && __ldg(R.qname_id + i_a) == idn && __ldg(R.qname_hash + i_a) == hn) A.mw = (uint32_t)nx | ((wang_hash(hn) & 1u) << 31);
                        if (t1 && __ldg(R.qname_id + i_b) == idn && __ldg(R.qname_hash + i_b) == hn) B.mw = (uint32_t)nx | ((wang_hash(hn) & 1u) << 31);
                    }
                }
            }
            // the reads' payload words; a read's overlap partner is another candidate of the column: its word comes by a shuffle
            const uint32_t wA = pay_word(ownA), wB = pay_word(ownB);
            const uint32_t xA = cA.aligned ? (wA | 0x80000000u) : 0u, xB = cB.aligned ? (wB | 0x80000000u) : 0u;
            const int64_t kA = (int64_t)(A.mw & 0x7FFFFFFFu) - U.lo, kB = (int64_t)(B.mw & 0x7FFFFFFFu) - U.lo;  // partner as a candidate index
            uint32_t pA = __shfl_sync(FULL, xA, (int)(kA & 31)), pB = 0u;
            if (two) {
                const uint32_t pAb = __shfl_sync(FULL, xB, (int)(kA & 31));
                const uint32_t pBa = __shfl_sync(FULL, xA, (int)(kB & 31)), pBb = __shfl_sync(FULL, xB, (int)(kB & 31));
                if (kA >= 32) pA = pAb;
                pB = kB >= 32 ? pBb : pBa;
                if (B.mw == VFK_NO_MATE || kB < 0 || kB >= n_cand) pB = 0u;
            }
            if (A.mw == VFK_NO_MATE || kA < 0 || kA >= n_cand) pA = 0u;
            int qA = 0, qB = 0;
            uint32_t mwA = 0u, mwB = 0u;  // the payload word the classification sees
            if (cA.aligned) { mwA = wA; qA = overlap_quality(wA, pA, i_a < (int64_t)(A.mw & 0x7FFFFFFFu), A.mw >> 31); }
            if (cB.aligned) { mwB = wB; qB = overlap_quality(wB, pB, i_b < (int64_t)(B.mw & 0x7FFFFFFFu), B.mw >> 31); }
            if (__any_sync(FULL, cA.hard || cB.hard)) {
                if (cA.hard) { mwA = wA; qA = hard_quality(R, P, U.col, U.stop, U.tid, U.hi, i_a, A.mw, cA.qpos, ownA.q, ownA.s); }
                if (cB.hard) { mwB = wB; qB = hard_quality(R, P, U.col, U.stop, U.tid, U.hi, i_b, B.mw, cB.qpos, ownB.q, ownB.s); }
            }
            const Eval e0 = finish_column(R, P, U, i_a, A.hot, cA, qA, mwA);
            Eval e1;
            e1.pk = e1.fl = e1.alt = e1.hard = 0u;
            if (two) e1 = finish_column(R, P, U, i_b, B.hot, cB, qB, mwB);
            const bool k0 = e0.pk != 0u, k1 = e1.pk != 0u;
            const uint32_t m0 = __ballot_sync(FULL, k0), m1 = __ballot_sync(FULL, k1);
            const uint32_t both_lists = (e0.pk & (e0.pk << 1) & PK_ALT) | (e1.pk & (e1.pk << 1) & PK_ALT);  // REF == ALT: a read in both lists
            // more contributions than lanes, an insertion a read supports more than once (pileups.py:267-290), REF == ALT
            const bool by_histogram = __popc(m0) + __popc(m1) > 32 || __any_sync(FULL, e0.alt > 1u || e1.alt > 1u || both_lists != 0u);
            if (by_histogram) {
                if constexpr (HIST_HERE) {
                    reduce_by_histogram<POSB>(e0, e1, U.kind, P.include_ambiguous, hist, hist + L::BINS / 2, lane, out + u, (uint32_t)n_cand);
                } else {
                    if (lane == 0) {
                        warp_list[atomicAdd(n_warp, 1u)] = (uint32_t)u;
                        store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind | VFK_ST_DEFERRED, (uint32_t)n_cand, 0u);
                    }
                }
                continue;
            }
            uint32_t pk = e0.pk;
            if (m1) {  // move slot B's contributions into lanes slot A left free
                const int n1 = __popc(m1);
                if (k1) stage[__popc(m1 & lt_mask)] = e1.pk;
                __syncwarp();
                if (!k0) {
                    const int f = __popc(~m0 & lt_mask);
                    if (f < n1) pk = stage[f];
                }
                __syncwarp();
            }
            SlotShape sh;
            sh.is_ref = pk & PK_REF;
            sh.is_alt = pk & PK_ALT;
            sh.R = __ballot_sync(FULL, sh.is_ref);
            sh.A = __ballot_sync(FULL, sh.is_alt);
            // four counters (each <= 64) in the byte lanes of one word, one reduction
            const uint32_t tally = __reduce_add_sync(FULL, e0.fl + e1.fl);
            const int n_cover = (int)(tally & 0xFFu), n_col = (int)((tally >> 8) & 0xFFu), n_amb = (int)(tally >> 24);
            int dp = (int)((tally >> 16) & 0xFFu);
            const int ac_ref = __popc(sh.R), ac_alt = __popc(sh.A);
            if (sh.R | sh.A) {
                uint32_t lt, eq, w;
                const uint32_t mq = pk & 0xFFu, bq = (pk >> 8) & 0xFFu, qp = (pk >> 16) & 0xFFFu;
                sh.R_low = sh.R & lt_mask; sh.A_low = sh.A & lt_mask;
                sh.r1 = (ac_ref - 1) >> 1; sh.a1 = (ac_alt - 1) >> 1;
                sh.dr = (uint32_t)(ac_ref >> 1) - (uint32_t)sh.r1; sh.da = (uint32_t)(ac_alt >> 1) - (uint32_t)sh.a1;
                sh.sh_r = 1u - sh.dr; sh.sh_a = 17u - sh.da;
                radix_compare<8>(mq, sh.R | sh.A, lt, eq);
                uint32_t term = radix_finish(mq, lt, eq, sh, w);
                med2[0][0] = ac_ref ? (int)(w & 0xFFFFu) : -1; med2[0][1] = ac_alt ? (int)(w >> 16) : -1;
                if (U.kind == VFK_KIND_SNV) {
                    radix_compare<8>(bq, sh.R | sh.A, lt, eq);
                    term |= radix_finish(bq, lt, eq, sh, w) << 10;
                    med2[1][0] = ac_ref ? (int)(w & 0xFFFFu) : -1; med2[1][1] = ac_alt ? (int)(w >> 16) : -1;
                }
                radix_compare<L::POS_BITS>(qp, sh.R | sh.A, lt, eq);
                term |= radix_finish(qp, lt, eq, sh, w) << 20;
                med2[2][0] = ac_ref ? (int)(w & 0xFFFFu) : -1; med2[2][1] = ac_alt ? (int)(w >> 16) : -1;
                if (ac_ref && ac_alt) {
                    const uint32_t sums = __reduce_add_sync(FULL, term);
                    const uint64_t base2 = (uint64_t)(ac_alt * ac_alt + ac_alt);
                    s2[0] = base2 + (sums & 0x3FFu);
                    if (U.kind == VFK_KIND_SNV) s2[1] = base2 + ((sums >> 10) & 0x3FFu);
                    s2[2] = base2 + (sums >> 20);
                }
            }
            if (U.kind == VFK_KIND_SNV && !P.include_ambiguous) dp -= n_amb;  // pileups.py:224-227
            if (lane == 0)
                store_counts(out + u, dp, n_amb, ac_ref, ac_alt, med2, s2, U.kind, (uint32_t)n_cand, (uint32_t)n_cover, (uint32_t)n_col);
        }
    }
}

// ------------------------------------------------------------------------------------------
// the list path, ONE cooperative launch: mate linking of the listed units' windows (three passes, grid barriers in between),
// then a warp per listed unit with shared-memory histograms, then a CTA per unit for columns deeper than the u16 bins allow.
// With empty lists (nearly every 30x / 60x call) every CTA leaves at the first line.
// ------------------------------------------------------------------------------------------
// The listed units that fit the u16 histograms, found again by their status word (VFK_ST_DEFERRED, written by the register
// kernel) in UNIT order: a CTA takes eight consecutive units at a time and its eight warps work on neighbouring columns at the
// same moment -- their windows overlap (a deep panel: many times over), so a read's bytes are fetched from DRAM once and
// found in L1 / L2 by the warps next door.  Taken in list order (scrambled by the register kernel's self-scheduling) the
// same units re-fetch every read once per unit.
// Work distribution: the unit range is cut into one contiguous segment per SM; a CTA reads the id of the SM it runs on and
// takes the next units of THAT segment (one per warp and visit), so that everything resident on an SM -- four CTAs, 32 warps --
// works on the same stretch of columns and shares its L1, and the whole grid touches ~148 stretches at a time, which L2 holds.
// An SM that runs out of its segment takes from the next one that still has units.
constexpr int MAX_SEGMENTS = 256;
__device__ __forceinline__ uint32_t sm_id() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
template <int POSB>
__device__ __forceinline__ void deep_units(const ReadsDev &R, const ResolvedUnit *__restrict__ res, int64_t n_units, const uint8_t *__restrict__ ins_seq,
                                           const vfk_params &P, vfk_counts *__restrict__ out, uint32_t bad, uint32_t *smem, uint32_t *seg_ctr, int n_seg) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS / 2;  // u16 bins
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *ref_slot = smem + (size_t)warp * 2 * WORDS;
    uint32_t *alt_slot = ref_slot + WORDS;
    const int64_t seg_len = (n_units + n_seg - 1) / n_seg;
    const int home = (int)(sm_id() % (uint32_t)n_seg);
    for (;;) {  // every warp takes its next unit by itself: no CTA barrier on this path
        long long u = -1;
        for (int first = 0; first < n_seg && u < 0; first += 32) {  // segments home, home + 1, ... 32 at a time
            const int k = first + lane;
            const int cand = (home + k) % n_seg;
            const int64_t have = min(seg_len, n_units - (int64_t)cand * seg_len);  // units of the segment (<= 0: none)
            const bool open = k < n_seg && (int64_t)*((volatile uint32_t *)seg_ctr + cand) < have;
            uint32_t m = __ballot_sync(FULL, open);
            while (m && u < 0) {
                const int l = __ffs(m) - 1;
                m &= m - 1u;
                const int c = __shfl_sync(FULL, cand, l);
                const int64_t c_have = __shfl_sync(FULL, have, l);
                uint32_t start = 0u;
                if (lane == 0) start = atomicAdd(seg_ctr + c, 1u);
                start = __shfl_sync(FULL, start, 0);
                if ((int64_t)start < c_have) u = (int64_t)c * seg_len + start;
            }
        }
        if (u < 0) break;
        const uint4 w = reinterpret_cast<const uint4 *>(out + u)[2];  // {med2[2][0], med2[2][1], status, n_cand}
        if (!(w.z & VFK_ST_DEFERRED) || w.w > (uint32_t)WARP_PATH_MAX_CAND) continue;
        Unit U;
        unit_from(res[u], ins_seq, U);
        unit_by_histogram<POSB>(R, P, U, ref_slot, alt_slot, lane, out + u, bad);
    }
}

// CTA per unit, u32 counters: columns deeper than the warp path's u16 bins allow
template <int POSB>
__device__ __forceinline__ void block_units(const ReadsDev &R, const ResolvedUnit *__restrict__ res, const uint8_t *__restrict__ ins_seq, const vfk_params &P,
                                            vfk_counts *__restrict__ out, const uint32_t *__restrict__ deferred, uint32_t nd, uint32_t *smem) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS;  // u32 bins
    __shared__ int red[6];
    __shared__ uint32_t red_status;
    __shared__ int s_med2[3][2];
    __shared__ unsigned long long s_s2[3];
    uint32_t *ref_slot = smem, *alt_slot = smem + WORDS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t d = blockIdx.x; d < nd; d += gridDim.x) {
        const int64_t u = deferred[d];
        Unit U;
        unit_from(res[u], ins_seq, U);
        __syncthreads();
        for (int w = threadIdx.x; w < 2 * WORDS; w += blockDim.x) smem[w] = 0u;
        if (threadIdx.x < 6) red[threadIdx.x] = 0;
        if (threadIdx.x == 0) red_status = U.kind;
        if (threadIdx.x < 6) s_med2[threadIdx.x >> 1][threadIdx.x & 1] = -1;
        if (threadIdx.x < 3) s_s2[threadIdx.x] = 0ull;
        __syncthreads();
        int dp = 0, n_amb = 0, ac_ref = 0, ac_alt = 0, n_cover = 0, n_col = 0;
        uint32_t status = U.kind;
        const bool with_bq = U.kind == VFK_KIND_SNV;
        for (int64_t i = U.lo + threadIdx.x; i < U.hi; i += blockDim.x) {
            const Contribution c = evaluate(R, P, U, i);
            n_cover += c.covers ? 1 : 0;
            n_col += c.in_col ? 1 : 0;
            dp += c.in_dp ? 1 : 0;
            n_amb += c.ambiguous ? 1 : 0;
            ac_ref += c.ref ? 1 : 0;
            ac_alt += (int)c.alt;
            accumulate<uint32_t, POSB>(ref_slot, alt_slot, c, with_bq, status);
        }
        dp = __reduce_add_sync(FULL, dp);
        n_amb = __reduce_add_sync(FULL, n_amb);
        ac_ref = __reduce_add_sync(FULL, ac_ref);
        ac_alt = __reduce_add_sync(FULL, ac_alt);
        n_cover = __reduce_add_sync(FULL, n_cover);
        n_col = __reduce_add_sync(FULL, n_col);
        status = __reduce_or_sync(FULL, status);
        if (lane == 0) {
            atomicAdd(&red[4], n_cover);
            atomicAdd(&red[5], n_col);
            atomicAdd(&red[0], dp); atomicAdd(&red[1], n_amb); atomicAdd(&red[2], ac_ref); atomicAdd(&red[3], ac_alt);
            atomicOr(&red_status, status);
        }
        __syncthreads();
        if (red[2] | red[3]) {  // warps 0..2 reduce one metric each
            int m0 = -1, m1 = -1;
            uint64_t s = 0;
            if (warp == 0) reduce_metric<uint32_t, 256>(ref_slot, alt_slot, L::MQ_OFF, lane, m0, m1, s);
            else if (warp == 1 && with_bq) reduce_metric<uint32_t, 256>(ref_slot, alt_slot, L::BQ_OFF, lane, m0, m1, s);
            else if (warp == 2) reduce_metric<uint32_t, POSB>(ref_slot, alt_slot, L::POS_OFF, lane, m0, m1, s);
            if (warp < 3 && lane == 0) { s_med2[warp][0] = m0; s_med2[warp][1] = m1; s_s2[warp] = s; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int med2[3][2];
            uint64_t s2[3];
            for (int m = 0; m < 3; ++m) { med2[m][0] = s_med2[m][0]; med2[m][1] = s_med2[m][1]; s2[m] = s_s2[m]; }
            int d_p = red[0];
            if (U.kind == VFK_KIND_SNV && !P.include_ambiguous) d_p -= red[1];
            store_counts(out + u, d_p, red[1], red[2], red[3], med2, s2, red_status & ~VFK_ST_DEFERRED, (uint32_t)(U.hi - U.lo),
                         (uint32_t)red[4], (uint32_t)red[5]);
        }
    }
}

// LINK: false when the link kernels have already run over every unit (debug hook)
template <int POSB, bool LINK>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
tail_kernel(const __grid_constant__ ReadsDev R, const ResolvedUnit *__restrict__ res, int64_t n_units, const uint8_t *__restrict__ ins_seq,
            const __grid_constant__ vfk_params P, vfk_counts *out, const uint32_t *__restrict__ block_list,
            const uint32_t *__restrict__ n_block, const uint32_t *__restrict__ warp_list, const uint32_t *__restrict__ n_warp,
            const uint32_t *__restrict__ batch_status, const uint32_t *__restrict__ n_tiled, uint32_t *seg_ctr, int n_seg,
            const __grid_constant__ LinkScratch Lk) {
    extern __shared__ __align__(16) uint32_t smem[];
    const uint32_t nw = *n_warp, nb = *n_block;  // written by the register / medium kernels (earlier launches)
    if (nw <= *n_tiled && nb == 0u) return;      // nothing listed, or the tile kernel has finished all of it: the whole grid leaves here
                                                 // (every CTA takes the same way out: no barrier is left waiting)
    if constexpr (LINK) {
        cg::grid_group grid = cg::this_grid();
        link_mark(R, P, res, warp_list, (int64_t)nw, Lk, out);
        link_mark(R, P, res, block_list, (int64_t)nb, Lk, out);
        grid.sync();
        link_reads<0>(R, P, Lk);
        grid.sync();
        link_reads<1>(R, P, Lk);
        grid.sync();
    }
    if (nw) deep_units<POSB>(R, res, n_units, ins_seq, P, out, *batch_status, smem, seg_ctr, n_seg);
    if (nb) block_units<POSB>(R, res, ins_seq, P, out, block_list, nb, smem);
    if constexpr (LINK) link_reads<2>(R, P, Lk);  // bucket heads and bitmap left clean for the next call
}

// ------------------------------------------------------------------------------------------
// host-side launchers (called from vfk_api.cu)
// ------------------------------------------------------------------------------------------
cudaError_t launch_link(const ReadsDev &R, const vfk_params &P, const void *resolved, const uint32_t *list, const uint32_t *n_list,
                        int64_t n_all, uint32_t *heads, uint32_t n_buckets, uint32_t *next, int32_t *column, uint32_t *claimed,
                        uint32_t *mate, int sm_count, cudaStream_t stream, int *launches);

// What a launch needs to know about its kernels and need not ask the runtime on every call (a call is a fraction of a
// millisecond; each attribute set / occupancy query costs the host several microseconds).
struct LaunchCache {
    bool ready;
    int warp_ctas, tail_ctas;
};

// Per-call device scratch (owned by the handle / a staging slot).
//   counters: @0 u64 unit work counter, @8 u32 block-list length, @12 u32 warp-list length, @20 u32 batch status
//   (VFK_ST_BADBATCH or 0), @24 u32 medium-list length, @28 u32 deep units listed, @32 u32 units finished by the tile kernel, @40 u32 pairs
//   handed out by request_kernel, @64 256 x u32 per-SM segment counters of the list path.   lists: 3 * n_units u32.
struct PileupScratch {
    void *resolved;   // n_units * 32 B
    void *counters;   // 64 B
    uint32_t *lists;  // 2 * n_units
    // link scratch
    uint32_t *heads;
    uint32_t n_buckets;
    uint32_t *next;
    int32_t *column;
    uint32_t *claimed;
    uint32_t *mate;
    cudaEvent_t *ev;  // timing (vfk_set_timing): [2] after locate, [3] after the register kernel, [4] after the list path; or NULL
    LaunchCache *cache;  // [4], one per position-range instantiation: function attributes set / occupancies queried once per handle
    // fetch list of the host entry point (request_kernel above); fetch_cap = 0: none
    uint32_t fetch_cap;         // pairs the list holds
    int fetch_by_index;         // requests name (read, read position) instead of pool offsets
    uint32_t *fetch_base;       // [n_units] device: first pair of a unit, or NO_FETCH
    void *fetch_req;            // [fetch_cap] x 16 B, page-locked host memory as the device sees it
    const uint16_t *fetched;    // [fetch_cap] device: the gathered bytes (copied in by fetch_between)
    cudaError_t (*fetch_between)(void *ctx, const uint32_t *pair_counter, cudaStream_t stream);  // host gather + copy, enqueued between the two kernels
    void *fetch_ctx;
};

constexpr int posb_index(int posb) { return posb == 256 ? 0 : posb == 512 ? 1 : posb == 1024 ? 2 : 3; }

template <int POSB, bool LINK>
static cudaError_t launch_tail(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, vfk_counts *out, const PileupScratch &S, int sm_count,
                               cudaStream_t stream, int *launches) {
    using L = Layout<POSB>;
    const size_t smem = std::max((size_t)WARPS_PER_CTA * 2 * (L::BINS / 2) * sizeof(uint32_t),  // a pair of u16 histograms per warp
                                 (size_t)2 * L::BINS * sizeof(uint32_t));                        // a pair of u32 histograms per CTA
    cudaError_t e;
    LaunchCache &C = S.cache[posb_index(POSB)];
    if (!C.tail_ctas) {
        if (smem > 40 * 1024) {  // with the kernel's static shared memory on top, 48 KB is the limit without the opt-in
            e = cudaFuncSetAttribute(tail_kernel<POSB, LINK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        // the occupancy query assumes the largest shared-memory carve-out: ask for it, so that the launch sees the same limit
        (void)cudaFuncSetAttribute(tail_kernel<POSB, LINK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int ctas = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, tail_kernel<POSB, LINK>, WARPS_PER_CTA * 32, smem);
        if (e != cudaSuccess) return e;
        C.tail_ctas = ctas < 1 ? 1 : ctas;
    }
    const int ctas = C.tail_ctas;
    // every CTA must be resident (grid barriers); more warps than units would only idle
    int64_t grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count * ctas, (V.n + WARPS_PER_CTA - 1) / WARPS_PER_CTA));
    const ResolvedUnit *res = (const ResolvedUnit *)S.resolved;
    const uint8_t *ins_seq = V.ins_seq;
    const uint32_t *n_block = (const uint32_t *)((char *)S.counters + 8), *n_warp = (const uint32_t *)((char *)S.counters + 12);
    const uint32_t *batch_status = (const uint32_t *)((char *)S.counters + 20);
    const uint32_t *block_list = S.lists, *warp_list = S.lists + V.n;
    int64_t n_units = V.n;
    const uint32_t *n_tiled = (const uint32_t *)((char *)S.counters + 32);
    uint32_t *seg_ctr = (uint32_t *)((char *)S.counters + 64);  // one work counter per SM segment (zeroed with the counters)
    int n_seg = std::min(sm_count, MAX_SEGMENTS);
    LinkScratch Lk;
    Lk.heads = S.heads; Lk.n_buckets = S.n_buckets; Lk.next = S.next; Lk.column = S.column; Lk.claimed = S.claimed; Lk.mate = S.mate;
    void *args[] = {(void *)&R, (void *)&res, (void *)&n_units, (void *)&ins_seq, (void *)&P, (void *)&out, (void *)&block_list, (void *)&n_block,
                    (void *)&warp_list, (void *)&n_warp, (void *)&batch_status, (void *)&n_tiled, (void *)&seg_ctr, (void *)&n_seg, (void *)&Lk};
    e = cudaLaunchCooperativeKernel((const void *)tail_kernel<POSB, LINK>, dim3((unsigned)grid), dim3(WARPS_PER_CTA * 32), args, smem, stream);
    if (e == cudaErrorCooperativeLaunchTooLarge && grid > sm_count) {  // the driver counts fewer resident CTAs than the query: one per SM always fits
        (void)cudaGetLastError();
        grid = sm_count;
        e = cudaLaunchCooperativeKernel((const void *)tail_kernel<POSB, LINK>, dim3((unsigned)grid), dim3(WARPS_PER_CTA * 32), args, smem, stream);
    }
    ++*launches;
    return e;
}

template <int POSB>
static cudaError_t launch_posb(ReadsDev R, const UnitsDev &V, const vfk_params &P, vfk_counts *out, const PileupScratch &S,
                               int sm_count, cudaStream_t stream, int *launches, bool link_all) {
    ResolvedUnit *res = (ResolvedUnit *)S.resolved;
    unsigned long long *work = (unsigned long long *)S.counters;
    uint32_t *n_block = (uint32_t *)((char *)S.counters + 8), *n_warp = (uint32_t *)((char *)S.counters + 12);
    const uint32_t *batch_status = (const uint32_t *)((char *)S.counters + 20);
    uint32_t *block_list = S.lists, *warp_list = S.lists + V.n, *medium_list = S.lists + 2 * V.n;
    uint32_t *n_medium = (uint32_t *)((char *)S.counters + 24);
    cudaError_t e;
    R.mate = S.mate;
    locate_kernel<<<(unsigned)((V.n + 255) / 256), 256, 0, stream>>>(R, V, res);
    ++*launches;
    if (S.ev) cudaEventRecord(S.ev[2], stream);
    const bool fetch = S.fetch_cap > 0u && !link_all;
    if (fetch) {
        uint32_t *pair_counter = (uint32_t *)((char *)S.counters + 40);
        request_kernel<<<(unsigned)((V.n + 7) / 8), 256, 0, stream>>>(R, res, V.n, S.fetch_base, (ulonglong2 *)S.fetch_req, pair_counter, S.fetch_cap,
                                                                     WarpSmem<POSB>::HIST_HERE ? (uint32_t)MED_CAP : 64u, S.fetch_by_index, batch_status);
        ++*launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = S.fetch_between(S.fetch_ctx, pair_counter, stream);
        if (e != cudaSuccess) return e;
    }
    int ctas_per_sm = 0;
    const size_t warp_smem = WarpSmem<POSB>::BYTES;
    LaunchCache &C = S.cache[posb_index(POSB)];
    if (!C.ready) {  // once per handle: function attributes and the register kernel's occupancy
        if (warp_smem > 40 * 1024) {
            e = link_all ? cudaFuncSetAttribute(pileup_warp_kernel<POSB, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem)
                         : cudaFuncSetAttribute(pileup_warp_kernel<POSB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(pileup_warp_kernel<POSB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)warp_smem);
            if (e != cudaSuccess) return e;
        }
        e = link_all ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, pileup_warp_kernel<POSB, false, false>, WARPS_PER_CTA * 32, warp_smem)
                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, pileup_warp_kernel<POSB, true, false>, WARPS_PER_CTA * 32, warp_smem);
        if (e != cudaSuccess) return e;
        C.warp_ctas = ctas_per_sm < 1 ? 1 : ctas_per_sm;
        if constexpr (WarpSmem<POSB>::HIST_HERE) {
            using L = Layout<POSB>;
            e = cudaFuncSetAttribute(medium_kernel<POSB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)WARPS_PER_CTA * (L::BINS * 4 + sizeof(MedSmem))));
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(tile_kernel<POSB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<POSB>));
            if (e != cudaSuccess) return e;
        }
        C.ready = true;
    }
    ctas_per_sm = C.warp_ctas;
    // persistent grid: a whole number of waves, never more CTAs than there is work
    int64_t want = (V.n + (int64_t)GRAB * WARPS_PER_CTA - 1) / ((int64_t)GRAB * WARPS_PER_CTA);
    int64_t grid = (int64_t)sm_count * ctas_per_sm;
    if (want < grid) grid = want < 1 ? 1 : want;
    if (link_all) {
        e = launch_link(R, P, res, nullptr, nullptr, V.n, S.heads, S.n_buckets, S.next, S.column, S.claimed, S.mate, sm_count, stream, launches);
        if (e != cudaSuccess) return e;
        pileup_warp_kernel<POSB, false, false><<<(unsigned)grid, WARPS_PER_CTA * 32, warp_smem, stream>>>(
            R, res, V.n, V.ins_seq, P, out, work, block_list, n_block, warp_list, n_warp, medium_list, n_medium, batch_status, nullptr, nullptr);
    } else if (fetch) {
        pileup_warp_kernel<POSB, true, true><<<(unsigned)grid, WARPS_PER_CTA * 32, warp_smem, stream>>>(
            R, res, V.n, V.ins_seq, P, out, work, block_list, n_block, warp_list, n_warp, medium_list, n_medium, batch_status, S.fetch_base, S.fetched);
    } else {
        pileup_warp_kernel<POSB, true, false><<<(unsigned)grid, WARPS_PER_CTA * 32, warp_smem, stream>>>(
            R, res, V.n, V.ins_seq, P, out, work, block_list, n_block, warp_list, n_warp, medium_list, n_medium, batch_status, nullptr, nullptr);
    }
    ++*launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (S.ev) cudaEventRecord(S.ev[3], stream);
    if constexpr (WarpSmem<POSB>::HIST_HERE) {
        if (!link_all) {  // columns with 65 .. MED_CAP candidate reads (the list is empty for a 30x batch: the CTAs leave at once)
            using L = Layout<POSB>;
            const size_t med_smem = (size_t)WARPS_PER_CTA * (L::BINS * 4 + sizeof(MedSmem));
            const int64_t med_grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count * 3, (V.n + WARPS_PER_CTA - 1) / WARPS_PER_CTA));
            medium_kernel<POSB><<<(unsigned)med_grid, WARPS_PER_CTA * 32, med_smem, stream>>>(R, res, V.ins_seq, P, out, medium_list, n_medium, warp_list, n_warp,
                                                                                              fetch ? S.fetch_base : nullptr, fetch ? S.fetched : nullptr);
            ++*launches;
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
    }
    if constexpr (WarpSmem<POSB>::HIST_HERE) {
        if (!link_all) {  // deep panels: tiles of neighbouring deep units, read-major (no deep unit listed: the CTAs leave at once)
            const size_t tile_smem = sizeof(TileSmem<POSB>);
            const int64_t tile_grid = std::max<int64_t>(1, std::min<int64_t>((int64_t)sm_count, (V.n + TILE_UNITS - 1) / TILE_UNITS));
            tile_kernel<POSB><<<(unsigned)tile_grid, TILE_THREADS, tile_smem, stream>>>(R, res, V.n, V.ins_seq, P, out, n_medium + 1, n_medium + 2);
            ++*launches;
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
        }
    }
    // the listed units: their windows linked, then the histogram paths -- one launch
    e = link_all ? launch_tail<POSB, false>(R, V, P, out, S, sm_count, stream, launches) : launch_tail<POSB, true>(R, V, P, out, S, sm_count, stream, launches);
    if (S.ev) cudaEventRecord(S.ev[4], stream);
    return e;
}

// locate + pileup kernels on a batch prep_kernel has prepared (R.ee / R.chunk_max / R.blk_incl valid, counters zeroed except the
// batch status word)
cudaError_t launch_pileup(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int max_l_qseq, vfk_counts *out,
                          const PileupScratch &S, int sm_count, cudaStream_t stream, int *launches, bool link_all) {
    if (V.n <= 0) return cudaSuccess;
    // read positions run 0 .. l_qseq (a read hanging in a trailing deletion reports l_qseq)
    if (max_l_qseq < 256) return launch_posb<256>(R, V, P, out, S, sm_count, stream, launches, link_all);
    if (max_l_qseq < 512) return launch_posb<512>(R, V, P, out, S, sm_count, stream, launches, link_all);
    if (max_l_qseq < 1024) return launch_posb<1024>(R, V, P, out, S, sm_count, stream, launches, link_all);
    return launch_posb<4096>(R, V, P, out, S, sm_count, stream, launches, link_all);
}

// the located record of every unit, for the list-shaped auxiliary (vfk_values.cu)
cudaError_t launch_locate_only(const ReadsDev &R, const UnitsDev &V, void *resolved, cudaStream_t stream, int *launches) {
    if (V.n <= 0) return cudaSuccess;
    locate_kernel<<<(unsigned)((V.n + 255) / 256), 256, 0, stream>>>(R, V, (ResolvedUnit *)resolved);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace vfk

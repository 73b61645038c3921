// vfk_pileup.cu -- the pileup kernels (sm_100a).
//
// Replaces vafator/pileups.py:106-362 + vafator/pileup_utils.py:64-92: for every unit (variant
// column x ALT allele) find the overlapping reads, classify them, and reduce the per-allele
// MQ / BQ / read-position distributions to integers: counts, twice-the-median, and twice the
// ALT mid-rank sum that scipy.stats.ranksums would compute (vafator/rank_sum_test.py:11).
//
// locate_kernel       : one thread per unit, 8-ary search of the position index.
// pileup_warp_kernel  : one warp per unit, lanes = reads spanning the column.  Persistent grid, units handed
//                       out in contiguous grabs from a global counter (shrinking towards the end).  Handles
//                       the shallow columns (<= 64 candidates, <= 32 of them contributing -- nearly every
//                       30x / 60x column): values stay in registers, ranks come from a bit-sliced (radix)
//                       comparison built on warp ballots.  No shared-memory histograms, no atomics, no
//                       calls, no stack.  Anything else (at 30x: the 3-4 units in 10 000 with more than 32
//                       contributing reads) is appended to one of two device-side lists:
// pileup_deep_kernel  : one warp per listed unit, histograms in shared memory (u16 counters, two per
//                       word) + warp scans;
// pileup_block_kernel : one CTA per listed unit, u32 counters, for columns with > 32767 candidates.
// pileup_column_kernel: (batches with a column store, the default) one warp per SNV unit, lane = sector of the column's
//                       window run; what it cannot serve goes to the kernels above through the same lists.
// Staged, opt-in, not yet run on a GPU (the default instantiations are untouched -- tools/sass_equal.py):
// pileup_column_kernel<.., TR = true> (VFK_COL_TRANSPOSED), <.., MQS = true> (VFK_MQ_SELECT), pileup_warp_kernel<.., LISTED = true>
// (VFK_LIST_SHALLOW).
#include <algorithm>
#include <cstdlib>
#include "vfk_device.cuh"

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
        for (int j = 0; j < 7; ++j) v[j] = __ldg(&R.sb[a + (j + 1) * step].x);
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
            if (a + j < b) c += __ldg(&R.sb[a + j].x) <= col ? 1 : 0;
        a += c;
    }
    ResolvedUnit r;
    r.col = col;
    r.meta = __ldg(V.meta + u);
    r.len = __ldg(V.length + u);
    r.ins_off = __ldg(V.seq_off + u);
    r.hi = (uint32_t)a;
    r.lo = a > c0 ? (uint32_t)__ldg(&R.sb[a - 1].y) : (uint32_t)a;
    r.stop = V.stop ? __ldg(V.stop + u) : 0x7FFFFFFF;
    r.c1 = (uint32_t)c1;
    return r;
}
__device__ __forceinline__ void store_resolved(ResolvedUnit *dst, const ResolvedUnit &r) {
    uint4 *o = reinterpret_cast<uint4 *>(dst);
    const uint4 *q = reinterpret_cast<const uint4 *>(&r);
    o[0] = q[0];
    o[1] = q[1];
}
// column store: the window of the column -> its run of sectors
__device__ __forceinline__ uint4 column_unit(const ReadsDev &R, int32_t tid, int32_t col, uint32_t meta) {
    const int32_t w = col / VFK_COL_WIDTH;
    uint32_t off = 0, cnt = 0;
    if (tid >= 0 && tid < R.n_contigs && col >= 0) {
        const int64_t w0 = __ldg(R.col_wbase + tid), w1 = __ldg(R.col_wbase + tid + 1);
        if (w0 + w < w1) {
            off = __ldg(R.col_off + w0 + w);
            cnt = __ldg(R.col_off + w0 + w + 1) - off;
        }
    }
    return make_uint4(off, cnt, (uint32_t)(col - w * VFK_COL_WIDTH), meta);
}

__global__ void __launch_bounds__(256)
locate_kernel(ReadsDev R, UnitsDev V, ResolvedUnit *__restrict__ res, uint4 *__restrict__ cunits,
              unsigned long long *__restrict__ cand_sum) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= V.n) return;
    const ResolvedUnit r = locate_unit(R, V, u);
    store_resolved(res + u, r);
    if (cunits) cunits[u] = column_unit(R, __ldg(V.tid + u), r.col, r.meta);
    if (cand_sum) {  // total candidate reads of the batch: lets the host entry pick its transfer strategy
        const unsigned active = __activemask();
        const unsigned mine = r.hi - r.lo;
        const unsigned tot = __reduce_add_sync(active, mine);
        if ((threadIdx.x & 31) == (__ffs(active) - 1)) atomicAdd(cand_sum, (unsigned long long)tot);
    }
}

// Column-store launches locate lazily: every unit only needs its window (no search) ...
__global__ void __launch_bounds__(256) column_units_kernel(ReadsDev R, UnitsDev V, uint4 *__restrict__ cunits) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= V.n) return;
    cunits[u] = column_unit(R, __ldg(V.tid + u), __ldg(V.pos + u), __ldg(V.meta + u));
}
// ... and only the units the column kernel left on the list are searched in the position index; columns too deep
// for the warp-per-unit histogram kernel move on to the block list (their slot is voided)
constexpr uint32_t LIST_VOID = 0xFFFFFFFFu;
__global__ void __launch_bounds__(256)
locate_list_kernel(ReadsDev R, UnitsDev V, ResolvedUnit *__restrict__ res, uint32_t *__restrict__ warp_list,
                   const uint32_t *__restrict__ n_warp, uint32_t *__restrict__ block_list, uint32_t *__restrict__ n_block) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)*n_warp) return;
    const uint32_t u = warp_list[i];
    const ResolvedUnit r = locate_unit(R, V, u);
    store_resolved(res + u, r);
    if (r.hi - r.lo > (uint32_t)WARP_PATH_MAX_CAND) {
        block_list[atomicAdd(n_block, 1u)] = u;
        warp_list[i] = LIST_VOID;
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

// The same lt / eq masks by selection: one round per DISTINCT value (smallest first) -- a warp minimum, a ballot of its
// holders.  Cheaper than the radix walk when a column holds few distinct values, which is the rule for mapping
// qualities (EXPERIMENTAL, not yet verified on a GPU: VFK_MQ_SELECT=1 uses it for the MQ metric of the column kernel).
__device__ __forceinline__ void select_compare(uint32_t v, uint32_t contributing, uint32_t lane_bit, uint32_t &lt, uint32_t &eq) {
    lt = 0u;
    eq = contributing;
    uint32_t rem = contributing, done = 0u;
    while (rem) {  // uniform across the warp
        const bool waiting = (rem & lane_bit) != 0u;
        const uint32_t m = __reduce_min_sync(FULL, waiting ? v : 0xFFFFFFFFu);
        const uint32_t g = __ballot_sync(FULL, waiting && v == m);
        if (g & lane_bit) {
            lt = done;
            eq = g;
        }
        done |= g;
        rem &= ~g;
    }
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
// Per-warp staging ring of the deep kernel: two stages of 32 read headers, filled by bulk async
// copies (one elected lane issues them) and consumed from shared memory.
struct StageRing {
    uint4 *hot;       // [2][32]
    uint64_t *bar;    // [2]
    uint32_t parity;  // bit s = phase the next wait on stage s expects
};

template <int POSB, bool STAGED>
__device__ __forceinline__ void unit_by_histogram(const ReadsDev &R, const vfk_params &P, const Unit &U, uint32_t *ref_slot,
                                               uint32_t *alt_slot, int lane, vfk_counts *out, StageRing *ring = nullptr) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS / 2;
    {
        uint4 *z = reinterpret_cast<uint4 *>(ref_slot);
        for (int w = lane; w < (2 * WORDS) / 4; w += 32) z[w] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
    int dp = 0, n_amb = 0, ac_ref = 0, ac_alt = 0, n_cover = 0, n_col = 0;
    uint32_t status = U.kind;
    const bool with_bq = U.kind == VFK_KIND_SNV;
    if (STAGED) {
        // chunk c of 32 candidates lives in stage c & 1; chunk c+1 is in flight while c is evaluated
        const int64_t n_cand = U.hi - U.lo;
        const int n_chunks = (int)((n_cand + 31) >> 5);
        auto issue = [&](int c) {
            const int s = c & 1;
            const int64_t first = U.lo + ((int64_t)c << 5);
            const uint32_t bytes = (uint32_t)min((int64_t)32, U.hi - first) * 16u;
            if (lane == 0) {
                mbar_expect_tx(ring->bar + s, bytes);
                bulk_copy_g2s(ring->hot + s * 32, R.hot + first, bytes, ring->bar + s);
            }
        };
        if (n_chunks > 0) issue(0);  // a unit without candidates must not touch the ring: its phase would run ahead
        uint32_t mw_next = (U.lo + lane < U.hi) ? __ldg(R.mate + U.lo + lane) : VFK_NO_MATE;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = c & 1;
            const int64_t i = U.lo + ((int64_t)c << 5) + lane;
            if (c + 1 < n_chunks) issue(c + 1);
            const uint32_t mw = mw_next;
            if (c + 1 < n_chunks && i + 32 < U.hi) mw_next = __ldg(R.mate + i + 32);
            mbar_wait(ring->bar + s, (ring->parity >> s) & 1u);
            ring->parity ^= 1u << s;
            if (i < U.hi) {
                const uint4 hot = ring->hot[s * 32 + lane];
                const Contribution c1 = evaluate(R, P, U, i, hot, mw);
                n_cover += c1.covers ? 1 : 0;
                n_col += c1.in_col ? 1 : 0;
                dp += c1.in_dp ? 1 : 0;
                n_amb += c1.ambiguous ? 1 : 0;
                ac_ref += c1.ref ? 1 : 0;
                ac_alt += (int)c1.alt;
                accumulate<uint16_t, POSB>(ref_slot, alt_slot, c1, with_bq, status);
            }
            __syncwarp();  // every lane has consumed stage s before it is refilled two chunks later
        }
    } else {
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
// LISTED = false: units 0 .. n_units-1.  LISTED = true (EXPERIMENTAL, not yet verified on a GPU; VFK_LIST_SHALLOW=1):
// the units the column kernel left on `warp_list` -- nearly all of them shallow columns it could not serve (a passing
// read inside a deletion, an unusable sector) -- are evaluated here instead of by the histogram kernel: entry k of the
// list is voided once its unit is done, what this kernel defers too (deep, heavy) stays for pileup_deep_kernel.
template <int POSB, bool LISTED>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, VFK_MIN_CTAS)
pileup_warp_kernel(ReadsDev R, const ResolvedUnit *__restrict__ res, int64_t n_units, const uint8_t *__restrict__ ins_seq,
                   vfk_params P, vfk_counts *__restrict__ out, unsigned long long *__restrict__ work_counter,
                   uint32_t *__restrict__ block_list, uint32_t *__restrict__ n_block, uint32_t *__restrict__ warp_list,
                   uint32_t *__restrict__ n_warp) {
    using L = Layout<POSB>;
    __shared__ uint32_t stage_all[WARPS_PER_CTA][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t *stage = stage_all[warp];
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_CTA;
    int64_t last_base = 0;
    if constexpr (LISTED) n_units = (int64_t)*n_warp;

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
        uint32_t uid = LIST_VOID;  // LISTED: the unit behind list entry base + lane
        if (lane < n_here) {
            if constexpr (LISTED) {
                uid = warp_list[base + lane];
                if (uid != LIST_VOID) {
                    const uint4 *q = reinterpret_cast<const uint4 *>(res + uid);
                    ra = q[0];  // written by locate_list_kernel (an earlier launch of this call)
                    rb = q[1];
                }
            } else {
                const uint4 *q = reinterpret_cast<const uint4 *>(res + base + lane);
                ra = __ldg(q);
                rb = __ldg(q + 1);
            }
        }
        // software pipeline: the header + mate word of the NEXT unit's first 32 candidates are
        // requested before the current unit is processed
        uint4 hot_n = make_uint4(0, 0, 0, 0);
        uint32_t mw_n = VFK_NO_MATE;
        {
            const uint32_t lo0 = __shfl_sync(FULL, rb.x, 0), hi0 = __shfl_sync(FULL, rb.y, 0);
            if (lo0 + lane < hi0) { hot_n = __ldg(R.hot + lo0 + lane); mw_n = __ldg(R.mate + lo0 + lane); }
        }
        for (int j = 0; j < n_here; ++j) {
            int64_t u = (int64_t)base + j;
            if constexpr (LISTED) u = (int64_t)__shfl_sync(FULL, uid, j);
            ResolvedUnit ru;
            ru.col = (int32_t)__shfl_sync(FULL, ra.x, j); ru.meta = __shfl_sync(FULL, ra.y, j);
            ru.len = (int32_t)__shfl_sync(FULL, ra.z, j); ru.ins_off = __shfl_sync(FULL, ra.w, j);
            ru.lo = __shfl_sync(FULL, rb.x, j); ru.hi = __shfl_sync(FULL, rb.y, j);
            ru.stop = (int32_t)__shfl_sync(FULL, rb.z, j); ru.c1 = __shfl_sync(FULL, rb.w, j);
            Unit U;
            unit_from(ru, ins_seq, U);
            const uint4 hot_c = hot_n;
            const uint32_t mw_c = mw_n;
            if (j + 1 < n_here) {
                const uint32_t lo1 = __shfl_sync(FULL, rb.x, j + 1), hi1 = __shfl_sync(FULL, rb.y, j + 1);
                if (lo1 + lane < hi1) { hot_n = __ldg(R.hot + lo1 + lane); mw_n = __ldg(R.mate + lo1 + lane); }
            }
            if constexpr (LISTED) {
                if (u == (int64_t)LIST_VOID) continue;  // an entry locate_list_kernel moved to the block list
            }
            const int64_t n_cand = U.hi - U.lo;
            int med2[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};
            uint64_t s2[3] = {0, 0, 0};
            if (n_cand == 0) {
                if (lane == 0) {
                    store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind, 0u, 0u);
                    if constexpr (LISTED) warp_list[base + j] = LIST_VOID;
                }
                continue;
            }
            if (n_cand > 64) {  // deep: handed to the histogram kernels through a device-side list
                if constexpr (!LISTED) {
                    if (lane == 0) {
                        if (n_cand > WARP_PATH_MAX_CAND) block_list[atomicAdd(n_block, 1u)] = (uint32_t)u;
                        else warp_list[atomicAdd(n_warp, 1u)] = (uint32_t)u;
                        store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind | VFK_ST_DEFERRED, (uint32_t)n_cand, 0u);
                    }
                }
                continue;  // LISTED: the entry stays on the list (its record already says deferred)
            }
            // ---- shallow: the reads that span the column are packed into the lanes (first chunk in place,
            // second chunk into the lanes the first one leaves free) and evaluated in one pass; a second
            // pass only runs for the overflow when more than 32 reads span the column ----
            bool act = lane < n_cand && (uint32_t)(U.col - (int32_t)hot_c.x) < hot_c.y;
            int64_t i_a = U.lo + lane;
            uint4 hot_a = hot_c;
            uint32_t mw_a = mw_c;
            Eval e0, e1;
            e0.pk = e0.fl = e0.alt = 0u;
            e1 = e0;
            if (n_cand > 32) {
                const bool has1 = lane + 32 < n_cand;
                uint4 hot_1 = make_uint4(0, 0, 0, 0);
                uint32_t mw_1 = VFK_NO_MATE;
                if (has1) { hot_1 = __ldg(R.hot + i_a + 32); mw_1 = __ldg(R.mate + i_a + 32); }
                const bool cov1 = has1 && (uint32_t)(U.col - (int32_t)hot_1.x) < hot_1.y;
                const uint32_t c0m = __ballot_sync(FULL, act), c1m = __ballot_sync(FULL, cov1);
                const int n_free = 32 - __popc(c0m), idx1 = __popc(c1m & lt_mask);
                const int n_move = min(n_free, __popc(c1m));
                if (cov1 && idx1 < n_move) stage[idx1] = (uint32_t)lane;
                __syncwarp();
                const int slot = __popc(~c0m & lt_mask);  // this lane's rank among the free lanes
                const bool adopt = !act && slot < n_move;
                const int src = adopt ? (int)stage[slot] : lane;
                __syncwarp();
                const uint32_t hx = __shfl_sync(FULL, hot_1.x, src), hy = __shfl_sync(FULL, hot_1.y, src);
                const uint32_t hz = __shfl_sync(FULL, hot_1.z, src), hw = __shfl_sync(FULL, hot_1.w, src);
                const uint32_t hm = __shfl_sync(FULL, mw_1, src);
                if (adopt) {
                    hot_a = make_uint4(hx, hy, hz, hw);
                    mw_a = hm;
                    i_a = U.lo + 32 + src;
                    act = true;
                }
                const bool over = cov1 && idx1 >= n_move;
                if (__any_sync(FULL, over)) {
                    if (over) e1 = evaluate_packed(R, P, U, U.lo + 32 + lane, hot_1, mw_1);
                }
            }
            if (act) e0 = evaluate_packed(R, P, U, i_a, hot_a, mw_a);
            const bool k0 = e0.pk != 0u, k1 = e1.pk != 0u;
            const uint32_t m0 = __ballot_sync(FULL, k0), m1 = __ballot_sync(FULL, k1);
            // a read can support an insertion more than once (pileups.py:267-290): such units take the deep path
            const bool heavy = U.kind == VFK_KIND_INS && __any_sync(FULL, e0.alt > 1u || e1.alt > 1u);
            uint32_t pk = e0.pk;
            if (m1) {  // move the overflow pass's contributions into lanes the first pass left free
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
            // rare: more contributions than lanes, a multiply supported insertion, or REF == ALT (a read in
            // both lists) -- re-done through the histograms
            if (heavy || __popc(m0) + __popc(m1) > 32 || (sh.R & sh.A)) {
                if constexpr (!LISTED) {
                    if (lane == 0) {
                        warp_list[atomicAdd(n_warp, 1u)] = (uint32_t)u;
                        store_counts(out + u, 0, 0, 0, 0, med2, s2, U.kind | VFK_ST_DEFERRED, (uint32_t)n_cand, 0u);
                    }
                }
                continue;
            }
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
            if (lane == 0) {
                store_counts(out + u, dp, n_amb, ac_ref, ac_alt, med2, s2, U.kind, (uint32_t)n_cand, (uint32_t)n_cover, (uint32_t)n_col);
                if constexpr (LISTED) warp_list[base + j] = LIST_VOID;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// warp per unit on the column store (SNV units): the reads of a column are ONE contiguous run of
// 32-byte sectors -- lane = sector, one 4-byte entry word + the 8-byte header per lane -- the read filter
// verdict, the query position and the overlap partner's whereabouts come with the sector, and the
// partner's base is a register of another lane.  No header gather, no CIGAR, no dependent loads.
// What it cannot serve goes to the read-major kernels through the same device-side lists: units that are
// not SNVs, windows deeper than 64 sectors, a passing read inside a deletion at the column (its
// base-quality filter looks at the next base and at the push timing of its mate), a sector flagged unusable.
// ------------------------------------------------------------------------------------------
#ifndef VFK_COL_MIN_CTAS
#define VFK_COL_MIN_CTAS 8
#endif
struct ColRead {  // one sector as a lane holds it
    uint32_t e;   // the 16-bit entry of the column
    uint32_t h0, h1;
};
// Sector k of the run [off, off + cnt) of a window, position p.  TR = false: the sector form (12 entries + 2 header
// words per 32-byte sector).  TR = true: the window-transposed form of the same run (vfkio_columns_transpose: cnt
// header pairs, then 12 rows of cnt entries) -- NOT YET VERIFIED ON A GPU; selected only by VFK_COL_TRANSPOSED=1
// together with a store built under the same setting (vafator_b200/device.py).
// CP = true (EXPERIMENTAL, VFK_HOST_GATHER=1, host entry point only): the run was gathered on the host for this very
// unit -- cnt header pairs, then the cnt entries of the unit's OWN position (one row), see vfk_api.cu.
template <bool TR, bool CP>
__device__ __forceinline__ ColRead load_col_read(const uint32_t *__restrict__ sectors, uint32_t off, uint32_t cnt, uint32_t k, uint32_t p) {
    ColRead c;
    if constexpr (CP) {
        const unsigned char *run = reinterpret_cast<const unsigned char *>(sectors + (size_t)off * 8u);
        const uint2 h = __ldg(reinterpret_cast<const uint2 *>(run) + k);
        c.e = (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(run + (size_t)cnt * 8u) + k);
        c.h0 = h.x;
        c.h1 = h.y;
    } else if constexpr (TR) {
        const unsigned char *run = reinterpret_cast<const unsigned char *>(sectors + (size_t)off * 8u);
        const uint2 h = __ldg(reinterpret_cast<const uint2 *>(run) + k);
        c.e = (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(run + (size_t)cnt * 8u) + (size_t)p * cnt + k);
        c.h0 = h.x;
        c.h1 = h.y;
    } else {
        const uint32_t *w = sectors + (size_t)(off + k) * 8u;
        c.e = (__ldg(w + (p >> 1)) >> ((p & 1u) * 16u)) & 0xFFFFu;
        const uint2 h = __ldg(reinterpret_cast<const uint2 *>(w + 6));
        c.h0 = h.x;
        c.h1 = h.y;
    }
    return c;
}
// htslib tweak_overlap_quality at one position (both reads show an aligned base): the quality this read keeps
__device__ __forceinline__ int overlap_quality(uint32_t mine, uint32_t theirs, uint32_t h0) {
    const int qi = (int)(mine & 0xFFu), qm = (int)(theirs & 0xFFu);
    const uint32_t sel = (h0 >> 29) & 1u;
    const uint32_t mymul = (h0 >> 30) & 1u ? sel : 1u - sel;  // the earlier record holds the hash slot
    if (((mine ^ theirs) & 0xF00u) == 0u) {
        const int s = min(qi + qm, 200);
        return mymul ? s : 0;
    }
    if (qi > qm) return (qi * 4) / 5;
    if (qi < qm) return 0;
    return mymul ? (qi * 4) / 5 : 0;
}
// one sector -> the packed contribution (Eval.pk / Eval.fl); `partner` is the partner's entry or 0
__device__ __forceinline__ void col_evaluate(const ColRead &c, uint32_t partner, uint32_t p, const vfk_params &P, uint32_t ref_code,
                                             uint32_t alt_code, uint32_t &pk, uint32_t &fl) {
    pk = 0u;
    fl = 0u;
    const uint32_t state = (c.e >> 12) & 3u;
    if (state == 0u) return;
    fl = FL_COVERS;
    if (!((c.h0 >> 28) & 1u)) return;
    fl |= FL_IN_COL;
    if (state != 1u) return;  // reference skip: in the column, nothing else (deletions never get here)
    int q = (int)(c.e & 0xFFu);
    if (((partner >> 12) & 3u) == 1u) q = overlap_quality(c.e, partner, c.h0);
    if (q < P.min_baseq) return;  // pysam pileup_base_qual_skip
    fl |= FL_IN_DP;
    const uint32_t code = (c.e >> 8) & 15u;
    if (code_is_ambiguous((int)code)) fl |= FL_AMBIGUOUS;
    const uint32_t lists = (code == ref_code ? PK_REF : 0u) | (code == alt_code ? PK_ALT : 0u);
    if (lists) {
        const uint32_t below = c.h1 & 0xFFFu & ((1u << p) - 1u);
        const uint32_t ins = ((c.h1 >> 28) & 1u) && p > ((c.h1 >> 12) & 15u) ? (c.h1 >> 16) & 0xFFFu : 0u;
        const uint32_t qpos = (c.h0 & 0xFFFu) + (uint32_t)__popc(below) + ins;
        pk = ((c.h0 >> 12) & 0xFFu) | ((uint32_t)q << 8) | (qpos << 16) | lists;
    }
}

template <int POSB, bool TR, bool MQS, bool CP>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, VFK_COL_MIN_CTAS)
pileup_column_kernel(ReadsDev R, const uint4 *__restrict__ cunits, const ResolvedUnit *__restrict__ res, int64_t n_units, vfk_params P,
                     vfk_counts *__restrict__ out,
                     unsigned long long *__restrict__ work_counter, uint32_t *__restrict__ block_list,
                     uint32_t *__restrict__ n_block, uint32_t *__restrict__ warp_list, uint32_t *__restrict__ n_warp) {
    using L = Layout<POSB>;
    __shared__ uint32_t stage_all[WARPS_PER_CTA][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t *stage = stage_all[warp];
    const int64_t n_warps = (int64_t)gridDim.x * WARPS_PER_CTA;
    int64_t last_base = 0;
    const uint32_t *__restrict__ sectors = R.col_sectors;

    for (;;) {
        unsigned long long base = 0;
        const int64_t left = n_units - last_base;  // as of this warp's previous visit
        const int grab = left > 16 * n_warps ? GRAB : left > 4 * n_warps ? 2 : 1;
        if (lane == 0) base = atomicAdd(work_counter, (unsigned long long)grab);
        base = __shfl_sync(FULL, base, 0);
        if ((int64_t)base >= n_units) break;
        last_base = (int64_t)base;
        const int n_here = (int)min((int64_t)grab, n_units - (int64_t)base);
        uint4 cu = make_uint4(0, 0, 0, 0);  // lane j keeps the column unit of unit base+j
        if (lane < n_here) cu = __ldg(cunits + base + lane);
        // software pipeline: the first 32 sectors of the NEXT unit are requested before the current one is processed
        ColRead next;
        next.e = next.h0 = next.h1 = 0u;
        {
            const uint32_t off0 = __shfl_sync(FULL, cu.x, 0), cnt0 = __shfl_sync(FULL, cu.y, 0), p0 = __shfl_sync(FULL, cu.z, 0);
            if ((uint32_t)lane < cnt0) next = load_col_read<TR, CP>(sectors, off0, cnt0, (uint32_t)lane, p0);
        }
        for (int j = 0; j < n_here; ++j) {
            const int64_t u = (int64_t)base + j;
            const uint32_t off = __shfl_sync(FULL, cu.x, j), cnt = __shfl_sync(FULL, cu.y, j), p = __shfl_sync(FULL, cu.z, j);
            const uint32_t meta = __shfl_sync(FULL, cu.w, j);
            const uint32_t kind = meta & 3u, ref_code = (meta >> 8) & 0xFFu, alt_code = (meta >> 16) & 0xFFu;
            ColRead c0 = next;
            if (j + 1 < n_here) {
                const uint32_t off1 = __shfl_sync(FULL, cu.x, j + 1), cnt1 = __shfl_sync(FULL, cu.y, j + 1);
                const uint32_t p1 = __shfl_sync(FULL, cu.z, j + 1);
                if ((uint32_t)lane < cnt1) next = load_col_read<TR, CP>(sectors, off1, cnt1, (uint32_t)lane, p1);
            }
            int med2[3][2] = {{-1, -1}, {-1, -1}, {-1, -1}};
            uint64_t s2[3] = {0, 0, 0};
            if (kind != VFK_KIND_SNV || cnt > 64u) {  // not for this kernel: the read-major kernels take it
                if (lane == 0) {
                    // located units (host entry point) are routed here, the others by locate_list_kernel
                    const uint32_t n_cand = res ? __ldg(&res[u].hi) - __ldg(&res[u].lo) : 0u;
                    if (n_cand > (uint32_t)WARP_PATH_MAX_CAND) block_list[atomicAdd(n_block, 1u)] = (uint32_t)u;
                    else warp_list[atomicAdd(n_warp, 1u)] = (uint32_t)u;
                    store_counts(out + u, 0, 0, 0, 0, med2, s2, kind | VFK_ST_DEFERRED, cnt, 0u);
                }
                continue;
            }
            if (cnt == 0u) {
                if (lane == 0) store_counts(out + u, 0, 0, 0, 0, med2, s2, kind, 0u, 0u);
                continue;
            }
            if ((uint32_t)lane >= cnt) c0.e = c0.h0 = c0.h1 = 0u;
            ColRead c1;
            c1.e = c1.h0 = c1.h1 = 0u;
            if (cnt > 32u && (uint32_t)lane + 32u < cnt) c1 = load_col_read<TR, CP>(sectors, off, cnt, 32u + (uint32_t)lane, p);
            // overlap partners: sector k + offset of the same run, i.e. an entry another lane holds
            const int off0 = (int)(int8_t)((c0.h0 >> 20) & 0xFFu), off1 = (int)(int8_t)((c1.h0 >> 20) & 0xFFu);
            const int k0p = lane + off0, k1p = 32 + lane + off1;
            const uint32_t a0 = __shfl_sync(FULL, c0.e, k0p & 31), b0 = __shfl_sync(FULL, c1.e, k0p & 31);
            uint32_t partner0 = off0 ? (k0p < 32 ? a0 : b0) : 0u, partner1 = 0u;
            if (cnt > 32u) {
                const uint32_t a1 = __shfl_sync(FULL, c0.e, k1p & 31), b1 = __shfl_sync(FULL, c1.e, k1p & 31);
                partner1 = off1 ? (k1p < 32 ? a1 : b1) : 0u;
            }
            // a passing read inside a deletion, or an unusable sector: the read-major path decides
            const bool in0 = ((c0.e >> 12) & 3u) != 0u && ((c0.h0 >> 28) & 1u), in1 = ((c1.e >> 12) & 3u) != 0u && ((c1.h0 >> 28) & 1u);
            const bool hard = (in0 && (((c0.e >> 12) & 3u) == 2u || (c0.h0 >> 31))) || (in1 && (((c1.e >> 12) & 3u) == 2u || (c1.h0 >> 31)));
            Eval e0, e1;
            e0.alt = e1.alt = 0u;
            col_evaluate(c0, partner0, p, P, ref_code, alt_code, e0.pk, e0.fl);
            col_evaluate(c1, partner1, p, P, ref_code, alt_code, e1.pk, e1.fl);
            const bool k0 = e0.pk != 0u, k1 = e1.pk != 0u;
            const uint32_t m0 = __ballot_sync(FULL, k0), m1 = __ballot_sync(FULL, k1);
            uint32_t pk = e0.pk;
            if (m1) {  // move the second chunk's contributions into lanes the first one left free
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
            if (__any_sync(FULL, hard) || __popc(m0) + __popc(m1) > 32 || (sh.R & sh.A)) {
                if (lane == 0) {
                    warp_list[atomicAdd(n_warp, 1u)] = (uint32_t)u;
                    store_counts(out + u, 0, 0, 0, 0, med2, s2, kind | VFK_ST_DEFERRED, cnt, 0u);
                }
                continue;
            }
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
                if constexpr (MQS) select_compare(mq, sh.R | sh.A, 1u << lane, lt, eq);
                else radix_compare<8>(mq, sh.R | sh.A, lt, eq);
                uint32_t term = radix_finish(mq, lt, eq, sh, w);
                med2[0][0] = ac_ref ? (int)(w & 0xFFFFu) : -1; med2[0][1] = ac_alt ? (int)(w >> 16) : -1;
                radix_compare<8>(bq, sh.R | sh.A, lt, eq);
                term |= radix_finish(bq, lt, eq, sh, w) << 10;
                med2[1][0] = ac_ref ? (int)(w & 0xFFFFu) : -1; med2[1][1] = ac_alt ? (int)(w >> 16) : -1;
                radix_compare<L::POS_BITS>(qp, sh.R | sh.A, lt, eq);
                term |= radix_finish(qp, lt, eq, sh, w) << 20;
                med2[2][0] = ac_ref ? (int)(w & 0xFFFFu) : -1; med2[2][1] = ac_alt ? (int)(w >> 16) : -1;
                if (ac_ref && ac_alt) {
                    const uint32_t sums = __reduce_add_sync(FULL, term);
                    const uint64_t base2 = (uint64_t)(ac_alt * ac_alt + ac_alt);
                    s2[0] = base2 + (sums & 0x3FFu);
                    s2[1] = base2 + ((sums >> 10) & 0x3FFu);
                    s2[2] = base2 + (sums >> 20);
                }
            }
            if (!P.include_ambiguous) dp -= n_amb;  // pileups.py:224-227
            if (lane == 0) store_counts(out + u, dp, n_amb, ac_ref, ac_alt, med2, s2, kind, cnt, (uint32_t)n_cover, (uint32_t)n_col);
        }
    }
}

// ------------------------------------------------------------------------------------------
// warp per listed unit, shared-memory histograms
// ------------------------------------------------------------------------------------------
template <int POSB>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pileup_deep_kernel(ReadsDev R, const ResolvedUnit *__restrict__ res, const uint8_t *__restrict__ ins_seq, vfk_params P,
                   vfk_counts *__restrict__ out, const uint32_t *__restrict__ warp_list, const uint32_t *__restrict__ n_warp,
                   uint32_t *__restrict__ deep_work) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS / 2;  // u16 bins
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *ref_slot = smem + (size_t)warp * 2 * WORDS;
    uint32_t *alt_slot = ref_slot + WORDS;
    __shared__ __align__(128) uint4 s_hot[WARPS_PER_CTA][2][32];
    __shared__ __align__(8) uint64_t s_bar[WARPS_PER_CTA][2];
    StageRing ring;
    ring.hot = &s_hot[warp][0][0];
    ring.bar = &s_bar[warp][0];
    ring.parity = 0u;
    if (lane == 0) {
        mbar_init(ring.bar + 0, 1);
        mbar_init(ring.bar + 1, 1);
        mbar_fence_init();
    }
    __syncwarp();
    const uint32_t n = *n_warp;
    // listed unit k goes to warp k first (no atomics when the list is short or empty), the rest through a counter
    const uint32_t n_warps = gridDim.x * WARPS_PER_CTA;
    uint32_t k = blockIdx.x * WARPS_PER_CTA + warp;
    for (;;) {
        if (k >= n) break;
        const uint32_t u = warp_list[k];
        if (u != LIST_VOID) {
            Unit U;
            unit_from(res[u], ins_seq, U);
            unit_by_histogram<POSB, true>(R, P, U, ref_slot, alt_slot, lane, out + u, &ring);
        }
        if (lane == 0) k = n_warps + atomicAdd(deep_work, 1u);
        k = __shfl_sync(FULL, k, 0);
    }
}

// ------------------------------------------------------------------------------------------
// CTA per unit, u32 counters: columns deeper than the warp path's u16 bins allow
// ------------------------------------------------------------------------------------------
template <int POSB>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
pileup_block_kernel(ReadsDev R, const ResolvedUnit *__restrict__ res, const uint8_t *__restrict__ ins_seq, vfk_params P,
                    vfk_counts *__restrict__ out, const uint32_t *__restrict__ deferred,
                    const uint32_t *__restrict__ n_deferred) {
    using L = Layout<POSB>;
    constexpr int WORDS = L::BINS;  // u32 bins
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ int red[6];
    __shared__ uint32_t red_status;
    __shared__ int s_med2[3][2];
    __shared__ unsigned long long s_s2[3];
    uint32_t *ref_slot = smem, *alt_slot = smem + WORDS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nd = *n_deferred;
    for (uint32_t d = blockIdx.x; d < nd; d += gridDim.x) {
        const int64_t u = deferred[d];
        Unit U;
        unit_from(res[u], ins_seq, U);
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
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// host-side launchers (called from vfk_api.cu)
// ------------------------------------------------------------------------------------------
cudaError_t launch_locate(const ReadsDev &R, const UnitsDev &V, void *resolved, unsigned long long *cand_sum, cudaStream_t stream,
                          int *launches) {
    if (V.n <= 0) return cudaSuccess;
    uint4 *cunits = R.col_off ? reinterpret_cast<uint4 *>((char *)resolved + (size_t)V.n * sizeof(ResolvedUnit)) : nullptr;
    locate_kernel<<<(unsigned)((V.n + 255) / 256), 256, 0, stream>>>(R, V, (ResolvedUnit *)resolved, cunits, cand_sum);
    ++*launches;
    return cudaGetLastError();
}

// set by the host entry point (vfk_api.cu) around a launch whose column runs it has gathered unit by unit
thread_local bool column_runs_gathered = false;

// the staged variants are chosen per call (tools/ab_staged.py switches them inside one process)
static bool env_flag(const char *name) {
    const char *v = getenv(name);
    return v && v[0] == '1';
}

// scratch words (32 bytes, device): @0 u64 unit work counter, @8 u32 block-list length, @12 u32 warp-list
// length, @16 u32 deep-kernel work counter, @24 u64 candidate sum (locate pass).  lists: 2 * n_units u32.
template <int POSB>
static cudaError_t launch_posb(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, vfk_counts *out, ResolvedUnit *res,
                               void *counters, uint32_t *lists, int sm_count, cudaStream_t stream, int *launches, bool located) {
    using L = Layout<POSB>;
    const size_t smem_deep = (size_t)WARPS_PER_CTA * 2 * (L::BINS / 2) * sizeof(uint32_t);
    const size_t smem_block = (size_t)2 * L::BINS * sizeof(uint32_t);
    unsigned long long *work = (unsigned long long *)counters;
    uint32_t *n_block = (uint32_t *)((char *)counters + 8), *n_warp = (uint32_t *)((char *)counters + 12);
    uint32_t *deep_work = (uint32_t *)((char *)counters + 16);
    uint32_t *block_list = lists, *warp_list = lists + V.n;
    cudaError_t e;
    {
        // static (staging ring) + dynamic (histograms) shared memory beyond 48 KB needs an opt-in
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, pileup_deep_kernel<POSB>);
        if (e != cudaSuccess) return e;
        if (smem_deep + fa.sharedSizeBytes > 48 * 1024) {
            e = cudaFuncSetAttribute(pileup_deep_kernel<POSB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_deep);
            if (e != cudaSuccess) return e;
        }
    }
    int ctas_per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, pileup_warp_kernel<POSB, false>, WARPS_PER_CTA * 32, 0);
    if (e != cudaSuccess) return e;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    // persistent grid: a whole number of waves, never more CTAs than there is work
    int64_t want = (V.n + (int64_t)GRAB * WARPS_PER_CTA - 1) / ((int64_t)GRAB * WARPS_PER_CTA);
    int64_t grid = (int64_t)sm_count * ctas_per_sm;
    if (want < grid) grid = want < 1 ? 1 : want;
    // EXPERIMENTAL (VFK_LIST_SHALLOW=1): the column kernel's list goes through the shallow warp kernel first; its work
    // counter is the u64 at byte 24 of the scratch words (the locate pass's candidate sum has been consumed by then)
    const bool list_shallow = env_flag("VFK_LIST_SHALLOW");
    e = cudaMemsetAsync(counters, 0, list_shallow ? 32 : 24, stream);
    if (e != cudaSuccess) return e;
    if (R.col_off) {  // column store: SNV units from contiguous sector runs, the rest through the lists
        // layout of the store the caller built: sector form unless both sides were told otherwise (see load_col_read)
        const bool transposed = env_flag("VFK_COL_TRANSPOSED");
        int col_ctas = 0;
        const bool mq_select = env_flag("VFK_MQ_SELECT");
        auto column_kernel = pileup_column_kernel<POSB, false, false, false>;
        if (column_runs_gathered && mq_select) column_kernel = pileup_column_kernel<POSB, false, true, true>;
        else if (column_runs_gathered) column_kernel = pileup_column_kernel<POSB, false, false, true>;
        else if (transposed && mq_select) column_kernel = pileup_column_kernel<POSB, true, true, false>;
        else if (transposed) column_kernel = pileup_column_kernel<POSB, true, false, false>;
        else if (mq_select) column_kernel = pileup_column_kernel<POSB, false, true, false>;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&col_ctas, column_kernel, WARPS_PER_CTA * 32, 0);
        if (e != cudaSuccess) return e;
        int64_t col_grid = (int64_t)sm_count * std::max(col_ctas, 1);
        if (want < col_grid) col_grid = want < 1 ? 1 : want;
        uint4 *cunits = reinterpret_cast<uint4 *>((char *)res + (size_t)V.n * sizeof(ResolvedUnit));
        if (!located) {  // lazy locate: every unit gets its window now, only the listed ones are searched afterwards
            column_units_kernel<<<(unsigned)((V.n + 255) / 256), 256, 0, stream>>>(R, V, cunits);
            ++*launches;
        }
        column_kernel<<<(unsigned)col_grid, WARPS_PER_CTA * 32, 0, stream>>>(R, cunits, located ? res : nullptr, V.n, P, out, work, block_list,
                                                                             n_block, warp_list, n_warp);
        if (!located) {
            ++*launches;
            locate_list_kernel<<<(unsigned)((V.n + 255) / 256), 256, 0, stream>>>(R, V, res, warp_list, n_warp, block_list, n_block);
        }
        if (list_shallow) {
            ++*launches;
            pileup_warp_kernel<POSB, true><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, stream>>>(
                R, res, V.n, V.ins_seq, P, out, (unsigned long long *)((char *)counters + 24), block_list, n_block, warp_list, n_warp);
        }
    } else {
        pileup_warp_kernel<POSB, false><<<(unsigned)grid, WARPS_PER_CTA * 32, 0, stream>>>(R, res, V.n, V.ins_seq, P, out, work,
                                                                                           block_list, n_block, warp_list, n_warp);
    }
    ++*launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    int deep_ctas = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&deep_ctas, pileup_deep_kernel<POSB>, WARPS_PER_CTA * 32, smem_deep);
    if (e != cudaSuccess) return e;
    if (deep_ctas < 1) deep_ctas = 1;
    int64_t deep_grid = std::min<int64_t>((int64_t)sm_count * deep_ctas, (V.n + WARPS_PER_CTA - 1) / WARPS_PER_CTA);
    pileup_deep_kernel<POSB><<<(unsigned)deep_grid, WARPS_PER_CTA * 32, smem_deep, stream>>>(R, res, V.ins_seq, P, out, warp_list,
                                                                                             n_warp, deep_work);
    ++*launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (R.n_reads > WARP_PATH_MAX_CAND) {  // only then can a column be too deep for the warp kernels
        pileup_block_kernel<POSB><<<(unsigned)sm_count, WARPS_PER_CTA * 32, smem_block, stream>>>(R, res, V.ins_seq, P, out,
                                                                                                  block_list, n_block);
        ++*launches;
        e = cudaGetLastError();
    }
    return e;
}

static cudaError_t launch_pileup_sized(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int max_l_qseq, vfk_counts *out,
                                       void *resolved, void *counters, uint32_t *lists, int sm_count, cudaStream_t stream,
                                       int *launches, bool located) {
    ResolvedUnit *res = (ResolvedUnit *)resolved;
    // read positions run 0 .. l_qseq (a read hanging in a trailing deletion reports l_qseq)
    if (max_l_qseq < 256) return launch_posb<256>(R, V, P, out, res, counters, lists, sm_count, stream, launches, located);
    if (max_l_qseq < 512) return launch_posb<512>(R, V, P, out, res, counters, lists, sm_count, stream, launches, located);
    if (max_l_qseq < 1024) return launch_posb<1024>(R, V, P, out, res, counters, lists, sm_count, stream, launches, located);
    return launch_posb<4096>(R, V, P, out, res, counters, lists, sm_count, stream, launches, located);
}

// the pileup kernels on units that launch_locate (or the host) has resolved
cudaError_t launch_pileup_resolved(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int max_l_qseq, vfk_counts *out,
                                   void *resolved, void *counters, uint32_t *lists, int sm_count, cudaStream_t stream,
                                   int *launches) {
    return launch_pileup_sized(R, V, P, max_l_qseq, out, resolved, counters, lists, sm_count, stream, launches, true);
}

cudaError_t launch_pileup(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int max_l_qseq, vfk_counts *out,
                          void *resolved_scratch, void *counters, uint32_t *lists, int sm_count, cudaStream_t stream,
                          int *launches) {
    if (V.n <= 0) return cudaSuccess;
    if (R.col_off && !getenv("VFK_EAGER_LOCATE"))  // column store: the position index is searched for the listed units only
        return launch_pileup_sized(R, V, P, max_l_qseq, out, resolved_scratch, counters, lists, sm_count, stream, launches, false);
    cudaError_t e = launch_locate(R, V, resolved_scratch, nullptr, stream, launches);
    if (e != cudaSuccess) return e;
    return launch_pileup_resolved(R, V, P, max_l_qseq, out, resolved_scratch, counters, lists, sm_count, stream, launches);
}

}  // namespace vfk

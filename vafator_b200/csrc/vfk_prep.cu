// vfk_prep.cu -- device-side read preparation: everything the pileup kernels need beyond the canonical arrays
// is derived here, on the device, inside every call (the reference starts at bam.pileup(), vafator/pileups.py:71-80,
// with nothing but the decoded records either).
//
//  offsets_*     (only for a batch that leaves its pool offsets NULL) exclusive prefix sums of n_cigar, (l_qseq + 1) / 2, l_qseq.
//  prep_kernel   a warp per 128 reads, four consecutive reads per lane: reference end (htslib bam_endpos) and CIGAR shape word
//                per read (vfk_device.cuh), the maximum of the end keys per chunk of 32 reads and per block of PREP_BLOCK reads;
//                checks the batch (sorted, starts >= 0, CIGAR inside its pool, lengths in range) and raises
//                VFK_ST_BADBATCH otherwise.
//  scan_kernel   inclusive running maximum over the block maxima: the top level of the position index.  A column's
//                first overlapping read is the first read whose end exceeds it -- found by a binary search over the
//                blocks, a walk over the chunk maxima of one block and one pass over a chunk (locate_kernel, vfk_pileup.cu).
//  link kernels  the mate-linking passes of vfk_link.cuh as separate launches, for the list-shaped auxiliary
//                (vfk_pileup_values) and the debug hook that links every unit; the pileup call runs them fused in tail_kernel.
#include "vfk_link.cuh"

namespace vfk {

// ---- CIGAR -> span + shape word (same classification for every kernel) ----
struct CigarInfo {
    uint32_t span;
    uint32_t ev;
};
__device__ __forceinline__ bool op_is_match(uint32_t op) { return op == OP_M || op == OP_EQ || op == OP_X; }
__device__ __forceinline__ bool op_consumes_ref(uint32_t op) { return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X; }

// One pass over the CIGAR: the reference span, and whether it reads [H*][S] M [ (I|D|N) M ] [S][H*] with every length in
// range -- then the packed event word, else EV_GENERAL.  A small state machine (a complex read costs its warp one pass of
// this loop with a few lanes active: the words are loaded once).
//   0 start: H -> 0, S -> 1, M -> 2     1 after the leading clip: M -> 2      2 after the first block: I/D/N -> 3, S -> 5, H -> 6
//   3 after the event: M -> 4           4 after the second block: S -> 5, H -> 6     5 after the trailing clip: H -> 6     6: H -> 6
__device__ __forceinline__ CigarInfo cigar_single(uint32_t w, int32_t l_qseq) {  // nine reads in ten: one op
    CigarInfo o;
    const uint32_t op = w & 15u, len = w >> 4;
    o.span = op_consumes_ref(op) ? len : 0u;
    if (!op_is_match(op)) o.ev = EV_GENERAL;
    else if ((int64_t)len == (int64_t)l_qseq) o.ev = EV_SIMPLE;
    else o.ev = (len == 0u || len >= 4096u || l_qseq > VFK_MAX_READ_LEN) ? EV_GENERAL : (len << 12);
    return o;
}
// The state machine as a table in constant memory: no branch in a step (a walk runs with whatever lanes hold a multi-op CIGAR).
//   CIGAR_STEP[op]: bits 0-23 the next state for each of the eight states (three bits each), bit 24 the op consumes the
//   reference, bit 25 it is a soft clip, bit 26 an aligned block (M = X), bit 27 an event (I D N)
constexpr uint32_t pack_states(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
    return (uint32_t)s0 | (uint32_t)s1 << 3 | (uint32_t)s2 << 6 | (uint32_t)s3 << 9 | (uint32_t)s4 << 12 | (uint32_t)s5 << 15 | (uint32_t)s6 << 18 | (uint32_t)s7 << 21;
}
constexpr uint32_t STEP_REF = 1u << 24, STEP_CLIP = 1u << 25, STEP_BLOCK = 1u << 26, STEP_EVENT = 1u << 27;
constexpr uint32_t NEXT_H = pack_states(0, 7, 6, 7, 6, 6, 6, 7), NEXT_S = pack_states(1, 7, 5, 7, 5, 7, 7, 7) | STEP_CLIP;
constexpr uint32_t NEXT_M = pack_states(2, 2, 7, 4, 7, 7, 7, 7) | STEP_BLOCK | STEP_REF, NEXT_E = pack_states(7, 7, 3, 7, 7, 7, 7, 7) | STEP_EVENT;
constexpr uint32_t NEXT_NONE = pack_states(7, 7, 7, 7, 7, 7, 7, 7);
__constant__ uint32_t CIGAR_STEP[16] = {
    /* M */ NEXT_M, /* I */ NEXT_E, /* D */ NEXT_E | STEP_REF, /* N */ NEXT_E | STEP_REF, /* S */ NEXT_S, /* H */ NEXT_H, /* P */ NEXT_NONE,
    /* = */ NEXT_M, /* X */ NEXT_M, NEXT_NONE, NEXT_NONE, NEXT_NONE, NEXT_NONE, NEXT_NONE, NEXT_NONE, NEXT_NONE};
struct WalkState {
    uint32_t span, lead, m1, ev, elen, state;
};
__device__ __forceinline__ void cigar_step(WalkState &s, uint32_t w) {
    const uint32_t op = w & 15u, len = w >> 4;
    const uint32_t t = CIGAR_STEP[op];
    s.span += (t & STEP_REF) ? len : 0u;
    uint32_t next = (t >> (3u * s.state)) & 7u;
    if ((t & (STEP_BLOCK | STEP_EVENT)) && len == 0u) next = 7u;  // an empty block / event is not of this shape
    s.lead = ((t & STEP_CLIP) && s.state == 0u) ? len : s.lead;
    s.m1 = ((t & STEP_BLOCK) && s.state <= 1u) ? len : s.m1;
    const bool event = next == 3u;  // I / D / N right after the first block
    s.ev = event ? (op == OP_I ? (uint32_t)EVK_INS : op == OP_D ? (uint32_t)EVK_DEL : (uint32_t)EVK_SKIP) : s.ev;
    s.elen = event ? len : s.elen;
    s.state = next;
}
// n >= 2 ops.  The first four are requested together and stepped through without a loop (nearly every multi-op CIGAR has two
// to four ops; a lane with fewer just sits out the steps it does not need), the rest in a loop.
__device__ __forceinline__ CigarInfo cigar_walk(const uint32_t *__restrict__ c, int n, int32_t l_qseq) {
    const uint32_t w0 = n > 0 ? __ldg(c) : 0u, w1 = n > 1 ? __ldg(c + 1) : 0u, w2 = n > 2 ? __ldg(c + 2) : 0u, w3 = n > 3 ? __ldg(c + 3) : 0u;
    WalkState s;
    s.span = s.lead = s.m1 = s.elen = s.state = 0u;
    s.ev = EVK_NONE;
    if (n > 0) cigar_step(s, w0);
    if (n > 1) cigar_step(s, w1);
    if (n > 2) cigar_step(s, w2);
    if (n > 3) cigar_step(s, w3);
#pragma unroll 1
    for (int k = 4; k < n; ++k) cigar_step(s, __ldg(c + k));
    CigarInfo o;
    o.span = s.span;
    const bool shaped = (s.state == 2u || (s.state >= 4u && s.state != 7u)) && s.m1 != 0u && s.m1 < 4096u && s.lead < 4096u && s.elen < 64u &&
                        l_qseq <= VFK_MAX_READ_LEN;
    o.ev = shaped ? (s.lead | (s.m1 << 12) | (s.ev << 24) | (s.elen << 26)) : EV_GENERAL;
    return o;
}
__device__ __forceinline__ CigarInfo cigar_info(const uint32_t *__restrict__ c, int n, int32_t l_qseq) {
    return n == 1 ? cigar_single(__ldg(c), l_qseq) : cigar_walk(c, n, l_qseq);
}

__device__ __forceinline__ int32_t contig_of(const int64_t *__restrict__ contig_off, int32_t n_contigs, int64_t i) {
    int32_t a = 0, b = n_contigs;  // last t with contig_off[t] <= i
    while (b - a > 1) {
        const int32_t m = (a + b) >> 1;
        if (__ldg(contig_off + m) <= i) a = m; else b = m;
    }
    return a;
}

// One CTA = PREP_BLOCK (1024) reads = 8 warps of 128 reads; a warp works alone (no CTA barrier after the first line): four
// CHUNKS of 32 consecutive reads, two at a time -- both chunks' loads (start, CIGAR count / offset, read length) are in flight
// before either is consumed.  Nine reads in ten have a one-op CIGAR and are finished on the spot; the others are put on the
// warp's shared-memory list and walked afterwards, ~10 lanes at once instead of two or three per chunk, their first four ops
// requested together.  Writes {end, shape word} per read, the maximum end key (tid << 32 | end) per chunk and per block:
// the two levels of the position index.  The last warp of a CTA to finish folds the eight warp keys into the block key.
struct PrepRead {
    int32_t p, n, lq;
    int64_t off;
};
__device__ __forceinline__ PrepRead prep_load(const ReadsDev &R, int64_t i, bool in) {
    PrepRead r;
    r.p = in ? __ldg(R.pos + i) : 0;
    r.n = in ? __ldg(R.n_cigar + i) : 0;
    r.off = in ? __ldg(R.cigar_off + i) : 0;
    r.lq = in ? __ldg(R.l_qseq + i) : 0;
    return r;
}
// negative start / count / offset, CIGAR outside its pool, read longer than the position histograms cover
__device__ __forceinline__ bool prep_malformed(const ReadsDev &R, const PrepRead &r) {
    return (r.p | r.n) < 0 || (uint64_t)r.off + (uint64_t)(uint32_t)r.n > (uint64_t)R.n_cigar_words || (uint32_t)r.lq > (uint32_t)VFK_MAX_READ_LEN;
}
// a warp whose reads straddle a contig boundary (one in tens of thousands): every read looks its own contig up and walks its own CIGAR
__device__ __noinline__ long long straddling_read(const ReadsDev &R, int64_t i, bool in, uint2 *__restrict__ ee_out, uint32_t *__restrict__ batch_status) {
    long long key = LLONG_MIN;
    if (in) {
        const PrepRead r = prep_load(R, i, true);
        const int32_t tid = contig_of(R.contig_off, R.n_contigs, i);
        bool bad = prep_malformed(R, r) || (i > __ldg(R.contig_off + tid) && __ldg(R.pos + i - 1) > r.p);
        CigarInfo ci;
        ci.span = 0u;
        ci.ev = EV_GENERAL;
        if (!bad && r.n > 0) ci = cigar_info(R.cigar + r.off, r.n, r.lq);
        if ((uint64_t)(uint32_t)r.p + ci.span > 0x7FFFFFFFull) { bad = true; ci.span = 0u; }
        const int32_t end = r.p + (int32_t)ci.span;
        ee_out[i] = make_uint2((uint32_t)end, ci.ev);
        if (bad) atomicOr(batch_status, VFK_ST_BADBATCH);
        key = ((long long)tid << 32) | (long long)(uint32_t)end;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
    return key;
}

constexpr int PREP_WARPS = PREP_THREADS / 32;           // 8
constexpr int PREP_WARP_READS = PREP_BLOCK / PREP_WARPS;  // 128: four chunks
// ---- the two ways a warp gets through its 128 reads.  Both leave: {end, shape} of every one-op read written, the reads with
// more ops on the warp's list (s_list / cnt), the maximum end of the finished reads of each chunk in s_cend.
//
// scalar: lane = read, four chunks two at a time; any alignment, any tail.
__device__ __forceinline__ void prep_chunks_scalar(const ReadsDev &R, int64_t w0, int lane, uint32_t lt_mask, int64_t first_of_contig,
                                                   uint2 *__restrict__ ee_out, uint8_t *s_list, int32_t *s_cend, uint32_t &cnt, bool &bad) {
    PrepRead r[2];
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
#pragma unroll
        for (int j = 0; j < 2; ++j) r[j] = prep_load(R, w0 + (2 * round + j) * 32 + lane, w0 + (2 * round + j) * 32 + lane < R.n_reads);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = 2 * round + j;
            const int64_t i = w0 + c * 32 + lane;
            const bool in = i < R.n_reads;
            // the start before this one (coordinate sorted inside the contig): the lane below has it, lane 0 reads it
            int32_t prev = __shfl_up_sync(FULL, r[j].p, 1);
            if (lane == 0) prev = (in && i > first_of_contig) ? __ldg(R.pos + i - 1) : INT32_MIN;
            int32_t end = -1;  // a read whose CIGAR is still to be walked does not vote
            bool multi = false;
            if (in) {
                bool b = prep_malformed(R, r[j]) || (prev > r[j].p && i > first_of_contig);
                multi = !b && r[j].n > 1;
                if (!multi) {
                    CigarInfo ci;
                    ci.span = 0u;
                    ci.ev = EV_GENERAL;
                    if (!b && r[j].n == 1) ci = cigar_single(__ldg(R.cigar + r[j].off), r[j].lq);
                    if ((uint64_t)(uint32_t)r[j].p + ci.span > 0x7FFFFFFFull) { b = true; ci.span = 0u; }
                    end = r[j].p + (int32_t)ci.span;
                    ee_out[i] = make_uint2((uint32_t)end, ci.ev);
                }
                bad = bad || b;
            }
            const uint32_t mm = __ballot_sync(FULL, multi);
            if (multi) s_list[cnt + __popc(mm & lt_mask)] = (uint8_t)(c * 32 + lane);
            cnt += __popc(mm);
            const int32_t e = __reduce_max_sync(FULL, end);
            if (lane == 0) s_cend[c] = e;
        }
    }
}
// vector: lane = four consecutive reads -- one 128-bit load per array (two for the 64-bit CIGAR offsets) serves the warp's 128
// reads, two 128-bit stores write their {end, shape} pairs.  Needs all 128 reads in range and 16-byte aligned arrays.
__device__ __forceinline__ void prep_chunks_vector(const ReadsDev &R, int64_t w0, int lane, uint32_t lt_mask, int64_t first_of_contig,
                                                   uint2 *__restrict__ ee_out, uint8_t *s_list, int32_t *s_cend, uint32_t &cnt, bool &bad) {
    const int64_t base = w0 + 4 * lane;
    const int4 p4 = __ldg(reinterpret_cast<const int4 *>(R.pos + base));
    const int4 n4 = __ldg(reinterpret_cast<const int4 *>(R.n_cigar + base));
    const int4 l4 = __ldg(reinterpret_cast<const int4 *>(R.l_qseq + base));
    const longlong2 o01 = __ldg(reinterpret_cast<const longlong2 *>(R.cigar_off + base));
    const longlong2 o23 = __ldg(reinterpret_cast<const longlong2 *>(R.cigar_off + base + 2));
    PrepRead r[4];
    r[0].p = p4.x; r[1].p = p4.y; r[2].p = p4.z; r[3].p = p4.w;
    r[0].n = n4.x; r[1].n = n4.y; r[2].n = n4.z; r[3].n = n4.w;
    r[0].lq = l4.x; r[1].lq = l4.y; r[2].lq = l4.z; r[3].lq = l4.w;
    r[0].off = o01.x; r[1].off = o01.y; r[2].off = o23.x; r[3].off = o23.y;
    // the start before the lane's first read: the lane below has it, lane 0 reads it
    int32_t prev = __shfl_up_sync(FULL, p4.w, 1);
    if (lane == 0) prev = w0 > first_of_contig ? __ldg(R.pos + w0 - 1) : INT32_MIN;
    bool b[4], multi[4];
    uint32_t word[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        b[j] = prep_malformed(R, r[j]) || ((j ? r[j - 1].p : prev) > r[j].p && base + j > first_of_contig);
        multi[j] = !b[j] && r[j].n > 1;
        word[j] = (!b[j] && r[j].n == 1) ? __ldg(R.cigar + r[j].off) : 0xFFFFFFFFu;  // the four CIGAR words in flight together
    }
    uint32_t out[8];
    int32_t e = -1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        CigarInfo ci;
        ci.span = 0u;
        ci.ev = EV_GENERAL;
        if (!b[j] && r[j].n == 1) ci = cigar_single(word[j], r[j].lq);
        if ((uint64_t)(uint32_t)r[j].p + ci.span > 0x7FFFFFFFull) { b[j] = true; ci.span = 0u; }
        const int32_t end = r[j].p + (int32_t)ci.span;
        out[2 * j] = (uint32_t)end;  // a multi-op read gets a placeholder here; its walker overwrites it after the warp barrier
        out[2 * j + 1] = ci.ev;
        e = multi[j] ? e : max(e, end);
        bad = bad || b[j];
        const uint32_t mm = __ballot_sync(FULL, multi[j]);
        if (multi[j]) s_list[cnt + __popc(mm & lt_mask)] = (uint8_t)(4 * lane + j);
        cnt += __popc(mm);
    }
    uint4 *o = reinterpret_cast<uint4 *>(ee_out + base);
    o[0] = make_uint4(out[0], out[1], out[2], out[3]);
    o[1] = make_uint4(out[4], out[5], out[6], out[7]);
    // a chunk = 32 reads = 8 lanes: maximum over each group of eight lanes
#pragma unroll
    for (int k = 1; k < 8; k <<= 1) e = max(e, __shfl_xor_sync(FULL, e, k));
    if ((lane & 7) == 0) s_cend[lane >> 3] = e;
}

#ifndef VFK_PREP_MIN_CTAS
#define VFK_PREP_MIN_CTAS 5  // 48 registers, 40 warps per SM
#endif
// VEC: the four per-read arrays the kernel streams are 16-byte aligned (checked by the launcher)
template <bool VEC>
__global__ void __launch_bounds__(PREP_THREADS, VFK_PREP_MIN_CTAS)
prep_kernel(const __grid_constant__ ReadsDev R, uint2 *__restrict__ ee_out, int64_t *__restrict__ chunk_max, int64_t *__restrict__ blk_max,
            uint32_t *__restrict__ batch_status) {
    __shared__ uint8_t s_list[PREP_WARPS][PREP_WARP_READS];  // per warp: its reads with more than one CIGAR op
    __shared__ int32_t s_cend[PREP_WARPS][4];                // per warp: maximum end of each of its chunks
    __shared__ long long s_wkey[PREP_WARPS];                 // per warp: its maximum end key
    __shared__ uint32_t s_arrived;
    if (threadIdx.x == 0) s_arrived = 0u;
    __syncthreads();  // the only CTA-wide barrier, before anything is in flight
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int64_t w0 = (int64_t)blockIdx.x * PREP_BLOCK + (int64_t)warp * PREP_WARP_READS;  // this warp's reads [w0, w0 + 128)
    const int64_t w1 = min(w0 + PREP_WARP_READS, R.n_reads) - 1;
    long long wkey = LLONG_MIN;
    if (w0 < R.n_reads) {
        int32_t t = 0;  // contig of the warp's first / last read
        if (lane < 2) t = contig_of(R.contig_off, R.n_contigs, lane ? w1 : w0);
        const int32_t tid0 = __shfl_sync(FULL, t, 0);
        if (tid0 != __shfl_sync(FULL, t, 1)) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int64_t i = w0 + c * 32 + lane;
                const long long key = straddling_read(R, i, i < R.n_reads, ee_out, batch_status);
                if (lane == 0 && (w0 + c * 32) < R.n_reads) chunk_max[(w0 >> 5) + c] = key;
                wkey = max(wkey, key);
            }
        } else {
            const int64_t first_of_contig = __ldg(R.contig_off + tid0);
            uint32_t cnt = 0u;
            bool bad = false;
            if (VEC && w0 + PREP_WARP_READS <= R.n_reads) prep_chunks_vector(R, w0, lane, lt_mask, first_of_contig, ee_out, s_list[warp], s_cend[warp], cnt, bad);
            else prep_chunks_scalar(R, w0, lane, lt_mask, first_of_contig, ee_out, s_list[warp], s_cend[warp], cnt, bad);
            __syncwarp();
            // the multi-op CIGARs of the warp's 128 reads, one per lane
            for (uint32_t k = lane; k < cnt; k += 32) {
                const uint32_t idx = s_list[warp][k];
                const int64_t i = w0 + idx;
                const int32_t p = __ldg(R.pos + i);
                CigarInfo ci = cigar_walk(R.cigar + __ldg(R.cigar_off + i), __ldg(R.n_cigar + i), __ldg(R.l_qseq + i));
                if ((uint64_t)(uint32_t)p + ci.span > 0x7FFFFFFFull) { bad = true; ci.span = 0u; }
                const int32_t end = p + (int32_t)ci.span;
                ee_out[i] = make_uint2((uint32_t)end, ci.ev);
                atomicMax(&s_cend[warp][idx >> 5], end);
            }
            if (__any_sync(FULL, bad) && lane == 0) atomicOr(batch_status, VFK_ST_BADBATCH);
            __syncwarp();
            const int32_t e = lane < 4 ? s_cend[warp][lane] : -1;
            const long long key = e < 0 ? LLONG_MIN : (((long long)tid0 << 32) | (long long)(uint32_t)e);
            if (lane < 4 && (w0 + lane * 32) < R.n_reads) chunk_max[(w0 >> 5) + lane] = key;
            const int32_t m = __reduce_max_sync(FULL, e);
            wkey = m < 0 ? LLONG_MIN : (((long long)tid0 << 32) | (long long)(uint32_t)m);
        }
    }
    // block key: the last warp to arrive folds the eight warp keys
    if (lane == 0) {
        s_wkey[warp] = wkey;
        __threadfence_block();
        if (atomicAdd(&s_arrived, 1u) == PREP_WARPS - 1) {
            __threadfence_block();
            long long m = LLONG_MIN;
#pragma unroll
            for (int w = 0; w < PREP_WARPS; ++w) m = max(m, ((volatile long long *)s_wkey)[w]);
            blk_max[blockIdx.x] = m;
        }
    }
}
static_assert(PREP_WARP_READS == 128, "a warp owns four chunks");

// blk_incl[0] = LLONG_MIN, blk_incl[b + 1] = max(blk_max[0..b]).  CTA c owns entries [1024 c, 1024 c + 1024): it first folds
// every entry before its own (a coalesced sweep -- a maximum needs no order, and the whole array is a few hundred KB of L2),
// then scans its 1 024 entries with that carry.  One launch, no chain between the CTAs.
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const int64_t *__restrict__ blk_max, int64_t n_blocks, int64_t *__restrict__ blk_incl) {
    __shared__ long long s_warp[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t first = (int64_t)blockIdx.x * SCAN_THREADS;
    long long carry = LLONG_MIN;
    for (int64_t k = t; k < first; k += SCAN_THREADS) carry = max(carry, (long long)__ldg(blk_max + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) carry = max(carry, __shfl_xor_sync(FULL, carry, o));
    if (lane == 0) s_warp[warp] = carry;
    __syncthreads();
    carry = s_warp[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) carry = max(carry, __shfl_xor_sync(FULL, carry, o));  // every thread: maximum of all entries before the CTA's
    __syncthreads();
    const int64_t k = first + t;
    long long v = k < n_blocks ? (long long)__ldg(blk_max + k) : LLONG_MIN;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v = max(v, u);
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    long long before = carry;
    for (int w = 0; w < warp; ++w) before = max(before, s_warp[w]);
    v = max(v, before);
    if (k < n_blocks) blk_incl[k + 1] = v;
    if (k == 0) blk_incl[0] = LLONG_MIN;
}

// ------------------------------------------------------------------------------------------
// mate linking as three launches (the list-shaped auxiliary vfk_pileup_values, and the debug hook that links every unit);
// the pileup call itself runs the passes fused inside tail_kernel (vfk_pileup.cu)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
link_mark_kernel(ReadsDev R, vfk_params P, const ResolvedUnit *__restrict__ res, const uint32_t *__restrict__ list, const uint32_t *__restrict__ n_list,
                 int64_t n_all, LinkScratch L) {
    link_mark(R, P, res, list, list ? (int64_t)*n_list : n_all, L);
}
template <int PHASE>
__global__ void __launch_bounds__(256) link_reads_kernel(ReadsDev R, vfk_params P, LinkScratch L) {
    link_reads<PHASE>(R, P, L);
}

// ------------------------------------------------------------------------------------------
// implicit pool offsets: a batch whose CIGAR / sequence / quality pools are contiguous in read order (what a decoder that
// appends record by record produces) may leave cigar_off / seq_off / qual_off NULL; they are then the exclusive prefix sums of
// n_cigar, (l_qseq + 1) / 2 and l_qseq, derived here.  The host entry point saves 8 bytes per read of bulk copy and the in-place
// reads of two offset arrays that way.  Three launches: per-block sums, a scan over the blocks, per-block apply.
// ------------------------------------------------------------------------------------------
constexpr int OFF_THREADS = 256, OFF_PER = 4;  // a block of 1024 reads, four consecutive reads per thread
struct Off3 {
    long long c, s, q;
};
__device__ __forceinline__ Off3 off3_add(const Off3 &a, const Off3 &b) { return Off3{a.c + b.c, a.s + b.s, a.q + b.q}; }
__device__ __forceinline__ Off3 off3_shfl_up(const Off3 &v, int d) {
    return Off3{__shfl_up_sync(FULL, v.c, d), __shfl_up_sync(FULL, v.s, d), __shfl_up_sync(FULL, v.q, d)};
}
__device__ __forceinline__ Off3 off3_of(const ReadsDev &R, int64_t i) {
    Off3 v{0, 0, 0};
    if (i < R.n_reads) {
        const long long n = max(__ldg(R.n_cigar + i), 0), l = max(__ldg(R.l_qseq + i), 0);  // negative lengths: prep_kernel refuses the batch
        v.c = n; v.s = (l + 1) >> 1; v.q = l;
    }
    return v;
}
// inclusive scan of one value per thread over the CTA; returns the inclusive prefix, *total = the CTA sum
__device__ __forceinline__ Off3 off3_cta_scan(Off3 v, Off3 *s_warp, Off3 *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Off3 u = off3_shfl_up(v, d);
        if (lane >= d) v = off3_add(v, u);
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    Off3 before{0, 0, 0}, all{0, 0, 0};
    for (int w = 0; w < OFF_THREADS / 32; ++w) {
        if (w < warp) before = off3_add(before, s_warp[w]);
        all = off3_add(all, s_warp[w]);
    }
    *total = all;
    return off3_add(v, before);
}
__global__ void __launch_bounds__(OFF_THREADS) offsets_sums_kernel(ReadsDev R, long long *__restrict__ blk_sums) {
    __shared__ Off3 s_warp[OFF_THREADS / 32];
    const int64_t i0 = ((int64_t)blockIdx.x * OFF_THREADS + threadIdx.x) * OFF_PER;
    Off3 v{0, 0, 0};
#pragma unroll
    for (int j = 0; j < OFF_PER; ++j) v = off3_add(v, off3_of(R, i0 + j));
    Off3 total;
    (void)off3_cta_scan(v, s_warp, &total);
    if (threadIdx.x == 0) {
        blk_sums[3 * (int64_t)blockIdx.x] = total.c;
        blk_sums[3 * (int64_t)blockIdx.x + 1] = total.s;
        blk_sums[3 * (int64_t)blockIdx.x + 2] = total.q;
    }
}
// exclusive scan of the block sums in place, ONE CTA: every thread owns a run of consecutive blocks
__global__ void __launch_bounds__(OFF_THREADS) offsets_scan_kernel(long long *__restrict__ blk_sums, int64_t n_blocks) {
    __shared__ Off3 s_warp[OFF_THREADS / 32];
    const int64_t per = (n_blocks + OFF_THREADS - 1) / OFF_THREADS;
    const int64_t k0 = (int64_t)threadIdx.x * per, k1 = min(k0 + per, n_blocks);
    Off3 v{0, 0, 0};
    for (int64_t k = k0; k < k1; ++k) v = off3_add(v, Off3{blk_sums[3 * k], blk_sums[3 * k + 1], blk_sums[3 * k + 2]});
    Off3 total;
    const Off3 incl = off3_cta_scan(v, s_warp, &total);
    Off3 run{incl.c - v.c, incl.s - v.s, incl.q - v.q};  // exclusive prefix of this thread's run
    for (int64_t k = k0; k < k1; ++k) {
        const Off3 x{blk_sums[3 * k], blk_sums[3 * k + 1], blk_sums[3 * k + 2]};
        blk_sums[3 * k] = run.c; blk_sums[3 * k + 1] = run.s; blk_sums[3 * k + 2] = run.q;
        run = off3_add(run, x);
    }
}
__global__ void __launch_bounds__(OFF_THREADS)
offsets_apply_kernel(ReadsDev R, const long long *__restrict__ blk_excl, int64_t *__restrict__ cigar_off, int64_t *__restrict__ seq_off,
                     int64_t *__restrict__ qual_off) {
    __shared__ Off3 s_warp[OFF_THREADS / 32];
    const int64_t i0 = ((int64_t)blockIdx.x * OFF_THREADS + threadIdx.x) * OFF_PER;
    Off3 x[OFF_PER], v{0, 0, 0};
#pragma unroll
    for (int j = 0; j < OFF_PER; ++j) {
        x[j] = off3_of(R, i0 + j);
        v = off3_add(v, x[j]);
    }
    Off3 total;
    const Off3 incl = off3_cta_scan(v, s_warp, &total);
    Off3 run{blk_excl[3 * (int64_t)blockIdx.x] + incl.c - v.c, blk_excl[3 * (int64_t)blockIdx.x + 1] + incl.s - v.s,
             blk_excl[3 * (int64_t)blockIdx.x + 2] + incl.q - v.q};
#pragma unroll
    for (int j = 0; j < OFF_PER; ++j) {
        if (i0 + j < R.n_reads) {
            cigar_off[i0 + j] = run.c;
            seq_off[i0 + j] = run.s;
            qual_off[i0 + j] = run.q;
        }
        run = off3_add(run, x[j]);
    }
}
cudaError_t launch_offsets(const ReadsDev &R, long long *blk_sums, int64_t *cigar_off, int64_t *seq_off, int64_t *qual_off, cudaStream_t stream,
                           int *launches) {
    if (R.n_reads <= 0) return cudaSuccess;
    const int64_t nb = (R.n_reads + OFF_THREADS * OFF_PER - 1) / (OFF_THREADS * OFF_PER);
    offsets_sums_kernel<<<(unsigned)nb, OFF_THREADS, 0, stream>>>(R, blk_sums);
    offsets_scan_kernel<<<1, OFF_THREADS, 0, stream>>>(blk_sums, nb);
    offsets_apply_kernel<<<(unsigned)nb, OFF_THREADS, 0, stream>>>(R, blk_sums, cigar_off, seq_off, qual_off);
    *launches += 3;
    return cudaGetLastError();
}

cudaError_t launch_prep(const ReadsDev &R, uint2 *ee_out, int64_t *chunk_max, int64_t *blk_max, int64_t *blk_incl, uint32_t *batch_status,
                        cudaStream_t stream, int *launches) {
    if (R.n_reads <= 0) return cudaSuccess;
    const int64_t nb = (R.n_reads + PREP_BLOCK - 1) / PREP_BLOCK;
    // the vector path reads four reads per lane with 128-bit loads: the streamed arrays must be 16-byte aligned (device allocations are)
    const bool aligned = (((uintptr_t)R.pos | (uintptr_t)R.n_cigar | (uintptr_t)R.l_qseq | (uintptr_t)R.cigar_off | (uintptr_t)ee_out) & 15u) == 0u;
    if (aligned) prep_kernel<true><<<(unsigned)nb, PREP_THREADS, 0, stream>>>(R, ee_out, chunk_max, blk_max, batch_status);
    else prep_kernel<false><<<(unsigned)nb, PREP_THREADS, 0, stream>>>(R, ee_out, chunk_max, blk_max, batch_status);
    scan_kernel<<<(unsigned)((nb + SCAN_THREADS - 1) / SCAN_THREADS), SCAN_THREADS, 0, stream>>>(blk_max, nb, blk_incl);
    *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_link(const ReadsDev &R, const vfk_params &P, const void *resolved, const uint32_t *list, const uint32_t *n_list,
                        int64_t n_all, uint32_t *heads, uint32_t n_buckets, uint32_t *next, int32_t *column, uint32_t *claimed,
                        uint32_t *mate, int sm_count, cudaStream_t stream, int *launches) {
    if (n_all <= 0 || R.n_reads <= 0) return cudaSuccess;
    LinkScratch L;
    L.heads = heads; L.n_buckets = n_buckets; L.next = next; L.column = column; L.claimed = claimed; L.mate = mate;
    const int64_t warps = list ? (int64_t)sm_count * 64 : n_all;
    const unsigned grid = (unsigned)std::min<int64_t>((warps + 7) / 8, (int64_t)sm_count * 8);
    const ResolvedUnit *res = (const ResolvedUnit *)resolved;
    link_mark_kernel<<<grid, 256, 0, stream>>>(R, P, res, list, n_list, n_all, L);
    const unsigned rgrid = (unsigned)sm_count * 8;
    link_reads_kernel<0><<<rgrid, 256, 0, stream>>>(R, P, L);
    link_reads_kernel<1><<<rgrid, 256, 0, stream>>>(R, P, L);
    link_reads_kernel<2><<<rgrid, 256, 0, stream>>>(R, P, L);
    ++*launches;
    *launches += 3;
    return cudaGetLastError();
}

}  // namespace vfk

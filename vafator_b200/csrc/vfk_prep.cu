// vfk_prep.cu -- device-side read preparation: everything the pileup kernels need beyond the canonical arrays
// is derived here, on the device, inside every call (the reference starts at bam.pileup(), vafator/pileups.py:71-80,
// with nothing but the decoded records either).
//
//  prep_kernel   one thread per read: walks the CIGAR once -> reference end (htslib bam_endpos), the shape word
//                (vfk_device.cuh) and, per block of PREP_BLOCK reads, the maximum of the end keys; checks the batch
//                (sorted, starts >= 0, CIGAR inside its pool, lengths in range) and raises VFK_ST_BADBATCH otherwise.
//  scan_kernel   inclusive running maximum over the block maxima: the position index.  A column's first
//                overlapping read is the first read whose end exceeds it -- found by a binary search over the
//                blocks and one pass over a block (locate_kernel, vfk_pileup.cu).
//  link kernels  htslib's mate-overlap pairing (sam.c overlap_push / overlap_remove) for the reads
//                inside the windows of the units on a device-side list -- the histogram paths (deep columns, and
//                the few shallow ones the register kernel hands on) take a read's partner from `mate[]`; the
//                register kernel pairs the reads of its column in registers instead.
//                Emulation of the sequential name hash: the records that reach overlap_push are chained per name
//                bucket (atomicExch on a bucket head), then every record replays the protocol over the records
//                that share its name, in file order: found -> pair and delete; not found -> stored if the mate is
//                still to come; an entry whose record has left the pileup buffer is dropped.
#include "vfk_device.cuh"

namespace vfk {

// ---- CIGAR -> span + shape word (same classification for every kernel) ----
struct CigarInfo {
    uint32_t span;
    uint32_t ev;
};
__device__ __forceinline__ bool op_is_match(uint32_t op) { return op == OP_M || op == OP_EQ || op == OP_X; }
__device__ __forceinline__ bool op_consumes_ref(uint32_t op) { return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X; }

// [H*][S] M [ (I|D|N) M ] [S][H*] with every length in range -> the packed event word
__device__ __forceinline__ uint32_t one_event_word(const uint32_t *__restrict__ c, int n, int32_t l_qseq) {
    int k = 0, e = n;
    while (k < e && (__ldg(c + k) & 15u) == OP_H) ++k;
    while (e > k && (__ldg(c + e - 1) & 15u) == OP_H) --e;
    uint32_t lead = 0;
    if (k < e && (__ldg(c + k) & 15u) == OP_S) lead = __ldg(c + k++) >> 4;
    if (e > k && (__ldg(c + e - 1) & 15u) == OP_S) --e;  // a trailing clip only lengthens the query
    const int ops = e - k;
    if (ops != 1 && ops != 3) return EV_GENERAL;
    const uint32_t w0 = __ldg(c + k);
    if (!op_is_match(w0 & 15u)) return EV_GENERAL;
    const uint32_t m1 = w0 >> 4;
    uint32_t ev = EVK_NONE, len = 0;
    if (ops == 3) {
        const uint32_t w1 = __ldg(c + k + 1), w2 = __ldg(c + k + 2);
        const uint32_t op = w1 & 15u;
        ev = op == OP_I ? EVK_INS : op == OP_D ? EVK_DEL : op == OP_N ? EVK_SKIP : EVK_NONE;
        len = w1 >> 4;
        if (ev == EVK_NONE || !op_is_match(w2 & 15u) || (w2 >> 4) == 0u || len == 0u) return EV_GENERAL;
    }
    if (m1 == 0u || m1 >= 4096u || lead >= 4096u || len >= 64u || l_qseq > VFK_MAX_READ_LEN) return EV_GENERAL;
    return lead | (m1 << 12) | (ev << 24) | (len << 26);
}

__device__ __forceinline__ CigarInfo cigar_info(const uint32_t *__restrict__ c, int n, int32_t l_qseq) {
    CigarInfo o;
    if (n == 1) {  // nine reads in ten
        const uint32_t w = __ldg(c);
        const uint32_t op = w & 15u, len = w >> 4;
        o.span = op_consumes_ref(op) ? len : 0u;
        o.ev = (op_is_match(op) && (int64_t)len == (int64_t)l_qseq) ? EV_SIMPLE : one_event_word(c, 1, l_qseq);
        return o;
    }
    uint32_t span = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t w = __ldg(c + k);
        if (op_consumes_ref(w & 15u)) span += w >> 4;
    }
    o.span = span;
    o.ev = n == 0 ? EV_GENERAL : one_event_word(c, n, l_qseq);
    return o;
}

__device__ __forceinline__ int32_t contig_of(const int64_t *__restrict__ contig_off, int32_t n_contigs, int64_t i) {
    int32_t a = 0, b = n_contigs;  // last t with contig_off[t] <= i
    while (b - a > 1) {
        const int32_t m = (a + b) >> 1;
        if (__ldg(contig_off + m) <= i) a = m; else b = m;
    }
    return a;
}

// status word of a call (device): VFK_ST_BADBATCH when the batch is malformed
__global__ void __launch_bounds__(PREP_BLOCK)
prep_kernel(ReadsDev R, const int32_t *__restrict__ pos_src, int32_t *__restrict__ pos_copy, int32_t *__restrict__ end_out,
            uint32_t *__restrict__ ev_out, int64_t *__restrict__ blk_max, uint32_t *__restrict__ batch_status) {
    __shared__ long long s_max[PREP_BLOCK / 32];
    __shared__ int32_t s_tid[2];
    const int64_t i = (int64_t)blockIdx.x * PREP_BLOCK + threadIdx.x;
    const int64_t b0 = (int64_t)blockIdx.x * PREP_BLOCK, b1 = min(b0 + PREP_BLOCK, R.n_reads) - 1;
    if (threadIdx.x < 2) s_tid[threadIdx.x] = contig_of(R.contig_off, R.n_contigs, threadIdx.x ? b1 : b0);
    __syncthreads();
    long long key = LLONG_MIN;
    bool bad = false;
    if (i < R.n_reads) {
        const int32_t p = __ldg(pos_src + i);
        const int32_t n = __ldg(R.n_cigar + i);
        const int64_t off = __ldg(R.cigar_off + i);
        const int32_t lq = __ldg(R.l_qseq + i);
        const int32_t tid = s_tid[0] == s_tid[1] ? s_tid[0] : contig_of(R.contig_off, R.n_contigs, i);
        bad = p < 0 || n < 0 || off < 0 || off + n > R.n_cigar_words || lq < 0 || lq > VFK_MAX_READ_LEN;
        // coordinate sorted inside the contig
        if (i > 0 && i > __ldg(R.contig_off + tid) && __ldg(pos_src + i - 1) > p) bad = true;
        CigarInfo ci;
        ci.span = 0u; ci.ev = EV_GENERAL;
        if (!bad) ci = cigar_info(R.cigar + off, n, lq);
        if ((uint64_t)p + ci.span > 0x7FFFFFFFull) { bad = true; ci.span = 0u; }
        if (pos_copy) pos_copy[i] = p;
        end_out[i] = p + (int32_t)ci.span;
        ev_out[i] = ci.ev;
        key = ((long long)tid << 32) | (long long)(uint32_t)(p + (int32_t)ci.span);
    }
    if (__any_sync(FULL, bad) && (threadIdx.x & 31) == 0) atomicOr(batch_status, VFK_ST_BADBATCH);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = s_max[0];
#pragma unroll
        for (int w = 1; w < PREP_BLOCK / 32; ++w) m = max(m, s_max[w]);
        blk_max[blockIdx.x] = m;
    }
}

// blk[0] = LLONG_MIN, blk[b + 1] = max(blk_max[0..b]) -- one CTA (the array has n_reads / 128 entries)
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const int64_t *__restrict__ blk_max, int64_t n_blocks, int64_t *__restrict__ blk_incl) {
    __shared__ long long s_part[SCAN_THREADS];
    const int t = threadIdx.x;
    const int64_t per = (n_blocks + SCAN_THREADS - 1) / SCAN_THREADS;
    const int64_t a = min((int64_t)t * per, n_blocks), b = min(a + per, n_blocks);
    long long m = LLONG_MIN;
    for (int64_t k = a; k < b; ++k) m = max(m, (long long)blk_max[k]);
    s_part[t] = m;
    __syncthreads();
    // inclusive scan of the per-thread maxima (Hillis-Steele, 10 steps)
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
        const long long v = t >= o ? s_part[t - o] : LLONG_MIN;
        __syncthreads();
        s_part[t] = max(s_part[t], v);
        __syncthreads();
    }
    long long run = t > 0 ? s_part[t - 1] : LLONG_MIN;
    if (t == 0) blk_incl[0] = LLONG_MIN;
    for (int64_t k = a; k < b; ++k) {
        run = max(run, (long long)blk_max[k]);
        blk_incl[k + 1] = run;
    }
}

// ------------------------------------------------------------------------------------------
// mate linking for the windows of listed units
// ------------------------------------------------------------------------------------------
constexpr uint32_t LINK_EMPTY = 0xFFFFFFFFu;
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct LinkScratch {
    uint32_t *heads;     // [n_buckets] bucket -> last inserted record, LINK_EMPTY when empty (restored after every call)
    uint32_t n_buckets;  // power of two
    uint32_t *next;      // [n_reads] chain
    int32_t *column;     // [n_reads] push_column of the chained records
    uint32_t *claimed;   // [n_reads / 32 + 1] bitmap: record handled by this call
    uint32_t *mate;      // [n_reads] result
};
__device__ __forceinline__ uint32_t bucket_of(const ReadsDev &R, const LinkScratch &L, int64_t i, int32_t tid) {
    return (uint32_t)mix64(__ldg(R.qname_id + i) ^ ((uint64_t)(uint32_t)tid << 40) ^ __ldg(R.qname_hash + i)) & (L.n_buckets - 1u);
}

// the window of a located unit as the linker sees it: its candidate reads plus the first record pushed after the column
// (bam_plp_auto reads one record ahead: that record's arrival can still pair with a read of the column)
__device__ __forceinline__ int64_t window_end(const ReadsDev &R, const vfk_params &P, const ResolvedUnit &r) {
    const int64_t nx = next_pushed(R, P, (int64_t)r.hi, (int64_t)r.c1, r.stop);
    return nx >= 0 ? nx + 1 : (int64_t)r.hi;
}

// phase: 0 = claim + chain, 1 = replay the protocol and write mate[], 2 = restore the bucket heads and the bitmap
template <int PHASE>
__global__ void __launch_bounds__(256)
link_kernel(ReadsDev R, vfk_params P, const ResolvedUnit *__restrict__ res, const uint32_t *__restrict__ list, const uint32_t *__restrict__ n_list,
            int64_t n_all, LinkScratch L) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n = list ? (int64_t)*n_list : n_all;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = w; k < n; k += n_warps) {
        const uint32_t u = list ? list[k] : (uint32_t)k;
        if (u == 0xFFFFFFFFu) continue;
        const ResolvedUnit r = res[u];
        const int64_t c1 = (int64_t)r.c1;
        if (r.hi <= r.lo) continue;
        int64_t last = 0;
        if (lane == 0) last = window_end(R, P, r);
        last = __shfl_sync(FULL, last, 0);
        const int32_t tid = contig_of(R.contig_off, R.n_contigs, (int64_t)r.lo);
        const int64_t c0 = __ldg(R.contig_off + tid);
        for (int64_t i = (int64_t)r.lo + lane; i < last; i += 32) {
            if (i >= (int64_t)r.hi && i != last - 1) continue;  // between the column and the next pushed record: filtered reads
            const uint32_t bit = 1u << (i & 31);
            if (PHASE == 0) {
                if (atomicOr(L.claimed + (i >> 5), bit) & bit) continue;  // another unit's warp has it
                L.mate[i] = VFK_NO_MATE;
                L.next[i] = LINK_EMPTY - 1u;  // not chained
                const int32_t column = push_column(R, P, i, c0);
                if (!P.ignore_overlaps || !reaches_hash(R, P, i, tid, column)) continue;
                L.column[i] = column;
                L.next[i] = atomicExch(L.heads + bucket_of(R, L, i, tid), (uint32_t)i);
            } else if (PHASE == 1) {
                if (L.next[i] == LINK_EMPTY - 1u) continue;
                if (L.mate[i] != VFK_NO_MATE) continue;  // written by its partner's thread (same value)
                // replay over the records sharing this name, in file order; stop once record i is settled
                const uint64_t id = __ldg(R.qname_id + i);
                const uint32_t qh = __ldg(R.qname_hash + i);
                const uint32_t head = L.heads[bucket_of(R, L, i, tid)];
                int64_t waiting = -1, cur = -1;
                uint32_t result = VFK_NO_MATE;
                for (;;) {
                    int64_t m = -1;  // smallest member beyond cur
                    for (uint32_t j = head; j != LINK_EMPTY; j = L.next[j]) {
                        if ((int64_t)j > cur && (m < 0 || (int64_t)j < m) && (int64_t)j >= c0 && (int64_t)j < c1 && __ldg(R.qname_id + j) == id &&
                            __ldg(R.qname_hash + j) == qh)
                            m = (int64_t)j;
                    }
                    if (m < 0) break;
                    if (waiting >= 0 && __ldg(R.end + waiting) < L.column[m]) waiting = -1;  // its record already left the buffer
                    if (waiting < 0) {
                        if (stores_itself(R, m)) waiting = m;
                    } else {
                        if (waiting == i || m == i) {
                            const uint32_t sel = wang_hash(__ldg(R.qname_hash + waiting)) & 1u;
                            result = (uint32_t)(waiting == i ? m : waiting) | (sel << 31);
                            break;
                        }
                        waiting = -1;
                    }
                    cur = m;
                    if (m >= i && waiting != i) break;  // record i has been through without a partner
                }
                L.mate[i] = result;
            } else {
                if (L.next[i] != LINK_EMPTY - 1u) L.heads[bucket_of(R, L, i, tid)] = LINK_EMPTY;
                L.claimed[i >> 5] = 0u;  // benign race: every writer stores zero
            }
        }
    }
}

cudaError_t launch_prep(const ReadsDev &R, const int32_t *pos_src, int32_t *pos_copy, int32_t *end_out, uint32_t *ev_out, int64_t *blk_max,
                        int64_t *blk_incl, uint32_t *batch_status, cudaStream_t stream, int *launches) {
    if (R.n_reads <= 0) return cudaSuccess;
    const int64_t nb = (R.n_reads + PREP_BLOCK - 1) / PREP_BLOCK;
    prep_kernel<<<(unsigned)nb, PREP_BLOCK, 0, stream>>>(R, pos_src, pos_copy, end_out, ev_out, blk_max, batch_status);
    scan_kernel<<<1, SCAN_THREADS, 0, stream>>>(blk_max, nb, blk_incl);
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
    link_kernel<0><<<grid, 256, 0, stream>>>(R, P, res, list, n_list, n_all, L);
    link_kernel<1><<<grid, 256, 0, stream>>>(R, P, res, list, n_list, n_all, L);
    link_kernel<2><<<grid, 256, 0, stream>>>(R, P, res, list, n_list, n_all, L);
    *launches += 3;
    return cudaGetLastError();
}

}  // namespace vfk

// vfk_prep.cu -- device-side read preparation: everything the pileup kernels need beyond the canonical arrays
// is derived here, on the device, inside every call (the reference starts at bam.pileup(), vafator/pileups.py:71-80,
// with nothing but the decoded records either).
//
//  prep_kernel   one thread per read: walks the CIGAR once -> reference end (htslib bam_endpos), the shape word
//                (vfk_device.cuh) and the maximum of the end keys per chunk of 32 reads and per block of PREP_BLOCK reads;
//                checks the batch (sorted, starts >= 0, CIGAR inside its pool, lengths in range) and raises
//                VFK_ST_BADBATCH otherwise.
//  scan_kernel   inclusive running maximum over the block maxima: the top level of the position index.  A column's
//                first overlapping read is the first read whose end exceeds it -- found by a binary search over the
//                blocks, a walk over the chunk maxima of one block and one pass over a chunk (locate_kernel, vfk_pileup.cu).
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
__device__ __forceinline__ CigarInfo cigar_walk(const uint32_t *__restrict__ c, int n, int32_t l_qseq) {
    CigarInfo o;
    uint32_t span = 0, lead = 0, m1 = 0, ev = EVK_NONE, elen = 0;
    int state = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t w = __ldg(c + k);
        const uint32_t op = w & 15u, len = w >> 4;
        if (op_consumes_ref(op)) span += len;
        int next = 7;  // 7 = not of this shape
        if (op == OP_H) next = (state == 0) ? 0 : (state == 2 || state >= 4) ? 6 : 7;
        else if (op == OP_S) {
            if (state == 0) { next = 1; lead = len; }
            else if (state == 2 || state == 4) next = 5;
        } else if (op_is_match(op)) {
            if (state <= 1) { next = (len == 0u) ? 7 : 2; m1 = len; }
            else if (state == 3) next = (len == 0u) ? 7 : 4;
        } else if (op == OP_I || op == OP_D || op == OP_N) {
            if (state == 2 && len != 0u) { next = 3; ev = op == OP_I ? EVK_INS : op == OP_D ? EVK_DEL : EVK_SKIP; elen = len; }
        }
        state = (state == 7) ? 7 : next;
    }
    o.span = span;
    const bool shaped = (state == 2 || state >= 4) && state != 7 && m1 != 0u && m1 < 4096u && lead < 4096u && elen < 64u && l_qseq <= VFK_MAX_READ_LEN;
    o.ev = shaped ? (lead | (m1 << 12) | (ev << 24) | (elen << 26)) : EV_GENERAL;
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

// One CTA = PREP_BLOCK (1024) reads, PREP_THREADS (256) threads: a warp takes four CHUNKS of 32 consecutive reads, every
// thread one read of each -- all of a thread's loads (four reads x starts, CIGAR counts / offsets, read lengths) are issued
// before the first is consumed, and the contig of the block is searched by one thread meanwhile.  Writes end / shape word per
// read, the maximum end key (tid << 32 | end) per chunk and per block: the two levels of the position index.
__global__ void __launch_bounds__(PREP_THREADS)
prep_kernel(ReadsDev R, const int32_t *__restrict__ pos_src, int32_t *__restrict__ pos_copy, int32_t *__restrict__ end_out,
            uint32_t *__restrict__ ev_out, int64_t *__restrict__ chunk_max, int64_t *__restrict__ blk_max,
            uint32_t *__restrict__ batch_status) {
    constexpr int PER = PREP_BLOCK / PREP_THREADS;  // reads per thread
    __shared__ long long s_max[PREP_THREADS / 32];
    __shared__ int32_t s_tid[2];
    const int64_t b0 = (int64_t)blockIdx.x * PREP_BLOCK, b1 = min(b0 + PREP_BLOCK, R.n_reads) - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t p[PER], n[PER], lq[PER], prev[PER];
    int64_t off[PER], idx[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        // chunk (warp * PER + j) of the block: consecutive reads per warp instruction, consecutive chunks per warp
        idx[j] = b0 + (int64_t)(warp * PER + j) * 32 + lane;
        const bool in = idx[j] < R.n_reads;
        p[j] = in ? __ldg(pos_src + idx[j]) : 0;
        n[j] = in ? __ldg(R.n_cigar + idx[j]) : 0;
        off[j] = in ? __ldg(R.cigar_off + idx[j]) : 0;
        lq[j] = in ? __ldg(R.l_qseq + idx[j]) : 0;
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {  // the start before this one (sortedness): the lane below has it, lane 0 reads it
        prev[j] = __shfl_up_sync(FULL, p[j], 1);
        if (lane == 0) prev[j] = (idx[j] > 0 && idx[j] < R.n_reads) ? __ldg(pos_src + idx[j] - 1) : 0;
    }
    if (threadIdx.x < 2) s_tid[threadIdx.x] = contig_of(R.contig_off, R.n_contigs, threadIdx.x ? b1 : b0);
    __syncthreads();
    const bool one_contig = s_tid[0] == s_tid[1];
    const int64_t first_of_contig = one_contig ? __ldg(R.contig_off + s_tid[0]) : 0;
    bool bad = false;
    uint32_t span[PER], ev[PER];
    uint32_t pending = 0u;  // reads of this thread whose CIGAR has more than one op
    // pass 1: checks + the one-op CIGARs, every lane busy
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        span[j] = 0u;
        ev[j] = EV_GENERAL;
        if (idx[j] < R.n_reads) {
            const bool b = p[j] < 0 || n[j] < 0 || off[j] < 0 || off[j] + n[j] > R.n_cigar_words || lq[j] < 0 || lq[j] > VFK_MAX_READ_LEN;
            bad = bad || b;
            if (!b && n[j] == 1) {
                const CigarInfo ci = cigar_single(__ldg(R.cigar + off[j]), lq[j]);
                span[j] = ci.span;
                ev[j] = ci.ev;
            } else if (!b && n[j] > 1) {
                pending |= 1u << j;
            }
        }
    }
    // pass 2: the other CIGARs, one per thread and round -- the few complex reads of a warp's four chunks share the rounds
    while (__any_sync(FULL, pending != 0u)) {
        if (pending) {
            const int j = __ffs(pending) - 1;
            pending &= pending - 1u;
            const int64_t o = j == 0 ? off[0] : j == 1 ? off[1] : j == 2 ? off[2] : off[3];
            const int nn = j == 0 ? n[0] : j == 1 ? n[1] : j == 2 ? n[2] : n[3];
            const int32_t l = j == 0 ? lq[0] : j == 1 ? lq[1] : j == 2 ? lq[2] : lq[3];
            const CigarInfo ci = cigar_walk(R.cigar + o, nn, l);
#pragma unroll
            for (int t = 0; t < PER; ++t)
                if (t == j) { span[t] = ci.span; ev[t] = ci.ev; }
        }
    }
    long long blk_key = LLONG_MIN;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int64_t i = idx[j];
        const bool in = i < R.n_reads;
        int32_t tid = s_tid[0];
        if (in) {
            bool b = false;
            if (!one_contig) tid = contig_of(R.contig_off, R.n_contigs, i);
            // coordinate sorted inside the contig
            if (i > (one_contig ? first_of_contig : __ldg(R.contig_off + tid)) && prev[j] > p[j]) b = true;
            if ((uint64_t)(uint32_t)p[j] + span[j] > 0x7FFFFFFFull) { b = true; span[j] = 0u; }
            if (pos_copy) pos_copy[i] = p[j];
            end_out[i] = p[j] + (int32_t)span[j];
            ev_out[i] = ev[j];
            bad = bad || b;
        }
        long long key;
        if (one_contig) {  // one REDUX instead of a 64-bit butterfly
            const int32_t e = __reduce_max_sync(FULL, in ? p[j] + (int32_t)span[j] : -1);
            key = e < 0 ? LLONG_MIN : (((long long)tid << 32) | (long long)(uint32_t)e);
        } else {
            key = in ? (((long long)tid << 32) | (long long)(uint32_t)(p[j] + (int32_t)span[j])) : LLONG_MIN;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
        }
        const int64_t chunk = (b0 >> 5) + warp * PER + j;
        if (lane == 0 && chunk * 32 < R.n_reads) chunk_max[chunk] = key;
        blk_key = max(blk_key, key);
    }
    if (__any_sync(FULL, bad) && lane == 0) atomicOr(batch_status, VFK_ST_BADBATCH);
    if (lane == 0) s_max[warp] = blk_key;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = s_max[0];
#pragma unroll
        for (int w = 1; w < PREP_THREADS / 32; ++w) m = max(m, s_max[w]);
        blk_max[blockIdx.x] = m;
    }
}

// blk_incl[0] = LLONG_MIN, blk_incl[b + 1] = max(blk_max[0..b]) -- one CTA walks the array (n_reads / 1024 entries) in tiles
// of 1024: coalesced load, warp scans + a scan of the warp totals, the running maximum carried from tile to tile
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) scan_kernel(const int64_t *__restrict__ blk_max, int64_t n_blocks, int64_t *__restrict__ blk_incl) {
    __shared__ long long s_warp[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    long long carry = LLONG_MIN;
    if (t == 0) blk_incl[0] = LLONG_MIN;
    for (int64_t base = 0; base < n_blocks; base += SCAN_THREADS) {
        const int64_t k = base + t;
        long long v = k < n_blocks ? (long long)blk_max[k] : LLONG_MIN;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long u = __shfl_up_sync(FULL, v, o);
            if (lane >= o) v = max(v, u);
        }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long u = __shfl_up_sync(FULL, w, o);
                if (lane >= o) w = max(w, u);
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const long long before = warp > 0 ? s_warp[warp - 1] : LLONG_MIN;
        v = max(max(v, before), carry);
        if (k < n_blocks) blk_incl[k + 1] = v;
        carry = max(carry, s_warp[31]);
        __syncthreads();
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

cudaError_t launch_prep(const ReadsDev &R, const int32_t *pos_src, int32_t *pos_copy, int32_t *end_out, uint32_t *ev_out, int64_t *chunk_max,
                        int64_t *blk_max, int64_t *blk_incl, uint32_t *batch_status, cudaStream_t stream, int *launches) {
    if (R.n_reads <= 0) return cudaSuccess;
    const int64_t nb = (R.n_reads + PREP_BLOCK - 1) / PREP_BLOCK;
    prep_kernel<<<(unsigned)nb, PREP_THREADS, 0, stream>>>(R, pos_src, pos_copy, end_out, ev_out, chunk_max, blk_max, batch_status);
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

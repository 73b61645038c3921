// vfk_link.cuh -- htslib's mate-overlap pairing (sam.c overlap_push / overlap_remove) for the reads inside the
// windows of the units on a device-side list: the histogram paths (deep columns, and the few shallow ones the register kernel
// hands on) take a read's partner from `mate[]`; the register kernel pairs the reads of its column in registers instead.
// Emulation of the sequential name hash: the records that reach overlap_push are chained per name bucket (atomicExch on a
// bucket head), then every record replays the protocol over the records that share its name, in file order: found -> pair
// and delete; not found -> stored if the mate is still to come; an entry whose record has left the pileup buffer is dropped.
#pragma once
#include "vfk_device.cuh"

namespace vfk {

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
    const int64_t nx = next_pushed(R, P, (int64_t)r.hi, contig_end(R, r.tid), r.stop);
    return nx >= 0 ? nx + 1 : (int64_t)r.hi;
}

// The linker in two kinds of passes.
//   link_mark   a warp per listed unit (list == nullptr: units 0 .. n-1): the reads of its window -- its candidate reads plus
//               the first record pushed after the column -- are marked in the `claimed` bitmap, a word (32 reads) per
//               atomic.  Windows of neighbouring units overlap many times over in a deep panel; a bit is set once.
//   link_reads  a thread per marked read, warps striding over the bitmap words.  PHASE 0 = chain the records that reach
//               overlap_push per name bucket, 1 = replay the protocol and write mate[], 2 = restore the bucket heads and
//               clear the bitmap.
// The scratch arrays are written and read inside one launch when the passes run fused (tail_kernel, grid barriers in
// between): every access to them is a plain (coherent) load or store, never the read-only path.
// `out` (may be nullptr): a listed unit whose record no longer carries VFK_ST_DEFERRED has been finished by the tile kernel
__device__ __forceinline__ void link_mark(const ReadsDev &R, const vfk_params &P, const ResolvedUnit *__restrict__ res, const uint32_t *list, int64_t n,
                                          const LinkScratch &L, const vfk_counts *out = nullptr) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = w; k < n; k += n_warps) {
        const uint32_t u = list ? list[k] : (uint32_t)k;
        if (u == 0xFFFFFFFFu) continue;
        if (out && !(out[u].status & VFK_ST_DEFERRED)) continue;
        const ResolvedUnit r = res[u];
        if (r.hi <= r.lo) continue;
        const int64_t lo = (int64_t)r.lo, hi = (int64_t)r.hi;
        for (int64_t word = (lo >> 5) + lane; word <= ((hi - 1) >> 5); word += 32) {
            uint32_t m = 0xFFFFFFFFu;
            if (word == (lo >> 5)) m &= 0xFFFFFFFFu << (lo & 31);
            if (word == ((hi - 1) >> 5)) m &= 0xFFFFFFFFu >> (31 - ((hi - 1) & 31));
            if ((L.claimed[word] & m) != m) atomicOr(L.claimed + word, m);
        }
        if (lane == 0) {
            const int64_t last = window_end(R, P, r);  // the first record pushed after the column, or hi
            if (last > hi) atomicOr(L.claimed + ((last - 1) >> 5), 1u << ((last - 1) & 31));
        }
    }
}

template <int PHASE>
__device__ __forceinline__ void link_one_read(const ReadsDev &R, const vfk_params &P, const LinkScratch &L, int64_t i, int32_t tid) {
    const int64_t c0 = __ldg(R.contig_off + tid), c1 = __ldg(R.contig_off + tid + 1);
    if (PHASE == 0) {
        L.mate[i] = VFK_NO_MATE;
        L.next[i] = LINK_EMPTY - 1u;  // not chained
        const int32_t column = push_column(R, P, i, c0);
        if (!P.ignore_overlaps || !reaches_hash(R, P, i, tid, column)) return;
        L.column[i] = column;
        L.next[i] = atomicExch(L.heads + bucket_of(R, L, i, tid), (uint32_t)i);
    } else if (PHASE == 1) {
        if (L.next[i] == LINK_EMPTY - 1u) return;
        if (L.mate[i] != VFK_NO_MATE) return;  // written by its partner's thread (same value)
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
            if (waiting >= 0 && ld_end(R, waiting) < L.column[m]) waiting = -1;  // its record already left the buffer
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
    }
}

template <int PHASE>
__device__ __forceinline__ void link_reads(const ReadsDev &R, const vfk_params &P, const LinkScratch &L) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_words = (R.n_reads >> 5) + 1;
    // Warp w owns the words w, w + n_warps, w + 2 n_warps, ...: marked reads come in runs (the windows of neighbouring units),
    // and a run of consecutive words spreads over as many warps.  A warp fetches 32 of its words at a time (one per lane)
    // and then visits those that have a bit set.
    for (int64_t k0 = 0; w + k0 * n_warps < n_words; k0 += 32) {
        const int64_t my_word = w + (k0 + lane) * n_warps;
        const uint32_t mine = my_word < n_words ? L.claimed[my_word] : 0u;
        uint32_t nz = __ballot_sync(FULL, mine != 0u);
        if (PHASE == 2 && mine != 0u) L.claimed[my_word] = 0u;  // every pass has seen the word by now (grid barrier before this pass)
        while (nz) {
            const int j = __ffs(nz) - 1;
            nz &= nz - 1u;
            const uint32_t bits = __shfl_sync(FULL, mine, j);
            const int64_t i = ((w + (k0 + j) * n_warps) << 5) + lane;
            if (!((bits >> lane) & 1u) || i >= R.n_reads) continue;
            int32_t a = 0, b = R.n_contigs;  // contig of read i: last t with contig_off[t] <= i
            while (b - a > 1) {
                const int32_t m = (a + b) >> 1;
                if (__ldg(R.contig_off + m) <= i) a = m; else b = m;
            }
            link_one_read<PHASE>(R, P, L, i, a);
        }
    }
}

}  // namespace vfk

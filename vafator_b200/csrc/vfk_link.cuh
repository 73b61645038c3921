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

// One pass of the linker over the units of a list (list == nullptr: units 0 .. n-1), warps striding over the units.
// phase: 0 = claim + chain, 1 = replay the protocol and write mate[], 2 = restore the bucket heads and the bitmap.
// The scratch arrays are written and read inside one launch when the passes run fused (tail_kernel, grid barriers in
// between): every access to them is a plain (coherent) load or store, never the read-only path.
template <int PHASE>
__device__ __forceinline__ void link_pass(const ReadsDev &R, const vfk_params &P, const ResolvedUnit *__restrict__ res, const uint32_t *list, int64_t n,
                                          const LinkScratch &L) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = w; k < n; k += n_warps) {
        const uint32_t u = list ? list[k] : (uint32_t)k;
        if (u == 0xFFFFFFFFu) continue;
        const ResolvedUnit r = res[u];
        if (r.hi <= r.lo) continue;
        const int64_t c1 = contig_end(R, r.tid);
        int64_t last = 0;
        if (lane == 0) last = window_end(R, P, r);
        last = __shfl_sync(FULL, last, 0);
        const int32_t tid = r.tid;
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
                L.claimed[i >> 5] = 0u;  // benign race: every writer stores zero
            }
        }
    }
}


}  // namespace vfk

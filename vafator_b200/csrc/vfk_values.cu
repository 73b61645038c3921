// vfk_values.cu -- auxiliary kernels for the reference's list-shaped interfaces.
//
//  values_kernel   per unit, the per-read values the reference would hold in all_mqs / all_bqs /
//                  all_positions (vafator/pileups.py:199-220), packed one word per value.  Needed when
//                  replicate BAMs are merged: the reference then ranks the ALT values of one BAM against
//                  the REF values of another (dict.update, vafator/annotator.py:285-287).
//  ranksum_kernel  twice the ALT mid-rank sum of two explicit integer lists, per metric byte lane --
//                  the integer core of scipy.stats.ranksums as vafator/rank_sum_test.py:6-12 calls it.
#include "vfk_device.cuh"

namespace vfk {

// packed value: mq | bq << 8 | qpos << 16 | ref << 28 | alt << 29  (same as the pileup kernel)
__global__ void __launch_bounds__(256)
values_kernel(ReadsDev R, const ResolvedUnit *__restrict__ res, int64_t n_units, const uint8_t *__restrict__ ins_seq, vfk_params P,
              int32_t max_per_unit, uint32_t *__restrict__ values, int32_t *__restrict__ n_values) {
    const int lane = threadIdx.x & 31;
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (u >= n_units) return;
    // the units have been located and the reads of their windows linked (locate_kernel, link kernels) by this call
    Unit U;
    unit_from(res[u], ins_seq, U);
    uint32_t *out = values + (size_t)u * max_per_unit;
    int n = 0;
    for (int64_t base = U.lo; base < U.hi; base += 32) {
        const int64_t i = base + lane;
        Contribution c;
        c.ref = false; c.alt = 0; c.mq = c.bq = c.qpos = 0;
        if (i < U.hi) c = evaluate(R, P, U, i);
        const uint32_t copies = c.alt > 0u ? c.alt : (c.ref ? 1u : 0u);  // max(ref, alt multiplicity)
        // exclusive prefix of `copies` over the warp
        uint32_t incl = copies;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        const uint32_t w = (uint32_t)c.mq | ((uint32_t)c.bq << 8) | ((uint32_t)(c.qpos & 0xFFF) << 16) |
                           (c.ref ? (1u << 28) : 0u) | (c.alt ? (1u << 29) : 0u);
        for (uint32_t k = 0; k < copies; ++k) {
            const int at = n + (int)(incl - copies + k);
            // a read in both lists (REF == ALT text) is one word with both flags; extra ALT copies
            // (the reference can count one read several times for an insertion) carry the ALT flag only
            if (at < max_per_unit) out[at] = k == 0 ? w : (w & ~(1u << 28));
        }
        n += (int)total;
    }
    if (lane == 0) n_values[u] = n;
}

// one warp per pair of lists; values are the packed words above, the metric is a bit field of them
__global__ void __launch_bounds__(256)
ranksum_kernel(int64_t n_pairs, const uint32_t *__restrict__ alt, const int64_t *__restrict__ alt_off,
               const uint32_t *__restrict__ ref, const int64_t *__restrict__ ref_off, unsigned long long *__restrict__ s2) {
    const int lane = threadIdx.x & 31;
    const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= n_pairs) return;
    const uint32_t *A = alt + alt_off[p], *B = ref + ref_off[p];
    const int64_t na = alt_off[p + 1] - alt_off[p], nb = ref_off[p + 1] - ref_off[p];
    const int shift[3] = {0, 8, 16};
    const uint32_t mask[3] = {0xFFu, 0xFFu, 0xFFFu};
    for (int m = 0; m < 3; ++m) {
        unsigned long long acc = 0;
        for (int64_t i = lane; i < na; i += 32) {
            const uint32_t v = (A[i] >> shift[m]) & mask[m];
            unsigned long long below = 0, ties = 0;
            for (int64_t j = 0; j < na; ++j) { const uint32_t w = (A[j] >> shift[m]) & mask[m]; below += w < v; ties += w == v; }
            for (int64_t j = 0; j < nb; ++j) { const uint32_t w = (B[j] >> shift[m]) & mask[m]; below += w < v; ties += w == v; }
            acc += 2ull * below + ties + 1ull;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) s2[p * 3 + m] = acc;
    }
}

cudaError_t launch_values(const ReadsDev &R, const void *resolved, int64_t n_units, const uint8_t *ins_seq, const vfk_params &P,
                          int32_t max_per_unit, uint32_t *values, int32_t *n_values, cudaStream_t stream, int *launches) {
    if (n_units <= 0) return cudaSuccess;
    const int64_t threads = n_units * 32;
    values_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(R, (const ResolvedUnit *)resolved, n_units, ins_seq, P, max_per_unit,
                                                                         values, n_values);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_ranksum(int64_t n_pairs, const uint32_t *alt, const int64_t *alt_off, const uint32_t *ref,
                           const int64_t *ref_off, unsigned long long *s2, cudaStream_t stream, int *launches) {
    if (n_pairs <= 0) return cudaSuccess;
    const int64_t threads = n_pairs * 32;
    ranksum_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(n_pairs, alt, alt_off, ref, ref_off, s2);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace vfk

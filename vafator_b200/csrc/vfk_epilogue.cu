// vfk_epilogue.cu -- FP64 statistics epilogue (sm_100a), one thread per (unit, statistic).
//
// Replaces, from the integer record the pileup kernel wrote:
//   vafator/rank_sum_test.py:6-12   scipy.stats.ranksums(x=alt, y=ref): mid-ranks, NO tie or
//                                   continuity correction, p = 2*ndtr(-|z|)        
//   vafator/power.py:39-50          pu = binom.cdf(ac, dp, eaf)
//   vafator/power.py:84-131         Carter-2012 k, d and power pw
// Rounding to 3/5 decimals is presentation and stays on the host.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "vfk.h"

namespace vfk {

// log(k!): exact products up to 20! (the largest factorial that is an exact double), lgamma beyond.  The
// epilogue reads it from a table covering every depth a shallow column can have; the table is filled on the
// device by the same function the fallback calls, so both agree bit for bit.
constexpr int LOG_FACTORIAL_TABLE = 1024;
__device__ double g_log_factorial[LOG_FACTORIAL_TABLE];
__constant__ double LOG_FACTORIAL_SMALL[21] = {0.0, 0.0, 0.6931471805599453, 1.791759469228055, 3.1780538303479458, 4.787491742782046, 6.579251212010101, 8.525161361065415, 10.60460290274525, 12.801827480081469, 15.104412573075516, 17.502307845873887, 19.987214495661885, 22.552163853123425, 25.19122118273868, 27.89927138384089, 30.671860106080672, 33.50507345013689, 36.39544520803305, 39.339884187199495, 42.335616460753485};
__device__ __forceinline__ double log_factorial_slow(int64_t k) { return k <= 20 ? LOG_FACTORIAL_SMALL[k] : lgamma((double)k + 1.0); }
__device__ __forceinline__ double log_factorial(int64_t k) {
    return k < LOG_FACTORIAL_TABLE ? __ldg(g_log_factorial + k) : log_factorial_slow(k);
}
__global__ void log_factorial_table_kernel() {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < LOG_FACTORIAL_TABLE) g_log_factorial[k] = log_factorial_slow(k);
}

// Everything about the success probability that the binomial terms share: evaluated once per unit.
struct Prob {
    double p, logp, log1mp, odds, inv_odds;
};
__device__ __forceinline__ Prob make_prob(double p) {
    Prob q;
    q.p = p;
    const bool interior = p > 0.0 && p < 1.0;
    q.logp = interior ? log(p) : 0.0;
    q.log1mp = interior ? log1p(-p) : 0.0;
    q.odds = interior ? p / (1.0 - p) : 1.0;
    q.inv_odds = interior ? (1.0 - p) / p : 1.0;
    return q;
}

// every thread evaluates several binomial terms for the same n (its depth): log(n!) is computed once
__device__ __forceinline__ double log_binom_pmf(int64_t k, int64_t n, const Prob &q, double lg_n1) {
    return lg_n1 - log_factorial(k) - log_factorial(n - k) + (double)k * q.logp + (double)(n - k) * q.log1mp;
}

__device__ bool dyadic_binom(int64_t k, int64_t n, double p, bool want_cdf, double &out);

__device__ double binom_pmf(int64_t k, int64_t n, const Prob &q, double lg_n1) {
    double exact;
    if (dyadic_binom(k, n, q.p, false, exact)) return exact;
    if (k < 0 || k > n) return 0.0;
    if (q.p <= 0.0) return k == 0 ? 1.0 : 0.0;
    if (q.p >= 1.0) return k == n ? 1.0 : 0.0;
    return exp(log_binom_pmf(k, n, q, lg_n1));
}

// Exact evaluation when p = a / 2^b and the numerators fit in 64 bits (b * n <= 62): the result is
// then the correctly rounded dyadic rational.  This matters for presentation, not for accuracy:
// with the default expected VAF of 0.5 values such as 1/64 sit exactly on a rounding tie of the
// 5-decimal INFO fields, and the reference (scipy/Boost) returns them exactly.
__device__ bool dyadic_binom(int64_t k, int64_t n, double p, bool want_cdf, double &out) {
    if (!(p > 0.0 && p < 1.0) || n < 0 || n > 62) return false;
    int b = 0;
    double a = p;
    while (b <= 10 && a != floor(a)) { a *= 2.0; ++b; }
    if (b > 10 || b == 0 || (int64_t)b * n > 62) return false;
    const uint64_t A = (uint64_t)a, B = (1ull << b) - A;
    if (k < 0) { out = 0.0; return true; }
    if (k > n) { out = want_cdf ? 1.0 : 0.0; return true; }
    uint64_t sum = 0, c = 1;  // c = C(n, i)
    for (int64_t i = 0; i <= k; ++i) {
        if (want_cdf || i == k) {
            uint64_t t = c;
            for (int64_t j = 0; j < i; ++j) t *= A;
            for (int64_t j = 0; j < n - i; ++j) t *= B;
            sum += t;
        }
        c = c * (uint64_t)(n - i) / (uint64_t)(i + 1);
    }
    out = ldexp((double)sum, -(int)(b * n));
    return true;
}

// lower tail P[X <= k] by direct summation of the probability mass, starting from the largest
// term of the tail that is summed so that every further term is smaller (no cancellation).  The term
// ratio is formed with a correctly rounded reciprocal (one ulp from the quotient, far inside the 1e-6
// parity tolerance) instead of two FP64 divisions per term.
__device__ double binom_cdf(int64_t k, int64_t n, const Prob &q, double lg_n1) {
    double exact;
    if (dyadic_binom(k, n, q.p, true, exact)) return exact;
    if (k < 0) return 0.0;
    if (k >= n) return 1.0;
    if (q.p <= 0.0) return 1.0;
    if (q.p >= 1.0) return 0.0;
    if ((double)k < (double)(n + 1) * q.p) {
        // k below the mode: sum k, k-1, ... downwards
        double term = exp(log_binom_pmf(k, n, q, lg_n1));
        double sum = term;
        for (int64_t i = k; i > 0; --i) {
            term *= (double)i * __drcp_rn((double)(n - i + 1)) * q.inv_odds;
            sum += term;
            if (term < sum * 1e-18) break;
        }
        return sum;
    }
    // upper tail k+1, k+2, ... then complement
    double term = exp(log_binom_pmf(k + 1, n, q, lg_n1));
    double sum = term;
    for (int64_t i = k + 1; i < n; ++i) {
        term *= (double)(n - i) * __drcp_rn((double)(i + 1)) * q.odds;
        sum += term;
        if (term < sum * 1e-18) break;
    }
    return 1.0 - sum;
}

// vafator/power.py:84-93 -- P(m, n) = 1 - binom.cdf(m-1, n, err/3), 1 when m < 1
__device__ __forceinline__ double carter_p(int64_t m, int64_t n, const Prob &e3, double lg_n1) {
    return m >= 1 ? 1.0 - binom_cdf(m - 1, n, e3, lg_n1) : 1.0;
}

// Carter-2012 k and d depend on the depth only (given FPR and error rate): the reference caches k per
// depth (vafator/power.py:106-114); here a table over depths [0, n_table) is filled once per
// (fpr, error_rate) and reused by every epilogue launch.
struct CarterEntry {
    double d;
    int32_t k;
    int32_t pad;
};

__device__ __forceinline__ CarterEntry carter_kd(int64_t dp, double fpr, const Prob &e3, double lg_n1) {
    // k = min{k >= 1 : P(k, dp) <= fpr}   (vafator/power.py:106-114), d as in :95-104
    int64_t k = 1;
    double pk = carter_p(1, dp, e3, lg_n1), pk_1 = 1.0;
    while (pk > fpr) {
        ++k;
        pk_1 = pk;
        pk = carter_p(k, dp, e3, lg_n1);
    }
    CarterEntry e;
    e.d = (fpr - pk) / (pk_1 - pk);
    e.k = (int32_t)k;
    e.pad = 0;
    return e;
}

__global__ void __launch_bounds__(128)
carter_table_kernel(int32_t n_table, double fpr, double error_rate, CarterEntry *__restrict__ table) {
    const int64_t dp = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (dp >= n_table) return;
    table[dp] = carter_kd(dp, fpr, make_prob(error_rate / 3.0), log_factorial_slow(dp));
}

// One thread per (unit, task); blockIdx.y is the task, so a warp runs one code path: tasks 0-2 the rank-sum
// z / p of MQ / BQ / position, task 3 pu, task 4 k + pw.  A unit's statistics are five independent chains of
// dependent FP64 operations: one thread per unit would run them back to back on a grid too small to hide
// the latency (100 k units = 21 warps per SM); split, the grid is five times larger and every chain a
// fifth as long.
constexpr int EPILOGUE_TASKS = 5;
__global__ void __launch_bounds__(128)
epilogue_kernel(int64_t n_units, const vfk_counts *__restrict__ counts, const double *__restrict__ eaf, double fpr,
                double error_rate, const CarterEntry *__restrict__ table, int32_t n_table, vfk_stats *__restrict__ out) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_units) return;
    const int task = blockIdx.y;
    const vfk_counts *c = counts + u;
    if (task < 3) {
        const int32_t ac_alt = c->ac_alt, ac_ref = c->ac_ref;
        const bool has_bq = (c->status & VFK_ST_KIND_MASK) == VFK_KIND_SNV;
        double z = nan(""), p = nan("");
        if (ac_alt > 0 && ac_ref > 0 && (task != 1 || has_bq)) {
            // scipy: s = sum(ranks of x); expected = n1*(n1+n2+1)/2.0; z = (s-expected)/sqrt(n1*n2*(n1+n2+1)/12.0)
            const double n1 = (double)ac_alt, n2 = (double)ac_ref;
            const double ranks = (double)c->s2[task] * 0.5;  // exact: multiples of 0.5 well below 2^53
            const double expected = n1 * (n1 + n2 + 1.0) / 2.0;
            const uint64_t prod = (uint64_t)ac_alt * (uint64_t)ac_ref * (uint64_t)(ac_alt + ac_ref + 1);
            z = (ranks - expected) / sqrt((double)prod / 12.0);
            p = erfc(fabs(z) * 0.70710678118654752440);  // 2*ndtr(-|z|)
        }
        out[u].z[task] = z;
        out[u].p[task] = p;
        return;
    }
    const int64_t dp = c->dp;
    const Prob f = make_prob(eaf[u]);
    const double lg_n1 = log_factorial(dp);
    if (task == 3) {
        out[u].pu = binom_cdf(c->ac_alt, dp, f, lg_n1);
        return;
    }
    const CarterEntry kd = (table && dp < n_table) ? table[dp] : carter_kd(dp, fpr, make_prob(error_rate / 3.0), lg_n1);
    const int64_t k = kd.k;
    out[u].pw = 1.0 - binom_cdf(k - 1, dp, f, lg_n1) + kd.d * binom_pmf(k, dp, f, lg_n1);
    out[u].k = (int32_t)k;
    out[u].pad = 0;
}

size_t carter_entry_size() { return sizeof(CarterEntry); }

cudaError_t launch_carter_table(int32_t n_table, double fpr, double error_rate, void *table, cudaStream_t stream, int *launches) {
    log_factorial_table_kernel<<<(LOG_FACTORIAL_TABLE + 127) / 128, 128, 0, stream>>>();  // per context, refilled with the Carter table
    ++*launches;
    carter_table_kernel<<<(unsigned)((n_table + 127) / 128), 128, 0, stream>>>(n_table, fpr, error_rate, (CarterEntry *)table);
    ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_epilogue(int64_t n_units, const vfk_counts *counts, const double *eaf, double fpr, double error_rate,
                            const void *table, int32_t n_table, vfk_stats *out, cudaStream_t stream, int *launches) {
    if (n_units <= 0) return cudaSuccess;
    const int block = 128;
    const int64_t grid = (n_units + block - 1) / block;
    epilogue_kernel<<<dim3((unsigned)grid, EPILOGUE_TASKS), block, 0, stream>>>(n_units, counts, eaf, fpr, error_rate,
                                                                              (const CarterEntry *)table, n_table, out);
    ++*launches;
    return cudaGetLastError();
}

}  // namespace vfk

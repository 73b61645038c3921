// vfk_api.cu -- the C ABI declared in include/vfk.h.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "vfk.h"
#ifdef _OPENMP
#include <omp.h>
#endif
#include "vfk_device.cuh"

namespace vfk {
cudaError_t launch_pileup(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int max_l_qseq, vfk_counts *out,
                          void *resolved_scratch, void *counters, uint32_t *lists, int sm_count, cudaStream_t stream,
                          int *launches);
cudaError_t launch_locate(const ReadsDev &R, const UnitsDev &V, void *resolved, unsigned long long *cand_sum, cudaStream_t stream,
                          int *launches);
cudaError_t launch_pileup_resolved(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int max_l_qseq, vfk_counts *out,
                                   void *resolved, void *counters, uint32_t *lists, int sm_count, cudaStream_t stream,
                                   int *launches);
cudaError_t launch_epilogue(int64_t n_units, const vfk_counts *counts, const double *eaf, double fpr, double error_rate,
                            const void *table, int32_t n_table, vfk_stats *out, cudaStream_t stream, int *launches);
cudaError_t launch_carter_table(int32_t n_table, double fpr, double error_rate, void *table, cudaStream_t stream, int *launches);
size_t carter_entry_size();
extern thread_local bool column_runs_gathered;  // vfk_pileup.cu: the next column launch reads per-unit gathered runs
cudaError_t launch_values(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int32_t max_per_unit, uint32_t *values,
                          int32_t *n_values, cudaStream_t stream, int *launches);
cudaError_t launch_ranksum(int64_t n_pairs, const uint32_t *alt, const int64_t *alt_off, const uint32_t *ref,
                           const int64_t *ref_off, unsigned long long *s2, cudaStream_t stream, int *launches);
}  // namespace vfk

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CU(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) return fail(VFK_ECUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return VFK_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(VFK_ENOMEM, "cudaMalloc: %s", cudaGetErrorString(e));
        cap = want;
        return VFK_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct PinBuf {  // page-locked host scratch
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return VFK_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e != cudaSuccess) return fail(VFK_ENOMEM, "cudaHostAlloc: %s", cudaGetErrorString(e));
        cap = want;
        return VFK_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// One of the two staging slots of the host entry point: its own stream, device staging buffers and
// scratch, so that the copies of one batch overlap the kernels of the previous one (double buffering).
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
    void *counters = nullptr;             // 32 B: work counter, deferred count, candidate sum
    unsigned long long *h_cand = nullptr;  // pinned host word for the candidate sum
    DevBuf deferred, resolved;
    PinBuf h_resolved;  // units located on the host (sparse variant sets), copied to `resolved`
    PinBuf h_col_gather, h_cunits2;  // VFK_HOST_GATHER=1: column runs gathered unit by unit, the units pointing at them
    DevBuf col_gather;
    DevBuf contig, hot, cold, mate, sb, cigar, payload;
    DevBuf tid, pos, meta, len, soff, ins, stop, eaf, counts, stats;
    void release() {
        DevBuf *bufs[] = {&deferred, &resolved, &contig, &hot, &cold, &mate, &sb, &cigar, &payload, &tid, &pos, &meta, &len,
                          &soff, &ins, &stop, &eaf, &counts, &stats};
        for (DevBuf *b : bufs) b->release();
        col_gather.release();
        h_col_gather.release();
        h_cunits2.release();
        h_resolved.release();
        if (counters) cudaFree(counters);
        if (h_cand) cudaFreeHost(h_cand);
        if (done) cudaEventDestroy(done);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct vfk_handle {
    int device = 0;
    int sm_count = 0;
    void *counters = nullptr;  // 32-byte scratch of the device-pointer API (layout: vfk_pileup.cu launch_posb)
    DevBuf deferred, resolved;  // per-launch scratch: deep-unit list, located units (32 B each)
    int64_t launches = 0;
    Slot slot[2];
    int next_slot = 0;
    int64_t zero_copy_batches = 0, zero_copy_hot_batches = 0, host_located_batches = 0, staged_bytes = 0, column_batches = 0;
    // Carter k / d per depth for the last (fpr, error_rate) seen, depths [0, CARTER_DEPTHS)
    DevBuf carter;
    double carter_fpr = -1.0, carter_err = -1.0;
    cudaEvent_t carter_ready = nullptr;
};

static const int32_t CARTER_DEPTHS = 4096;

// make sure the table matches (fpr, error_rate); work on `st` is ordered after its fill
static int ensure_carter(vfk_handle *h, double fpr, double error_rate, cudaStream_t st) {
    if (!h->carter_ready) CU(cudaEventCreateWithFlags(&h->carter_ready, cudaEventDisableTiming));
    if (h->carter_fpr != fpr || h->carter_err != error_rate || !h->carter.p) {
        CU(cudaDeviceSynchronize());  // rare: launches still reading the previous table must finish first
        int rc = h->carter.reserve((size_t)CARTER_DEPTHS * vfk::carter_entry_size());
        if (rc) return rc;
        int launches = 0;
        cudaError_t e = vfk::launch_carter_table(CARTER_DEPTHS, fpr, error_rate, h->carter.p, st, &launches);
        h->launches += launches;
        if (e != cudaSuccess) return fail(VFK_ECUDA, "carter table launch: %s", cudaGetErrorString(e));
        CU(cudaEventRecord(h->carter_ready, st));
        h->carter_fpr = fpr;
        h->carter_err = error_rate;
    } else {
        CU(cudaStreamWaitEvent(st, h->carter_ready, 0));
    }
    return VFK_OK;
}

extern "C" int vfk_version(void) { return VFK_VERSION; }
extern "C" const char *vfk_last_error(void) { return g_err; }

extern "C" int vfk_create(int device, vfk_handle **out) {
    if (!out) return fail(VFK_EINVAL, "vfk_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail(VFK_ENODEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(VFK_EINVAL, "vfk_create: device index out of range");
    CU(cudaSetDevice(device));
    vfk_handle *h = new (std::nothrow) vfk_handle();
    if (!h) return fail(VFK_ENOMEM, "vfk_create: out of host memory");
    h->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    {
        // The path gathers single 32-byte sectors (one payload sector per (read, column) pair):
        // ask L2 not to promote misses to 64/128-byte DRAM fetches.  A hint; failure is harmless.
        size_t gran = 32;
        if (const char *env = getenv("VFK_L2_FETCH_GRANULARITY")) gran = (size_t)atoi(env);
        if (gran == 32 || gran == 64 || gran == 128) (void)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
        (void)cudaGetLastError();
    }
    for (Slot &s : h->slot) {
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        CU(cudaMalloc(&s.counters, 32));
        CU(cudaMallocHost((void **)&s.h_cand, sizeof(unsigned long long)));
    }
    CU(cudaMalloc(&h->counters, 32));
    *out = h;
    return VFK_OK;
}

extern "C" int vfk_destroy(vfk_handle *h) {
    if (!h) return VFK_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (Slot &s : h->slot) s.release();
    h->deferred.release();
    h->resolved.release();
    h->carter.release();
    if (h->carter_ready) cudaEventDestroy(h->carter_ready);
    if (h->counters) cudaFree(h->counters);
    delete h;
    return VFK_OK;
}

extern "C" int vfk_sync(vfk_handle *h) {
    if (!h) return fail(VFK_EINVAL, "vfk_sync: NULL handle");
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    return VFK_OK;
}

extern "C" int64_t vfk_launch_count(vfk_handle *h) { return h ? h->launches : 0; }

static int check_batches(const vfk_read_batch *r, const vfk_variant_batch *v, const vfk_params *p) {
    if (!r || !v || !p) return fail(VFK_EINVAL, "NULL batch or params");
    if (r->n_reads < 0 || v->n_units < 0 || r->n_contigs < 0) return fail(VFK_EINVAL, "negative size");
    if (r->n_reads > 0x7FFFFFFF) return fail(VFK_EUNSUPPORTED, "more than 2^31-1 reads in one batch: split it");
    if (r->max_l_qseq > VFK_MAX_READ_LEN) return fail(VFK_EUNSUPPORTED, "read longer than VFK_MAX_READ_LEN");
    if (!r->contig_off) return fail(VFK_EINVAL, "contig_off is NULL");
    if (r->n_reads > 0 && (!r->hot || !r->cold || !r->mate || !r->sb || !r->payload))
        return fail(VFK_EINVAL, "read batch has a NULL array");
    if (v->n_units > 0 && (!v->tid || !v->pos || !v->meta || !v->length || !v->seq_off || !v->ins_seq))
        return fail(VFK_EINVAL, "variant batch has a NULL array");
    return VFK_OK;
}

static void to_dev(const vfk_read_batch *r, const vfk_variant_batch *v, vfk::ReadsDev &R, vfk::UnitsDev &V) {
    R.hot = (const uint4 *)r->hot;
    R.cold = (const uint4 *)r->cold;
    R.mate = r->mate;
    R.sb = (const int2 *)r->sb;
    R.cigar = r->cigar;
    R.payload = r->payload;
    R.contig_off = r->contig_off;
    R.n_reads = r->n_reads;
    R.n_contigs = r->n_contigs;
    const bool columns = r->col_off && r->col_wbase && r->col_sectors;
    R.col_wbase = columns ? r->col_wbase : nullptr;
    R.col_off = columns ? r->col_off : nullptr;
    R.col_sectors = columns ? (const uint32_t *)r->col_sectors : nullptr;
    V.n = v->n_units;
    V.tid = v->tid;
    V.pos = v->pos;
    V.meta = v->meta;
    V.length = v->length;
    V.seq_off = v->seq_off;
    V.ins_seq = v->ins_seq;
    V.stop = v->stop;
}

extern "C" int vfk_pileup(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                          const vfk_params *params, vfk_counts *out, void *stream) {
    if (!h) return fail(VFK_EINVAL, "vfk_pileup: NULL handle");
    int rc = check_batches(reads, units, params);
    if (rc) return rc;
    if (units->n_units == 0) return VFK_OK;
    if (!out) return fail(VFK_EINVAL, "vfk_pileup: out is NULL");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream
    rc = h->deferred.reserve((size_t)units->n_units * 2 * sizeof(uint32_t));
    if (rc) return rc;
    rc = h->resolved.reserve((size_t)units->n_units * 48);  // located unit + column unit
    if (rc) return rc;
    vfk::ReadsDev R;
    vfk::UnitsDev V;
    to_dev(reads, units, R, V);
    int launches = 0;
    cudaError_t e = vfk::launch_pileup(R, V, *params, reads->max_l_qseq, out, h->resolved.p, h->counters, (uint32_t *)h->deferred.p,
                                       h->sm_count, st, &launches);
    h->launches += launches;
    if (e != cudaSuccess) return fail(VFK_ECUDA, "pileup launch: %s", cudaGetErrorString(e));
    return VFK_OK;
}

extern "C" int vfk_epilogue(vfk_handle *h, int64_t n_units, const vfk_counts *counts, const double *eaf, double fpr,
                            double error_rate, vfk_stats *out, void *stream) {
    if (!h) return fail(VFK_EINVAL, "vfk_epilogue: NULL handle");
    if (n_units < 0) return fail(VFK_EINVAL, "vfk_epilogue: negative size");
    if (n_units == 0) return VFK_OK;
    if (!counts || !eaf || !out) return fail(VFK_EINVAL, "vfk_epilogue: NULL array");
    if (!(fpr >= 0.0) || !(error_rate >= 0.0)) return fail(VFK_EINVAL, "vfk_epilogue: fpr / error_rate must be >= 0");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;  // NULL = the CUDA default stream
    int rc = ensure_carter(h, fpr, error_rate, st);
    if (rc) return rc;
    int launches = 0;
    cudaError_t e = vfk::launch_epilogue(n_units, counts, eaf, fpr, error_rate, h->carter.p, CARTER_DEPTHS, out, st, &launches);
    h->launches += launches;
    if (e != cudaSuccess) return fail(VFK_ECUDA, "epilogue launch: %s", cudaGetErrorString(e));
    return VFK_OK;
}

extern "C" int vfk_pileup_values(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                                 const vfk_params *params, int32_t max_per_unit, uint32_t *values, int32_t *n_values,
                                 void *stream) {
    if (!h) return fail(VFK_EINVAL, "vfk_pileup_values: NULL handle");
    int rc = check_batches(reads, units, params);
    if (rc) return rc;
    if (units->n_units == 0) return VFK_OK;
    if (!values || !n_values || max_per_unit < 1) return fail(VFK_EINVAL, "vfk_pileup_values: bad output buffer");
    CU(cudaSetDevice(h->device));
    vfk::ReadsDev R;
    vfk::UnitsDev V;
    to_dev(reads, units, R, V);
    int launches = 0;
    cudaError_t e = vfk::launch_values(R, V, *params, max_per_unit, values, n_values, (cudaStream_t)stream, &launches);
    h->launches += launches;
    if (e != cudaSuccess) return fail(VFK_ECUDA, "values launch: %s", cudaGetErrorString(e));
    return VFK_OK;
}

extern "C" int vfk_ranksum_values(vfk_handle *h, int64_t n_pairs, const uint32_t *alt, const int64_t *alt_off,
                                  const uint32_t *ref, const int64_t *ref_off, uint64_t *s2, void *stream) {
    if (!h) return fail(VFK_EINVAL, "vfk_ranksum_values: NULL handle");
    if (n_pairs < 0) return fail(VFK_EINVAL, "vfk_ranksum_values: negative size");
    if (n_pairs == 0) return VFK_OK;
    if (!alt || !alt_off || !ref || !ref_off || !s2) return fail(VFK_EINVAL, "vfk_ranksum_values: NULL array");
    CU(cudaSetDevice(h->device));
    int launches = 0;
    cudaError_t e = vfk::launch_ranksum(n_pairs, alt, alt_off, ref, ref_off, (unsigned long long *)s2, (cudaStream_t)stream,
                                        &launches);
    h->launches += launches;
    if (e != cudaSuccess) return fail(VFK_ECUDA, "ranksum launch: %s", cudaGetErrorString(e));
    return VFK_OK;
}

#define STAGE(buf, src, bytes)                                                                   \
    do {                                                                                         \
        size_t b_ = (size_t)(bytes);                                                             \
        int rc_ = (buf).reserve(b_ ? b_ : 16);                                                   \
        if (rc_) return rc_;                                                                     \
        if (b_) CU(cudaMemcpyAsync((buf).p, (src), b_, cudaMemcpyHostToDevice, st));             \
        h->staged_bytes += (int64_t)b_;                                                          \
    } while (0)

// device-visible alias of a page-locked host buffer, or NULL
static const void *pinned_alias(const void *host) {
    cudaPointerAttributes attr;
    const void *dev = nullptr;
    if (host && cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost) dev = attr.devicePointer;
    (void)cudaGetLastError();
    return dev;
}

// Host twin of locate_kernel (vfk_pileup.cu) for the host entry point: when the variant set is sparse the
// position index (8 bytes per read) is not worth shipping -- the units are located here, from the
// caller's host arrays, and only the 32-byte records cross the bus.  VCF order makes consecutive units
// neighbours in the index, so each search gallops forward from the previous answer.
static unsigned long long locate_host(const vfk_read_batch *r, const vfk_variant_batch *v, vfk::ResolvedUnit *out, uint4 *cunits) {
    const int64_t nu = v->n_units;
    const int32_t *sb = r->sb;
    unsigned long long total = 0;
#pragma omp parallel reduction(+ : total)
    {
        int nt = 1, me = 0;
#ifdef _OPENMP
        nt = omp_get_num_threads();
        me = omp_get_thread_num();
#endif
        const int64_t u0 = nu * me / nt, u1 = nu * (me + 1) / nt;
        int32_t prev_tid = -1, prev_col = 0;
        int64_t prev_a = 0;
        for (int64_t u = u0; u < u1; ++u) {
            const int32_t tid = v->tid[u], col = v->pos[u];
            const bool known = tid >= 0 && tid < r->n_contigs;
            const int64_t c0 = known ? r->contig_off[tid] : 0, c1 = known ? r->contig_off[tid + 1] : 0;
            int64_t a = c0, b = c1;  // first read of the contig starting beyond the column lies in [a, b]
            if (known && tid == prev_tid && col >= prev_col) {
                a = prev_a;
                int64_t step = 32;
                while (a + step < c1 && sb[2 * (a + step)] <= col) {
                    a += step;
                    step <<= 1;
                }
                b = a + step < c1 ? a + step : c1;
            }
            while (a < b) {
                const int64_t m = (a + b) >> 1;
                if (sb[2 * m] > col) b = m; else a = m + 1;
            }
            vfk::ResolvedUnit q;
            q.col = col;
            q.meta = v->meta[u];
            q.len = v->length[u];
            q.ins_off = v->seq_off[u];
            q.hi = (uint32_t)a;
            q.lo = a > c0 ? (uint32_t)sb[2 * (a - 1) + 1] : (uint32_t)a;
            q.stop = v->stop ? v->stop[u] : 0x7FFFFFFF;
            q.c1 = (uint32_t)c1;
            out[u] = q;
            if (cunits) {  // column store: the run of sectors of the column's window (as locate_kernel does)
                const int32_t w = col / VFK_COL_WIDTH;
                uint32_t off = 0, cnt = 0;
                if (known && col >= 0 && r->col_wbase[tid] + w < r->col_wbase[tid + 1]) {
                    off = r->col_off[r->col_wbase[tid] + w];
                    cnt = r->col_off[r->col_wbase[tid] + w + 1] - off;
                }
                cunits[u] = make_uint4(off, cnt, (uint32_t)(col - w * VFK_COL_WIDTH), q.meta);
            }
            total += q.hi - q.lo;
            prev_tid = known ? tid : -1;
            prev_col = col;
            prev_a = a;
        }
    }
    return total;
}

// Host half of VFK_HOST_GATHER (see vfk_annotate_host_async): per column unit {first sector of the window's run, sectors,
// position in the window, meta} -> a compact run of its own: the header pairs of the run, then the entries of ITS
// position (10 bytes per read instead of 32), padded to 32 bytes; units_out points the units at their compact runs.
extern "C" int vfk_host_gather_columns(const uint32_t *units_in, int64_t n_units, const uint8_t *sectors, int transposed,
                                       uint32_t *units_out, uint8_t *out, int64_t out_cap, int64_t *out_bytes) {
    if (!units_in || !units_out || !out_bytes || n_units < 0 || (!sectors && n_units)) return fail(VFK_EINVAL, "vfk_host_gather_columns: NULL argument");
    const uint4 *cu = reinterpret_cast<const uint4 *>(units_in);
    uint4 *cu2 = reinterpret_cast<uint4 *>(units_out);
    uint64_t acc = 0;  // in 32-byte sectors
    for (int64_t u = 0; u < n_units; ++u) {
        const uint32_t cnt = cu[u].y, kind = cu[u].w & 3u;
        const uint32_t g = (kind == VFK_KIND_SNV && cnt <= 64u) ? cnt : 0u;
        // not served by the column kernel: no run (other kinds), or a count that only says "too deep" (the kernel may
        // still prefetch lanes of it: offset 0 is always backed by 4 KB)
        cu2[u] = make_uint4(g ? (uint32_t)acc : 0u, kind != VFK_KIND_SNV ? 0u : (cnt <= 64u ? cnt : 65u), cu[u].z, cu[u].w);
        acc += (10u * g + 31u) / 32u;
    }
    if (acc >= (1ull << 32)) return fail(VFK_EUNSUPPORTED, "gathered column runs exceed 2^32 sectors");
    const int64_t bytes = (int64_t)(acc < 128 ? 128 : acc) * 32;
    *out_bytes = bytes;
    if (!out) return VFK_OK;  // sizing pass
    if (out_cap < bytes) return fail(VFK_EINVAL, "vfk_host_gather_columns: output buffer too small");
#pragma omp parallel for schedule(static)
    for (int64_t u = 0; u < n_units; ++u) {
        const uint32_t g = (cu2[u].y >= 1u && cu2[u].y <= 64u) ? cu2[u].y : 0u;
        if (!g) continue;
        const uint32_t p = cu[u].z;
        const uint8_t *src = sectors + (size_t)cu[u].x * 32;
        uint8_t *hdr = out + (size_t)cu2[u].x * 32, *ent = hdr + (size_t)g * 8;
        if (transposed) {
            memcpy(hdr, src, (size_t)g * 8);
            memcpy(ent, src + (size_t)g * 8 + (size_t)p * g * 2, (size_t)g * 2);
        } else {
            for (uint32_t k = 0; k < g; ++k) {
                memcpy(hdr + (size_t)k * 8, src + (size_t)k * 32 + 24, 8);
                memcpy(ent + (size_t)k * 2, src + (size_t)k * 32 + 2 * p, 2);
            }
        }
    }
    return VFK_OK;
}

extern "C" int vfk_wait(vfk_handle *h, int ticket) {
    if (!h || ticket < 0 || ticket > 1) return fail(VFK_EINVAL, "vfk_wait: bad handle or ticket");
    Slot &s = h->slot[ticket];
    if (!s.busy) return VFK_OK;
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(s.done));
    s.busy = false;
    return VFK_OK;
}

extern "C" int vfk_annotate_host_async(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                                       const vfk_params *params, const double *eaf, double fpr, double error_rate,
                                       vfk_counts *counts_out, vfk_stats *stats_out, int *ticket) {
    if (!h) return fail(VFK_EINVAL, "vfk_annotate_host: NULL handle");
    int rc = check_batches(reads, units, params);
    if (rc) return rc;
    const int64_t nu = units->n_units, nr = reads->n_reads;
    const int si = h->next_slot;
    if (ticket) *ticket = si;
    if (nu == 0) return VFK_OK;
    if (!eaf || !counts_out || !stats_out) return fail(VFK_EINVAL, "vfk_annotate_host: NULL array");
    rc = vfk_wait(h, si);  // the slot's previous batch must have left it
    if (rc) return rc;
    h->next_slot ^= 1;
    CU(cudaSetDevice(h->device));
    Slot &S = h->slot[si];
    cudaStream_t st = S.stream;
    unsigned long long *cand_sum = (unsigned long long *)((char *)S.counters + 24);
    // ---- locate.  Sparse variant set (the usual case: a VCF against a whole BAM): on the host, no index
    // transfer; otherwise the index and the unit arrays are staged and the locate kernel runs.
    rc = S.deferred.reserve((size_t)nu * 2 * sizeof(uint32_t));
    if (rc) return rc;
    rc = S.resolved.reserve((size_t)nu * 48);  // located unit + column unit
    if (rc) return rc;
    STAGE(S.ins, units->ins_seq, (size_t)(units->n_ins_seq > 0 ? units->n_ins_seq : 1));
    vfk_read_batch dr = *reads;
    dr.col_wbase = nullptr;  // the host entry point reads the read-major arrays only
    dr.col_off = nullptr;
    dr.col_sectors = nullptr;
    vfk_variant_batch dv = *units;
    dv.ins_seq = (const uint8_t *)S.ins.p;
    int launches = 0;
    CU(cudaMemsetAsync(S.counters, 0, 32, st));
    const char *env_loc = getenv("VFK_HOST_LOCATE");
    const bool host_locate = env_loc ? env_loc[0] == '1' : nu * 64 <= nr;
    unsigned long long candidates = 0;
    // The column store takes part when it can be gathered in place: units located on the host (its window table
    // is a host array) and every buffer the kernels touch page-locked.
    const void *z_col = (reads->col_wbase && reads->col_off && reads->col_sectors) ? pinned_alias(reads->col_sectors) : nullptr;
    const char *env_col = getenv("VFK_HOST_COLUMNS");
    const bool host_columns = host_locate && z_col && pinned_alias(reads->payload) && pinned_alias(reads->cold) &&
                              pinned_alias(reads->hot) && pinned_alias(reads->mate) && !(env_col && env_col[0] == '0') &&
                              !getenv("VFK_ZERO_COPY_PAYLOAD") && !getenv("VFK_ZERO_COPY_HOT");
    if (host_locate) {
        const size_t per_unit = sizeof(vfk::ResolvedUnit) + (host_columns ? sizeof(uint4) : 0);
        rc = S.h_resolved.reserve((size_t)nu * per_unit);
        if (rc) return rc;
        uint4 *h_cunits = host_columns ? reinterpret_cast<uint4 *>((char *)S.h_resolved.p + (size_t)nu * sizeof(vfk::ResolvedUnit)) : nullptr;
        candidates = locate_host(reads, units, (vfk::ResolvedUnit *)S.h_resolved.p, h_cunits);
        CU(cudaMemcpyAsync(S.resolved.p, S.h_resolved.p, (size_t)nu * per_unit, cudaMemcpyHostToDevice, st));
        h->staged_bytes += (int64_t)nu * (int64_t)per_unit;
        ++h->host_located_batches;
    } else {
        STAGE(S.contig, reads->contig_off, (size_t)(reads->n_contigs + 1) * sizeof(int64_t));
        STAGE(S.sb, reads->sb, (size_t)nr * 8);
        STAGE(S.tid, units->tid, (size_t)nu * 4);
        STAGE(S.pos, units->pos, (size_t)nu * 4);
        STAGE(S.meta, units->meta, (size_t)nu * 4);
        STAGE(S.len, units->length, (size_t)nu * 4);
        STAGE(S.soff, units->seq_off, (size_t)nu * 4);
        if (units->stop) STAGE(S.stop, units->stop, (size_t)nu * 4);
        dr.contig_off = (const int64_t *)S.contig.p;
        dr.sb = (const int32_t *)S.sb.p;
        dv.tid = (const int32_t *)S.tid.p;
        dv.pos = (const int32_t *)S.pos.p;
        dv.meta = (const uint32_t *)S.meta.p;
        dv.length = (const int32_t *)S.len.p;
        dv.seq_off = (const uint32_t *)S.soff.p;
        dv.stop = units->stop ? (const int32_t *)S.stop.p : nullptr;
    }
    vfk::ReadsDev R;
    vfk::UnitsDev V;
    to_dev(&dr, &dv, R, V);
    if (!host_locate) {
        cudaError_t e = vfk::launch_locate(R, V, S.resolved.p, cand_sum, st, &launches);
        if (e != cudaSuccess) return fail(VFK_ECUDA, "locate launch: %s", cudaGetErrorString(e));
    }
    // ---- transfer strategy for the bulk of the batch (payload ~84 %, headers ~10 %, cold records + CIGARs
    // ~6 %): a sparse variant set touches a few percent of those bytes, so when the caller's buffers are
    // page-locked the kernel gathers the sectors it needs straight from host memory over PCIe
    // ("zero copy") instead of staging everything.  Decided from the located candidate count.
    const void *z_payload = pinned_alias(reads->payload), *z_cold = pinned_alias(reads->cold),
               *z_cigar = reads->n_cigar_words ? pinned_alias(reads->cigar) : reads->cigar;
    bool zero_copy = z_payload && z_cold && (z_cigar || reads->n_cigar_words == 0);
    if (const char *env = getenv("VFK_ZERO_COPY_PAYLOAD")) zero_copy = zero_copy && env[0] == '1';
    if (zero_copy && !getenv("VFK_ZERO_COPY_PAYLOAD")) {
        if (!host_locate) {
            CU(cudaMemcpyAsync(S.h_cand, cand_sum, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            candidates = *S.h_cand;
        }
        // ~64 bytes cross the bus per candidate read when gathering; staging moves everything once
        const double gathered = (double)candidates * 64.0;
        const double staged = (double)reads->n_sectors * 32.0 + (double)nr * 16.0 + (double)reads->n_cigar_words * 4.0;
        zero_copy = gathered * 4.0 < staged;
    }
    const bool use_columns = host_columns && zero_copy;
    if (use_columns) {  // SNV units: one contiguous run of sectors per column, read in place; the rest as before
        R.col_wbase = reads->col_wbase;  // not dereferenced on the device (the units are located)
        R.col_off = (const uint32_t *)pinned_alias(reads->col_off);
        if (!R.col_off) R.col_off = (const uint32_t *)z_col;  // only its presence matters to the launcher
        R.col_sectors = (const uint32_t *)z_col;
        ++h->column_batches;
    }
    // EXPERIMENTAL (VFK_HOST_GATHER=1, not yet verified on a GPU): instead of letting the column kernel read the sector
    // runs in place over PCIe (32 bytes per read of the column, of which it uses 10), the host gathers, unit by unit,
    // the header pairs of the run and the entries of the unit's own position into one compact page-locked buffer that
    // crosses the bus as a single bulk copy; the units are re-pointed at their compact runs (pileup_column_kernel<.., CP>).
    bool gathered = false;
    {
        const char *env_g = getenv("VFK_HOST_GATHER");
        if (use_columns && env_g && env_g[0] == '1') {
            const char *env_t = getenv("VFK_COL_TRANSPOSED");
            const bool src_transposed = env_t && env_t[0] == '1';  // form of the caller's store (device.pack_reads)
            const uint32_t *cu = reinterpret_cast<const uint32_t *>((const char *)S.h_resolved.p + (size_t)nu * sizeof(vfk::ResolvedUnit));
            rc = S.h_cunits2.reserve((size_t)nu * sizeof(uint4));
            if (rc) return rc;
            int64_t need = 0;
            rc = vfk_host_gather_columns(cu, nu, reads->col_sectors, src_transposed ? 1 : 0, (uint32_t *)S.h_cunits2.p, nullptr, 0, &need);
            if (rc) return rc;
            const size_t bytes = (size_t)need;
            rc = S.h_col_gather.reserve(bytes);
            if (rc) return rc;
            rc = S.col_gather.reserve(bytes);
            if (rc) return rc;
            rc = vfk_host_gather_columns(cu, nu, reads->col_sectors, src_transposed ? 1 : 0, (uint32_t *)S.h_cunits2.p,
                                         (uint8_t *)S.h_col_gather.p, need, &need);
            if (rc) return rc;
            CU(cudaMemcpyAsync(S.col_gather.p, S.h_col_gather.p, bytes, cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync((char *)S.resolved.p + (size_t)nu * sizeof(vfk::ResolvedUnit), S.h_cunits2.p, (size_t)nu * sizeof(uint4),
                               cudaMemcpyHostToDevice, st));  // after the copy of the located units: replaces their column part
            h->staged_bytes += (int64_t)bytes + (int64_t)nu * (int64_t)sizeof(uint4);
            R.col_sectors = (const uint32_t *)S.col_gather.p;
            gathered = true;
        }
    }
    // ---- stage 2
    const void *z_hot = pinned_alias(reads->hot), *z_mate = pinned_alias(reads->mate);
    const char *env_hot = getenv("VFK_ZERO_COPY_HOT");  // the 20-byte headers follow the payload unless told otherwise
    const bool zero_copy_hot = zero_copy && z_hot && z_mate && !(env_hot && env_hot[0] == '0');
    if (!zero_copy_hot) {
        STAGE(S.hot, reads->hot, (size_t)nr * sizeof(vfk_read_hot));
        STAGE(S.mate, reads->mate, (size_t)nr * 4);
    }
    STAGE(S.eaf, eaf, (size_t)nu * sizeof(double));
    if (zero_copy) {
        R.cold = (const uint4 *)z_cold;
        R.cigar = (const uint32_t *)z_cigar;
        R.payload = (const uint8_t *)z_payload;
        ++h->zero_copy_batches;
        if (zero_copy_hot) ++h->zero_copy_hot_batches;
    } else {
        STAGE(S.cold, reads->cold, (size_t)nr * sizeof(vfk_read_cold));
        STAGE(S.cigar, reads->cigar, (size_t)reads->n_cigar_words * 4);
        STAGE(S.payload, reads->payload, (size_t)reads->n_sectors * 32);
        R.cold = (const uint4 *)S.cold.p;
        R.cigar = (const uint32_t *)S.cigar.p;
        R.payload = (const uint8_t *)S.payload.p;
    }
    R.hot = zero_copy_hot ? (const uint4 *)z_hot : (const uint4 *)S.hot.p;
    R.mate = zero_copy_hot ? (const uint32_t *)z_mate : (const uint32_t *)S.mate.p;
    rc = S.counts.reserve((size_t)nu * sizeof(vfk_counts));
    if (rc) return rc;
    rc = S.stats.reserve((size_t)nu * sizeof(vfk_stats));
    if (rc) return rc;
    vfk::column_runs_gathered = gathered;
    cudaError_t e = vfk::launch_pileup_resolved(R, V, *params, reads->max_l_qseq, (vfk_counts *)S.counts.p, S.resolved.p, S.counters,
                                                (uint32_t *)S.deferred.p, h->sm_count, st, &launches);
    vfk::column_runs_gathered = false;
    if (e != cudaSuccess) return fail(VFK_ECUDA, "pileup launch: %s", cudaGetErrorString(e));
    rc = ensure_carter(h, fpr, error_rate, st);
    if (rc) return rc;
    e = vfk::launch_epilogue(nu, (const vfk_counts *)S.counts.p, (const double *)S.eaf.p, fpr, error_rate, h->carter.p, CARTER_DEPTHS,
                             (vfk_stats *)S.stats.p, st, &launches);
    if (e != cudaSuccess) return fail(VFK_ECUDA, "epilogue launch: %s", cudaGetErrorString(e));
    h->launches += launches;
    CU(cudaMemcpyAsync(counts_out, S.counts.p, (size_t)nu * sizeof(vfk_counts), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(stats_out, S.stats.p, (size_t)nu * sizeof(vfk_stats), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(S.done, st));
    S.busy = true;
    return VFK_OK;
}

extern "C" int vfk_annotate_host(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                                 const vfk_params *params, const double *eaf, double fpr, double error_rate,
                                 vfk_counts *counts_out, vfk_stats *stats_out) {
    int ticket = 0;
    int rc = vfk_annotate_host_async(h, reads, units, params, eaf, fpr, error_rate, counts_out, stats_out, &ticket);
    if (rc) return rc;
    return vfk_wait(h, ticket);
}

extern "C" int64_t vfk_zero_copy_batches(vfk_handle *h) { return h ? h->zero_copy_batches : 0; }
extern "C" int64_t vfk_column_batches(vfk_handle *h) { return h ? h->column_batches : 0; }
extern "C" int vfk_transfer_stats(vfk_handle *h, int64_t out[4]) {
    if (!h || !out) return fail(VFK_EINVAL, "vfk_transfer_stats: NULL argument");
    out[0] = h->staged_bytes;
    out[1] = h->zero_copy_batches;
    out[2] = h->zero_copy_hot_batches;
    out[3] = h->host_located_batches;
    return VFK_OK;
}

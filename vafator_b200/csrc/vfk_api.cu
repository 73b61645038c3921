// vfk_api.cu -- the C ABI declared in include/vfk.h.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <omp.h>
#include "vfk.h"
#include "vfk_device.cuh"

namespace vfk {
struct LaunchCache {
    bool ready;
    int warp_ctas, tail_ctas;
};
struct PileupScratch {
    void *resolved;
    void *counters;
    uint32_t *lists;
    uint32_t *heads;
    uint32_t n_buckets;
    uint32_t *next;
    int32_t *column;
    uint32_t *claimed;
    uint32_t *mate;
    cudaEvent_t *ev;
    LaunchCache *cache;
    uint32_t fetch_cap;
    int fetch_by_index;
    uint32_t *fetch_base;
    void *fetch_req;
    const uint16_t *fetched;
    cudaError_t (*fetch_between)(void *ctx, const uint32_t *pair_counter, cudaStream_t stream);
    void *fetch_ctx;
};
cudaError_t launch_prep(const ReadsDev &R, uint2 *ee_out, int64_t *chunk_max, int64_t *blk_max, int64_t *blk_incl, uint32_t *batch_status,
                        cudaStream_t stream, int *launches);
cudaError_t launch_offsets(const ReadsDev &R, long long *blk_sums, int64_t *cigar_off, int64_t *seq_off, int64_t *qual_off, cudaStream_t stream,
                           int *launches);
cudaError_t launch_pileup(const ReadsDev &R, const UnitsDev &V, const vfk_params &P, int max_l_qseq, vfk_counts *out,
                          const PileupScratch &S, int sm_count, cudaStream_t stream, int *launches, bool link_all);
cudaError_t launch_locate_only(const ReadsDev &R, const UnitsDev &V, void *resolved, cudaStream_t stream, int *launches);
cudaError_t launch_link(const ReadsDev &R, const vfk_params &P, const void *resolved, const uint32_t *list, const uint32_t *n_list,
                        int64_t n_all, uint32_t *heads, uint32_t n_buckets, uint32_t *next, int32_t *column, uint32_t *claimed,
                        uint32_t *mate, int sm_count, cudaStream_t stream, int *launches);
cudaError_t launch_epilogue(int64_t n_units, const vfk_counts *counts, const double *eaf, double fpr, double error_rate,
                            const void *table, int32_t n_table, vfk_stats *out, cudaStream_t stream, int *launches);
cudaError_t launch_carter_table(int32_t n_table, double fpr, double error_rate, void *table, cudaStream_t stream, int *launches);
size_t carter_entry_size();
cudaError_t launch_values(const ReadsDev &R, const void *resolved, int64_t n_units, const uint8_t *ins_seq, const vfk_params &P,
                          int32_t max_per_unit, uint32_t *values, int32_t *n_values, cudaStream_t stream, int *launches);
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

static const size_t COUNTER_BYTES = 64 + 256 * 4;  // per-call counters + one work counter per SM segment (vfk_pileup.cu)

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


// Per-stream working memory of one call: the words prep_kernel derives per read, the located units, the device-side
// lists and the link scratch.  The bucket heads and the claim bitmap are kept clean by the link kernels themselves
// (set once when (re)allocated).
struct Work {
    DevBuf off_c, off_s, off_q, off_sums;  // derived pool offsets of a batch that leaves them NULL
    DevBuf ee, chunk_max, blk_max, blk_incl, resolved, lists, next, column, claimed, mate, heads;
    uint32_t n_buckets = 0;
    void *counters = nullptr;  // COUNTER_BYTES, layout in vfk_pileup.cu (PileupScratch)
    cudaEvent_t tev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // phase boundaries when timing is on
    bool ev_epilogue = false;
    int reserve(int64_t n_reads, int64_t n_units, cudaStream_t st) {
        const size_t nr = (size_t)(n_reads > 0 ? n_reads : 1), nu = (size_t)(n_units > 0 ? n_units : 1);
        const size_t nb = (nr + vfk::PREP_BLOCK - 1) / vfk::PREP_BLOCK;
        int rc;
        if ((rc = ee.reserve(nr * 8)) || (rc = chunk_max.reserve((nr / 32 + 1) * 8)) || (rc = blk_max.reserve(nb * 8)) || (rc = blk_incl.reserve((nb + 1) * 8)) ||
            (rc = resolved.reserve(nu * sizeof(vfk::ResolvedUnit))) || (rc = lists.reserve(nu * 3 * sizeof(uint32_t))) ||
            (rc = next.reserve(nr * 4)) || (rc = column.reserve(nr * 4)) || (rc = mate.reserve(nr * 4)))
            return rc;
        const size_t claim_bytes = (nr / 32 + 1) * 4;
        if (claim_bytes > claimed.cap) {
            if ((rc = claimed.reserve(claim_bytes))) return rc;
            if (cudaMemsetAsync(claimed.p, 0, claimed.cap, st) != cudaSuccess) return fail(VFK_ECUDA, "cudaMemsetAsync(claimed)");
        }
        uint64_t want = 1024;
        while (want < 2 * (uint64_t)nr && want < (1ull << 28)) want <<= 1;
        if (want > n_buckets) {
            if ((rc = heads.reserve((size_t)want * 4))) return rc;
            if (cudaMemsetAsync(heads.p, 0xFF, (size_t)want * 4, st) != cudaSuccess) return fail(VFK_ECUDA, "cudaMemsetAsync(heads)");
            n_buckets = (uint32_t)want;
        }
        return VFK_OK;
    }
    vfk::PileupScratch scratch() const {
        vfk::PileupScratch s;
        s.resolved = resolved.p; s.counters = counters; s.lists = (uint32_t *)lists.p;
        s.heads = (uint32_t *)heads.p; s.n_buckets = n_buckets; s.next = (uint32_t *)next.p; s.column = (int32_t *)column.p;
        s.claimed = (uint32_t *)claimed.p; s.mate = (uint32_t *)mate.p;
        s.ev = nullptr;
        s.cache = nullptr;
        s.fetch_cap = 0u; s.fetch_by_index = 0; s.fetch_base = nullptr; s.fetch_req = nullptr; s.fetched = nullptr; s.fetch_between = nullptr; s.fetch_ctx = nullptr;
        return s;
    }
    void release() {
        DevBuf *bufs[] = {&off_c, &off_s, &off_q, &off_sums, &ee, &chunk_max, &blk_max, &blk_incl, &resolved, &lists, &next, &column, &claimed, &mate, &heads};
        for (DevBuf *b : bufs) b->release();
        n_buckets = 0;
        if (counters) cudaFree(counters);
        counters = nullptr;
        for (cudaEvent_t &e : tev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
    }
};

// The host's part of a fetch list (vfk_pileup.cu request_kernel): between the two kernels of a call, in stream order, the bytes
// the device asked for are gathered from the caller's arrays into a page-locked buffer that one bulk copy takes to the device.
struct FetchJob {
    const ulonglong2 *req = nullptr;  // written by request_kernel
    uint16_t *out = nullptr;          // quality | sequence byte << 8 per pair
    const uint32_t *h_total = nullptr;
    uint32_t cap = 0;
    const uint8_t *qual = nullptr, *seq = nullptr;
    const int64_t *qual_off = nullptr, *seq_off = nullptr;  // the caller's offset arrays when requests go by read index, else NULL
    int64_t n_reads = 0, n_qual = 0, n_seq = 0;
    int threads = 1;
    int64_t *pairs = nullptr, *nanos = nullptr;  // handle statistics: pairs gathered, wall time of the gathers
};

static void CUDART_CB fetch_cb(void *p) {
    const FetchJob &j = *(const FetchJob *)p;
    const double t0 = omp_get_wtime();
    const int64_t n = (int64_t)(*j.h_total < j.cap ? *j.h_total : j.cap);
    const int64_t n_blk = (n + 63) / 64;
    const unsigned long long none = ~0ull;
    // a thread owns a contiguous run of 64-pair blocks and works one block ahead: the addresses of block b + 1 are computed and
    // every line they name is requested before the bytes of block b are picked up (two random cache lines per pair; nothing
    // but the number of lines in flight per core bounds this loop)
#pragma omp parallel num_threads(j.threads)
    {
        const int64_t t = omp_get_thread_num(), nt = omp_get_num_threads();
        const int64_t b0 = n_blk * t / nt, b1 = n_blk * (t + 1) / nt;
        uint64_t qa[2][64], sa[2][64];
        auto ask = [&](int64_t b) {
            const int64_t p0 = b * 64, p1 = p0 + 64 < n ? p0 + 64 : n;
            uint64_t *qq = qa[b & 1], *ss = sa[b & 1];
            for (int64_t q = p0; q < p1; ++q) {
                const ulonglong2 r = j.req[q];
                uint64_t a = none, c = 0;
                if (r.x != none) {
                    if (j.qual_off) {
                        if ((int64_t)r.x < j.n_reads) { a = (uint64_t)j.qual_off[r.x] + r.y; c = (uint64_t)j.seq_off[r.x] + (r.y >> 1); }
                    } else {
                        a = r.x; c = r.y;
                    }
                    if (a >= (uint64_t)j.n_qual || c >= (uint64_t)j.n_seq) a = none;  // a malformed batch must not take the host down
                }
                qq[q - p0] = a; ss[q - p0] = c;
                if (a != none) {
                    __builtin_prefetch(j.qual + a);
                    __builtin_prefetch(j.seq + c);
                }
            }
        };
        if (b0 < b1) ask(b0);
        for (int64_t b = b0; b < b1; ++b) {
            if (b + 1 < b1) ask(b + 1);
            const int64_t p0 = b * 64, p1 = p0 + 64 < n ? p0 + 64 : n;
            const uint64_t *qq = qa[b & 1], *ss = sa[b & 1];
            for (int64_t q = p0; q < p1; ++q) {
                const uint64_t a = qq[q - p0];
                j.out[q] = a == none ? (uint16_t)0 : (uint16_t)(j.qual[a] | ((uint16_t)j.seq[ss[q - p0]] << 8));
            }
        }
    }
    if (j.pairs) __atomic_fetch_add(j.pairs, n, __ATOMIC_RELAXED);
    if (j.nanos) __atomic_fetch_add(j.nanos, (int64_t)((omp_get_wtime() - t0) * 1e9), __ATOMIC_RELAXED);
}

// One of the two staging slots of the host entry point: its own stream, working memory and device staging
// buffers, so that the copies of one batch overlap the kernels of the previous one (double buffering).
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    bool busy = false;
    uint32_t *h_status = nullptr;  // pinned: [0] the batch status word of the slot's last batch, [1] pairs on its fetch list
    Work work;
    // Pageable unit arrays / result buffers would make their "asynchronous" copies synchronous (the call would return only once
    // the whole batch is through, and the two slots would never overlap): they go through page-locked bounce buffers, the
    // results reaching the caller's arrays in vfk_wait.
    PinBuf h_in, h_counts, h_stats;
    vfk_counts *user_counts = nullptr;
    vfk_stats *user_stats = nullptr;
    int64_t user_n = 0;
    PinBuf fetch_req, fetch_out;   // fetch list: requests (written by the device), gathered bytes
    DevBuf fetch_base, fetched;
    FetchJob job;
    DevBuf contig, pos, flag, mapq, l_qseq, n_cigar, cigar_off, seq_off, qual_off, cigar, seq, qual, mtid, mpos, isize, qname_id, qname_hash;
    DevBuf tid, upos, meta, len, soff, ins, stop, eaf, counts, stats;
    void release() {
        DevBuf *bufs[] = {&contig, &pos, &flag, &mapq, &l_qseq, &n_cigar, &cigar_off, &seq_off, &qual_off, &cigar, &seq, &qual, &mtid, &mpos,
                          &isize, &qname_id, &qname_hash, &tid, &upos, &meta, &len, &soff, &ins, &stop, &eaf, &counts, &stats};
        for (DevBuf *b : bufs) b->release();
        fetch_base.release(); fetched.release(); fetch_req.release(); fetch_out.release();
        h_in.release(); h_counts.release(); h_stats.release();
        work.release();
        if (h_status) cudaFreeHost(h_status);
        if (done) cudaEventDestroy(done);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct vfk_handle {
    int device = 0;
    int sm_count = 0;
    Work work;  // the device-pointer API (vfk_pileup, vfk_pileup_values)
    int64_t launches = 0;
    Slot slot[VFK_SLOTS];
    int next_slot = 0;
    int64_t staged_bytes = 0, d2h_bytes = 0, zero_copy_batches = 0, staged_batches = 0;
    bool timing = false;     // vfk_set_timing
    Work *timed = nullptr;   // the working set of the last timed call
    vfk::LaunchCache launch_cache[4] = {};  // per position-range instantiation of the pileup kernels (vfk_pileup.cu)
    bool link_all = false;   // debug hook (VFK_DEBUG_LINK_ALL=1 when the handle is created): partners from the link kernels everywhere
    bool stage_all = false;  // VFK_HOST_STAGE_ALL=1 when the handle is created: the host entry point copies every array
    bool fetch = true;       // in-place batches get their base / quality bytes through a fetch list (VFK_HOST_FETCH=0 when the handle is created: off)
    double fetch_slack = 1.15;  // list capacity = units x (sampled depth x slack + 6); VFK_HOST_FETCH_SLACK at creation (tests: a short list)
    int fetch_threads = 1;   // host threads of the gather (VFK_HOST_THREADS at creation, else up to 16 of the cores)
    int64_t fetch_pairs = 0, fetch_batches = 0, fetch_nanos = 0;
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
extern "C" int vfk_slot_count(void) { return VFK_SLOTS; }
extern "C" const char *vfk_last_error(void) { return g_err; }

static bool env_on(const char *name) {
    const char *v = getenv(name);
    return v && v[0] == '1';
}

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
    h->link_all = env_on("VFK_DEBUG_LINK_ALL");
    h->stage_all = env_on("VFK_HOST_STAGE_ALL");
    {
        const char *v;
        if ((v = getenv("VFK_HOST_FETCH_SLACK")) && atof(v) > 0.0) h->fetch_slack = atof(v);
        int t = omp_get_max_threads();
        if ((v = getenv("VFK_HOST_THREADS")) && atoi(v) > 0) t = atoi(v);
        else if (t > 16) t = 16;
        h->fetch_threads = t < 1 ? 1 : t;
        // A pair costs a gathering thread ~32 ns and saves the link two requests (~7 ns): with fewer than six threads for this
        // GPU the gather is longer than the link time it saves (measured: eight ranks on a 32-core host, four threads each: 19.0 ms
        // per C3 step with the lists, 14.1 ms in place; 16 threads, one GPU: 34.8 against 52.0 ms).  VFK_HOST_FETCH=1 / 0 decides outright.
        v = getenv("VFK_HOST_FETCH");
        h->fetch = v && v[0] == '1' ? true : v && v[0] == '0' ? false : h->fetch_threads >= 6;
    }
    {
        // The path gathers single 32-byte sectors (one quality byte, one sequence byte per (read, column) pair):
        // ask L2 not to promote misses to 64/128-byte DRAM fetches.  A hint; failure is harmless.
        (void)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
        (void)cudaGetLastError();
    }
    for (Slot &s : h->slot) {
        CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        CU(cudaMalloc(&s.work.counters, COUNTER_BYTES));
        CU(cudaMallocHost((void **)&s.h_status, 2 * sizeof(uint32_t)));
        s.h_status[0] = s.h_status[1] = 0u;
    }
    CU(cudaMalloc(&h->work.counters, COUNTER_BYTES));
    *out = h;
    return VFK_OK;
}

extern "C" int vfk_destroy(vfk_handle *h) {
    if (!h) return VFK_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (Slot &s : h->slot) s.release();
    h->work.release();
    h->carter.release();
    if (h->carter_ready) cudaEventDestroy(h->carter_ready);
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
    if (r->n_cigar_words < 0 || r->n_seq_bytes < 0 || r->n_qual_bytes < 0) return fail(VFK_EINVAL, "negative pool size");
    if (r->n_reads > 0x7FFFFFFF) return fail(VFK_EUNSUPPORTED, "more than 2^31-1 reads in one batch: split it");
    if (v->n_units > 0x7FFFFFFF) return fail(VFK_EUNSUPPORTED, "more than 2^31-1 units in one batch: split it");
    if (r->n_cigar_words >= ((int64_t)1 << 32)) return fail(VFK_EUNSUPPORTED, "CIGAR pool exceeds 2^32 words: split the batch");
    if (r->max_l_qseq > VFK_MAX_READ_LEN) return fail(VFK_EUNSUPPORTED, "read longer than VFK_MAX_READ_LEN");
    if (!r->contig_off) return fail(VFK_EINVAL, "contig_off is NULL");
    if (r->n_reads > 0 && (!r->pos || !r->flag || !r->mapq || !r->l_qseq || !r->n_cigar || !r->cigar || !r->seq || !r->qual || !r->mtid ||
                           !r->mpos || !r->isize || !r->qname_id || !r->qname_hash))
        return fail(VFK_EINVAL, "read batch has a NULL array");
    // the three pool offset arrays: all given, or all NULL (pools contiguous in read order: the offsets are derived on the device)
    if (r->n_reads > 0 && ((!r->cigar_off) != (!r->seq_off) || (!r->cigar_off) != (!r->qual_off)))
        return fail(VFK_EINVAL, "cigar_off / seq_off / qual_off must be all given or all NULL");
    if (v->n_units > 0 && (!v->tid || !v->pos || !v->meta || !v->length || !v->seq_off || !v->ins_seq))
        return fail(VFK_EINVAL, "variant batch has a NULL array");
    return VFK_OK;
}

static void to_dev(const vfk_read_batch *r, const vfk_variant_batch *v, const Work &w, vfk::ReadsDev &R, vfk::UnitsDev &V) {
    R.pos = r->pos; R.flag = r->flag; R.mapq = r->mapq; R.l_qseq = r->l_qseq; R.n_cigar = r->n_cigar;
    R.cigar_off = r->cigar_off; R.seq_off = r->seq_off; R.qual_off = r->qual_off;
    R.cigar = r->cigar; R.seq = r->seq; R.qual = r->qual;
    R.mtid = r->mtid; R.mpos = r->mpos; R.isize = r->isize; R.qname_id = (const uint64_t *)r->qname_id; R.qname_hash = r->qname_hash;
    R.contig_off = r->contig_off;
    R.n_reads = r->n_reads;
    R.n_cigar_words = r->n_cigar_words; R.n_seq_bytes = r->n_seq_bytes; R.n_qual_bytes = r->n_qual_bytes;
    R.n_contigs = r->n_contigs;
    R.ee = (const uint2 *)w.ee.p;
    R.chunk_max = (const int64_t *)w.chunk_max.p;
    R.blk_incl = (const int64_t *)w.blk_incl.p;
    R.mate = (const uint32_t *)w.mate.p;
    V.n = v->n_units;
    V.tid = v->tid; V.pos = v->pos; V.meta = v->meta; V.length = v->length; V.seq_off = v->seq_off; V.ins_seq = v->ins_seq; V.stop = v->stop;
}

// read preparation on `st`: counters zeroed, spans / shapes / position index derived (R's canonical pointers are device visible)
static int enqueue_prep(vfk_handle *h, Work &w, vfk::ReadsDev &R, cudaStream_t st, int *launches) {
    if (!R.cigar_off && R.n_reads > 0) {  // implicit pool offsets (check_batches: then all three are NULL)
        const size_t nr = (size_t)R.n_reads;
        int rc;
        if ((rc = w.off_c.reserve(nr * 8)) || (rc = w.off_s.reserve(nr * 8)) || (rc = w.off_q.reserve(nr * 8)) ||
            (rc = w.off_sums.reserve((nr / 1024 + 1) * 24)))
            return rc;
        cudaError_t e = vfk::launch_offsets(R, (long long *)w.off_sums.p, (int64_t *)w.off_c.p, (int64_t *)w.off_s.p, (int64_t *)w.off_q.p, st, launches);
        if (e != cudaSuccess) return fail(VFK_ECUDA, "offsets launch: %s", cudaGetErrorString(e));
        R.cigar_off = (const int64_t *)w.off_c.p; R.seq_off = (const int64_t *)w.off_s.p; R.qual_off = (const int64_t *)w.off_q.p;
    }
    if (h->timing) {
        for (cudaEvent_t &e : w.tev)
            if (!e) CU(cudaEventCreate(&e));
        w.ev_epilogue = false;
        h->timed = &w;
        CU(cudaEventRecord(w.tev[0], st));
    }
    CU(cudaMemsetAsync(w.counters, 0, COUNTER_BYTES, st));
    cudaError_t e = vfk::launch_prep(R, (uint2 *)w.ee.p, (int64_t *)w.chunk_max.p, (int64_t *)w.blk_max.p, (int64_t *)w.blk_incl.p,
                                     (uint32_t *)((char *)w.counters + 20), st, launches);
    if (e != cudaSuccess) return fail(VFK_ECUDA, "prep launch: %s", cudaGetErrorString(e));
    if (h->timing) CU(cudaEventRecord(w.tev[1], st));
    return VFK_OK;
}

static vfk::PileupScratch scratch_of(vfk_handle *h, Work &w) {
    vfk::PileupScratch s = w.scratch();
    if (h->timing) s.ev = w.tev;
    s.cache = h->launch_cache;
    return s;
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
    rc = h->work.reserve(reads->n_reads, units->n_units, st);
    if (rc) return rc;
    vfk::ReadsDev R;
    vfk::UnitsDev V;
    to_dev(reads, units, h->work, R, V);
    int launches = 0;
    rc = enqueue_prep(h, h->work, R, st, &launches);
    if (rc) return rc;
    cudaError_t e = vfk::launch_pileup(R, V, *params, reads->max_l_qseq, out, scratch_of(h, h->work), h->sm_count, st, &launches, h->link_all);
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
    cudaStream_t st = (cudaStream_t)stream;
    rc = h->work.reserve(reads->n_reads, units->n_units, st);
    if (rc) return rc;
    vfk::ReadsDev R;
    vfk::UnitsDev V;
    to_dev(reads, units, h->work, R, V);
    int launches = 0;
    rc = enqueue_prep(h, h->work, R, st, &launches);
    if (rc) return rc;
    const vfk::PileupScratch S = h->work.scratch();
    cudaError_t e = vfk::launch_locate_only(R, V, S.resolved, st, &launches);
    if (e == cudaSuccess)
        e = vfk::launch_link(R, *params, S.resolved, nullptr, nullptr, V.n, S.heads, S.n_buckets, S.next, S.column, S.claimed, S.mate,
                             h->sm_count, st, &launches);
    if (e == cudaSuccess) e = vfk::launch_values(R, S.resolved, V.n, V.ins_seq, *params, max_per_unit, values, n_values, st, &launches);
    h->launches += launches;
    if (e != cudaSuccess) return fail(VFK_ECUDA, "values launch: %s", cudaGetErrorString(e));
    return VFK_OK;
}

extern "C" int vfk_prep_words(vfk_handle *h, const vfk_read_batch *reads, int32_t *end_out, uint32_t *shape_out, uint32_t *status_out,
                              void *stream) {
    if (!h) return fail(VFK_EINVAL, "vfk_prep_words: NULL handle");
    vfk_variant_batch none;
    memset(&none, 0, sizeof(none));
    vfk_params prm;
    memset(&prm, 0, sizeof(prm));
    int rc = check_batches(reads, &none, &prm);
    if (rc) return rc;
    if (reads->n_reads == 0) return VFK_OK;
    if (!end_out || !shape_out || !status_out) return fail(VFK_EINVAL, "vfk_prep_words: NULL output");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    rc = h->work.reserve(reads->n_reads, 1, st);
    if (rc) return rc;
    vfk::ReadsDev R;
    vfk::UnitsDev V;
    to_dev(reads, &none, h->work, R, V);
    int launches = 0;
    rc = enqueue_prep(h, h->work, R, st, &launches);
    h->launches += launches;
    if (rc) return rc;
    // the kernels keep {end, shape} as one 8-byte pair per read: two strided copies split it
    CU(cudaMemcpy2DAsync(end_out, 4, h->work.ee.p, 8, 4, (size_t)reads->n_reads, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpy2DAsync(shape_out, 4, (const char *)h->work.ee.p + 4, 8, 4, (size_t)reads->n_reads, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(status_out, (char *)h->work.counters + 20, 4, cudaMemcpyDeviceToDevice, st));
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

// device-visible alias of a page-locked host buffer, or NULL
static const void *pinned_alias(const void *host) {
    cudaPointerAttributes attr;
    const void *dev = nullptr;
    if (host && cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost) dev = attr.devicePointer;
    (void)cudaGetLastError();
    return dev;
}

extern "C" int vfk_wait(vfk_handle *h, int ticket) {
    if (!h || ticket < 0 || ticket >= VFK_SLOTS) return fail(VFK_EINVAL, "vfk_wait: bad handle or ticket");
    Slot &s = h->slot[ticket];
    if (!s.busy) return VFK_OK;
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(s.done));
    s.busy = false;
    if (s.user_counts) memcpy(s.user_counts, s.h_counts.p, (size_t)s.user_n * sizeof(vfk_counts));
    if (s.user_stats) memcpy(s.user_stats, s.h_stats.p, (size_t)s.user_n * sizeof(vfk_stats));
    s.user_counts = nullptr;
    s.user_stats = nullptr;
    if (*s.h_status & VFK_ST_BADBATCH)
        return fail(VFK_EINVAL, "malformed read batch (not coordinate sorted, negative start, CIGAR outside its pool or read too long)");
    return VFK_OK;
}

// Reads starting inside (col - max_l_qseq, col], averaged over 64 evenly spaced units: the depth of the variant columns
// as far as two binary searches per sampled unit can tell (HOST arrays).
static double sampled_depth(const vfk_read_batch *r, const vfk_variant_batch *v) {
    const int64_t nu = v->n_units;
    const int32_t w = r->max_l_qseq > 0 ? r->max_l_qseq : 1;
    double sum = 0.0;
    int n = 0;
    for (int s = 0; s < 64 && s < nu; ++s) {
        const int64_t u = nu <= 64 ? s : (int64_t)((2 * s + 1) * (double)nu / 128.0);
        const int32_t tid = v->tid[u], col = v->pos[u];
        if (tid < 0 || tid >= r->n_contigs) continue;
        const int32_t *b = r->pos + r->contig_off[tid], *e = r->pos + r->contig_off[tid + 1];
        const int32_t *hi = b, *lo = b;
        for (int64_t len = e - b; len > 0;) {  // first start beyond col
            const int64_t half = len >> 1;
            if (hi[half] <= col) { hi += half + 1; len -= half + 1; } else len = half;
        }
        for (int64_t len = e - b; len > 0;) {  // first start beyond col - w
            const int64_t half = len >> 1;
            if ((int64_t)lo[half] <= (int64_t)col - w) { lo += half + 1; len -= half + 1; } else len = half;
        }
        sum += (double)(hi - lo);
        ++n;
    }
    return n ? sum / n : 0.0;
}

// fetch list, between request_kernel and the register kernel (stream order): how many pairs, the host's gather, the bytes back
static cudaError_t fetch_between(void *ctx, const uint32_t *pair_counter, cudaStream_t st) {
    Slot &S = *(Slot *)ctx;
    cudaError_t e = cudaMemcpyAsync(S.h_status + 1, pair_counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaLaunchHostFunc(st, fetch_cb, &S.job);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(S.fetched.p, S.fetch_out.p, (size_t)S.job.cap * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
}

// copy `bytes` of a host array into the slot's staging buffer; *dst = the device copy
#define STAGE(buf, src, bytes, dst)                                                              \
    do {                                                                                         \
        size_t b_ = (size_t)(bytes);                                                             \
        rc = (buf).reserve(b_ ? b_ : 16);                                                        \
        if (rc) goto failed;                                                                     \
        if (b_ && cudaMemcpyAsync((buf).p, (src), b_, cudaMemcpyHostToDevice, st) != cudaSuccess) { \
            rc = fail(VFK_ECUDA, "cudaMemcpyAsync(" #src ") failed");                            \
            goto failed;                                                                         \
        }                                                                                        \
        h->staged_bytes += (int64_t)b_;                                                          \
        (dst) = (decltype(dst))(buf).p;                                                          \
    } while (0)

// the same for a small host array that may be pageable: through the slot's page-locked bounce arena (at *used), so that the copy
// is asynchronous whatever memory the caller handed over
#define STAGE_SMALL(buf, src, bytes, dst)                                                        \
    do {                                                                                         \
        size_t n_ = (size_t)(bytes);                                                             \
        const void *from_ = (src);                                                               \
        if (n_ && !pinned_alias(from_)) {                                                        \
            memcpy((char *)S.h_in.p + arena_used, from_, n_);                                    \
            from_ = (char *)S.h_in.p + arena_used;                                               \
            arena_used += (n_ + 63) & ~(size_t)63;                                               \
        }                                                                                        \
        STAGE(buf, from_, n_, dst);                                                              \
    } while (0)

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
    rc = vfk_wait(h, si);  // the slot's previous batch must have left it (its status belongs to that batch's own wait)
    (void)rc;
    CU(cudaSetDevice(h->device));
    Slot &S = h->slot[si];
    cudaStream_t st = S.stream;
    vfk_read_batch dr = *reads;
    vfk_variant_batch dv = *units;
    int launches = 0;
    size_t arena_used = 0;
    cudaError_t e = cudaSuccess;
    // ---- how the read arrays reach the kernels.  Page-locked arrays: only what the position index is built from (starts,
    // CIGARs, read lengths: ~25 bytes per read) is copied; flags, qualities, bases, mate fields and name ids are read in place
    // over PCIe, and only for the reads that overlap a variant column.  Pageable arrays: everything is staged.
    const bool implicit = !reads->cigar_off;  // pool offsets derived on the device: nothing of them crosses the bus
    const void *z[11] = {pinned_alias(reads->flag), pinned_alias(reads->mapq), pinned_alias(reads->seq_off), pinned_alias(reads->qual_off),
                         pinned_alias(reads->seq), pinned_alias(reads->qual), pinned_alias(reads->mtid), pinned_alias(reads->mpos),
                         pinned_alias(reads->isize), pinned_alias(reads->qname_id), pinned_alias(reads->qname_hash)};
    bool zero_copy = !h->stage_all && nr > 0;
    double depth = 0.0;
    vfk::PileupScratch scratch;
    for (int k = 0; k < 11; ++k) zero_copy = zero_copy && (z[k] != nullptr || (implicit && (k == 2 || k == 3)));
    if (zero_copy) {
        // In place pays per (unit, candidate read) pair -- about 256 bytes cross the bus for each (a dozen arrays, 64-byte requests) --
        // a bulk copy pays per read.  Sparse variant sets against 30x / 60x reads touch a fifth of the batch: in place.  A deep
        // panel touches every read many times over: staged.  The depth is estimated from a sample of the units.
        depth = sampled_depth(reads, units);
        const double touches = (double)nu * depth;
        const double per_read = 43.0 + (double)(reads->n_seq_bytes + reads->n_qual_bytes) / (double)nr;
        zero_copy = touches * 256.0 < (double)nr * per_read;
    }
    rc = S.work.reserve(nr, nu, st);
    if (rc) goto failed;
    STAGE(S.contig, reads->contig_off, (size_t)(reads->n_contigs + 1) * sizeof(int64_t), dr.contig_off);  // (a few hundred bytes: synchronous at worst on an idle stream)
    STAGE(S.pos, reads->pos, (size_t)nr * 4, dr.pos);
    STAGE(S.n_cigar, reads->n_cigar, (size_t)nr * 4, dr.n_cigar);
    if (!implicit) STAGE(S.cigar_off, reads->cigar_off, (size_t)nr * 8, dr.cigar_off);
    STAGE(S.l_qseq, reads->l_qseq, (size_t)nr * 4, dr.l_qseq);
    STAGE(S.cigar, reads->cigar, (size_t)reads->n_cigar_words * 4, dr.cigar);
    if (zero_copy) {
        dr.flag = (const uint16_t *)z[0]; dr.mapq = (const uint8_t *)z[1]; dr.seq_off = (const int64_t *)z[2]; dr.qual_off = (const int64_t *)z[3];
        dr.seq = (const uint8_t *)z[4]; dr.qual = (const uint8_t *)z[5]; dr.mtid = (const int32_t *)z[6]; dr.mpos = (const int32_t *)z[7];
        dr.isize = (const int32_t *)z[8]; dr.qname_id = (const uint64_t *)z[9]; dr.qname_hash = (const uint32_t *)z[10];
        ++h->zero_copy_batches;
    } else {
        STAGE(S.flag, reads->flag, (size_t)nr * 2, dr.flag);
        STAGE(S.mapq, reads->mapq, (size_t)nr, dr.mapq);
        if (!implicit) {
            STAGE(S.seq_off, reads->seq_off, (size_t)nr * 8, dr.seq_off);
            STAGE(S.qual_off, reads->qual_off, (size_t)nr * 8, dr.qual_off);
        }
        STAGE(S.seq, reads->seq, (size_t)reads->n_seq_bytes, dr.seq);
        STAGE(S.qual, reads->qual, (size_t)reads->n_qual_bytes, dr.qual);
        STAGE(S.mtid, reads->mtid, (size_t)nr * 4, dr.mtid);
        STAGE(S.mpos, reads->mpos, (size_t)nr * 4, dr.mpos);
        STAGE(S.isize, reads->isize, (size_t)nr * 4, dr.isize);
        STAGE(S.qname_id, reads->qname_id, (size_t)nr * 8, dr.qname_id);
        STAGE(S.qname_hash, reads->qname_hash, (size_t)nr * 4, dr.qname_hash);
        ++h->staged_batches;
    }
    rc = S.h_in.reserve((size_t)nu * (6 * 4 + 8) + (size_t)(units->n_ins_seq > 0 ? units->n_ins_seq : 1) + (size_t)(reads->n_contigs + 1) * 8 + 16 * 64);
    if (rc) goto failed;
    STAGE_SMALL(S.tid, units->tid, (size_t)nu * 4, dv.tid);
    STAGE_SMALL(S.upos, units->pos, (size_t)nu * 4, dv.pos);
    STAGE_SMALL(S.meta, units->meta, (size_t)nu * 4, dv.meta);
    STAGE_SMALL(S.len, units->length, (size_t)nu * 4, dv.length);
    STAGE_SMALL(S.soff, units->seq_off, (size_t)nu * 4, dv.seq_off);
    STAGE_SMALL(S.ins, units->ins_seq, (size_t)(units->n_ins_seq > 0 ? units->n_ins_seq : 1), dv.ins_seq);
    if (units->stop) STAGE_SMALL(S.stop, units->stop, (size_t)nu * 4, dv.stop);
    {
        const double *d_eaf = nullptr;
        STAGE_SMALL(S.eaf, eaf, (size_t)nu * sizeof(double), d_eaf);
        rc = S.counts.reserve((size_t)nu * sizeof(vfk_counts));
        if (rc) goto failed;
        rc = S.stats.reserve((size_t)nu * sizeof(vfk_stats));
        if (rc) goto failed;
        vfk::ReadsDev R;
        vfk::UnitsDev V;
        to_dev(&dr, &dv, S.work, R, V);
        rc = enqueue_prep(h, S.work, R, st, &launches);
        if (rc) goto failed;
        scratch = scratch_of(h, S.work);
        if (zero_copy && h->fetch && !h->link_all && nu < ((int64_t)1 << 24)) {
            // the reads' base / quality bytes at the columns: asked for by the device, gathered by the host, copied in bulk
            double want = (double)nu * (depth * h->fetch_slack + 6.0) + 1024.0;
            if (want > (double)nu * 192.0) want = (double)nu * 192.0;  // (the deepest column that asks: the medium kernel's)
            if (want > 4.0e9) want = 4.0e9;
            const uint32_t cap = (uint32_t)want;
            if ((rc = S.fetch_req.reserve((size_t)cap * 16)) || (rc = S.fetch_out.reserve((size_t)cap * 2)) || (rc = S.fetched.reserve((size_t)cap * 2)) ||
                (rc = S.fetch_base.reserve((size_t)nu * 4)))
                goto failed;
            FetchJob &J = S.job;
            J.req = (const ulonglong2 *)S.fetch_req.p; J.out = (uint16_t *)S.fetch_out.p; J.h_total = S.h_status + 1; J.cap = cap;
            J.qual = reads->qual; J.seq = reads->seq;
            J.qual_off = implicit ? nullptr : reads->qual_off; J.seq_off = implicit ? nullptr : reads->seq_off;
            J.n_reads = nr; J.n_qual = reads->n_qual_bytes; J.n_seq = reads->n_seq_bytes;
            J.threads = h->fetch_threads; J.pairs = &h->fetch_pairs; J.nanos = &h->fetch_nanos;
            scratch.fetch_cap = cap; scratch.fetch_by_index = implicit ? 0 : 1;
            scratch.fetch_base = (uint32_t *)S.fetch_base.p; scratch.fetch_req = (void *)pinned_alias(S.fetch_req.p);
            scratch.fetched = (const uint16_t *)S.fetched.p; scratch.fetch_between = fetch_between; scratch.fetch_ctx = &S;
            if (!scratch.fetch_req) scratch.fetch_cap = 0u;
            else { ++h->fetch_batches; h->staged_bytes += (int64_t)cap * 2; }
        }
        e = vfk::launch_pileup(R, V, *params, reads->max_l_qseq, (vfk_counts *)S.counts.p, scratch, h->sm_count, st, &launches,
                               h->link_all);
        if (e != cudaSuccess) { rc = fail(VFK_ECUDA, "pileup launch: %s", cudaGetErrorString(e)); goto failed; }
        rc = ensure_carter(h, fpr, error_rate, st);
        if (rc) goto failed;
        e = vfk::launch_epilogue(nu, (const vfk_counts *)S.counts.p, d_eaf, fpr, error_rate, h->carter.p, CARTER_DEPTHS, (vfk_stats *)S.stats.p,
                                 st, &launches);
        if (e != cudaSuccess) { rc = fail(VFK_ECUDA, "epilogue launch: %s", cudaGetErrorString(e)); goto failed; }
        if (h->timing) {
            cudaEventRecord(S.work.tev[5], st);
            S.work.ev_epilogue = true;
        }
    }
    h->launches += launches;
    {
        void *c_to = counts_out, *s_to = stats_out;
        if (!pinned_alias(counts_out)) {
            rc = S.h_counts.reserve((size_t)nu * sizeof(vfk_counts));
            if (rc) goto failed;
            c_to = S.h_counts.p;
        }
        if (!pinned_alias(stats_out)) {
            rc = S.h_stats.reserve((size_t)nu * sizeof(vfk_stats));
            if (rc) goto failed;
            s_to = S.h_stats.p;
        }
        if (cudaMemcpyAsync(c_to, S.counts.p, (size_t)nu * sizeof(vfk_counts), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(s_to, S.stats.p, (size_t)nu * sizeof(vfk_stats), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(S.h_status, (char *)S.work.counters + 20, sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) {
            rc = fail(VFK_ECUDA, "cudaMemcpyAsync(results) failed");
            goto failed;
        }
        S.user_counts = c_to == (void *)counts_out ? nullptr : counts_out;
        S.user_stats = s_to == (void *)stats_out ? nullptr : stats_out;
        S.user_n = nu;
    }
    h->d2h_bytes += (int64_t)nu * (int64_t)(sizeof(vfk_counts) + sizeof(vfk_stats)) + 4;
    CU(cudaEventRecord(S.done, st));
    S.busy = true;
    h->next_slot = (h->next_slot + 1) % VFK_SLOTS;
    return VFK_OK;
failed:
    // copies and kernels already enqueued may still read the caller's buffers: they must have finished before the
    // caller gets the error (and may free them); the slot stays free
    h->launches += launches;
    cudaStreamSynchronize(st);
    (void)cudaGetLastError();
    return rc;
}

extern "C" int vfk_annotate_host(vfk_handle *h, const vfk_read_batch *reads, const vfk_variant_batch *units,
                                 const vfk_params *params, const double *eaf, double fpr, double error_rate,
                                 vfk_counts *counts_out, vfk_stats *stats_out) {
    int ticket = 0;
    int rc = vfk_annotate_host_async(h, reads, units, params, eaf, fpr, error_rate, counts_out, stats_out, &ticket);
    if (rc) return rc;
    return vfk_wait(h, ticket);
}

extern "C" int vfk_set_timing(vfk_handle *h, int on) {
    if (!h) return fail(VFK_EINVAL, "vfk_set_timing: NULL handle");
    h->timing = on != 0;
    if (!h->timing) h->timed = nullptr;
    return VFK_OK;
}

extern "C" int vfk_last_timing(vfk_handle *h, double ms[5]) {
    if (!h || !ms) return fail(VFK_EINVAL, "vfk_last_timing: NULL argument");
    if (!h->timed) return fail(VFK_EINVAL, "vfk_last_timing: no timed call (vfk_set_timing first)");
    Work &w = *h->timed;
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(w.tev[w.ev_epilogue ? 5 : 4]));
    for (int k = 0; k < 5; ++k) {
        float t = 0.f;
        ms[k] = 0.0;
        if (k == 4 && !w.ev_epilogue) break;
        CU(cudaEventElapsedTime(&t, w.tev[k], w.tev[k + 1]));
        ms[k] = (double)t;
    }
    return VFK_OK;
}

extern "C" int vfk_fetch_gather(const uint64_t *req, int64_t n_pairs, const uint8_t *qual, int64_t n_qual_bytes, const uint8_t *seq,
                                int64_t n_seq_bytes, const int64_t *qual_off, const int64_t *seq_off, int64_t n_reads, int threads, uint16_t *out) {
    if (n_pairs < 0 || n_pairs > 0xFFFFFFFFll) return fail(VFK_EINVAL, "vfk_fetch_gather: bad pair count");
    if (n_pairs == 0) return VFK_OK;
    if (!req || !qual || !seq || !out || (!qual_off) != (!seq_off)) return fail(VFK_EINVAL, "vfk_fetch_gather: NULL array");
    const uint32_t total = (uint32_t)n_pairs;
    FetchJob j;
    j.req = (const ulonglong2 *)req; j.out = out; j.h_total = &total; j.cap = total;
    j.qual = qual; j.seq = seq; j.qual_off = qual_off; j.seq_off = seq_off;
    j.n_reads = n_reads; j.n_qual = n_qual_bytes; j.n_seq = n_seq_bytes;
    j.threads = threads > 0 ? threads : 1;
    fetch_cb(&j);
    return VFK_OK;
}

extern "C" int vfk_fetch_stats(vfk_handle *h, int64_t out[4]) {
    if (!h || !out) return fail(VFK_EINVAL, "vfk_fetch_stats: NULL argument");
    out[0] = __atomic_load_n(&h->fetch_pairs, __ATOMIC_RELAXED);
    out[1] = h->fetch_batches;
    out[2] = __atomic_load_n(&h->fetch_nanos, __ATOMIC_RELAXED);
    out[3] = h->fetch_threads;
    return VFK_OK;
}

extern "C" int vfk_transfer_stats(vfk_handle *h, int64_t out[4]) {
    if (!h || !out) return fail(VFK_EINVAL, "vfk_transfer_stats: NULL argument");
    out[0] = h->staged_bytes;
    out[1] = h->zero_copy_batches;
    out[2] = h->staged_batches;
    out[3] = h->d2h_bytes;
    return VFK_OK;
}

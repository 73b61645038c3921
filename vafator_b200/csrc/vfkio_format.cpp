// vfkio_format.cpp -- the INFO text of one single-BAM sample for a run of records, from the per-record / per-row arrays the
// host composition holds (vafator_b200/engine.py SingleBamColumns).  What the reference builds field by field with Python string
// formatting (vafator/annotator.py:304-420): "{sample}_af=..;{sample}_ac=..;..{sample}_pos=..[;{sample}_rsmq=..;{sample}_rsmq_pv=..]..".
// No statistic is computed here: every number arrives as the device produced it (and numpy rounded it); this file only prints,
// with Python's own float formatting -- repr(float): the shortest digit string that round-trips, fixed notation while
// -4 < decimal exponent <= 16, else d.ddde+XX -- and Python's round(x, 5) for the allele frequency (correctly rounded on the
// exact binary value, as glibc's printf("%.5f") rounds).  tests/test_host_logic.py compares it with the Python formatter string
// for string; vfkio_py_repr is exported for a differential test against repr() itself.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <omp.h>
#include <string>
#include <vector>
#include "vfk_io.h"

int vfkio_fail(int code, const char *msg);

namespace {

// repr(float) into `out` (<= 32 bytes); returns the length
int py_repr(double x, char *out) {
    if (std::isnan(x)) { memcpy(out, "nan", 3); return 3; }
    if (std::isinf(x)) { const char *s = x < 0 ? "-inf" : "inf"; const int n = (int)strlen(s); memcpy(out, s, (size_t)n); return n; }
    char *o = out;
    if (std::signbit(x)) { *o++ = '-'; x = -x; }
    if (x == 0.0) { memcpy(o, "0.0", 3); return (int)(o - out) + 3; }
    char sci[40];
    const auto res = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific);  // shortest round-trip digits: d[.ddd]e[+-]XX
    *res.ptr = '\0';  // (to_chars does not terminate; at most 24 of the 40 bytes are used)
    char digits[24];
    int nd = 0, exp10 = 0;
    const char *p = sci;
    for (; p < res.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    if (p < res.ptr) exp10 = atoi(p + 1);
    const int decpt = exp10 + 1;  // value = 0.d1d2.. x 10^decpt
    if (decpt <= -4 || decpt > 16) {
        *o++ = digits[0];
        if (nd > 1) { *o++ = '.'; memcpy(o, digits + 1, (size_t)(nd - 1)); o += nd - 1; }
        *o++ = 'e';
        int e = decpt - 1;
        *o++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        o += snprintf(o, 8, "%02d", e);
    } else if (decpt <= 0) {
        *o++ = '0'; *o++ = '.';
        for (int k = 0; k < -decpt; ++k) *o++ = '0';
        memcpy(o, digits, (size_t)nd); o += nd;
    } else if (decpt >= nd) {
        memcpy(o, digits, (size_t)nd); o += nd;
        for (int k = 0; k < decpt - nd; ++k) *o++ = '0';
        *o++ = '.'; *o++ = '0';
    } else {
        memcpy(o, digits, (size_t)decpt); o += decpt;
        *o++ = '.';
        memcpy(o, digits + decpt, (size_t)(nd - decpt)); o += nd - decpt;
    }
    return (int)(o - out);
}

inline void put_float(std::string &s, double x) {
    char b[40];
    s.append(b, (size_t)py_repr(x, b));
}
inline void put_int(std::string &s, long long v) {
    char b[24];
    const auto r = std::to_chars(b, b + sizeof(b), v);
    s.append(b, (size_t)(r.ptr - b));
}
// Python's round(x, 5): correctly rounded to five decimals on the exact value, then the nearest double
inline double py_round5(double x) {
    if (!std::isfinite(x)) return x;
    char b[64];
    snprintf(b, sizeof(b), "%.5f", x);
    return strtod(b, nullptr);
}

constexpr int KIND_SNV = 0;
// "0" = the reference's Counter misses the allele, "nan" = key present with an empty list, else the median
inline void put_median(std::string &s, int code, long long med2) {
    if (code == 2) put_float(s, (double)med2 / 2.0);
    else if (code == 1) s.append("nan");
    else s.push_back('0');
}

}  // namespace

extern "C" int vfkio_py_repr(double x, char *out32) { return py_repr(x, out32); }

extern "C" void vfkio_free(void *p) { free(p); }

extern "C" int vfkio_format_info(const vfkio_format_in *in, char **text_out, int64_t *offsets_out) {
    if (!in || !text_out || !offsets_out) return vfkio_fail(VFK_EINVAL, "vfkio_format_info: NULL argument");
    const int64_t n_rec = in->n_rec;
    if (n_rec < 0) return vfkio_fail(VFK_EINVAL, "vfkio_format_info: negative size");
    *text_out = nullptr;
    offsets_out[0] = 0;
    if (n_rec == 0) return VFK_OK;
    const std::string sample(in->sample ? in->sample : "");
    static const char *const FIXED[11] = {"af", "ac", "n", "dp", "eaf", "pu", "pw", "k", "bq", "mq", "pos"};
    static const char *const RS[3] = {"rsmq", "rsbq", "rspos"};  // emission order (vafator/annotator.py:350-376)
    static const int RS_METRIC[3] = {0, 1, 2};                   // index into the device record's (mq, bq, pos)
    std::string key[11], rs_key[3], rs_pv[3];
    for (int k = 0; k < 11; ++k) key[k] = (k ? ";" : "") + sample + "_" + FIXED[k] + "=";
    for (int m = 0; m < 3; ++m) { rs_key[m] = ";" + sample + "_" + RS[m] + "="; rs_pv[m] = ";" + sample + "_" + RS[m] + "_pv="; }
    const int64_t *first = in->first;
    const int nt = std::max(1, std::min(omp_get_max_threads(), (int)std::min<int64_t>(64, (n_rec + 4095) / 4096)));
    std::vector<std::string> part((size_t)nt);
    std::vector<int64_t> len((size_t)n_rec);
#pragma omp parallel num_threads(nt)
    {
        const int t = omp_get_thread_num();
        const int64_t i0 = n_rec * t / nt, i1 = n_rec * (t + 1) / nt;
        std::string &s = part[(size_t)t];
        s.reserve((size_t)(i1 - i0) * 260);
        // medians per metric: 0 = mq, 1 = bq, 2 = pos in the device record; printed bq, mq, pos
        static const int PRINT_ORDER[3] = {1, 0, 2};
        for (int64_t i = i0; i < i1; ++i) {
            const size_t at = s.size();
            const int64_t r0 = first[i], r1 = first[i + 1];
            const bool has = in->has_rec[i] != 0, snv = in->kind_rec[i] == KIND_SNV, indel = has && in->kind_rec[i] > KIND_SNV;
            const long long dp = in->dp[i], ac_ref = in->ac_ref[i];
            const bool ref_ok = has && ac_ref > 0;
            s += key[0];
            for (int64_t r = r0; r < r1; ++r) {
                if (r > r0) s.push_back(',');
                if (dp > 0) put_float(s, py_round5((double)in->ac[r] / (double)dp)); else s.append("0.0");
            }
            s += key[1];
            for (int64_t r = r0; r < r1; ++r) { if (r > r0) s.push_back(','); put_int(s, in->ac[r]); }
            s += key[2]; put_int(s, in->n[i]);
            s += key[3]; put_int(s, dp);
            s += key[4]; put_float(s, in->eaf[i]);
            s += key[5];
            for (int64_t r = r0; r < r1; ++r) { if (r > r0) s.push_back(','); put_float(s, in->pu[r]); }
            s += key[6]; if (r1 > r0) put_float(s, in->pw[r0]); else s.append("0.0");
            s += key[7]; if (r1 > r0) put_int(s, in->k[r0]); else s.push_back('1');
            for (int q = 0; q < 3; ++q) {
                const int m = PRINT_ORDER[q];
                s += key[8 + q];
                int code = ref_ok ? 2 : 0;                       // SNV: value or "0"
                if (indel) code = ac_ref > 0 ? 2 : 1;            // indel: the REF key always exists
                if (m == 1 && !snv) code = 0;                    // indel metrics carry no base qualities
                put_median(s, code, in->med2_ref[3 * i + m]);
                for (int64_t r = r0; r < r1; ++r) {
                    const bool alt_ok = has && in->ac[r] > 0;
                    int c = alt_ok ? 2 : 0;
                    // an indel ALT list is keyed ALT.upper() and looked up with the literal text
                    if (indel) c = in->present_row[r] ? ((alt_ok && in->alt_idx[r] == 0) ? 2 : 1) : 0;
                    if (m == 1 && !snv) c = 0;
                    s.push_back(',');
                    put_median(s, c, in->med2_alt[3 * r + m]);
                }
            }
            // rank sums: reported when both lists hold values and the statistic is a number
            for (int e = 0; e < 3; ++e) {
                const int m = RS_METRIC[e];
                if (m == 1 && !snv) continue;
                bool any = false;
                for (int64_t r = r0; r < r1 && !any; ++r)
                    any = ref_ok && has && in->ac[r] > 0 && !std::isnan(in->z[3 * r + m]) && !std::isnan(in->p[3 * r + m]);
                if (!any) continue;
                for (int pass = 0; pass < 2; ++pass) {
                    s += pass ? rs_pv[e] : rs_key[e];
                    bool firstv = true;
                    for (int64_t r = r0; r < r1; ++r) {
                        if (!(ref_ok && in->ac[r] > 0 && !std::isnan(in->z[3 * r + m]) && !std::isnan(in->p[3 * r + m]))) continue;
                        if (!firstv) s.push_back(',');
                        firstv = false;
                        put_float(s, pass ? in->p[3 * r + m] : in->z[3 * r + m]);
                    }
                }
            }
            len[(size_t)i] = (int64_t)(s.size() - at);
        }
    }
    int64_t total = 0;
    for (int64_t i = 0; i < n_rec; ++i) { offsets_out[i] = total; total += len[(size_t)i]; }
    offsets_out[n_rec] = total;
    char *text = (char *)malloc((size_t)total + 1);
    if (!text) return vfkio_fail(VFK_ENOMEM, "vfkio_format_info: out of memory");
    size_t at = 0;
    for (const std::string &s : part) { memcpy(text + at, s.data(), s.size()); at += s.size(); }
    text[total] = 0;
    *text_out = text;
    return VFK_OK;
}

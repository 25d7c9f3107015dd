/*
 * arp_json.cu -- host-side emitter of the atom-atom part of the contact JSON.
 *
 * Replaces, for the records of arp_pairs_fetch, the per-contact Python of
 *   InteractionComplex.get_contacts                       interactions.py:172-196
 *   json.dump(contacts, fp, indent=4, sort_keys=True)     process_protein_cli.py:187-188
 * byte for byte: one dict per record with the keys bgn, contact, distance, end, interacting_entities,
 * type (sorted), four spaces per level.  The caller renders every ATOM once (the 'bgn' / 'end' object,
 * utils.make_pymol_json + label_comp_type, as json.dumps leaves it at nesting depth 2); the emitter joins
 * the pieces: a sizing pass, a prefix sum, then the entries are written by several host threads.
 * No CUDA in this file: it is pure C++ behind the same C ABI.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/arpeggio_cuda.h"

namespace {

/* the names of get_contacts (interactions.py:178-180), SIFt bit order */
const char* const kContact[ARP_SIFT_NBITS] = {"clash", "covalent", "vdw_clash", "vdw", "proximal", "hbond", "weak_hbond", "xbond",
                                              "ionic", "metal_complex", "aromatic", "hydrophobic", "carbonyl", "polar", "weak_polar"};
/* __get_contact_type (interactions.py:643-691) */
const char* const kClass[8] = {"INTRA_NON_SELECTION", "INTRA_SELECTION", "INTER", "SELECTION_WATER", "NON_SELECTION_WATER",
                               "WATER_WATER", "INTRA_BINDING_SITE", ""};

const char kOpen[]     = "    {\n        \"bgn\": ";
const char kContactK[] = ",\n        \"contact\": [";
const char kItem[]     = "\n            \"";
const char kContactE[] = "\n        ]";
const char kDistance[] = ",\n        \"distance\": ";
const char kEnd[]      = ",\n        \"end\": ";
const char kEntity[]   = ",\n        \"interacting_entities\": \"";
const char kType[]     = "\",\n        \"type\": \"atom-atom\"\n    }";

/* float.__repr__ of round(np.float64(dist), 2) (interactions.py:190): NumPy rounds as rint(x * 100) / 100, the
   json encoder prints the shortest string that reads back as the same double */
int format_distance(float dist, char* out)
{
    const double v = rint((double)dist * 100.0) / 100.0;
    if (v != v) return (int)(stpcpy(out, "NaN") - out);
    if (isinf(v)) return (int)(stpcpy(out, v > 0 ? "Infinity" : "-Infinity") - out);
    if (fabs(v) < 1e13) {
        const long long k = llrint(v * 100.0);
        if ((double)k / 100.0 == v) {                     /* k / 100 is what rint(..) / 100 produced */
            char* p = out;
            unsigned long long a = (unsigned long long)(k < 0 ? -k : k);
            if (k < 0 || (k == 0 && signbit(v))) *p++ = '-';
            char tmp[24];
            int n = 0;
            unsigned long long whole = a / 100;
            do { tmp[n++] = (char)('0' + whole % 10); whole /= 10; } while (whole);
            while (n) *p++ = tmp[--n];
            const int frac = (int)(a % 100);
            *p++ = '.';
            *p++ = (char)('0' + frac / 10);
            if (frac % 10) *p++ = (char)('0' + frac % 10);
            return (int)(p - out);
        }
    }
    /* general case: shortest round-trip digits, then Python's repr layout */
    char buf[40];
    int prec = 1;
    for (; prec <= 17; ++prec) {
        snprintf(buf, sizeof buf, "%.*e", prec - 1, v);
        if (strtod(buf, nullptr) == v) break;
    }
    char digits[24];
    int nd = 0;
    const char* e = strchr(buf, 'e');
    for (const char* q = buf; q < e; ++q) if (*q >= '0' && *q <= '9') digits[nd++] = *q;
    while (nd > 1 && digits[nd - 1] == '0') --nd;
    const int exp10 = atoi(e + 1);
    char* p = out;
    if (buf[0] == '-') *p++ = '-';
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {
            *p++ = '0'; *p++ = '.';
            for (int z = 0; z < -exp10 - 1; ++z) *p++ = '0';
            for (int d = 0; d < nd; ++d) *p++ = digits[d];
        } else {
            for (int d = 0; d <= exp10; ++d) *p++ = d < nd ? digits[d] : '0';
            *p++ = '.';
            if (nd > exp10 + 1) for (int d = exp10 + 1; d < nd; ++d) *p++ = digits[d];
            else *p++ = '0';
        }
    } else {
        *p++ = digits[0];
        if (nd > 1) { *p++ = '.'; for (int d = 1; d < nd; ++d) *p++ = digits[d]; }
        p += sprintf(p, "e%c%02d", exp10 < 0 ? '-' : '+', abs(exp10));
    }
    return (int)(p - out);
}

inline uint64_t entry_size(const arp_pair& r, const uint32_t* frag_len)
{
    char num[48];
    uint64_t n = sizeof kOpen - 1 + frag_len[r.i] + sizeof kContactK - 1;
    const uint32_t m = r.mask & 0x7fffu;
    int items = 0;
    for (int b = 0; b < ARP_SIFT_NBITS; ++b)
        if (m >> b & 1u) { n += sizeof kItem - 1 + strlen(kContact[b]) + 1; ++items; }
    n += items ? (uint64_t)(items - 1) + sizeof kContactE - 1 : 1;        /* commas between items; "[]" when empty */
    n += sizeof kDistance - 1 + (uint64_t)format_distance(r.dist, num);
    n += sizeof kEnd - 1 + frag_len[r.j];
    n += sizeof kEntity - 1 + strlen(kClass[(r.mask >> ARP_CLASS_SHIFT) & 7u]) + sizeof kType - 1;
    return n;
}

inline char* put(char* p, const char* s, size_t n) { memcpy(p, s, n); return p + n; }

char* entry_write(const arp_pair& r, const char* const* frag, const uint32_t* frag_len, char* p)
{
    p = put(p, kOpen, sizeof kOpen - 1);
    p = put(p, frag[r.i], frag_len[r.i]);
    p = put(p, kContactK, sizeof kContactK - 1);
    const uint32_t m = r.mask & 0x7fffu;
    bool first = true;
    for (int b = 0; b < ARP_SIFT_NBITS; ++b) {
        if (!(m >> b & 1u)) continue;
        if (!first) *p++ = ',';
        first = false;
        p = put(p, kItem, sizeof kItem - 1);
        p = put(p, kContact[b], strlen(kContact[b]));
        *p++ = '"';
    }
    if (first) *p++ = ']';
    else p = put(p, kContactE, sizeof kContactE - 1);
    p = put(p, kDistance, sizeof kDistance - 1);
    p += format_distance(r.dist, p);
    p = put(p, kEnd, sizeof kEnd - 1);
    p = put(p, frag[r.j], frag_len[r.j]);
    p = put(p, kEntity, sizeof kEntity - 1);
    const char* c = kClass[(r.mask >> ARP_CLASS_SHIFT) & 7u];
    p = put(p, c, strlen(c));
    p = put(p, kType, sizeof kType - 1);
    return p;
}

template <class F>
void parallel_ranges(uint64_t n, int threads, F f)
{
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n / 4096 + 1) threads = (int)(n / 4096 + 1);
    if (threads == 1) { f(0, 0, n); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back(f, t, n * t / threads, n * (t + 1) / threads);
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int arp_pairs_json_size(const arp_pair* rec, uint64_t n, int32_t n_atoms, const uint32_t* frag_len, int threads,
                        uint64_t* bytes)
{
    if (!bytes || (n && (!rec || !frag_len))) return ARP_E_INVALID_ARG;
    std::vector<uint64_t> part((size_t)(threads < 1 ? 1 : threads) + 1, 0);
    std::vector<int> bad(part.size(), 0);
    parallel_ranges(n, threads, [&](int t, uint64_t lo, uint64_t hi) {
        uint64_t s = 0;
        for (uint64_t r = lo; r < hi; ++r) {
            if (rec[r].i < 0 || rec[r].i >= n_atoms || rec[r].j < 0 || rec[r].j >= n_atoms) { bad[t] = 1; return; }
            s += entry_size(rec[r], frag_len);
        }
        part[t] = s;
    });
    uint64_t total = n ? 2 * (n - 1) : 0;                 /* ",\n" between entries */
    for (size_t t = 0; t < part.size(); ++t) { if (bad[t]) return ARP_E_INVALID_ARG; total += part[t]; }
    *bytes = total;
    return ARP_OK;
}

int arp_pairs_json_write(const arp_pair* rec, uint64_t n, int32_t n_atoms, const char* const* frag,
                         const uint32_t* frag_len, int threads, char* dst, uint64_t cap, uint64_t* written)
{
    if (n && (!rec || !frag || !frag_len || !dst)) return ARP_E_INVALID_ARG;
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n / 4096 + 1) threads = (int)(n / 4096 + 1);
    /* sizing pass per thread range, then every thread writes its range at its offset */
    std::vector<uint64_t> part((size_t)threads, 0);
    std::vector<int> bad((size_t)threads, 0);
    parallel_ranges(n, threads, [&](int t, uint64_t lo, uint64_t hi) {
        uint64_t s = 0;
        for (uint64_t r = lo; r < hi; ++r) {
            if (rec[r].i < 0 || rec[r].i >= n_atoms || rec[r].j < 0 || rec[r].j >= n_atoms) { bad[t] = 1; return; }
            s += entry_size(rec[r], frag_len) + 2;
        }
        part[t] = s;
    });
    uint64_t total = 0;
    std::vector<uint64_t> off((size_t)threads, 0);
    for (int t = 0; t < threads; ++t) { if (bad[t]) return ARP_E_INVALID_ARG; off[t] = total; total += part[t]; }
    if (n) total -= 2;                                    /* no separator after the last entry */
    if (total > cap) return ARP_E_CAPACITY;
    parallel_ranges(n, threads, [&](int t, uint64_t lo, uint64_t hi) {
        char* p = dst + off[t];
        for (uint64_t r = lo; r < hi; ++r) {
            p = entry_write(rec[r], frag, frag_len, p);
            if (r + 1 < n) { *p++ = ','; *p++ = '\n'; }
        }
    });
    if (written) *written = total;
    return ARP_OK;
}

}  /* extern "C" */

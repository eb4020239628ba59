// sstar_b200 -- native VCF scanner (host code).
//
// Replaces what the reference gets from scikit-allel (un-vendored dependency):
//     allel.read_vcf(vcf, alt_number=1, samples=ind, region=chr_name)      sstar/utils/utils.py:80
// i.e. per chromosome variants/POS (int32), variants/REF, variants/ALT (first alternate allele) and
// calldata/GT as int8 [V, N, 2] with -1 for a missing allele -- for ALL samples of the file in
// header order; the population split and the filters of read_data happen afterwards
// (sstar_b200/utils/utils.py on the host, ingest_kernels.cuh on the device).
//
// The file (plain text or gzip, through zlib) is read into memory once, data lines are indexed,
// and the lines are parsed in parallel by std::thread workers.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "sstar_b200.h"

namespace {

struct ChromData {
    std::string name;
    std::vector<size_t> line_begin, line_end;  // byte ranges of this chromosome's data lines
    std::vector<int32_t> pos;
    std::vector<int8_t> gt;  // [V, N, 2]
    std::vector<int32_t> ref_off, alt_off;  // [V + 1]
    std::vector<char> ref_bytes, alt_bytes;
};

}  // namespace

struct sstar_b200_vcf {
    std::vector<std::string> samples;
    std::vector<ChromData> chroms;
};

namespace {

thread_local std::string g_vcf_err;

inline const char *find_byte(const char *p, const char *end, char c) {
    const void *q = memchr(p, c, (size_t)(end - p));
    return q ? (const char *)q : end;
}

// one GT field [p, e) -> two alleles; anything after ':' is ignored
inline void parse_gt(const char *p, const char *e, int8_t *out) {
    int vals[2] = {-1, -1};
    int k = 0;
    while (k < 2 && p < e && *p != ':') {
        if (*p >= '0' && *p <= '9') {
            int v = 0;
            while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); if (v > 127) v = 127; ++p; }
            vals[k] = v;
        } else {
            while (p < e && *p != '|' && *p != '/' && *p != ':') ++p;  // '.', empty or junk -> missing
        }
        ++k;
        if (p < e && (*p == '|' || *p == '/')) ++p; else break;
    }
    out[0] = (int8_t)vals[0];
    out[1] = (int8_t)vals[1];
}

struct LineFields {
    const char *ref, *alt;
    int32_t ref_len, alt_len;
};

void parse_line(const char *p, const char *e, int n_samples, int32_t *pos_out, int8_t *gt_out, LineFields *lf) {
    if (e > p && e[-1] == '\r') --e;
    int field = 0;
    int sample = 0;
    memset(gt_out, 0xff, (size_t)n_samples * 2);  // -1 everywhere
    *pos_out = 0;
    lf->ref = lf->alt = p; lf->ref_len = lf->alt_len = 0;
    while (p <= e) {
        const char *t = find_byte(p, e, '\t');
        if (field == 1) {
            long v = 0; bool neg = false; const char *q = p;
            if (q < t && *q == '-') { neg = true; ++q; }
            while (q < t && *q >= '0' && *q <= '9') { v = v * 10 + (*q - '0'); ++q; }
            *pos_out = (int32_t)(neg ? -v : v);
        } else if (field == 3) {
            lf->ref = p; lf->ref_len = (int32_t)(t - p);
        } else if (field == 4) {
            const char *c = find_byte(p, t, ',');
            lf->alt = p; lf->alt_len = (int32_t)(c - p);
        } else if (field >= 9) {
            if (sample < n_samples) parse_gt(p, t, gt_out + (size_t)sample * 2);
            ++sample;
        }
        ++field;
        if (t >= e) break;
        p = t + 1;
    }
}

}  // namespace

extern "C" const char *sstar_b200_vcf_last_error(void) { return g_vcf_err.c_str(); }

extern "C" int sstar_b200_vcf_scan(const char *path, const char *chr_name, int nthreads, sstar_b200_vcf **out) {
    if (!path || !out) { g_vcf_err = "null argument"; return SSTAR_B200_EINVAL; }
    *out = nullptr;
    gzFile gz = gzopen(path, "rb");
    if (!gz) { g_vcf_err = std::string("cannot open ") + path; return SSTAR_B200_EINVAL; }
    gzbuffer(gz, 1 << 20);
    std::vector<char> buf;
    {
        size_t cap = (size_t)64 << 20, len = 0;
        buf.resize(cap);
        for (;;) {
            if (len == cap) { cap *= 2; buf.resize(cap); }
            const size_t want = std::min(cap - len, (size_t)1 << 30);
            const int got = gzread(gz, buf.data() + len, (unsigned)want);
            if (got < 0) { gzclose(gz); g_vcf_err = std::string("read error in ") + path; return SSTAR_B200_EINVAL; }
            if (got == 0) break;
            len += (size_t)got;
        }
        buf.resize(len);
    }
    gzclose(gz);

    auto *h = new sstar_b200_vcf();
    const char *b = buf.data(), *end = b + buf.size();
    std::unordered_map<std::string, int> chrom_id;
    bool have_header = false;
    const std::string want = chr_name ? chr_name : "";
    for (const char *p = b; p < end;) {
        const char *nl = find_byte(p, end, '\n');
        if (nl > p) {
            if (p[0] == '#' && nl - p >= 2 && p[1] == '#') {
                // meta line
            } else if (nl - p >= 6 && memcmp(p, "#CHROM", 6) == 0) {
                h->samples.clear();
                const char *e = nl;
                if (e > p && e[-1] == '\r') --e;
                int field = 0;
                for (const char *q = p; q <= e;) {
                    const char *t = find_byte(q, e, '\t');
                    if (field >= 9) h->samples.emplace_back(q, (size_t)(t - q));
                    ++field;
                    if (t >= e) break;
                    q = t + 1;
                }
                have_header = true;
            } else {
                const char *t = find_byte(p, nl, '\t');
                if (!chr_name || ((size_t)(t - p) == want.size() && memcmp(p, want.data(), want.size()) == 0)) {
                    std::string name(p, (size_t)(t - p));
                    auto it = chrom_id.find(name);
                    int id;
                    if (it == chrom_id.end()) {
                        id = (int)h->chroms.size();
                        chrom_id.emplace(name, id);
                        h->chroms.emplace_back();
                        h->chroms.back().name = name;
                    } else id = it->second;
                    h->chroms[id].line_begin.push_back((size_t)(p - b));
                    h->chroms[id].line_end.push_back((size_t)(nl - b));
                }
            }
        }
        p = nl + 1;
    }
    if (!have_header) {
        delete h;
        g_vcf_err = std::string(path) + " has no #CHROM header line.";
        return SSTAR_B200_EINVAL;
    }
    const int N = (int)h->samples.size();
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    for (auto &c : h->chroms) {
        const size_t V = c.line_begin.size();
        c.pos.resize(V);
        c.gt.resize(V * (size_t)N * 2);
        std::vector<LineFields> lf(V);
        auto work = [&](size_t a, size_t z) {
            for (size_t i = a; i < z; ++i)
                parse_line(b + c.line_begin[i], b + c.line_end[i], N, &c.pos[i], c.gt.data() + i * (size_t)N * 2, &lf[i]);
        };
        const int use = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, V / 256));
        if (use <= 1) work(0, V);
        else {
            std::vector<std::thread> th;
            for (int k = 0; k < use; ++k) th.emplace_back(work, V * k / use, V * (k + 1) / use);
            for (auto &t : th) t.join();
        }
        c.ref_off.resize(V + 1); c.alt_off.resize(V + 1);
        size_t rl = 0, al = 0;
        for (size_t i = 0; i < V; ++i) { c.ref_off[i] = (int32_t)rl; c.alt_off[i] = (int32_t)al; rl += lf[i].ref_len; al += lf[i].alt_len; }
        c.ref_off[V] = (int32_t)rl; c.alt_off[V] = (int32_t)al;
        c.ref_bytes.resize(rl + 1); c.alt_bytes.resize(al + 1);
        for (size_t i = 0; i < V; ++i) {
            memcpy(c.ref_bytes.data() + c.ref_off[i], lf[i].ref, (size_t)lf[i].ref_len);
            memcpy(c.alt_bytes.data() + c.alt_off[i], lf[i].alt, (size_t)lf[i].alt_len);
        }
        c.line_begin.clear(); c.line_begin.shrink_to_fit();
        c.line_end.clear(); c.line_end.shrink_to_fit();
    }
    *out = h;
    return 0;
}

extern "C" void sstar_b200_vcf_free(sstar_b200_vcf *h) { delete h; }
extern "C" int32_t sstar_b200_vcf_n_samples(const sstar_b200_vcf *h) { return (int32_t)h->samples.size(); }
extern "C" const char *sstar_b200_vcf_sample(const sstar_b200_vcf *h, int32_t i) {
    return (i >= 0 && (size_t)i < h->samples.size()) ? h->samples[i].c_str() : nullptr;
}
extern "C" int32_t sstar_b200_vcf_n_chroms(const sstar_b200_vcf *h) { return (int32_t)h->chroms.size(); }
extern "C" const char *sstar_b200_vcf_chrom(const sstar_b200_vcf *h, int32_t c) {
    return (c >= 0 && (size_t)c < h->chroms.size()) ? h->chroms[c].name.c_str() : nullptr;
}
extern "C" int64_t sstar_b200_vcf_n_variants(const sstar_b200_vcf *h, int32_t c) {
    return (c >= 0 && (size_t)c < h->chroms.size()) ? (int64_t)h->chroms[c].pos.size() : -1;
}
extern "C" const int32_t *sstar_b200_vcf_pos(const sstar_b200_vcf *h, int32_t c) { return h->chroms[c].pos.data(); }
extern "C" const int8_t *sstar_b200_vcf_gt(const sstar_b200_vcf *h, int32_t c) { return h->chroms[c].gt.data(); }
extern "C" const char *sstar_b200_vcf_alleles(const sstar_b200_vcf *h, int32_t c, int32_t which, const int32_t **off) {
    const ChromData &d = h->chroms[c];
    *off = which ? d.alt_off.data() : d.ref_off.data();
    return which ? d.alt_bytes.data() : d.ref_bytes.data();
}

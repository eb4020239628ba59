// sstar_b200 -- native VCF scanner (host code).
//
// Replaces what the reference gets from scikit-allel (un-vendored dependency):
//     allel.read_vcf(vcf, alt_number=1, samples=ind, region=chr_name)      sstar/utils/utils.py:80
// i.e. per chromosome variants/POS (int32), variants/REF, variants/ALT (first alternate allele) and
// calldata/GT as int8 [V, N, 2] with -1 for a missing allele -- for ALL samples of the file in
// header order; the population split and the filters of read_data happen afterwards
// (sstar_b200/utils/utils.py on the host, ingest_kernels.cuh on the device).
//
// The file (plain text or gzip, through zlib) is STREAMED in blocks of a few dozen MB: the data lines
// of a block that belong to the wanted chromosome are indexed and parsed in parallel by std::thread
// workers, lines of other chromosomes are skipped before anything is buffered, and only the listed
// samples' columns are parsed and stored (allel.read_vcf(samples=..., region=...) does the same:
// sstar/utils/utils.py:80) -- memory is the kept calls plus one block, whatever the file holds.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "sstar_b200.h"

namespace {

// A byte buffer that grows WITHOUT being zero-filled (std::vector::resize touches every new byte on the scanning
// thread; here the pages are first touched by the parser threads that write them, or by the read that fills them).
struct RawBuf {
    char *p = nullptr;
    size_t cap = 0;
    RawBuf() = default;
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    RawBuf(RawBuf &&o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    RawBuf &operator=(RawBuf &&o) noexcept {
        if (this != &o) { free(p); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
        return *this;
    }
    ~RawBuf() { free(p); }
    bool ensure(size_t n) {  // capacity >= n, contents kept; geometric growth
        if (n <= cap) return true;
        size_t nc = cap ? cap : n;
        while (nc < n) nc *= 2;
        char *q = (char *)realloc(p, nc);
        if (!q) return false;
        p = q;
        cap = nc;
        return true;
    }
};

struct ChromData {
    std::string name;
    std::vector<int32_t> pos;
    RawBuf gt;  // int8 [V, N, 2]
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

// one allele character of the common call shape: a digit or '.'
inline bool allele_char(char c) { return (unsigned)(c - '0') < 10u || c == '.'; }
inline int8_t allele_value(char c) { return c == '.' ? (int8_t)-1 : (int8_t)(c - '0'); }

// sel[c] = output slot of header column c, or -1 (column not wanted); n_cols = header columns,
// n_samples = output slots, last_col = the last wanted column (the rest of the line is not scanned)
void parse_line(const char *p, const char *e, const int32_t *sel, int n_cols, int last_col, int n_samples,
                int32_t *pos_out, int8_t *gt_out, LineFields *lf) {
    if (e > p && e[-1] == '\r') --e;
    if (n_samples > 0) memset(gt_out, 0xff, (size_t)n_samples * 2);  // -1 everywhere
    *pos_out = 0;
    lf->ref = lf->alt = p; lf->ref_len = lf->alt_len = 0;
    // the nine fixed fields: CHROM POS ID REF ALT QUAL FILTER INFO FORMAT
    for (int field = 0; field < 9; ++field) {
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
        }
        if (t >= e) return;  // a line without sample fields
        p = t + 1;
    }
    // the sample fields.  Nearly all of them are "a|b" / "a/b" with one character per allele and nothing
    // behind: those are decoded in place (no search for the field's end); anything else -- haploid calls,
    // multi-digit alleles, FORMAT fields behind the GT -- takes the general parser.
    for (int sample = 0;; ++sample) {
        const int32_t slot = sample < n_cols ? sel[sample] : -1;
        const char *t;
        if (e - p >= 3 && (p[1] == '|' || p[1] == '/') && (p + 3 == e || p[3] == '\t') && allele_char(p[0]) &&
            allele_char(p[2])) {
            if (slot >= 0) {
                gt_out[(size_t)slot * 2] = allele_value(p[0]);
                gt_out[(size_t)slot * 2 + 1] = allele_value(p[2]);
            }
            t = p + 3;
        } else {
            t = find_byte(p, e, '\t');
            if (slot >= 0) parse_gt(p, t, gt_out + (size_t)slot * 2);
        }
        if (sample >= last_col || t >= e) break;
        p = t + 1;
    }
}

// The workers of one scan: created once, woken per phase (a block goes through four parallel phases -- fill, line
// ends, parse, and for BGZF the inflate; spawning sixteen threads for each cost as much as the lighter phases took).
// run(n, f): f(k) for k in [0, n) on workers 0 .. n-1, the caller waits for all of them; n <= 1 runs f(0) inline.
class Workers {
public:
    explicit Workers(int n) : size_(n < 1 ? 1 : n) {}
    Workers(const Workers &) = delete;
    Workers &operator=(const Workers &) = delete;
    ~Workers() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        wake_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return size_; }
    void run(int n, const std::function<void(int)> &f) {
        if (n <= 1) { f(0); return; }
        if (n > size_) {  // (callers size n by the pool; more items than workers are taken in rounds)
            for (int k0 = 0; k0 < n; k0 += size_) {
                const int m = std::min(size_, n - k0);
                if (m == 1) f(k0);
                else run(m, [&](int k) { f(k0 + k); });
            }
            return;
        }
        if (th_.empty())  // started at the first phase that wants them (a small file never does)
            for (int k = 0; k < size_; ++k) th_.emplace_back([this, k] { loop(k); });
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &f;
            active_ = n;
            pending_ = n;
            ++generation_;
        }
        wake_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
        job_ = nullptr;
    }

private:
    void loop(int k) {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            wake_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            if (k < active_) {
                const std::function<void(int)> *f = job_;
                lk.unlock();
                (*f)(k);
                lk.lock();
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    const int size_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable wake_, done_;
    const std::function<void(int)> *job_ = nullptr;
    int active_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

// ---------------------------------------------------------------------------------------------
// BGZF (bgzip, what tabix-indexed VCFs are stored in): a gzip file made of independent members of at most
// 64 KB of text, each carrying its compressed size in a 'BC' extra field (SAM specification, section 4.1).
// The members of a batch are located by walking those sizes and inflated in PARALLEL, each into its place of
// the batch's text (its length is the member's ISIZE trailer); CRC32 and length of every member are checked.
// A plain gzip stream (one member, no sizes) cannot be split and goes through zlib's gzread on one thread.
// ---------------------------------------------------------------------------------------------
inline uint32_t le16(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
inline uint32_t le32(const unsigned char *p) { return le16(p) | (le16(p + 2) << 16); }

// total size of the BGZF member starting at p (n bytes available), 0 = not enough bytes to tell, -1 = not BGZF
long bgzf_member_size(const unsigned char *p, size_t n) {
    if (n < 12) return 0;
    if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return -1;
    const size_t xlen = le16(p + 10);
    if (n < 12 + xlen) return 0;
    for (size_t o = 12; o + 4 <= 12 + xlen;) {
        const size_t slen = le16(p + o + 2);
        if (p[o] == 'B' && p[o + 1] == 'C' && slen == 2 && o + 6 <= 12 + xlen) return (long)le16(p + o + 4) + 1;
        o += 4 + slen;
    }
    return -1;
}

struct BgzfReader {
    int fd = -1;
    Workers *pool = nullptr;
    RawBuf cbuf, obuf;          // compressed bytes read ahead; carry buffer (a member that did not fit the caller's space)
    size_t clen = 0, cpos = 0;  // bytes in cbuf, first unconsumed one
    size_t olen = 0, opos = 0;
    bool ceof = false, at_end = false;
    std::string err;
    struct Member { const unsigned char *src; uint32_t csize, isize; size_t off; };
    std::vector<Member> batch;

    // The next batch of members, inflated into `dst` (at most `cap` bytes of text; members are never split) ->
    // bytes written; 0 = the file has ended, or the next member does not fit `cap`; -1 = error (err set).
    long inflate_batch(char *dst, size_t cap) {
        const size_t ccap = (size_t)4 << 20;  // compressed bytes read ahead (a batch: 64+ members)
        if (!cbuf.ensure(ccap)) { err = "out of memory"; return -1; }
        for (;;) {
            if (cpos > 0) { memmove(cbuf.p, cbuf.p + cpos, clen - cpos); clen -= cpos; cpos = 0; }
            while (!ceof && clen < cbuf.cap) {
                const ssize_t got = read(fd, cbuf.p + clen, cbuf.cap - clen);
                if (got < 0) { err = "read error"; return -1; }
                if (got == 0) ceof = true;
                clen += (size_t)got;
            }
            if (clen == 0) { at_end = true; return 0; }  // end of the file
            batch.clear();
            size_t total = 0;
            bool full = false;
            const unsigned char *base = reinterpret_cast<const unsigned char *>(cbuf.p);
            while (cpos < clen) {
                const long sz = bgzf_member_size(base + cpos, clen - cpos);
                if (sz < 0) { err = "not a BGZF member (mixed gzip / BGZF stream?)"; return -1; }
                if (sz == 0 || cpos + (size_t)sz > clen) {
                    if (ceof) { err = "truncated BGZF member"; return -1; }
                    break;  // the member continues in the bytes not read yet
                }
                const size_t xlen = le16(base + cpos + 10);
                if ((size_t)sz < 12 + xlen + 8) { err = "malformed BGZF member"; return -1; }
                const uint32_t isize = le32(base + cpos + sz - 4);
                if (isize > (1u << 16)) { err = "malformed BGZF member (more than 64 KB of text)"; return -1; }
                if (total + isize > cap) { full = true; break; }
                batch.push_back({base + cpos, (uint32_t)sz, isize, total});
                total += isize;
                cpos += (size_t)sz;
            }
            if (batch.empty()) {
                if (full) return 0;  // the caller's space is smaller than the next member
                err = "malformed BGZF member";
                return -1;
            }
            std::atomic<bool> bad(false);
            auto work = [&](size_t a, size_t z) {
                z_stream zs;
                memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) { bad = true; return; }
                for (size_t i = a; i < z; ++i) {
                    const Member &m = batch[i];
                    if (m.isize == 0) continue;  // (the empty member that ends a BGZF file)
                    const size_t xlen = le16(m.src + 10);
                    inflateReset(&zs);
                    zs.next_in = const_cast<unsigned char *>(m.src + 12 + xlen);
                    zs.avail_in = (uInt)(m.csize - 12 - xlen - 8);
                    zs.next_out = reinterpret_cast<unsigned char *>(dst + m.off);
                    zs.avail_out = m.isize;
                    const int r = inflate(&zs, Z_FINISH);
                    if (r != Z_STREAM_END || zs.avail_out != 0 ||
                        (uint32_t)crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<unsigned char *>(dst + m.off), m.isize) !=
                            le32(m.src + m.csize - 8))
                        bad = true;
                }
                inflateEnd(&zs);
            };
            const int use = (int)std::min<size_t>((size_t)pool->size(), std::max<size_t>(1, batch.size() / 8));
            pool->run(use, [&](int k) { work(batch.size() * (size_t)k / use, batch.size() * (size_t)(k + 1) / use); });
            if (bad) { err = "corrupt BGZF member (inflate / CRC32 / length check failed)"; return -1; }
            if (total > 0) return (long)total;
            // (only empty members so far: read on)
        }
    }

    // up to `want` bytes of text into dst; 0 at the end, -1 on error.  Whole members are inflated straight into
    // dst; only when the next member is larger than the space left does it go through the 64 KB carry buffer.
    long fill(char *dst, size_t want) {
        if (opos == olen) {
            const long got = inflate_batch(dst, want);
            if (got != 0 || at_end) return got;
            if (!obuf.ensure((size_t)1 << 16)) { err = "out of memory"; return -1; }
            const long one = inflate_batch(obuf.p, (size_t)1 << 16);  // (a member holds at most 64 KB)
            if (one <= 0) return one;
            olen = (size_t)one;
            opos = 0;
        }
        const size_t n = std::min(want, olen - opos);
        memcpy(dst, obuf.p + opos, n);
        opos += n;
        return (long)n;
    }
};

}  // namespace

extern "C" const char *sstar_b200_vcf_last_error(void) { return g_vcf_err.c_str(); }

extern "C" int sstar_b200_vcf_scan_ex(const char *path, const char *chr_name, const char *const *samples,
                                      int32_t n_samples, int nthreads, sstar_b200_vcf **out) {
    if (!path || !out || (n_samples > 0 && !samples)) { g_vcf_err = "null argument"; return SSTAR_B200_EINVAL; }
    *out = nullptr;
    // plain text is read straight into the block (read(2)); only a gzip stream goes through zlib
    int fd = open(path, O_RDONLY);
    if (fd < 0) { g_vcf_err = std::string("cannot open ") + path; return SSTAR_B200_EINVAL; }
    unsigned char magic[64] = {0};
    const ssize_t nm = pread(fd, magic, sizeof magic, 0);
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 64) nt = 64;
    gzFile gz = nullptr;
    BgzfReader bg;  // BGZF members are inflated in parallel; any other gzip stream by zlib on this thread
    const bool bgzf = nm >= 18 && bgzf_member_size(magic, (size_t)nm) > 0;
    Workers pool(nt);
    if (bgzf) {
        bg.fd = fd;
        bg.pool = &pool;
    } else if (nm >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        gz = gzdopen(fd, "rb");  // (owns fd from here on)
        if (!gz) { close(fd); g_vcf_err = std::string("cannot open ") + path; return SSTAR_B200_EINVAL; }
        gzbuffer(gz, 1 << 20);
    }
    auto close_input = [&]() { if (gz) gzclose(gz); else close(fd); };
    // plain text in a regular file: the block is filled by the worker threads, a slice of the file each (pread)
    struct stat sb;
    const bool sliced = !bgzf && !gz && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode);
    size_t file_off = 0;

    auto *h = new sstar_b200_vcf();
    std::unordered_set<std::string> wanted;
    const bool subset = samples != nullptr;
    for (int32_t i = 0; subset && i < n_samples; ++i)
        if (samples[i]) wanted.insert(samples[i]);
    std::unordered_map<std::string, int> chrom_id;
    std::vector<int32_t> sel;  // header column -> output slot (or -1)
    int n_cols = 0, last_col = -1, N = 0;
    bool have_header = false;
    const std::string want = chr_name ? chr_name : "";

    struct LineRef { size_t begin, end; int chrom; size_t row; };
    std::vector<LineRef> lines;
    std::vector<LineFields> lf;
    std::vector<std::vector<size_t>> nl_parts;  // newline offsets of the block, one list per worker slice
    size_t block = (size_t)32 << 20;
    if (const char *env = getenv("SSTAR_B200_VCF_BLOCK")) {  // tests shrink the block to cross its edges
        const long long v = atoll(env);
        if (v >= 64) block = (size_t)v;
    }
    RawBuf buf;
    if (!buf.ensure(block)) { close_input(); delete h; g_vcf_err = "out of memory"; return SSTAR_B200_EINVAL; }
    size_t len = 0;  // bytes in buf
    std::string last_name;  // chromosome of the previous kept line (lines of one chromosome come in runs)
    int last_id = -1;
    bool eof = false;
    int rc = 0;
    // SSTAR_B200_VCF_TIMING=1: milliseconds per phase on stderr (read / line ends / line loop / parse / alleles)
    static const bool timing = [] { const char *e = getenv("SSTAR_B200_VCF_TIMING"); return e && atoi(e) != 0; }();
    double phase_ms[5] = {0, 0, 0, 0, 0};
    auto tick = [] { return std::chrono::steady_clock::now(); };
    auto lap = [&](int i, std::chrono::steady_clock::time_point &t) {
        if (!timing) return;
        const auto n = tick();
        phase_ms[i] += std::chrono::duration<double, std::milli>(n - t).count();
        t = n;
    };
    while (!eof || len > 0) {
        auto t_phase = tick();
        while (!eof && len < buf.cap) {  // top the block up
            size_t wantb = std::min(buf.cap - len, (size_t)1 << 30);
            long got;
            if (sliced) {
                const size_t left = (size_t)sb.st_size > file_off ? (size_t)sb.st_size - file_off : 0;
                wantb = std::min(wantb, left);
                const int use = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, wantb >> 21));  // >= 2 MB a thread
                std::atomic<long> total(0);
                std::atomic<bool> failed(false);
                char *dst = buf.p + len;
                pool.run(use, [&](int k) {
                    size_t a = wantb * (size_t)k / use;
                    const size_t z = wantb * (size_t)(k + 1) / use;
                    while (a < z) {
                        const ssize_t g = pread(fd, dst + a, z - a, (off_t)(file_off + a));
                        if (g < 0) { failed = true; return; }
                        if (g == 0) return;  // (the file shrank under us: the short count below ends the scan)
                        a += (size_t)g;
                        total += (long)g;
                    }
                });
                got = failed ? -1 : (total == (long)wantb ? (long)wantb : -1);
                if (wantb == 0) got = 0;
                if (got > 0) file_off += (size_t)got;
            } else {
                got = bgzf ? bg.fill(buf.p + len, wantb)
                      : gz ? (long)gzread(gz, buf.p + len, (unsigned)wantb) : (long)read(fd, buf.p + len, wantb);
            }
            if (got < 0) {
                g_vcf_err = std::string("read error in ") + path + (bgzf && !bg.err.empty() ? ": " + bg.err : "");
                rc = SSTAR_B200_EINVAL;
                break;
            }
            if (got == 0) eof = true;
            len += (size_t)got;
        }
        if (rc) break;
        lap(0, t_phase);
        const char *b = buf.p, *end = b + len;
        // the block's complete lines end at the last newline (at EOF the unterminated tail is a line too)
        const char *stop = end;
        if (!eof) {
            while (stop > b && stop[-1] != '\n') --stop;
            if (stop == b) {  // one line longer than the block: grow and read on
                if (!buf.ensure(buf.cap * 2)) { g_vcf_err = "out of memory"; rc = SSTAR_B200_EINVAL; break; }
                continue;
            }
        }
        lines.clear();
        // line ends of the block: every worker lists the newlines of its slice, the slices are walked in order
        const size_t span = (size_t)(stop - b);
        const int iuse = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, span >> 20));  // >= 1 MB a thread
        if ((int)nl_parts.size() < iuse) nl_parts.resize(iuse);
        pool.run(iuse, [&](int k) {
            std::vector<size_t> &v = nl_parts[k];
            v.clear();
            const char *q = b + span * (size_t)k / iuse, *z = b + span * (size_t)(k + 1) / iuse;
            while (q < z) {
                const char *nl = find_byte(q, z, '\n');
                if (nl >= z) break;
                v.push_back((size_t)(nl - b));
                q = nl + 1;
            }
        });
        lap(1, t_phase);
        int part = 0;
        size_t part_i = 0;
        for (const char *p = b; p < stop;) {
            while (part < iuse && part_i >= nl_parts[part].size()) { ++part; part_i = 0; }
            const char *nl = part < iuse ? b + nl_parts[part][part_i++] : stop;  // (at EOF the tail has no newline)
            if (nl > p) {
                if (p[0] == '#' && nl - p >= 2 && p[1] == '#') {
                    // meta line
                } else if (nl - p >= 6 && memcmp(p, "#CHROM", 6) == 0) {
                    h->samples.clear();
                    sel.clear();
                    const char *e = nl;
                    if (e > p && e[-1] == '\r') --e;
                    int field = 0;
                    for (const char *q = p; q <= e;) {
                        const char *t = find_byte(q, e, '\t');
                        if (field >= 9) {
                            std::string name(q, (size_t)(t - q));
                            if (!subset || wanted.count(name)) {
                                sel.push_back((int32_t)h->samples.size());
                                h->samples.push_back(std::move(name));
                            } else sel.push_back(-1);
                        }
                        ++field;
                        if (t >= e) break;
                        q = t + 1;
                    }
                    n_cols = (int)sel.size();
                    N = (int)h->samples.size();
                    last_col = -1;
                    for (int c = 0; c < n_cols; ++c) if (sel[c] >= 0) last_col = c;
                    have_header = true;
                } else if (have_header) {
                    const char *t = find_byte(p, nl, '\t');
                    if (!chr_name || ((size_t)(t - p) == want.size() && memcmp(p, want.data(), want.size()) == 0)) {
                        int id;
                        if (last_id >= 0 && (size_t)(t - p) == last_name.size() && memcmp(p, last_name.data(), last_name.size()) == 0) {
                            id = last_id;
                        } else {
                            std::string name(p, (size_t)(t - p));
                            auto it = chrom_id.find(name);
                            if (it == chrom_id.end()) {
                                id = (int)h->chroms.size();
                                chrom_id.emplace(name, id);
                                h->chroms.emplace_back();
                                h->chroms.back().name = name;
                                h->chroms.back().ref_off.push_back(0);
                                h->chroms.back().alt_off.push_back(0);
                            } else id = it->second;
                            last_name = std::move(name);
                            last_id = id;
                        }
                        ChromData &c = h->chroms[id];
                        lines.push_back({(size_t)(p - b), (size_t)(nl - b), id, c.pos.size()});
                        c.pos.push_back(0);
                    }
                } else {
                    g_vcf_err = std::string(path) + " has no #CHROM header line.";
                    rc = SSTAR_B200_EINVAL;
                    break;
                }
            }
            p = nl + 1;
        }
        if (rc) break;
        lap(2, t_phase);
        // parse the block's kept lines in parallel, straight into the chromosomes' arrays
        for (auto &c : h->chroms)
            if (!c.gt.ensure(c.pos.size() * (size_t)N * 2)) { g_vcf_err = "out of memory"; rc = SSTAR_B200_EINVAL; }
        if (rc) break;
        lf.resize(lines.size());
        const size_t L = lines.size();
        auto work = [&](size_t a, size_t z) {
            for (size_t i = a; i < z; ++i) {
                ChromData &c = h->chroms[lines[i].chrom];
                parse_line(b + lines[i].begin, b + lines[i].end, sel.data(), n_cols, last_col, N, &c.pos[lines[i].row],
                           reinterpret_cast<int8_t *>(c.gt.p) + lines[i].row * (size_t)N * 2, &lf[i]);
            }
        };
        const int use = (int)std::min<size_t>((size_t)nt, std::max<size_t>(1, L / 256));
        pool.run(use, [&](int k) { work(L * (size_t)k / use, L * (size_t)(k + 1) / use); });
        lap(3, t_phase);
        for (size_t i = 0; i < L; ++i) {  // the alleles are copied out before the block is recycled
            ChromData &c = h->chroms[lines[i].chrom];
            c.ref_bytes.insert(c.ref_bytes.end(), lf[i].ref, lf[i].ref + lf[i].ref_len);
            c.alt_bytes.insert(c.alt_bytes.end(), lf[i].alt, lf[i].alt + lf[i].alt_len);
            c.ref_off.push_back((int32_t)c.ref_bytes.size());
            c.alt_off.push_back((int32_t)c.alt_bytes.size());
        }
        lap(4, t_phase);
        const size_t used = (size_t)(stop - b);
        memmove(buf.p, buf.p + used, len - used);
        len -= used;
        if (eof) len = 0;
    }
    close_input();
    if (timing)
        fprintf(stderr, "[vcf_scan] %d threads: read %.2f  line ends %.2f  line loop %.2f  parse %.2f  alleles %.2f ms\n", nt,
                phase_ms[0], phase_ms[1], phase_ms[2], phase_ms[3], phase_ms[4]);
    if (!rc && !have_header) {
        g_vcf_err = std::string(path) + " has no #CHROM header line.";
        rc = SSTAR_B200_EINVAL;
    }
    if (rc) { delete h; return rc; }
    for (auto &c : h->chroms) {  // (offset arrays keep their [V + 1] shape; the byte arrays one spare byte)
        c.ref_bytes.push_back(0);
        c.alt_bytes.push_back(0);
    }
    *out = h;
    return 0;
}

extern "C" int sstar_b200_vcf_scan(const char *path, const char *chr_name, int nthreads, sstar_b200_vcf **out) {
    return sstar_b200_vcf_scan_ex(path, chr_name, nullptr, 0, nthreads, out);
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
extern "C" const int8_t *sstar_b200_vcf_gt(const sstar_b200_vcf *h, int32_t c) {
    return reinterpret_cast<const int8_t *>(h->chroms[c].gt.p);
}
extern "C" const char *sstar_b200_vcf_alleles(const sstar_b200_vcf *h, int32_t c, int32_t which, const int32_t **off) {
    const ChromData &d = h->chroms[c];
    *off = which ? d.alt_off.data() : d.ref_off.data();
    return which ? d.alt_bytes.data() : d.ref_bytes.data();
}

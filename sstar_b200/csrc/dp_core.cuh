// sstar_b200 -- the per-unit chain DP, written once for device and host.
//
// These two functions are the arithmetic of the score kernel.  They compile for the GPU (called
// from score_kernels.cuh) and, unchanged, for the host, where tests/emul builds them with g++ and
// checks them against the oracle without a GPU.  Nothing in the product calls the host build.
//
// Recurrence (sstar/sstar.py:195-217).  With v_i = max(M[i], 0):
//     M[j] = max_{i<j, p_j-p_i>=10} s(j,i) + v_i
// and for carried genotypes in {1,2}:  g_i == g_j -> s = (p_j - p_i) + bonus
//                                      g_i != g_j -> s = penalty      (only if max_mismatch >= 1)
// so  M[j] = max( p_j + bonus + max_{g_i=g_j}(v_i - p_i) ,  penalty + max_{g_i!=g_j} v_i ):
// four running maxima (A0, A1, B0, B1).  Positions ascend, so the eligible predecessors of site j
// are a prefix of the sites before it: a site is "committed" to the maxima once the current site is
// >= 10 bp ahead.  At most 10 distinct integer positions fit in 9 bp, so at most 10 sites wait; the
// newest waits in registers, older ones in a 16-slot thread-local ring.
//
// "-inf" is a sentinel NEG with |real values| < |NEG| / 4: invalid candidates stay below THR
// without selects, m = max(same + p + bonus, other + penalty) is valid iff m > THR, and
// v = max(m, 0) is 0 for an invalid m -- exactly max(M[i] + s, s) of the reference.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SSTAR_HD __host__ __device__ __forceinline__
#define SSTAR_HD_COLD __host__ __device__ __noinline__  // rare paths: keep them out of the hot kernel body
#else
#define SSTAR_HD inline
#define SSTAR_HD_COLD inline
#endif

namespace sstar {

constexpr int kMinPairDistance = 10;  // sstar/sstar.py:195: pairs closer than 10 bp never link
constexpr int kRingSlots = 16;        // >= 10 waiting sites
constexpr uint32_t kListSentinel = 0xffffffffu;  // list terminator: a position beyond any window
constexpr int64_t kNegInf64 = INT64_MIN;

template <typename acc_t> struct Sentinel;
template <> struct Sentinel<int32_t> {
    static constexpr int32_t NEG = -(1 << 30);
    static constexpr int32_t THR = -(1 << 29);
};
template <> struct Sentinel<int64_t> {
    static constexpr int64_t NEG = -(int64_t(1) << 62);
    static constexpr int64_t THR = -(int64_t(1) << 61);
};
// int32 arithmetic is exact while |any chain score| < 2^28
constexpr int64_t kSmallBound = int64_t(1) << 28;

template <typename T> SSTAR_HD T imax(T a, T b) { return a > b ? a : b; }
SSTAR_HD int ctz32(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// |any chain score| inside a span of `span` bp holding `nvar` variants
SSTAR_HD int64_t score_bound(int64_t span, int64_t nvar, int32_t bonus, int32_t penalty) {
    const int64_t mag = (bonus < 0 ? -(int64_t)bonus : (int64_t)bonus) +
                        (penalty < 0 ? -(int64_t)penalty : (int64_t)penalty);
    return span + (nvar + 1) * mag;
}

// ---------------------------------------------------------------------------------------------
// List DP (int32).  `list` holds one column's carried reference-absent sites of a window GROUP in
// ascending position, entry = (position relative to the group's first variant) << 1 | (genotype==2);
// consecutive entries are `ls` words apart (a strided shared-memory column on the device, 1 on the
// host).  Scores the window [plo, phi] (relative positions,
// closed).  `a` is the list cursor: windows of a group are visited with non-decreasing plo, so it
// only moves forward.  The caller terminates the list with kListSentinel at index cnt.
// ---------------------------------------------------------------------------------------------
// Scores exactly the n sites starting at `base` (a run of one column's list).  n >= 1.
// kShift = 1: entries (position << 1 | genotype==2).  kShift = 8: general entries (position << 8 |
// genotype byte) of a unit whose genotypes all are 1 or 2 -- the same DP, decoded differently.
// kShift = 2: signed entries (position << 2 | sign << 1 | |genotype|==2) of a unit without a negative genotype.
template <int kShift>
SSTAR_HD bool entry_is_two(uint32_t e) { return kShift == 8 ? (e & 0xffu) == 2u : (e & 1u) != 0; }

// true when every one of the n general entries holds genotype 1 or 2
SSTAR_HD bool general_run_is_plain(const uint32_t *base, int ls, int n) {
    bool plain = true;
    for (int k = 0; k < n; ++k) plain = plain && ((base[k * ls] & 0xffu) - 1u) < 2u;
    return plain;
}

template <int kShift = 1>
SSTAR_HD int64_t list_dp_run(const uint32_t *base, int ls, int n, int32_t bonus, int32_t pen_eff) {
    constexpr int32_t NEG = Sentinel<int32_t>::NEG, THR = Sentinel<int32_t>::THR;
    int32_t ring[kRingSlots];  // v of waiting sites older than k-1: thread-local memory, touched only
                               // when carried sites crowd within 9 bp (a few percent of the sites)
    const uint32_t *pk = base;  // next site to score
    int c = 0;                  // index of the next site to commit
    int32_t A0 = NEG, A1 = NEG, B0 = NEG, B1 = NEG, top = NEG;
    uint32_t pe = 0;  // entry k-1 and its v
    int32_t pv = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t e = *pk;
        const int32_t p = (int32_t)(e >> kShift);
        if (c == k - 1) {  // common case: only the previous site waits, in registers
            const int32_t pc = (int32_t)(pe >> kShift);
            if (p - pc >= kMinPairDistance) {
                if (entry_is_two<kShift>(pe)) { A1 = imax(A1, pv - pc); B1 = imax(B1, pv); }
                else         { A0 = imax(A0, pv - pc); B0 = imax(B0, pv); }
                c = k;
            } else {
                ring[(k - 1) & (kRingSlots - 1)] = pv;  // k-1 keeps waiting: park its v
            }
        } else if (c < k) {  // crowded sites: older ones wait in the ring
            while (c < k) {
                uint32_t ec;
                int32_t vc;
                if (c == k - 1) { ec = pe; vc = pv; }
                else            { ec = base[c * ls]; vc = ring[c & (kRingSlots - 1)]; }
                const int32_t pc = (int32_t)(ec >> kShift);
                if (p - pc < kMinPairDistance) break;
                if (entry_is_two<kShift>(ec)) { A1 = imax(A1, vc - pc); B1 = imax(B1, vc); }
                else         { A0 = imax(A0, vc - pc); B0 = imax(B0, vc); }
                ++c;
            }
            if (c < k) ring[(k - 1) & (kRingSlots - 1)] = pv;
        }
        const bool cls = entry_is_two<kShift>(e);
        const int32_t same = cls ? A1 : A0;
        const int32_t other = cls ? B0 : B1;
        const int32_t m = imax(same + (p + bonus), other + pen_eff);
        top = imax(top, m);
        pe = e;
        pv = imax(m, (int32_t)0);
        pk += ls;
    }
    return (n <= 2 || top <= THR) ? kNegInf64 : (int64_t)top;
}

// ---------------------------------------------------------------------------------------------
// General-genotype list DP: the same prefix-max recurrence for ANY int8 genotype (missing calls
// give negative dosages, multi-allelic sites values > 2: sstar/utils/utils.py:305-306,
// sstar/sstar.py:163-164,179).  Entry = (relative position << 8) | (genotype & 0xff).  The running
// maxima are kept per distinct genotype value (class) in small thread-local tables:
//     M[j] = max( p_j + bonus + A[g_j] ,  penalty + max_{g' != g_j, |g' - g_j| <= max_mm} B[g'] ).
// Returns false when a unit holds more than kMaxClasses distinct values (the caller routes the
// window to the pairwise kernel).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxClasses = 8;
constexpr int kGeneralPosBits = 23;  // relative positions below 2^23 (8 Mb per window group)

template <int kShift>
SSTAR_HD void list_window_range_shifted(const uint32_t *list, int ls, int &a, int &b, uint32_t plo, uint32_t phi,
                                        bool b_monotone) {
    const uint32_t *pa = list + a * ls;
    while ((*pa >> kShift) < plo) { pa += ls; ++a; }
    if (!b_monotone || b < a) b = a;
    const uint32_t *pb = list + b * ls;
    while ((*pb >> kShift) <= phi) { pb += ls; ++b; }
}

SSTAR_HD_COLD bool list_dp_run_general(const uint32_t *base, int ls, int n, int32_t bonus, int32_t penalty, int32_t max_mm,
                                  int64_t &score_out) {
    constexpr int32_t NEG = Sentinel<int32_t>::NEG, THR = Sentinel<int32_t>::THR;
    int32_t A[kMaxClasses], B[kMaxClasses], val[kMaxClasses];
    int32_t ring_v[kRingSlots];
    int8_t ring_c[kRingSlots];  // class of the waiting sites (found when they were scored)
    int K = 0;
    int c = 0;  // next site to commit
    int32_t top = NEG;
    auto class_of = [&](int32_t g) -> int {
        for (int i = 0; i < K; ++i)
            if (val[i] == g) return i;
        return -1;
    };
    for (int k = 0; k < n; ++k) {
        const uint32_t e = base[k * ls];
        const int32_t p = (int32_t)(e >> 8);
        const int32_t g = (int32_t)(int8_t)(e & 0xffu);
        while (c < k) {  // sites >= 10 bp behind become eligible predecessors
            const uint32_t ec = base[c * ls];
            const int32_t pc = (int32_t)(ec >> 8);
            if (p - pc < kMinPairDistance) break;
            const int ci = ring_c[c & (kRingSlots - 1)];
            const int32_t vc = ring_v[c & (kRingSlots - 1)];
            A[ci] = imax(A[ci], vc - pc);
            B[ci] = imax(B[ci], vc);
            ++c;
        }
        int idx = class_of(g);
        if (idx < 0) {
            if (K == kMaxClasses) return false;
            idx = K++;
            val[idx] = g; A[idx] = NEG; B[idx] = NEG;
        }
        int32_t other = NEG;
        for (int i = 0; i < K; ++i) {
            const int32_t d = val[i] > g ? val[i] - g : g - val[i];
            if (i != idx && d <= max_mm) other = imax(other, B[i]);
        }
        const int32_t m = imax(A[idx] + (p + bonus), other + penalty);
        top = imax(top, m);
        ring_v[k & (kRingSlots - 1)] = imax(m, (int32_t)0);
        ring_c[k & (kRingSlots - 1)] = (int8_t)idx;
    }
    score_out = (n <= 2 || top <= THR) ? kNegInf64 : (int64_t)top;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Small signed genotypes: units whose carried genotypes all lie in {-2, -1, 1, 2} -- what missing calls
// give (an allele of -1: sstar/utils/utils.py:305-306 sums the alleles, so "./." is -2, a missing
// haploid call or "./0" -1).  Signed entries = (relative position << 2) | (g < 0) << 1 | (|g| == 2), built
// from three bit-planes without touching the genotype bytes; the low two bits are the class:
// 0: g = 1, 1: g = 2, 2: g = -1, 3: g = -2.  A unit without a negative genotype is scored by
// list_dp_run<2> (the four-maxima DP reads bit 0 and the position only); the others by
// list_dp_run_signed: four classes, their running maxima in REGISTERS (list_dp_run_general keeps
// them in thread-local tables, whose loads and stores are the latency of that loop), waiting sites as
// in list_dp_run.  `compat` comes from signed_compat_mask(max_mismatch).
// ---------------------------------------------------------------------------------------------
constexpr int kSignedShift = 2;
SSTAR_HD int signed_class_value(int ci) { return (ci & 2) ? -1 - (ci & 1) : 1 + (ci & 1); }
// bit (4 * i + j): classes i != j whose genotype values differ by at most max_mm (sstar.py:202-206)
SSTAR_HD uint32_t signed_compat_mask(int32_t max_mm) {
    uint32_t m = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            const int vi = signed_class_value(i), vj = signed_class_value(j);
            const int d = vi > vj ? vi - vj : vj - vi;
            if (i != j && d <= max_mm) m |= 1u << (4 * i + j);
        }
    return m;
}

SSTAR_HD_COLD int64_t list_dp_run_signed(const uint32_t *base, int ls, int n, int32_t bonus, int32_t penalty, uint32_t compat) {
    constexpr int32_t NEG = Sentinel<int32_t>::NEG, THR = Sentinel<int32_t>::THR;
    int32_t ring[kRingSlots];  // v of waiting sites older than k-1 (crowded sites only)
    int32_t A0 = NEG, A1 = NEG, A2 = NEG, A3 = NEG, B0 = NEG, B1 = NEG, B2 = NEG, B3 = NEG, top = NEG;
    auto commit = [&](uint32_t ec, int32_t vc) {  // the site becomes an eligible predecessor
        const int32_t d = vc - (int32_t)(ec >> kSignedShift);
        const int ci = (int)(ec & 3u);
        if (ci == 0)      { A0 = imax(A0, d); B0 = imax(B0, vc); }
        else if (ci == 1) { A1 = imax(A1, d); B1 = imax(B1, vc); }
        else if (ci == 2) { A2 = imax(A2, d); B2 = imax(B2, vc); }
        else              { A3 = imax(A3, d); B3 = imax(B3, vc); }
    };
    const uint32_t *pk = base;
    int c = 0;        // next site to commit
    uint32_t pe = 0;  // entry k-1 and its v
    int32_t pv = 0;
    for (int k = 0; k < n; ++k) {
        const uint32_t e = *pk;
        const int32_t p = (int32_t)(e >> kSignedShift);
        if (c == k - 1) {  // common case: only the previous site waits, in registers
            if (p - (int32_t)(pe >> kSignedShift) >= kMinPairDistance) { commit(pe, pv); c = k; }
            else ring[(k - 1) & (kRingSlots - 1)] = pv;
        } else if (c < k) {  // crowded sites: older ones wait in the ring
            while (c < k) {
                const uint32_t ec = (c == k - 1) ? pe : base[c * ls];
                const int32_t vc = (c == k - 1) ? pv : ring[c & (kRingSlots - 1)];
                if (p - (int32_t)(ec >> kSignedShift) < kMinPairDistance) break;
                commit(ec, vc);
                ++c;
            }
            if (c < k) ring[(k - 1) & (kRingSlots - 1)] = pv;
        }
        const int ci = (int)(e & 3u);
        const int32_t same = ci == 0 ? A0 : (ci == 1 ? A1 : (ci == 2 ? A2 : A3));
        const uint32_t cm = compat >> (4 * ci);
        int32_t other = NEG;
        if (cm & 1u) other = imax(other, B0);
        if (cm & 2u) other = imax(other, B1);
        if (cm & 4u) other = imax(other, B2);
        if (cm & 8u) other = imax(other, B3);
        const int32_t m = imax(same + (p + bonus), other + penalty);
        top = imax(top, m);
        pe = e;
        pv = imax(m, (int32_t)0);
        pk += ls;
    }
    return (n <= 2 || top <= THR) ? kNegInf64 : (int64_t)top;
}

// list_window_range_shifted<2> for signed entries that also keeps, for both cursors, the number of entries
// passed whose genotype is negative: a unit is "plain" (list_dp_run<2> takes it) iff the two counts agree.
SSTAR_HD void list_window_range_signed(const uint32_t *list, int ls, int &a, int &b, int &odd_a, int &odd_b, uint32_t plo,
                                       uint32_t phi) {
    const uint32_t *pa = list + a * ls;
    while ((*pa >> kSignedShift) < plo) { odd_a += (int)((*pa >> 1) & 1u); pa += ls; ++a; }
    if (b < a) { b = a; odd_b = odd_a; }
    const uint32_t *pb = list + b * ls;
    while ((*pb >> kSignedShift) <= phi) { odd_b += (int)((*pb >> 1) & 1u); pb += ls; ++b; }
}

// list_window_range_shifted<8> for general entries that also keeps, for both cursors, the number of
// entries passed whose genotype is not 1 / 2: a unit is "plain" (the four-maxima DP takes it) iff the
// two counts agree -- O(1) per unit instead of a scan of its entries.
SSTAR_HD void list_window_range_general(const uint32_t *list, int ls, int &a, int &b, int &odd_a, int &odd_b, uint32_t plo,
                                        uint32_t phi) {
    const uint32_t *pa = list + a * ls;
    while ((*pa >> 8) < plo) { odd_a += (int)(((*pa & 0xffu) - 1u) >= 2u); pa += ls; ++a; }
    if (b < a) { b = a; odd_b = odd_a; }
    const uint32_t *pb = list + b * ls;
    while ((*pb >> 8) <= phi) { odd_b += (int)(((*pb & 0xffu) - 1u) >= 2u); pb += ls; ++b; }
}

// list_window_range_shifted moves the cursors to the window [plo, phi]: a = first entry with
// position >= plo, b = first entry with position > phi.  The list ends with kListSentinel, so no
// bound checks.  With b_monotone the end cursor continues from where the previous window left it
// (windows whose ends never decrease); otherwise it restarts at a.
SSTAR_HD void list_window_range(const uint32_t *list, int ls, int &a, int &b, uint32_t plo, uint32_t phi,
                                bool b_monotone) {
    list_window_range_shifted<1>(list, ls, a, b, plo, phi, b_monotone);
}

SSTAR_HD void list_dp_window(const uint32_t *list, int ls, int &a, uint32_t plo, uint32_t phi, int32_t bonus,
                             int32_t pen_eff, int &n_out, int64_t &score_out) {
    int b = a;
    list_window_range(list, ls, a, b, plo, phi, false);
    n_out = b - a;
    score_out = n_out <= 2 ? kNegInf64 : list_dp_run(list + a * ls, ls, n_out, bonus, pen_eff);
}

// ---------------------------------------------------------------------------------------------
// Walk DP: the same recurrence straight from the word-major bit-planes in global memory, for
// units the list path cannot take (very long windows, a column with more sites than list slots,
// scores beyond int32).  Column t of plane word g is plane[g * Tp + t].  The waiting sites live in
// two small per-thread arrays.
// ---------------------------------------------------------------------------------------------
template <typename acc_t, typename LoadU32, typename LoadPos>
SSTAR_HD void walk_dp_unit(LoadU32 ld, LoadPos ldpos, const uint32_t *carried, const uint32_t *two, int32_t Tp,
                           int t, int32_t lo, int32_t hi, int32_t bonus_i, int32_t penalty_i, int32_t max_mm,
                           int &n_out, int64_t &score_out) {
    constexpr acc_t NEG = Sentinel<acc_t>::NEG, THR = Sentinel<acc_t>::THR;
    const acc_t bonus = (acc_t)bonus_i;
    const acc_t pen_eff = (max_mm >= 1) ? (acc_t)penalty_i : NEG;
    acc_t A0 = NEG, A1 = NEG, B0 = NEG, B1 = NEG, top = NEG;
    acc_t wait_p[kRingSlots], wait_v[kRingSlots];  // v << 1 | class
    int q = 0, head = 0, n = 0;
    const int64_t p0 = ldpos(lo);
    const int gi0 = lo >> 5, gi1 = (hi - 1) >> 5;
    for (int gi = gi0; gi <= gi1; ++gi) {
        uint32_t c = ld(carried + (size_t)gi * Tp + t);
        if (gi == gi0) c &= 0xffffffffu << (lo & 31);
        if (gi == gi1) c &= 0xffffffffu >> (31 - ((hi - 1) & 31));
        if (c == 0) continue;
        const uint32_t tw = ld(two + (size_t)gi * Tp + t);
        while (c) {
            const int b = ctz32(c);
            c &= c - 1;
            const acc_t p = (acc_t)(ldpos((gi << 5) + b) - p0);
            const int cls = (int)((tw >> b) & 1u);
            while (q > 0) {
                const acc_t pi = wait_p[head];
                if (p - pi < kMinPairDistance) break;
                const acc_t vv = wait_v[head];
                const acc_t v = vv >> 1;
                if (vv & 1) { A1 = imax(A1, v - pi); B1 = imax(B1, v); }
                else        { A0 = imax(A0, v - pi); B0 = imax(B0, v); }
                head = (head + 1) & (kRingSlots - 1);
                --q;
            }
            const acc_t same = cls ? A1 : A0;
            const acc_t other = cls ? B0 : B1;
            const acc_t m = imax(same + (p + bonus), other + pen_eff);
            top = imax(top, m);
            const int tail = (head + q) & (kRingSlots - 1);
            wait_p[tail] = p;
            wait_v[tail] = (imax(m, (acc_t)0) << 1) | (acc_t)cls;
            ++q;
            ++n;
        }
    }
    n_out = n;
    score_out = (n <= 2 || top <= THR) ? kNegInf64 : (int64_t)top;
}

// ---------------------------------------------------------------------------------------------
// Walk DP, register form: the same walk with list_dp_run's waiting-site scheme -- the previous site
// waits in registers (the common case: carried sites >= 10 bp apart are committed straight from
// there), only crowded sites are parked in the thread-local ring.  For planes that sit in shared
// memory (the fused kernel), where the loads are cheap and the local-memory ring of walk_dp_unit
// was the latency.
// ---------------------------------------------------------------------------------------------
template <typename acc_t, typename LoadU32, typename LoadPos>
SSTAR_HD void walk_dp_unit_reg(LoadU32 ld, LoadPos ldpos, const uint32_t *carried, const uint32_t *two, int32_t Tp,
                               int t, int32_t lo, int32_t hi, int32_t bonus_i, int32_t penalty_i, int32_t max_mm,
                               int &n_out, int64_t &score_out) {
    constexpr acc_t NEG = Sentinel<acc_t>::NEG, THR = Sentinel<acc_t>::THR;
    const acc_t bonus = (acc_t)bonus_i;
    const acc_t pen_eff = (max_mm >= 1) ? (acc_t)penalty_i : NEG;
    acc_t A0 = NEG, A1 = NEG, B0 = NEG, B1 = NEG, top = NEG;
    acc_t ring_p[kRingSlots], ring_v[kRingSlots];  // parked sites older than the previous one: p, v << 1 | class
    acc_t pp = 0, pv = 0;                          // the previous site
    int k = 0, c = 0;                              // sites seen, next site to commit
    const auto p0 = ldpos(lo);  // (a loader returning uint32 keeps the position arithmetic 32-bit: differences
                                //  inside a window are never negative and below 2^32)
    const int gi0 = lo >> 5, gi1 = (hi - 1) >> 5;
    for (int gi = gi0; gi <= gi1; ++gi) {
        uint32_t cb = ld(carried + (size_t)gi * Tp + t);
        if (gi == gi0) cb &= 0xffffffffu << (lo & 31);
        if (gi == gi1) cb &= 0xffffffffu >> (31 - ((hi - 1) & 31));
        if (cb == 0) continue;
        const uint32_t tw = ld(two + (size_t)gi * Tp + t);
        while (cb) {
            const int b = ctz32(cb);
            cb &= cb - 1;
            const acc_t p = (acc_t)(ldpos((gi << 5) + b) - p0);
            const int cls = (int)((tw >> b) & 1u);
            if (c == k - 1) {  // common case: only the previous site waits, in registers
                if (p - pp >= kMinPairDistance) {
                    const acc_t v = pv >> 1;
                    if (pv & 1) { A1 = imax(A1, v - pp); B1 = imax(B1, v); }
                    else        { A0 = imax(A0, v - pp); B0 = imax(B0, v); }
                    c = k;
                } else {
                    ring_p[(k - 1) & (kRingSlots - 1)] = pp;  // it keeps waiting: park it
                    ring_v[(k - 1) & (kRingSlots - 1)] = pv;
                }
            } else if (c < k) {  // crowded sites: older ones wait in the ring
                while (c < k) {
                    acc_t pc, vc;
                    if (c == k - 1) { pc = pp; vc = pv; }
                    else            { pc = ring_p[c & (kRingSlots - 1)]; vc = ring_v[c & (kRingSlots - 1)]; }
                    if (p - pc < kMinPairDistance) break;
                    const acc_t v = vc >> 1;
                    if (vc & 1) { A1 = imax(A1, v - pc); B1 = imax(B1, v); }
                    else        { A0 = imax(A0, v - pc); B0 = imax(B0, v); }
                    ++c;
                }
                if (c < k) {
                    ring_p[(k - 1) & (kRingSlots - 1)] = pp;
                    ring_v[(k - 1) & (kRingSlots - 1)] = pv;
                }
            }
            const acc_t same = cls ? A1 : A0;
            const acc_t other = cls ? B0 : B1;
            const acc_t m = imax(same + (p + bonus), other + pen_eff);
            top = imax(top, m);
            pp = p;
            pv = (imax(m, (acc_t)0) << 1) | (acc_t)cls;
            ++k;
        }
    }
    n_out = k;
    score_out = (k <= 2 || top <= THR) ? kNegInf64 : (int64_t)top;
}

// ---------------------------------------------------------------------------------------------
// General-genotype walk DP: walk_dp_unit for ANY int8 genotype (missing calls, multi-allelic
// dosages), straight from a carried plane plus the genotype bytes; int64 accumulators.  Per-class
// running maxima as in list_dp_run_general (<= kMaxClasses distinct values per unit, else false:
// the caller takes the class-table DP below).  `ldgt(k)` returns the genotype byte of row k.
// ---------------------------------------------------------------------------------------------
template <typename LoadU32, typename LoadPos, typename LoadGt>
SSTAR_HD_COLD bool walk_dp_unit_general(LoadU32 ld, LoadPos ldpos, LoadGt ldgt, const uint32_t *carried, int32_t Tp, int t,
                                        int32_t lo, int32_t hi, int32_t bonus, int32_t penalty, int32_t max_mm,
                                        int &n_out, int64_t &score_out) {
    constexpr int64_t NEG = Sentinel<int64_t>::NEG, THR = Sentinel<int64_t>::THR;
    int64_t A[kMaxClasses], B[kMaxClasses];
    int32_t val[kMaxClasses];
    int64_t wait_p[kRingSlots], wait_v[kRingSlots];
    int8_t wait_c[kRingSlots];
    int K = 0, q = 0, head = 0, n = 0;
    int64_t top = NEG;
    const int64_t p0 = ldpos(lo);
    const int gi0 = lo >> 5, gi1 = (hi - 1) >> 5;
    for (int gi = gi0; gi <= gi1; ++gi) {
        uint32_t c = ld(carried + (size_t)gi * Tp + t);
        if (gi == gi0) c &= 0xffffffffu << (lo & 31);
        if (gi == gi1) c &= 0xffffffffu >> (31 - ((hi - 1) & 31));
        while (c) {
            const int b = ctz32(c);
            c &= c - 1;
            const int32_t row = (gi << 5) + b;
            const int64_t p = ldpos(row) - p0;
            const int32_t g = (int32_t)ldgt(row);
            while (q > 0) {  // sites >= 10 bp behind become eligible predecessors
                const int64_t pi = wait_p[head];
                if (p - pi < kMinPairDistance) break;
                const int ci = wait_c[head];
                const int64_t v = wait_v[head];
                A[ci] = imax(A[ci], v - pi);
                B[ci] = imax(B[ci], v);
                head = (head + 1) & (kRingSlots - 1);
                --q;
            }
            int idx = -1;
            for (int i = 0; i < K; ++i)
                if (val[i] == g) idx = i;
            if (idx < 0) {
                if (K == kMaxClasses) return false;
                idx = K++;
                val[idx] = g; A[idx] = NEG; B[idx] = NEG;
            }
            int64_t other = NEG;
            for (int i = 0; i < K; ++i) {
                const int32_t d = val[i] > g ? val[i] - g : g - val[i];
                if (i != idx && d <= max_mm) other = imax(other, B[i]);
            }
            const int64_t m = imax(A[idx] + (p + bonus), other + penalty);
            top = imax(top, m);
            const int tail = (head + q) & (kRingSlots - 1);
            wait_p[tail] = p;
            wait_v[tail] = imax(m, (int64_t)0);
            wait_c[tail] = (int8_t)idx;
            ++q;
            ++n;
        }
    }
    n_out = n;
    score_out = (n <= 2 || top <= THR) ? kNegInf64 : top;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Class-table DP: the same prefix-max recurrence with the running maxima in two tables indexed
// DIRECTLY by the genotype byte (256 entries each, caller-provided memory: shared memory on the
// device) -- any number of distinct genotype values, any window length, bounded state.  Sites are
// fed one at a time in ascending position (`site`), so the caller can stream a window of any size.
// The last resort of the fused kernel (a unit with more than kMaxClasses values, a window larger
// than its tile); int64 throughout.
// ---------------------------------------------------------------------------------------------
constexpr int kClassTableSize = 256;

struct StreamDP {
    int64_t *A, *B;            // [kClassTableSize], index = genotype byte + 128
    int64_t *wait_p, *wait_v;  // [kRingSlots]
    int32_t *wait_g;           // [kRingSlots]
    int head, q, n;
    int64_t top;

    SSTAR_HD void begin() { head = 0; q = 0; n = 0; top = Sentinel<int64_t>::NEG; }  // (tables reset by the caller)

    SSTAR_HD void site(int64_t p, int32_t g, int32_t bonus, int32_t penalty, int32_t max_mm) {
        constexpr int64_t NEG = Sentinel<int64_t>::NEG;
        while (q > 0) {
            const int64_t pi = wait_p[head];
            if (p - pi < kMinPairDistance) break;
            const int ci = wait_g[head] + 128;
            const int64_t v = wait_v[head];
            A[ci] = imax(A[ci], v - pi);
            B[ci] = imax(B[ci], v);
            head = (head + 1) & (kRingSlots - 1);
            --q;
        }
        int64_t other = NEG;
        const int64_t g_lo = imax((int64_t)g - (int64_t)imax(max_mm, (int32_t)0), (int64_t)-128);
        const int64_t g_hi = -imax(-((int64_t)g + (int64_t)imax(max_mm, (int32_t)0)), (int64_t)-127);
        for (int64_t x = g_lo; x <= g_hi; ++x)
            if (x != g) other = imax(other, B[x + 128]);
        const int64_t m = imax(A[g + 128] + (p + bonus), other + penalty);
        top = imax(top, m);
        const int tail = (head + q) & (kRingSlots - 1);
        wait_p[tail] = p;
        wait_v[tail] = imax(m, (int64_t)0);
        wait_g[tail] = g;
        ++q;
        ++n;
    }

    SSTAR_HD int64_t result() const {
        return (n <= 2 || top <= Sentinel<int64_t>::THR) ? kNegInf64 : top;
    }
};

// ---------------------------------------------------------------------------------------------
// Walk DP with backpointers: the sites on the unit's best chain (the legacy S*_SNP_number / S*_SNPs
// columns) out of the O(n) prefix-max form.  The rule on the literal
// recurrence -- backpointer of j = the FIRST i attaining M[j]; the chain is extended there when
// M[i] >= 0, else it starts afresh with the pair (i, j); it ends at the FIRST j attaining the final
// max -- survives the four running maxima when each keeps the index of its first maximiser (sites
// are committed in index order, updates are strict), and the two candidates of site j (same class,
// other class) tie-break to the smaller index.
// Per site the caller's arrays get: row[k] = its row relative to `lo` (bit 15: M[k] is valid and >= 0;
// windows of at most kChainMaxRows rows), back[k] = -1 (no valid predecessor) or (first << 1 | extended).  Returns false when the unit has more than kChainCap
// sites (the caller then takes the pairwise kernels).  Genotypes in {0,1,2} only.
// ---------------------------------------------------------------------------------------------
constexpr int kChainCap = 128;
constexpr int kChainMaxRows = 32767;  // rows per window the chain DP takes: row[] keeps bit 15 for "M >= 0"

template <typename acc_t, typename LoadU32, typename LoadPos>
SSTAR_HD bool walk_dp_unit_chain(LoadU32 ld, LoadPos ldpos, const uint32_t *carried, const uint32_t *two, int32_t Tp,
                                 int t, int32_t lo, int32_t hi, int32_t bonus_i, int32_t penalty_i, int32_t max_mm,
                                 uint16_t *row, int16_t *back, int &n_out, int &top_j_out, int64_t &score_out) {
    constexpr acc_t NEG = Sentinel<acc_t>::NEG, THR = Sentinel<acc_t>::THR;
    const acc_t bonus = (acc_t)bonus_i;
    const acc_t pen_eff = (max_mm >= 1) ? (acc_t)penalty_i : NEG;
    acc_t A0 = NEG, A1 = NEG, B0 = NEG, B1 = NEG;  // the four running maxima ...
    int iA0 = -1, iA1 = -1, iB0 = -1, iB1 = -1;    // ... and the index of the first site attaining each
    acc_t ring_p[kRingSlots], ring_v[kRingSlots];  // parked sites older than the previous one: p, v << 1 | class
    acc_t pp = 0, pv = 0;                          // the previous site (waits in registers)
    acc_t top = NEG;
    int k = 0, c = 0, top_j = -1;                  // sites seen, next site to commit
    const auto p0 = ldpos(lo);
    const int gi0 = lo >> 5, gi1 = (hi - 1) >> 5;
    auto commit = [&](int idx, acc_t pc, acc_t vc) {  // site idx becomes an eligible predecessor (index order, strict updates)
        const acc_t v = vc >> 1;
        if (vc & 1) { if (v - pc > A1) { A1 = v - pc; iA1 = idx; } if (v > B1) { B1 = v; iB1 = idx; } }
        else        { if (v - pc > A0) { A0 = v - pc; iA0 = idx; } if (v > B0) { B0 = v; iB0 = idx; } }
    };
    for (int gi = gi0; gi <= gi1; ++gi) {
        uint32_t cb = ld(carried + (size_t)gi * Tp + t);
        if (gi == gi0) cb &= 0xffffffffu << (lo & 31);
        if (gi == gi1) cb &= 0xffffffffu >> (31 - ((hi - 1) & 31));
        if (cb == 0) continue;
        const uint32_t tw = ld(two + (size_t)gi * Tp + t);
        while (cb) {
            const int b = ctz32(cb);
            cb &= cb - 1;
            if (k >= kChainCap) return false;
            const int32_t r = (gi << 5) + b;
            const acc_t p = (acc_t)(ldpos(r) - p0);
            const int cls = (int)((tw >> b) & 1u);
            if (c == k - 1) {  // common case: only the previous site waits, in registers
                if (p - pp >= kMinPairDistance) { commit(k - 1, pp, pv); c = k; }
                else { ring_p[(k - 1) & (kRingSlots - 1)] = pp; ring_v[(k - 1) & (kRingSlots - 1)] = pv; }
            } else if (c < k) {  // crowded sites: older ones wait in the ring
                while (c < k) {
                    acc_t pc, vc;
                    if (c == k - 1) { pc = pp; vc = pv; }
                    else            { pc = ring_p[c & (kRingSlots - 1)]; vc = ring_v[c & (kRingSlots - 1)]; }
                    if (p - pc < kMinPairDistance) break;
                    commit(c, pc, vc);
                    ++c;
                }
                if (c < k) { ring_p[(k - 1) & (kRingSlots - 1)] = pp; ring_v[(k - 1) & (kRingSlots - 1)] = pv; }
            }
            const acc_t ca = (cls ? A1 : A0) + (p + bonus), cb2 = (cls ? B0 : B1) + pen_eff;
            const int ia = cls ? iA1 : iA0, ib = cls ? iB0 : iB1;
            const acc_t m = imax(ca, cb2);
            const bool valid = m > THR;
            int first = -1;
            if (valid) first = ca > cb2 ? ia : (cb2 > ca ? ib : (ia < ib ? ia : ib));
            // back: -1 (no valid predecessor) or first << 1 | (M[first] >= 0: the chain is extended there)
            back[k] = first < 0 ? (int16_t)-1 : (int16_t)((first << 1) | (int)(row[first] >> 15));
            row[k] = (uint16_t)((uint32_t)(r - lo) | ((valid && m >= 0) ? 0x8000u : 0u));
            if (valid && m > top) { top = m; top_j = k; }
            pp = p;
            pv = (imax(m, (acc_t)0) << 1) | (acc_t)cls;
            ++k;
        }
    }
    n_out = k;
    top_j_out = (k <= 2) ? -1 : top_j;
    score_out = (k <= 2 || top_j < 0) ? kNegInf64 : (int64_t)top;
    return true;
}

// sites on the chain that ends at top_j (0 when there is none)
SSTAR_HD int chain_length(const int16_t *back, int top_j) {
    int cnt = 0;
    for (int j = top_j; j >= 0;) {
        ++cnt;
        const int b = back[j];
        if (b < 0) break;            // (cannot happen for a site with a valid M: kept for safety)
        if (b & 1) j = b >> 1;       // extension of the chain ending at its first maximiser
        else { ++cnt; break; }       // fresh start with the pair (first, j)
    }
    return cnt;
}

// rows (relative to lo) of the chain's sites, ascending, into out[0 .. len)
template <typename Out>
SSTAR_HD void chain_rows(const uint16_t *row, const int16_t *back, int top_j, int len, Out out) {
    int k = len - 1;
    for (int j = top_j; j >= 0 && k >= 0;) {
        out(k--, (uint16_t)(row[j] & 0x7fffu));
        const int b = back[j];
        if (b < 0) break;
        if (b & 1) j = b >> 1;
        else { if (k >= 0) out(k--, (uint16_t)(row[b >> 1] & 0x7fffu)); break; }
    }
}

}  // namespace sstar

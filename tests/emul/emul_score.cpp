// Host emulation of score_group_kernel's control flow around the REAL DP core (dp_core.cuh,
// compiled here for the host).  TEST INFRASTRUCTURE: lets `pytest -m "not gpu"` check the device
// arithmetic against the oracle without a GPU.  One "thread" = one column; the planes, the group
// logic and the list build mirror sstar_b200/csrc/{prep,score}_kernels.cuh in scalar form.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "../../sstar_b200/csrc/dp_core.cuh"

using namespace sstar;

namespace {
struct HostLd { uint32_t operator()(const uint32_t *p) const { return *p; } };
struct HostPos {
    const int32_t *pos;
    int64_t operator()(int32_t k) const { return (int64_t)pos[k]; }
};
int32_t lower_bound_i64(const int32_t *pos, int64_t V, int64_t x) {
    int64_t lo = 0, hi = V;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((int64_t)pos[mid] < x) lo = mid + 1; else hi = mid;
    }
    return (int32_t)lo;
}
}  // namespace

extern "C" int emul_score_windows(const int8_t *ref, const int8_t *tgt, const int32_t *pos, int64_t V, int32_t R,
                                  int32_t T, const int32_t *ws, const int32_t *we, int32_t W, int32_t bonus,
                                  int32_t max_mm, int32_t penalty, int32_t wg, int32_t L, int32_t words_cap,
                                  int64_t *score, int32_t *nsnp, uint8_t *slow /*[W]*/, int32_t *path /*[W,T]*/) {
    const int64_t G = (V + 31) / 32;
    const int32_t Tp = (T + 31) / 32 * 32;
    std::vector<uint32_t> absent(G > 0 ? G : 1, 0), exotic(G > 0 ? G : 1, 0);
    std::vector<uint32_t> carried((size_t)(G > 0 ? G : 1) * Tp, 0), two((size_t)(G > 0 ? G : 1) * Tp, 0);
    for (int64_t k = 0; k < V; ++k) {
        int sum = 0;
        for (int r = 0; r < R; ++r) sum += ref[k * R + r];
        if (sum != 0) continue;
        absent[k >> 5] |= 1u << (k & 31);
        for (int t = 0; t < T; ++t) {
            const int b = tgt[k * T + t];
            if (b != 0) carried[(size_t)(k >> 5) * Tp + t] |= 1u << (k & 31);
            if (b == 2) two[(size_t)(k >> 5) * Tp + t] |= 1u << (k & 31);
            // group level as the pack kernel raises it: 1 = other genotypes, all in {-2,-1} (the planes then
            // hold |g| == 2 and the sign, for the whole group); 2 = anything else (the bytes are read)
            if ((unsigned)b > 2u) exotic[k >> 5] |= (b == -1 || b == -2) ? 1u : 2u;
        }
    }
    std::vector<uint32_t> neg((size_t)(G > 0 ? G : 1) * Tp, 0xdeadbeefu);  // (unwritten where the group is clean)
    for (int64_t k = 0; k < V; ++k) {
        if (!exotic[k >> 5]) continue;
        const uint32_t bit = 1u << (k & 31);
        for (int t = 0; t < T; ++t) {
            uint32_t &nw = neg[(size_t)(k >> 5) * Tp + t], &tw = two[(size_t)(k >> 5) * Tp + t];
            if ((k & 31) == 0) nw = 0;
            const int b = tgt[k * T + t];
            if (!(absent[k >> 5] & bit)) continue;
            if (b < 0) nw |= bit;
            if (b == -2) tw |= bit;
        }
    }
    const uint32_t compat = signed_compat_mask(max_mm);
    std::vector<int32_t> wlo(W), whi(W);
    for (int w = 0; w < W; ++w) {
        wlo[w] = lower_bound_i64(pos, V, ws[w]);
        whi[w] = std::max(wlo[w], lower_bound_i64(pos, V, (int64_t)we[w] + 1));
    }
    memset(slow, 0, W);
    const int32_t pen_eff = (max_mm >= 1) ? penalty : Sentinel<int32_t>::NEG;
    std::vector<uint32_t> list(L + 1), spos;
    auto window_absent = [&](int32_t lo, int32_t hi, int32_t &na, uint32_t &ex) {
        na = 0; ex = 0;
        if (hi <= lo) return;
        for (int gi = lo >> 5; gi <= ((hi - 1) >> 5); ++gi) {
            uint32_t rm = 0xffffffffu;
            if (gi == (lo >> 5)) rm &= 0xffffffffu << (lo & 31);
            if (gi == ((hi - 1) >> 5)) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
            na += __builtin_popcount(absent[gi] & rm);
            ex |= exotic[gi];
        }
    };
    for (int w0 = 0; w0 < W; w0 += wg) {
        const int nwin = std::min(wg, W - w0);
        // greedy sub-groups, as in score_group_kernel
        int s0 = 0;
        while (s0 < nwin) {
            int32_t rlo = INT32_MAX, rhi = -1, prev_lo = -1, prev_hi = -1, prev_pf = INT32_MIN, maxnv = 0, p0 = 0, plast = INT32_MIN;
            bool mono_hi = true, alone = false;
            int s1 = s0;
            for (; s1 < nwin; ++s1) {
                const int32_t lo = wlo[w0 + s1], hi = whi[w0 + s1];
                if (hi <= lo) continue;
                const int32_t pf = pos[lo], pl = pos[hi - 1];
                const int32_t n_rlo = std::min(rlo, lo), n_rhi = std::max(rhi, hi);
                const int32_t n_p0 = lo < rlo ? pf : p0, n_plast = std::max(plast, pl), n_maxnv = std::max(maxnv, hi - lo);
                const bool fits = lo >= prev_lo && pf >= prev_pf && pl >= pf &&
                                  ((n_rhi - 1) >> 5) - (n_rlo >> 5) + 1 <= words_cap &&
                                  score_bound((int64_t)n_plast - (int64_t)n_p0, n_maxnv, bonus, penalty) < kSmallBound;
                if (!fits) {
                    if (rhi < 0) { alone = true; ++s1; }
                    break;
                }
                mono_hi = mono_hi && hi >= prev_hi;
                prev_lo = lo; prev_hi = hi; prev_pf = pf;
                rlo = n_rlo; rhi = n_rhi; p0 = n_p0; plast = n_plast; maxnv = n_maxnv;
            }
            if (alone) {
                for (int i = s0; i < s1; ++i) {
                    const int w = w0 + i;
                    const int32_t lo = wlo[w], hi = whi[w];
                    int32_t na; uint32_t ex;
                    window_absent(lo, hi, na, ex);
                    if (ex) { slow[w] = 1; continue; }
                    for (int t = 0; t < T; ++t) {
                        int n = 0; int64_t sc = kNegInf64; int pth = 0;
                        if (hi > lo) {
                            const int64_t wspan = (int64_t)pos[hi - 1] - (int64_t)pos[lo];
                            if (score_bound(wspan, hi - lo, bonus, penalty) < kSmallBound) {
                                walk_dp_unit<int32_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, n, sc);
                                pth = 3;
                            } else {
                                walk_dp_unit<int64_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, n, sc);
                                pth = 4;
                            }
                        }
                        score[(size_t)w * T + t] = sc; nsnp[(size_t)w * T + t] = (hi - lo) - na + n; path[(size_t)w * T + t] = pth;
                    }
                }
            } else if (rhi < 0) {
                for (int i = s0; i < s1; ++i)
                    for (int t = 0; t < T; ++t) {
                        score[(size_t)(w0 + i) * T + t] = kNegInf64; nsnp[(size_t)(w0 + i) * T + t] = 0; path[(size_t)(w0 + i) * T + t] = 0;
                    }
            } else {
                const int gi0 = rlo >> 5, nwr = ((rhi - 1) >> 5) - gi0 + 1;
                const int64_t span = (int64_t)plast - p0;
                const int32_t base = gi0 << 5;
                const int32_t npos = (int32_t)std::min<int64_t>((int64_t)nwr << 5, V - base);
                spos.assign((size_t)nwr << 5, 0);
                for (int i = 0; i < npos; ++i) spos[i] = (uint32_t)(pos[base + i] - p0);
                bool any_ex = false;
                for (int i = s0; i < s1; ++i) {
                    int32_t na; uint32_t ex;
                    window_absent(wlo[w0 + i], whi[w0 + i], na, ex);
                    any_ex = any_ex || ex;
                }
                const bool gen_span = any_ex && span < (int64_t(1) << kGeneralPosBits);
                uint32_t ex_or = 0;
                for (int i = s0; i < s1; ++i) {
                    int32_t na; uint32_t ex;
                    window_absent(wlo[w0 + i], whi[w0 + i], na, ex);
                    ex_or |= ex;
                }
                const bool sgn = ex_or == 1u;  // every other genotype of the sub-group is -1 / -2: signed entries from the planes
                const bool gen = !sgn && gen_span;
                for (int t = 0; t < T; ++t) {
                    int cnt = 0;
                    bool col_neg = false;
                    for (int j = 0; j < nwr; ++j) {
                        uint32_t c = carried[(size_t)(gi0 + j) * Tp + t];
                        if (j == 0) c &= 0xffffffffu << (rlo & 31);
                        if (j == nwr - 1) c &= 0xffffffffu >> (31 - ((rhi - 1) & 31));
                        const uint32_t tw = two[(size_t)(gi0 + j) * Tp + t];
                        while (c) {
                            const int b = __builtin_ctz(c);
                            c &= c - 1;
                            uint32_t e = (spos[(j << 5) + b] << 1) | ((tw >> b) & 1u);
                            if (gen) e = (spos[(j << 5) + b] << 8) | (uint32_t)(uint8_t)tgt[(size_t)(base + (j << 5) + b) * T + t];
                            if (sgn) {  // signed entries from the three planes (the sign plane only where the group is flagged)
                                const uint32_t ngw = exotic[gi0 + j] ? neg[(size_t)(gi0 + j) * Tp + t] : 0u;
                                const uint32_t t2 = (tw >> b) & 1u, sg = (ngw >> b) & 1u;
                                const int gb = tgt[(size_t)(base + (j << 5) + b) * T + t];
                                if (gb != (sg ? -(int)(1 + t2) : (int)(1 + t2))) return -8;  // the planes reproduce the byte
                                e = (spos[(j << 5) + b] << 2) | (sg << 1) | t2;
                                col_neg = col_neg || sg;
                            }
                            if (cnt < L) list[cnt] = e;
                            ++cnt;
                        }
                    }
                    const bool overflow = cnt >= L;  // one slot is the terminator
                    if (!overflow) list[cnt] = kListSentinel;
                    int cur = 0, cur_b = 0, odd_a = 0, odd_b = 0;
                    const bool balanced = (s1 - s0) > 1 && mono_hi;  // the kernel's two-pass path
                    for (int i = s0; i < s1; ++i) {
                        const int w = w0 + i;
                        const int32_t lo = wlo[w], hi = whi[w];
                        int32_t na; uint32_t ex;
                        window_absent(lo, hi, na, ex);
                        if (ex && !gen && !sgn) { slow[w] = 1; continue; }
                        int n = 0; int64_t sc = kNegInf64;
                        if (hi > lo) {
                            const uint32_t plo = (uint32_t)(pos[lo] - p0), phi = (uint32_t)(pos[hi - 1] - p0);
                            if (gen) {
                                if (overflow) { slow[w] = 1; continue; }
                                if (!balanced) cur_b = cur;
                                if (balanced) list_window_range_general(list.data(), 1, cur, cur_b, odd_a, odd_b, plo, phi);
                                else list_window_range_shifted<8>(list.data(), 1, cur, cur_b, plo, phi, false);
                                n = cur_b - cur;
                                // as on the device: in the balanced pass the units whose genotypes all are 1 / 2 (the
                                // cursors' counts of other genotypes agree) take the four-maxima DP on general
                                // entries, the others the per-class DP
                                if (balanced && (odd_a == odd_b) != general_run_is_plain(list.data() + cur, 1, n)) return -7;
                                if (n > 2 && balanced && odd_a == odd_b)
                                    sc = list_dp_run<8>(list.data() + cur, 1, n, bonus, pen_eff);
                                else if (n > 2 && !list_dp_run_general(list.data() + cur, 1, n, bonus, penalty, max_mm, sc)) { slow[w] = 1; continue; }
                            } else if (sgn) {
                                // a column without a negative site runs the clean body on signed entries; the others
                                // count the negative entries their cursors pass (balanced pass) or take the four-class DP
                                if (overflow && col_neg) { slow[w] = 1; continue; }
                                if (overflow) walk_dp_unit<int32_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, n, sc);
                                else if (balanced) {
                                    if (col_neg) list_window_range_signed(list.data(), 1, cur, cur_b, odd_a, odd_b, plo, phi);
                                    else list_window_range_shifted<kSignedShift>(list.data(), 1, cur, cur_b, plo, phi, true);
                                    n = cur_b - cur;
                                    if (n > 2 && odd_a == odd_b) sc = list_dp_run<kSignedShift>(list.data() + cur, 1, n, bonus, pen_eff);
                                    else if (n > 2) sc = list_dp_run_signed(list.data() + cur, 1, n, bonus, penalty, compat);
                                } else {
                                    cur_b = cur;
                                    list_window_range_shifted<kSignedShift>(list.data(), 1, cur, cur_b, plo, phi, false);
                                    n = cur_b - cur;
                                    if (n > 2 && col_neg) sc = list_dp_run_signed(list.data() + cur, 1, n, bonus, penalty, compat);
                                    else if (n > 2) sc = list_dp_run<kSignedShift>(list.data() + cur, 1, n, bonus, pen_eff);
                                }
                            } else if (!overflow && balanced) {
                                list_window_range(list.data(), 1, cur, cur_b, plo, phi, true);
                                n = cur_b - cur;
                                if (n > 2) sc = list_dp_run(list.data() + cur, 1, n, bonus, pen_eff);
                            } else if (!overflow) list_dp_window(list.data(), 1, cur, plo, phi, bonus, pen_eff, n, sc);
                            else walk_dp_unit<int32_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, n, sc);
                        }
                        score[(size_t)w * T + t] = sc;
                        nsnp[(size_t)w * T + t] = (hi - lo) - na + n;
                        path[(size_t)w * T + t] = sgn ? 6 : gen ? 5 : overflow ? 2 : 1;
                    }
                }
            }
            s0 = s1;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Scalar emulation of the fused kernel's unit paths (fused_kernels.cuh): the clean walk DP, the
// general walk DP (<= kMaxClasses genotype values) and the class-table streaming DP behind it.
// (score_group emulation above: 1 list, 2 walk after a list overflow, 3 / 4 walk alone, 5 general bytes, 6 small signed genotypes from the planes)
// path: 0 empty window, 1 walk int32, 2 walk int64, 3 general walk, 4 class-table stream.
// force_stream != 0 sends every unit through the streaming DP (the oversize-window route).
// ---------------------------------------------------------------------------------------------
namespace {
struct HostGt {
    const int8_t *tgt; int32_t T, t;
    int operator()(int32_t k) const { return (int)tgt[(size_t)k * T + t]; }
};
}  // namespace

extern "C" int emul_fused_windows(const int8_t *ref, const int8_t *tgt, const int32_t *pos, int64_t V, int32_t R,
                                  int32_t T, const int32_t *ws, const int32_t *we, int32_t W, int32_t bonus,
                                  int32_t max_mm, int32_t penalty, int32_t force_stream, int64_t *score, int32_t *nsnp,
                                  int32_t *path) {
    const int64_t G = (V + 31) / 32;
    const int32_t Tp = (T + 31) / 32 * 32;
    std::vector<uint32_t> absent(G > 0 ? G : 1, 0), exotic(G > 0 ? G : 1, 0);
    std::vector<uint32_t> carried((size_t)(G > 0 ? G : 1) * Tp, 0), two((size_t)(G > 0 ? G : 1) * Tp, 0);
    for (int64_t k = 0; k < V; ++k) {
        int sum = 0;
        for (int r = 0; r < R; ++r) sum += ref[k * R + r];
        if (sum != 0) continue;
        absent[k >> 5] |= 1u << (k & 31);
        for (int t = 0; t < T; ++t) {
            const int b = tgt[k * T + t];
            if (b != 0) carried[(size_t)(k >> 5) * Tp + t] |= 1u << (k & 31);
            if (b == 2) two[(size_t)(k >> 5) * Tp + t] |= 1u << (k & 31);
            if ((unsigned)b > 2u) exotic[k >> 5] = 1;
        }
    }
    std::vector<int64_t> A(kClassTableSize), B(kClassTableSize), wp(kRingSlots), wv(kRingSlots);
    std::vector<int32_t> wgt(kRingSlots);
    for (int w = 0; w < W; ++w) {
        const int32_t lo = lower_bound_i64(pos, V, ws[w]);
        const int32_t hi = std::max(lo, lower_bound_i64(pos, V, (int64_t)we[w] + 1));
        int32_t na = 0; uint32_t ex = 0;
        if (hi > lo)
            for (int gi = lo >> 5; gi <= ((hi - 1) >> 5); ++gi) {
                uint32_t rm = 0xffffffffu;
                if (gi == (lo >> 5)) rm &= 0xffffffffu << (lo & 31);
                if (gi == ((hi - 1) >> 5)) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                na += __builtin_popcount(absent[gi] & rm);
                ex |= exotic[gi];
            }
        for (int t = 0; t < T; ++t) {
            int n = 0; int64_t sc = kNegInf64; int pth = 0;
            bool done = hi <= lo;
            if (!done && !force_stream) {
                if (ex) {
                    done = walk_dp_unit_general(HostLd(), HostPos{pos}, HostGt{tgt, T, t}, carried.data(), Tp, t, lo, hi, bonus,
                                                penalty, max_mm, n, sc);
                    pth = 3;
                } else {
                    const int64_t span = (int64_t)pos[hi - 1] - (int64_t)pos[lo];
                    if (score_bound(span, hi - lo, bonus, penalty) < kSmallBound) {
                        walk_dp_unit_reg<int32_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, n, sc);
                        pth = 1;
                    } else {
                        walk_dp_unit_reg<int64_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, n, sc);
                        pth = 2;
                    }
                    done = true;
                }
            }
            if (!done) {  // class tables indexed by the genotype byte, rows streamed in order
                StreamDP dp;
                dp.A = A.data(); dp.B = B.data(); dp.wait_p = wp.data(); dp.wait_v = wv.data(); dp.wait_g = wgt.data();
                std::fill(A.begin(), A.end(), Sentinel<int64_t>::NEG);
                std::fill(B.begin(), B.end(), Sentinel<int64_t>::NEG);
                dp.begin();
                const int64_t p0 = pos[lo];
                for (int32_t k = lo; k < hi; ++k) {
                    if (!((absent[k >> 5] >> (k & 31)) & 1u)) continue;
                    const int g = tgt[(size_t)k * T + t];
                    if (g != 0) dp.site((int64_t)pos[k] - p0, g, bonus, penalty, max_mm);
                }
                n = dp.n; sc = dp.result(); pth = 4;
            }
            score[(size_t)w * T + t] = sc; nsnp[(size_t)w * T + t] = (hi - lo) - na + n; path[(size_t)w * T + t] = pth;
        }
    }
    return 0;
}


// ---------------------------------------------------------------------------------------------
// The O(n) chain DP with backpointers (dp_core.cuh: walk_dp_unit_chain) per unit: score, chain
// length and the positions on the chain.  fast[u] = 0 when the unit has more than kChainCap sites.
// ---------------------------------------------------------------------------------------------
extern "C" int emul_chain_windows(const int8_t *ref, const int8_t *tgt, const int32_t *pos, int64_t V, int32_t R,
                                  int32_t T, const int32_t *ws, const int32_t *we, int32_t W, int32_t bonus,
                                  int32_t max_mm, int32_t penalty, int64_t *score, int32_t *set_len,
                                  int32_t *set_pos /*[W*T*kChainCap]*/, uint8_t *fast) {
    const int64_t G = (V + 31) / 32;
    const int32_t Tp = (T + 31) / 32 * 32;
    std::vector<uint32_t> carried((size_t)(G > 0 ? G : 1) * Tp, 0), two((size_t)(G > 0 ? G : 1) * Tp, 0);
    for (int64_t k = 0; k < V; ++k) {
        int sum = 0;
        for (int r = 0; r < R; ++r) sum += ref[k * R + r];
        if (sum != 0) continue;
        for (int t = 0; t < T; ++t) {
            const int b = tgt[k * T + t];
            if (b != 0) carried[(size_t)(k >> 5) * Tp + t] |= 1u << (k & 31);
            if (b == 2) two[(size_t)(k >> 5) * Tp + t] |= 1u << (k & 31);
        }
    }
    uint16_t row[kChainCap];
    int16_t back[kChainCap];
    for (int w = 0; w < W; ++w) {
        const int32_t lo = lower_bound_i64(pos, V, ws[w]);
        const int32_t hi = std::max(lo, lower_bound_i64(pos, V, (int64_t)we[w] + 1));
        for (int t = 0; t < T; ++t) {
            const size_t u = (size_t)w * T + t;
            score[u] = kNegInf64; set_len[u] = 0; fast[u] = 1;
            if (hi <= lo) continue;
            if (hi - lo > kChainMaxRows) { fast[u] = 0; continue; }
            int n = 0, top_j = -1;
            int64_t sc = kNegInf64;
            const int64_t span = (int64_t)pos[hi - 1] - (int64_t)pos[lo];
            bool ok;
            if (score_bound(span, hi - lo, bonus, penalty) < kSmallBound)
                ok = walk_dp_unit_chain<int32_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, row, back, n, top_j, sc);
            else
                ok = walk_dp_unit_chain<int64_t>(HostLd(), HostPos{pos}, carried.data(), two.data(), Tp, t, lo, hi, bonus, penalty, max_mm, row, back, n, top_j, sc);
            if (!ok) { fast[u] = 0; continue; }
            score[u] = sc;
            const int len = chain_length(back, top_j);
            set_len[u] = len;
            int32_t *dst = set_pos + u * kChainCap;
            chain_rows(row, back, top_j, len, [&](int k, uint16_t r) { dst[k] = pos[lo + r]; });
        }
    }
    return 0;
}

// sstar_b200 -- chain-DP kernels: K2 (prefix-max, thread per unit) and K3 (pairwise, warp per unit).
#pragma once
#include "common.cuh"

namespace sstar {

// ------------------------------------------------------------------------------------------
// K2: prefix-max chain DP (dp_core.cuh), one THREAD per target column, one CTA per GROUP of up
// to kMaxGroup consecutive windows x <=128 columns.
//
// Overlapping windows (50 kb / 10 kb: 5x) share most of their variants, so the CTA turns the
// union variant range of a run of windows into per-column site lists ONCE -- each thread pulls
// its column's plane words (coalesced across the warp), and writes
// (relative position << 1 | genotype==2) for every carried reference-absent site into its own
// shared-memory column -- and then every window of the run is a cursor walk over that list: no
// bit scanning, no position lookups, no global loads inside the DP.
//
// The group is consumed as SUB-GROUPS: the longest runs of windows that ascend in variant index
// and position, whose union range fits the staged position tile and whose scores fit int32.
// Regular overlapping windows give one sub-group; disjoint or unordered ones give shorter runs; a
// window that does not fit even alone, and a column with more sites than list slots, are walked
// straight from the planes (walk_dp_unit).  Inside a sub-group of several windows the DP runs
// in two passes: every column locates its windows' runs with two forward-only cursors and
// finishes the units with n <= 2 (NaN), the rest are queued and shared evenly by the CTA's
// threads.  Genotypes outside {0,1,2} switch the sub-group to general entries and the per-class
// DP (out of line).  Results are identical on every path.
// ------------------------------------------------------------------------------------------
struct GroupShared {
    int32_t lo[kMaxGroup], hi[kMaxGroup];
    int32_t pfirst[kMaxGroup], plast[kMaxGroup];
    uint32_t plo[kMaxGroup], phi[kMaxGroup];
    int32_t nabs[kMaxGroup];
    uint32_t ex[kMaxGroup];
    int32_t red_nabs;
    uint32_t red_ex;
    int32_t n_work;
    int32_t n_work_back;  // general body: units holding other genotypes than 1 / 2, queued from the far end
    uint32_t exm[2];      // signed body: bit j = word j of the sub-group's union range belongs to a flagged group
};

// words between consecutive entries of one column's list: the CTA's thread count, or less when the
// panel is narrower than a warp (host and device must agree: the launch sizes the list area by it)
__host__ __device__ inline int score_list_stride(int T, int bd) { return T > 16 ? bd : (T > 8 ? 16 : 8); }

struct LdgU32 { __device__ __forceinline__ uint32_t operator()(const uint32_t *p) const { return __ldg(p); } };
struct LdgPos {
    const int32_t *pos;
    __device__ __forceinline__ int64_t operator()(int32_t k) const { return (int64_t)__ldg(pos + k); }
};

struct GroupCtx {
    GroupShared *g;
    uint32_t *s_list, *s_pos;
    int tid, bd, chunk, t, w0, nwin, L;  // w0 / nwin: the sub-group this call lists and scores
    int io;                              // its first window's slot in the GroupShared arrays
    int32_t rlo, rhi, gi0, nwr, pen_eff;
    bool mono_hi;
};

// List build + DP of one window group, in one of three modes.
//   0  genotypes in {0,1,2}: entries (relative position << 1 | genotype==2), the four-maxima DP.
//   2  some window holds other genotypes and they all are -1 / -2 (missing calls; group levels <= 1): signed
//      entries (relative position << 2 | sign << 1 | |g|==2) from the planes -- "|g| == 2" and the sign plane
//      the pack kernel writes for flagged groups, no genotype byte is fetched.  A column without a negative
//      site (nearly all of them when missing calls are sparse) runs the clean body on these entries; the
//      others count the negative entries their cursors pass, units without one still take the four-maxima
//      DP, the rest the four-class DP with the classes in registers (list_dp_run_signed).
//   1  any other byte: entries (relative position << 8 | genotype byte, fetched from the target matrix) and
//      the per-class DP with thread-local tables (list_dp_run_general).
template <int kMode, bool kNarrow = false>
__device__ __forceinline__ void group_body(const ScoreArgs &a, const GroupCtx &c) {
    constexpr bool kGen = kMode == 1;     // general entries, genotype bytes
    constexpr bool kSigned = kMode == 2;  // signed entries from three planes
    constexpr bool kTwoQueues = kGen || kSigned;
    GroupShared &g = *c.g;
    const int32_t *glo = g.lo + c.io, *ghi = g.hi + c.io, *gnabs = g.nabs + c.io;
    const uint32_t *gplo = g.plo + c.io, *gphi = g.phi + c.io, *gex = g.ex + c.io;
    uint32_t *s_list = c.s_list;
    uint32_t *s_work = c.s_pos;  // [wg * bd] unit worklist, once the lists are built
    const int tid = c.tid, bd = c.bd, chunk = c.chunk, t = c.t, w0 = c.w0, nwin = c.nwin, L = c.L;
    const int ls = kNarrow ? score_list_stride(a.T, bd) : bd;  // words between consecutive entries of a column's list
    constexpr int kShift = kGen ? 8 : (kSigned ? kSignedShift : 1);
    constexpr int kBatch = kSigned ? 4 : 8;  // plane words fetched per round trip (the signed body holds three planes' words: 8 would cost the kernel 80 registers)

    // this thread's column: plane words -> site list.  Shared memory is addressed with 32-bit
    // shared-window addresses (ld/st.shared) and the write cursor is clamped to a spill slot
    // (index L) instead of branching on overflow.
    if (kSigned) {
        // which words of the union range have a sign plane (flagged groups): a bit mask in shared memory, so
        // that the sign words are fetched next to the other planes' instead of behind a flag load each
        for (int j0 = tid & ~31; j0 < 64; j0 += bd) {
            const int j = j0 + (tid & 31);
            const uint32_t m = __ballot_sync(0xffffffffu, j < c.nwr && __ldg(a.exotic + c.gi0 + j) != 0u);
            if ((tid & 31) == 0) g.exm[j0 >> 5] = m;
        }
        __syncthreads();
    }
    int cnt = 0;
    bool col_neg = false;  // signed body: this column's list holds a negative genotype
    if (t < a.T) {
        const uint32_t *cp = a.carried + (size_t)c.gi0 * a.Tp + t;
        const uint32_t *tp = a.two + (size_t)c.gi0 * a.Tp + t;
        const uint32_t *np = a.neg + (size_t)c.gi0 * a.Tp + t;
        const uint32_t first_mask = 0xffffffffu << (c.rlo & 31);
        const uint32_t last_mask = 0xffffffffu >> (31 - ((c.rhi - 1) & 31));
        const uint32_t pos_addr = (uint32_t)__cvta_generic_to_shared(c.s_pos);
        const uint32_t step = (uint32_t)ls * 4u;
        uint32_t w_addr = (uint32_t)__cvta_generic_to_shared(s_list + tid);
        const uint32_t w_end = w_addr + (uint32_t)L * step;  // the spill slot
        const int nwr = c.nwr;
        uint32_t neg_any = 0;
        const uint32_t exm0 = kSigned ? g.exm[0] : 0u, exm1 = kSigned ? g.exm[1] : 0u;
        for (int jb = 0; jb < nwr; jb += kBatch) {
            uint32_t cw[kBatch], tw[kBatch], ng[kSigned ? kBatch : 1];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) cw[u] = (jb + u < nwr) ? __ldg(cp + (size_t)(jb + u) * a.Tp) : 0u;
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                if (jb + u == 0) cw[u] &= first_mask;
                if (jb + u == nwr - 1) cw[u] &= last_mask;
                tw[u] = (!kGen && jb + u < nwr) ? __ldg(tp + (size_t)(jb + u) * a.Tp) : 0u;  // not chained to cw
                // the sign plane exists only where the group is flagged (clean groups never write it)
                if (kSigned) {
                    const uint32_t has = (((jb + u) & 32) ? exm1 : exm0) >> ((jb + u) & 31);  // (no bit beyond nwr)
                    ng[u] = (has & 1u) ? __ldg(np + (size_t)(jb + u) * a.Tp) : 0u;
                    neg_any |= ng[u] & cw[u];
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                uint32_t cbits = cw[u];
                const uint32_t sp = pos_addr + ((uint32_t)(jb + u) << 7);
                cnt += __popc(cbits);
                while (cbits) {
                    const int b = __ffs(cbits) - 1;
                    cbits &= cbits - 1;
                    uint32_t e;
                    if (kGen) {
                        e = (uint32_t)(((jb + u) << 5) + b);  // row within the tile; completed below
                    } else {
                        uint32_t rel;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(rel) : "r"(sp + 4u * (uint32_t)b));
                        if (kSigned) e = (rel << 2) | (((ng[kSigned ? u : 0] >> b) & 1u) << 1) | ((tw[u] >> b) & 1u);
                        else e = (rel << 1) | ((tw[u] >> b) & 1u);
                    }
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(w_addr), "r"(e) : "memory");
                    w_addr = min(w_addr + step, w_end);
                }
            }
        }
        col_neg = neg_any != 0;
    }
    if (kGen && t < a.T) {
        // general entries: (relative position << 8 | genotype byte).  The bytes are fetched here,
        // in a loop of independent iterations (several loads in flight), not one exposed memory
        // latency per site inside the bit loop above.
        const int m = min(cnt, L);
        const int8_t *col = a.tgt + (size_t)(c.gi0 << 5) * a.T + t;
#pragma unroll 4
        for (int i = 0; i < m; ++i) {
            const uint32_t row = s_list[i * ls + tid];
            const int8_t gb = col[(size_t)row * a.T];
            s_list[i * ls + tid] = (c.s_pos[row] << 8) | (uint32_t)(uint8_t)gb;
        }
    }
    const bool overflow = cnt >= L;  // one slot is the terminator
    const uint32_t compat = kSigned ? signed_compat_mask(a.max_mm) : 0u;
    if (t < a.T && !overflow) s_list[cnt * ls + tid] = kListSentinel;
    if (tid == 0) { g.n_work = 0; g.n_work_back = 0; }
    __syncthreads();  // every list is built: the position tile becomes the worklist

    // a unit the general DP cannot take (more than kMaxClasses genotype values, or its column
    // overflowed the list): its window goes to the pairwise kernel, queued once
    auto bail_window = [&](int w) {
        if (atomicExch(&a.win_bail[w], 1u) == 0u) a.slow_list[a.w_begin + atomicAdd(&a.hdr->n_slow[a.slot], 1)] = w;
    };
    // windows with genotypes outside {0,1,2} that the list path is not taking in general mode
    auto route_exotic = [&](int i) -> bool {
        if (kTwoQueues || !gex[i]) return false;
        if (chunk == 0 && tid == 0) a.slow_list[a.w_begin + atomicAdd(&a.hdr->n_slow[a.slot], 1)] = w0 + i;
        return true;
    };

    if (nwin > 1 && c.mono_hi) {
        // ---- balanced path.  Pass 1, one thread per column: both list cursors only move forward,
        // so locating every window's run [a, a + n) costs O(list length); units with n <= 2 are
        // NaN and finished here (most units of sparse haplotype data), the others are queued.
        // Pass 2: the CTA's threads share the queued units evenly, so a warp's DP loop no longer
        // runs for the longest of 32 unrelated columns in every window.
        int cur_a = 0, cur_b = 0, odd_a = 0, odd_b = 0;  // (general body: entries passed whose genotype is not 1 / 2)
        for (int i = 0; i < nwin; ++i) {
            const int w = w0 + i;
            const int32_t lo = glo[i], hi = ghi[i];
            const bool ex = route_exotic(i);
            int n = 0;
            bool queued = false;
            int64_t score = kNegInf64;
            if (t < a.T && !ex && hi > lo) {
                if (!overflow) {
                    if (kGen) list_window_range_general(s_list + tid, ls, cur_a, cur_b, odd_a, odd_b, gplo[i], gphi[i]);
                    else if (kSigned && col_neg) list_window_range_signed(s_list + tid, ls, cur_a, cur_b, odd_a, odd_b, gplo[i], gphi[i]);
                    else list_window_range_shifted<kShift>(s_list + tid, ls, cur_a, cur_b, gplo[i], gphi[i], true);
                    n = cur_b - cur_a;
                    queued = n > 2;
                } else if (kGen || (kSigned && col_neg)) {
                    bail_window(w);  // (a signed column without a negative site walks its clean planes below)
                } else {
                    walk_dp_unit<int32_t>(LdgU32(), LdgPos{a.pos}, a.carried, a.two, a.Tp, t, lo, hi, a.bonus, a.penalty,
                                          a.max_mm, n, score);
                }
            }
            // warp-aggregated append of (column, window, first entry, length).  In the general body the
            // units whose genotypes all are 1 / 2 (most of them) queue from the front and take the
            // four-maxima DP, the others queue from the far end and take the per-class DP: the
            // warps of pass 2 then run one body at a time.
            const bool other = kTwoQueues && queued && odd_b != odd_a;  // some genotype of the unit is not 1 / 2
            const uint32_t qm = __ballot_sync(0xffffffffu, queued && !other);
            const uint32_t om = kTwoQueues ? __ballot_sync(0xffffffffu, other) : 0u;
            const uint32_t item = (uint32_t)tid | ((uint32_t)i << 7) | ((uint32_t)cur_a << 11) | ((uint32_t)n << 21);
            if (qm) {
                const int lane = tid & 31;
                int basei = 0;
                if (lane == __ffs(qm) - 1) basei = atomicAdd(&g.n_work, __popc(qm));
                basei = __shfl_sync(0xffffffffu, basei, __ffs(qm) - 1);
                if (queued && !other) s_work[basei + __popc(qm & ((1u << lane) - 1u))] = item;
            }
            if (om) {
                const int lane = tid & 31;
                int basei = 0;
                if (lane == __ffs(om) - 1) basei = atomicAdd(&g.n_work_back, __popc(om));
                basei = __shfl_sync(0xffffffffu, basei, __ffs(om) - 1);
                if (other) s_work[nwin * bd - 1 - (basei + __popc(om & ((1u << lane) - 1u)))] = item;
            }
            if (t < a.T && !ex) {
                const size_t o = (size_t)w * a.T + t;
                if (!queued) a.score[o] = score;  // NaN (n <= 2, empty window) or the overflow column's score
                a.nsnp[o] = (hi - lo) - gnabs[i] + n;
            }
        }
        __syncthreads();
        const int n_work = g.n_work;
        for (int u = tid; u < n_work; u += bd) {
            const uint32_t e = s_work[u];
            const int col = e & 127, i = (e >> 7) & 15, ua = (e >> 11) & 1023, n = e >> 21;
            const int64_t score = list_dp_run<kShift>(s_list + col + ua * ls, ls, n, a.bonus, c.pen_eff);
            a.score[(size_t)(w0 + i) * a.T + chunk * bd + col] = score;
        }
        if (kTwoQueues) {
            const int n_back = g.n_work_back;
            for (int u = tid; u < n_back; u += bd) {
                const uint32_t e = s_work[nwin * bd - 1 - u];
                const int col = e & 127, i = (e >> 7) & 15, ua = (e >> 11) & 1023, n = e >> 21;
                int64_t score;
                if (kSigned) {
                    score = list_dp_run_signed(s_list + col + ua * ls, ls, n, a.bonus, a.penalty, compat);
                } else if (!list_dp_run_general(s_list + col + ua * ls, ls, n, a.bonus, a.penalty, a.max_mm, score)) {
                    bail_window(w0 + i);
                    continue;
                }
                a.score[(size_t)(w0 + i) * a.T + chunk * bd + col] = score;
            }
        }
        return;
    }

    // ---- direct path (a single window, or window ends that do not ascend)
    int cur = 0;
    for (int i = 0; i < nwin; ++i) {
        const int w = w0 + i;
        if (route_exotic(i) || t >= a.T) continue;
        const int32_t lo = glo[i], hi = ghi[i];
        int n = 0;
        int64_t score = kNegInf64;
        if (hi > lo) {
            if (overflow) {
                if (kGen || (kSigned && col_neg)) { bail_window(w); continue; }
                walk_dp_unit<int32_t>(LdgU32(), LdgPos{a.pos}, a.carried, a.two, a.Tp, t, lo, hi, a.bonus, a.penalty,
                                      a.max_mm, n, score);
            } else {
                int b = cur;
                list_window_range_shifted<kShift>(s_list + tid, ls, cur, b, gplo[i], gphi[i], false);
                n = b - cur;
                if (n > 2) {
                    if (kSigned && col_neg) score = list_dp_run_signed(s_list + tid + cur * ls, ls, n, a.bonus, a.penalty, compat);
                    else if (!kGen) score = list_dp_run<kShift>(s_list + tid + cur * ls, ls, n, a.bonus, c.pen_eff);
                    else if (!list_dp_run_general(s_list + tid + cur * ls, ls, n, a.bonus, a.penalty, a.max_mm, score)) {
                        bail_window(w);
                        continue;
                    }
                }
            }
        }
        const size_t o = (size_t)w * a.T + t;
        a.score[o] = score;
        a.nsnp[o] = (hi - lo) - gnabs[i] + n;
    }
}

// the general instantiation lives out of line: it is rare, and its per-class tables must not
// cost the common kernel body registers or stack
template <bool kNarrow>
__device__ __noinline__ void group_body_general(const ScoreArgs &a, const GroupCtx &c) { group_body<1, kNarrow>(a, c); }
template <bool kNarrow>
__device__ __noinline__ void group_body_signed(const ScoreArgs &a, const GroupCtx &c) { group_body<2, kNarrow>(a, c); }

// grid.x = tchunks * ngroups; dynamic smem: list [L + 1][bd] | max(positions [words_cap*32], worklist [wg*bd])
// kNarrow (panels of at most 16 columns) lives in a kernel of its own: there the lists of a column
// sit 8 or 16 words apart instead of blockDim.x, which the common kernel must not pay registers for.
template <bool kNarrow>
__device__ __forceinline__ void score_group_body(const ScoreArgs &a, int wg, int L, int words_cap) {
    extern __shared__ __align__(16) uint8_t dyn[];
    __shared__ GroupShared g;
    pdl_launch_dependents();
    const int tid = threadIdx.x, bd = blockDim.x;
    const int chunk = blockIdx.x % a.tchunks, grp = blockIdx.x / a.tchunks;
    const int w0 = a.w_begin + grp * wg;
    const int nwin = min(wg, a.w_end - w0);
    const int t = chunk * bd + tid;
    // the position tile is only read while the lists are built, the unit worklist only after:
    // they share one region (a barrier separates the two phases)
    uint32_t *s_list = reinterpret_cast<uint32_t *>(dyn);
    // narrow panels keep their lists 8 or 16 words apart instead of a whole warp's 32 (less shared
    // memory per CTA: replicate batches of a dozen columns then keep three times the CTAs resident)
    const int ls = kNarrow ? score_list_stride(a.T, bd) : bd;
    uint32_t *s_pos = s_list + (size_t)(L + 1) * ls;  // row L of the list area is the spill slot

    pdl_wait();  // bit-planes and window bounds come from the prep launch
    if (tid < nwin) {
        g.lo[tid] = a.win_lo[w0 + tid];
        g.hi[tid] = a.win_hi[w0 + tid];
        g.pfirst[tid] = a.win_pfirst[w0 + tid];  // (undefined for empty windows, never used then)
        g.plast[tid] = a.win_plast[w0 + tid];
        g.nabs[tid] = 0;
        g.ex[tid] = 0;
    }
    __syncthreads();
    const int32_t pen_eff = (a.max_mm >= 1) ? a.penalty : Sentinel<int32_t>::NEG;

    // The group is consumed as one or more SUB-GROUPS [s0, s1): the longest run of windows that
    // ascend (in variant index and in position), whose union variant range fits the staged
    // position tile and whose scores fit int32.  Regular overlapping windows give one sub-group;
    // disjoint or unordered windows simply give shorter ones.  A window that does not fit even
    // alone is walked straight from the planes.
    int s0 = 0;
    while (s0 < nwin) {
        int32_t rlo = INT32_MAX, rhi = -1, prev_lo = -1, prev_hi = -1, prev_pfirst = INT32_MIN, maxnv = 0, p0 = 0,
                plast = INT32_MIN;
        bool mono_hi = true, alone = false;
        int s1 = s0;
        for (; s1 < nwin; ++s1) {
            const int32_t lo = g.lo[s1], hi = g.hi[s1];
            if (hi <= lo) continue;  // empty windows ride along
            const int32_t n_rlo = min(rlo, lo), n_rhi = max(rhi, hi);
            const int32_t n_p0 = lo < rlo ? g.pfirst[s1] : p0, n_plast = max(plast, g.plast[s1]);
            const int32_t n_maxnv = max(maxnv, hi - lo);
            const bool fits = lo >= prev_lo && g.pfirst[s1] >= prev_pfirst && g.plast[s1] >= g.pfirst[s1] &&
                              ((n_rhi - 1) >> 5) - (n_rlo >> 5) + 1 <= words_cap &&
                              score_bound((int64_t)n_plast - (int64_t)n_p0, n_maxnv, a.bonus, a.penalty) < kSmallBound;
            if (!fits) {
                if (rhi < 0) { alone = true; ++s1; }  // not even by itself: walk it
                break;
            }
            mono_hi = mono_hi && hi >= prev_hi;
            prev_lo = lo; prev_hi = hi; prev_pfirst = g.pfirst[s1];
            rlo = n_rlo; rhi = n_rhi; p0 = n_p0; plast = n_plast; maxnv = n_maxnv;
        }
        const int nsub = s1 - s0;

        if (alone) {
            // ---- window by window from global memory
            for (int i = s0; i < s1; ++i) {
                const int w = w0 + i;
                const int32_t lo = g.lo[i], hi = g.hi[i];
                if (tid == 0) { g.red_nabs = 0; g.red_ex = 0; }
                __syncthreads();
                if (hi > lo) {
                    const int wi0 = lo >> 5, wi1 = (hi - 1) >> 5;
                    int na = 0;
                    uint32_t ex = 0;
                    for (int gi = wi0 + tid; gi <= wi1; gi += bd) {
                        uint32_t rm = 0xffffffffu;
                        if (gi == wi0) rm &= 0xffffffffu << (lo & 31);
                        if (gi == wi1) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                        na += __popc(__ldg(a.absent_bits + gi) & rm);
                        ex |= __ldg(a.exotic + gi);
                    }
                    na = __reduce_add_sync(0xffffffffu, na);
                    ex = __reduce_or_sync(0xffffffffu, ex);
                    if ((tid & 31) == 0 && (na | ex)) {
                        atomicAdd(&g.red_nabs, na);
                        if (ex) atomicOr(&g.red_ex, ex);
                    }
                }
                __syncthreads();
                const int32_t nabs = g.red_nabs;
                const uint32_t wex = g.red_ex;
                __syncthreads();  // everyone has read the reduction before the next window resets it
                if (wex) {
                    if (chunk == 0 && tid == 0) a.slow_list[a.w_begin + atomicAdd(&a.hdr->n_slow[a.slot], 1)] = w;
                    continue;
                }
                if (t >= a.T) continue;
                int n = 0;
                int64_t score = kNegInf64;
                if (hi > lo) {
                    const int64_t wspan = (int64_t)g.plast[i] - (int64_t)g.pfirst[i];
                    if (score_bound(wspan, hi - lo, a.bonus, a.penalty) < kSmallBound)
                        walk_dp_unit<int32_t>(LdgU32(), LdgPos{a.pos}, a.carried, a.two, a.Tp, t, lo, hi, a.bonus,
                                              a.penalty, a.max_mm, n, score);
                    else
                        walk_dp_unit<int64_t>(LdgU32(), LdgPos{a.pos}, a.carried, a.two, a.Tp, t, lo, hi, a.bonus,
                                              a.penalty, a.max_mm, n, score);
                }
                const size_t o = (size_t)w * a.T + t;
                a.score[o] = score;
                a.nsnp[o] = (hi - lo) - nabs + n;
            }
        } else if (rhi < 0) {
            // ---- nothing but empty windows: NaN, 0 SNPs (sstar.py:191-193 with n = 0)
            if (t < a.T)
                for (int i = s0; i < s1; ++i) {
                    const size_t o = (size_t)(w0 + i) * a.T + t;
                    a.score[o] = kNegInf64;
                    a.nsnp[o] = 0;
                }
        } else {
            // ---- list path: positions of the union range, relative to its first variant
            const int gi0 = rlo >> 5, nwr = ((rhi - 1) >> 5) - gi0 + 1;
            const int64_t span = (int64_t)plast - (int64_t)p0;
            const int32_t base = gi0 << 5;
            const int32_t npos = min(nwr << 5, a.V - base);
#pragma unroll 4
            for (int i = tid; i < npos; i += bd) s_pos[i] = (uint32_t)(__ldg(a.pos + base + i) - p0);
            if (tid < nsub) {
                const int32_t lo = g.lo[s0 + tid], hi = g.hi[s0 + tid];
                g.plo[s0 + tid] = hi > lo ? (uint32_t)(g.pfirst[s0 + tid] - p0) : 1u;
                g.phi[s0 + tid] = hi > lo ? (uint32_t)(g.plast[s0 + tid] - p0) : 0u;
            }
            // #absent variants and the exotic flag per window: a warp per window, lanes over its words
            for (int i = s0 + (tid >> 5); i < s1; i += bd >> 5) {
                const int32_t lo = g.lo[i], hi = g.hi[i];
                int na = 0;
                uint32_t ex = 0;
                if (hi > lo) {
                    const int wi0 = lo >> 5, wi1 = (hi - 1) >> 5;
                    for (int gi = wi0 + (tid & 31); gi <= wi1; gi += 32) {
                        uint32_t rm = 0xffffffffu;
                        if (gi == wi0) rm &= 0xffffffffu << (lo & 31);
                        if (gi == wi1) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                        na += __popc(__ldg(a.absent_bits + gi) & rm);
                        ex |= __ldg(a.exotic + gi);
                    }
                }
                na = __reduce_add_sync(0xffffffffu, na);
                ex = __reduce_or_sync(0xffffffffu, ex);
                if ((tid & 31) == 0) { g.nabs[i] = na; g.ex[i] = ex; }  // ex: OR of the groups' levels (0 clean, 1 = -1 / -2 only, >= 2 any byte)
            }
            __syncthreads();

            // Windows holding genotypes outside {0,1,2}: the whole sub-group takes the general body.
            // Only when it spans more than 2^23 bp do such windows go to the pairwise kernel.
            uint32_t ex_or = 0;
            for (int i = s0; i < s1; ++i) ex_or |= g.ex[i];
            const bool any_ex = ex_or != 0;
            GroupCtx c;
            c.g = &g; c.s_list = s_list; c.s_pos = s_pos;
            c.tid = tid; c.bd = bd; c.chunk = chunk; c.t = t; c.w0 = w0 + s0; c.nwin = nsub; c.L = L; c.io = s0;
            c.rlo = rlo; c.rhi = rhi; c.gi0 = gi0; c.nwr = nwr; c.pen_eff = pen_eff; c.mono_hi = mono_hi;
            // other genotypes, all -1 / -2: signed entries from the planes; any other byte: general entries
            // (only a sub-group spanning more than 2^23 bp sends such windows to the pairwise kernel)
            if (ex_or == 1u) group_body_signed<kNarrow>(a, c);
            else if (any_ex && span < (int64_t(1) << kGeneralPosBits)) group_body_general<kNarrow>(a, c);
            else group_body<0, kNarrow>(a, c);
        }
        s0 = s1;
        if (s0 < nwin) __syncthreads();  // the lists / worklist are reused by the next sub-group
    }
}

__global__ void __launch_bounds__(kScoreMaxThreads)
score_group_kernel(const __grid_constant__ ScoreArgs a, int wg, int L, int words_cap) {
    score_group_body<false>(a, wg, L, words_cap);
}

__global__ void __launch_bounds__(32)
score_group_narrow_kernel(const __grid_constant__ ScoreArgs a, int wg, int L, int words_cap) {
    score_group_body<true>(a, wg, L, words_cap);
}

// ------------------------------------------------------------------------------------------
// K3: the literal pairwise recurrence (sstar/sstar.py:195-217), one WARP per unit, persistent
// grid, any int8 genotype.  Lanes stride over i < j; the row max is two REDUX steps.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t warp_max_i64(int64_t v) {
    const int32_t hi = (int32_t)(v >> 32);
    const int32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t lo = (hi == mh) ? (uint32_t)v : 0u;
    const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
    return ((int64_t)mh << 32) | (int64_t)ml;
}

__global__ void __launch_bounds__(kPairWarps * 32) score_pairwise_kernel(ScoreArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_maxvw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_maxvw = 0;
    pdl_wait();  // slow_list / n_slow come from the previous launches
    const int n_slow = a.force_slow ? a.w_end - a.w_begin : a.hdr->n_slow[a.slot];
    if (n_slow <= 0) return;
    auto slow_window = [&](int64_t i) -> int { return a.force_slow ? a.w_begin + (int)i : a.slow_list[a.w_begin + i]; };
    __syncthreads();
    {   // largest routed window decides where the lists live
        int mv = 0;
        for (int i = threadIdx.x; i < n_slow; i += blockDim.x) {
            const int w = slow_window(i);
            mv = max(mv, a.win_hi[w] - a.win_lo[w]);
        }
        mv = __reduce_max_sync(0xffffffffu, mv);
        if (lane == 0 && mv > 0) atomicMax(&s_maxvw, mv);
    }
    __syncthreads();
    const int64_t total = (int64_t)n_slow * a.T;
    const uint32_t cap = (uint32_t)((s_maxvw + 7) & ~7);
    if (cap == 0) {  // every routed window is empty
        for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
             it += (int64_t)gridDim.x * blockDim.x) {
            const size_t o = (size_t)slow_window(it / a.T) * a.T + (size_t)(it % a.T);
            a.score[o] = kNegInf64;
            a.nsnp[o] = 0;
        }
        return;
    }
    int64_t gw = (int64_t)blockIdx.x * kPairWarps + warp;
    int64_t nw = (int64_t)gridDim.x * kPairWarps;
    uint8_t *slab;
    uint32_t slab_cap;
    if (cap <= (uint32_t)a.pair_smem_cap) {
        slab_cap = (uint32_t)a.pair_smem_cap;
        slab = smem + (size_t)warp * a.pair_smem_cap * 13;
    } else {
        slab_cap = cap;
        const int64_t nslabs = (int64_t)(a.arena_entries / cap);
        if (nslabs < nw) nw = nslabs;
        if (gw >= nw) return;
        slab = a.arena + (size_t)gw * cap * 13;
    }
    int64_t *M = reinterpret_cast<int64_t *>(slab);
    int32_t *P = reinterpret_cast<int32_t *>(slab + (size_t)slab_cap * 8);
    int8_t *Gt = reinterpret_cast<int8_t *>(slab + (size_t)slab_cap * 12);

    for (int64_t it = gw; it < total; it += nw) {
        const int w = slow_window(it / a.T);
        const int t = (int)(it % a.T);
        const int32_t lo = a.win_lo[w], hi = a.win_hi[w];
        int n = 0, nabs = 0;
        if (hi > lo) {
            // Lanes fetch 32 words of the column's carried plane (and of the absent mask) in one
            // round trip; the warp then walks the words that hold a site.  Genotype bytes are read
            // for carried sites only.
            const int wi0 = lo >> 5, wi1 = (hi - 1) >> 5;
            for (int gb = wi0; gb <= wi1; gb += 32) {
                const int gl = gb + lane;
                uint32_t my_cw = 0, my_ab = 0;
                if (gl <= wi1) {
                    uint32_t rm = 0xffffffffu;
                    if (gl == wi0) rm &= 0xffffffffu << (lo & 31);
                    if (gl == wi1) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                    my_cw = __ldg(a.carried + (size_t)gl * a.Tp + t) & rm;
                    my_ab = __ldg(a.absent_bits + gl) & rm;
                }
                nabs += __reduce_add_sync(0xffffffffu, __popc(my_ab));
                uint32_t live = __ballot_sync(0xffffffffu, my_cw != 0);
                while (live) {
                    const int src = __ffs(live) - 1;
                    live &= live - 1;
                    const uint32_t cw = __shfl_sync(0xffffffffu, my_cw, src);
                    if ((cw >> lane) & 1u) {
                        const int k = ((gb + src) << 5) + lane;
                        const int idx = n + __popc(cw & ((1u << lane) - 1u));
                        P[idx] = __ldg(a.pos + k);
                        Gt[idx] = a.tgt[(size_t)k * a.T + t];
                    }
                    n += __popc(cw);
                }
            }
        }
        __syncwarp();
        int64_t top = kNegInf64;
        if (n > 2) {
            for (int j = 0; j < n; ++j) {
                const int32_t pj = P[j];
                const int gj = Gt[j];
                int64_t best = kNegInf64;
                for (int i = lane; i < j; i += 32) {
                    const int64_t d = (int64_t)pj - (int64_t)P[i];
                    int qd = gj - (int)Gt[i];
                    qd = qd < 0 ? -qd : qd;
                    if (d >= kMinPairDistance && qd <= a.max_mm) {
                        const int64_t s = (qd == 0) ? d + a.bonus : (int64_t)a.penalty;
                        best = max(best, s + max(M[i], (int64_t)0));  // == max(M[i] + s, s)
                    }
                }
                best = warp_max_i64(best);
                if (lane == 0) M[j] = best;
                top = max(top, best);
                __syncwarp();
            }
        }
        if (lane == 0) {
            const size_t o = (size_t)w * a.T + t;
            a.score[o] = top;  // INT64_MIN == SSTAR_B200_NAN (n <= 2 or no valid pair)
            a.nsnp[o] = (hi - lo) - nabs + n;
        }
        __syncwarp();
    }
}

}  // namespace sstar

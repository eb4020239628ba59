// sstar_b200 -- chain-DP kernels: K2 (prefix-max, thread per unit) and K3 (pairwise, warp per unit).
#pragma once
#include "common.cuh"

namespace sstar {

// ------------------------------------------------------------------------------------------
// K2: prefix-max chain DP, one THREAD per (window, column) unit; CTA = (window, <=128 columns).
//
// With v_i = max(M[i], 0) the recurrence of sstar/sstar.py:208-212 reads
//     M[j] = max_{i<j, p_j-p_i>=10} s(j,i) + v_i .
// For carried genotypes in {1,2}:  g_i == g_j -> s = (p_j - p_i) + bonus
//                                  g_i != g_j -> s = penalty          (only if max_mismatch >= 1)
// hence  M[j] = max( p_j + bonus + max_{g_i=g_j}(v_i - p_i) ,  penalty + max_{g_i!=g_j} v_i ):
// four running maxima (A0,A1,B0,B1).  Positions ascend, so eligible predecessors form a prefix:
// the previous site waits in registers and is committed to the maxima as soon as the current site
// is >= 10 bp ahead; only when carried sites crowd within 9 bp do entries queue in a per-thread
// shared-memory ring (<= 10 distinct integer positions fit in 9 bp, ring has 16 slots).
//
// "-inf" is a sentinel NEG with |real values| < |NEG|/4, so invalid candidates stay below THR
// without any select: m = max(same + p + bonus, other + penalty) is valid iff m > THR, and
// v = max(m, 0) is 0 for an invalid m -- exactly max(M[i] + s, s) of the reference.
// Arithmetic is int32 when the window's score bound allows it (decided per window), else int64.
// ------------------------------------------------------------------------------------------
template <typename acc_t> struct Sentinel;
template <> struct Sentinel<int32_t> {
    static constexpr int32_t NEG = -(1 << 30);
    static constexpr int32_t THR = -(1 << 29);
};
template <> struct Sentinel<int64_t> {
    static constexpr int64_t NEG = -(int64_t(1) << 62);
    static constexpr int64_t THR = -(int64_t(1) << 61);
};

// Per-unit DP.  kStaged: the unit's plane words and the window's positions were staged in shared
// memory by the CTA (thread-private columns car[j*bd+tid] / two[j*bd+tid], shared rel. positions);
// otherwise everything is read from global memory (very long windows, unknown size hint).
// Each lane walks ITS OWN set bits across word boundaries, so a warp iterates max_lane(n) times,
// not sum_words max_lane(popc).
template <typename acc_t, bool kStaged>
__device__ __forceinline__ void prefix_dp_unit(const ScoreArgs &a, int t, int32_t lo, int32_t hi, int32_t p0,
                                               uint8_t *ring_raw, const uint32_t *s_car, const uint32_t *s_two,
                                               const uint32_t *s_pos, int tid, int bd, int &n_out,
                                               int64_t &score_out) {
    constexpr acc_t NEG = Sentinel<acc_t>::NEG, THR = Sentinel<acc_t>::THR;
    uint32_t *ring_p = reinterpret_cast<uint32_t *>(ring_raw);                          // [slots][bd]
    acc_t *ring_v = reinterpret_cast<acc_t *>(ring_raw + (size_t)kRingSlots * bd * 4);  // [slots][bd]
    acc_t A0 = NEG, A1 = NEG, B0 = NEG, B1 = NEG, top = NEG;
    const acc_t bonus = (acc_t)a.bonus;
    const acc_t pen_eff = (a.max_mm >= 1) ? (acc_t)a.penalty : NEG;
    bool pend = false;  // the previous carried site, not yet an eligible predecessor
    acc_t pend_p = 0, pend_v = 0;
    int pend_c = 0;
    int q = 0, head = 0;  // older pending sites (crowded positions only)
    int n = 0;

    const int gi0 = lo >> 5, nw = ((hi - 1) >> 5) - gi0 + 1;
    const int lo_bit = lo & 31;
    const uint32_t first_mask = 0xffffffffu << lo_bit;
    const uint32_t last_mask = 0xffffffffu >> (31 - ((hi - 1) & 31));
    auto load_car = [&](int j) -> uint32_t {
        if (kStaged) return s_car[j * bd + tid];
        uint32_t c = __ldg(a.carried + (size_t)(gi0 + j) * a.Tp + t);
        if (j == 0) c &= first_mask;
        if (j == nw - 1) c &= last_mask;
        return c;
    };
    int j = 0;
    uint32_t c = load_car(0), tw = 0;
    bool need_tw = true;
    if (!kStaged) n = 0;
    for (;;) {
        while (c == 0) {
            if (++j >= nw) goto done;
            c = load_car(j);
            need_tw = true;
        }
        if (need_tw) {
            tw = kStaged ? s_two[j * bd + tid] : __ldg(a.two + (size_t)(gi0 + j) * a.Tp + t);
            need_tw = false;
            n += __popc(c);
        }
        {
            const int b = __ffs(c) - 1;
            c &= c - 1;
            const int idx = (j << 5) + b - lo_bit;  // variant index relative to lo
            const acc_t p = kStaged ? (acc_t)s_pos[idx]
                                    : (acc_t)((int64_t)__ldg(a.pos + lo + idx) - (int64_t)p0);
            const int cls = (tw >> b) & 1;
            while (q > 0) {  // rare: crowded sites that have fallen >= 10 bp behind
                const acc_t pi = (acc_t)ring_p[head * bd + tid];
                if (p - pi < kMinPairDistance) break;
                const acc_t vv = ring_v[head * bd + tid];
                const acc_t v = vv >> 1;
                if (vv & 1) { A1 = max(A1, v - pi); B1 = max(B1, v); }
                else        { A0 = max(A0, v - pi); B0 = max(B0, v); }
                head = (head + 1) & (kRingSlots - 1);
                --q;
            }
            if (pend) {
                if (q == 0 && p - pend_p >= kMinPairDistance) {
                    if (pend_c) { A1 = max(A1, pend_v - pend_p); B1 = max(B1, pend_v); }
                    else        { A0 = max(A0, pend_v - pend_p); B0 = max(B0, pend_v); }
                } else {
                    const int tail = (head + q) & (kRingSlots - 1);
                    ring_p[tail * bd + tid] = (uint32_t)pend_p;  // 0 <= p < 2^32
                    ring_v[tail * bd + tid] = (pend_v << 1) | (acc_t)pend_c;
                    ++q;
                }
            }
            const acc_t same = cls ? A1 : A0;
            const acc_t other = cls ? B0 : B1;
            const acc_t m = max(same + (p + bonus), other + pen_eff);
            top = max(top, m);
            pend = true;
            pend_p = p;
            pend_v = max(m, (acc_t)0);
            pend_c = cls;
        }
    }
done:
    n_out = n;
    score_out = (n <= 2 || top <= THR) ? kNegInf64 : (int64_t)top;
}

// grid: x = column chunk, (y, z) = window; dynamic smem: ring | car tile | two tile | positions
__global__ void __launch_bounds__(kScoreMaxThreads)
score_prefix_kernel(const __grid_constant__ ScoreArgs a, int words_cap, int pos_cap) {
    extern __shared__ __align__(16) uint8_t dyn[];
    __shared__ int s_nabs;
    __shared__ uint32_t s_ex;
    pdl_launch_dependents();
    const int tid = threadIdx.x, bd = blockDim.x;
    const int w = blockIdx.z * gridDim.y + blockIdx.y;
    if (w >= a.W) return;
    const int t = blockIdx.x * bd + tid;
    uint8_t *ring_raw = dyn;
    uint32_t *s_car = reinterpret_cast<uint32_t *>(dyn + (size_t)kRingSlots * bd * 12);
    uint32_t *s_two = s_car + (size_t)words_cap * bd;
    uint32_t *s_pos = s_two + (size_t)words_cap * bd;
    if (tid == 0) { s_nabs = 0; s_ex = 0; }
    pdl_wait();  // bit-planes and window bounds come from the prep launch
    const int32_t lo = a.win_lo[w], hi = a.win_hi[w];
    __syncthreads();

    int n = 0;
    int32_t p0 = 0;
    bool staged = false;
    if (hi > lo) {
        const int gi0 = lo >> 5, gi1 = (hi - 1) >> 5, nw = gi1 - gi0 + 1;
        // window-level quantities, once per CTA: #absent variants and the exotic flag
        int na = 0;
        uint32_t ex = 0;
        for (int gi = gi0 + tid; gi <= gi1; gi += bd) {
            uint32_t rm = 0xffffffffu;
            if (gi == gi0) rm &= 0xffffffffu << (lo & 31);
            if (gi == gi1) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
            na += __popc(__ldg(a.absent_bits + gi) & rm);
            ex |= __ldg(a.exotic + gi);
        }
        na = __reduce_add_sync(0xffffffffu, na);
        ex = __reduce_or_sync(0xffffffffu, ex);
        if ((tid & 31) == 0 && (na | ex)) {
            atomicAdd(&s_nabs, na);
            if (ex) atomicOr(&s_ex, ex);
        }
        p0 = __ldg(a.pos + lo);
        staged = nw <= words_cap && (hi - lo) <= pos_cap;
        if (staged) {
            // every thread stages ITS column (coalesced across the warp, all loads in flight at once)
            if (t < a.T) {
                const uint32_t *cp = a.carried + (size_t)gi0 * a.Tp + t;
                const uint32_t *tp = a.two + (size_t)gi0 * a.Tp + t;
                for (int jb = 0; jb < nw; jb += 4) {
                    uint32_t cw[4], tw[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) cw[u] = (jb + u < nw) ? __ldg(cp + (size_t)(jb + u) * a.Tp) : 0u;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (jb + u == 0) cw[u] &= 0xffffffffu << (lo & 31);
                        if (jb + u == nw - 1) cw[u] &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                        tw[u] = cw[u] ? __ldg(tp + (size_t)(jb + u) * a.Tp) : 0u;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (jb + u < nw) {
                            s_car[(jb + u) * bd + tid] = cw[u];
                            s_two[(jb + u) * bd + tid] = tw[u];
                            n += __popc(cw[u]);
                        }
                    }
                }
            }
            for (int i = tid; i < hi - lo; i += bd) s_pos[i] = (uint32_t)(__ldg(a.pos + lo + i) - p0);
        }
    }
    __syncthreads();
    if (s_ex) {  // genotypes outside {0,1,2}: the whole window goes to the pairwise kernel
        if (tid == 0 && blockIdx.x == 0) a.slow_list[atomicAdd(&a.hdr->n_slow, 1)] = w;
        return;
    }
    if (t >= a.T) return;

    int64_t score = kNegInf64;
    if (hi > lo) {
        const int64_t span = (int64_t)__ldg(a.pos + hi - 1) - (int64_t)p0;
        const int64_t mag = (int64_t)(a.bonus < 0 ? -(int64_t)a.bonus : a.bonus) +
                            (int64_t)(a.penalty < 0 ? -(int64_t)a.penalty : a.penalty);
        const int64_t bound = span + (int64_t)(hi - lo + 1) * mag;  // |any chain score| <= bound
        const bool small = bound < (int64_t(1) << 28);
        if (staged) {
            int n2 = 0;  // n was counted while staging
            if (small) prefix_dp_unit<int32_t, true>(a, t, lo, hi, p0, ring_raw, s_car, s_two, s_pos, tid, bd, n2, score);
            else       prefix_dp_unit<int64_t, true>(a, t, lo, hi, p0, ring_raw, s_car, s_two, s_pos, tid, bd, n2, score);
        } else {
            if (small) prefix_dp_unit<int32_t, false>(a, t, lo, hi, p0, ring_raw, s_car, s_two, s_pos, tid, bd, n, score);
            else       prefix_dp_unit<int64_t, false>(a, t, lo, hi, p0, ring_raw, s_car, s_two, s_pos, tid, bd, n, score);
        }
    }
    const size_t o = (size_t)w * a.T + t;
    a.score[o] = score;
    a.nsnp[o] = (hi - lo) - s_nabs + n;
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
    const int n_slow = a.hdr->n_slow;
    if (n_slow <= 0) return;
    __syncthreads();
    {   // largest routed window decides where the lists live
        int mv = 0;
        for (int i = threadIdx.x; i < n_slow; i += blockDim.x) {
            const int w = a.slow_list[i];
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
            const size_t o = (size_t)a.slow_list[it / a.T] * a.T + (size_t)(it % a.T);
            a.score[o] = kNegInf64;
            a.nsnp[o] = 0;
        }
        return;
    }
    int64_t gw = (int64_t)blockIdx.x * kPairWarps + warp;
    int64_t nw = (int64_t)gridDim.x * kPairWarps;
    uint8_t *slab;
    uint32_t slab_cap;
    if (cap <= (uint32_t)kPairSmemCap) {
        slab_cap = kPairSmemCap;
        slab = smem + (size_t)warp * kPairSmemCap * 13;
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
        const int w = a.slow_list[it / a.T];
        const int t = (int)(it % a.T);
        const int32_t lo = a.win_lo[w], hi = a.win_hi[w];
        int n = 0, nabs = 0;
        if (hi > lo) {
            for (int gi = lo >> 5; gi <= ((hi - 1) >> 5); ++gi) {
                const int k = (gi << 5) + lane;
                const uint32_t ab = __ldg(a.absent_bits + gi);
                const bool isab = k >= lo && k < hi && ((ab >> lane) & 1);
                int g = 0;
                if (isab) g = a.tgt[(size_t)k * a.T + t];
                const uint32_t ab_m = __ballot_sync(0xffffffffu, isab);
                const uint32_t car_m = __ballot_sync(0xffffffffu, g != 0);
                nabs += __popc(ab_m);
                if (g != 0) {
                    const int idx = n + __popc(car_m & ((1u << lane) - 1u));
                    P[idx] = __ldg(a.pos + k);
                    Gt[idx] = (int8_t)g;
                }
                n += __popc(car_m);
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

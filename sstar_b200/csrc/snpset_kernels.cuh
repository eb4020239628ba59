// sstar_b200 -- S* SNP sets: which sites the best chain of a (window, column) unit runs through.
//
// An extension of the scoring path: the current reference never backtracks (its table ends at
// Region_ind_SNP_number, sstar/feature_vector_preprocessor.py:167-181); the only pins are its
// legacy fixtures tests/exp_results/test.score.{unphased,phased}.exp.results.tsv with the columns
// S*_SNP_number and S*_SNPs.  The rule that reproduces them on the
// literal recurrence of sstar/sstar.py:195-217:
//     M[j] = max_{i<j} ( s(j,i) + max(M[i], 0) )        backpointer = the FIRST i attaining the max;
//     at that i the chain is extended when M[i] >= 0, else it starts afresh with the pair (i, j);
//     the chain ends at the FIRST j attaining max_j M[j].
//
// snp_count_kernel, one thread per unit (a CTA per window x run of columns): #absent rows and the exotic
// flag of the window, then -- in clean windows (genotypes in {0,1,2}, int32-safe scores, at most 32,767
// rows) -- the O(n) prefix-max DP with backpointers right there (dp_core.cuh: walk_dp_unit_chain: each
// running maximum keeps its first maximiser; thread-local backpointers): score, SNP count and chain
// length of the unit in ONE walk of its plane words.  Units it cannot take (a window with other
// genotypes, more than kChainCap sites) get their n from popcounts; those with n > 2 are queued.
// snp_set_kernel, one warp per QUEUED unit (the literal O(n^2) recurrence):
// the site list comes from the planes (genotype bytes are
// fetched only in windows holding values outside {0,1,2}), lanes stride over i < j, the row max is
// a REDUX pair, the first i attaining it a REDUX min.  The result is ragged, so the work is two
// passes around a scan (lengths -> offsets -> fill): the fill pass re-walks the units that have a chain
// (snp_chain_fill_kernel, a thread per unit, for the units the count kernel scored; snp_set_kernel<true>
// over the queue for the others).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sstar {

constexpr int kSetEntryBytes = 17;   // per list entry: M int64, P int32, B int32, G int8
constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 16;
constexpr int kScanPerBlock = kScanThreads * kScanPerThread;

struct SnpSetArgs {
    const int32_t *pos;
    const int8_t *tgt;
    const uint32_t *absent_bits, *exotic, *carried, *two;
    const int32_t *win_lo, *win_hi;
    const int32_t *win_pfirst, *win_plast;  // position of each window's first / last variant
    uint32_t *win_ex;        // [W] count pass: the window holds genotypes outside {0,1,2}
    int32_t *queue;          // [W * T] from the front: units with n > 2 that the count kernel could not score itself;
                             //          from the far end: units it scored that have a chain (the fill pass's work list)
    int32_t *queue_n;        // [2] their numbers (zeroed before the count pass)
    int32_t T, Tp, W;
    int32_t bonus, max_mm, penalty;
    int64_t *score;          // lengths pass: [W, T] (INT64_MIN = NaN)
    int32_t *nsnp;           // lengths pass: [W, T]
    int32_t *set_len;        // lengths pass: [W * T] sites on the best chain (0 when the score is NaN)
    const int64_t *set_off;  // fill pass: [W * T + 1]
    int32_t *set_pos;        // fill pass: positions of the chain sites, ascending, unit after unit
    int64_t capacity;        // fill pass: entries set_pos holds
    uint8_t *arena;
    uint64_t arena_bytes;
    int32_t smem_cap;        // list entries per warp that fit the launch's shared memory (larger windows: arena)
};

__device__ __forceinline__ int64_t set_warp_max_i64(int64_t v) {
    const int32_t hi = (int32_t)(v >> 32);
    const int32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t lo = (hi == mh) ? (uint32_t)v : 0u;
    const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
    return ((int64_t)mh << 32) | (int64_t)ml;
}

constexpr int kCountThreads = 128;

struct ChainLdg { __device__ __forceinline__ uint32_t operator()(const uint32_t *p) const { return __ldg(p); } };
struct ChainPos {
    const int32_t *pos;
    __device__ __forceinline__ int64_t operator()(int32_t k) const { return (int64_t)__ldg(pos + k); }
};

// a window whose units the O(n) chain DP takes (the same test in the count and the fill kernel)
__device__ __forceinline__ bool chain_window(const SnpSetArgs &a, int w, int32_t lo, int32_t hi, uint32_t ex) {
    return hi > lo && ex == 0 && hi - lo <= kChainMaxRows &&
           score_bound((int64_t)a.win_plast[w] - (int64_t)a.win_pfirst[w], hi - lo, a.bonus, a.penalty) < kSmallBound;
}

// grid.x = W * ceil(T / blockDim.x): one window x one run of columns per CTA
__global__ void __launch_bounds__(kCountThreads) snp_count_kernel(const __grid_constant__ SnpSetArgs a) {
    __shared__ int s_nabs, s_nwork;
    __shared__ uint32_t s_ex;
    __shared__ uint8_t s_work[kCountThreads];  // threads (columns of this CTA) whose unit takes the chain DP
    const int tid = threadIdx.x, bd = blockDim.x;
    const int tchunks = (a.T + bd - 1) / bd;
    const int w = blockIdx.x / tchunks, t = (blockIdx.x % tchunks) * bd + tid;
    if (tid == 0) { s_nabs = 0; s_ex = 0; s_nwork = 0; }
    __syncthreads();
    const int32_t lo = a.win_lo[w], hi = a.win_hi[w];
    const int wi0 = lo >> 5, wi1 = (hi - 1) >> 5;
    const uint32_t m0 = 0xffffffffu << (lo & 31), m1 = 0xffffffffu >> (31 - ((hi - 1) & 31));
    if (hi > lo) {
        int na = 0;
        uint32_t ex = 0;
        for (int gi = wi0 + tid; gi <= wi1; gi += bd) {
            const uint32_t rm = (gi == wi0 ? m0 : 0xffffffffu) & (gi == wi1 ? m1 : 0xffffffffu);
            na += __popc(__ldg(a.absent_bits + gi) & rm);
            ex |= __ldg(a.exotic + gi);
        }
        na = __reduce_add_sync(0xffffffffu, na);
        ex = __reduce_or_sync(0xffffffffu, ex);
        if ((tid & 31) == 0 && (na | ex)) {
            atomicAdd(&s_nabs, na);
            if (ex) atomicOr(&s_ex, 1u);
        }
    }
    __syncthreads();
    if (blockIdx.x % tchunks == 0 && tid == 0) a.win_ex[w] = s_ex;
    int n = 0;
    if (t < a.T && hi > lo) {  // n from popcounts: two thirds of the units of sparse data end here (n <= 2: NaN)
        const uint32_t *cp = a.carried + (size_t)wi0 * a.Tp + t;
        for (int gi = wi0; gi <= wi1; ++gi, cp += a.Tp) {
            const uint32_t rm = (gi == wi0 ? m0 : 0xffffffffu) & (gi == wi1 ? m1 : 0xffffffffu);
            n += __popc(__ldg(cp) & rm);
        }
    }
    // Units with n > 2 in a window the O(n) chain DP takes go on the CTA's work list and are then
    // shared out evenly: a third of the units of sparse data have work, and a warp whose lanes each ran
    // their own column would idle two lanes in three.
    const bool dp_unit = t < a.T && n > 2 && n <= kChainCap && chain_window(a, w, lo, hi, s_ex);
    const bool queued = t < a.T && n > 2 && !dp_unit;  // other genotypes in the window, too many sites, int64 scores
    const int lane = tid & 31;
    {
        const uint32_t dm = __ballot_sync(0xffffffffu, dp_unit);
        if (dm) {
            const int leader = __ffs(dm) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&s_nwork, __popc(dm));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (dp_unit) s_work[base + __popc(dm & ((1u << lane) - 1u))] = (uint8_t)tid;
        }
    }
    if (t < a.T) {
        const size_t o = (size_t)w * a.T + t;
        a.nsnp[o] = (hi - lo) - s_nabs + n;
        if (!queued && !dp_unit) { a.score[o] = INT64_MIN; a.set_len[o] = 0; }  // n <= 2 (or an empty window): NaN
    }
    const uint32_t qm = __ballot_sync(0xffffffffu, queued);
    if (qm) {
        const int leader = __ffs(qm) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(a.queue_n, __popc(qm));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (queued) a.queue[base + __popc(qm & ((1u << lane) - 1u))] = w * a.T + t;
    }
    __syncthreads();
    // ---- the work list: score and chain length by the O(n) DP with backpointers
    const int n_work = s_nwork;
    for (int u0 = 0; u0 < n_work; u0 += bd) {
        const int u = u0 + tid;
        bool chained = false;
        int tu = 0;
        if (u < n_work) {
            tu = (blockIdx.x % tchunks) * bd + s_work[u];
            uint16_t row[kChainCap];
            int16_t back[kChainCap];
            int top_j = -1, n2;
            int64_t score;
            walk_dp_unit_chain<int32_t>(ChainLdg(), ChainPos{a.pos}, a.carried, a.two, a.Tp, tu, lo, hi, a.bonus, a.penalty, a.max_mm,
                                        row, back, n2, top_j, score);
            const int len = chain_length(back, top_j);
            const size_t o = (size_t)w * a.T + tu;
            a.score[o] = score;
            a.set_len[o] = len;
            chained = len > 0;
        }
        const uint32_t cm = __ballot_sync(0xffffffffu, chained);  // the fill pass's work list, from the far end
        if (cm) {
            const int leader = __ffs(cm) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(a.queue_n + 1, __popc(cm));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (chained) a.queue[(int64_t)a.W * a.T - 1 - (base + __popc(cm & ((1u << lane) - 1u)))] = w * a.T + tu;
        }
    }
}

// ---- fill pass of the units the count kernel scored itself: a thread per unit of its work list
constexpr int kChainThreads = 128;
__global__ void __launch_bounds__(kChainThreads) snp_chain_fill_kernel(const __grid_constant__ SnpSetArgs a) {
    const int64_t total = a.queue_n[1];
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t it = a.queue[(int64_t)a.W * a.T - 1 - q];
        const int64_t off = a.set_off[it];
        const int len = (int)(a.set_off[it + 1] - off);
        if (len == 0 || off + len > a.capacity) continue;
        const int w = (int)(it / a.T), t = (int)(it % a.T);
        const int32_t lo = a.win_lo[w], hi = a.win_hi[w];
        uint16_t row[kChainCap];
        int16_t back[kChainCap];
        int n = 0, top_j = -1;
        int64_t score;
        walk_dp_unit_chain<int32_t>(ChainLdg(), ChainPos{a.pos}, a.carried, a.two, a.Tp, t, lo, hi, a.bonus, a.penalty, a.max_mm,
                                    row, back, n, top_j, score);
        int32_t *dst = a.set_pos + off;
        const int32_t *psrc = a.pos + lo;
        chain_rows(row, back, top_j, len, [&](int k, uint16_t r) { dst[k] = __ldg(psrc + r); });
    }
}

template <bool kFill>
__global__ void __launch_bounds__(kPairWarps * 32) snp_set_kernel(const __grid_constant__ SnpSetArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_maxvw;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (*a.queue_n == 0) return;  // the count pass scored every unit itself (clean data)
    if (threadIdx.x == 0) s_maxvw = 0;
    __syncthreads();
    {   // the largest window decides where the lists live
        int mv = 0;
        for (int w = threadIdx.x; w < a.W; w += blockDim.x) mv = max(mv, a.win_hi[w] - a.win_lo[w]);
        mv = __reduce_max_sync(0xffffffffu, mv);
        if (lane == 0 && mv > 0) atomicMax(&s_maxvw, mv);
    }
    __syncthreads();
    const int64_t total = *a.queue_n;
    const uint32_t cap = (uint32_t)((s_maxvw + 7) & ~7);
    if (cap == 0 || total == 0) return;  // the count pass has written every other unit
    int64_t gw = (int64_t)blockIdx.x * kPairWarps + warp;
    int64_t nw = (int64_t)gridDim.x * kPairWarps;
    uint8_t *slab;
    uint32_t slab_cap;
    if (cap <= (uint32_t)a.smem_cap) {
        slab_cap = (uint32_t)a.smem_cap;
        slab = smem + (size_t)warp * a.smem_cap * kSetEntryBytes;
    } else {
        slab_cap = cap;
        const int64_t nslabs = (int64_t)(a.arena_bytes / ((uint64_t)cap * kSetEntryBytes));
        if (nslabs < nw) nw = nslabs;
        if (gw >= nw) return;
        slab = a.arena + (size_t)gw * cap * kSetEntryBytes;
    }
    int64_t *M = reinterpret_cast<int64_t *>(slab);
    int32_t *P = reinterpret_cast<int32_t *>(slab + (size_t)slab_cap * 8);
    int32_t *B = reinterpret_cast<int32_t *>(slab + (size_t)slab_cap * 12);
    int8_t *Gt = reinterpret_cast<int8_t *>(slab + (size_t)slab_cap * 16);

    for (int64_t q = gw; q < total; q += nw) {
        const int64_t it = a.queue[q];
        int64_t off = 0;
        int len = 0;
        if (kFill) {
            off = a.set_off[it];
            len = (int)(a.set_off[it + 1] - off);
            if (len == 0 || off + len > a.capacity) continue;  // warp-uniform
        }
        const int w = (int)(it / a.T);
        const int t = (int)(it % a.T);
        const int32_t lo = a.win_lo[w], hi = a.win_hi[w];
        const bool wex = a.win_ex[w] != 0;
        int n = 0;
        {   // the unit's carried, reference-absent sites: one plane word per 32 variants.  Lanes
            // fetch 32 words of the column at once (one round trip), then the warp walks them.
            const int wi0 = lo >> 5, wi1 = (hi - 1) >> 5;
            for (int gb = wi0; gb <= wi1; gb += 32) {
                const int gl = gb + lane;
                uint32_t my_cw = 0, my_tw = 0;
                if (gl <= wi1) {
                    my_cw = __ldg(a.carried + (size_t)gl * a.Tp + t);
                    my_tw = __ldg(a.two + (size_t)gl * a.Tp + t);
                    if (gl == wi0) my_cw &= 0xffffffffu << (lo & 31);
                    if (gl == wi1) my_cw &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                }
                uint32_t live = __ballot_sync(0xffffffffu, my_cw != 0);  // words holding a site
                while (live) {
                    const int src = __ffs(live) - 1;
                    live &= live - 1;
                    const uint32_t cw = __shfl_sync(0xffffffffu, my_cw, src);
                    const uint32_t tw = __shfl_sync(0xffffffffu, my_tw, src);
                    if ((cw >> lane) & 1u) {
                        const int k = ((gb + src) << 5) + lane;
                        const int idx = n + __popc(cw & ((1u << lane) - 1u));
                        P[idx] = __ldg(a.pos + k);
                        Gt[idx] = wex ? a.tgt[(size_t)k * a.T + t] : (int8_t)(1 + ((tw >> lane) & 1u));
                    }
                    n += __popc(cw);
                }
            }
        }
        __syncwarp();
        int64_t top = INT64_MIN;
        int top_j = -1;
        if (n > 2) {
            for (int j = 0; j < n; ++j) {
                const int32_t pj = P[j];
                const int gj = Gt[j];
                int64_t best = INT64_MIN;
                int bi = INT32_MAX;
                for (int i = lane; i < j; i += 32) {
                    const int64_t d = (int64_t)pj - (int64_t)P[i];
                    int qd = gj - (int)Gt[i];
                    qd = qd < 0 ? -qd : qd;
                    if (d >= kMinPairDistance && qd <= a.max_mm) {
                        const int64_t s = (qd == 0) ? d + a.bonus : (int64_t)a.penalty;
                        const int64_t cand = s + max(M[i], (int64_t)0);
                        if (cand > best) { best = cand; bi = i; }  // ascending i: the lane's first maximiser
                    }
                }
                const int64_t m = set_warp_max_i64(best);
                const int first = __reduce_min_sync(0xffffffffu, (best == m && bi != INT32_MAX) ? bi : INT32_MAX);
                if (lane == 0) {
                    M[j] = m;
                    // extension of the chain ending at `first`, or a fresh start with the pair (first, j)
                    B[j] = (first == INT32_MAX) ? INT32_MIN : (M[first] >= 0 ? first : -(first + 2));
                }
                if (m > top) { top = m; top_j = j; }
                __syncwarp();
            }
        }
        if (!kFill) {
            if (lane == 0) {
                int cnt = 0;
                if (top_j >= 0) {
                    int j = top_j;
                    for (;;) {
                        ++cnt;
                        const int b = B[j];
                        if (b >= 0) j = b; else { ++cnt; break; }
                    }
                }
                a.score[it] = top;
                a.set_len[it] = cnt;
            }
        } else if (lane == 0 && top_j >= 0) {
            int k = len - 1, j = top_j;
            while (k >= 0) {
                a.set_pos[off + k] = P[j];
                --k;
                const int b = B[j];
                if (b >= 0) j = b;
                else { if (k >= 0) a.set_pos[off + k] = P[-b - 2]; break; }
            }
        }
        __syncwarp();
    }
}

// ---- exclusive scan int32 [n] -> int64 [n + 1] ------------------------------------------------
__device__ __forceinline__ int64_t scan_block_excl(int64_t v, int64_t *total, int64_t *s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int64_t tt = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, tt, d);
            if (lane >= d) tt += y;
        }
        s_warp[lane] = tt;  // inclusive over warps
    }
    __syncthreads();
    *total = s_warp[(blockDim.x >> 5) - 1];
    const int64_t r = x - v + (warp ? s_warp[warp - 1] : 0);
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_sum_kernel(const int32_t *v, int64_t n, int64_t *block_sum) {
    __shared__ int64_t s_warp[32];
    const int64_t i0 = (int64_t)blockIdx.x * kScanPerBlock + (int64_t)threadIdx.x * kScanPerThread;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) s += (i0 + k < n) ? v[i0 + k] : 0;
    int64_t total;
    scan_block_excl(s, &total, s_warp);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = total;
}

// in-place exclusive scan of the block sums, one CTA; the grand total goes to *total_out
__global__ void __launch_bounds__(1024) scan_blocks_kernel(int64_t *block_sum, int64_t nb, int64_t *total_out) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < nb ? block_sum[i] : 0;
        int64_t total;
        const int64_t ex = scan_block_excl(v, &total, s_warp);
        const int64_t carry = s_carry;
        if (i < nb) block_sum[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = s_carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const int32_t *v, int64_t n, const int64_t *block_base,
                                                                  int64_t *off) {
    __shared__ int64_t s_warp[32];
    const int64_t i0 = (int64_t)blockIdx.x * kScanPerBlock + (int64_t)threadIdx.x * kScanPerThread;
    int32_t x[kScanPerThread];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) { x[k] = (i0 + k < n) ? v[i0 + k] : 0; s += x[k]; }
    int64_t total;
    int64_t run = block_base[blockIdx.x] + scan_block_excl(s, &total, s_warp);
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        if (i0 + k < n) off[i0 + k] = run;
        run += x[k];
    }
}

}  // namespace sstar

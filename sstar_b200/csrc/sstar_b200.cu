// sstar_b200 -- hand-written sm_100a CUDA kernels + C-ABI for the S* hot path.
//
// Pipeline (all integer work, HBM/ALU bound, no tensor cores -- see DESIGN.md):
//
//   K0 window_bounds   pos + closed windows            -> variant index range [lo, hi) per window
//                      (sstar/generators/window_data_generator.py:113, the O(W*V) mask)
//   K1 pack            int8 genotype matrices, read once-> absent mask + 2 bit-planes per column
//                      (sstar/sstar.py:87-92 signed ref row-sum; :163-164 carried mask)
//   K2 score_prefix    one THREAD per (window, column)  -> S* by an exact O(n) prefix-max form of
//                      the chain DP, valid when carried genotypes are in {1, 2}
//   K3 score_pairwise  one WARP per (window, column)    -> the literal O(n^2) recurrence of
//                      sstar/sstar.py:195-217 with a REDUX row max; any int8 genotype
//
// Results are bit-identical to the reference: scores are exact integers (int64), NaN is
// reported as SSTAR_B200_NAN.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "sstar_b200.h"

namespace {

constexpr int kMinPairDistance = 10;  // sstar/sstar.py:195
constexpr int64_t kNegInf = INT64_MIN;
constexpr int64_t kNegHalf = -(int64_t(1) << 61);  // anything below is "-inf" in the prefix DP
constexpr int64_t kNegInit = -(int64_t(1) << 62);

constexpr int kPackThreads = 256;
constexpr int kScoreThreads = 128;
constexpr int kRingSlots = 16;  // >= 10: at most 10 distinct integer positions lie within 9 bp
constexpr int kPairWarps = 4;
constexpr int kPairSmemCap = 1024;  // list entries per warp kept in shared memory

thread_local char g_err[512] = "";
thread_local int g_launches = 0;

int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            snprintf(g_err, sizeof(g_err), "%s failed: %s", #expr, cudaGetErrorString(e__)); \
            return SSTAR_B200_ECUDA;                                                \
        }                                                                           \
    } while (0)

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
struct WsHeader {  // first 64 bytes of the workspace
    uint32_t status;
    int32_t max_win_variants;  // device-computed max(hi - lo)
    int32_t n_slow;            // windows routed to the pairwise kernel
    int32_t pad[13];
};

struct WsLayout {
    int64_t G;    // 32-variant groups
    int32_t Tp;   // plane row pitch in words (T rounded up to 32)
    size_t off_absent, off_exotic, off_carried, off_two, off_win_lo, off_win_hi, off_slow, off_arena;
    size_t arena_entries;
    size_t total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

WsLayout make_layout(int64_t V, int32_t T, int32_t W, int32_t max_win_variants) {
    WsLayout L;
    L.G = (V + 31) / 32;
    L.Tp = (int32_t)align_up((size_t)(T > 0 ? T : 1), 32);
    size_t off = align_up(sizeof(WsHeader), 256);
    size_t g = (size_t)(L.G > 0 ? L.G : 1);
    L.off_absent = off;  off = align_up(off + g * 4, 256);
    L.off_exotic = off;  off = align_up(off + g * 4, 256);
    L.off_carried = off; off = align_up(off + g * (size_t)L.Tp * 4, 256);
    L.off_two = off;     off = align_up(off + g * (size_t)L.Tp * 4, 256);
    size_t w = (size_t)(W > 0 ? W : 1);
    L.off_win_lo = off;  off = align_up(off + w * 4, 256);
    L.off_win_hi = off;  off = align_up(off + w * 4, 256);
    L.off_slow = off;    off = align_up(off + w * 4, 256);
    // pairwise list arena: at least one slab of the worst-case window must fit
    size_t worst = (size_t)(V > 0 ? V : 1);
    size_t want;
    if (max_win_variants > 0) {
        size_t cap = align_up((size_t)max_win_variants, 8);
        want = cap * 592;  // 148 SMs x 4 warps
        if (want > (size_t)16 << 20) want = (size_t)16 << 20;
        if (want < cap) want = cap;
    } else {
        want = worst > ((size_t)4 << 20) ? worst : ((size_t)4 << 20);
        if (want > worst * 592) want = worst * 592;
    }
    L.arena_entries = align_up(want, 8);
    L.off_arena = off;   off = align_up(off + L.arena_entries * 13, 256);
    L.total = off;
    return L;
}

// ---------------------------------------------------------------------------------------------
// K0: window bounds.  One thread per window, lower_bound on the (segment of the) position array.
// ---------------------------------------------------------------------------------------------
struct BoundsArgs {
    const int32_t *pos;
    const int32_t *win_start, *win_end;  // [W] or null (then fixed_start/fixed_end)
    const int32_t *seg;                  // [W+1] search segment per window, or null (0..V)
    int32_t fixed_start, fixed_end;
    int32_t V, W;
    int32_t *win_lo, *win_hi, *slow_list;
    WsHeader *hdr;
    int force_slow;
};

__device__ __forceinline__ int32_t lower_bound_dev(const int32_t *__restrict__ pos, int32_t lo,
                                                   int32_t hi, int64_t x) {
    while (lo < hi) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if ((int64_t)__ldg(pos + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void window_bounds_kernel(BoundsArgs a) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    int span = 0;
    if (w < a.W) {
        int32_t sb = a.seg ? a.seg[w] : 0;
        int32_t se = a.seg ? a.seg[w + 1] : a.V;
        int64_t s = a.win_start ? a.win_start[w] : a.fixed_start;
        int64_t e = a.win_end ? a.win_end[w] : a.fixed_end;
        int32_t lo = lower_bound_dev(a.pos, sb, se, s);
        int32_t hi = lower_bound_dev(a.pos, lo, se, e + 1);
        if (hi < lo) hi = lo;
        a.win_lo[w] = lo;
        a.win_hi[w] = hi;
        span = hi - lo;
        if (a.force_slow) a.slow_list[w] = w;
    }
    span = __reduce_max_sync(0xffffffffu, span);
    if ((threadIdx.x & 31) == 0 && span > 0) atomicMax(&a.hdr->max_win_variants, span);
    if (a.force_slow && w == 0) a.hdr->n_slow = a.W;
}

// ---------------------------------------------------------------------------------------------
// K1: pack.  One CTA per `gpc` groups of 32 variants.  The CTA's rows of both matrices are one
// contiguous byte span each (row-major), staged into shared memory with 128-bit loads whatever
// R and T are; bit-planes come out word-major [G][Tp] so K2's lanes (= columns) read coalesced.
// ---------------------------------------------------------------------------------------------
struct PackArgs {
    const int8_t *ref, *tgt;
    int32_t V, R, T, Tp;
    int32_t G, gpc;
    int stage;                  // 1: tile staged in smem; 0: read global directly (huge R+T)
    uint32_t ref_tile_bytes;    // smem bytes reserved for the ref tile (incl. alignment slack)
    uint32_t tgt_tile_bytes;
    uint32_t *absent_bits, *exotic, *carried, *two;
};

__device__ __forceinline__ uint4 ldg_stream_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Copies nbytes from global to shared keeping the source's 16-byte phase, so the body moves as
// aligned 128-bit vectors for ANY base pointer / row pitch.  Returns where byte 0 landed.
__device__ __forceinline__ const int8_t *stage_span(uint8_t *sdst, const int8_t *gsrc, size_t nbytes,
                                                    int tid, int nthreads) {
    const uint32_t phase = (uint32_t)((uintptr_t)gsrc & 15);
    uint8_t *s0 = sdst + phase;
    size_t head = (16 - phase) & 15;
    if (head > nbytes) head = nbytes;
    if ((size_t)tid < head) s0[tid] = (uint8_t)gsrc[tid];
    const size_t nvec = (nbytes - head) >> 4;
    const uint4 *gv = reinterpret_cast<const uint4 *>(gsrc + head);
    uint4 *sv = reinterpret_cast<uint4 *>(s0 + head);
    for (size_t i = tid; i < nvec; i += nthreads) sv[i] = ldg_stream_v4(gv + i);
    const size_t done = head + (nvec << 4);
    if (done + tid < nbytes) s0[done + tid] = (uint8_t)gsrc[done + tid];  // tail < 16 bytes
    return reinterpret_cast<const int8_t *>(s0);
}

__global__ void __launch_bounds__(kPackThreads) pack_kernel(PackArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kPackThreads / 32;
    const int32_t g0 = blockIdx.x * a.gpc;
    const int ng = min(a.gpc, a.G - g0);
    const int32_t k0 = g0 * 32;
    const int nrows = min(ng * 32, a.V - k0);

    uint8_t *s_absent = smem;                                   // [gpc*32]
    uint32_t *s_flags = reinterpret_cast<uint32_t *>(smem + a.gpc * 32);  // [gpc]
    uint8_t *s_tiles = smem + a.gpc * 32 + align16(a.gpc * 4);
    const int8_t *rbase = a.ref + (size_t)k0 * a.R;
    const int8_t *tbase = a.tgt + (size_t)k0 * a.T;
    if (a.stage) {
        rbase = stage_span(s_tiles, rbase, (size_t)nrows * a.R, tid, kPackThreads);
        tbase = stage_span(s_tiles + a.ref_tile_bytes, tbase, (size_t)nrows * a.T, tid, kPackThreads);
    }
    if (tid < ng) s_flags[tid] = 0;
    __syncthreads();

    // A: signed row sums of the reference panel -> absent flag per variant (sstar.py:87-88)
    for (int r = warp; r < ng * 32; r += nwarps) {
        int sum = 0;
        if (r < nrows) {
            const int8_t *row = rbase + (size_t)r * a.R;
            for (int c = lane; c < a.R; c += 32) sum += row[c];
        }
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) s_absent[r] = (r < nrows) && (sum == 0);
    }
    __syncthreads();

    // B: per (group, column): 32 rows -> carried / genotype==2 bit words; flag other genotypes
    const int nitems = ng * a.T;
    for (int it = tid; it < nitems; it += kPackThreads) {
        const int gl = it / a.T, t = it - gl * a.T;
        const int rmax = min(32, nrows - gl * 32);
        const int8_t *col = tbase + (size_t)(gl * 32) * a.T + t;
        uint32_t car = 0, two = 0, ex = 0;
        for (int r = 0; r < rmax; ++r) {
            const int b = col[(size_t)r * a.T];
            const uint32_t ab = s_absent[gl * 32 + r];
            car |= ((uint32_t)(b != 0) & ab) << r;
            two |= ((uint32_t)(b == 2) & ab) << r;
            ex |= (uint32_t)((unsigned)b > 2u) & ab;
        }
        const size_t o = (size_t)(g0 + gl) * a.Tp + t;
        a.carried[o] = car;
        a.two[o] = two;
        if (ex) atomicOr(&s_flags[gl], 1u);
    }
    __syncthreads();

    // C: absent bit word + exotic flag per group
    for (int gl = warp; gl < ng; gl += nwarps) {
        const uint32_t bits = __ballot_sync(0xffffffffu, s_absent[gl * 32 + lane] != 0);
        if (lane == 0) {
            a.absent_bits[g0 + gl] = bits;
            a.exotic[g0 + gl] = s_flags[gl];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K2: prefix-max chain DP, one thread per (window, column).
//
// With v_i = max(M[i], 0) the recurrence of sstar.py:208-212 is M[j] = max_{i<j} s(j,i) + v_i.
// For carried genotypes in {1,2}:   same genotype  -> s = (p_j - p_i) + bonus
//                                   other genotype -> s = penalty            (if max_mismatch >= 1)
// so  M[j] = max( p_j + bonus + max_{i: g_i = g_j}(v_i - p_i),  penalty + max_{i: g_i != g_j} v_i )
// over the i with p_j - p_i >= 10.  Positions ascend, so eligible i form a prefix: sites enter the
// four running maxima through a small per-thread ring once they are >= 10 bp behind.
// ---------------------------------------------------------------------------------------------
struct ScoreArgs {
    const int32_t *pos;
    const int8_t *tgt;
    const uint32_t *absent_bits, *exotic, *carried, *two;
    const int32_t *win_lo, *win_hi;
    int32_t *slow_list;
    WsHeader *hdr;
    int32_t T, Tp, W, tchunks;
    int32_t bonus, max_mm, penalty;
    int64_t *score;
    int32_t *nsnp;
    // pairwise only
    uint8_t *arena;
    uint64_t arena_entries;
};

__global__ void __launch_bounds__(kScoreThreads) score_prefix_kernel(ScoreArgs a) {
    __shared__ int32_t ring_p[kRingSlots][kScoreThreads];
    __shared__ int64_t ring_v[kRingSlots][kScoreThreads];
    const int tid = threadIdx.x;
    const int w = blockIdx.x / a.tchunks;
    const int t = (blockIdx.x - w * a.tchunks) * kScoreThreads + tid;
    if (t >= a.T) return;
    const int32_t lo = a.win_lo[w], hi = a.win_hi[w];

    int64_t bestA0 = kNegInit, bestA1 = kNegInit, bestB0 = kNegInit, bestB1 = kNegInit;
    int64_t top = kNegInit;
    int n = 0, nabs = 0, head = 0, cnt = 0;
    uint32_t ex = 0;
    const bool mm_ok = a.max_mm >= 1;

    if (hi > lo) {
        const int gi0 = lo >> 5, gi1 = (hi - 1) >> 5;
        for (int gi = gi0; gi <= gi1; ++gi) {
            uint32_t rm = 0xffffffffu;
            if (gi == gi0) rm &= 0xffffffffu << (lo & 31);
            if (gi == gi1) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
            nabs += __popc(__ldg(a.absent_bits + gi) & rm);
            ex |= __ldg(a.exotic + gi);
            uint32_t c = __ldg(a.carried + (size_t)gi * a.Tp + t) & rm;
            if (!c) continue;
            const uint32_t tw = __ldg(a.two + (size_t)gi * a.Tp + t);
            n += __popc(c);
            while (c) {
                const int b = __ffs(c) - 1;
                c &= c - 1;
                const int32_t p = __ldg(a.pos + (gi << 5) + b);
                const int cls = (tw >> b) & 1;
                // sites now >= 10 bp behind become eligible predecessors
                while (cnt > 0) {
                    const int32_t pi = ring_p[head][tid];
                    if (p - pi < kMinPairDistance) break;
                    const int64_t vv = ring_v[head][tid];
                    const int64_t v = vv >> 1;
                    if (vv & 1) { bestA1 = max(bestA1, v - pi); bestB1 = max(bestB1, v); }
                    else        { bestA0 = max(bestA0, v - pi); bestB0 = max(bestB0, v); }
                    head = (head + 1) & (kRingSlots - 1);
                    --cnt;
                }
                const int64_t same = cls ? bestA1 : bestA0;
                const int64_t other = cls ? bestB0 : bestB1;
                int64_t m = kNegInit;
                if (same > kNegHalf) m = same + p + a.bonus;
                if (mm_ok && other > kNegHalf) m = max(m, other + a.penalty);
                int64_t v = 0;
                if (m > kNegHalf) { top = max(top, m); v = max(m, (int64_t)0); }
                const int tail = (head + cnt) & (kRingSlots - 1);
                ring_p[tail][tid] = p;
                ring_v[tail][tid] = (v << 1) | cls;
                ++cnt;
            }
        }
    }
    if (ex) {  // uniform over the CTA: the whole window goes to the pairwise kernel
        if (tid == 0 && blockIdx.x == w * a.tchunks) {
            const int idx = atomicAdd(&a.hdr->n_slow, 1);
            a.slow_list[idx] = w;
        }
        return;
    }
    const size_t o = (size_t)w * a.T + t;
    a.score[o] = (n <= 2 || top <= kNegHalf) ? kNegInf : top;
    a.nsnp[o] = (hi - lo) - nabs + n;
}

// ---------------------------------------------------------------------------------------------
// K3: literal pairwise recurrence, one warp per (window, column), any int8 genotype.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t warp_max_i64(int64_t v) {
    // two REDUX steps: max of the high words, then max of the low words among the winners
    const int32_t hi = (int32_t)(v >> 32);
    const int32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t lo = (hi == mh) ? (uint32_t)v : 0u;
    const uint32_t ml = __reduce_max_sync(0xffffffffu, lo);
    return ((int64_t)mh << 32) | (int64_t)ml;
}

__global__ void __launch_bounds__(kPairWarps * 32) score_pairwise_kernel(ScoreArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_slow = a.hdr->n_slow;
    if (n_slow <= 0) return;
    const int64_t total = (int64_t)n_slow * a.T;
    const int32_t maxvw = a.hdr->max_win_variants;
    const uint32_t cap = (uint32_t)((maxvw + 7) & ~7);
    if (cap == 0) {  // every window is empty: all units are NaN with the ref-polymorphic count 0
        for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total;
             it += (int64_t)gridDim.x * blockDim.x) {
            const int w = a.slow_list[it / a.T];
            const size_t o = (size_t)w * a.T + (size_t)(it % a.T);
            a.score[o] = kNegInf;
            a.nsnp[o] = 0;
        }
        return;
    }
    int64_t gw = (int64_t)blockIdx.x * kPairWarps + warp;
    int64_t nw = (int64_t)gridDim.x * kPairWarps;
    uint8_t *slab;
    if (cap <= (uint32_t)kPairSmemCap) {
        slab = smem + (size_t)warp * kPairSmemCap * 13;
    } else {
        const int64_t nslabs = (int64_t)(a.arena_entries / cap);
        if (nslabs < nw) nw = nslabs;
        if (gw >= nw) return;
        slab = a.arena + (size_t)gw * cap * 13;
    }
    const uint32_t slab_cap = cap <= (uint32_t)kPairSmemCap ? (uint32_t)kPairSmemCap : cap;
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
                const bool inr = k >= lo && k < hi;
                const bool isab = inr && ((ab >> lane) & 1);
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
        int64_t top = kNegInf;
        if (n > 2) {
            for (int j = 0; j < n; ++j) {
                const int32_t pj = P[j];
                const int gj = Gt[j];
                int64_t best = kNegInf;
                for (int i = lane; i < j; i += 32) {
                    const int64_t d = (int64_t)pj - (int64_t)P[i];
                    int q = gj - (int)Gt[i];
                    q = q < 0 ? -q : q;
                    if (d >= kMinPairDistance && q <= a.max_mm) {
                        const int64_t s = (q == 0) ? d + a.bonus : (int64_t)a.penalty;
                        const int64_t mi = M[i];
                        best = max(best, s + max(mi, (int64_t)0));  // == max(M[i] + s, s)
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
            a.score[o] = top;  // kNegInf == SSTAR_B200_NAN (n <= 2 or no valid pair)
            a.nsnp[o] = (hi - lo) - nabs + n;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

int check_shapes(int64_t V, int32_t R, int32_t T, int32_t W) {
    if (V < 0 || V > (int64_t)INT32_MAX - 64) return fail(SSTAR_B200_EINVAL, "V out of range%s");
    if (R < 1) return fail(SSTAR_B200_EINVAL, "R must be >= 1%s");
    if (T < 1) return fail(SSTAR_B200_EINVAL, "T must be >= 1%s");
    if (W < 0) return fail(SSTAR_B200_EINVAL, "W must be >= 0%s");
    return 0;
}

int launch_pack(cudaStream_t st, const int8_t *d_ref, const int8_t *d_tgt, int64_t V, int32_t R,
                int32_t T, uint8_t *ws, const WsLayout &L) {
    if (V == 0) return 0;
    PackArgs p;
    p.ref = d_ref; p.tgt = d_tgt;
    p.V = (int32_t)V; p.R = R; p.T = T; p.Tp = L.Tp; p.G = (int32_t)L.G;
    p.absent_bits = reinterpret_cast<uint32_t *>(ws + L.off_absent);
    p.exotic = reinterpret_cast<uint32_t *>(ws + L.off_exotic);
    p.carried = reinterpret_cast<uint32_t *>(ws + L.off_carried);
    p.two = reinterpret_cast<uint32_t *>(ws + L.off_two);
    // groups per CTA: keep the staged tile under ~64 KB, at most 8 groups, and keep >= ~4 CTAs per SM
    const size_t per_group = (size_t)32 * ((size_t)R + (size_t)T);
    int gpc = (int)((size_t)(64 << 10) / (per_group ? per_group : 1));
    if (gpc > 8) gpc = 8;
    while (gpc > 1 && (L.G + gpc - 1) / gpc < 592) --gpc;
    if (gpc < 1) gpc = 1;
    p.gpc = gpc;
    p.ref_tile_bytes = (uint32_t)align_up((size_t)gpc * 32 * R + 32, 16);
    p.tgt_tile_bytes = (uint32_t)align_up((size_t)gpc * 32 * T + 32, 16);
    size_t fixed = (size_t)gpc * 32 + align_up((size_t)gpc * 4, 16);
    size_t smem = fixed + p.ref_tile_bytes + p.tgt_tile_bytes;
    p.stage = 1;
    if (smem > (size_t)200 << 10) {  // rows too wide to stage: read global directly
        p.stage = 0;
        smem = fixed;
    }
    cudaError_t e = cudaFuncSetAttribute(pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaFuncSetAttribute(pack): %s", cudaGetErrorString(e));
    const unsigned grid = (unsigned)((L.G + gpc - 1) / gpc);
    pack_kernel<<<grid, kPackThreads, smem, st>>>(p);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "pack_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

int launch_score(cudaStream_t st, const int8_t *d_tgt, const int32_t *d_pos, int64_t V, int32_t T,
                 const int32_t *d_ws, const int32_t *d_we, const int32_t *d_seg, int32_t fixed_s,
                 int32_t fixed_e, int32_t W, int32_t bonus, int32_t max_mm, int32_t penalty,
                 int64_t *d_score, int32_t *d_nsnp, uint8_t *ws, const WsLayout &L, int algo) {
    if (W == 0) return 0;
    WsHeader *hdr = reinterpret_cast<WsHeader *>(ws);
    cudaError_t e = cudaMemsetAsync(hdr, 0, sizeof(WsHeader), st);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaMemsetAsync(header): %s", cudaGetErrorString(e));

    BoundsArgs b;
    b.pos = d_pos; b.win_start = d_ws; b.win_end = d_we; b.seg = d_seg;
    b.fixed_start = fixed_s; b.fixed_end = fixed_e;
    b.V = (int32_t)V; b.W = W;
    b.win_lo = reinterpret_cast<int32_t *>(ws + L.off_win_lo);
    b.win_hi = reinterpret_cast<int32_t *>(ws + L.off_win_hi);
    b.slow_list = reinterpret_cast<int32_t *>(ws + L.off_slow);
    b.hdr = hdr;
    b.force_slow = (algo == SSTAR_B200_ALGO_PAIRWISE);
    window_bounds_kernel<<<(W + 127) / 128, 128, 0, st>>>(b);
    ++g_launches;

    ScoreArgs s;
    s.pos = d_pos; s.tgt = d_tgt;
    s.absent_bits = reinterpret_cast<const uint32_t *>(ws + L.off_absent);
    s.exotic = reinterpret_cast<const uint32_t *>(ws + L.off_exotic);
    s.carried = reinterpret_cast<const uint32_t *>(ws + L.off_carried);
    s.two = reinterpret_cast<const uint32_t *>(ws + L.off_two);
    s.win_lo = b.win_lo; s.win_hi = b.win_hi; s.slow_list = b.slow_list; s.hdr = hdr;
    s.T = T; s.Tp = L.Tp; s.W = W;
    s.tchunks = (T + kScoreThreads - 1) / kScoreThreads;
    s.bonus = bonus; s.max_mm = max_mm; s.penalty = penalty;
    s.score = d_score; s.nsnp = d_nsnp;
    s.arena = ws + L.off_arena; s.arena_entries = L.arena_entries;

    if (algo != SSTAR_B200_ALGO_PAIRWISE) {
        const int64_t grid = (int64_t)W * s.tchunks;
        if (grid > INT32_MAX) return fail(SSTAR_B200_EINVAL, "W * ceil(T/128) exceeds the grid limit%s");
        score_prefix_kernel<<<(unsigned)grid, kScoreThreads, 0, st>>>(s);
        ++g_launches;
    }
    const size_t psmem = (size_t)kPairWarps * kPairSmemCap * 13;
    e = cudaFuncSetAttribute(score_pairwise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaFuncSetAttribute(pairwise): %s", cudaGetErrorString(e));
    int64_t units = (int64_t)W * T;
    int64_t pgrid = (units + kPairWarps - 1) / kPairWarps;
    if (pgrid > 148 * 4) pgrid = 148 * 4;
    score_pairwise_kernel<<<(unsigned)pgrid, kPairWarps * 32, psmem, st>>>(s);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "score kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

int check_common(int device, int64_t V, int32_t R, int32_t T, int32_t W, const void *ws,
                 size_t ws_bytes, const sstar_b200_options *opts, WsLayout *L, int *algo,
                 bool pack_only = false) {
    g_launches = 0;
    int rc = check_shapes(V, R, T, W);
    if (rc) return rc;
    *algo = opts ? opts->algo : SSTAR_B200_ALGO_AUTO;
    if (*algo != SSTAR_B200_ALGO_AUTO && *algo != SSTAR_B200_ALGO_PAIRWISE)
        return fail(SSTAR_B200_EINVAL, "unknown algo%s");
    const int32_t hint = opts ? opts->max_win_variants : 0;
    if (hint < 0) return fail(SSTAR_B200_EINVAL, "max_win_variants must be >= 0%s");
    *L = make_layout(V, T, W, hint);
    if (!ws) return fail(SSTAR_B200_EINVAL, "workspace pointer is null%s");
    // the packed planes sit at the front of the workspace and do not depend on W / the hint
    if (ws_bytes < (pack_only ? L->off_win_lo : L->total))
        return fail(SSTAR_B200_EWORKSPACE, "workspace too small%s");
    if (((uintptr_t)ws & 255) != 0) return fail(SSTAR_B200_EINVAL, "workspace must be 256-byte aligned%s");
    (void)device;
    return 0;
}

}  // namespace

extern "C" {

int sstar_b200_version(void) { return SSTAR_B200_ABI_VERSION; }

const char *sstar_b200_last_error(void) { return g_err; }

int sstar_b200_last_launch_count(void) { return g_launches; }

int sstar_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

size_t sstar_b200_workspace_bytes(int64_t V, int32_t R, int32_t T, int32_t W, int32_t max_win_variants) {
    (void)R;
    if (V < 0 || T < 1 || W < 0) return 0;
    return make_layout(V, T, W, max_win_variants < 0 ? 0 : max_win_variants).total;
}

int sstar_b200_pack(int device, void *stream, const int8_t *d_ref_gt, const int8_t *d_tgt_gt, int64_t V,
                    int32_t R, int32_t T, void *d_workspace, size_t workspace_bytes) {
    WsLayout L; int algo;
    int rc = check_common(device, V, R, T, 0, d_workspace, workspace_bytes, nullptr, &L, &algo, true);
    if (rc) return rc;
    if (V > 0 && (!d_ref_gt || !d_tgt_gt)) return fail(SSTAR_B200_EINVAL, "null genotype pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    return launch_pack((cudaStream_t)stream, d_ref_gt, d_tgt_gt, V, R, T, (uint8_t *)d_workspace, L);
}

int sstar_b200_score_packed(int device, void *stream, const int8_t *d_tgt_gt, const int32_t *d_pos,
                            int64_t V, int32_t R, int32_t T, const int32_t *d_win_start,
                            const int32_t *d_win_end, int32_t W, int32_t match_bonus,
                            int32_t max_mismatch, int32_t mismatch_penalty, int64_t *d_score,
                            int32_t *d_nsnp, void *d_workspace, size_t workspace_bytes,
                            const sstar_b200_options *opts) {
    WsLayout L; int algo;
    int rc = check_common(device, V, R, T, W, d_workspace, workspace_bytes, opts, &L, &algo);
    if (rc) return rc;
    if (W > 0 && (!d_win_start || !d_win_end || !d_score || !d_nsnp))
        return fail(SSTAR_B200_EINVAL, "null window/output pointer%s");
    if (V > 0 && (!d_pos || !d_tgt_gt)) return fail(SSTAR_B200_EINVAL, "null pos/tgt pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    return launch_score((cudaStream_t)stream, d_tgt_gt, d_pos, V, T, d_win_start, d_win_end, nullptr, 0, 0,
                        W, match_bonus, max_mismatch, mismatch_penalty, d_score, d_nsnp,
                        (uint8_t *)d_workspace, L, algo);
}

int sstar_b200_score_windows_ex(int device, void *stream, const int8_t *d_ref_gt, const int8_t *d_tgt_gt,
                                const int32_t *d_pos, int64_t V, int32_t R, int32_t T,
                                const int32_t *d_win_start, const int32_t *d_win_end, int32_t W,
                                int32_t match_bonus, int32_t max_mismatch, int32_t mismatch_penalty,
                                int64_t *d_score, int32_t *d_nsnp, void *d_workspace,
                                size_t workspace_bytes, const sstar_b200_options *opts) {
    WsLayout L; int algo;
    int rc = check_common(device, V, R, T, W, d_workspace, workspace_bytes, opts, &L, &algo);
    if (rc) return rc;
    if (W > 0 && (!d_win_start || !d_win_end || !d_score || !d_nsnp))
        return fail(SSTAR_B200_EINVAL, "null window/output pointer%s");
    if (V > 0 && (!d_pos || !d_tgt_gt || !d_ref_gt)) return fail(SSTAR_B200_EINVAL, "null input pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_pack(st, d_ref_gt, d_tgt_gt, V, R, T, (uint8_t *)d_workspace, L);
    if (rc) return rc;
    return launch_score(st, d_tgt_gt, d_pos, V, T, d_win_start, d_win_end, nullptr, 0, 0, W, match_bonus,
                        max_mismatch, mismatch_penalty, d_score, d_nsnp, (uint8_t *)d_workspace, L, algo);
}

int sstar_b200_score_windows(int device, void *stream, const int8_t *d_ref_gt, const int8_t *d_tgt_gt,
                             const int32_t *d_pos, int64_t V, int32_t R, int32_t T,
                             const int32_t *d_win_start, const int32_t *d_win_end, int32_t W,
                             int32_t match_bonus, int32_t max_mismatch, int32_t mismatch_penalty,
                             int64_t *d_score, int32_t *d_nsnp, void *d_workspace, size_t workspace_bytes) {
    return sstar_b200_score_windows_ex(device, stream, d_ref_gt, d_tgt_gt, d_pos, V, R, T, d_win_start,
                                       d_win_end, W, match_bonus, max_mismatch, mismatch_penalty, d_score,
                                       d_nsnp, d_workspace, workspace_bytes, nullptr);
}

int sstar_b200_score_replicates(int device, void *stream, const int8_t *d_ref_gt, const int8_t *d_tgt_gt,
                                const int32_t *d_pos, int64_t V, int32_t R, int32_t T,
                                const int32_t *d_rep_offset, int32_t Nrep, int32_t win_start,
                                int32_t win_end, int32_t match_bonus, int32_t max_mismatch,
                                int32_t mismatch_penalty, int64_t *d_score, int32_t *d_nsnp,
                                void *d_workspace, size_t workspace_bytes, const sstar_b200_options *opts) {
    WsLayout L; int algo;
    int rc = check_common(device, V, R, T, Nrep, d_workspace, workspace_bytes, opts, &L, &algo);
    if (rc) return rc;
    if (Nrep > 0 && (!d_rep_offset || !d_score || !d_nsnp))
        return fail(SSTAR_B200_EINVAL, "null replicate/output pointer%s");
    if (V > 0 && (!d_pos || !d_tgt_gt || !d_ref_gt)) return fail(SSTAR_B200_EINVAL, "null input pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t st = (cudaStream_t)stream;
    rc = launch_pack(st, d_ref_gt, d_tgt_gt, V, R, T, (uint8_t *)d_workspace, L);
    if (rc) return rc;
    return launch_score(st, d_tgt_gt, d_pos, V, T, nullptr, nullptr, d_rep_offset, win_start, win_end, Nrep,
                        match_bonus, max_mismatch, mismatch_penalty, d_score, d_nsnp,
                        (uint8_t *)d_workspace, L, algo);
}

}  // extern "C"

namespace {
struct HostScratch {
    size_t off_ref, off_tgt, off_pos, off_ws, off_we, off_score, off_nsnp, off_work, total;
};
HostScratch make_host_scratch(int64_t V, int32_t R, int32_t T, int32_t W, int32_t hint) {
    HostScratch h;
    size_t off = 0;
    size_t v = (size_t)(V > 0 ? V : 1), w = (size_t)(W > 0 ? W : 1);
    h.off_ref = off;   off = align_up(off + v * (size_t)R, 256);
    h.off_tgt = off;   off = align_up(off + v * (size_t)T, 256);
    h.off_pos = off;   off = align_up(off + v * 4, 256);
    h.off_ws = off;    off = align_up(off + w * 4, 256);
    h.off_we = off;    off = align_up(off + w * 4, 256);
    h.off_score = off; off = align_up(off + w * (size_t)T * 8, 256);
    h.off_nsnp = off;  off = align_up(off + w * (size_t)T * 4, 256);
    h.off_work = off;  off += make_layout(V, T, W, hint).total;
    h.total = off;
    return h;
}
}  // namespace

extern "C" size_t sstar_b200_host_scratch_bytes(int64_t V, int32_t R, int32_t T, int32_t W,
                                                int32_t max_win_variants) {
    if (V < 0 || R < 1 || T < 1 || W < 0) return 0;
    return make_host_scratch(V, R, T, W, max_win_variants < 0 ? 0 : max_win_variants).total;
}

extern "C" int sstar_b200_score_windows_host(int device, void *stream, const int8_t *h_ref_gt,
                                             const int8_t *h_tgt_gt, const int32_t *h_pos, int64_t V,
                                             int32_t R, int32_t T, const int32_t *h_win_start,
                                             const int32_t *h_win_end, int32_t W, int32_t match_bonus,
                                             int32_t max_mismatch, int32_t mismatch_penalty,
                                             int64_t *h_score, int32_t *h_nsnp, void *d_scratch,
                                             size_t scratch_bytes, const sstar_b200_options *opts) {
    g_launches = 0;
    int rc = check_shapes(V, R, T, W);
    if (rc) return rc;
    const int32_t hint = opts ? opts->max_win_variants : 0;
    if (hint < 0) return fail(SSTAR_B200_EINVAL, "max_win_variants must be >= 0%s");
    if (!d_scratch) return fail(SSTAR_B200_EINVAL, "scratch pointer is null%s");
    if (((uintptr_t)d_scratch & 255) != 0) return fail(SSTAR_B200_EINVAL, "scratch must be 256-byte aligned%s");
    const HostScratch h = make_host_scratch(V, R, T, W, hint);
    if (scratch_bytes < h.total) return fail(SSTAR_B200_EWORKSPACE, "scratch too small%s");
    if (W > 0 && (!h_win_start || !h_win_end || !h_score || !h_nsnp))
        return fail(SSTAR_B200_EINVAL, "null window/output pointer%s");
    if (V > 0 && (!h_pos || !h_tgt_gt || !h_ref_gt)) return fail(SSTAR_B200_EINVAL, "null input pointer%s");
    uint8_t *d = (uint8_t *)d_scratch;
    cudaStream_t st = (cudaStream_t)stream;
    {
        DeviceGuard guard(device);
        if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
        if (V > 0) {
            CUDA_TRY(cudaMemcpyAsync(d + h.off_pos, h_pos, (size_t)V * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d + h.off_ref, h_ref_gt, (size_t)V * R, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d + h.off_tgt, h_tgt_gt, (size_t)V * T, cudaMemcpyHostToDevice, st));
        }
        if (W > 0) {
            CUDA_TRY(cudaMemcpyAsync(d + h.off_ws, h_win_start, (size_t)W * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d + h.off_we, h_win_end, (size_t)W * 4, cudaMemcpyHostToDevice, st));
        }
    }
    rc = sstar_b200_score_windows_ex(device, stream, (const int8_t *)(d + h.off_ref),
                                     (const int8_t *)(d + h.off_tgt), (const int32_t *)(d + h.off_pos), V,
                                     R, T, (const int32_t *)(d + h.off_ws), (const int32_t *)(d + h.off_we),
                                     W, match_bonus, max_mismatch, mismatch_penalty,
                                     (int64_t *)(d + h.off_score), (int32_t *)(d + h.off_nsnp),
                                     d + h.off_work, scratch_bytes - h.off_work, opts);
    if (rc) return rc;
    if (W > 0) {
        DeviceGuard guard(device);
        CUDA_TRY(cudaMemcpyAsync(h_score, d + h.off_score, (size_t)W * T * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(h_nsnp, d + h.off_nsnp, (size_t)W * T * 4, cudaMemcpyDeviceToHost, st));
    }
    return 0;
}

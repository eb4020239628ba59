// sstar_b200 -- the ONE-LAUNCH path for small problems: row staging, packing, window bounds and
// the chain DP in a single kernel, no inter-kernel (or inter-CTA) edge.
//
// On a 10 Mb chromosome (BASELINE configs[1]: 5 MB of genotypes, 1,000 windows) the three-launch
// pipeline is bound by its dependent launches and by the planes' round trip through global memory,
// not by bytes: 23 us for what HBM moves in under 1 us.  Here the work is cut by ROWS, so a CTA
// knows what to load from its block index alone:
//   * CTA b owns the "home" rows [b*S, (b+1)*S) and, at its first instruction, hands the int8 rows
//     [b*S, b*S + S + H) of both matrices plus their positions to the TMA engine (cp.async.bulk +
//     mbarrier; one contiguous span per matrix; H = the largest window, so every window that starts
//     in the home rows ends inside the tile).  Neighbouring CTAs re-read the H halo rows through
//     the L2, which the whole input fits many times over.
//   * while the tile is in flight the CTA finds ITS windows -- those whose first variant is a home
//     row: pos[b*S - 1] < start <= pos[(b+1)*S - 1] -- by scanning the window list (any order,
//     duplicates allowed; O(W) compares per CTA, a few per thread at the sizes this path takes);
//   * the tile is packed in shared memory (a lane per reference row for the signed row sums ->
//     absent words; SIMD-within-word carried / genotype planes, four lanes per 32 x 4 block);
//   * window bounds are a 32-ary search in the STAGED positions (shared memory: no global latency),
//     and one thread per (window, column) unit walks the planes (dp_core.cuh: walk_dp_unit_reg;
//     walk_dp_unit_general when the window holds genotypes outside {0,1,2}); results are stored.
// Replicate batches (one window per replicate, disjoint row ranges, positions restarting) take
// the same kernel with a CTA per `wg` consecutive replicates: there no row is read twice.
//
// Whatever a CTA cannot do from its tile it still does ITSELF, from global memory, with bounded
// state: a window that does not end inside the tile (an understated size hint) and a unit with more
// than kMaxClasses distinct genotype values are streamed row by row through the class-table DP
// (dp_core.cuh: StreamDP, tables indexed by the genotype byte in shared memory), one warp per unit.
// So there is no slow list, no follow-up launch and no cross-CTA state.
#pragma once
#include "common.cuh"
#include "prep_kernels.cuh"
#include "score_kernels.cuh"

namespace sstar {

constexpr int kFusedThreads = 512;       // chromosome mode: one CTA per SM
constexpr int kFusedRepThreads = 256;    // replicate batches: more, smaller CTAs
constexpr int kFusedMaxSlots = 16;     // windows scored at a time: 512 threads / columns per window, a warp each for their bounds
constexpr int kFusedListCap = 1024;    // windows examined per round of the list scan
// shared memory the streaming fallback needs: per warp two class tables + the waiting ring, and a
// bitmask of reference-absent rows for one chunk of a window
constexpr uint32_t kFusedStreamWarpBytes = 2u * kClassTableSize * 8u + kRingSlots * (8u + 8u + 4u);
constexpr uint32_t kFusedStreamMinBytes = (kFusedThreads / 32) * kFusedStreamWarpBytes + 4096u + 4u * kFusedListCap * 4u;

struct FusedArgs {
    const int8_t *ref, *tgt;
    const int32_t *pos;
    const int32_t *win_start, *win_end;  // [W] (chromosome mode) or null (replicates: fixed_start / fixed_end)
    const int32_t *seg;                  // [W+1] row segment per window (replicates), or null
    int32_t fixed_start, fixed_end;
    int32_t V, R, T, W;
    int32_t home_rows;   // S: home rows per CTA, a multiple of 32 (chromosome mode)
    int32_t wg;          // replicates per CTA (replicate mode)
    int32_t tc;          // plane pitch in words: T rounded up to a multiple of 32
    int32_t cap_groups;  // tile capacity in 32-row groups
    uint32_t ref_tile_bytes, tgt_tile_bytes;  // shared memory reserved per matrix (incl. alignment slack)
    uint32_t smem_bytes;                      // the launch's dynamic shared memory
    int32_t bonus, max_mm, penalty;
    int32_t early_trigger;  // SSTAR_B200_FLAG_INDEPENDENT: the next flagged launch may start under this one
    int64_t *score;
    int32_t *nsnp;
};

// Phase timeline of the fused kernel (tools/exp_fused_timing.py builds with -DSSTAR_B200_FUSED_TIMING):
// thread 0 of every CTA stamps its SM's cycle counter at the phase edges (cheap; deltas inside a CTA only).
#ifdef SSTAR_B200_FUSED_TIMING
__device__ unsigned long long g_fused_stamps[4096 * 8];
#define FUSED_STAMP(i)                                                                            \
    do {                                                                                          \
        if (threadIdx.x == 0 && blockIdx.x < 4096) {                                              \
            g_fused_stamps[blockIdx.x * 8 + (i)] = (unsigned long long)clock64();                 \
        }                                                                                         \
    } while (0)
#else
#define FUSED_STAMP(i) do { } while (0)
#endif

struct LdsU32 { __device__ __forceinline__ uint32_t operator()(const uint32_t *p) const { return *p; } };
struct LdsPos {
    const int32_t *pos;
    __device__ __forceinline__ int64_t operator()(int32_t k) const { return (int64_t)pos[k]; }
};
struct LdsPos32 {  // positions as uint32: differences inside a window (< 2^32, never negative) stay exact
    const int32_t *pos;
    __device__ __forceinline__ uint32_t operator()(int32_t k) const { return (uint32_t)pos[k]; }
};
struct LdsGt {  // genotype byte of tile row k, column t (whole rows staged at `phase`)
    const uint8_t *tile;
    uint32_t phase, T, t;
    __device__ __forceinline__ int operator()(int32_t k) const { return (int)(int8_t)tile[phase + (uint32_t)k * T + t]; }
};

struct FusedShared {
    int32_t n_list, n_stream, max_hi;
};

// 32-ary lower bound in shared memory: first index in [lo, hi) with pos[idx] >= x
__device__ __forceinline__ int32_t warp_lower_bound_smem(const int32_t *pos, int32_t lo, int32_t hi, int64_t x, int lane) {
    while (hi - lo > 32) {
        const int32_t step = (hi - lo + 32) / 33;
        const int32_t idx = lo + step * (lane + 1) - 1;
        const bool below = idx < hi && (int64_t)pos[idx] < x;
        const int c = __popc(__ballot_sync(0xffffffffu, below));
        const int32_t new_hi = lo + step * (c + 1) - 1;
        if (c > 0) lo = lo + step * c;
        if (c < 32 && new_hi < hi) hi = new_hi;
    }
    const int32_t idx = lo + lane;
    const bool below = idx < hi && (int64_t)pos[idx] < x;
    return lo + __popc(__ballot_sync(0xffffffffu, below));
}

// ---- the streaming fallback: window w, any size, any genotypes; all threads of the CTA, shared
// memory from offset 0 (the tile is gone afterwards)
template <int kThreads>
__device__ __noinline__ void fused_stream_window(const FusedArgs &a, uint8_t *smem, int w) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kThreads / 32;
    __shared__ int32_t s_range[2];
    if (warp < 2) {  // the window's true row range, from global memory
        const int32_t sb = a.seg ? a.seg[w] : 0;
        const int32_t se = a.seg ? a.seg[w + 1] : a.V;
        const int64_t x = warp ? (int64_t)(a.win_end ? a.win_end[w] : a.fixed_end) + 1
                               : (int64_t)(a.win_start ? a.win_start[w] : a.fixed_start);
        const int32_t b = warp_lower_bound(a.pos, sb, se, x, lane);
        if (lane == 0) s_range[warp] = b;
    }
    __syncthreads();
    const int32_t lo = s_range[0], hi = max(s_range[0], s_range[1]);
    uint8_t *wbase = smem + (size_t)warp * kFusedStreamWarpBytes;
    StreamDP dp;
    dp.A = reinterpret_cast<int64_t *>(wbase);
    dp.B = dp.A + kClassTableSize;
    dp.wait_p = dp.B + kClassTableSize;
    dp.wait_v = dp.wait_p + kRingSlots;
    dp.wait_g = reinterpret_cast<int32_t *>(dp.wait_v + kRingSlots);
    uint32_t *mask = reinterpret_cast<uint32_t *>(smem + (size_t)nwarps * kFusedStreamWarpBytes);
    // (the window lists sit at the very end of the dynamic area and stay live)
    const uint32_t mask_words = (a.smem_bytes - nwarps * kFusedStreamWarpBytes - 4u * kFusedListCap * 4u) / 4u;
    const int32_t chunk_rows = (int32_t)min((uint32_t)0x7fffffe0u, mask_words * 32u);
    const int64_t p0 = hi > lo ? (int64_t)a.pos[lo] : 0;
    // units are streamed chunk by chunk; a unit's DP state lives in its warp's tables across the
    // chunks, so a round of `nwarps` units walks the whole window before the next round starts
    for (int t0 = 0; t0 < a.T; t0 += nwarps) {
        const int t = t0 + warp;
        const bool live = t < a.T;
        for (int i = lane; i < kClassTableSize; i += 32) { dp.A[i] = Sentinel<int64_t>::NEG; dp.B[i] = Sentinel<int64_t>::NEG; }
        dp.begin();
        int nabs = 0;
        __syncwarp();
        for (int32_t c0 = lo;; c0 += chunk_rows) {
            const int32_t c1 = (int32_t)min((int64_t)hi, (int64_t)c0 + chunk_rows);
            __syncthreads();  // the previous chunk's mask is done with
            // reference-absent rows of the chunk: a warp per row, lanes over the R columns
            for (int32_t i = tid; i < (c1 - c0 + 31) / 32; i += kThreads) mask[i] = 0u;
            __syncthreads();
            for (int32_t r = c0 + warp; r < c1; r += nwarps) {
                const int8_t *row = a.ref + (size_t)r * a.R;
                int sum = 0;
                for (int c = lane; c < a.R; c += 32) sum += row[c];
                sum = __reduce_add_sync(0xffffffffu, sum);
                if (lane == 0 && sum == 0) atomicOr(&mask[(r - c0) >> 5], 1u << ((r - c0) & 31));
            }
            __syncthreads();
            if (live) {
                for (int32_t k0 = c0; k0 < c1; k0 += 32) {
                    const int32_t k = k0 + lane;
                    const bool ab = k < c1 && ((mask[(k - c0) >> 5] >> ((k - c0) & 31)) & 1u);
                    const int gt = ab ? (int)a.tgt[(size_t)k * a.T + t] : 0;
                    const int64_t p = ab ? (int64_t)a.pos[k] - p0 : 0;
                    nabs += __popc(__ballot_sync(0xffffffffu, ab));
                    uint32_t car = __ballot_sync(0xffffffffu, ab && gt != 0);
                    while (car) {
                        const int src = __ffs(car) - 1;
                        car &= car - 1;
                        dp.site(__shfl_sync(0xffffffffu, p, src), __shfl_sync(0xffffffffu, gt, src), a.bonus, a.penalty,
                                a.max_mm);
                        __syncwarp();  // every lane runs the same site; identical values are stored
                    }
                }
            }
            if (c1 >= hi) break;
        }
        if (live && lane == 0) {
            const size_t o = (size_t)w * a.T + t;
            a.score[o] = dp.result();
            a.nsnp[o] = (hi - lo) - nabs + dp.n;
        }
    }
    __syncthreads();
}

// ---- pack of the staged tile: rows [0, nrows) of ng groups -> absent / exotic words and the two
// planes [ng][tc], all in shared memory.  Same arithmetic as pack_tile (prep_kernels.cuh), laid out
// for a CTA that has ONE tile to do and wants it done fast: a lane per reference row, four lanes per
// 32-row x 4-column block of the target rows.
template <int kThreads>
__device__ __forceinline__ void fused_pack(const FusedArgs &a, const uint8_t *s_ref, uint32_t rph, const uint8_t *s_tgt,
                                           uint32_t tph, int ng, int nrows, uint8_t *s_flag, uint32_t *s_absent,
                                           uint32_t *s_exotic, uint32_t *s_carried, uint32_t *s_two) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kThreads / 32;
    // A: signed row sums, one lane per row (consecutive rows are R bytes apart: for R = 100 that is 25
    // words, coprime with the 32 banks)
    {
        const int nwr = (a.R + 3) >> 2;
        const uint32_t last_mask = (a.R & 3) ? ((1u << (8 * (a.R & 3))) - 1u) : 0xffffffffu;
        const bool al = ((rph | (uint32_t)a.R) & 3u) == 0;
        for (int r = tid; r < ng * 32; r += kThreads) {
            int sum = 0;
            if (r < nrows) {
                const uint32_t row = rph + (uint32_t)r * (uint32_t)a.R;
                if (al) {
                    const uint32_t *vp = reinterpret_cast<const uint32_t *>(s_ref + row);
#pragma unroll 4
                    for (int i = 0; i < nwr; ++i) sum = __dp4a((int)vp[i], 0x01010101, sum);
                } else {
                    for (int i = 0; i < nwr; ++i) {
                        uint32_t w = lds_word<false>(s_ref, row + 4u * i);
                        if (i == nwr - 1) w &= last_mask;
                        sum = __dp4a((int)w, 0x01010101, sum);
                    }
                }
            }
            s_flag[r] = (r < nrows) && (sum == 0);
        }
    }
    __syncthreads();
    for (int gl = warp; gl < ng; gl += nwarps) {
        const uint32_t bits = __ballot_sync(0xffffffffu, s_flag[gl * 32 + lane] != 0);
        if (lane == 0) { s_absent[gl] = bits; s_exotic[gl] = 0u; }
    }
    __syncthreads();
    // B: a block of 32 rows x 4 columns per FOUR lanes: lane j folds rows 8j .. 8j+7 ("byte != 0" and
    // "byte == 2" of four columns into bit 0 of each byte), the four exchange their accumulators
    // and lane j transposes out the word of column j.
    const int nq = (a.T + 3) >> 2;
    const int nitems = ng * nq * 4;
    const bool aligned = ((tph | (uint32_t)a.T) & 3u) == 0;
    for (int it0 = warp * 32; it0 < nitems; it0 += kThreads) {
        const int it = it0 + lane;
        const bool valid = it < nitems;
        const int blk = it & 3, pair = it >> 2;
        const int gl = valid ? pair / nq : 0, q = valid ? pair - gl * nq : 0;
        const uint32_t base = tph + (uint32_t)(gl * 32 + blk * 8) * (uint32_t)a.T + 4u * q;
        uint32_t an = 0, at = 0, e_or = 0, e_and = 0;
        if (valid) {
#pragma unroll
            for (int rr = 7; rr >= 0; --rr) {  // row 7 first: it ends up in bit 7 of each byte
                const uint32_t off = base + (uint32_t)rr * (uint32_t)a.T;
                const uint32_t w = aligned ? lds_word<true>(s_tgt, off) : lds_word<false>(s_tgt, off);
                const uint32_t h = w >> 1;
                an = an + an + ((w | h) & 0x01010101u);
                at = at + at + (h & 0x01010101u);
                e_or |= w;
                e_and |= w & h;
            }
        }
        uint32_t bad = (e_or & 0xfcfcfcfcu) | (e_and & 0x01010101u);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 1);
        bad |= __shfl_xor_sync(0xffffffffu, bad, 2);
        uint32_t ax = 0;
        if (bad && valid) {
            // a byte outside {0,1,2} somewhere in the 32 x 4 block (missing calls, multi-allelic dosages -- or
            // the padding bytes of a ragged last quad / tail rows): the four lanes redo their rows with the
            // EXACT byte tests, plus the plane of bytes outside {0,1,2} for the group's exotic flag
            an = 0; at = 0;
#pragma unroll
            for (int rr = 7; rr >= 0; --rr) {
                const uint32_t off = base + (uint32_t)rr * (uint32_t)a.T;
                const uint32_t w = aligned ? lds_word<true>(s_tgt, off) : lds_word<false>(s_tgt, off);
                an = an + an + (bytes_nonzero(w) >> 7);
                at = at + at + (bytes_zero(w ^ 0x02020202u) >> 7);
                ax = ax + ax + ((bytes_nonzero(w & 0xfcfcfcfcu) | bytes_zero(w ^ 0x03030303u)) >> 7);
            }
        }
        const bool any_bad = __any_sync(0xffffffffu, bad != 0);
        const int l0 = lane & ~3;
        uint32_t nz[4], tw[4], xw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            nz[j] = __shfl_sync(0xffffffffu, an, l0 + j);
            tw[j] = __shfl_sync(0xffffffffu, at, l0 + j);
        }
        if (any_bad) {
#pragma unroll
            for (int j = 0; j < 4; ++j) xw[j] = __shfl_sync(0xffffffffu, ax, l0 + j);
        }
        uint32_t car[4], two[4], exo[4];
        transpose4(nz[0], nz[1], nz[2], nz[3], car);
        transpose4(tw[0], tw[1], tw[2], tw[3], two);
        transpose4(xw[0], xw[1], xw[2], xw[3], exo);
        const int c = blk;  // this lane's column within the quad
        if (valid && 4 * q + c < a.T) {
            const uint32_t ab = s_absent[gl];  // (rows past the matrix's end have no absent bit)
            const uint32_t m_nz = car[c], m_two = two[c];
            if (exo[c] & ab) atomicOr(&s_exotic[gl], 1u);
            s_carried[gl * a.tc + 4 * q + c] = m_nz & ab;
            s_two[gl * a.tc + 4 * q + c] = m_two & ab;
        }
    }
    __syncthreads();
}

template <int kThreads>
__device__ __forceinline__ void fused_body(const FusedArgs &a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ FusedShared g;
    __shared__ __align__(8) uint64_t s_bar[2];  // [0] the positions have landed, [1] the two matrices
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kThreads / 32;
    constexpr int kPerThread = (kFusedListCap + kThreads - 1) / kThreads;  // windows a thread examines per round
    const bool replicates = a.seg != nullptr;
    if (a.early_trigger) pdl_launch_dependents();  // (never followed by a wait: flagged calls are independent)
    FUSED_STAMP(0);

    // ---- shared-memory map
    uint8_t *s_ref = smem;
    uint8_t *s_tgt = s_ref + a.ref_tile_bytes;
    int32_t *s_pos = reinterpret_cast<int32_t *>(s_tgt + a.tgt_tile_bytes);
    uint32_t *s_carried = reinterpret_cast<uint32_t *>(s_pos + a.cap_groups * 32);
    uint32_t *s_two = s_carried + (size_t)a.cap_groups * a.tc;
    uint32_t *s_absent = s_two + (size_t)a.cap_groups * a.tc;
    uint32_t *s_exotic = s_absent + a.cap_groups;
    uint8_t *s_flag = reinterpret_cast<uint8_t *>(s_exotic + a.cap_groups);
    // the CTA's windows of a round: index, then (start, end) turned into (lo, hi); and the stream list --
    // at the very end of the dynamic area (the streaming fallback's bitmask stops before them)
    int32_t *s_wlist = reinterpret_cast<int32_t *>(smem + a.smem_bytes - 4 * kFusedListCap * 4);
    int32_t *s_wa = s_wlist + kFusedListCap;  // start -> lo
    int32_t *s_wb = s_wa + kFusedListCap;     // end -> hi; -1: the window does not end inside the tile, -2: a unit gave up
    int32_t *s_slist = s_wb + kFusedListCap;

    // ---- the tile: rows [k0, k0 + nrows)
    int32_t k0, want_rows, w_first = 0, w_count = 0;
    if (replicates) {
        w_first = (int)blockIdx.x * a.wg;
        w_count = min(a.wg, a.W - w_first);
        const int32_t rA = a.seg[w_first], rB = a.seg[w_first + w_count];
        k0 = rA & ~31;
        want_rows = ((rB - k0) + 31) & ~31;
    } else {
        k0 = (int)blockIdx.x * a.home_rows;
        want_rows = a.cap_groups * 32;
    }
    const int nrows = max(0, min(min(want_rows, a.cap_groups * 32), a.V - k0));
    const int8_t *gref = a.ref + (size_t)k0 * a.R;
    const int8_t *gtgt = a.tgt + (size_t)k0 * a.T;
    const int32_t *gpos = a.pos + k0;
    const uint32_t ref_bytes = (uint32_t)nrows * (uint32_t)a.R, tgt_bytes = (uint32_t)nrows * (uint32_t)a.T;
    const uint32_t pos_bytes = (uint32_t)nrows * 4u;
    const bool rbulk = nrows > 0 && (((uintptr_t)gref & 15) == 0) && ((ref_bytes & 15u) == 0);
    const bool tbulk = nrows > 0 && (((uintptr_t)gtgt & 15) == 0) && ((tgt_bytes & 15u) == 0);
    const bool pbulk = nrows > 0 && (((uintptr_t)gpos & 15) == 0) && ((pos_bytes & 15u) == 0);
    auto issue_tile = [&]() {  // thread 0: the three spans to the TMA engine, the (small) positions first
        // (a tile staged AGAIN lands on shared memory the streaming fallback wrote through the generic
        //  proxy: order those writes before the async proxy's)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (pbulk) {
            mbar_expect_tx(&s_bar[0], pos_bytes);
            bulk_g2s(s_pos, gpos, pos_bytes, &s_bar[0]);
        }
        if (rbulk || tbulk) {
            mbar_expect_tx(&s_bar[1], (rbulk ? ref_bytes : 0u) + (tbulk ? tgt_bytes : 0u));
            if (tbulk) bulk_g2s(s_tgt, gtgt, tgt_bytes, &s_bar[1]);
            if (rbulk) bulk_g2s(s_ref, gref, ref_bytes, &s_bar[1]);
        }
    };
    uint32_t rph = 0, tph = 0;
    auto stage_rest = [&]() {  // all threads: the spans the bulk engine cannot take
        if (!rbulk && nrows > 0) rph = stage_span(s_ref, gref, ref_bytes, tid, kThreads);
        if (!tbulk && nrows > 0) tph = stage_span(s_tgt, gtgt, tgt_bytes, tid, kThreads);
        if (!pbulk)
            for (int i = tid; i < nrows; i += kThreads) s_pos[i] = gpos[i];
    };
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
        issue_tile();  // the loads depend on nothing but the block index: they start right here
        g.n_list = 0;
        g.n_stream = 0;
        g.max_hi = 0;
    }
    // ---- chromosome mode: every global read the window scan needs is issued NOW, next to the tile:
    // the position just left of the home rows and this thread's share of the window coordinates
    int64_t p_left = INT64_MIN, p_right = INT64_MAX;  // windows with p_left < start <= p_right are this CTA's
    int32_t my_ws[kPerThread], my_we[kPerThread];
    auto load_coords = [&](int wb) {
#pragma unroll
        for (int j = 0; j < kPerThread; ++j) {
            const int w = wb + tid + j * kThreads;
            const bool in = w < min(a.W, wb + kFusedListCap);
            my_ws[j] = in ? a.win_start[w] : 0;
            my_we[j] = in ? a.win_end[w] : 0;
        }
    };
    if (!replicates) {
        if (k0 > 0) p_left = (int64_t)a.pos[k0 - 1];
        load_coords(0);
    }
    stage_rest();
    bool pend_pos = pbulk, pend_mat = rbulk || tbulk;  // bulk copies in flight: they must land before the CTA may exit
    uint32_t par_pos = 0, par_mat = 0;
    const int slots = kThreads / a.T;  // windows scored at a time
    const int slot = tid / a.T, t = tid - slot * a.T;
    int packed = 0;  // groups of the tile packed so far
    __syncthreads();  // barriers initialised, list counters zeroed, staged spans visible

    const int n_rounds = replicates ? (w_count + kFusedListCap - 1) / kFusedListCap : (a.W + kFusedListCap - 1) / kFusedListCap;
    for (int round = 0; round < n_rounds; ++round) {
        const int wb = round * kFusedListCap;
        if (round > 0 && !replicates) load_coords(wb);
        if (pend_pos) {  // the staged positions: the right edge of the home rows, and the bounds below
            mbar_wait(&s_bar[0], par_pos);
            par_pos ^= 1u;
            pend_pos = false;
        }
        // ---- this CTA's windows of the round, with their coordinates
        int n_list;
        if (replicates) {
            n_list = min(kFusedListCap, w_count - wb);
            for (int i = tid; i < n_list; i += kThreads) {
                s_wlist[i] = w_first + wb + i;
                s_wa[i] = a.fixed_start;
                s_wb[i] = a.fixed_end;
            }
            __syncthreads();
        } else {
            if ((int64_t)k0 + a.home_rows < (int64_t)a.V) p_right = (int64_t)s_pos[a.home_rows - 1];
#pragma unroll
            for (int j = 0; j < kPerThread; ++j) {
                const int w = wb + tid + j * kThreads;
                if (w < min(a.W, wb + kFusedListCap) && (int64_t)my_ws[j] > p_left && (int64_t)my_ws[j] <= p_right) {
                    const int i = atomicAdd(&g.n_list, 1);
                    s_wlist[i] = w;
                    s_wa[i] = my_ws[j];
                    s_wb[i] = my_we[j];
                }
            }
            __syncthreads();
            n_list = g.n_list;
        }
        FUSED_STAMP(1);
        if (n_list > 0) {
            // ---- bounds of every window of the round in the STAGED positions (a warp per window),
            // while the matrices are still landing; a window that may run past the tile is marked
            // for the streaming fallback
            for (int i = warp; i < n_list; i += nwarps) {
                const int w = s_wlist[i];
                int32_t sb = 0, se = nrows;
                bool inside = true;
                if (replicates) {
                    sb = a.seg[w] - k0;
                    se = a.seg[w + 1] - k0;
                    inside = se <= nrows;  // the replicate's rows are all staged
                    se = min(se, nrows);
                    sb = min(sb, se);
                }
                const int64_t xs = (int64_t)s_wa[i], xe = (int64_t)s_wb[i];
                const int32_t lo = warp_lower_bound_smem(s_pos, sb, se, xs, lane);
                int32_t hi = warp_lower_bound_smem(s_pos, lo, se, xe + 1, lane);
                if (hi < lo) hi = lo;
                if (!replicates && hi == nrows && k0 + nrows < a.V && nrows > 0 && (int64_t)s_pos[nrows - 1] <= xe)
                    inside = false;  // its last staged row is still inside the window, and rows follow
                __syncwarp();
                if (lane == 0) {
                    s_wa[i] = lo;
                    s_wb[i] = inside ? hi : -1;
                    if (inside && hi > lo) atomicMax(&g.max_hi, hi);
                }
            }
            __syncthreads();
            FUSED_STAMP(7);
            // ---- pack the rows the windows reach (a tile staged again after a fallback is packed again)
            const int need = (g.max_hi + 31) >> 5;
            if (need > packed) {
                if (pend_mat) {
                    mbar_wait(&s_bar[1], par_mat);
                    par_mat ^= 1u;
                    pend_mat = false;
                }
                FUSED_STAMP(2);
                fused_pack<kThreads>(a, s_ref, rph, s_tgt, tph, need, min(nrows, need * 32), s_flag, s_absent, s_exotic,
                                     s_carried, s_two);
                packed = need;
            }
            FUSED_STAMP(3);
        }
        for (int b0 = 0; b0 < n_list; b0 += slots) {
            // ---- one thread per (window, column) unit: #absent rows and the exotic flag of its window
            // from the words the pack left, then the walk DP
            const int i = b0 + slot;
            if (i < n_list && slot < slots) {
                const int32_t lo = s_wa[i], hi = s_wb[i];
                if (hi < 0) {
                    if (t == 0 && hi == -1) s_slist[atomicAdd(&g.n_stream, 1)] = s_wlist[i];
                } else {
                    int n = 0, nabs = 0;
                    int64_t score = kNegInf64;
                    bool done = true;
                    if (hi > lo) {
                        uint32_t ex = 0;
                        const int wi0 = lo >> 5, wi1 = (hi - 1) >> 5;
                        for (int gi = wi0; gi <= wi1; ++gi) {
                            uint32_t rm = 0xffffffffu;
                            if (gi == wi0) rm &= 0xffffffffu << (lo & 31);
                            if (gi == wi1) rm &= 0xffffffffu >> (31 - ((hi - 1) & 31));
                            nabs += __popc(s_absent[gi] & rm);
                            ex |= s_exotic[gi];
                        }
                        if (ex) {
                            done = walk_dp_unit_general(LdsU32(), LdsPos{s_pos}, LdsGt{s_tgt, tph, (uint32_t)a.T, (uint32_t)t},
                                                        s_carried, a.tc, t, lo, hi, a.bonus, a.penalty, a.max_mm, n, score);
                            // more than kMaxClasses genotype values in this unit: the window is streamed (queued once)
                            if (!done && atomicMin(&s_wb[i], -2) >= 0) s_slist[atomicAdd(&g.n_stream, 1)] = s_wlist[i];
                        } else {
                            const int64_t span = (int64_t)s_pos[hi - 1] - (int64_t)s_pos[lo];
                            if (score_bound(span, hi - lo, a.bonus, a.penalty) < kSmallBound)
                                walk_dp_unit_reg<int32_t>(LdsU32(), LdsPos32{s_pos}, s_carried, s_two, a.tc, t, lo, hi, a.bonus,
                                                          a.penalty, a.max_mm, n, score);
                            else
                                walk_dp_unit_reg<int64_t>(LdsU32(), LdsPos{s_pos}, s_carried, s_two, a.tc, t, lo, hi, a.bonus,
                                                          a.penalty, a.max_mm, n, score);
                        }
                    }
                    if (done) {
                        const size_t o = (size_t)s_wlist[i] * a.T + t;
                        a.score[o] = score;
                        a.nsnp[o] = (hi - lo) - nabs + n;
                    }
                }
            }
        }
        FUSED_STAMP(5);
        __syncthreads();
        // ---- windows the tile could not serve (whole window recomputed, whatever was stored before)
        const int n_stream = g.n_stream;
        if (n_stream > 0) {
            if (pend_mat) {  // the fallback takes over the tile's shared memory: nothing may still be landing in it
                mbar_wait(&s_bar[1], par_mat);
                par_mat ^= 1u;
                pend_mat = false;
            }
            __syncthreads();
            for (int i = 0; i < n_stream; ++i) fused_stream_window<kThreads>(a, smem, s_slist[i]);
            packed = 0;
            if (round + 1 < n_rounds) {  // (more rounds only beyond kFusedListCap windows) stage the tile again
                if (tid == 0) issue_tile();
                stage_rest();
                pend_pos = pbulk;
                pend_mat = rbulk || tbulk;
            }
        }
        if (round + 1 < n_rounds) {
            __syncthreads();
            if (tid == 0) { g.n_list = 0; g.n_stream = 0; g.max_hi = 0; }
            __syncthreads();
        }
    }
    if (pend_pos) mbar_wait(&s_bar[0], par_pos);  // (a CTA that owns no window still lets its tile land)
    if (pend_mat) mbar_wait(&s_bar[1], par_mat);
    FUSED_STAMP(6);
}

__global__ void __launch_bounds__(kFusedThreads, 1) fused_score_kernel(const __grid_constant__ FusedArgs a) {
    fused_body<kFusedThreads>(a);
}

// the same body within 64 registers: two CTAs per SM, for streams of independent calls
// (SSTAR_B200_FLAG_INDEPENDENT) whose consecutive launches overlap
__global__ void __launch_bounds__(kFusedThreads, 2) fused_score_pair_kernel(const __grid_constant__ FusedArgs a) {
    fused_body<kFusedThreads>(a);
}

// replicate batches: 256 threads, several CTAs per SM
__global__ void __launch_bounds__(kFusedRepThreads, 3) fused_score_rep_kernel(const __grid_constant__ FusedArgs a) {
    fused_body<kFusedRepThreads>(a);
}

}  // namespace sstar

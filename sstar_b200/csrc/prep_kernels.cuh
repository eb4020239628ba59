// sstar_b200 -- "prep" launch: genotype packing (K1) and window bounds (K0) in ONE grid.
//
// CTAs [0, n_pack) pack 32-variant groups into bit-planes; CTAs [n_pack, ...) resolve window
// bounds with a warp-cooperative 32-ary search.  The two jobs are independent, so fusing them
// takes the latency-bound search off the critical path of small problems.
#pragma once
#include "common.cuh"

namespace sstar {

// ------------------------------------------------------------------------------------------
// K0: window bounds -- sstar/generators/window_data_generator.py:113 (pos >= start) & (pos <= end)
// becomes [lo, hi) = [lower_bound(start), lower_bound(end + 1)) on the ascending position array.
// One warp per window; each round 32 lanes probe 32 pivots, shrinking the range ~33x.
// ------------------------------------------------------------------------------------------
// pos / win_start / win_end of a bounds CTA may be the device copies that fetch_kernel -- the grid this
// one is a programmatic dependent of -- was still writing when this grid started (in-place host entry):
// they are read after griddepcontrol.wait with ordinary coherent loads (ld.relaxed.gpu), never through
// the non-coherent path (__ldg / const __restrict__), which PTX only allows for data that stays
// read-only for the whole kernel.
__device__ __forceinline__ int32_t ld_coherent(const int32_t *p) {
    int32_t v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int32_t warp_lower_bound(const int32_t *pos, int32_t lo, int32_t hi,
                                                    int64_t x, int lane) {
    // invariant: pos[k] < x for k < lo, pos[k] >= x for k >= hi
    while (hi - lo > 32) {
        const int32_t step = (hi - lo + 32) / 33;
        const int32_t idx = lo + step * (lane + 1) - 1;
        const bool below = idx < hi && (int64_t)ld_coherent(pos + idx) < x;
        const int c = __popc(__ballot_sync(0xffffffffu, below));  // monotone: the first c lanes
        const int32_t new_hi = lo + step * (c + 1) - 1;  // pivot of lane c, the first one not below x
        if (c > 0) lo = lo + step * c;
        if (c < 32 && new_hi < hi) hi = new_hi;
    }
    const int32_t idx = lo + lane;
    const bool below = idx < hi && (int64_t)ld_coherent(pos + idx) < x;
    return lo + __popc(__ballot_sync(0xffffffffu, below));
}

// Fetch CTA `blk` of n_fetch: its share of pos | win_start | win_end, mapped host memory -> device.
__device__ __forceinline__ void fetch_block(const BoundsArgs &a, int blk, int nthreads, int tid) {
    const int64_t V = a.V, W = a.W, nw = V + 2 * W;
    const int64_t per = (((nw + a.n_fetch - 1) / a.n_fetch) + 3) & ~(int64_t)3;
    const int64_t i0 = (int64_t)blk * per, i1 = min(nw, i0 + per);
    int32_t *dpos = const_cast<int32_t *>(a.pos), *dws = const_cast<int32_t *>(a.win_start),
            *dwe = const_cast<int32_t *>(a.win_end);
#pragma unroll 4
    for (int64_t i = i0 + tid; i < i1; i += nthreads) {
        if (i < V) dpos[i] = a.src_pos[i];
        else if (i < V + W) dws[i - V] = a.src_ws[i - V];
        else dwe[i - V - W] = a.src_we[i - V - W];
    }
}

__device__ __forceinline__ void bounds_block(const BoundsArgs &a, int block, int nthreads, int tid) {
    const int lane = tid & 31;
    const int w = block * (nthreads >> 5) + (tid >> 5);
    if (block == 0 && tid < kMaxSlots) a.hdr->n_slow[tid] = 0;
    if (a.n_fetch) pdl_wait();  // pos / windows come from fetch_kernel, the launch before this grid
    if (w < a.W) {
        const int32_t sb = a.seg ? a.seg[w] : 0;
        const int32_t se = a.seg ? a.seg[w + 1] : a.V;
        const int64_t s = a.win_start ? ld_coherent(a.win_start + w) : a.fixed_start;
        const int64_t e = a.win_end ? ld_coherent(a.win_end + w) : a.fixed_end;
        const int32_t lo = warp_lower_bound(a.pos, sb, se, s, lane);
        int32_t hi = warp_lower_bound(a.pos, lo, se, e + 1, lane);
        if (hi < lo) hi = lo;
        if (lane < 2 && hi > lo) {  // first / last position, so the score kernel need not chase them
            const int32_t v = ld_coherent(a.pos + (lane == 0 ? lo : hi - 1));
            (lane == 0 ? a.win_pfirst : a.win_plast)[w] = v;
        }
        if (lane == 0) {
            a.win_lo[w] = lo;
            a.win_hi[w] = hi;
            a.win_bail[w] = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K1: pack.  One CTA per `gpc` groups of 32 variants.
//
// Staging: the CTA's rows of a row-major int8 matrix are ONE contiguous byte span whatever the
// row pitch is, so it moves global -> shared as aligned 128-bit streaming loads (the source's
// 16-byte phase is kept; head/tail bytes go separately).
// Compute (SIMD within a 32-bit word, 4 genotype bytes at a time):
//   A  ref rows   : signed row sum by dp4a, eight lanes per row, four rows per warp step
//                                                                       -> absent flag  (sstar.py:87-88)
//   B  tgt columns: a thread owns 4 adjacent columns x 32 rows; per row one (unaligned-capable)
//                   word read, "byte != 0" and "byte == 2" folded to bit 0 of each byte and
//                   shifted into per-8-row accumulators; PRMT transposes them into 4 column
//                   words per plane; masked by the group's absent word   (sstar.py:163-164)
//   Genotypes outside {0,1,2} are detected by two running ORs; the warp redoes such a 32x4 block with
//   ballots (a lane per row: exact planes on |b|, plus the sign plane of a flagged group) and raises the
//   group's level (1: the other bytes all are -1 / -2, 2: anything else).
// Output planes are word-major [G][Tp]: 128-bit coalesced stores here, coalesced loads in K2.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_v4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// ---- bulk asynchronous copy (TMA engine, SASS: UBLKCP) + mbarrier completion ----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// tx bytes announced by the thread that issues the copy, the (single) arrival made separately
__device__ __forceinline__ void mbar_add_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

// Fallback staging for spans the bulk engine cannot take (base not 16-byte aligned, ragged tail
// tile): 128-bit streaming loads keeping the source's 16-byte phase.  Returns that phase, i.e.
// the offset from sdst at which byte 0 of the span landed.
__device__ __forceinline__ uint32_t stage_span(uint8_t *sdst, const int8_t *gsrc, uint32_t nbytes, int tid,
                                               int nthreads) {
    const uint32_t phase = (uint32_t)((uintptr_t)gsrc & 15);
    uint8_t *s0 = sdst + phase;
    uint32_t head = (16 - phase) & 15;
    if (head > nbytes) head = nbytes;
    if ((uint32_t)tid < head) s0[tid] = (uint8_t)gsrc[tid];
    const uint32_t nvec = (nbytes - head) >> 4;
    const uint4 *gv = reinterpret_cast<const uint4 *>(gsrc + head);
    uint4 *sv = reinterpret_cast<uint4 *>(s0 + head);
    for (uint32_t i = tid; i < nvec; i += nthreads) sv[i] = ldg_stream_v4(gv + i);
    const uint32_t done = head + (nvec << 4);
    if (done + tid < nbytes) s0[done + tid] = (uint8_t)gsrc[done + tid];  // tail < 16 bytes
    return phase;
}

// 32-bit read at an arbitrary byte offset of a 4-byte-aligned shared region.
template <bool kAligned>
__device__ __forceinline__ uint32_t lds_word(const uint8_t *region, uint32_t off) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(region + (off & ~3u));
    if (kAligned) return w[0];
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

// SIMD-within-word byte tests, exact for any byte value: bit 7 of each byte of the result
__device__ __forceinline__ uint32_t bytes_nonzero(uint32_t w) {
    return (((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u;
}
__device__ __forceinline__ uint32_t bytes_zero(uint32_t w) { return ~(((w & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w) & 0x80808080u; }

// 4x4 byte transpose: a0..a3 hold (rows 0-7, 8-15, 16-23, 24-31) x (4 columns, one per byte);
// c[k] becomes the 32-row word of column k.
__device__ __forceinline__ void transpose4(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t c[4]) {
    const uint32_t t0 = __byte_perm(a0, a1, 0x5140), t1 = __byte_perm(a2, a3, 0x5140);
    const uint32_t t2 = __byte_perm(a0, a1, 0x7362), t3 = __byte_perm(a2, a3, 0x7362);
    c[0] = __byte_perm(t0, t1, 0x5410);
    c[1] = __byte_perm(t0, t1, 0x7632);
    c[2] = __byte_perm(t2, t3, 0x5410);
    c[3] = __byte_perm(t2, t3, 0x7632);
}

// A 32-row x 4-column block holding a byte outside {0,1,2} (missing calls give -1 / -2: sstar/utils/utils.py:305-306
// sums alleles of -1; multi-allelic dosages; garbage in the padding of a ragged last quad) is redone by the WHOLE
// warp, a lane per row: each lane reads its row's four bytes and sixteen ballots give the four columns' words of
// "b != 0", "|b| == 2", "b < 0" and "|b| > 2" -- exact for any byte, ~70 warp instructions per block, where a
// lane redoing its own block alone kept 31 lanes waiting for 700 (a handful of missing calls per tile, the
// common case of real data, doubled the tile's pack time).  Only when more than kDenseOddLanes blocks of a
// warp step are odd do the lanes redo their own blocks in parallel (SIMD within a word on |b|), leaving the
// ballots the blocks that hold |b| > 2.  The block's level: 2 when a reference-absent row
// of a real column holds |b| > 2, else 1 when one holds a negative byte, else 0.  The tile remembers which
// blocks were redone (their sign words are written) in kRedoWords * 32 bits.
constexpr int kRedoWords = 64;
constexpr int kDenseOddLanes = 8;  // more odd blocks than this in a warp step: per-lane redo (700 instructions flat) beats the ballots (70 a block)

// The per-lane redo of a block on |b| and the sign (dense case above), out of line: the clean path must not
// pay its registers.  Returns true when some byte of the block has |b| > 2.
// Rows are addressed as in pack_tile (kRowPhase: each row segment keeps its own 16-byte phase).
template <bool kAligned, bool kRowPhase>
__device__ __forceinline__ bool pack_block_signed(const uint8_t *s_tgt, uint32_t base, uint32_t tpitch, uint32_t tgt_phase,
                                               uint32_t pstep, uint32_t row0, uint32_t col_off, uint32_t car[4],
                                               uint32_t two[4], uint32_t neg[4]) {
    // two sweeps over the block's rows (the planes on |b|, then the sign words): fewer values live at a time
    uint32_t e_or = 0, e_and = 0;
    {
        uint32_t nz[4], tw[4];
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
            uint32_t an = 0, at = 0;
#pragma unroll
            for (int rr = 7; rr >= 0; --rr) {
                const uint32_t r = (uint32_t)(blk * 8 + rr);
                const uint32_t off = kRowPhase ? (row0 + r) * tpitch + ((tgt_phase + (row0 + r) * pstep) & 15u) + col_off
                                               : base + r * tpitch;
                const uint32_t w = lds_word<kAligned>(s_tgt, off);
                const uint32_t n01 = (w >> 7) & 0x01010101u;
                const uint32_t m = (w ^ ((n01 << 8) - n01)) + n01;  // |b| per byte
                const uint32_t h = m >> 1;
                an = an + an + ((m | h) & 0x01010101u);  // |b| in {1,2,3} -> bit 0
                at = at + at + (h & 0x01010101u);        // |b| in {2,3}   -> bit 0
                e_or |= m;
                e_and |= m & h;
            }
            nz[blk] = an;
            tw[blk] = at;
        }
        transpose4(nz[0], nz[1], nz[2], nz[3], car);
        transpose4(tw[0], tw[1], tw[2], tw[3], two);
    }
    {
        uint32_t ng[4];
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
            uint32_t ag = 0;
#pragma unroll
            for (int rr = 7; rr >= 0; --rr) {
                const uint32_t r = (uint32_t)(blk * 8 + rr);
                const uint32_t off = kRowPhase ? (row0 + r) * tpitch + ((tgt_phase + (row0 + r) * pstep) & 15u) + col_off
                                               : base + r * tpitch;
                ag = ag + ag + ((lds_word<kAligned>(s_tgt, off) >> 7) & 0x01010101u);
            }
            ng[blk] = ag;
        }
        transpose4(ng[0], ng[1], ng[2], ng[3], neg);
    }
    return ((e_or & 0xfcfcfcfcu) | (e_and & 0x01010101u)) != 0;
}

// The staged target tile holds columns [c0, c0 + ncols) of the CTA's rows with a row pitch of
// `tpitch` bytes (whole rows: c0 = 0, ncols = tpitch = T; wide panels: one column chunk).
// kRowPhase: every row segment was staged keeping its own 16-byte phase (column chunks of a matrix
// whose row length is not a multiple of 16): row r starts at r * tpitch + ((tgt_phase + r * T) & 15).
// kChunked = false (whole rows) takes pitch, origin and width from T itself.
// kLaneRows: a handful of reference columns (R <= 16, replicate batches): one LANE sums a row.
template <bool kAligned, bool kRowPhase = false, bool kChunked = false, bool kLaneRows = false>
__device__ __forceinline__ void pack_tile(const PackArgs &a, uint8_t *smem, int32_t g0, int ng, int nrows,
                                          uint32_t ref_phase, uint32_t tgt_phase, int32_t tpitch_ = 0, int32_t c0_ = 0,
                                          int32_t ncols_ = 0) {
    const int32_t tpitch = kChunked ? tpitch_ : a.T, c0 = kChunked ? c0_ : 0, ncols = kChunked ? ncols_ : a.T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kPackThreads / 32;
    uint8_t *s_flag = smem;                                                  // [gpc*32] absent per row
    uint32_t *s_word = reinterpret_cast<uint32_t *>(smem + a.gpc * 32);      // [gpc] absent word
    uint32_t *s_exo = s_word + a.gpc;                                        // [gpc] exotic flag
    const uint8_t *s_ref = smem + a.gpc * 32 + align16(a.gpc * 8);
    const uint8_t *s_tgt = s_ref + a.ref_tile_bytes;
    __shared__ uint32_t s_redo[kRedoWords + 1];  // (last word: some group of the tile is flagged) bit per 32x4 block of the tile: redone out of line (its sign words are written)

    // A: signed row sums of the reference panel (dp4a over the widest aligned vector).  A warp takes
    // four rows at a time, eight lanes per row: four independent chains of loads per warp and a
    // three-step shuffle sum instead of one row, one REDUX.
    const uint32_t ralign = (ref_phase | (uint32_t)a.R) & 15u;
    if (kLaneRows) {
        const int nwr = (a.R + 3) >> 2;
        const uint32_t last_mask = (a.R & 3) ? ((1u << (8 * (a.R & 3))) - 1u) : 0xffffffffu;
        for (int r = tid; r < ng * 32; r += kPackThreads) {
            int sum = 0;
            if (r < nrows) {
                const uint32_t row = ref_phase + (uint32_t)r * (uint32_t)a.R;
                for (int i = 0; i < nwr; ++i) {
                    uint32_t w = lds_word<false>(s_ref, row + 4u * i);
                    if (i == nwr - 1) w &= last_mask;
                    sum = __dp4a((int)w, 0x01010101, sum);
                }
            }
            s_flag[r] = (r < nrows) && (sum == 0);
        }
    } else {
    constexpr int kLpr = 8, kRpw = 32 / kLpr;  // lanes per row (4 and 16 measure the same or worse), rows per warp step
    const int sub = lane / kLpr, l8 = lane % kLpr;
    for (int r4 = warp * kRpw; r4 < ng * 32; r4 += nwarps * kRpw) {
        const int r = r4 + sub;
        int sum = 0;
        if (r < nrows) {
            const uint32_t row = ref_phase + (uint32_t)r * (uint32_t)a.R;
            if (ralign == 0) {
                const uint4 *vp = reinterpret_cast<const uint4 *>(s_ref + row);
                const int nv = a.R >> 4;
#pragma unroll 2
                for (int i = l8; i < nv; i += kLpr) {
                    const uint4 v = vp[i];
                    sum = __dp4a((int)v.x, 0x01010101, sum);
                    sum = __dp4a((int)v.y, 0x01010101, sum);
                    sum = __dp4a((int)v.z, 0x01010101, sum);
                    sum = __dp4a((int)v.w, 0x01010101, sum);
                }
            } else if ((ralign & 7u) == 0) {
                const uint2 *vp = reinterpret_cast<const uint2 *>(s_ref + row);
                const int nv = a.R >> 3;
#pragma unroll 4
                for (int i = l8; i < nv; i += kLpr) {
                    const uint2 v = vp[i];
                    sum = __dp4a((int)v.x, 0x01010101, sum);
                    sum = __dp4a((int)v.y, 0x01010101, sum);
                }
            } else if ((ralign & 3u) == 0) {
                const uint32_t *vp = reinterpret_cast<const uint32_t *>(s_ref + row);
                const int nv = a.R >> 2;
#pragma unroll 4
                for (int i = l8; i < nv; i += kLpr) sum = __dp4a((int)vp[i], 0x01010101, sum);
            } else {
                const int nwr = (a.R + 3) >> 2;
                const uint32_t last_mask = (a.R & 3) ? ((1u << (8 * (a.R & 3))) - 1u) : 0xffffffffu;
                for (int i = l8; i < nwr; i += kLpr) {
                    uint32_t w = lds_word<false>(s_ref, row + 4u * i);
                    if (i == nwr - 1) w &= last_mask;
                    sum = __dp4a((int)w, 0x01010101, sum);
                }
            }
        }
#pragma unroll
        for (int d = 1; d < kLpr; d <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
        if (l8 == 0) s_flag[r] = (r < nrows) && (sum == 0);
    }
    }
    __syncthreads();
    for (int gl = warp; gl < ng; gl += nwarps) {
        const uint32_t bits = __ballot_sync(0xffffffffu, s_flag[gl * 32 + lane] != 0);
        if (lane == 0) {
            s_word[gl] = bits;
            s_exo[gl] = 0;
            if (c0 == 0) a.absent_bits[g0 + gl] = bits;  // (every column chunk of a group computes the same word)
        }
    }
    if (tid <= kRedoWords) s_redo[tid] = 0u;
    __syncthreads();

    // B: 4 columns x 32 rows per thread (a warp's 32 consecutive items per step, the trip count warp-uniform)
    const int nq = (ncols + 3) >> 2;
    const int nitems = ng * nq;
    const uint32_t pstep = (uint32_t)a.T & 15u;
    for (int it0 = tid - lane; it0 < nitems; it0 += kPackThreads) {
        const bool valid = it0 + lane < nitems;
        const int it = valid ? it0 + lane : nitems - 1;  // (idle lanes of the last step shadow the last item, without stores)
        const int gl = it / nq, q = it - gl * nq;
        const uint32_t base = tgt_phase + (uint32_t)(gl * 32) * (uint32_t)tpitch + 4u * q;
        auto row_off = [&](int r) -> uint32_t {  // r: row within the group
            if (!kRowPhase) return base + (uint32_t)r * (uint32_t)tpitch;
            const uint32_t rr = (uint32_t)(gl * 32 + r);
            return rr * (uint32_t)tpitch + ((tgt_phase + rr * pstep) & 15u) + 4u * q;
        };
        uint32_t nz[4], tw[4], e_or = 0, e_and = 0;
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
            uint32_t an = 0, at = 0;
#pragma unroll
            for (int rr = 7; rr >= 0; --rr) {  // row 7 first: it ends up in bit 7 of each byte
                const uint32_t w = lds_word<kAligned>(s_tgt, row_off(blk * 8 + rr));
                const uint32_t h = w >> 1;
                an = an + an + ((w | h) & 0x01010101u);  // byte in {1,2,3} -> bit 0
                at = at + at + (h & 0x01010101u);        // byte in {2,3}   -> bit 0
                e_or |= w;
                e_and |= w & h;
            }
            nz[blk] = an;
            tw[blk] = at;
        }
        const uint32_t ab = s_word[gl];
        uint32_t car[4], two[4];
        transpose4(nz[0], nz[1], nz[2], nz[3], car);
        transpose4(tw[0], tw[1], tw[2], tw[3], two);
        const int ncol = min(4, ncols - 4 * q);
        const size_t o = (size_t)(g0 + gl) * a.Tp + c0 + 4 * q;
        // a byte outside {0,1,2} somewhere in this lane's block: the warp redoes such blocks one by one
        const bool odd = valid && (((e_or & 0xfcfcfcfcu) | (e_and & 0x01010101u)) != 0);
        uint32_t oddmask = __ballot_sync(0xffffffffu, odd);
        if (oddmask) {
            uint32_t neg[4] = {0u, 0u, 0u, 0u}, big = 0;
            if (__popc(oddmask) > kDenseOddLanes) {
                // many of the warp's blocks are odd (missing calls at a percent of all calls): every such lane redoes
                // its own block on |b| and the sign with the bit tests of the clean path -- with n = b >> 7 (0 / 1 per
                // byte) and s = n * 0xff, |b| = (b ^ s) + n never carries across bytes (|b| <= 0x80) -- which is exact
                // while every |b| <= 2; the blocks holding a larger one are left to the ballots below
                bool left = false;
                if (odd) left = pack_block_signed<kAligned, kRowPhase>(s_tgt, base, (uint32_t)tpitch, tgt_phase, pstep,
                                                                       (uint32_t)(gl * 32), 4u * q, car, two, neg);
                oddmask = __ballot_sync(0xffffffffu, left);
            }
            while (oddmask) {
                const int src = __ffs(oddmask) - 1;
                oddmask &= oddmask - 1;
                const int its = it0 + src, gls = its / nq, qs = its - gls * nq;  // lane src's block; this lane reads row `lane` of it
                const uint32_t rr = (uint32_t)(gls * 32 + lane);
                const uint32_t off = kRowPhase ? rr * (uint32_t)tpitch + ((tgt_phase + rr * pstep) & 15u) + 4u * qs
                                               : tgt_phase + rr * (uint32_t)tpitch + 4u * qs;
                const uint32_t w = lds_word<kAligned>(s_tgt, off);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int b = (int)(w << (24 - 8 * c)) >> 24;  // the row's byte of column c, sign-extended
                    const int mag = b < 0 ? -b : b;
                    const uint32_t m_nz = __ballot_sync(0xffffffffu, b != 0);
                    const uint32_t m_tw = __ballot_sync(0xffffffffu, mag == 2);
                    const uint32_t m_ng = __ballot_sync(0xffffffffu, b < 0);
                    const uint32_t m_bg = __ballot_sync(0xffffffffu, mag > 2);
                    if (lane == src) {
                        car[c] = m_nz;
                        two[c] = m_tw;
                        neg[c] = m_ng;
                        if (c < ncol) big |= m_bg;  // (rows past the matrix's end have no absent bit)
                    }
                }
            }
            if (odd) {
                uint32_t level = 0;
                if (big & ab) level = 2u;
                else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < ncol && (neg[c] & ab)) level = 1u;
                }
                // column chunks of a group cannot see one another's blocks, and the tile remembers kRedoWords * 32
                // redone blocks: such groups read the genotype bytes (level 2), never the sign plane
                if (level && (kChunked || it >= kRedoWords * 32)) level = 2u;
                if (it < kRedoWords * 32) atomicOr(&s_redo[it >> 5], 1u << (it & 31));
                if (level) {
                    s_redo[kRedoWords] = 1u;  // some group of the tile is flagged
                    atomicOr(&s_exo[gl], level);
                }
                if (ncol == 4) {
                    *reinterpret_cast<uint4 *>(a.neg + o) = make_uint4(neg[0] & ab, neg[1] & ab, neg[2] & ab, neg[3] & ab);
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        if (c < ncol) a.neg[o + c] = neg[c] & ab;
                }
            }
        }
        if (!valid) continue;
        if (ncol == 4) {
            *reinterpret_cast<uint4 *>(a.carried + o) = make_uint4(car[0] & ab, car[1] & ab, car[2] & ab, car[3] & ab);
            *reinterpret_cast<uint4 *>(a.two + o) = make_uint4(two[0] & ab, two[1] & ab, two[2] & ab, two[3] & ab);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (c >= ncol) continue;
                a.carried[o + c] = car[c] & ab;
                a.two[o + c] = two[c] & ab;
            }
        }
    }
    __syncthreads();
    // A flagged group's sign plane is read for ALL its columns: the blocks that took the clean path store
    // zero words now (clean groups -- all of them on clean data -- never write or read the plane).
    if (!kChunked && s_redo[kRedoWords]) {
        for (int it = tid; it < nitems; it += kPackThreads) {
            const int gl = it / nq, q = it - gl * nq;
            if (s_exo[gl] == 0 || it >= kRedoWords * 32 || ((s_redo[it >> 5] >> (it & 31)) & 1u)) continue;
            const int ncol = min(4, ncols - 4 * q);
            const size_t o = (size_t)(g0 + gl) * a.Tp + 4 * q;
            if (ncol == 4) {
                *reinterpret_cast<uint4 *>(a.neg + o) = make_uint4(0u, 0u, 0u, 0u);
            } else {
                for (int c = 0; c < ncol; ++c) a.neg[o + c] = 0u;
            }
        }
    }
    if (tid < ng) {
        if (kChunked) { if (s_exo[tid]) atomicOr(&a.exotic[g0 + tid], s_exo[tid]); }  // zeroed by the host; chunks share the flag
        else a.exotic[g0 + tid] = s_exo[tid];
    }
}

// Fallback for rows too wide to stage (32 * (R + T) bytes beyond shared memory): byte-wise from
// global memory, same outputs.
__device__ __forceinline__ void pack_tile_direct(const PackArgs &a, uint8_t *smem, int32_t g0, int ng, int nrows) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kPackThreads / 32;
    uint8_t *s_flag = smem;
    uint32_t *s_word = reinterpret_cast<uint32_t *>(smem + a.gpc * 32);
    uint32_t *s_exo = s_word + a.gpc;
    const int8_t *rbase = a.ref + (size_t)g0 * 32 * a.R;
    const int8_t *tbase = a.tgt + (size_t)g0 * 32 * a.T;
    for (int r = warp; r < ng * 32; r += nwarps) {
        int sum = 0;
        if (r < nrows) {
            const int8_t *row = rbase + (size_t)r * a.R;
            for (int c = lane; c < a.R; c += 32) sum += row[c];
        }
        sum = __reduce_add_sync(0xffffffffu, sum);
        if (lane == 0) s_flag[r] = (r < nrows) && (sum == 0);
    }
    __syncthreads();
    for (int gl = warp; gl < ng; gl += nwarps) {
        const uint32_t bits = __ballot_sync(0xffffffffu, s_flag[gl * 32 + lane] != 0);
        if (lane == 0) {
            s_word[gl] = bits;
            s_exo[gl] = 0;
            a.absent_bits[g0 + gl] = bits;
        }
    }
    __syncthreads();
    const int nitems = ng * a.T;
    for (int it = tid; it < nitems; it += kPackThreads) {
        const int gl = it / a.T, t = it - gl * a.T;
        const int rmax = min(32, nrows - gl * 32);
        const int8_t *col = tbase + (size_t)(gl * 32) * a.T + t;
        const uint32_t ab = s_word[gl];
        uint32_t car = 0, two = 0, ex = 0;
        for (int r = 0; r < rmax; ++r) {
            const int b = col[(size_t)r * a.T];
            const uint32_t isab = (ab >> r) & 1u;
            car |= ((uint32_t)(b != 0) & isab) << r;
            two |= ((uint32_t)(b == 2) & isab) << r;
            ex |= (uint32_t)((unsigned)b > 2u) & isab;
        }
        const size_t o = (size_t)(g0 + gl) * a.Tp + t;
        a.carried[o] = car;
        a.two[o] = two;
        if (ex) atomicOr(&s_exo[gl], 2u);  // (level 2: the genotype bytes are read, never the sign plane)
    }
    __syncthreads();
    if (tid < ng) a.exotic[g0 + tid] = s_exo[tid];
}

// mode: 0 = staged word path (whole rows), 1 = direct byte path, 2 = staged word path, one column chunk
// of the target rows per tile (wide panels)
// (kWide compiles the column-chunk mode in: it lives in a kernel of its own so that the common
// kernel keeps its registers)
template <bool kWide, bool kLaneRows = false>
__device__ __forceinline__ void prep_body(const PackArgs &p, const BoundsArgs &b, int n_pack, int mode, uint8_t *smem,
                                          uint64_t *s_bar) {
    pdl_launch_dependents();  // let the next kernel's CTAs get scheduled while this grid drains
    const int blk = (int)blockIdx.x;
    if (blk >= n_pack) {
        bounds_block(b, blk - n_pack, kPackThreads, threadIdx.x);
        return;
    }
    const int tid = threadIdx.x;
    // A pack CTA takes tiles blockIdx.x, blockIdx.x + n_pack, ...: one tile each in the usual grid;
    // a grid cut short (p.n_tiles > n_pack) walks the matrix front to back with a bounded number of
    // tiles in flight -- what a matrix read in place over PCIe wants.
    uint32_t parity = 0;
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        fence_mbar_init();
    }
    for (int tile = blk; tile < p.n_tiles; tile += n_pack) {
        const int cchunk = (kWide && mode == 2) ? tile % p.cchunks : 0;
        const int32_t g0 = p.g_begin + ((kWide && mode == 2) ? tile / p.cchunks : tile) * p.gpc;
        const int ng = min(p.gpc, p.g_end - g0);
        const int32_t k0 = g0 * 32;
        const int nrows = min(ng * 32, p.V - k0);
        if (mode == 1) {
            pack_tile_direct(p, smem, g0, ng, nrows);
            __syncthreads();
            continue;
        }
        uint8_t *s_ref = smem + p.gpc * 32 + align16(p.gpc * 8);
        uint8_t *s_tgt = s_ref + p.ref_tile_bytes;
        const int8_t *gref = p.ref + (size_t)k0 * p.R;
        const uint32_t ref_bytes = (uint32_t)nrows * (uint32_t)p.R;
        if (kWide && mode == 2) {
            // Wide panel: the group's reference rows (one span) and ONE column chunk of its target
            // rows (a 16-byte aligned segment per row: the host takes this mode only when T and the
            // chunk width are multiples of 16) go to the TMA engine; the chunks of a group each
            // redo the reference row sums rather than wait for one another.
            const int32_t c0 = cchunk * p.ccols, ncols = min(p.ccols, p.T - c0);
            const bool rbulk = (((uintptr_t)gref & 15) == 0) && ((ref_bytes & 15u) == 0);
            const int8_t *gt0 = p.tgt + (size_t)k0 * p.T + c0;  // first row's segment
            const uint32_t tph0 = (uint32_t)((uintptr_t)gt0 & 15);
            const bool seg16 = ((tph0 | (uint32_t)p.T) & 15u) == 0;  // every segment 16-byte aligned (chunks start at multiples of 128)
            uint32_t rph = 0;
            if (tid == 0 && rbulk) {
                mbar_add_tx(&s_bar[0], ref_bytes);
                bulk_g2s(s_ref, gref, ref_bytes, &s_bar[0]);
            }
            int fallback_row = -1;
            if (tid < nrows) {
                // the 16-byte aligned superset of the row's segment: the row keeps its phase in the tile
                const int8_t *src = gt0 + (size_t)tid * p.T;
                const uint32_t ph = (uint32_t)((uintptr_t)src & 15);
                const uint32_t bytes = (ncols + ph + 15u) & ~15u;
                if (src - ph + bytes <= p.tgt + (size_t)p.V * p.T) {
                    mbar_add_tx(&s_bar[0], bytes);
                    bulk_g2s(s_tgt + (uint32_t)tid * (uint32_t)p.tpitch, src - ph, bytes, &s_bar[0]);
                } else {
                    fallback_row = tid;  // the superset would run past the end of the matrix (its last row)
                }
            }
            if (!rbulk) rph = stage_span(s_ref, gref, ref_bytes, tid, kPackThreads);
            {   // at most the matrix's last row: one warp stages it with vector loads
                const uint32_t fm = __ballot_sync(0xffffffffu, fallback_row >= 0);
                if (tid < 32 && fm) {
                    const int r = __ffs(fm) - 1;
                    stage_span(s_tgt + (uint32_t)r * (uint32_t)p.tpitch, gt0 + (size_t)r * p.T, (uint32_t)ncols, tid, 32);
                }
            }
            __syncthreads();  // every copy is announced before the barrier's one arrival
            if (tid == 0) mbar_arrive(&s_bar[0]);
            mbar_wait(&s_bar[0], parity);
            parity ^= 1u;
            if (seg16) pack_tile<true, false, true>(p, smem, g0, ng, nrows, rph, 0u, p.tpitch, c0, ncols);
            else if (((tph0 | (uint32_t)p.T) & 3u) == 0) pack_tile<true, true, true>(p, smem, g0, ng, nrows, rph, tph0, p.tpitch, c0, ncols);
            else pack_tile<false, true, true>(p, smem, g0, ng, nrows, rph, tph0, p.tpitch, c0, ncols);
            __syncthreads();
            continue;
        }
        const int8_t *gtgt = p.tgt + (size_t)k0 * p.T;
        const uint32_t tgt_bytes = (uint32_t)nrows * (uint32_t)p.T;
        uint32_t rph = 0, tph = 0;
        const bool bulk = ((((uintptr_t)gref | (uintptr_t)gtgt) & 15) == 0) && (((ref_bytes | tgt_bytes) & 15u) == 0);
        if (bulk) {
            // one thread hands both spans to the TMA engine; no registers, no issue slots for the copy
            if (tid == 0) {
                mbar_expect_tx(&s_bar[0], ref_bytes + tgt_bytes);
                bulk_g2s(s_ref, gref, ref_bytes, &s_bar[0]);
                bulk_g2s(s_tgt, gtgt, tgt_bytes, &s_bar[0]);
            }
            __syncthreads();  // barrier initialised before anyone polls it
            mbar_wait(&s_bar[0], parity);
            parity ^= 1u;
        } else {
            rph = stage_span(s_ref, gref, ref_bytes, tid, kPackThreads);
            tph = stage_span(s_tgt, gtgt, tgt_bytes, tid, kPackThreads);
            __syncthreads();
        }
        const bool aligned = ((tph | (uint32_t)p.T) & 3u) == 0;
        if (aligned) pack_tile<true, false, false, kLaneRows>(p, smem, g0, ng, nrows, rph, tph);
        else         pack_tile<false, false, false, kLaneRows>(p, smem, g0, ng, nrows, rph, tph);
        __syncthreads();  // the tile and its flags are done with before the next one lands
    }
}

// in-place host entry: pos | win_start | win_end from mapped host memory into the scratch; the prep
// grid behind it starts right away (its pack CTAs need none of this)
__global__ void __launch_bounds__(kPackThreads) fetch_kernel(const __grid_constant__ BoundsArgs b) {
    pdl_launch_dependents();
    fetch_block(b, (int)blockIdx.x, kPackThreads, threadIdx.x);
}

__global__ void __launch_bounds__(kPackThreads)
prep_kernel(const __grid_constant__ PackArgs p, const __grid_constant__ BoundsArgs b, int n_pack, int mode) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    prep_body<false>(p, b, n_pack, mode, smem, s_bar);
}

__global__ void __launch_bounds__(kPackThreads)
prep_wide_kernel(const __grid_constant__ PackArgs p, const __grid_constant__ BoundsArgs b, int n_pack, int mode) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    prep_body<true>(p, b, n_pack, mode, smem, s_bar);
}

// reference panels of at most 16 columns (whole-row tiles only)
__global__ void __launch_bounds__(kPackThreads)
prep_narrow_kernel(const __grid_constant__ PackArgs p, const __grid_constant__ BoundsArgs b, int n_pack, int mode) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    prep_body<false, true>(p, b, n_pack, mode, smem, s_bar);
}

}  // namespace sstar

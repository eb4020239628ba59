// sstar_b200 -- ingest filters on the device.
//
// Input: the raw calls of one chromosome as the VCF scanner (or scikit-allel) gives them,
// GT int8 [V, N, 2] (-1 = missing allele) for all N samples in header order, plus the sample
// indices of the reference and target populations.  Output: exactly what read_data hands to
// the tiler (sstar/utils/utils.py:225-308), compacted on the device:
//   * optional ancestral-allele polarisation, decided per variant on the host
//     (mode 0 keep, 1 flip g -> |g - 1|, 2 drop)                                  utils.py:363-423
//   * rows where EVERY sample of a population has a missing call are dropped (tested on the raw
//     calls, per population; keeping the positions both populations share = dropping a row that
//     either population drops)                                                    utils.py:105-107,203-214
//   * rows hom-alt (all alleles equal and > 0) in every reference AND every target sample are
//     dropped                                                                     utils.py:282-291
//   * phased: [V, N, 2] -> [V', N*2]; unphased: allele sum -> [V', N]              utils.py:293-306
//
// Three launches: keep flags + counts per block of 256 variants (one warp per variant) ->
// exclusive scan of the block counts -> compacting write.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sstar {

constexpr int kIngestWarps = 8;
constexpr int kIngestRowsPerBlock = 256;

struct IngestArgs {
    const int8_t *gt;     // [V, N, 2]
    const int32_t *pos;   // [V]
    const int8_t *mode;   // [V] or null
    const int32_t *ref_cols, *tgt_cols;
    int64_t V;
    int32_t N, n_ref, n_tgt, phased;
    int32_t ref_listed, tgt_listed;  // len(ref_samples) / len(tgt_samples) of the ind files: what the fixed-site filter compares with
    uint8_t *keep;        // [V]  bit 0: row survives, bit 1: row survives the missing / ancestral filters
    int32_t *block_base;  // [nb] counts, then offsets
    int8_t *ref_out, *tgt_out;
    int32_t *pos_out;
    int64_t *counts;      // [0] rows kept, [1] rows shared by both populations before the fixed filter
};

__device__ __forceinline__ int ingest_flip(int a, bool flip) {
    if (!flip) return a;
    const int d = a - 1;
    return d < 0 ? -d : d;
}

// (#samples with a missing raw call, #samples hom-alt after polarisation) of one population
__device__ __forceinline__ void ingest_pop_counts(const int8_t *row, const int32_t *cols, int n, bool flip, int lane,
                                                  int &n_missing, int &n_homalt) {
    int miss = 0, hom = 0;
    for (int i = lane; i < n; i += 32) {
        const char2 g = *reinterpret_cast<const char2 *>(row + 2 * (size_t)cols[i]);
        miss += (g.x < 0 || g.y < 0);
        const int a0 = ingest_flip(g.x, flip), a1 = ingest_flip(g.y, flip);
        hom += (a0 > 0 && a1 == a0);
    }
    n_missing = __reduce_add_sync(0xffffffffu, miss);
    n_homalt = __reduce_add_sync(0xffffffffu, hom);
}

__global__ void __launch_bounds__(kIngestWarps * 32) ingest_flag_kernel(const __grid_constant__ IngestArgs a) {
    __shared__ int s_kept, s_shared;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_kept = 0; s_shared = 0; }
    __syncthreads();
    const int64_t k0 = (int64_t)blockIdx.x * kIngestRowsPerBlock;
    int kept = 0, shared = 0;
    for (int r = warp; r < kIngestRowsPerBlock; r += kIngestWarps) {
        const int64_t k = k0 + r;
        if (k >= a.V) break;
        const int mode = a.mode ? a.mode[k] : 0;
        const int8_t *row = a.gt + (size_t)k * a.N * 2;
        int mr, hr, mt, ht;
        ingest_pop_counts(row, a.ref_cols, a.n_ref, mode == 1, lane, mr, hr);
        ingest_pop_counts(row, a.tgt_cols, a.n_tgt, mode == 1, lane, mt, ht);
        const bool sh = mode != 2 && mr != a.n_ref && mt != a.n_tgt;
        const bool kp = sh && !(hr == a.ref_listed && ht == a.tgt_listed);  // utils.py:282-291 compares with len(samples)
        if (lane == 0) a.keep[k] = (uint8_t)((kp ? 1 : 0) | (sh ? 2 : 0));
        kept += kp;
        shared += sh;
    }
    if (lane == 0) { atomicAdd(&s_kept, kept); atomicAdd(&s_shared, shared); }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.block_base[blockIdx.x] = s_kept;
        if (s_shared) atomicAdd(reinterpret_cast<unsigned long long *>(a.counts + 1), (unsigned long long)s_shared);
    }
}

// exclusive scan of the per-block counts, one CTA; total -> counts[0]
__global__ void __launch_bounds__(1024) ingest_scan_kernel(const __grid_constant__ IngestArgs a, int64_t nb) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int64_t v = i < nb ? a.block_base[i] : 0;
        int64_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t t = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, t, d);
                if (lane >= d) t += y;
            }
            s_warp[lane] = t;
        }
        __syncthreads();
        const int64_t carry = s_carry;
        if (i < nb) a.block_base[i] = (int32_t)(carry + x - v + (warp ? s_warp[warp - 1] : 0));
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.counts[0] = s_carry;
}

__device__ __forceinline__ void ingest_write_pop(const int8_t *row, const int32_t *cols, int n, bool flip, bool phased,
                                                 int8_t *out, int lane) {
    for (int i = lane; i < n; i += 32) {
        const char2 g = *reinterpret_cast<const char2 *>(row + 2 * (size_t)cols[i]);
        const int a0 = ingest_flip(g.x, flip), a1 = ingest_flip(g.y, flip);
        if (phased) { out[2 * i] = (int8_t)a0; out[2 * i + 1] = (int8_t)a1; }
        else out[i] = (int8_t)(a0 + a1);
    }
}

__global__ void __launch_bounds__(kIngestWarps * 32) ingest_write_kernel(const __grid_constant__ IngestArgs a) {
    __shared__ int s_off[kIngestRowsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k0 = (int64_t)blockIdx.x * kIngestRowsPerBlock;
    // intra-block offsets of the kept rows: thread r owns row r
    {
        const int r = threadIdx.x;
        const int flag = (k0 + r < a.V) ? (a.keep[k0 + r] & 1) : 0;
        const uint32_t m = __ballot_sync(0xffffffffu, flag);
        __shared__ int s_wsum[kIngestWarps];
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s_wsum[w];
        s_off[r] = flag ? before + __popc(m & ((1u << lane) - 1u)) : -1;
    }
    __syncthreads();
    const int64_t base = a.block_base[blockIdx.x];
    const int R = a.n_ref * (a.phased ? 2 : 1), T = a.n_tgt * (a.phased ? 2 : 1);
    for (int r = warp; r < kIngestRowsPerBlock; r += kIngestWarps) {
        const int off = s_off[r];
        if (off < 0) continue;
        const int64_t k = k0 + r, j = base + off;
        const bool flip = a.mode && a.mode[k] == 1;
        const int8_t *row = a.gt + (size_t)k * a.N * 2;
        ingest_write_pop(row, a.ref_cols, a.n_ref, flip, a.phased, a.ref_out + (size_t)j * R, lane);
        ingest_write_pop(row, a.tgt_cols, a.n_tgt, flip, a.phased, a.tgt_out + (size_t)j * T, lane);
        if (lane == 0) a.pos_out[j] = a.pos[k];
    }
}

}  // namespace sstar

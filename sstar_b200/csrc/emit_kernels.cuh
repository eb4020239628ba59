// sstar_b200 -- result emitter: the feature table as TSV text, formatted on the device.
//
// Reproduces, byte for byte, what the reference gets from
//     pd.DataFrame(rows).to_csv(output_file, sep="\t", index=False, na_rep="NA")
// (sstar/preprocess.py:122, sstar/simulate.py:195) for rows built by
// FeatureVectorPreprocessor._fmt_res (sstar/feature_vector_preprocessor.py:167-181):
//     Chromosome \t Start \t End \t Sample \t S*_score \t Region_ind_SNP_number [\t Replicate] \n
// S*_score is a float column: integral values print as "51470.0", NaN as "NA"; the other columns
// are integers / strings.  Rows come in (window, column) order, columns optionally permuted.
//
// Three launches: row lengths per block of 256 rows -> exclusive scan of the block totals ->
// format.  A block formats its rows into shared memory and copies them out with coalesced byte
// stores, so the global writes are one contiguous span per block.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sstar {

constexpr int kEmitThreads = 256;  // rows per block
constexpr int64_t kEmitMaxAbsScore = 10000000000000000LL;  // 1e16: repr(float) turns to exponent form there

struct EmitArgs {
    const int64_t *score;      // [W, T], INT64_MIN = NaN
    const int32_t *nsnp;       // [W, T]
    const int32_t *win_start;  // [W]
    const int32_t *win_end;    // [W]
    const int32_t *replicate;  // [W] or null (no Replicate column)
    const int32_t *col_order;  // [T] row k of a window shows column col_order[k]; null = identity
    const char *labels;        // sample labels, concatenated
    const int32_t *label_off;  // [T + 1]
    // optional prediction columns (sstar/quantile_regression.py:108-135): entry v of the table
    // belongs to Region_ind_SNP_number == v (clamped to n_pred - 1)
    const char *pred_text;      // decimal text of the predictions, concatenated; null = no such column
    const int32_t *pred_off;    // [n_pred + 1]
    const double *pred_value;   // [n_pred]
    int32_t n_pred;
    int32_t mode;        // 0: feature rows (+ prediction column when pred_text), 1: BED rows of the units
                         //    whose score exceeds the prediction: chrom, start - 1, end, sample
    int32_t sample_major;  // 0: rows ordered (window, sample k), 1: (sample k, window)
    int32_t W, T, blocks_per_win, row_cap;  // blocks_per_win: blocks per OUTER index; row_cap: smem bytes per row
    int32_t chrom_len;
    char chrom[64];
    int64_t *block_base;  // [W * blocks_per_win]: block totals, then (after the scan) block offsets
    char *out;
    int64_t capacity;
    int64_t *total;    // bytes written (device scalar)
    uint32_t *status;  // bit 0: out too small, bit 1: a score beyond 1e16 (host must format that table)
};

__device__ __forceinline__ int ndigits_u64(uint64_t v) {
    int n = 1;
    while (v >= 10000) { v /= 10000; n += 4; }
    if (v >= 1000) return n + 3;
    if (v >= 100) return n + 2;
    if (v >= 10) return n + 1;
    return n;
}
__device__ __forceinline__ int len_i64(int64_t v) {
    return v < 0 ? 1 + ndigits_u64((uint64_t)(-(v + 1)) + 1u) : ndigits_u64((uint64_t)v);
}
// writes v at p, returns the number of bytes
__device__ __forceinline__ int put_i64(char *p, int64_t v) {
    uint64_t u;
    int n = 0;
    if (v < 0) { p[n++] = '-'; u = (uint64_t)(-(v + 1)) + 1u; } else u = (uint64_t)v;
    const int d = ndigits_u64(u);
    for (int i = d - 1; i >= 0; --i) { p[n + i] = (char)('0' + (int)(u % 10)); u /= 10; }
    return n + d;
}

struct EmitRow {
    int w, col;
    bool valid;
};
__device__ __forceinline__ EmitRow emit_row(const EmitArgs &a) {
    EmitRow r;
    const int outer = blockIdx.x / a.blocks_per_win;
    const int inner = (blockIdx.x % a.blocks_per_win) * kEmitThreads + threadIdx.x;
    int k;  // position of the sample in the printed order
    if (a.sample_major) { k = outer; r.w = inner; r.valid = inner < a.W; }
    else                { k = inner; r.w = outer; r.valid = inner < a.T; }
    if (!r.valid) r.w = 0;
    r.col = r.valid ? (a.col_order ? a.col_order[k] : k) : 0;
    return r;
}
__device__ __forceinline__ int emit_pred_index(const EmitArgs &a, const EmitRow &r) {
    const int v = a.nsnp[(size_t)r.w * a.T + r.col];
    return v < 0 ? 0 : (v >= a.n_pred ? a.n_pred - 1 : v);
}
__device__ __forceinline__ int emit_row_len(const EmitArgs &a, const EmitRow &r, int64_t s) {
    if (a.mode == 1) {  // BED row, only where the observed score exceeds the predicted one
        if (s == INT64_MIN || !((double)s > a.pred_value[emit_pred_index(a, r)])) return 0;
        return a.chrom_len + 1 + len_i64((int64_t)a.win_start[r.w] - 1) + 1 + len_i64(a.win_end[r.w]) + 1 +
               (a.label_off[r.col + 1] - a.label_off[r.col]) + 1;
    }
    int n = a.chrom_len + 1 + len_i64(a.win_start[r.w]) + 1 + len_i64(a.win_end[r.w]) + 1 +
            (a.label_off[r.col + 1] - a.label_off[r.col]) + 1;
    n += (s == INT64_MIN) ? 2 : len_i64(s) + 2;
    n += 1 + len_i64(a.nsnp[(size_t)r.w * a.T + r.col]);
    if (a.replicate) n += 1 + len_i64(a.replicate[r.w]);
    if (a.pred_text) {
        const int pi = emit_pred_index(a, r);
        n += 1 + ((s == INT64_MIN) ? 2 : a.pred_off[pi + 1] - a.pred_off[pi]);
    }
    return n + 1;
}

// block-wide exclusive scan of one int per thread (kEmitThreads threads); returns the exclusive
// prefix, the block total lands in *total
__device__ __forceinline__ int block_excl_scan(int v, int *total, int *s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int t = lane < kEmitThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, t, d);
            if (lane >= d) t += y;
        }
        if (lane < kEmitThreads / 32) s_warp[lane] = t;  // inclusive over warps
    }
    __syncthreads();
    *total = s_warp[kEmitThreads / 32 - 1];
    return x - v + (warp ? s_warp[warp - 1] : 0);
}

__global__ void __launch_bounds__(kEmitThreads) emit_len_kernel(const __grid_constant__ EmitArgs a) {
    __shared__ int s_warp[kEmitThreads / 32];
    const EmitRow r = emit_row(a);
    int n = 0;
    if (r.valid) {
        const int64_t s = a.score[(size_t)r.w * a.T + r.col];
        if (s != INT64_MIN && (s >= kEmitMaxAbsScore || s <= -kEmitMaxAbsScore)) atomicOr(a.status, 2u);
        n = emit_row_len(a, r, s);
    }
    int total;
    block_excl_scan(n, &total, s_warp);
    if (threadIdx.x == 0) a.block_base[blockIdx.x] = total;
}

// exclusive scan of the block totals, one CTA
__global__ void __launch_bounds__(1024) emit_scan_kernel(const __grid_constant__ EmitArgs a, int64_t nb) {
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
        if (i < nb) a.block_base[i] = carry + x - v + (warp ? s_warp[warp - 1] : 0);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *a.total = s_carry;
        if (s_carry > a.capacity) atomicOr(a.status, 1u);
    }
}

__global__ void __launch_bounds__(kEmitThreads) emit_format_kernel(const __grid_constant__ EmitArgs a) {
    extern __shared__ __align__(16) char s_text[];
    __shared__ int s_warp[kEmitThreads / 32];
    if (*a.status) return;  // table not representable here or buffer too small: nothing is written
    const EmitRow r = emit_row(a);
    int n = 0;
    int64_t s = 0;
    if (r.valid) {
        s = a.score[(size_t)r.w * a.T + r.col];
        n = emit_row_len(a, r, s);
    }
    int total;
    const int off = block_excl_scan(n, &total, s_warp);
    if (r.valid && n > 0) {
        char *p = s_text + off;
        for (int i = 0; i < a.chrom_len; ++i) *p++ = a.chrom[i];
        *p++ = '\t';
        p += put_i64(p, (int64_t)a.win_start[r.w] - (a.mode == 1 ? 1 : 0));
        *p++ = '\t';
        p += put_i64(p, a.win_end[r.w]);
        *p++ = '\t';
        const int l0 = a.label_off[r.col], l1 = a.label_off[r.col + 1];
        for (int i = l0; i < l1; ++i) *p++ = a.labels[i];
        if (a.mode == 0) {
            *p++ = '\t';
            if (s == INT64_MIN) { *p++ = 'N'; *p++ = 'A'; }
            else { p += put_i64(p, s); *p++ = '.'; *p++ = '0'; }
            *p++ = '\t';
            p += put_i64(p, a.nsnp[(size_t)r.w * a.T + r.col]);
            if (a.replicate) { *p++ = '\t'; p += put_i64(p, a.replicate[r.w]); }
            if (a.pred_text) {
                *p++ = '\t';
                if (s == INT64_MIN) { *p++ = 'N'; *p++ = 'A'; }
                else {
                    const int pi = emit_pred_index(a, r);
                    for (int i = a.pred_off[pi]; i < a.pred_off[pi + 1]; ++i) *p++ = a.pred_text[i];
                }
            }
        }
        *p++ = '\n';
    }
    __syncthreads();
    char *dst = a.out + a.block_base[blockIdx.x];
    for (int i = threadIdx.x; i < total; i += kEmitThreads) dst[i] = s_text[i];
}

}  // namespace sstar

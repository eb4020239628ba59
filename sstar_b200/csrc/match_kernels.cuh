// sstar_b200 -- source match rates of inferred tracts.
//
// The reference (sstar/match.py:430-565, run_match) walks tracts x source individuals in Python:
// for the VCF positions inside a tract, dosage match 1 - |x - y| / P between the target (diploid
// ALT dosage, or one haplotype's allele * P for phased labels, :120-162) and each source's diploid
// dosage, summed over the sites where both are called (:62-73), divided by the number of sites in
// the tract; the tract's rate is the mean over the sources.
//
// Here one warp takes one TRACT and returns, for every source, the INTEGER numerator
//     N = sum over callable sites of (P - |x - y|)
// so that the host forms (N / P) / count and the mean over sources exactly as the reference does
// (for P = 2 every term of the reference's float sum is a multiple of 1/2, hence exact).
// Layout: the calls of one variant are one contiguous row of 2N bytes, so the lanes of the warp
// spread over (row, source): Sp = the source count rounded up to a power of two (<= 32) sources
// side by side -- their columns of a row sit in a few 32-byte sectors -- and 32 / Sp consecutive rows;
// the tract's target column is read once per row (a broadcast within the lanes of that row) instead
// of once per (row, source) pair, and each needed sector of a row once per tract.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sstar {

struct MatchArgs {
    const int8_t *gt;           // [V, N, 2], -1 = missing allele
    int32_t N;
    const int32_t *tract_lo;    // [K] variant index range of each tract
    const int32_t *tract_hi;
    const int32_t *tract_tgt;   // [K] target sample column
    const int32_t *tract_hap;   // [K] haplotype index (0 / 1) or -1 for the diploid dosage
    const int32_t *src_cols;    // [S]
    int32_t K, S, P;
    int64_t *numer;             // [K, S]
};

// scikit-allel's to_n_alt(): number of non-reference allele calls; a genotype with any allele < 0
// is missing (-1)
__device__ __forceinline__ int match_diploid_dosage(char2 g) {
    if (g.x < 0 || g.y < 0) return -1;
    return (g.x > 0) + (g.y > 0);
}

__global__ void __launch_bounds__(256) match_kernel(const __grid_constant__ MatchArgs a) {
    const int lane = threadIdx.x & 31;
    const int k = (int)((int64_t)blockIdx.x * 8 + (threadIdx.x >> 5));
    if (k >= a.K) return;
    const int32_t lo = a.tract_lo[k], hi = a.tract_hi[k];
    const int tcol = a.tract_tgt[k], hap = a.tract_hap[k];
    int Sp = 1;
    while (Sp < a.S && Sp < 32) Sp <<= 1;
    const int rows_per = 32 / Sp, r = lane / Sp, sl = lane - r * Sp;
    const size_t pitch = (size_t)a.N * 2;
    for (int s0 = 0; s0 < a.S; s0 += 32) {  // (more than 32 sources: blocks of 32)
        const int s = s0 + sl;
        const bool live = s < a.S;
        const int scol = live ? a.src_cols[s] : 0;
        int64_t acc = 0;
#pragma unroll 4
        for (int32_t i = lo + r; i < hi; i += rows_per) {
            const int8_t *row = a.gt + (size_t)i * pitch;
            const char2 tg = *reinterpret_cast<const char2 *>(row + 2 * (size_t)tcol);
            const char2 sg = *reinterpret_cast<const char2 *>(row + 2 * (size_t)scol);
            int x;
            if (hap < 0) x = match_diploid_dosage(tg);
            else {
                const int al = hap == 0 ? tg.x : tg.y;
                x = al >= 0 ? al * a.P : -1;
            }
            const int y = match_diploid_dosage(sg);
            if (live && x >= 0 && y >= 0) {
                const int d = x > y ? x - y : y - x;
                acc += a.P - d;
            }
        }
        for (int o = Sp; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);  // the lanes of one source
        if (live && r == 0) a.numer[(size_t)k * a.S + s] = acc;
    }
}

}  // namespace sstar

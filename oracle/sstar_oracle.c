/* CPU oracle for the S* hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Scalar C restatement of the reference algorithm (xin-huang/sstar, paths relative
 * to /root/reference), literal pairwise formulation:
 *   - window slicing by closed interval      sstar/generators/window_data_generator.py:113-116
 *   - signed ref row-sum == 0 "absent" mask   sstar/sstar.py:87-92
 *   - carried sites of the focal column       sstar/sstar.py:162-167
 *   - |union(ref-polymorphic, carried)|       sstar/sstar.py:171
 *   - pair scores (+bonus / penalty / -inf)   sstar/sstar.py:195-203
 *   - O(n^2) max-plus chain, NaN rules        sstar/sstar.py:191-193,205-217
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 * Pinned against the reference's golden vectors by tests/test_oracle.py.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_NAN INT64_MIN
#define NEG_INF INT64_MIN

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* One (window, column) unit given the window's member rows. Returns the score or ORACLE_NAN.
 * With back != NULL it also records the chain (pinned only by the
 * reference's legacy fixtures tests/exp_results/test.score.*.exp.results.tsv): back[j] = i when the
 * chain ending at i is extended (first i attaining the max, extension iff best[i] >= 0),
 * -(i + 2) when the pair (i, j) starts a chain; *end_j = first j attaining the final max. */
static int64_t chain_score(const int32_t *p, const int8_t *g, int n, int64_t bonus, int64_t max_mm,
                           int64_t penalty, int64_t *best /* scratch[n] */, int32_t *back /* scratch[n] or NULL */,
                           int *end_j) {
    if (end_j) *end_j = -1;
    if (n <= 2) return ORACLE_NAN;
    for (int j = 0; j < n; ++j) {
        int64_t top = NEG_INF;
        int32_t link = INT32_MIN;
        for (int i = 0; i < j; ++i) {
            int64_t d = (int64_t)p[j] - (int64_t)p[i];
            if (d < 0) d = -d;
            int64_t q = (int64_t)g[j] - (int64_t)g[i];
            if (q < 0) q = -q;
            int64_t s;
            if (d < 10) continue;
            if (q == 0) s = d + bonus;
            else if (q <= max_mm) s = penalty;
            else continue;
            int64_t cand = s;
            int extend = best[i] != NEG_INF && best[i] >= 0;
            if (extend) cand = best[i] + s;
            if (cand > top) { top = cand; link = extend ? i : -(i + 2); }
        }
        best[j] = top;
        if (back) back[j] = link;
    }
    int64_t out = NEG_INF;
    for (int j = 0; j < n; ++j)
        if (best[j] > out) { out = best[j]; if (end_j) *end_j = j; }
    return out == NEG_INF ? ORACLE_NAN : out;
}

static int is_sorted(const int32_t *pos, int64_t V) {
    for (int64_t k = 1; k < V; ++k)
        if (pos[k] < pos[k - 1]) return 0;
    return 1;
}

static int64_t lower_bound(const int32_t *pos, int64_t V, int64_t x) { /* first k: pos[k] >= x */
    int64_t lo = 0, hi = V;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (pos[mid] < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* score[W*T] (ORACLE_NAN = NaN), nsnp[W*T]; with set_off != NULL also the chain of every unit:
 * set_off[W*T + 1], and (when set_pos != NULL) the chain positions, ascending, for the units whose
 * entries end within `capacity`. Returns 0, or -1 on allocation failure. */
static int run_windows(const int8_t *ref_gt, const int8_t *tgt_gt, const int32_t *pos,
                       int64_t V, int32_t R, int32_t T, const int32_t *win_start,
                       const int32_t *win_end, int32_t W, int32_t match_bonus,
                       int32_t max_mismatch, int32_t mismatch_penalty, int64_t *score,
                       int32_t *nsnp, int64_t *set_off, int32_t *set_pos, int64_t capacity) {
    int sorted = is_sorted(pos, V);
    int64_t *members = (int64_t *)malloc(sizeof(int64_t) * (size_t)(V > 0 ? V : 1));
    uint8_t *absent = (uint8_t *)malloc((size_t)(V > 0 ? V : 1));
    int32_t *p = (int32_t *)malloc(sizeof(int32_t) * (size_t)(V > 0 ? V : 1));
    int8_t *g = (int8_t *)malloc((size_t)(V > 0 ? V : 1));
    int32_t *uni = (int32_t *)malloc(sizeof(int32_t) * (size_t)(V > 0 ? V : 1));
    int64_t *best = (int64_t *)malloc(sizeof(int64_t) * (size_t)(V > 0 ? V : 1));
    int32_t *back = (int32_t *)malloc(sizeof(int32_t) * (size_t)(V > 0 ? V : 1));
    if (!members || !absent || !p || !g || !uni || !best || !back) return -1;
    int64_t cursor = 0;

    for (int64_t k = 0; k < V; ++k) {
        int64_t s = 0;
        for (int32_t r = 0; r < R; ++r) s += ref_gt[k * R + r];
        absent[k] = (s == 0);
    }
    for (int32_t w = 0; w < W; ++w) {
        int64_t m = 0;
        if (sorted) {
            int64_t lo = lower_bound(pos, V, win_start[w]);
            int64_t hi = lower_bound(pos, V, (int64_t)win_end[w] + 1);
            for (int64_t k = lo; k < hi; ++k) members[m++] = k;
        } else {
            for (int64_t k = 0; k < V; ++k)
                if (pos[k] >= win_start[w] && pos[k] <= win_end[w]) members[m++] = k;
        }
        for (int32_t t = 0; t < T; ++t) {
            int n = 0;
            int64_t nu = 0;
            for (int64_t a = 0; a < m; ++a) {
                int64_t k = members[a];
                if (!absent[k]) {
                    uni[nu++] = pos[k];
                } else if (tgt_gt[k * T + t] != 0) {
                    p[n] = pos[k];
                    g[n] = tgt_gt[k * T + t];
                    uni[nu++] = pos[k];
                    ++n;
                }
            }
            /* |union|: distinct positions */
            qsort(uni, (size_t)nu, sizeof(int32_t), cmp_i32);
            int32_t distinct = 0;
            for (int64_t a = 0; a < nu; ++a)
                if (a == 0 || uni[a] != uni[a - 1]) ++distinct;
            nsnp[(int64_t)w * T + t] = distinct;
            int end_j = -1;
            score[(int64_t)w * T + t] = chain_score(p, g, n, match_bonus, max_mismatch, mismatch_penalty, best,
                                                    set_off ? back : NULL, set_off ? &end_j : NULL);
            if (set_off) {
                int64_t len = 0;
                if (end_j >= 0) {
                    for (int j = end_j;;) { ++len; if (back[j] >= 0) j = back[j]; else { ++len; break; } }
                }
                set_off[(int64_t)w * T + t] = cursor;
                if (set_pos && len > 0 && cursor + len <= capacity) {
                    int64_t k = len - 1;
                    for (int j = end_j;;) {
                        set_pos[cursor + k--] = p[j];
                        if (back[j] >= 0) j = back[j];
                        else { set_pos[cursor + k] = p[-back[j] - 2]; break; }
                    }
                }
                cursor += len;
            }
        }
    }
    if (set_off) set_off[(int64_t)W * T] = cursor;
    free(members); free(absent); free(p); free(g); free(uni); free(best); free(back);
    return 0;
}

int sstar_oracle_score_windows(const int8_t *ref_gt, const int8_t *tgt_gt, const int32_t *pos,
                               int64_t V, int32_t R, int32_t T, const int32_t *win_start,
                               const int32_t *win_end, int32_t W, int32_t match_bonus,
                               int32_t max_mismatch, int32_t mismatch_penalty, int64_t *score,
                               int32_t *nsnp) {
    return run_windows(ref_gt, tgt_gt, pos, V, R, T, win_start, win_end, W, match_bonus, max_mismatch,
                       mismatch_penalty, score, nsnp, NULL, NULL, 0);
}

int sstar_oracle_snp_sets(const int8_t *ref_gt, const int8_t *tgt_gt, const int32_t *pos,
                          int64_t V, int32_t R, int32_t T, const int32_t *win_start,
                          const int32_t *win_end, int32_t W, int32_t match_bonus,
                          int32_t max_mismatch, int32_t mismatch_penalty, int64_t *score,
                          int32_t *nsnp, int64_t *set_off, int32_t *set_pos, int64_t capacity) {
    return run_windows(ref_gt, tgt_gt, pos, V, R, T, win_start, win_end, W, match_bonus, max_mismatch,
                       mismatch_penalty, score, nsnp, set_off, set_pos, capacity);
}

int sstar_oracle_version(void) { return 1; }

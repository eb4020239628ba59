// sstar_b200 -- host side of the C-ABI (include/sstar_b200.h) for the S* hot path.
//
// Pipeline per call (all integer work, HBM/ALU bound, no tensor cores -- see DESIGN.md):
//
//   prep_kernel          K1 pack: int8 genotype matrices, read once -> absent mask + 2 bit-planes
//                        K0 bounds (same grid): closed windows -> variant index ranges [lo, hi)
//   score_prefix_kernel  K2: one THREAD per (window, column): exact O(n) prefix-max chain DP
//   score_pairwise_kernel K3: one WARP per (window, column): literal O(n^2) recurrence, any genotype
//
// K2 and K3 are chained with programmatic dependent launch so their launch latency hides under
// the predecessor.  Results are bit-identical to the reference (int64 scores, INT64_MIN = NaN).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <mutex>

#include "sstar_b200.h"
#include "common.cuh"
#include "prep_kernels.cuh"
#include "score_kernels.cuh"
#include "emit_kernels.cuh"
#include "ingest_kernels.cuh"
#include "match_kernels.cuh"
#include "snpset_kernels.cuh"
#include "fused_kernels.cuh"

using namespace sstar;

namespace {

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

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// SMs of the CURRENT device (read once per device; every launch heuristic sizes its grid by it).
// kNumSM (common.cuh) is only the sizing constant of the device-independent workspace layout.
int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSM;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSM;
    cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

// Pack CTAs an SM holds when only registers and threads bound them (small tiles): what "one wave" of the pack
// grid means for small problems.  Asked of the runtime once per device -- the kernel's register count is
// the maximum over its call tree, out-of-line redo paths included -- never below 2.
int pack_ctas_per_sm() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 4;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n > 0) return n;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, prep_kernel, kPackThreads, 0) != cudaSuccess || n < 2) n = n < 2 ? 2 : 4;
    if (n > 6) n = 6;
    cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

// cudaFuncSetAttribute is per device and PROCESS-wide (a smaller value set by another host thread
// would lower the limit under a launch already sized for the larger one): the largest dynamic
// shared-memory size granted to each kernel is remembered per device for the whole process and only
// ever raised, under one lock.
constexpr int kAttrDevices = 64;
struct SmemGrant { std::atomic<size_t> bytes[kAttrDevices]; };
std::mutex g_grant_mutex;
template <typename Kernel>
cudaError_t grant_smem(SmemGrant &g, Kernel kernel, size_t smem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= kAttrDevices) {
        std::lock_guard<std::mutex> lock(g_grant_mutex);
        return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 << 10));
    }
    if (smem <= g.bytes[dev].load(std::memory_order_acquire)) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_grant_mutex);
    if (smem <= g.bytes[dev].load(std::memory_order_relaxed)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) g.bytes[dev].store(smem, std::memory_order_release);
    return e;
}

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
    L.off_neg = off;     off = align_up(off + g * (size_t)L.Tp * 4, 256);
    size_t w = (size_t)(W > 0 ? W : 1);
    L.off_win_lo = off;  off = align_up(off + w * 4, 256);
    L.off_win_hi = off;  off = align_up(off + w * 4, 256);
    L.off_win_pfirst = off; off = align_up(off + w * 4, 256);
    L.off_win_plast = off;  off = align_up(off + w * 4, 256);
    L.off_slow = off;    off = align_up(off + w * 4, 256);
    L.off_bail = off;    off = align_up(off + w * 4, 256);
    // pairwise list arena: at least one slab of the worst-case window must fit
    size_t worst = (size_t)(V > 0 ? V : 1);
    size_t want;
    if (max_win_variants > 0) {
        size_t cap = align_up((size_t)max_win_variants, 8);
        want = cap * (size_t)(kNumSM * kPairWarps);
        if (want > (size_t)16 << 20) want = (size_t)16 << 20;
        if (want < cap) want = cap;
    } else {
        want = worst > ((size_t)4 << 20) ? worst : ((size_t)4 << 20);
        if (want > worst * (size_t)(kNumSM * kPairWarps)) want = worst * (size_t)(kNumSM * kPairWarps);
    }
    // whatever the hint says, one slab of a window spanning every variant always fits: an
    // understated hint costs parallelism in the pairwise kernel, never results
    if (want < worst + 8) want = worst + 8;
    L.arena_entries = align_up(want, 8);
    L.off_arena = off;   off = align_up(off + L.arena_entries * 13, 256);
    L.total = off;
    return L;
}

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

template <typename... KArgs, typename... Args>
cudaError_t launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st,
                   bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    ++g_launches;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// One "prep" grid: pack CTAs for the 32-variant groups [g_begin, g_end) (if any) followed by
// window-bounds CTAs for W windows (if W > 0).
int launch_prep(cudaStream_t st, int64_t g_begin, int64_t g_end, const int8_t *d_ref, const int8_t *d_tgt,
                int64_t V, int32_t R, int32_t T, const int32_t *d_pos, const int32_t *d_ws, const int32_t *d_we,
                const int32_t *d_seg, int32_t fixed_s, int32_t fixed_e, int32_t W, uint8_t *ws,
                const WsLayout &L, int pack_ctas = 0, const int32_t *const *fetch_src = nullptr) {
    PackArgs p = {};
    BoundsArgs b = {};
    if (fetch_src && W > 0) {  // {pos, win_start, win_end} in mapped host memory
        b.src_pos = fetch_src[0]; b.src_ws = fetch_src[1]; b.src_we = fetch_src[2];
        const int64_t words = V + 2 * (int64_t)W;
        b.n_fetch = (int32_t)((words + 2047) / 2048);
        if (b.n_fetch > kMaxFetchCtas) b.n_fetch = kMaxFetchCtas;
    }
    int n_pack = 0, mode = 0;
    size_t smem = 0;
    const int n_bounds = W > 0 ? (W + kBoundsWindowsPerCta - 1) / kBoundsWindowsPerCta : 0;
    if (g_end > g_begin) {
        const int64_t ngroups = g_end - g_begin;
        p.ref = d_ref; p.tgt = d_tgt;
        p.V = (int32_t)V; p.R = R; p.T = T; p.Tp = L.Tp;
        p.g_begin = (int32_t)g_begin; p.g_end = (int32_t)g_end;
        p.absent_bits = reinterpret_cast<uint32_t *>(ws + L.off_absent);
        p.exotic = reinterpret_cast<uint32_t *>(ws + L.off_exotic);
        p.carried = reinterpret_cast<uint32_t *>(ws + L.off_carried);
        p.two = reinterpret_cast<uint32_t *>(ws + L.off_two);
        p.neg = reinterpret_cast<uint32_t *>(ws + L.off_neg);
        // groups per CTA: staged tile <= ~64 KB, at most 8 groups; small problems get the fewest
        // groups per CTA that still fit the whole grid in ONE wave (the pack kernel's CTAs per SM by registers)
        const size_t per_group = (size_t)32 * ((size_t)R + (size_t)T);
        int gpc_max = (int)((size_t)(64 << 10) / per_group);
        if (gpc_max > 32) gpc_max = 32;  // narrow panels (a few dozen columns): tiles of up to 1,024 variants
        if (gpc_max < 1) gpc_max = 1;
        const int nsm = num_sms();
        const int64_t wave = (int64_t)pack_ctas_per_sm() * nsm - n_bounds;
        int gpc = 1;
        while (gpc < gpc_max && (ngroups + gpc - 1) / gpc > (wave > nsm ? wave : nsm)) ++gpc;
        if ((ngroups + gpc - 1) / gpc > 4 * (int64_t)pack_ctas_per_sm() * nsm) gpc = gpc_max;  // many waves anyway: biggest tiles
        p.gpc = gpc;
        p.ref_tile_bytes = (uint32_t)align_up((size_t)gpc * 32 * R + 32, 16);
        p.tgt_tile_bytes = (uint32_t)align_up((size_t)gpc * 32 * T + 32, 16);
        const size_t fixed = (size_t)gpc * 32 + align16((uint32_t)gpc * 8);
        smem = fixed + p.ref_tile_bytes + p.tgt_tile_bytes;
        n_pack = (int)((ngroups + gpc - 1) / gpc);
        // Wide panels: once three 32-row tiles no longer fit an SM, stage the group's reference rows
        // with ONE column chunk of its target rows (chunks of a multiple of 128 columns, tile <= 73
        // KB).  Only a reference panel too wide for that keeps whole rows (while one tile fits) or
        // the byte path.
        const size_t ref_b = align_up((size_t)32 * R + 32, 16);
        if (smem > (size_t)73 << 10 && ref_b <= ((size_t)40 << 10)) {
            mode = 2;
            p.gpc = 1;
            // chunk width: 2,048 or 1,024 columns give the 256 threads whole rounds of 4-column items
            // (phase B); otherwise the widest multiple of 128 that keeps the tile within 73 KB
            // (row segments that are not 16-byte aligned -- T not a multiple of 16 -- are staged by
            // vector loads keeping each row's phase: 16 spare bytes per row)
            const bool seg16 = (T & 15) == 0 && ((uintptr_t)d_tgt & 15) == 0;
            int32_t ccols = (int32_t)((((size_t)73 << 10) - ref_b - 64 - (seg16 ? 0 : 32 * 16)) / 32 / 128 * 128);
            if (ccols >= 2048) ccols = 2048;
            else if (ccols >= 1024) ccols = 1024;
            p.cchunks = (T + ccols - 1) / ccols;
            p.ccols = ccols;
            p.tpitch = seg16 ? ccols : ccols + 16;
            p.ref_tile_bytes = (uint32_t)ref_b;
            p.tgt_tile_bytes = (uint32_t)(32 * p.tpitch + 16);
            smem = (size_t)32 + align16(8u) + p.ref_tile_bytes + p.tgt_tile_bytes;
            if (ngroups * p.cchunks > (int64_t)INT32_MAX) return fail(SSTAR_B200_EINVAL, "too many pack tiles%s");
            n_pack = (int)(ngroups * p.cchunks);
            // the chunks of a group OR their exotic flags together
            cudaError_t em = cudaMemsetAsync(p.exotic + g_begin, 0, (size_t)ngroups * 4, st);
            if (em != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(em));
        } else if (smem > (size_t)200 << 10) {  // rows too wide to stage: byte path straight from global
            mode = 1;
            smem = fixed;
        }
        p.n_tiles = n_pack;
        if (pack_ctas > 0 && pack_ctas < n_pack) n_pack = pack_ctas;
    }
    if (W > 0) {
        b.pos = d_pos; b.win_start = d_ws; b.win_end = d_we; b.seg = d_seg;
        b.fixed_start = fixed_s; b.fixed_end = fixed_e;
        b.V = (int32_t)V; b.W = W;
        b.win_lo = reinterpret_cast<int32_t *>(ws + L.off_win_lo);
        b.win_hi = reinterpret_cast<int32_t *>(ws + L.off_win_hi);
        b.win_pfirst = reinterpret_cast<int32_t *>(ws + L.off_win_pfirst);
        b.win_plast = reinterpret_cast<int32_t *>(ws + L.off_win_plast);
        b.win_bail = reinterpret_cast<uint32_t *>(ws + L.off_bail);
        b.hdr = reinterpret_cast<WsHeader *>(ws);
    }
    if (n_pack + n_bounds == 0) return 0;
    const int n_fetch = b.n_fetch;
    // three instances of one body: whole-row tiles (the common one), column chunks
    // for wide panels, and whole-row tiles with one lane per reference row for panels of <= 16
    // reference columns (replicate batches)
    static SmemGrant prep_grant, wide_grant, narrow_grant;
    auto kernel = prep_kernel;
    SmemGrant *grant = &prep_grant;
    if (mode == 2) { kernel = prep_wide_kernel; grant = &wide_grant; }
    else if (mode == 0 && n_pack > 0 && R <= 16) { kernel = prep_narrow_kernel; grant = &narrow_grant; }
    cudaError_t e = grant_smem(*grant, kernel, smem);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaFuncSetAttribute(prep): %s", cudaGetErrorString(e));
    if (n_fetch > 0) {  // the fetch first; the prep grid is its programmatic dependent
        e = launch(fetch_kernel, (unsigned)n_fetch, kPackThreads, 0, st, false, b);
        if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "fetch_kernel launch: %s", cudaGetErrorString(e));
    }
    e = launch(kernel, (unsigned)(n_pack + n_bounds), kPackThreads, smem, st, n_fetch > 0, p, b, n_pack, mode);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "prep_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// CTAs of the pack grid when the matrices are read in place over PCIe: 32 already saturate the
// link (profiles/r01_side_measurements.txt); a short grid keeps the tiles arriving front to back
constexpr int kInPlacePackCtas = 64;

struct ScoreCall {
    const int8_t *d_tgt;
    const int32_t *d_pos;
    int64_t V;
    int32_t T;
    int32_t bonus, max_mm, penalty;
    int64_t *d_score;
    int32_t *d_nsnp;
    uint8_t *ws;
    int algo;
    int32_t max_win_variants;
    bool disjoint;  // every window is its own replicate (positions restart): one window per CTA
};

// Scores windows [w_begin, w_end) (bounds and planes must be in the workspace): the grouped
// prefix-max kernel, then the pairwise kernel for whatever it routed to the slow list.
int launch_score(cudaStream_t st, const ScoreCall &c, const WsLayout &L, int32_t w_begin, int32_t w_end, int slot) {
    const int32_t W = w_end - w_begin;
    if (W <= 0) return 0;
    if (slot < 0 || slot >= kMaxSlots) return fail(SSTAR_B200_EINVAL, "internal: bad score slot%s");
    uint8_t *ws = c.ws;
    ScoreArgs s = {};
    s.pos = c.d_pos; s.tgt = c.d_tgt;
    s.absent_bits = reinterpret_cast<const uint32_t *>(ws + L.off_absent);
    s.exotic = reinterpret_cast<const uint32_t *>(ws + L.off_exotic);
    s.carried = reinterpret_cast<const uint32_t *>(ws + L.off_carried);
    s.two = reinterpret_cast<const uint32_t *>(ws + L.off_two);
    s.neg = reinterpret_cast<const uint32_t *>(ws + L.off_neg);
    s.win_lo = reinterpret_cast<const int32_t *>(ws + L.off_win_lo);
    s.win_hi = reinterpret_cast<const int32_t *>(ws + L.off_win_hi);
    s.win_pfirst = reinterpret_cast<const int32_t *>(ws + L.off_win_pfirst);
    s.win_plast = reinterpret_cast<const int32_t *>(ws + L.off_win_plast);
    s.slow_list = reinterpret_cast<int32_t *>(ws + L.off_slow);
    s.win_bail = reinterpret_cast<uint32_t *>(ws + L.off_bail);
    s.hdr = reinterpret_cast<WsHeader *>(ws);
    s.V = (int32_t)c.V; s.T = c.T; s.Tp = L.Tp;
    s.w_begin = w_begin; s.w_end = w_end; s.slot = slot;
    s.force_slow = (c.algo == SSTAR_B200_ALGO_PAIRWISE);
    s.bonus = c.bonus; s.max_mm = c.max_mm; s.penalty = c.penalty;
    s.score = c.d_score; s.nsnp = c.d_nsnp;
    s.arena = ws + L.off_arena; s.arena_entries = L.arena_entries;
    const int32_t T = c.T;

    if (!s.force_slow) {
        int bd = (T + 31) / 32 * 32;
        if (bd > kScoreMaxThreads) bd = kScoreMaxThreads;
        const int tchunks = (T + bd - 1) / bd;
        s.tchunks = tchunks;
        // Windows per CTA: up to 16 consecutive ones (the list build and the per-CTA set-up are
        // shared by them; the kernel itself cuts a group into sub-groups that fit its position
        // tile), but keep at least two full waves of CTAs (7 resident per SM) so the tail stays short.
        int words_cap = kScoreWordsCap;
        int wg = kMaxGroup;
        while (wg > 1 && (int64_t)((W + wg - 1) / wg) * tchunks < 2 * 7 * num_sms()) wg >>= 1;
        if (c.disjoint) wg = 1;  // replicates share no variants (and their positions restart)
        const int64_t ngroups = (W + wg - 1) / wg;
        if (ngroups * tchunks > (int64_t)INT32_MAX) return fail(SSTAR_B200_EINVAL, "too many (window group, column chunk) tiles%s");
        const int ls = score_list_stride(T, bd);      // words between a column's list entries
        int Lslots = kScoreListBytes / (ls * 4) - 1;  // list slots per column (one more row is the spill slot)
        if (Lslots > 1000) Lslots = 1000;             // (the worklist keeps a unit's first entry in 10 bits)
        size_t pos_b = (size_t)words_cap * 32 * 4;    // position tile
        if (wg == 1 && c.max_win_variants > 0) {
            // one window per CTA and its size is known: neither a column's list nor the position
            // tile can outgrow it -- narrow launches (replicate batches) then keep more CTAs resident
            if (Lslots > c.max_win_variants + 1) Lslots = c.max_win_variants + 1;
            const int need_words = (c.max_win_variants + 31) / 32 + 1;
            if (need_words < words_cap) words_cap = need_words;  // (a larger window than promised is walked from the planes)
            pos_b = (size_t)words_cap * 32 * 4;
        }
        const size_t ring_b = (size_t)wg * bd * 4;  // worklist (shares the position tile's region)
        const size_t smem = (size_t)(Lslots + 1) * ls * 4 + (ring_b > pos_b ? ring_b : pos_b);  // +1: spill slot
        static SmemGrant group_grant, narrow_grant;
        const bool narrow = ls != bd;
        cudaError_t e = narrow ? grant_smem(narrow_grant, score_group_narrow_kernel, smem)
                               : grant_smem(group_grant, score_group_kernel, smem);
        if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaFuncSetAttribute(score_group): %s", cudaGetErrorString(e));
        e = launch(narrow ? score_group_narrow_kernel : score_group_kernel, (unsigned)(ngroups * tchunks), (unsigned)bd, smem,
                   st, true, s, wg, Lslots, words_cap);
        if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "score_group_kernel launch: %s", cudaGetErrorString(e));
    }
    // list entries per warp kept in shared memory: the largest window when the caller told us (more
    // resident warps to hide the gather's latency), else 1024; larger windows use the arena
    int32_t pcap = kPairSmemCap;
    if (c.max_win_variants > 0 && c.max_win_variants < kPairSmemCap)
        pcap = (int32_t)align_up((size_t)(c.max_win_variants < 64 ? 64 : c.max_win_variants), 8);
    s.pair_smem_cap = pcap;
    const size_t psmem = (size_t)kPairWarps * pcap * 13;
    static SmemGrant pair_grant;
    cudaError_t ep = grant_smem(pair_grant, score_pairwise_kernel, psmem);
    if (ep != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaFuncSetAttribute(pairwise): %s", cudaGetErrorString(ep));
    int per_sm = (int)(((size_t)200 << 10) / (psmem + 1024));
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 4 || !s.force_slow) per_sm = 4;  // behind the grouped kernel the slow list is short or empty: keep the launch light
    int64_t pgrid = ((int64_t)W * T + kPairWarps - 1) / kPairWarps;
    if (pgrid > (int64_t)num_sms() * per_sm) pgrid = (int64_t)num_sms() * per_sm;
    cudaError_t e = launch(score_pairwise_kernel, (unsigned)pgrid, kPairWarps * 32, psmem, st, true, s);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "score_pairwise_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// The one-launch path (fused_kernels.cuh).  Returns 1 when it scored the call, 0 when the shapes do
// not qualify (the caller then takes the pack -> score launches), a negative code on error.
int launch_fused(cudaStream_t st, const int8_t *d_ref, const int8_t *d_tgt, const int32_t *d_pos, int64_t V, int32_t R,
                 int32_t T, const int32_t *d_ws, const int32_t *d_we, const int32_t *d_seg, int32_t fixed_s,
                 int32_t fixed_e, int32_t W, int32_t bonus, int32_t max_mm, int32_t penalty, int64_t *d_score,
                 int32_t *d_nsnp, int32_t hint, int32_t flags) {
    if (flags & SSTAR_B200_FLAG_NO_FUSED) return 0;
    if (T > 128 || W <= 0) return 0;  // (a window's columns must fit one CTA: T <= the replicate kernel's 256 threads)
    const bool force = (flags & SSTAR_B200_FLAG_FORCE_FUSED) != 0;
    const bool disjoint = d_seg != nullptr;  // replicates: no row is shared by two windows
    const bool indep = (flags & SSTAR_B200_FLAG_INDEPENDENT) != 0;
    // independent calls in a stream take the 64-register instance, so that the next call's CTA fits an
    // SM beside this one's (two 512-thread CTAs and two tiles per SM)
    static const int pair_env = [] { const char *e = getenv("SSTAR_B200_FUSED_PAIR"); return e ? atoi(e) : 1; }();
    if (!force) {
        if (hint <= 0) return 0;  // the tile is sized by the largest window
        // overlapping windows re-read rows through the L2: only while the whole input sits there
        static const int64_t max_bytes = [] {
            const char *e = getenv("SSTAR_B200_FUSED_MAX_MB");
            return (int64_t)(e ? atoll(e) : 16) << 20;
        }();
        if (!disjoint && V * ((int64_t)R + T) > max_bytes) return 0;
    }
    const int32_t tc = (T + 31) / 32 * 32;
    const int64_t rows1 = hint > 0 ? hint : 256;
    FusedArgs a = {};
    a.tc = tc;
    auto bytes_for = [&](int64_t c) {
        return align_up((size_t)c * 32 * R + 32, 16) + align_up((size_t)c * 32 * T + 32, 16) + (size_t)c * 32 * 4 +
               2 * (size_t)c * tc * 4 + 2 * (size_t)c * 4 + (size_t)c * 32 + 16 + 4 * (size_t)kFusedListCap * 4;
    };
    size_t limit = (size_t)100 << 10;  // two CTAs per SM at least
    int64_t cap = 0;
    unsigned grid = 1;
    if (disjoint) {
        // a CTA per `wg` consecutive replicates (as many as it scores at a time), their rows side by side
        int32_t wg = kFusedRepThreads / T;
        if (wg > kFusedMaxSlots) wg = kFusedMaxSlots;
        if (wg < 1) wg = 1;
        for (;; wg = wg / 2) {
            cap = (rows1 * wg + 31) / 32 + 1;  // + the range's phase inside its first group
            if (bytes_for(cap) <= limit || wg == 1) break;
        }
        a.wg = wg;
        grid = (unsigned)((W + wg - 1) / wg);
    } else {
        // a CTA per S home rows; its tile adds the largest window's rows behind them
        // S: one CTA per SM when the problem allows (a single wave, every SM busy from the first
        // instruction), but no more windows per CTA than it scores at a time
        static const int32_t home_env = [] {
            const char *e = getenv("SSTAR_B200_FUSED_S");
            const int v = e ? atoi(e) : 0;
            return v >= 32 ? v / 32 * 32 : 0;
        }();
        int32_t slots = kFusedThreads / T;
        if (slots > kFusedMaxSlots) slots = kFusedMaxSlots;
        const int64_t s_max = V > 0 ? (V * slots / (W > 0 ? W : 1)) / 32 * 32 : 32;  // ~`slots` windows per CTA
        int32_t S = 0;
        // a call of its own: one 512-thread CTA per SM, one wave, every SM busy from the first instruction.
        // One of a stream of independent calls (tried first when flagged): the SMs are kept busy by the
        // neighbouring calls, so a CTA takes as many windows as it scores at a time (fewer CTAs, less
        // halo re-read and re-packed) and its tile leaves room for a CTA of the next call beside it
        for (int stream_mode = (indep && pair_env) ? 1 : 0; stream_mode >= 0; --stream_mode) {
            S = home_env;
            if (S == 0) {
                const int nsm = num_sms();
                S = stream_mode ? (int32_t)(s_max < 1024 ? s_max : 1024) : (int32_t)(((V + nsm - 1) / nsm + 31) / 32 * 32);
                if (S > s_max) S = (int32_t)s_max;
                if (S < 64) S = 64;
                if (S > 1024) S = 1024;
            }
            limit = stream_mode ? ((size_t)111 << 10) : ((size_t)160 << 10);
            for (;; S = stream_mode ? S - 32 : (S / 2 + 31) / 32 * 32) {
                cap = (S + rows1 + 31) / 32;
                if (bytes_for(cap) <= limit || S == 32) break;
            }
            if (bytes_for(cap) <= limit) break;
        }
        a.home_rows = S;
        a.wg = 0;
        grid = (unsigned)(V > 0 ? (V + S - 1) / S : 1);
        if ((int64_t)grid * W > ((int64_t)64 << 20) && !force) return 0;  // the list scan is O(W) per CTA
    }
    if (bytes_for(cap) > limit) {
        if (!force) return 0;
        while (cap > 2 && bytes_for(cap) > limit) --cap;  // forced: windows beyond the tile are streamed
        if (!disjoint && cap * 32 < a.home_rows) { a.home_rows = (int32_t)(cap * 32); grid = (unsigned)(V > 0 ? (V + a.home_rows - 1) / a.home_rows : 1); }
        if (bytes_for(cap) > ((size_t)200 << 10)) return 0;
    }
    size_t smem = bytes_for(cap);
    a.cap_groups = (int32_t)cap;
    a.ref_tile_bytes = (uint32_t)align_up((size_t)cap * 32 * R + 32, 16);
    a.tgt_tile_bytes = (uint32_t)align_up((size_t)cap * 32 * T + 32, 16);
    if (smem < kFusedStreamMinBytes) smem = kFusedStreamMinBytes;
    smem = align_up(smem, 16);
    a.smem_bytes = (uint32_t)smem;
    a.ref = d_ref; a.tgt = d_tgt; a.pos = d_pos; a.win_start = d_ws; a.win_end = d_we; a.seg = d_seg;
    a.fixed_start = fixed_s; a.fixed_end = fixed_e;
    a.V = (int32_t)V; a.R = R; a.T = T; a.W = W;
    a.bonus = bonus; a.max_mm = max_mm; a.penalty = penalty;
    a.score = d_score; a.nsnp = d_nsnp;
    a.early_trigger = indep ? 1 : 0;
    static SmemGrant fused_grant, rep_grant, pair_grant;
    const bool pair = indep && !disjoint && pair_env && 2 * (smem + 2048) <= ((size_t)227 << 10);
    cudaError_t e = disjoint ? grant_smem(rep_grant, fused_score_rep_kernel, smem)
                    : pair   ? grant_smem(pair_grant, fused_score_pair_kernel, smem)
                             : grant_smem(fused_grant, fused_score_kernel, smem);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaFuncSetAttribute(fused): %s", cudaGetErrorString(e));
    e = disjoint ? launch(fused_score_rep_kernel, grid, kFusedRepThreads, smem, st, indep, a)
        : pair   ? launch(fused_score_pair_kernel, grid, kFusedThreads, smem, st, indep, a)
                 : launch(fused_score_kernel, grid, kFusedThreads, smem, st, indep, a);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "fused_score_kernel launch: %s", cudaGetErrorString(e));
    return 1;
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

// Per (host thread, device): one copy stream and a ring of events for the host pipeline.  Created
// on first use and kept for the life of the thread (not caller-visible memory).
constexpr int kMaxChunks = 16;
constexpr int kMaxDevices = 64;
// A stream of independent host calls (SSTAR_B200_FLAG_INDEPENDENT): the hazards of a scratch buffer
// that comes round again are tracked HERE, per scratch -- its next call's copies wait for the event
// recorded behind the kernels of its previous call, nothing else.
constexpr int kMaxScratchSlots = 32;
struct ScratchSlot {
    const void *scratch = nullptr;
    cudaEvent_t in = nullptr, in2 = nullptr, done = nullptr;
    bool live = false;  // `done` is recorded and nothing unflagged has used the scratch since
    uint64_t tick = 0;
};
struct PipeCtx {
    bool ready = false;
    cudaStream_t copy = nullptr, copy2 = nullptr;
    cudaEvent_t entry = nullptr, done = nullptr;
    cudaEvent_t chunk[kMaxChunks] = {};
    ScratchSlot slots[kMaxScratchSlots];
    uint64_t tick = 0;
};
thread_local PipeCtx g_pipe[kMaxDevices];

int pipe_ctx(int device, PipeCtx **out) {
    if (device < 0 || device >= kMaxDevices) return fail(SSTAR_B200_EINVAL, "device index out of range%s");
    PipeCtx &c = g_pipe[device];
    if (!c.ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c.copy2, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&c.entry, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
        for (int i = 0; i < kMaxChunks; ++i) CUDA_TRY(cudaEventCreateWithFlags(&c.chunk[i], cudaEventDisableTiming));
        c.ready = true;
    }
    *out = &c;
    return 0;
}

// the slot of `scratch` (least recently used one recycled); *fresh: it carries no recorded `done`
int scratch_slot(PipeCtx *pc, const void *scratch, ScratchSlot **out, bool *fresh) {
    ScratchSlot *hit = nullptr, *lru = &pc->slots[0];
    for (ScratchSlot &sl : pc->slots) {
        if (sl.scratch == scratch) { hit = &sl; break; }
        if (sl.tick < lru->tick) lru = &sl;
    }
    if (!hit) {
        hit = lru;
        hit->scratch = scratch;
        hit->live = false;
        if (!hit->in) {
            CUDA_TRY(cudaEventCreateWithFlags(&hit->in, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&hit->in2, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&hit->done, cudaEventDisableTiming));
        }
    }
    hit->tick = ++pc->tick;
    *fresh = !hit->live;
    *out = hit;
    return 0;
}
// an unflagged call is about to use `scratch`: whatever a flagged call recorded for it is void
void scratch_forget(PipeCtx *pc, const void *scratch) {
    for (ScratchSlot &sl : pc->slots)
        if (sl.scratch == scratch) sl.live = false;
}

// pack (optional) + bounds + score, the common body of every device-pointer entry point
int run_pipeline(int device, cudaStream_t st, bool do_pack, const int8_t *d_ref, const int8_t *d_tgt,
                 const int32_t *d_pos, int64_t V, int32_t R, int32_t T, const int32_t *d_ws,
                 const int32_t *d_we, const int32_t *d_seg, int32_t fixed_s, int32_t fixed_e, int32_t W,
                 int32_t bonus, int32_t max_mm, int32_t penalty, int64_t *d_score, int32_t *d_nsnp,
                 uint8_t *ws, const WsLayout &L, int algo, int32_t max_win_variants, int32_t flags = 0) {
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    if (do_pack && algo == SSTAR_B200_ALGO_AUTO) {  // small problems, replicate batches: one launch does it all
        const int frc = launch_fused(st, d_ref, d_tgt, d_pos, V, R, T, d_ws, d_we, d_seg, fixed_s, fixed_e, W, bonus, max_mm,
                                     penalty, d_score, d_nsnp, max_win_variants, flags);
        if (frc != 0) return frc < 0 ? frc : 0;
    }
    const ScoreCall c = {d_tgt, d_pos, V, T, bonus, max_mm, penalty, d_score, d_nsnp, ws, algo, max_win_variants,
                         d_seg != nullptr};
    int rc = launch_prep(st, 0, do_pack ? L.G : 0, d_ref, d_tgt, V, R, T, d_pos, d_ws, d_we, d_seg, fixed_s, fixed_e, W,
                         ws, L);
    if (rc) return rc;
    return launch_score(st, c, L, 0, W, 0);
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
    return launch_prep((cudaStream_t)stream, 0, L.G, d_ref_gt, d_tgt_gt, V, R, T, nullptr, nullptr, nullptr, nullptr,
                       0, 0, 0, (uint8_t *)d_workspace, L);
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
    return run_pipeline(device, (cudaStream_t)stream, false, nullptr, d_tgt_gt, d_pos, V, R, T, d_win_start,
                        d_win_end, nullptr, 0, 0, W, match_bonus, max_mismatch, mismatch_penalty, d_score,
                        d_nsnp, (uint8_t *)d_workspace, L, algo, opts ? opts->max_win_variants : 0);
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
    return run_pipeline(device, (cudaStream_t)stream, true, d_ref_gt, d_tgt_gt, d_pos, V, R, T, d_win_start,
                        d_win_end, nullptr, 0, 0, W, match_bonus, max_mismatch, mismatch_penalty, d_score,
                        d_nsnp, (uint8_t *)d_workspace, L, algo, opts ? opts->max_win_variants : 0,
                        opts ? opts->flags : 0);
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
    return run_pipeline(device, (cudaStream_t)stream, true, d_ref_gt, d_tgt_gt, d_pos, V, R, T, nullptr, nullptr,
                        d_rep_offset, win_start, win_end, Nrep, match_bonus, max_mismatch, mismatch_penalty,
                        d_score, d_nsnp, (uint8_t *)d_workspace, L, algo, opts ? opts->max_win_variants : 0,
                        opts ? opts->flags : 0);
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

namespace {

bool mapped_pointer(const void *h, void **dptr) {
    if (!h) { *dptr = nullptr; return true; }  // absent because V == 0 or W == 0
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, h) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    *dptr = at.devicePointer;
    return true;
}

}  // namespace

// The end-to-end call.  PCIe is the bound (the genotype matrices are 100x the kernels' HBM time),
// so the job is to keep the link busy and hide everything else under it.  With page-locked, mapped
// host buffers:
//   * matrices of 32 MB and more go host->device in row chunks on an internal copy stream (copy
//     engine: the full PCIe rate); while chunk k+1 is in flight the caller's stream packs chunk k
//     and scores every window whose variants are now all packed (windows sorted by end), so only
//     the last chunk's kernels are exposed.  pos / windows are small and latency-critical (the
//     bounds search chases them): they are copied first.
//   * smaller matrices are packed IN PLACE: copies of a few MB do not reach the link rate (and
//     crawl on some hosts), the SMs' own reads of mapped host memory hold 34-42 GB/s everywhere.
//     One prep grid does it all -- fetch CTAs bring pos / windows into the scratch, pack CTAs
//     stream the matrices over PCIe tile by tile, bounds CTAs wait for the fetch -- and no copy
//     engine or second stream is involved.
//   * result tables below 8 MB are stored by the score kernels straight into the mapped host
//     buffers (posted PCIe writes, no D2H copy latency at the end); larger ones go back as D2H
//     copies (per chunk on the chunked path, overlapping the inbound traffic).
// Pageable buffers or SSTAR_B200_FLAG_STAGED_COPIES take plain H2D / D2H copies on `stream`.
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
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    WsLayout L; int algo;
    rc = check_common(device, V, R, T, W, d + h.off_work, scratch_bytes - h.off_work, opts, &L, &algo);
    if (rc) return rc;
    uint8_t *ws = d + h.off_work;
    int8_t *d_ref = (int8_t *)(d + h.off_ref), *d_tgt = (int8_t *)(d + h.off_tgt);
    int32_t *d_pos = (int32_t *)(d + h.off_pos);

    // the pipelined path needs page-locked inputs (truly asynchronous copies next to the caller's
    // stream) and mapped output tables (stored in place by the kernels)
    void *m_score = nullptr, *m_nsnp = nullptr, *m_ref = nullptr, *m_tgt = nullptr, *m_pos = nullptr, *m_ws = nullptr,
         *m_we = nullptr;
    bool mapped = !(opts && (opts->flags & SSTAR_B200_FLAG_STAGED_COPIES));
    mapped = mapped && mapped_pointer(h_pos, &m_pos) && mapped_pointer(h_win_start, &m_ws) &&
             mapped_pointer(h_win_end, &m_we) && mapped_pointer(h_score, &m_score) && mapped_pointer(h_nsnp, &m_nsnp) &&
             mapped_pointer(h_ref_gt, &m_ref) && mapped_pointer(h_tgt_gt, &m_tgt);

    if (!mapped) {  // staged: copy in, run, copy out, all on the caller's stream
        if (device >= 0 && device < kMaxDevices && g_pipe[device].ready) scratch_forget(&g_pipe[device], d_scratch);
        if (V > 0) {
            CUDA_TRY(cudaMemcpyAsync(d_pos, h_pos, (size_t)V * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_ref, h_ref_gt, (size_t)V * R, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d_tgt, h_tgt_gt, (size_t)V * T, cudaMemcpyHostToDevice, st));
        }
        if (W > 0) {
            CUDA_TRY(cudaMemcpyAsync(d + h.off_ws, h_win_start, (size_t)W * 4, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaMemcpyAsync(d + h.off_we, h_win_end, (size_t)W * 4, cudaMemcpyHostToDevice, st));
        }
        rc = run_pipeline(device, st, true, d_ref, d_tgt, d_pos, V, R, T, (const int32_t *)(d + h.off_ws),
                          (const int32_t *)(d + h.off_we), nullptr, 0, 0, W, match_bonus, max_mismatch,
                          mismatch_penalty, (int64_t *)(d + h.off_score), (int32_t *)(d + h.off_nsnp), ws, L, algo, hint,
                          opts ? opts->flags : 0);
        if (rc) return rc;
        if (W > 0) {
            CUDA_TRY(cudaMemcpyAsync(h_score, d + h.off_score, (size_t)W * T * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_nsnp, d + h.off_nsnp, (size_t)W * T * 4, cudaMemcpyDeviceToHost, st));
        }
        return 0;
    }

    // ---- pipelined path on mapped host buffers
    PipeCtx *pc;
    rc = pipe_ctx(device, &pc);
    if (rc) return rc;
    const int32_t flags = opts ? opts->flags : 0;
    // small result tables are stored in place by the kernels (no D2H copy latency at the end);
    // large ones (>= 8 MB) go through per-chunk D2H copies, which the copy engine moves faster
    // than SM stores to host memory (measured on the C4 shard: 5.66 vs 6.08 ms)
    const bool copy_out = (flags & SSTAR_B200_FLAG_COPY_OUTPUTS) != 0 || (int64_t)W * T * 12 >= ((int64_t)8 << 20);
    // How do the genotype matrices cross PCIe?  Large inputs go through the copy engine in chunks
    // (55 GB/s, overlapped with the kernels).  Copies of a few MB do not reach that rate -- 31-50
    // GB/s depending on the host, ~12 GB/s on some -- while the SMs' own reads of mapped host
    // memory hold 34-42 GB/s everywhere, so inputs below 32 MB are packed straight out of host
    // memory: the pack CTAs' TMA tile loads ARE the transfer, and there is nothing to wait for.
    const int64_t in_bytes = V * ((int64_t)R + T);
    const bool zero_copy_in = (flags & SSTAR_B200_FLAG_ZERO_COPY_INPUTS) != 0 ||
                              (!(flags & SSTAR_B200_FLAG_COPY_INPUTS) && in_bytes < ((int64_t)32 << 20));
    // chunk the variant rows: >= 8 MB of genotypes per chunk (below that the per-chunk launches and
    // event hops cost more than the overlap wins: measured on C2, 4.9 MB), whole 256-row blocks
    // (any pack tile size divides 256 rows), at most kMaxChunks
    const int64_t row_bytes = (int64_t)R + T;
    int nchunks = (int)((V * row_bytes) / ((int64_t)8 << 20));
    if (nchunks > kMaxChunks) nchunks = kMaxChunks;
    if (nchunks < 1 || (flags & SSTAR_B200_FLAG_SINGLE_CHUNK)) nchunks = 1;
    int64_t rows_per = ((V + nchunks - 1) / nchunks + 255) / 256 * 256;
    if (rows_per < 256) rows_per = 256;
    nchunks = V > 0 ? (int)((V + rows_per - 1) / rows_per) : 1;

    if ((flags & SSTAR_B200_FLAG_INDEPENDENT) && in_bytes < ((int64_t)32 << 20) && !copy_out) {
        // One of a STREAM of independent host calls (own host buffers and own scratch each; the
        // caller cycles through a few scratch buffers).  The inputs go through the copy engine on
        // the internal copy stream -- queued right behind the previous call's copies, so the link
        // stays busy while the caller's stream scores that call -- and the kernels store both
        // tables in place.  The only thing this call's copies wait for is the previous call that
        // used the SAME scratch (its recorded `done`); a scratch seen for the first time waits for
        // everything the caller has queued.
        ScratchSlot *sl;
        bool fresh;
        rc = scratch_slot(pc, d_scratch, &sl, &fresh);
        if (rc) return rc;
        static const int two_streams = [] { const char *e = getenv("SSTAR_B200_HOST_STREAMS"); return e ? atoi(e) : 1; }() > 1;
        cudaStream_t c1 = pc->copy, c2 = two_streams ? pc->copy2 : pc->copy;
        if (fresh) {
            CUDA_TRY(cudaEventRecord(pc->entry, st));
            CUDA_TRY(cudaStreamWaitEvent(c1, pc->entry, 0));
            if (c2 != c1) CUDA_TRY(cudaStreamWaitEvent(c2, pc->entry, 0));
        } else {
            CUDA_TRY(cudaStreamWaitEvent(c1, sl->done, 0));
            if (c2 != c1) CUDA_TRY(cudaStreamWaitEvent(c2, sl->done, 0));
        }
        // (SSTAR_B200_HOST_STREAMS=2: the reference matrix on one copy stream, everything else on a
        // second one -- measured equal, 112.8 vs 114.1 us per C2 call: the link, not the per-copy cost,
        // is the limit)
        if (V > 0) {
            CUDA_TRY(cudaMemcpyAsync(d_ref, h_ref_gt, (size_t)V * R, cudaMemcpyHostToDevice, c1));
            CUDA_TRY(cudaMemcpyAsync(d_pos, h_pos, (size_t)V * 4, cudaMemcpyHostToDevice, c2));
            CUDA_TRY(cudaMemcpyAsync(d_tgt, h_tgt_gt, (size_t)V * T, cudaMemcpyHostToDevice, c2));
        }
        if (W > 0) {
            CUDA_TRY(cudaMemcpyAsync(d + h.off_ws, h_win_start, (size_t)W * 4, cudaMemcpyHostToDevice, c2));
            CUDA_TRY(cudaMemcpyAsync(d + h.off_we, h_win_end, (size_t)W * 4, cudaMemcpyHostToDevice, c2));
        }
        CUDA_TRY(cudaEventRecord(sl->in, c1));
        CUDA_TRY(cudaStreamWaitEvent(st, sl->in, 0));
        if (c2 != c1) {
            CUDA_TRY(cudaEventRecord(sl->in2, c2));
            CUDA_TRY(cudaStreamWaitEvent(st, sl->in2, 0));
        }
        rc = run_pipeline(device, st, true, d_ref, d_tgt, d_pos, V, R, T, (const int32_t *)(d + h.off_ws),
                          (const int32_t *)(d + h.off_we), nullptr, 0, 0, W, match_bonus, max_mismatch,
                          mismatch_penalty, (int64_t *)m_score, (int32_t *)m_nsnp, ws, L, algo, hint, flags);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(sl->done, st));
        sl->live = true;
        return 0;
    }
    scratch_forget(pc, d_scratch);
    if (zero_copy_in) {
        // Small input read in place, in ONE prep grid on the caller's stream and without the copy
        // engine.  Its first CTAs fetch pos / windows from mapped host memory; a short run of pack
        // CTAs walks the matrices front to back, their tile loads being the transfer; the bounds
        // CTAs wait for the fetch and resolve the windows.
        const int32_t *const fetch_src[3] = {(const int32_t *)m_pos, (const int32_t *)m_ws, (const int32_t *)m_we};
        rc = launch_prep(st, 0, L.G, (const int8_t *)m_ref, (const int8_t *)m_tgt, V, R, T, d_pos,
                         (const int32_t *)(d + h.off_ws), (const int32_t *)(d + h.off_we), nullptr, 0, 0, W, ws, L,
                         kInPlacePackCtas, fetch_src);
        if (rc) return rc;
        int64_t *z_score = copy_out ? (int64_t *)(d + h.off_score) : (int64_t *)m_score;
        int32_t *z_nsnp = copy_out ? (int32_t *)(d + h.off_nsnp) : (int32_t *)m_nsnp;
        const ScoreCall zc = {(const int8_t *)m_tgt, d_pos, V, T, match_bonus, max_mismatch, mismatch_penalty, z_score,
                              z_nsnp, ws, algo, hint, false};
        rc = launch_score(st, zc, L, 0, W, 0);
        if (rc) return rc;
        if (copy_out && W > 0) {
            CUDA_TRY(cudaMemcpyAsync(h_score, z_score, (size_t)W * T * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_nsnp, z_nsnp, (size_t)W * T * 4, cudaMemcpyDeviceToHost, st));
        }
        return 0;
    }
    // windows can be scored chunk by chunk only if they are sorted by end
    bool sorted = nchunks > 1;
    for (int32_t w = 1; w < W && sorted; ++w) sorted = h_win_end[w] >= h_win_end[w - 1];
    // the copy stream starts after whatever the caller already queued (it reuses the scratch)
    CUDA_TRY(cudaEventRecord(pc->entry, st));
    CUDA_TRY(cudaStreamWaitEvent(pc->copy, pc->entry, 0));
    if (V > 0) CUDA_TRY(cudaMemcpyAsync(d_pos, h_pos, (size_t)V * 4, cudaMemcpyHostToDevice, st));
    if (W > 0) {
        CUDA_TRY(cudaMemcpyAsync(d + h.off_ws, h_win_start, (size_t)W * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d + h.off_we, h_win_end, (size_t)W * 4, cudaMemcpyHostToDevice, st));
    }
    // bounds first: they only need pos and the windows
    rc = launch_prep(st, 0, 0, nullptr, nullptr, V, R, T, d_pos, (const int32_t *)(d + h.off_ws),
                     (const int32_t *)(d + h.off_we), nullptr, 0, 0, W, ws, L);
    if (rc) return rc;
    int64_t *o_score = copy_out ? (int64_t *)(d + h.off_score) : (int64_t *)m_score;
    int32_t *o_nsnp = copy_out ? (int32_t *)(d + h.off_nsnp) : (int32_t *)m_nsnp;
    const ScoreCall c = {d_tgt, d_pos, V, T, match_bonus, max_mismatch, mismatch_penalty, o_score, o_nsnp, ws, algo, hint,
                         false};
    int32_t w_done = 0;
    for (int k = 0; k < nchunks; ++k) {
        const int64_t r0 = (int64_t)k * rows_per, r1 = (r0 + rows_per < V) ? r0 + rows_per : V;
        if (r1 > r0) {
            CUDA_TRY(cudaMemcpyAsync(d_ref + r0 * R, h_ref_gt + r0 * R, (size_t)(r1 - r0) * R, cudaMemcpyHostToDevice, pc->copy));
            CUDA_TRY(cudaMemcpyAsync(d_tgt + r0 * T, h_tgt_gt + r0 * T, (size_t)(r1 - r0) * T, cudaMemcpyHostToDevice, pc->copy));
            CUDA_TRY(cudaEventRecord(pc->chunk[k], pc->copy));
            CUDA_TRY(cudaStreamWaitEvent(st, pc->chunk[k], 0));
            rc = launch_prep(st, r0 / 32, (r1 + 31) / 32, d_ref, d_tgt, V, R, T, d_pos, nullptr, nullptr, nullptr, 0, 0, 0,
                             ws, L);
            if (rc) return rc;
        }
        // windows whose last variant lies in the rows packed so far: end < pos[r1] (all, at the end)
        int32_t w_ready = W;
        if (k + 1 < nchunks) {
            if (!sorted) continue;
            const int64_t limit = h_pos[r1];
            int32_t lo = w_done, hi = W;
            while (lo < hi) {
                const int32_t mid = lo + (hi - lo) / 2;
                if ((int64_t)h_win_end[mid] < limit) lo = mid + 1; else hi = mid;
            }
            w_ready = lo;
        }
        rc = launch_score(st, c, L, w_done, w_ready, k);
        if (rc) return rc;
        if (copy_out && w_ready > w_done) {
            CUDA_TRY(cudaMemcpyAsync(h_score + (size_t)w_done * T, o_score + (size_t)w_done * T,
                                     (size_t)(w_ready - w_done) * T * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpyAsync(h_nsnp + (size_t)w_done * T, o_nsnp + (size_t)w_done * T,
                                     (size_t)(w_ready - w_done) * T * 4, cudaMemcpyDeviceToHost, st));
        }
        w_done = w_ready;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Result emitter
// ---------------------------------------------------------------------------------------------
namespace {
int emit_row_cap(int32_t chrom_len, int32_t max_label_len, int has_replicate) {
    // chrom \t start \t end \t label \t score(sign + 17 digits + ".0") \t nsnp [\t replicate] \n
    return chrom_len + 1 + 11 + 1 + 11 + 1 + max_label_len + 1 + 20 + 1 + 11 + (has_replicate ? 12 : 0) + 1;
}
}  // namespace

extern "C" size_t sstar_b200_tsv_max_bytes(int32_t W, int32_t T, int32_t chrom_len, int32_t max_label_len,
                                           int32_t has_replicate) {
    if (W < 0 || T < 1 || chrom_len < 0 || max_label_len < 0) return 0;
    return (size_t)W * (size_t)T * (size_t)emit_row_cap(chrom_len, max_label_len, has_replicate) + 16;
}

extern "C" size_t sstar_b200_tsv_workspace_bytes(int32_t W, int32_t T) {
    if (W < 0 || T < 1) return 0;
    // enough for either row order: blocks of 256 rows along the inner index
    const size_t nb_w = (size_t)(W > 0 ? W : 1) * (size_t)((T + kEmitThreads - 1) / kEmitThreads);
    const size_t nb_s = (size_t)T * (size_t)(((W > 0 ? W : 1) + kEmitThreads - 1) / kEmitThreads);
    return 256 + align_up((nb_w > nb_s ? nb_w : nb_s) * 8, 256);
}

extern "C" int sstar_b200_format_table(int device, void *stream, const int64_t *d_score, const int32_t *d_nsnp,
                                       const int32_t *d_win_start, const int32_t *d_win_end, int32_t W, int32_t T,
                                       const char *chrom, const char *d_labels, const int32_t *d_label_off,
                                       int32_t max_label_len, const int32_t *d_replicate,
                                       const int32_t *d_col_order, const sstar_b200_table_options *topts,
                                       char *d_out, size_t out_capacity, void *d_workspace, size_t workspace_bytes) {
    g_launches = 0;
    if (W < 0 || T < 1) return fail(SSTAR_B200_EINVAL, "bad table shape%s");
    if (!chrom) return fail(SSTAR_B200_EINVAL, "chrom is null%s");
    const size_t clen = strlen(chrom);
    if (clen >= sizeof(EmitArgs::chrom)) return fail(SSTAR_B200_EINVAL, "chromosome name longer than 63 bytes%s");
    if (!d_workspace || ((uintptr_t)d_workspace & 255)) return fail(SSTAR_B200_EINVAL, "workspace null or not 256-byte aligned%s");
    if (workspace_bytes < sstar_b200_tsv_workspace_bytes(W, T)) return fail(SSTAR_B200_EWORKSPACE, "workspace too small%s");
    if (W > 0 && (!d_score || !d_nsnp || !d_win_start || !d_win_end || !d_labels || !d_label_off || !d_out))
        return fail(SSTAR_B200_EINVAL, "null table pointer%s");
    if (max_label_len < 0) return fail(SSTAR_B200_EINVAL, "max_label_len must be >= 0%s");
    const int mode = topts ? topts->mode : 0;
    const int sample_major = topts ? topts->sample_major : 0;
    if (mode != 0 && mode != 1) return fail(SSTAR_B200_EINVAL, "unknown table mode%s");
    const bool with_pred = topts && topts->d_pred_value;
    if (mode == 1 && !with_pred) return fail(SSTAR_B200_EINVAL, "BED rows need the prediction table%s");
    if (with_pred && (topts->n_pred < 1 || (mode == 0 && (!topts->d_pred_text || !topts->d_pred_off))))
        return fail(SSTAR_B200_EINVAL, "incomplete prediction table%s");
    const int max_pred_len = with_pred && mode == 0 ? topts->max_pred_len : 0;
    if (max_pred_len < 0) return fail(SSTAR_B200_EINVAL, "max_pred_len must be >= 0%s");
    const int row_cap = emit_row_cap((int32_t)clen, max_label_len, d_replicate != nullptr) + 1 + (max_pred_len > 2 ? max_pred_len : 2);
    const size_t smem = (size_t)row_cap * kEmitThreads;
    if (smem > (size_t)200 << 10) return fail(SSTAR_B200_EINVAL, "sample labels too long for the emitter%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = (uint8_t *)d_workspace;
    EmitArgs a = {};
    a.score = d_score; a.nsnp = d_nsnp; a.win_start = d_win_start; a.win_end = d_win_end;
    a.replicate = d_replicate; a.col_order = d_col_order; a.labels = d_labels; a.label_off = d_label_off;
    a.mode = mode; a.sample_major = sample_major ? 1 : 0;
    if (with_pred) {
        a.pred_text = mode == 0 ? topts->d_pred_text : nullptr;
        a.pred_off = topts->d_pred_off; a.pred_value = topts->d_pred_value; a.n_pred = topts->n_pred;
    }
    a.W = W; a.T = T; a.row_cap = row_cap;
    const int inner = a.sample_major ? W : T, outer = a.sample_major ? T : W;
    a.blocks_per_win = (inner + kEmitThreads - 1) / kEmitThreads;
    a.chrom_len = (int32_t)clen;
    memcpy(a.chrom, chrom, clen);
    a.total = reinterpret_cast<int64_t *>(ws);          // header: total bytes, status
    a.status = reinterpret_cast<uint32_t *>(ws + 8);
    a.block_base = reinterpret_cast<int64_t *>(ws + 256);
    a.out = d_out; a.capacity = (int64_t)out_capacity;
    CUDA_TRY(cudaMemsetAsync(ws, 0, 256, st));
    const int64_t nb = (int64_t)outer * a.blocks_per_win;
    if (nb > (int64_t)INT32_MAX) return fail(SSTAR_B200_EINVAL, "table too large for one emitter call%s");
    if (nb == 0 || W == 0) return 0;
    static SmemGrant emit_grant;
    CUDA_TRY(grant_smem(emit_grant, emit_format_kernel, smem));
    cudaError_t e = launch(emit_len_kernel, (unsigned)nb, kEmitThreads, 0, st, false, a);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "emit_len_kernel launch: %s", cudaGetErrorString(e));
    e = launch(emit_scan_kernel, 1u, 1024u, 0, st, false, a, nb);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "emit_scan_kernel launch: %s", cudaGetErrorString(e));
    e = launch(emit_format_kernel, (unsigned)nb, kEmitThreads, smem, st, false, a);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "emit_format_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int sstar_b200_format_tsv(int device, void *stream, const int64_t *d_score, const int32_t *d_nsnp,
                                     const int32_t *d_win_start, const int32_t *d_win_end, int32_t W, int32_t T,
                                     const char *chrom, const char *d_labels, const int32_t *d_label_off,
                                     int32_t max_label_len, const int32_t *d_replicate,
                                     const int32_t *d_col_order, char *d_out, size_t out_capacity,
                                     void *d_workspace, size_t workspace_bytes) {
    return sstar_b200_format_table(device, stream, d_score, d_nsnp, d_win_start, d_win_end, W, T, chrom, d_labels,
                                   d_label_off, max_label_len, d_replicate, d_col_order, nullptr, d_out, out_capacity,
                                   d_workspace, workspace_bytes);
}

// ---------------------------------------------------------------------------------------------
// Ingest filters
// ---------------------------------------------------------------------------------------------
extern "C" size_t sstar_b200_ingest_workspace_bytes(int64_t V) {
    if (V < 0) return 0;
    const size_t nb = (size_t)((V + kIngestRowsPerBlock - 1) / kIngestRowsPerBlock);
    return 256 + align_up((size_t)(V > 0 ? V : 1), 256) + align_up((nb > 0 ? nb : 1) * 4, 256);
}

extern "C" int sstar_b200_ingest(int device, void *stream, const int8_t *d_gt, const int32_t *d_pos,
                                 const int8_t *d_mode, int64_t V, int32_t N, const int32_t *d_ref_cols,
                                 int32_t n_ref, const int32_t *d_tgt_cols, int32_t n_tgt, int32_t is_phased,
                                 int8_t *d_ref_out, int8_t *d_tgt_out, int32_t *d_pos_out, void *d_workspace,
                                 size_t workspace_bytes) {
    return sstar_b200_ingest_ex(device, stream, d_gt, d_pos, d_mode, V, N, d_ref_cols, n_ref, n_ref, d_tgt_cols, n_tgt,
                                n_tgt, is_phased, d_ref_out, d_tgt_out, d_pos_out, d_workspace, workspace_bytes);
}

extern "C" int sstar_b200_ingest_ex(int device, void *stream, const int8_t *d_gt, const int32_t *d_pos,
                                    const int8_t *d_mode, int64_t V, int32_t N, const int32_t *d_ref_cols,
                                    int32_t n_ref, int32_t n_ref_listed, const int32_t *d_tgt_cols, int32_t n_tgt,
                                    int32_t n_tgt_listed, int32_t is_phased, int8_t *d_ref_out, int8_t *d_tgt_out,
                                    int32_t *d_pos_out, void *d_workspace, size_t workspace_bytes) {
    g_launches = 0;
    if (V < 0 || V > (int64_t)INT32_MAX - 64) return fail(SSTAR_B200_EINVAL, "V out of range%s");
    if (N < 1 || n_ref < 1 || n_tgt < 1) return fail(SSTAR_B200_EINVAL, "at least one sample per population is required%s");
    if (!d_workspace || ((uintptr_t)d_workspace & 255)) return fail(SSTAR_B200_EINVAL, "workspace null or not 256-byte aligned%s");
    if (workspace_bytes < sstar_b200_ingest_workspace_bytes(V)) return fail(SSTAR_B200_EWORKSPACE, "workspace too small%s");
    if (!d_ref_cols || !d_tgt_cols) return fail(SSTAR_B200_EINVAL, "null column list%s");
    if (V > 0 && (!d_gt || !d_pos || !d_ref_out || !d_tgt_out || !d_pos_out)) return fail(SSTAR_B200_EINVAL, "null pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = (uint8_t *)d_workspace;
    CUDA_TRY(cudaMemsetAsync(ws, 0, 256, st));
    if (V == 0) return 0;
    const int64_t nb = (V + kIngestRowsPerBlock - 1) / kIngestRowsPerBlock;
    IngestArgs a = {};
    a.gt = d_gt; a.pos = d_pos; a.mode = d_mode; a.ref_cols = d_ref_cols; a.tgt_cols = d_tgt_cols;
    a.V = V; a.N = N; a.n_ref = n_ref; a.n_tgt = n_tgt; a.phased = is_phased ? 1 : 0;
    a.ref_listed = n_ref_listed; a.tgt_listed = n_tgt_listed;
    a.counts = reinterpret_cast<int64_t *>(ws);
    a.keep = ws + 256;
    a.block_base = reinterpret_cast<int32_t *>(ws + 256 + align_up((size_t)V, 256));
    a.ref_out = d_ref_out; a.tgt_out = d_tgt_out; a.pos_out = d_pos_out;
    cudaError_t e = launch(ingest_flag_kernel, (unsigned)nb, kIngestWarps * 32, 0, st, false, a);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "ingest_flag_kernel launch: %s", cudaGetErrorString(e));
    e = launch(ingest_scan_kernel, 1u, 1024u, 0, st, false, a, nb);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "ingest_scan_kernel launch: %s", cudaGetErrorString(e));
    e = launch(ingest_write_kernel, (unsigned)nb, kIngestWarps * 32, 0, st, false, a);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "ingest_write_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Source match rates
// ---------------------------------------------------------------------------------------------
extern "C" int sstar_b200_match_numerators(int device, void *stream, const int8_t *d_gt, int64_t V, int32_t N,
                                           const int32_t *d_tract_lo, const int32_t *d_tract_hi,
                                           const int32_t *d_tract_tgt, const int32_t *d_tract_hap, int32_t K,
                                           const int32_t *d_src_cols, int32_t S, int32_t ploidy, int64_t *d_numer) {
    g_launches = 0;
    if (V < 0 || V > (int64_t)INT32_MAX - 64 || N < 1) return fail(SSTAR_B200_EINVAL, "bad genotype array shape%s");
    if (K < 0 || S < 1 || ploidy < 1) return fail(SSTAR_B200_EINVAL, "bad tract / source / ploidy count%s");
    if (K == 0) return 0;
    if (!d_gt || !d_tract_lo || !d_tract_hi || !d_tract_tgt || !d_tract_hap || !d_src_cols || !d_numer)
        return fail(SSTAR_B200_EINVAL, "null pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    MatchArgs a = {};
    a.gt = d_gt; a.N = N; a.tract_lo = d_tract_lo; a.tract_hi = d_tract_hi; a.tract_tgt = d_tract_tgt;
    a.tract_hap = d_tract_hap; a.src_cols = d_src_cols; a.K = K; a.S = S; a.P = ploidy; a.numer = d_numer;
    const int64_t blocks = ((int64_t)K + 7) / 8;  // a warp per tract
    cudaError_t e = launch(match_kernel, (unsigned)blocks, 256u, 0, (cudaStream_t)stream, false, a);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "match_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}

// ---- S* SNP sets (optional extension; see snpset_kernels.cuh) --------------------------------
namespace {
struct SetLayout {
    size_t off_len, off_blocks, off_queue, off_qn, total;
    uint64_t arena_bytes;
    int64_t n, nb;
};
// The set kernels keep 17 bytes per list entry where the pairwise scorer keeps 13: their arena is
// the workspace's last block grown in place, then the per-unit lengths and the scan's block sums.
SetLayout make_set_layout(const WsLayout &L, int32_t W, int32_t T) {
    SetLayout S;
    S.arena_bytes = (uint64_t)L.arena_entries * kSetEntryBytes;
    S.n = (int64_t)W * T;
    S.nb = (S.n + kScanPerBlock - 1) / kScanPerBlock;
    size_t off = align_up(L.off_arena + (size_t)S.arena_bytes, 256);
    S.off_len = off;    off = align_up(off + (size_t)(S.n > 0 ? S.n : 1) * 4, 256);
    S.off_blocks = off; off = align_up(off + (size_t)(S.nb > 0 ? S.nb : 1) * 8, 256);
    S.off_queue = off;  off = align_up(off + (size_t)(S.n > 0 ? S.n : 1) * 4, 256);
    S.off_qn = off;     off += 256;
    S.total = off;
    return S;
}

int launch_snp_sets(cudaStream_t st, bool fill, const int8_t *d_tgt, const int32_t *d_pos, int32_t T, int32_t W,
                    int32_t bonus, int32_t max_mm, int32_t penalty, int64_t *d_score, int32_t *d_nsnp,
                    const int64_t *d_set_off, int32_t *d_set_pos, int64_t capacity, uint8_t *ws, const WsLayout &L,
                    const SetLayout &S, int32_t hint) {
    SnpSetArgs a = {};
    a.pos = d_pos; a.tgt = d_tgt;
    a.absent_bits = reinterpret_cast<const uint32_t *>(ws + L.off_absent);
    a.exotic = reinterpret_cast<const uint32_t *>(ws + L.off_exotic);
    a.carried = reinterpret_cast<const uint32_t *>(ws + L.off_carried);
    a.two = reinterpret_cast<const uint32_t *>(ws + L.off_two);
    a.win_lo = reinterpret_cast<const int32_t *>(ws + L.off_win_lo);
    a.win_hi = reinterpret_cast<const int32_t *>(ws + L.off_win_hi);
    a.win_pfirst = reinterpret_cast<const int32_t *>(ws + L.off_win_pfirst);
    a.win_plast = reinterpret_cast<const int32_t *>(ws + L.off_win_plast);
    a.win_ex = reinterpret_cast<uint32_t *>(ws + L.off_bail);
    a.queue = reinterpret_cast<int32_t *>(ws + S.off_queue);
    a.queue_n = reinterpret_cast<int32_t *>(ws + S.off_qn);
    a.T = T; a.Tp = L.Tp; a.W = W; a.bonus = bonus; a.max_mm = max_mm; a.penalty = penalty;
    a.score = d_score; a.nsnp = d_nsnp;
    a.set_len = reinterpret_cast<int32_t *>(ws + S.off_len);
    a.set_off = d_set_off; a.set_pos = d_set_pos; a.capacity = capacity;
    a.arena = ws + L.off_arena; a.arena_bytes = S.arena_bytes;
    // list entries per warp kept in shared memory: the largest window when the caller told us
    // (more resident warps to hide the gather's latency), else the pairwise kernel's 1024
    int32_t smem_cap = kPairSmemCap;
    if (hint > 0 && hint < kPairSmemCap) smem_cap = (int32_t)align_up((size_t)(hint < 64 ? 64 : hint), 8);
    a.smem_cap = smem_cap;
    const size_t smem = (size_t)kPairWarps * smem_cap * kSetEntryBytes;
    int per_sm = (int)(((size_t)200 << 10) / (smem + 1024));
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (S.n + kPairWarps - 1) / kPairWarps;
    if (grid > (int64_t)num_sms() * per_sm) grid = (int64_t)num_sms() * per_sm;
    cudaError_t e;
    if (!fill) {  // n, SNP counts and the NaN units from the planes; the rest is queued for the warps
        e = cudaMemsetAsync(a.queue_n, 0, 8, st);
        if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        int bd = (T + 31) / 32 * 32;
        if (bd > kCountThreads) bd = kCountThreads;
        const int64_t cgrid = (int64_t)W * ((T + bd - 1) / bd);
        if (cgrid > (int64_t)INT32_MAX) return fail(SSTAR_B200_EINVAL, "too many (window, column chunk) tiles%s");
        e = launch(snp_count_kernel, (unsigned)cgrid, (unsigned)bd, 0, st, false, a);
        if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "snp_count_kernel launch: %s", cudaGetErrorString(e));
    }
    if (fill) {  // the units the count kernel scored itself: their chains, a thread per unit
        int64_t cgrid = (S.n + kChainThreads - 1) / kChainThreads;
        if (cgrid > (int64_t)num_sms() * 16) cgrid = (int64_t)num_sms() * 16;
        e = launch(snp_chain_fill_kernel, (unsigned)cgrid, kChainThreads, 0, st, false, a);
        if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "snp_chain_fill_kernel launch: %s", cudaGetErrorString(e));
    }
    if (fill) {
        static SmemGrant grant;
        e = grant_smem(grant, snp_set_kernel<true>, smem);
        if (e == cudaSuccess) e = launch(snp_set_kernel<true>, (unsigned)grid, kPairWarps * 32, smem, st, false, a);
    } else {
        static SmemGrant grant;
        e = grant_smem(grant, snp_set_kernel<false>, smem);
        if (e == cudaSuccess) e = launch(snp_set_kernel<false>, (unsigned)grid, kPairWarps * 32, smem, st, false, a);
    }
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "snp_set_kernel launch: %s", cudaGetErrorString(e));
    return 0;
}
}  // namespace

extern "C" size_t sstar_b200_snp_set_workspace_bytes(int64_t V, int32_t R, int32_t T, int32_t W,
                                                     int32_t max_win_variants) {
    (void)R;
    if (V < 0 || T < 1 || W < 0) return 0;
    return make_set_layout(make_layout(V, T, W, max_win_variants < 0 ? 0 : max_win_variants), W, T).total;
}

extern "C" int sstar_b200_snp_set_offsets(int device, void *stream, const int8_t *d_ref_gt, const int8_t *d_tgt_gt,
                                          const int32_t *d_pos, int64_t V, int32_t R, int32_t T,
                                          const int32_t *d_win_start, const int32_t *d_win_end, int32_t W,
                                          int32_t match_bonus, int32_t max_mismatch, int32_t mismatch_penalty,
                                          int64_t *d_score, int32_t *d_nsnp, int64_t *d_set_off, void *d_workspace,
                                          size_t workspace_bytes, const sstar_b200_options *opts) {
    WsLayout L; int algo;
    int rc = check_common(device, V, R, T, W, d_workspace, workspace_bytes, opts, &L, &algo);
    if (rc) return rc;
    const SetLayout S = make_set_layout(L, W, T);
    if (workspace_bytes < S.total) return fail(SSTAR_B200_EWORKSPACE, "workspace too small for SNP sets%s");
    if (S.n > (int64_t)INT32_MAX) return fail(SSTAR_B200_EINVAL, "more than 2^31 - 1 units in one SNP-set call%s");
    if (!d_set_off) return fail(SSTAR_B200_EINVAL, "null offsets pointer%s");
    if (W > 0 && (!d_win_start || !d_win_end || !d_score || !d_nsnp))
        return fail(SSTAR_B200_EINVAL, "null window/output pointer%s");
    if (V > 0 && (!d_pos || !d_tgt_gt || !d_ref_gt)) return fail(SSTAR_B200_EINVAL, "null input pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = (uint8_t *)d_workspace;
    if (S.n == 0) {
        cudaError_t e = cudaMemsetAsync(d_set_off, 0, 8, st);
        if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
        return 0;
    }
    rc = launch_prep(st, 0, L.G, d_ref_gt, d_tgt_gt, V, R, T, d_pos, d_win_start, d_win_end, nullptr, 0, 0, W, ws, L);
    if (rc) return rc;
    rc = launch_snp_sets(st, false, d_tgt_gt, d_pos, T, W, match_bonus, max_mismatch, mismatch_penalty, d_score, d_nsnp,
                         nullptr, nullptr, 0, ws, L, S, opts ? opts->max_win_variants : 0);
    if (rc) return rc;
    const int32_t *len = reinterpret_cast<const int32_t *>(ws + S.off_len);
    int64_t *blocks = reinterpret_cast<int64_t *>(ws + S.off_blocks);
    cudaError_t e = launch(scan_sum_kernel, (unsigned)S.nb, kScanThreads, 0, st, false, len, S.n, blocks);
    if (e == cudaSuccess) e = launch(scan_blocks_kernel, 1u, 1024u, 0, st, false, blocks, S.nb, d_set_off + S.n);
    if (e == cudaSuccess)
        e = launch(scan_apply_kernel, (unsigned)S.nb, kScanThreads, 0, st, false, len, S.n, (const int64_t *)blocks, d_set_off);
    if (e != cudaSuccess) return fail(SSTAR_B200_ECUDA, "scan launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int sstar_b200_snp_set_fill(int device, void *stream, const int8_t *d_tgt_gt, const int32_t *d_pos,
                                       int64_t V, int32_t R, int32_t T, int32_t W, int32_t match_bonus,
                                       int32_t max_mismatch, int32_t mismatch_penalty, const int64_t *d_set_off,
                                       int32_t *d_set_pos, int64_t capacity, void *d_workspace,
                                       size_t workspace_bytes, const sstar_b200_options *opts) {
    WsLayout L; int algo;
    int rc = check_common(device, V, R, T, W, d_workspace, workspace_bytes, opts, &L, &algo);
    if (rc) return rc;
    const SetLayout S = make_set_layout(L, W, T);
    if (workspace_bytes < S.total) return fail(SSTAR_B200_EWORKSPACE, "workspace too small for SNP sets%s");
    if (capacity < 0) return fail(SSTAR_B200_EINVAL, "capacity must be >= 0%s");
    if (S.n == 0 || capacity == 0) return 0;
    if (!d_set_off || !d_set_pos || !d_pos || !d_tgt_gt) return fail(SSTAR_B200_EINVAL, "null pointer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return fail(SSTAR_B200_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(guard.err));
    return launch_snp_sets((cudaStream_t)stream, true, d_tgt_gt, d_pos, T, W, match_bonus, max_mismatch,
                           mismatch_penalty, nullptr, nullptr, d_set_off, d_set_pos, capacity,
                           (uint8_t *)d_workspace, L, S, opts ? opts->max_win_variants : 0);
}

#ifdef SSTAR_B200_FUSED_TIMING
// debug build only (tools/exp_fused_timing.py): the fused kernel's phase stamps, [4096][8] u64
extern "C" int sstar_b200_debug_fused_stamps(unsigned long long *h_out) {
    CUDA_TRY(cudaMemcpyFromSymbol(h_out, sstar::g_fused_stamps, sizeof(unsigned long long) * 4096 * 8));
    return 0;
}
#endif

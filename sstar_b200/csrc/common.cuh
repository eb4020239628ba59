// sstar_b200 -- shared definitions of the CUDA kernels and the host-side C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dp_core.cuh"

namespace sstar {

constexpr int kPackThreads = 256;
constexpr int kScoreMaxThreads = 128;
constexpr int kMaxGroup = 16;       // windows one score CTA walks at most
constexpr int kScoreWordsCap = 64;  // staged position tile of a window group: 64 words = 2048 variants
constexpr int kScoreListBytes = 22 << 10;  // shared memory for the per-column site lists of one CTA
constexpr int kPairWarps = 4;
constexpr int kPairSmemCap = 1024;  // pairwise list entries per warp kept in shared memory
constexpr int kNumSM = 148;  // sizing constant of the device-independent workspace layout only; launches use the device's SM count

constexpr int kMaxSlots = 32;       // independent score launches (window chunks) per call

// First 256 bytes of the workspace.
struct WsHeader {
    uint32_t status;
    int32_t pad[31];
    int32_t n_slow[kMaxSlots];  // per score launch: windows routed to the pairwise kernel
};

struct WsLayout {
    int64_t G;   // 32-variant groups
    int32_t Tp;  // plane row pitch in words (T rounded up to 32)
    size_t off_absent, off_exotic, off_carried, off_two, off_neg, off_win_lo, off_win_hi, off_win_pfirst, off_win_plast, off_slow, off_bail,
        off_arena;
    size_t arena_entries;
    size_t total;
};

struct BoundsArgs {
    const int32_t *pos;
    const int32_t *win_start, *win_end;  // [W] or null (then fixed_start / fixed_end)
    const int32_t *seg;                  // [W+1] search segment per window, or null (0..V)
    int32_t fixed_start, fixed_end;
    int32_t V, W;
    int32_t *win_lo, *win_hi, *slow_list;
    int32_t *win_pfirst, *win_plast;  // position of the first / last variant of each non-empty window
    uint32_t *win_bail;               // zeroed here: set once a window is routed to the pairwise kernel by a unit
    WsHeader *hdr;
    int force_slow;
    // In-place host entry (n_fetch > 0): pos / win_start / win_end above are device COPIES that
    // fetch_kernel -- the launch right before this one, n_fetch CTAs -- fills from mapped host memory (no
    // copy engine involved: small DMA copies are slow on some hosts and queue behind the SMs' own PCIe
    // reads).  The prep grid is a programmatic dependent launch: its pack CTAs start under the fetch,
    // its bounds CTAs wait for it (griddepcontrol.wait).
    const int32_t *src_pos, *src_ws, *src_we;
    int32_t n_fetch;
};
constexpr int kMaxFetchCtas = 32;

struct PackArgs {
    const int8_t *ref, *tgt;
    int32_t V, R, T, Tp;
    int32_t g_begin, g_end, gpc;  // 32-variant groups [g_begin, g_end) packed by this launch
    int32_t n_tiles;              // tiles of gpc groups; the grid's pack CTAs stride over them
    int32_t cchunks, ccols;       // wide panels (mode 2): a tile is one group x ccols target columns, cchunks per group
    int32_t tpitch;               // ... staged with this row pitch (ccols, + 16 when the segments keep their 16-byte phase)
    uint32_t ref_tile_bytes;  // smem bytes reserved per tile (incl. alignment slack)
    uint32_t tgt_tile_bytes;
    uint32_t *absent_bits, *exotic, *carried, *two;
    // sign plane: written (for every column of the group) only where exotic[g] != 0, see pack_tile
    uint32_t *neg;
};

struct ScoreArgs {
    const int32_t *pos;
    const int8_t *tgt;
    const uint32_t *absent_bits, *exotic, *carried, *two;
    const uint32_t *neg;  // sign plane, valid where exotic[g] != 0
    const int32_t *win_lo, *win_hi, *win_pfirst, *win_plast;
    int32_t *slow_list;
    uint32_t *win_bail;
    WsHeader *hdr;
    int32_t V, T, Tp, tchunks;
    int32_t w_begin, w_end;  // windows scored by this launch
    int32_t slot;            // its counter in WsHeader::n_slow; its slow-list segment starts at w_begin
    int32_t force_slow;      // every window of the launch goes to the pairwise kernel
    int32_t bonus, max_mm, penalty;
    int64_t *score;
    int32_t *nsnp;
    uint8_t *arena;  // pairwise only
    uint64_t arena_entries;
    int32_t pair_smem_cap;  // pairwise only: list entries per warp that fit the launch's shared memory
};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-
// serialization attribute may start while its predecessor drains; it must not touch the
// predecessor's outputs before pdl_wait().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

constexpr int kBoundsWindowsPerCta = kPackThreads / 32;

}  // namespace sstar

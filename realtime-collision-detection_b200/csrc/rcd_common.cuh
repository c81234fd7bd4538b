// Common device-side types and helpers for the collision core (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rcd.h"

namespace rcd {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 FULL_MASK = 0xffffffffu;

// ---- constants of the reference (src/collision/collision_detection.py:19-28,
//      src/collision/warning_system.py:18-27) -------------------------------------------------
constexpr double SAFE_DISTANCE_DEFAULT = 5.0;
constexpr double MAX_WARNING_TIME = 10.0;
constexpr double MAX_RELATIVE_SPEED = 50.0;
constexpr double W_DISTANCE = 0.3, W_TIME = 0.3, W_SPEED = 0.2, W_ANGLE = 0.1, W_TYPE = 0.1;
constexpr double RISK_LOW = 0.3, RISK_MEDIUM = 0.6, RISK_HIGH = 0.8;
constexpr int PREDICT_OFFSETS = 20;       // np.arange(0, 10, 0.5), :730
constexpr float PREDICT_RADIUS = 100.0f;  // :802
constexpr int PREDICT_STEPS = 10;         // int(1.0 / 0.1), :821 + :322

// ---- uniform grid ---------------------------------------------------------------------------
// Linear cell key, x fastest: key = (cz * ny + cy) * nx + cx.  Cells are clamped to the grid,
// which is a monotone, non-expanding map of cell coordinates: two objects within H of each
// other are always within floor(H / cell) + 1 cells after clamping ("grid").
struct GridParams {
    float ox, oy, oz;  // origin
    float cell;        // cell edge in x and y (any size: queries walk the cells under their bounding box)
    float inv_cell;
    float cell_z;      // cell edge in z (coarser: the worlds of the reference are ~100 m high)
    float inv_cell_z;
    int nx, ny, nz;
    u32 ncells;
};

__host__ __device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ int cell_coord(float x, float o, float inv_cell, int n) {
    float f = floorf((x - o) * inv_cell);
    // clamp in float first: huge / non-finite coordinates must not overflow the int conversion
    f = fminf(fmaxf(f, 0.0f), (float)(n - 1));
    return (int)f;
}
__device__ __forceinline__ u32 cell_key(const GridParams &g, float x, float y, float z) {
    int cx = cell_coord(x, g.ox, g.inv_cell, g.nx);
    int cy = cell_coord(y, g.oy, g.inv_cell, g.ny);
    int cz = cell_coord(z, g.oz, g.inv_cell_z, g.nz);
    return (u32)((cz * g.ny + cy) * g.nx + cx);
}

// ---- packed per-object state in cell order (3 planes of float4 = 48 B/object) ---------------
//   P0 = {x, y, z, size}   P1 = {vx, vy, vz, heading}   P2 = {ax, ay, az, meta}
// meta (bit pattern of P2.w): bits 0-7 type, bits 8-9 pattern / has-history, bit 16 owned.
constexpr u32 META_OWNED = 1u << 16;
__host__ __device__ __forceinline__ u32 meta_type(u32 m) { return m & 0xffu; }
__host__ __device__ __forceinline__ u32 meta_pattern(u32 m) { return (m >> 8) & 0x3u; }

// ---- frame counters (device) ------------------------------------------------------------------
struct Counters {
    unsigned long long n_candidates;
    unsigned long long n_potential;
    unsigned long long n_pairs;
    unsigned long long n_high_risk;
    unsigned long long n_alerts[4];
    unsigned long long n_exact;
    unsigned long long n_query_hits;  // rcd_query_radius
    // blocks of the pair queue handed out by k_pairs / length of the queue between k_narrow and k_exact
    // (reset before every step)
    unsigned long long n_qa_blocks, n_q3, n_q3p, n_q3u;
    unsigned long long n_overflow;  // parts of tiles k_pairs left to its overflow pass (pair queue full)
    unsigned long long n_items;     // work items of k_pairs (k_tile_plan)
    unsigned long long n_fallback;  // resolved entries the exact stage had to redo in full (expected 0)
    unsigned long long n_tests;     // (query, neighbour) tests of the S1 filter (k_pairs), padding of last chunks included
    unsigned long long n_detect_end;  // every record with predicted = 0 lies in front of this position of the pair buffer
};

// ---- warp / block helpers -----------------------------------------------------------------------
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_minf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// streaming loads / stores (read-once inputs, write-once outputs): keep them out of L1
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ u32 ld_stream(const u32 *p) { return __ldcs(p); }

}  // namespace rcd

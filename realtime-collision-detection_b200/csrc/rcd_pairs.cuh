// Pair enumeration + narrow phase + classification, one fused kernel per frame mode.
//
// A tile is TQ consecutive objects in cell order (one thread per querying object).  The tile
// finds the cell rows that can hold neighbours of any of its queries, flattens their contiguous
// spans and streams them through shared memory in double-buffered chunks (cp.async).  Work is
// split in two phases so that the expensive part runs with full warps:
//   filter : every query tests every staged neighbour with one broadcast LDS.128 and a squared
//            distance; survivors are compacted with __ballot_sync into a per-warp queue of
//            (query, neighbour) pairs in shared memory;
//   heavy  : whenever 32 pairs are queued, each lane takes one pair and runs the narrow phase
//            (temporal filter / 20 predicted offsets x 10 samples) in fp32 as a conservative
//            pre-filter; pairs that survive are decided in fp64 (rcd_exact.cuh), emitted through
//            one 64-bit atomic cursor and classified (alert priority) on the spot.
//
// Replaces the reference's per-vehicle Python loops:
//   detect : CollisionDetector.detect_collisions      src/collision/collision_detection.py:110-191
//   predict: CollisionPredictionModel.predict_collisions                                  :572-865
//   compute-node: SpatialIndex.query_nearby + CollisionDetector.detect_collisions
//                                                       src/compute/compute_node.py:98-119, 229-321
//   alert class: AlertManager.process_collision_risks/_get_priority
//                                                       src/collision/warning_system.py:259-311
#pragma once
#include "rcd_common.cuh"
#include "rcd_exact.cuh"

namespace rcd {

constexpr int TQ = 128;           // queries (threads) per tile
constexpr int NW = TQ / 32;       // warps per tile
constexpr int CH = 256;           // neighbours staged per chunk
constexpr int MAX_ROWS = TQ;      // cell rows handled per batch (one thread computes one row span)
constexpr int ROW_SCAN_MAX = 32;  // rows up to this many cells wide are scanned cell by cell
constexpr int QCAP = 64;          // per-warp pair queue (<= 31 carried + 32 pushed)

struct PairParams {
    u32 n;
    GridParams g;
    const float4 *P0, *P1, *P2;
    const u32 *keys;         // sorted cell keys
    const u32 *sorted_slot;  // cell order -> upload slot
    const u32 *in_id;        // upload slot -> caller id
    const u32 *cell_start, *cell_end;
    float R, T;              // search radius / time window (detect)
    int steps;               // int(T / 0.1)
    float pt, threshold;     // compute-node: prediction_time, risk_threshold
    int count_candidates;    // predict: also count the (i, j, m) radius hits (diagnostic, slower)
    rcd_pair *out;
    unsigned long long out_cap;
    Counters *counters;
    u32 *cand_count;         // per upload slot
};

// relative guard band of the fp32 radius test (fp32 error of d2 is < 1e-6 relative)
constexpr float BAND_R2 = 2.0e-5f;

// shared-memory state of one tile
struct StageBuf {
    float4 p0[CH], p1[CH], p2[CH];
    u32 pos[CH];  // position in cell order of the staged object
};
struct TileShared {
    StageBuf buf[2];
    float4 q0[TQ], q1[TQ], q2[TQ];  // the tile's own (querying) objects
    unsigned short queue[NW][QCAP];
    u32 cand[TQ];                   // candidates found in the heavy phase (per query)
    u32 row_lo[MAX_ROWS];
    u32 row_prefix[MAX_ROWS + 1];
    int red_i[NW][6];
    float red_f[NW];
    u32 scan[NW];
    u32 n_pot, n_exact;
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void emit_pair(const PairParams &P, u32 slot_i, u32 slot_j, double ttc, double dist,
                                          double rs, double risk, double mx, double my, double mz, double tcl,
                                          double dcl, int prio, int offset, bool predicted) {
    unsigned long long pos = atomicAdd(&P.counters->n_pairs, 1ULL);
    if (risk > 0.7) atomicAdd(&P.counters->n_high_risk, 1ULL);
    if (prio >= 0) atomicAdd(&P.counters->n_alerts[prio], 1ULL);
    if (pos < P.out_cap) {
        rcd_pair r;
        r.i = P.in_id[slot_i];
        r.j = P.in_id[slot_j];
        r.ttc = (float)ttc;
        r.distance = (float)dist;
        r.rel_speed = (float)rs;
        r.risk = (float)risk;
        r.cx = (float)mx; r.cy = (float)my; r.cz = (float)mz;
        r.t_closest = (float)tcl;
        r.d_closest = (float)dcl;
        r.priority = (int8_t)prio;
        r.offset = (uint8_t)offset;
        r.predicted = predicted ? 1 : 0;
        r.reserved = 0;
        P.out[pos] = r;
    }
}


// ---- cold paths: fp64 re-evaluation, kept out of line so the fp32 loops stay lean in registers ----
__device__ __noinline__ bool exact_within_radius(float ax, float ay, float az, float bx, float by, float bz, float R) {
    return within_radius_d(ax, ay, az, bx, by, bz, (double)R);
}

__device__ __noinline__ void exact_detect(const PairParams &P, TileShared &sh, const StageBuf &buf, u32 ql, u32 jj,
                                          u32 si, float T, int steps) {
    atomicAdd(&sh.n_exact, 1u);
    ObjD A = widen(sh.q0[ql], sh.q1[ql], sh.q2[ql]), B = widen(buf.p0[jj], buf.p1[jj], buf.p2[jj]);
    DetectResultD r = detect_pair_d(A, B, (double)T, steps);
    if (r.potential) atomicAdd(&sh.n_pot, 1u);
    if (r.hit)
        emit_pair(P, P.sorted_slot[si], P.sorted_slot[buf.pos[jj]], r.ttc, r.dist, r.rs, r.risk, r.mx, r.my, r.mz,
                  r.tc, r.cd, r.priority, 255, false);
}

struct PredictBest {
    double risk;
    double ttc, dist, rs, mx, my, mz;
    int m;
};

__device__ __noinline__ bool exact_predict_radius(TileShared &sh, const StageBuf &buf, u32 ql, u32 jj, u32 pattern,
                                                  int m) {
    atomicAdd(&sh.n_exact, 1u);
    ObjD A = widen(sh.q0[ql], sh.q1[ql], sh.q2[ql]);
    const float4 b0 = buf.p0[jj];
    double cx, cy, cz;
    predict_centre_d(A, pattern, 0.5 * (double)m, cx, cy, cz);
    return within_radius_d(cx, cy, cz, b0.x, b0.y, b0.z, (double)PREDICT_RADIUS);
}

__device__ __noinline__ void exact_predict(TileShared &sh, const StageBuf &buf, u32 ql, u32 jj, u32 pattern, int m,
                                           PredictBest *best) {
    atomicAdd(&sh.n_exact, 1u);
    ObjD A = widen(sh.q0[ql], sh.q1[ql], sh.q2[ql]), B = widen(buf.p0[jj], buf.p1[jj], buf.p2[jj]);
    PredictResultD r = predict_pair_d(A, B, pattern, m);
    if (r.hit && r.risk > best->risk) {  // strict >, offsets ascending (:862)
        best->risk = r.risk; best->ttc = r.ttc; best->dist = r.dist; best->rs = r.rs;
        best->mx = r.mx; best->my = r.my; best->mz = r.mz; best->m = m;
    }
}

__device__ __noinline__ void emit_predict(const PairParams &P, const StageBuf &buf, u32 si, u32 jj,
                                          const PredictBest *b) {
    emit_pair(P, P.sorted_slot[si], P.sorted_slot[buf.pos[jj]], b->ttc, b->dist, b->rs, b->risk, b->mx, b->my, b->mz,
              0.5 * (double)b->m, 0.0, priority_d(b->risk, b->ttc), b->m, true);
}

__device__ __noinline__ void exact_compute_node(const PairParams &P, TileShared &sh, const StageBuf &buf, u32 ql,
                                                u32 jj, u32 si) {
    atomicAdd(&sh.n_exact, 1u);
    ObjD A = widen(sh.q0[ql], sh.q1[ql], sh.q2[ql]), B = widen(buf.p0[jj], buf.p1[jj], buf.p2[jj]);
    ComputeNodeResultD r = compute_node_pair_d(A, B, (double)P.pt, (double)P.threshold);
    if (r.hit)
        emit_pair(P, P.sorted_slot[si], P.sorted_slot[buf.pos[jj]], r.ttc, r.fut, r.rs, r.risk, r.mx, r.my, r.mz, 0.0,
                  0.0, -1, 255, false);
}

// ---- detect: stages 1-4 for one queued pair ----------------------------------------------------
// a* = querying object, b* = neighbour, si / sj = their positions in cell order
__device__ __forceinline__ void heavy_detect(const PairParams &P, TileShared &sh, const StageBuf &buf, u32 ql, u32 jj,
                                             const float4 &a0, const float4 &a1, const float4 &a2, const float4 &b0,
                                             const float4 &b1, const float4 &b2, u32 si, float R, float T,
                                             int steps) {
    const float R2 = R * R;
    float dx = b0.x - a0.x, dy = b0.y - a0.y, dz = b0.z - a0.z;  // rel_position = other - self
    float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 >= R2 * (1.0f - BAND_R2)) {  // the filter could not decide the radius test (spatial_index.py:268)
        atomicAdd(&sh.n_exact, 1u);
        if (!exact_within_radius(a0.x, a0.y, a0.z, b0.x, b0.y, b0.z, R)) return;
        atomicAdd(&sh.cand[ql], 1u);
    }
    float rvx = a1.x - b1.x, rvy = a1.y - b1.y, rvz = a1.z - b1.z;  // rel_velocity = self - other
    float rs2 = rvx * rvx + rvy * rvy + rvz * rvz;
    if (rs2 < 0.0099f) return;  // rel_speed < 0.1 with margin (0.1^2 = 0.01)
    float dot = dx * rvx + dy * rvy + dz * rvz;
    float edot = 4.0e-6f * sqrtf(d2 * rs2) + 1.0e-20f;
    // dot > 0: either (dot > 0 and cur > 5) or time_to_closest < 0 rejects the pair (:273, :280)
    if (dot > edot) return;
    if (-dot > T * rs2 * (1.0f + 1.0e-5f) + edot) return;  // time_to_closest > time_window
    float tc = fmaxf(-dot, 0.0f) / rs2;
    float rax = a2.x - b2.x, ray = a2.y - b2.y, raz = a2.z - b2.z;
    float h = 0.5f * tc * tc;
    float ex = rvx * tc + rax * h - dx, ey = rvy * tc + ray * h - dy, ez = rvz * tc + raz * h - dz;
    float cd2 = ex * ex + ey * ey + ez * ez;
    float safe = (a0.w + b0.w) * 0.5f + 5.0f;
    float tcerr = edot / rs2 + 4.0e-6f * tc;
    float band = 2.0e-3f + 2.0f * (sqrtf(rs2) + sqrtf(rax * rax + ray * ray + raz * raz) * tc) * tcerr;
    float thr = safe + band;
    if (cd2 > thr * thr) return;
    // ---- survivor: decide everything in fp64, in the reference's operation order -------------
    exact_detect(P, sh, buf, ql, jj, si, T, steps);
}

// ---- predict: 20 offsets x (radius test, <= 10 samples), max-risk merge ------------------------
template <bool COUNT_CAND>
__device__ __forceinline__ void heavy_predict(const PairParams &P, TileShared &sh, const StageBuf &buf, u32 ql, u32 jj,
                                              const float4 &a0, const float4 &a1, const float4 &a2, const float4 &b0,
                                              const float4 &b1, const float4 &b2, u32 si, u32 pattern) {
    const float R2 = PREDICT_RADIUS * PREDICT_RADIUS;
    const float fv = (pattern >= RCD_PAT_CONSTANT_VELOCITY) ? 1.0f : 0.0f;
    const float fa = (pattern == RCD_PAT_ACCELERATING) ? 1.0f : 0.0f;
    // centre_i(t) = p_i + uv t + ua t^2/2 (pattern: stationary / constant_velocity / accelerating, :728-761)
    const float uvx = a1.x * fv, uvy = a1.y * fv, uvz = a1.z * fv;
    const float uax = a2.x * fa, uay = a2.y * fa, uaz = a2.z * fa;
    float dx = b0.x - a0.x, dy = b0.y - a0.y, dz = b0.z - a0.z;
    float d2 = dx * dx + dy * dy + dz * dz;
    // g(t) = centre_i(t) - predicted_j(t) = -d + cv t + ca t^2/2   (:814)
    float cvx = uvx - b1.x, cvy = uvy - b1.y, cvz = uvz - b1.z;
    float cax = uax - b2.x, cay = uay - b2.y, caz = uaz - b2.z;
    // the 10 samples advance both vehicles with their own v, a (:326-327, quirk Q6)
    float rvx = a1.x - b1.x, rvy = a1.y - b1.y, rvz = a1.z - b1.z;
    float rax = a2.x - b2.x, ray = a2.y - b2.y, raz = a2.z - b2.z;
    float rvn = sqrtf(rvx * rvx + rvy * rvy + rvz * rvz);
    float ran = sqrtf(rax * rax + ray * ray + raz * raz);
    float safe = (a0.w + b0.w) * 0.5f + 5.0f;
    float safe_b = safe + 2.0e-3f + 1.0e-6f * sqrtf(d2);
    float safe_b2 = safe_b * safe_b;
    // the samples move the pair by at most |rv|*0.9 + |ra|*0.405 from the offset state
    float hr = safe_b + rvn * 0.9f + ran * 0.405f;
    float hr2 = hr * hr;
    if (!COUNT_CAND) {
        // whole-pair rejection: |g(t)| >= |-d + cv t| - |ca| t^2/2 on [0, 9.5]
        float cv2 = cvx * cvx + cvy * cvy + cvz * cvz;
        float ts = (cv2 > 1.0e-12f) ? fminf(fmaxf((dx * cvx + dy * cvy + dz * cvz) / cv2, 0.0f), 9.5f) : 0.0f;
        float lx = cvx * ts - dx, ly = cvy * ts - dy, lz = cvz * ts - dz;
        float lim = hr + sqrtf(cax * cax + cay * cay + caz * caz) * 45.125f + 1.0e-3f * sqrtf(d2) + 1.0e-2f;
        if (lx * lx + ly * ly + lz * lz > lim * lim) return;
    }
    PredictBest best;
    best.risk = -1.0;
    best.m = -1;
    u32 ncand = 0;
#pragma unroll 1
    for (int m = 0; m < PREDICT_OFFSETS; ++m) {
        float t = 0.5f * (float)m;
        float h = 0.5f * t * t;
        float gx = cvx * t + cax * h - dx, gy = cvy * t + cay * h - dy, gz = cvz * t + caz * h - dz;
        float g2 = gx * gx + gy * gy + gz * gz;
        if (!COUNT_CAND && g2 > hr2) continue;
        // e = centre_i(t) - p_j: others are looked up at their CURRENT positions (:801-803)
        float ex = uvx * t + uax * h - dx, ey = uvy * t + uay * h - dy, ez = uvz * t + uaz * h - dz;
        float c2 = ex * ex + ey * ey + ez * ez;
        if (c2 > R2 * (1.0f + BAND_R2)) continue;
        if (c2 >= R2 * (1.0f - BAND_R2)) {
            if (!exact_predict_radius(sh, buf, ql, jj, pattern, m)) continue;
        }
        ++ncand;
        if (COUNT_CAND && g2 > hr2) continue;
        bool maybe = false;
#pragma unroll
        for (int k = 0; k < PREDICT_STEPS; ++k) {
            float tau = 0.1f * (float)k;
            float hh = 0.5f * tau * tau;
            float rx = gx + rvx * tau + rax * hh, ry = gy + rvy * tau + ray * hh, rz = gz + rvz * tau + raz * hh;
            maybe |= (rx * rx + ry * ry + rz * rz <= safe_b2);
        }
        if (!maybe) continue;
        exact_predict(sh, buf, ql, jj, pattern, m, &best);
    }
    if (COUNT_CAND && ncand) atomicAdd(&sh.cand[ql], ncand);
    if (best.m >= 0) emit_predict(P, buf, si, jj, &best);
}

// ---- compute-node pair function ------------------------------------------------------------------
__device__ __forceinline__ void heavy_compute_node(const PairParams &P, TileShared &sh, const StageBuf &buf, u32 ql,
                                                   u32 jj, const float4 &a0, const float4 &a1, const float4 &a2,
                                                   const float4 &b0, const float4 &b1, const float4 &b2, u32 si,
                                                   u32 sj) {
    const float R2 = P.R * P.R;
    float dx = b0.x - a0.x, dy = b0.y - a0.y, dz = b0.z - a0.z;
    float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 >= R2 * (1.0f - BAND_R2)) {
        atomicAdd(&sh.n_exact, 1u);
        if (!exact_within_radius(a0.x, a0.y, a0.z, b0.x, b0.y, b0.z, P.R)) return;
        atomicAdd(&sh.cand[ql], 1u);  // query_nearby returns the querying vehicle too (quirk Q8)
    }
    if (sj == si) return;                          // compute_node.py:251-252
    if (d2 > 2500.0f * (1.0f + BAND_R2)) return;   // current_distance > 50
    if (meta_pattern(__float_as_uint(a2.w)) == 0u || meta_pattern(__float_as_uint(b2.w)) == 0u)
        return;                                    // predict_position needs >= 2 samples (:202-203)
    float rvx = a1.x - b1.x, rvy = a1.y - b1.y, rvz = a1.z - b1.z;
    float fx = dx - rvx * P.pt, fy = dy - rvy * P.pt, fz = dz - rvz * P.pt;  // future_j - future_i
    float fut2 = fx * fx + fy * fy + fz * fz;
    float rs2 = rvx * rvx + rvy * rvy + rvz * rvz;
    if (fut2 > d2 * (1.0f + 1.0e-4f) + 1.0e-6f && d2 > 16.0f * (1.0f + 1.0e-4f)) return;  // moving apart
    float fut = fmaxf(sqrtf(fut2), 0.1f);
    if (0.4f * sqrtf(rs2) < P.threshold * fut * (1.0f - 1.0e-4f)) return;  // risk < threshold
    exact_compute_node(P, sh, buf, ql, jj, si);
}

// lower bound in the sorted key array
__device__ __forceinline__ u32 lower_bound_keys(const u32 *__restrict__ keys, u32 n, u32 key) {
    u32 lo = 0, hi = n;
    while (lo < hi) {
        u32 mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int MODE, bool COUNT_CAND>
__device__ __forceinline__ void heavy_dispatch(const PairParams &P, TileShared &sh, const StageBuf &buf, u32 entry,
                                               u32 tile_base) {
    const u32 ql = entry >> 8, jj = entry & 0xffu;
    const float4 a0 = sh.q0[ql], a1 = sh.q1[ql], a2 = sh.q2[ql];
    const float4 b0 = buf.p0[jj], b1 = buf.p1[jj], b2 = buf.p2[jj];
    const u32 si = tile_base + ql, sj = buf.pos[jj];
    if (MODE == RCD_MODE_DETECT) {
        heavy_detect(P, sh, buf, ql, jj, a0, a1, a2, b0, b1, b2, si, P.R, P.T, P.steps);
    } else if (MODE == RCD_MODE_PREDICT) {
        const u32 pattern = meta_pattern(__float_as_uint(a2.w));
        if (pattern == RCD_PAT_NO_HISTORY)  // history < 2 -> detect_collisions(id) with defaults (:590-592)
            heavy_detect(P, sh, buf, ql, jj, a0, a1, a2, b0, b1, b2, si, PREDICT_RADIUS, 10.0f, 100);
        else
            heavy_predict<COUNT_CAND>(P, sh, buf, ql, jj, a0, a1, a2, b0, b1, b2, si, pattern);
    } else {
        heavy_compute_node(P, sh, buf, ql, jj, a0, a1, a2, b0, b1, b2, si, sj);
    }
}

template <int MODE, bool COUNT_CAND>
__global__ void __launch_bounds__(TQ, 4) k_pairs(PairParams P) {
    __shared__ TileShared sh;

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 tile_base = blockIdx.x * TQ;
    const u32 s = tile_base + tid;
    const bool valid = s < P.n;
    const GridParams g = P.g;

    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0, p2 = p0;
    int cx = 0, cy = 0, cz = 0;
    u32 meta = 0;
    if (valid) {
        p0 = P.P0[s];
        p1 = P.P1[s];
        p2 = P.P2[s];
        meta = __float_as_uint(p2.w);
        u32 key = P.keys[s];
        cx = (int)(key % (u32)g.nx);
        u32 row = key / (u32)g.nx;
        cy = (int)(row % (u32)g.ny);
        cz = (int)(row / (u32)g.ny);
    }
    sh.q0[tid] = p0;
    sh.q1[tid] = p1;
    sh.q2[tid] = p2;
    sh.cand[tid] = 0;
    if (tid == 0) { sh.n_pot = 0; sh.n_exact = 0; }
    const bool owned = valid && (meta & META_OWNED);
    const u32 pattern = meta_pattern(meta);
    // queries that take the radius-R test in the filter (detect-like); the others are predict queries
    const bool radius_query = (MODE != RCD_MODE_PREDICT) || pattern == RCD_PAT_NO_HISTORY;
    const float Rq = (MODE == RCD_MODE_PREDICT) ? PREDICT_RADIUS : P.R;
    const float R2_hi = Rq * Rq * (1.0f + BAND_R2), R2_lo = Rq * Rq * (1.0f - BAND_R2);

    // reach of this query: how far a neighbour can be and still matter
    float reach;
    if (!radius_query) {
        float fv = (pattern >= RCD_PAT_CONSTANT_VELOCITY) ? 1.0f : 0.0f;
        float fa = (pattern == RCD_PAT_ACCELERATING) ? 1.0f : 0.0f;
        float travel = sqrtf(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z) * fv * 9.5f +
                       sqrtf(p2.x * p2.x + p2.y * p2.y + p2.z * p2.z) * fa * 45.125f;
        reach = (PREDICT_RADIUS + travel) * (1.0f + 1.0e-5f) + 1.0e-2f;
    } else {
        reach = Rq * (1.0f + 1.0e-5f) + 1.0e-3f;
    }
    if (!(reach < 1.0e30f)) reach = 1.0e30f;  // non-finite velocity: scan everything
    // filter threshold on the squared distance
    const float pass2 = radius_query ? R2_hi : reach * reach;
    u32 ncand = 0;  // candidates decided by the filter itself

    // a tile that crosses a cell-row boundary is processed as two groups (first row / the rest)
    // so that each group's cell box stays tight
    const u32 first_row = P.keys[tile_base] / (u32)g.nx;
    const int my_group = (valid && ((u32)(cy + cz * g.ny) != first_row)) ? 1 : 0;
    const int ngroups = __syncthreads_or(my_group) ? 2 : 1;  // also publishes sh.q*, sh.cand

    for (int grp = 0; grp < ngroups; ++grp) {
        const bool active = owned && my_group == grp;
        // ---- block reduction: cell box and maximum reach of the active queries ----------------
        int r0 = active ? cx : 0x7fffffff, r1 = active ? cx : -1;
        int r2 = active ? cy : 0x7fffffff, r3 = active ? cy : -1;
        int r4 = active ? cz : 0x7fffffff, r5 = active ? cz : -1;
        float rh = active ? reach : 0.0f;
        r0 = warp_min(r0); r1 = warp_max(r1); r2 = warp_min(r2); r3 = warp_max(r3);
        r4 = warp_min(r4); r5 = warp_max(r5); rh = warp_maxf(rh);
        __syncthreads();  // previous group's readers of red_* are done
        if (lane == 0) {
            sh.red_i[warp][0] = r0; sh.red_i[warp][1] = r1; sh.red_i[warp][2] = r2;
            sh.red_i[warp][3] = r3; sh.red_i[warp][4] = r4; sh.red_i[warp][5] = r5;
            sh.red_f[warp] = rh;
        }
        __syncthreads();
        int cxmin = 0x7fffffff, cxmax = -1, cymin = 0x7fffffff, cymax = -1, czmin = 0x7fffffff, czmax = -1;
        float hmax = 0.0f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            cxmin = min(cxmin, sh.red_i[w][0]); cxmax = max(cxmax, sh.red_i[w][1]);
            cymin = min(cymin, sh.red_i[w][2]); cymax = max(cymax, sh.red_i[w][3]);
            czmin = min(czmin, sh.red_i[w][4]); czmax = max(czmax, sh.red_i[w][5]);
            hmax = fmaxf(hmax, sh.red_f[w]);
        }
        if (cxmax < 0) continue;  // no active query in this group (uniform across the block)
        // |ci - cj| <= floor(H / cell) + 1 for clamped floor() cells (the reference's own bound,
        // spatial_index.py:246-248)
        float srf = floorf(hmax * g.inv_cell + 1.0e-3f) + 1.0f;
        int sr = (srf < 1.0e6f) ? (int)srf : 1000000;
        const int x0 = max(cxmin - sr, 0), x1 = min(cxmax + sr, g.nx - 1);
        const int y0 = max(cymin - sr, 0), y1 = min(cymax + sr, g.ny - 1);
        const int z0 = max(czmin - sr, 0), z1 = min(czmax + sr, g.nz - 1);
        const int ny_span = y1 - y0 + 1;
        const int nrows = ny_span * (z1 - z0 + 1);

        for (int rbase = 0; rbase < nrows; rbase += MAX_ROWS) {
            // ---- span of one cell row per thread ---------------------------------------------
            u32 lo = 0, cnt = 0;
            if ((int)tid + rbase < nrows) {
                int rr = rbase + (int)tid;
                int yy = y0 + rr % ny_span, zz = z0 + rr / ny_span;
                u32 c0 = (u32)((zz * g.ny + yy) * g.nx + x0), c1 = c0 + (u32)(x1 - x0);
                u32 first = 0xffffffffu, last = 0;
                if (x1 - x0 < ROW_SCAN_MAX) {
                    for (u32 c = c0; c <= c1; ++c) {
                        u32 st = P.cell_start[c], en = P.cell_end[c];
                        if (en > st) { first = min(first, st); last = max(last, en); }
                    }
                } else {
                    first = lower_bound_keys(P.keys, P.n, c0);
                    last = lower_bound_keys(P.keys, P.n, c1 + 1);
                }
                if (last > first) { lo = first; cnt = last - first; }
            }
            // exclusive scan of cnt over the block
            u32 incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 t = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= (u32)o) incl += t;
            }
            __syncthreads();  // previous batch's readers of row_* / scan are done
            if (lane == 31) sh.scan[warp] = incl;
            __syncthreads();
            u32 wbase = 0;
            for (u32 w = 0; w < warp; ++w) wbase += sh.scan[w];
            sh.row_lo[tid] = lo;
            sh.row_prefix[tid] = wbase + incl - cnt;
            if (tid == TQ - 1) sh.row_prefix[MAX_ROWS] = wbase + incl;
            __syncthreads();
            const u32 total = sh.row_prefix[MAX_ROWS];
            const u32 nchunks = (total + CH - 1) / CH;

            // stage chunk c of the flattened spans into buffer c & 1 (cp.async, 2 objects per thread)
            auto stage = [&](u32 c) {
                StageBuf &b = sh.buf[c & 1u];
#pragma unroll
                for (int e = 0; e < CH / TQ; ++e) {
                    u32 slot = tid + e * TQ;
                    u32 f = c * CH + slot;
                    if (f < total) {
                        int a = 0, z = MAX_ROWS - 1;  // last row r with prefix[r] <= f
                        while (a < z) {
                            int mid = (a + z + 1) >> 1;
                            if (sh.row_prefix[mid] <= f) a = mid; else z = mid - 1;
                        }
                        u32 src = sh.row_lo[a] + (f - sh.row_prefix[a]);
                        cp_async16(&b.p0[slot], P.P0 + src);
                        cp_async16(&b.p1[slot], P.P1 + src);
                        cp_async16(&b.p2[slot], P.P2 + src);
                        b.pos[slot] = src;
                    }
                }
                cp_async_commit();
            };

            if (nchunks) stage(0);
            for (u32 c = 0; c < nchunks; ++c) {
                if (c + 1 < nchunks) {
                    stage(c + 1);
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                const StageBuf &b = sh.buf[c & 1u];
                const u32 m = min((u32)CH, total - c * CH);
                // ---- filter: one query per thread against every staged neighbour ------------------
                u32 qcount = 0;  // warp-uniform
                for (u32 jj = 0; jj < m; ++jj) {
                    const float4 b0 = b.p0[jj];
                    float dx = b0.x - p0.x, dy = b0.y - p0.y, dz = b0.z - p0.z;
                    float d2 = dx * dx + dy * dy + dz * dz;
                    bool pass = active && d2 <= pass2;
                    if (MODE != RCD_MODE_COMPUTE_NODE) pass = pass && (b.pos[jj] != s);  // strip self (:224-225)
                    if (pass && radius_query && d2 < R2_lo) ++ncand;  // certainly within the radius
                    const u32 mask = __ballot_sync(FULL_MASK, pass);
                    if (mask) {
                        if (pass) sh.queue[warp][qcount + __popc(mask & lanemask_lt())] = (unsigned short)((tid << 8) | jj);
                        qcount += __popc(mask);
                        __syncwarp();
                        if (qcount >= 32) {
                            heavy_dispatch<MODE, COUNT_CAND>(P, sh, b, sh.queue[warp][qcount - 32 + lane], tile_base);
                            qcount -= 32;
                            __syncwarp();
                        }
                    }
                }
                if (qcount) {  // drain the tail before the buffer is recycled
                    if (lane < qcount) heavy_dispatch<MODE, COUNT_CAND>(P, sh, b, sh.queue[warp][lane], tile_base);
                    __syncwarp();
                }
                __syncthreads();
            }
        }
    }

    // ---- per-object candidate count + frame totals ------------------------------------------------
    __syncthreads();
    ncand += sh.cand[tid];
    if (owned && P.cand_count) P.cand_count[P.sorted_slot[s]] = ncand;
    unsigned long long c = warp_sum((unsigned long long)ncand);
    if (lane == 0 && c) atomicAdd(&P.counters->n_candidates, c);
    if (tid == 0) {
        if (sh.n_pot) atomicAdd(&P.counters->n_potential, (unsigned long long)sh.n_pot);
        if (sh.n_exact) atomicAdd(&P.counters->n_exact, (unsigned long long)sh.n_exact);
    }
}

// -------------------------------------------------------------------------------------------------
// Radius query for explicit points: SpatialIndex.get_nearby_vehicles (spatial_index.py:229-271)
// and compute_node.SpatialIndex.query_nearby (compute_node.py:98-119).  One warp per query;
// hits are appended as (query, upload slot) pairs, grouped and ordered on the host.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_query_radius(u32 nq, const float *__restrict__ qx, const float *__restrict__ qy, const float *__restrict__ qz,
               float radius, GridParams g, u32 n, const float4 *__restrict__ P0, const u32 *__restrict__ keys,
               const u32 *__restrict__ sorted_slot, uint2 *__restrict__ hits, unsigned long long cap,
               Counters *counters) {
    const u32 lane = threadIdx.x & 31u;
    const u32 qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (qi >= nq) return;
    const float x = qx[qi], y = qy[qi], z = qz[qi];
    const float R2 = radius * radius;
    int cx = cell_coord(x, g.ox, g.inv_cell, g.nx), cy = cell_coord(y, g.oy, g.inv_cell, g.ny),
        cz = cell_coord(z, g.oz, g.inv_cell, g.nz);
    // a query point may lie outside the grid: widen the stencil by its distance to the box
    float ex = fmaxf(fmaxf(g.ox - x, x - (g.ox + g.nx * g.cell)), 0.0f);
    float ey = fmaxf(fmaxf(g.oy - y, y - (g.oy + g.ny * g.cell)), 0.0f);
    float ez = fmaxf(fmaxf(g.oz - z, z - (g.oz + g.nz * g.cell)), 0.0f);
    (void)ex; (void)ey; (void)ez;  // clamping is non-expanding: the plain stencil already suffices
    float srf = floorf(radius * g.inv_cell + 1.0e-3f) + 1.0f;
    int sr = (srf < 1.0e6f) ? (int)srf : 1000000;
    const int x0 = max(cx - sr, 0), x1 = min(cx + sr, g.nx - 1);
    const int y0 = max(cy - sr, 0), y1 = min(cy + sr, g.ny - 1);
    const int z0 = max(cz - sr, 0), z1 = min(cz + sr, g.nz - 1);
    for (int zz = z0; zz <= z1; ++zz)
        for (int yy = y0; yy <= y1; ++yy) {
            u32 c0 = (u32)((zz * g.ny + yy) * g.nx + x0), c1 = c0 + (u32)(x1 - x0);
            u32 first = lower_bound_keys(keys, n, c0), last = lower_bound_keys(keys, n, c1 + 1);
            for (u32 s = first + lane; s < last; s += 32) {
                float4 b = P0[s];
                float dx = b.x - x, dy = b.y - y, dz = b.z - z;
                float d2 = dx * dx + dy * dy + dz * dz;
                bool in = d2 <= R2 * (1.0f - BAND_R2);
                if (!in && d2 <= R2 * (1.0f + BAND_R2)) in = within_radius_d(x, y, z, b.x, b.y, b.z, (double)radius);
                if (in) {
                    unsigned long long pos = atomicAdd(&counters->n_query_hits, 1ULL);
                    if (pos < cap) hits[pos] = make_uint2(qi, sorted_slot[s]);
                }
            }
        }
}

// -------------------------------------------------------------------------------------------------
// Trajectory pattern classifier (collision_detection.py:623-711), one thread per object, fp64,
// streaming over the (already time-ordered) samples in the reference's summation order.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_classify_patterns(u32 n, u32 stride, const double *__restrict__ samples /* [n][stride][4] */,
                    const u32 *__restrict__ count, uint8_t *__restrict__ out) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 cnt = min(count[i], stride);
    if (cnt < 2) { out[i] = RCD_PAT_NO_HISTORY; return; }
    const double *h = samples + (size_t)i * stride * 4;
    double svx = 0, svy = 0, svz = 0, sax = 0, say = 0, saz = 0;
    u32 nv = 0, na = 0;
    double pvx = 0, pvy = 0, pvz = 0, pvt = 0;
    double lx = h[0], ly = h[1], lz = h[2], lt = h[3];
    for (u32 k = 1; k < cnt; ++k) {
        double x = h[4 * k], y = h[4 * k + 1], z = h[4 * k + 2], t = h[4 * k + 3];
        double dt = dsub(t, lt);
        if (dt > 0) {
            double vx = __ddiv_rn(dsub(x, lx), dt), vy = __ddiv_rn(dsub(y, ly), dt), vz = __ddiv_rn(dsub(z, lz), dt);
            if (nv > 0) {
                double dtv = dsub(t, pvt);
                if (dtv > 0) {
                    sax = dadd(sax, __ddiv_rn(dsub(vx, pvx), dtv));
                    say = dadd(say, __ddiv_rn(dsub(vy, pvy), dtv));
                    saz = dadd(saz, __ddiv_rn(dsub(vz, pvz), dtv));
                    ++na;
                }
            }
            svx = dadd(svx, vx); svy = dadd(svy, vy); svz = dadd(svz, vz);
            ++nv;
            pvx = vx; pvy = vy; pvz = vz; pvt = t;
        }
        lx = x; ly = y; lz = z; lt = t;
    }
    if (nv == 0) { out[i] = RCD_PAT_STATIONARY; return; }
    svx = __ddiv_rn(svx, (double)nv); svy = __ddiv_rn(svy, (double)nv); svz = __ddiv_rn(svz, (double)nv);
    if (na) { sax = __ddiv_rn(sax, (double)na); say = __ddiv_rn(say, (double)na); saz = __ddiv_rn(saz, (double)na); }
    double speed = mag3_d(svx, svy, svz), accel = mag3_d(sax, say, saz);
    out[i] = speed < 0.1 ? RCD_PAT_STATIONARY : (accel < 0.1 ? RCD_PAT_CONSTANT_VELOCITY : RCD_PAT_ACCELERATING);
}

// -------------------------------------------------------------------------------------------------
// Spatial slabs: select / pack / append halo objects.  Record = 13 x u32:
// 11 floats (px..heading), meta (type | pattern << 8), id.
// -------------------------------------------------------------------------------------------------
constexpr int HALO_WORDS = 13;
constexpr int MAX_PEERS = 64;

struct SlabParams {
    int n_peers, self;
    float lo[MAX_PEERS], hi[MAX_PEERS];
    float halo;
};

// phase 0: count per peer; phase 1: pack at base[p] + cursor[p]++
__global__ void __launch_bounds__(256)
k_halo_pack(u32 n_owned, InputState in, SlabParams sp, int phase, unsigned long long *__restrict__ counts,
            const unsigned long long *__restrict__ base, u32 *__restrict__ out, unsigned long long cap) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned) return;
    float x = in.px[i];
    for (int p = 0; p < sp.n_peers; ++p) {
        if (p == sp.self) continue;
        if (x >= sp.lo[p] - sp.halo && x < sp.hi[p] + sp.halo) {
            unsigned long long k = atomicAdd(&counts[p], 1ULL);
            if (phase == 1) {
                unsigned long long pos = base[p] + k;
                if (pos < cap) {
                    u32 *r = out + pos * HALO_WORDS;
                    r[0] = __float_as_uint(x);
                    r[1] = __float_as_uint(in.py[i]);
                    r[2] = __float_as_uint(in.pz[i]);
                    r[3] = __float_as_uint(in.vx[i]);
                    r[4] = __float_as_uint(in.vy[i]);
                    r[5] = __float_as_uint(in.vz[i]);
                    r[6] = __float_as_uint(in.ax[i]);
                    r[7] = __float_as_uint(in.ay[i]);
                    r[8] = __float_as_uint(in.az[i]);
                    r[9] = __float_as_uint(in.size[i]);
                    r[10] = __float_as_uint(in.heading[i]);
                    r[11] = (u32)in.type[i] | ((u32)in.pattern[i] << 8);
                    r[12] = in.id ? in.id[i] : i;
                }
            }
        }
    }
}

struct MutableState {
    float *f[11];
    uint8_t *type, *pattern;
    u32 *id;
};

__global__ void __launch_bounds__(256)
k_halo_append(const u32 *__restrict__ rec, u32 n_rec, u32 dst_base, MutableState st) {
    u32 r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    const u32 *w = rec + (size_t)r * HALO_WORDS;
    u32 d = dst_base + r;
#pragma unroll
    for (int k = 0; k < 11; ++k) st.f[k][d] = __uint_as_float(w[k]);
    st.type[d] = (uint8_t)(w[11] & 0xffu);
    st.pattern[d] = (uint8_t)((w[11] >> 8) & 0xffu);
    st.id[d] = w[12];
}

__global__ void __launch_bounds__(256) k_iota(u32 *p, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

}  // namespace rcd

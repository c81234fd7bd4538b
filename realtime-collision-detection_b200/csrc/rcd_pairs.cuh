// Pair enumeration + narrow phase + classification, one fused kernel per frame mode.
//
// A tile is TQ consecutive objects in cell order (one thread per querying object).  The tile
// finds the cell rows that can hold neighbours of any of its queries, flattens their contiguous
// spans, and streams them through shared memory in chunks of CH objects; every query tests
// every staged neighbour (broadcast LDS.128, no divergence in the filter).  Survivors of the
// fp32 pre-filter are re-evaluated in fp64 (rcd_exact.cuh), emitted with a single 64-bit atomic
// cursor, and classified (alert priority) on the spot.
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

constexpr int TQ = 128;        // queries (threads) per tile
constexpr int CH = 128;        // neighbours staged per chunk
constexpr int MAX_ROWS = TQ;   // cell rows handled per batch (one thread computes one row span)
constexpr int ROW_SCAN_MAX = 32;  // rows up to this many cells wide are scanned cell by cell

struct PairParams {
    u32 n;
    GridParams g;
    const float4 *P0, *P1, *P2;
    const u32 *keys;         // sorted cell keys
    const u32 *sorted_slot;  // cell order -> upload slot
    const u32 *in_id;        // upload slot -> caller id (null: identity)
    const u32 *cell_start, *cell_end;
    float R, T;              // search radius / time window (detect)
    int steps;               // int(T / 0.1)
    float pt, threshold;     // compute-node: prediction_time, risk_threshold
    rcd_pair *out;
    unsigned long long out_cap;
    Counters *counters;
    u32 *cand_count;         // per upload slot
};

// relative guard band of the fp32 radius test (fp32 error of d2 is < 1e-6 relative)
constexpr float BAND_R2 = 2.0e-5f;

__device__ __forceinline__ void emit_pair(const PairParams &P, u32 slot_i, u32 slot_j, double ttc, double dist,
                                          double rs, double risk, double mx, double my, double mz, double tcl,
                                          double dcl, int prio, int offset, bool predicted) {
    unsigned long long pos = atomicAdd(&P.counters->n_pairs, 1ULL);
    if (risk > 0.7) atomicAdd(&P.counters->n_high_risk, 1ULL);
    if (prio >= 0) atomicAdd(&P.counters->n_alerts[prio], 1ULL);
    if (pos < P.out_cap) {
        rcd_pair r;
        r.i = P.in_id ? P.in_id[slot_i] : slot_i;
        r.j = P.in_id ? P.in_id[slot_j] : slot_j;
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

// per-thread query state
struct Query {
    float4 p0, p1, p2;
    u32 slot;     // upload slot
    u32 s;        // position in cell order
    u32 pattern;
    // predict: centre motion coefficients (pattern 0: 0,0; 1: v,0; 2: v,a)
    float uvx, uvy, uvz, uax, uay, uaz;
    float reach2;  // (R + travel)^2 with slack: pairs farther than this can never be candidates
    u32 ncand, npot, nexact;
};

// ---- detect: stages 1-4 for one staged neighbour -----------------------------------------------
__device__ __forceinline__ void test_detect(const PairParams &P, Query &q, const float4 &b0, const float4 &b1,
                                            const float4 &b2, u32 sj, float R, float T, int steps) {
    const float R2 = R * R;
    float dx = b0.x - q.p0.x, dy = b0.y - q.p0.y, dz = b0.z - q.p0.z;  // rel_position = other - self
    float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > R2 * (1.0f + BAND_R2)) return;
    if (sj == q.s) return;  // _spatial_filtering strips self (:224-225)
    if (d2 >= R2 * (1.0f - BAND_R2)) {
        ++q.nexact;
        if (!within_radius_d(q.p0.x, q.p0.y, q.p0.z, b0.x, b0.y, b0.z, (double)R)) return;
    }
    ++q.ncand;
    float rvx = q.p1.x - b1.x, rvy = q.p1.y - b1.y, rvz = q.p1.z - b1.z;  // rel_velocity = self - other
    float rs2 = rvx * rvx + rvy * rvy + rvz * rvz;
    if (rs2 < 0.0099f) return;  // rel_speed < 0.1 with margin (0.1^2 = 0.01)
    float dot = dx * rvx + dy * rvy + dz * rvz;
    float edot = 4.0e-6f * sqrtf(d2 * rs2) + 1.0e-20f;
    // dot > 0: either (dot > 0 and cur > 5) or time_to_closest < 0 rejects the pair (:273, :280)
    if (dot > edot) return;
    if (-dot > T * rs2 * (1.0f + 1.0e-5f) + edot) return;  // time_to_closest > time_window
    float tc = fmaxf(-dot, 0.0f) / rs2;
    float rax = q.p2.x - b2.x, ray = q.p2.y - b2.y, raz = q.p2.z - b2.z;
    float h = 0.5f * tc * tc;
    float ex = rvx * tc + rax * h - dx, ey = rvy * tc + ray * h - dy, ez = rvz * tc + raz * h - dz;
    float cd2 = ex * ex + ey * ey + ez * ez;
    float safe = (q.p0.w + b0.w) * 0.5f + 5.0f;
    float tcerr = edot / rs2 + 4.0e-6f * tc;
    float band = 2.0e-3f + 2.0f * (sqrtf(rs2) + sqrtf(rax * rax + ray * ray + raz * raz) * tc) * tcerr;
    float thr = safe + band;
    if (cd2 > thr * thr) return;
    // ---- survivor: decide everything in fp64, in the reference's operation order -------------
    ++q.nexact;
    ObjD A = widen(q.p0, q.p1, q.p2), B = widen(b0, b1, b2);
    DetectResultD r = detect_pair_d(A, B, (double)T, steps);
    if (r.potential) ++q.npot;
    if (r.hit)
        emit_pair(P, q.slot, P.sorted_slot[sj], r.ttc, r.dist, r.rs, r.risk, r.mx, r.my, r.mz, r.tc, r.cd, r.priority, 255,
                  false);
}

// ---- predict: 20 offsets x (radius test, <=10 samples), max-risk merge --------------------------
__device__ __forceinline__ void test_predict(const PairParams &P, Query &q, const float4 &b0, const float4 &b1,
                                             const float4 &b2, u32 sj) {
    float dx = b0.x - q.p0.x, dy = b0.y - q.p0.y, dz = b0.z - q.p0.z;
    float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > q.reach2) return;
    if (sj == q.s) return;
    const float R2 = PREDICT_RADIUS * PREDICT_RADIUS;
    float rvx = q.p1.x - b1.x, rvy = q.p1.y - b1.y, rvz = q.p1.z - b1.z;
    float rax = q.p2.x - b2.x, ray = q.p2.y - b2.y, raz = q.p2.z - b2.z;
    float rvn = sqrtf(rvx * rvx + rvy * rvy + rvz * rvz);
    float ran = sqrtf(rax * rax + ray * ray + raz * raz);
    float safe = (q.p0.w + b0.w) * 0.5f + 5.0f;
    float safe_b = safe + 2.0e-3f + 1.0e-6f * sqrtf(d2);
    float safe_b2 = safe_b * safe_b;
    // the 10 samples move the pair by at most |rv|*0.9 + |ra|*0.405 from the offset state
    float hr = safe_b + rvn * 0.9f + ran * 0.405f;
    float hr2 = hr * hr;
    double best_risk = -1.0;
    PredictResultD best;
    int best_m = -1;
#pragma unroll 1
    for (int m = 0; m < PREDICT_OFFSETS; ++m) {
        float t = 0.5f * (float)m;
        float h = 0.5f * t * t;
        // e = centre_i(t) - p_j   (others are looked up at their CURRENT positions, :801-803)
        float ex = q.uvx * t + q.uax * h - dx, ey = q.uvy * t + q.uay * h - dy, ez = q.uvz * t + q.uaz * h - dz;
        float c2 = ex * ex + ey * ey + ez * ez;
        if (c2 > R2 * (1.0f + BAND_R2)) continue;
        if (c2 >= R2 * (1.0f - BAND_R2)) {
            ++q.nexact;
            ObjD A = widen(q.p0, q.p1, q.p2);
            double cx, cy, cz;
            predict_centre_d(A, q.pattern, 0.5 * (double)m, cx, cy, cz);
            if (!within_radius_d(cx, cy, cz, b0.x, b0.y, b0.z, (double)PREDICT_RADIUS)) continue;
        }
        ++q.ncand;
        // g = centre_i(t) - predicted_j(t)  (:814)
        float gx = ex - (b1.x * t + b2.x * h), gy = ey - (b1.y * t + b2.y * h), gz = ez - (b1.z * t + b2.z * h);
        float g2 = gx * gx + gy * gy + gz * gz;
        if (g2 > hr2) continue;
        bool maybe = false;
#pragma unroll
        for (int k = 0; k < PREDICT_STEPS; ++k) {
            float tau = 0.1f * (float)k;
            float hh = 0.5f * tau * tau;
            float rx = gx + rvx * tau + rax * hh, ry = gy + rvy * tau + ray * hh, rz = gz + rvz * tau + raz * hh;
            maybe |= (rx * rx + ry * ry + rz * rz <= safe_b2);
        }
        if (!maybe) continue;
        ++q.nexact;
        ObjD A = widen(q.p0, q.p1, q.p2), B = widen(b0, b1, b2);
        PredictResultD r = predict_pair_d(A, B, q.pattern, m);
        if (r.hit && r.risk > best_risk) {  // strict >, offsets ascending (:862)
            best_risk = r.risk;
            best = r;
            best_m = m;
        }
    }
    if (best_m >= 0)
        emit_pair(P, q.slot, P.sorted_slot[sj], best.ttc, best.dist, best.rs, best.risk, best.mx, best.my, best.mz,
                  0.5 * (double)best_m, 0.0, priority_d(best.risk, best.ttc), best_m, true);
}

// ---- compute-node pair function ------------------------------------------------------------------
__device__ __forceinline__ void test_compute_node(const PairParams &P, Query &q, const float4 &b0,
                                                  const float4 &b1, const float4 &b2, u32 sj) {
    const float R2 = P.R * P.R;
    float dx = b0.x - q.p0.x, dy = b0.y - q.p0.y, dz = b0.z - q.p0.z;
    float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 > R2 * (1.0f + BAND_R2)) return;
    if (d2 >= R2 * (1.0f - BAND_R2)) {
        ++q.nexact;
        if (!within_radius_d(q.p0.x, q.p0.y, q.p0.z, b0.x, b0.y, b0.z, (double)P.R)) return;
    }
    ++q.ncand;               // query_nearby returns the querying vehicle too (quirk Q8)
    if (sj == q.s) return;   // compute_node.py:251-252
    if (d2 > 2500.0f * (1.0f + BAND_R2)) return;  // current_distance > 50
    u32 mj = __float_as_uint(b2.w);
    if (q.pattern == 0u || meta_pattern(mj) == 0u) return;  // predict_position needs >= 2 samples
    float rvx = q.p1.x - b1.x, rvy = q.p1.y - b1.y, rvz = q.p1.z - b1.z;
    float fx = dx - rvx * P.pt, fy = dy - rvy * P.pt, fz = dz - rvz * P.pt;  // future_j - future_i
    float fut2 = fx * fx + fy * fy + fz * fz;
    float rs2 = rvx * rvx + rvy * rvy + rvz * rvz;
    if (fut2 > d2 * (1.0f + 1.0e-4f) + 1.0e-6f && d2 > 16.0f * (1.0f + 1.0e-4f)) return;  // moving apart
    float fut = fmaxf(sqrtf(fut2), 0.1f);
    if (0.4f * sqrtf(rs2) < P.threshold * fut * (1.0f - 1.0e-4f)) return;  // risk < threshold
    ++q.nexact;
    ObjD A = widen(q.p0, q.p1, q.p2), B = widen(b0, b1, b2);
    ComputeNodeResultD r = compute_node_pair_d(A, B, (double)P.pt, (double)P.threshold);
    if (r.hit)
        emit_pair(P, q.slot, P.sorted_slot[sj], r.ttc, r.fut, r.rs, r.risk, r.mx, r.my, r.mz, 0.0, 0.0, -1, 255, false);
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

template <int MODE>
__global__ void __launch_bounds__(TQ) k_pairs(PairParams P) {
    __shared__ float4 sP0[CH], sP1[CH], sP2[CH];
    __shared__ u32 sSlot[CH];
    __shared__ u32 sRowLo[MAX_ROWS];
    __shared__ u32 sRowPrefix[MAX_ROWS + 1];
    __shared__ int sRedI[TQ / 32][6];
    __shared__ float sRedF[TQ / 32];
    __shared__ u32 sScan[TQ / 32];

    const u32 tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const u32 s = blockIdx.x * TQ + tid;
    const bool valid = s < P.n;
    const GridParams g = P.g;

    Query q;
    q.s = s;
    q.ncand = q.npot = q.nexact = 0;
    q.slot = 0;
    q.pattern = 0;
    q.p0 = q.p1 = q.p2 = make_float4(0.f, 0.f, 0.f, 0.f);
    int cx = 0, cy = 0, cz = 0;
    u32 meta = 0;
    if (valid) {
        q.p0 = P.P0[s];
        q.p1 = P.P1[s];
        q.p2 = P.P2[s];
        q.slot = P.sorted_slot[s];
        meta = __float_as_uint(q.p2.w);
        q.pattern = meta_pattern(meta);
        u32 key = P.keys[s];
        cx = (int)(key % (u32)g.nx);
        u32 row = key / (u32)g.nx;
        cy = (int)(row % (u32)g.ny);
        cz = (int)(row / (u32)g.ny);
    }
    const bool owned = valid && (meta & META_OWNED);

    // reach of this query: how far a neighbour can be and still become a candidate
    float reach;
    if (MODE == RCD_MODE_PREDICT && q.pattern != RCD_PAT_NO_HISTORY) {
        float fv = (q.pattern >= RCD_PAT_CONSTANT_VELOCITY) ? 1.0f : 0.0f;
        float fa = (q.pattern == RCD_PAT_ACCELERATING) ? 1.0f : 0.0f;
        q.uvx = q.p1.x * fv; q.uvy = q.p1.y * fv; q.uvz = q.p1.z * fv;
        q.uax = q.p2.x * fa; q.uay = q.p2.y * fa; q.uaz = q.p2.z * fa;
        float travel = sqrtf(q.uvx * q.uvx + q.uvy * q.uvy + q.uvz * q.uvz) * 9.5f +
                       sqrtf(q.uax * q.uax + q.uay * q.uay + q.uaz * q.uaz) * 45.125f;
        reach = (PREDICT_RADIUS + travel) * (1.0f + 1.0e-5f) + 1.0e-2f;
    } else {
        q.uvx = q.uvy = q.uvz = q.uax = q.uay = q.uaz = 0.0f;
        float R = (MODE == RCD_MODE_PREDICT) ? PREDICT_RADIUS : P.R;
        reach = R * (1.0f + 1.0e-5f) + 1.0e-3f;
    }
    if (!(reach < 1.0e30f)) reach = 1.0e30f;  // non-finite velocity: scan everything
    q.reach2 = reach * reach;

    // a tile that crosses a cell-row boundary is processed as two groups (first row / the rest)
    // so that each group's cell box stays tight
    const u32 first_row = P.keys[blockIdx.x * TQ] / (u32)g.nx;
    const int my_group = (valid && ((u32)(cy + cz * g.ny) != first_row)) ? 1 : 0;
    const int ngroups = __syncthreads_or(my_group) ? 2 : 1;

    for (int grp = 0; grp < ngroups; ++grp) {
        const bool active = owned && my_group == grp;
        // ---- block reduction: cell box and maximum reach of the active queries ----------------
        int r0 = active ? cx : 0x7fffffff, r1 = active ? cx : -1;
        int r2 = active ? cy : 0x7fffffff, r3 = active ? cy : -1;
        int r4 = active ? cz : 0x7fffffff, r5 = active ? cz : -1;
        float rh = active ? reach : 0.0f;
        r0 = warp_min(r0); r1 = warp_max(r1); r2 = warp_min(r2); r3 = warp_max(r3);
        r4 = warp_min(r4); r5 = warp_max(r5); rh = warp_maxf(rh);
        __syncthreads();  // previous group's readers of sRed* are done
        if (lane == 0) {
            sRedI[warp][0] = r0; sRedI[warp][1] = r1; sRedI[warp][2] = r2;
            sRedI[warp][3] = r3; sRedI[warp][4] = r4; sRedI[warp][5] = r5;
            sRedF[warp] = rh;
        }
        __syncthreads();
        int cxmin = 0x7fffffff, cxmax = -1, cymin = 0x7fffffff, cymax = -1, czmin = 0x7fffffff, czmax = -1;
        float hmax = 0.0f;
#pragma unroll
        for (int w = 0; w < TQ / 32; ++w) {
            cxmin = min(cxmin, sRedI[w][0]); cxmax = max(cxmax, sRedI[w][1]);
            cymin = min(cymin, sRedI[w][2]); cymax = max(cymax, sRedI[w][3]);
            czmin = min(czmin, sRedI[w][4]); czmax = max(czmax, sRedI[w][5]);
            hmax = fmaxf(hmax, sRedF[w]);
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
            __syncthreads();  // previous batch's readers of sRow* / sScan are done
            if (lane == 31) sScan[warp] = incl;
            __syncthreads();
            u32 wbase = 0;
            for (u32 w = 0; w < warp; ++w) wbase += sScan[w];
            sRowLo[tid] = lo;
            sRowPrefix[tid] = wbase + incl - cnt;
            if (tid == TQ - 1) sRowPrefix[MAX_ROWS] = wbase + incl;
            __syncthreads();
            const u32 total = sRowPrefix[MAX_ROWS];

            for (u32 base = 0; base < total; base += CH) {
                // ---- stage CH neighbours -------------------------------------------------------
                u32 f = base + tid;
                if (f < total) {
                    // last row r with prefix[r] <= f
                    int a = 0, b = MAX_ROWS - 1;
                    while (a < b) {
                        int mid = (a + b + 1) >> 1;
                        if (sRowPrefix[mid] <= f) a = mid; else b = mid - 1;
                    }
                    u32 src = sRowLo[a] + (f - sRowPrefix[a]);
                    sP0[tid] = P.P0[src];
                    sP1[tid] = P.P1[src];
                    sP2[tid] = P.P2[src];
                    sSlot[tid] = src;
                }
                __syncthreads();
                if (active) {
                    const u32 m = min((u32)CH, total - base);
                    for (u32 jj = 0; jj < m; ++jj) {
                        const float4 b0 = sP0[jj];
                        const u32 sj = sSlot[jj];
                        if (MODE == RCD_MODE_DETECT) {
                            float dx = b0.x - q.p0.x, dy = b0.y - q.p0.y, dz = b0.z - q.p0.z;
                            if (dx * dx + dy * dy + dz * dz > q.reach2) continue;
                            test_detect(P, q, b0, sP1[jj], sP2[jj], sj, P.R, P.T, P.steps);
                        } else if (MODE == RCD_MODE_PREDICT) {
                            float dx = b0.x - q.p0.x, dy = b0.y - q.p0.y, dz = b0.z - q.p0.z;
                            if (dx * dx + dy * dy + dz * dz > q.reach2) continue;
                            if (q.pattern == RCD_PAT_NO_HISTORY)  // history < 2 -> detect_collisions(id) (:590-592)
                                test_detect(P, q, b0, sP1[jj], sP2[jj], sj, PREDICT_RADIUS, 10.0f, 100);
                            else
                                test_predict(P, q, b0, sP1[jj], sP2[jj], sj);
                        } else {
                            float dx = b0.x - q.p0.x, dy = b0.y - q.p0.y, dz = b0.z - q.p0.z;
                            if (dx * dx + dy * dy + dz * dz > q.reach2) continue;
                            test_compute_node(P, q, b0, sP1[jj], sP2[jj], sj);
                        }
                    }
                }
                __syncthreads();
            }
        }
    }

    // ---- per-object candidate count + frame totals ------------------------------------------------
    if (owned && P.cand_count) P.cand_count[q.slot] = q.ncand;
    unsigned long long c = warp_sum((unsigned long long)q.ncand);
    unsigned long long p = warp_sum((unsigned long long)q.npot);
    unsigned long long e = warp_sum((unsigned long long)q.nexact);
    if (lane == 0) {
        if (c) atomicAdd(&P.counters->n_candidates, c);
        if (p) atomicAdd(&P.counters->n_potential, p);
        if (e) atomicAdd(&P.counters->n_exact, e);
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

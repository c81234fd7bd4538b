// fp64 re-evaluation of a pair, operation by operation in the reference's order.
//
// The fp32 kernels only *pre-filter*: a pair is dropped in fp32 only when a threshold is missed
// by more than a guard band that dominates the fp32 rounding error.  Everything else comes here
// and is decided in IEEE double, with explicit __dmul_rn/__dadd_rn so that nvcc cannot contract
// a*b+c into an FMA (CPython evaluates every operation separately).  Squares are x*x: glibc's
// pow(x, 2.0) that `x ** 2` goes through differs from x*x in the last ulp for ~0.08 % of inputs
//  -- a 1e-16 relative effect, irrelevant unless a value sits within one ulp of a
// threshold.
#pragma once
#include "rcd_common.cuh"

namespace rcd {

struct ObjD {  // one object's state widened to double
    double px, py, pz, vx, vy, vz, ax, ay, az, size, heading;
    u32 type;
};

__device__ __forceinline__ ObjD widen(const float4 &p0, const float4 &p1, const float4 &p2) {
    ObjD o;
    o.px = p0.x; o.py = p0.y; o.pz = p0.z; o.size = p0.w;
    o.vx = p1.x; o.vy = p1.y; o.vz = p1.z; o.heading = p1.w;
    o.ax = p2.x; o.ay = p2.y; o.az = p2.z;
    o.type = meta_type(__float_as_uint(p2.w));
    return o;
}

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
// x / 2 as the reference writes it: the same real number rounded once, so x * 0.5 gives the same bits without the
// division sequence
__device__ __forceinline__ double half_d(double x) { return __dmul_rn(x, 0.5); }

// collision_detection.py:391-406 (and spatial_index.py:285-300)
__device__ __forceinline__ double dist3_d(double x1, double y1, double z1, double x2, double y2, double z2) {
    double dx = dsub(x1, x2), dy = dsub(y1, y2), dz = dsub(z1, z2);
    return __dsqrt_rn(dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz)));
}
// collision_detection.py:408-418
__device__ __forceinline__ double mag3_d(double x, double y, double z) {
    return __dsqrt_rn(dadd(dadd(dmul(x, x), dmul(y, y)), dmul(z, z)));
}
// collision_detection.py:433-449: p + v*t + 0.5*a*t*t with Python precedence
__device__ __forceinline__ double pos1_d(double p, double v, double a, double t) {
    return dadd(dadd(p, dmul(v, t)), dmul(dmul(dmul(0.5, a), t), t));
}
// collision_detection.py:484-496
__device__ __forceinline__ double safe_d(double s1, double s2) {
    return dadd(half_d(dadd(s1, s2)), SAFE_DISTANCE_DEFAULT);
}

// collision_detection.py:344-389 + :498-513
__device__ __noinline__ double risk_level_d(double heading_i, double heading_j, u32 type_i, u32 type_j,
                                            double collision_time, double distance, double safe,
                                            double rel_speed) {
    double heading_diff = fabs(dsub(heading_i, heading_j));
    double angle_factor = sin(heading_diff);
    double type_factor = (type_i == type_j) ? 0.5 : 0.8;
    double distance_factor = dsub(1.0, __ddiv_rn(distance, safe));
    double r = __ddiv_rn(collision_time, MAX_WARNING_TIME);
    double time_factor = dsub(1.0, r < 1.0 ? r : 1.0);
    double s = __ddiv_rn(rel_speed, MAX_RELATIVE_SPEED);
    double speed_factor = s < 1.0 ? s : 1.0;
    double risk = dadd(dadd(dadd(dadd(dmul(W_DISTANCE, distance_factor), dmul(W_TIME, time_factor)),
                                 dmul(W_SPEED, speed_factor)),
                            dmul(W_ANGLE, angle_factor)),
                       dmul(W_TYPE, type_factor));
    double m = risk < 1.0 ? risk : 1.0;
    return m > 0.0 ? m : 0.0;
}

// warning_system.py:259-311
__device__ __forceinline__ int priority_d(double risk, double ttc) {
    if (risk < RISK_LOW) return -1;
    if (risk >= RISK_HIGH && ttc < 3.0) return 3;
    if (risk >= RISK_HIGH || ttc < 5.0) return 2;
    if (risk >= RISK_MEDIUM) return 1;
    return 0;
}

struct HitD {
    int k;           // first sample index inside the safe distance, -1 if none
    double dist;     // distance at that sample
    double mx, my, mz;  // midpoint
};

// collision_detection.py:296-342: samples t = k*0.1, k < steps; first with distance <= safe.
// (pi*, vi*, ai*) / (pj*, vj*, aj*) are the two start states.
__device__ __noinline__ HitD precise_hit_d(double pix, double piy, double piz, double pjx, double pjy,
                                           double pjz, const ObjD &a, const ObjD &b, double safe,
                                           int steps, double time_step = 0.1) {
    HitD h;
    h.k = -1;
    h.dist = 0; h.mx = 0; h.my = 0; h.mz = 0;
    for (int k = 0; k < steps; ++k) {
        double t = dmul((double)k, time_step);
        double xi = pos1_d(pix, a.vx, a.ax, t), yi = pos1_d(piy, a.vy, a.ay, t), zi = pos1_d(piz, a.vz, a.az, t);
        double xj = pos1_d(pjx, b.vx, b.ax, t), yj = pos1_d(pjy, b.vy, b.ay, t), zj = pos1_d(pjz, b.vz, b.az, t);
        double d = dist3_d(xi, yi, zi, xj, yj, zj);
        if (d <= safe) {
            h.k = k; h.dist = d;
            h.mx = half_d(dadd(xi, xj));
            h.my = half_d(dadd(yi, yj));
            h.mz = half_d(dadd(zi, zj));
            return h;
        }
    }
    return h;
}

struct DetectResultD {
    bool potential;  // passed stage 2 (temporal filter)
    bool hit;        // emitted
    double tc, cd;   // stage-2 values
    double ttc, dist, rs, risk, mx, my, mz;
    int priority;
};

// stages 2-4 of detect(i) for one candidate j (collision_detection.py:244-292, :296-389)
__device__ __noinline__ DetectResultD detect_pair_d(const ObjD &a, const ObjD &b, double T, int steps) {
    DetectResultD r;
    r.potential = false; r.hit = false;
    r.tc = r.cd = r.ttc = r.dist = r.rs = r.risk = r.mx = r.my = r.mz = 0.0;
    r.priority = -1;
    double cur = dist3_d(a.px, a.py, a.pz, b.px, b.py, b.pz);
    double rvx = dsub(a.vx, b.vx), rvy = dsub(a.vy, b.vy), rvz = dsub(a.vz, b.vz);
    double rpx = dsub(b.px, a.px), rpy = dsub(b.py, a.py), rpz = dsub(b.pz, a.pz);
    double rs = mag3_d(rvx, rvy, rvz);
    if (rs < 0.1) return r;
    double dot = dadd(dadd(dmul(rpx, rvx), dmul(rpy, rvy)), dmul(rpz, rvz));
    if (dot > 0 && cur > SAFE_DISTANCE_DEFAULT) return r;
    double tc = __ddiv_rn(-dot, dmul(rs, rs));
    if (tc < 0 || tc > T) return r;
    double cd = dist3_d(pos1_d(a.px, a.vx, a.ax, tc), pos1_d(a.py, a.vy, a.ay, tc), pos1_d(a.pz, a.vz, a.az, tc),
                        pos1_d(b.px, b.vx, b.ax, tc), pos1_d(b.py, b.vy, b.ay, tc), pos1_d(b.pz, b.vz, b.az, tc));
    double safe = safe_d(a.size, b.size);
    if (cd > safe) return r;
    r.potential = true; r.tc = tc; r.cd = cd;
    HitD h = precise_hit_d(a.px, a.py, a.pz, b.px, b.py, b.pz, a, b, safe, steps);
    if (h.k < 0) return r;
    double ct = dmul((double)h.k, 0.1);
    r.hit = true;
    r.ttc = ct; r.dist = h.dist; r.rs = rs;
    r.mx = h.mx; r.my = h.my; r.mz = h.mz;
    r.risk = risk_level_d(a.heading, b.heading, a.type, b.type, ct, h.dist, safe, rs);
    r.priority = priority_d(r.risk, ct);
    return r;
}

// centre of the predicted vehicle at offset t (collision_detection.py:728-761)
__device__ __forceinline__ void predict_centre_d(const ObjD &a, u32 pattern, double t, double &cx, double &cy,
                                                 double &cz) {
    if (pattern == RCD_PAT_STATIONARY) {
        cx = a.px; cy = a.py; cz = a.pz;
    } else if (pattern == RCD_PAT_CONSTANT_VELOCITY) {
        cx = dadd(a.px, dmul(a.vx, t)); cy = dadd(a.py, dmul(a.vy, t)); cz = dadd(a.pz, dmul(a.vz, t));
    } else {
        cx = pos1_d(a.px, a.vx, a.ax, t); cy = pos1_d(a.py, a.vy, a.ay, t); cz = pos1_d(a.pz, a.vz, a.az, t);
    }
}

struct PredictResultD {
    bool hit;
    double ttc, dist, rs, risk, mx, my, mz;
};

// one (i, j, m) of predict(i): collision_detection.py:801-842 after the radius test
__device__ __noinline__ PredictResultD predict_pair_d(const ObjD &a, const ObjD &b, u32 pattern, int m) {
    PredictResultD r;
    r.hit = false;
    r.ttc = r.dist = r.rs = r.risk = r.mx = r.my = r.mz = 0.0;
    double t = dmul(0.5, (double)m);
    double cx, cy, cz;
    predict_centre_d(a, pattern, t, cx, cy, cz);
    double qx = pos1_d(b.px, b.vx, b.ax, t), qy = pos1_d(b.py, b.vy, b.ay, t), qz = pos1_d(b.pz, b.vz, b.az, t);
    double safe = safe_d(a.size, b.size);
    HitD h = precise_hit_d(cx, cy, cz, qx, qy, qz, a, b, safe, PREDICT_STEPS);
    if (h.k < 0) return r;
    double ct = dmul((double)h.k, 0.1);
    double rs = mag3_d(dsub(a.vx, b.vx), dsub(a.vy, b.vy), dsub(a.vz, b.vz));
    r.hit = true;
    r.risk = risk_level_d(a.heading, b.heading, a.type, b.type, ct, h.dist, safe, rs);
    r.ttc = dadd(ct, t);
    r.dist = h.dist; r.rs = rs; r.mx = h.mx; r.my = h.my; r.mz = h.mz;
    return r;
}

// exact radius test of the broad phase (spatial_index.py:261-269)
__device__ __noinline__ bool within_radius_d(double cx, double cy, double cz, double px, double py, double pz,
                                             double R) {
    return dist3_d(cx, cy, cz, px, py, pz) <= R;
}

struct ComputeNodeResultD {
    bool hit;
    double ttc, fut, rs, risk, mx, my, mz;
};

// compute-node pair function (src/compute/compute_node.py:258-317); `** 0.5` evaluated as sqrt
__device__ __noinline__ ComputeNodeResultD compute_node_pair_d(const ObjD &a, const ObjD &b, double pt,
                                                               double threshold) {
    ComputeNodeResultD r;
    r.hit = false;
    r.ttc = r.fut = r.rs = r.risk = r.mx = r.my = r.mz = 0.0;
    const double vehicle_radius = 2.0;
    double cur = dist3_d(a.px, a.py, a.pz, b.px, b.py, b.pz);
    if (cur > 50.0) return r;
    double fix = dadd(a.px, dmul(a.vx, pt)), fiy = dadd(a.py, dmul(a.vy, pt)), fiz = dadd(a.pz, dmul(a.vz, pt));
    double fjx = dadd(b.px, dmul(b.vx, pt)), fjy = dadd(b.py, dmul(b.vy, pt)), fjz = dadd(b.pz, dmul(b.vz, pt));
    double fut = dist3_d(fix, fiy, fiz, fjx, fjy, fjz);
    double rs = mag3_d(dsub(a.vx, b.vx), dsub(a.vy, b.vy), dsub(a.vz, b.vz));
    if (fut > cur && cur > dmul(vehicle_radius, 2.0)) return r;
    double min_distance = fut > 0.1 ? fut : 0.1;
    double rl = __ddiv_rn(dmul(__ddiv_rn(dmul(vehicle_radius, 2.0), min_distance), rs), 10.0);
    double risk = rl < 1.0 ? rl : 1.0;
    if (risk < threshold) return r;
    double ttc = pt;
    if (fut < dmul(vehicle_radius, 2.0)) {
        if (cur > fut) {
            double ratio = __ddiv_rn(dsub(cur, dmul(vehicle_radius, 2.0)), dsub(cur, fut));
            double v = dmul(pt, ratio);
            ttc = v > 0.1 ? v : 0.1;
        }
    }
    r.hit = true;
    r.ttc = ttc; r.fut = fut; r.rs = rs; r.risk = risk;
    r.mx = half_d(dadd(fix, fjx));
    r.my = half_d(dadd(fiy, fjy));
    r.mz = half_d(dadd(fiz, fjz));
    return r;
}

}  // namespace rcd

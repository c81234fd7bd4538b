// Device side of batched ingest: scatter decoded vehicle records (include/rcd.h, rcd_record) into the
// frame's SoA state and append their (x, y, z, t) float64 samples to the trajectory rings.  Replaces, per
// message, CollisionDetector.update_vehicle + CollisionPredictionModel.update_trajectory as called
// from EarlyWarningSystem._handle_vehicle_position (src/collision/warning_system.py:638-678,
// collision_detection.py:74-85, 553-570).
// Algorithmic bytes per record: 72 read + 50 written (+ 32 into the ring + 4 ring counter r/w).
#pragma once
#include "rcd_common.cuh"
#include "rcd_pairs.cuh"  // Sample64, MutableState

namespace rcd {

constexpr int APPLY_THREADS = 256;
constexpr int RECORD_WORDS = sizeof(rcd_record) / 8;  // 9 x 8 bytes

// A block stages its 256 records through shared memory with coalesced 8-byte loads (consecutive threads
// read consecutive words), then each thread takes one record.  Only records whose `seq` equals the
// launch's are applied: a batch with several messages for one vehicle is applied in `max_seq + 1`
// launches, which keeps "the last message wins" and the order of the ring samples deterministic.
__global__ void __launch_bounds__(APPLY_THREADS)
k_apply_records(u32 n, const unsigned long long *__restrict__ recs, u32 seq, u32 n_objects, MutableState st,
                Sample64 *__restrict__ hist, u32 *__restrict__ count, u32 cap, u32 H) {
    __shared__ unsigned long long s_rec[APPLY_THREADS * RECORD_WORDS + APPLY_THREADS / 8];
    const u32 base = blockIdx.x * APPLY_THREADS;
    const u32 m = min((u32)APPLY_THREADS, n - base);
    const unsigned long long *src = recs + (size_t)base * RECORD_WORDS;
    for (u32 w = threadIdx.x; w < m * RECORD_WORDS; w += APPLY_THREADS) s_rec[w + w / 72] = __ldcs(src + w);
    __syncthreads();
    if (threadIdx.x >= m) return;
    // (one padding word per 8 records keeps the 9-word stride off the same banks)
    auto word = [&](u32 k) { const u32 w = threadIdx.x * RECORD_WORDS + k; return s_rec[w + w / 72]; };
    const unsigned long long tail = word(8);
    const u32 slot = (u32)(tail & 0xffffffffu);
    const u32 type = (u32)((tail >> 32) & 0xffu), rseq = (u32)((tail >> 40) & 0xffu);
    if (rseq != seq || slot >= n_objects) return;
    const double x = __longlong_as_double((long long)word(0)), y = __longlong_as_double((long long)word(1)),
                 z = __longlong_as_double((long long)word(2)), t = __longlong_as_double((long long)word(3));
    st.f[0][slot] = (float)x;
    st.f[1][slot] = (float)y;
    st.f[2][slot] = (float)z;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned long long w = word(4 + k);
        st.f[3 + 2 * k][slot] = __uint_as_float((u32)(w & 0xffffffffu));
        st.f[4 + 2 * k][slot] = __uint_as_float((u32)(w >> 32));
    }
    st.type[slot] = (uint8_t)type;
    if (hist) {
        const u32 c = count[slot];
        Sample64 v;
        v.x = x; v.y = y; v.z = z; v.t = t;
        hist[(size_t)(c % H) * cap + slot] = v;
        count[slot] = c + 1;
    }
}

// new slots: ids = slot index, default pattern
__global__ void __launch_bounds__(256) k_init_slots(u32 first, u32 n, u32 *__restrict__ id, uint8_t *__restrict__ pattern) {
    u32 i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        id[i] = i;
        pattern[i] = RCD_PAT_ACCELERATING;
    }
}

}  // namespace rcd

// tracks.cuh -- SURVEY 8f-2: data association and track lifecycle on the device.
//
// Restates the non-first-frame body of ObstacleTrack::cloudCallback (reference MOT.cpp:176-219) with
// fill_with_linear_interpolation (:593-619), updateObstacleQueue (:586-591), registerNewObstacle (:507-543, minus the
// RViz colour) and unregisterOldObstacle (:545-584).  The association is order dependent by construction -- a centroid
// takes the FIRST registered track (registration order) whose last XY lies within id_threshold, matching is not
// exclusive, and a track registered for centroid k is matchable by centroid k+1 of the same frame -- so centroids are
// processed one after the other by ONE CTA; what is parallel is the search over the tracks (block-wide arg-min of the
// matching slot) -- O(K*T) compares, ~10^6 at config c5.
//
// Track state (device resident, per handle): ids[T], rings[T][L] (x, y, z, intensity=time; oldest first, the layout of
// stack_obj, MOT.h:107), m_state[T][4] (the IHGP carry).  meta: [0] n_tracks, [1] next_obj_num, [2] max occurrence.
#pragma once
#include "common.cuh"

namespace mot {

constexpr int ASSOC_THREADS = 1024;
enum { TM_NTRACKS = 0, TM_NEXT_ID = 1, TM_MAX_OCC = 2, TM_OVERFLOW = 3, TM_N = 4 };

__device__ __forceinline__ float ref_euc_dist_xy(float ax, float ay, float bx, float by) {  // euc_dist (MOT.cpp:1025-1028), z = 0 on both sides
    const double dx = __dsub_rn((double)ax, (double)bx), dy = __dsub_rn((double)ay, (double)by);
    return __double2float_rn(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 0.0)));
}

__global__ void __launch_bounds__(ASSOC_THREADS) k_associate(const float4* __restrict__ centroids, int K, int L, int max_tracks, float id_threshold,
                                                              float dt_gp, int* __restrict__ ids, float4* __restrict__ rings,
                                                              double* __restrict__ m_state, int* __restrict__ meta, int* __restrict__ seen,
                                                              int* __restrict__ this_ids, int* __restrict__ slot_of_entry,
                                                              int* __restrict__ occurrence) {
    __shared__ int s_best[ASSOC_THREADS / 32];
    __shared__ int s_n;
    if (threadIdx.x == 0) {
        s_n = meta[TM_NTRACKS];
        meta[TM_OVERFLOW] = 0;  // centroids dropped by THIS call because the table was full (the purge may free slots later)
    }
    for (int t = threadIdx.x; t < max_tracks; t += ASSOC_THREADS) seen[t] = 0;
    __syncthreads();
    int max_occ = 0;
    for (int k = 0; k < K; ++k) {
        const float4 obj = centroids[k];
        const int n = s_n;
        // first registered track whose last observation is within id_threshold (MOT.cpp:184-207)
        int best = 0x7fffffff;
        for (int t = threadIdx.x; t < n; t += ASSOC_THREADS) {
            const float4 last = rings[(size_t)t * L + (L - 1)];
            if (ref_euc_dist_xy(obj.x, obj.y, last.x, last.y) < id_threshold) { best = t; break; }  // t ascending per thread
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(kFull, best, o));
        if (lane_id() == 0) s_best[warp_id()] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < ASSOC_THREADS / 32; ++w) best = min(best, s_best[w]);
            int slot;
            if (best != 0x7fffffff) {
                slot = best;
                float4* ring = rings + (size_t)slot * L;
                const float4 last0 = ring[L - 1];
                if (__fsub_rn(obj.w, last0.w) > __fmul_rn(3.0f, dt_gp)) {
                    // fill_with_linear_interpolation (MOT.cpp:593-619)
                    const double dx_total = (double)__fsub_rn(obj.x, last0.x), dy_total = (double)__fsub_rn(obj.y, last0.y);
                    const double dt_total = (double)__fsub_rn(obj.w, last0.w);
                    const int lost = (int)round(__ddiv_rn(dt_total, (double)dt_gp)) - 1;
                    for (int j = 0; j < lost; ++j) {
                        const float4 lc = ring[L - 1];
                        float4 c;
                        c.x = __double2float_rn(__dadd_rn((double)lc.x, __ddiv_rn(dx_total, (double)lost)));
                        c.y = __double2float_rn(__dadd_rn((double)lc.y, __ddiv_rn(dy_total, (double)lost)));
                        c.z = __double2float_rn(__dadd_rn((double)lc.z, __ddiv_rn(0.0, (double)lost)));
                        c.w = __fadd_rn(lc.w, dt_gp);
                        for (int i = 0; i + 1 < L; ++i) ring[i] = ring[i + 1];
                        ring[L - 1] = c;
                    }
                }
                for (int i = 0; i + 1 < L; ++i) ring[i] = ring[i + 1];  // updateObstacleQueue (MOT.cpp:586-591)
                ring[L - 1] = obj;
                this_ids[k] = ids[slot];
            } else if (n < max_tracks) {
                // registerNewObstacle (MOT.cpp:507-543): ring filled with the centroid, fresh GP state (m = 0)
                slot = n;
                const int id = meta[TM_NEXT_ID];
                meta[TM_NEXT_ID] = id + 1;
                ids[slot] = id;
                for (int i = 0; i < L; ++i) rings[(size_t)slot * L + i] = obj;
                for (int i = 0; i < 4; ++i) m_state[(size_t)slot * 4 + i] = 0.0;
                this_ids[k] = id;
                s_n = n + 1;
            } else {
                slot = -1;  // track table full: reported to the host, the centroid is skipped
                meta[TM_OVERFLOW] += 1;
                this_ids[k] = -1;
            }
            slot_of_entry[k] = slot;
            int occ = 0;
            if (slot >= 0) { occ = seen[slot]; seen[slot] = occ + 1; }
            occurrence[k] = slot >= 0 ? occ : 0x7fffffff;
            if (slot >= 0 && occ > max_occ) max_occ = occ;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        meta[TM_NTRACKS] = s_n;
        meta[TM_MAX_OCC] = max_occ;
    }
}

// unregisterOldObstacle (MOT.cpp:545-584): drop tracks whose last observation is older than `period` seconds; the
// survivors keep their order.  One CTA; source and destination are the two halves of a ping-pong buffer.
__global__ void __launch_bounds__(ASSOC_THREADS) k_tracks_purge(const int* __restrict__ ids_in, const float4* __restrict__ rings_in,
                                                                 const double* __restrict__ m_in, int* __restrict__ ids_out,
                                                                 float4* __restrict__ rings_out, double* __restrict__ m_out, int L, double now,
                                                                 double period, int* __restrict__ meta) {
    __shared__ int scratch[36];
    __shared__ int s_base;
    const int n = meta[TM_NTRACKS];
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int tb = 0; tb < n; tb += ASSOC_THREADS) {
        const int t = tb + threadIdx.x;
        int keep = 0;
        if (t < n) keep = !(now - (double)rings_in[(size_t)t * L + (L - 1)].w > period);
        int total;
        const int excl = block_exclusive_scan(keep, scratch, &total);
        const int base = s_base;
        if (keep) {
            const int d = base + excl;
            ids_out[d] = ids_in[t];
            for (int i = 0; i < L; ++i) rings_out[(size_t)d * L + i] = rings_in[(size_t)t * L + i];
            for (int i = 0; i < 4; ++i) m_out[(size_t)d * 4 + i] = m_in[(size_t)t * 4 + i];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) meta[TM_NTRACKS] = s_base;
}

}  // namespace mot

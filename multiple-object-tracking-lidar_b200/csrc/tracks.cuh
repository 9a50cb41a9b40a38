// tracks.cuh -- SURVEY 8f-2: data association and track lifecycle on the device.
//
// Restates the non-first-frame body of ObstacleTrack::cloudCallback (reference MOT.cpp:176-219) with
// fill_with_linear_interpolation (:593-619), updateObstacleQueue (:586-591), registerNewObstacle (:507-543, minus the
// RViz colour) and unregisterOldObstacle (:545-584).  The association is order dependent by construction -- a centroid
// takes the FIRST registered track (registration order) whose last XY lies within id_threshold, matching is not
// exclusive, and a track registered for centroid k is matchable by centroid k+1 of the same frame -- so centroids are
// processed one after the other by ONE CTA; what is parallel is the search over the tracks (block-wide arg-min of the
// matching slot) -- O(K*T) compares, ~10^6 at config c5 -- and, in the round-2 form (k_associate_fast + k_tracks_apply, used for
// tables of up to 8192 tracks), everything a match implies: ring shifts, interpolation fills, registrations.
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

// ---- the same association with the per-centroid step cut down to what is really sequential --------------------------------------
// k_associate walks the centroids one after the other and lets thread 0 do everything a match implies -- reading the track's last
// observation, shifting its ring of L entries, the interpolation fill, the id lookup -- in global memory: ~3.5 us per centroid,
// 3.5 ms for 1,000 (config c5), ten times the rest of the tracker step.  What the NEXT centroid's decision depends on is only the
// table of last observations (a matched track's last observation becomes the centroid; a new track appends one) -- so that table
// lives in shared memory (16 B per track), the sequential loop touches nothing else (two CTA barriers per centroid, ~0.2 us), and
// everything a match implies is recorded per entry (slot, occurrence, the last observation it replaced, a link to the same
// track's next entry of this frame) and applied afterwards by k_tracks_apply, one warp per track, entries of a track in order.
constexpr int AF_MAX_TRACKS = 8192;  // tracks the shared-memory table holds (16 + 4 + 2 bytes each); larger tables keep k_associate
constexpr size_t af_smem_bytes(int cap) { return (size_t)cap * (sizeof(float4) + sizeof(int) + sizeof(unsigned short)); }

constexpr int AF_CHUNK = 128;  // entries whose records are kept in shared memory between two flushes
// `match_below`: the reference's test float(sqrt(dx^2 + dy^2 + 0)) < id_threshold is monotone in the double s = dx^2 + dy^2 + 0
// (correctly rounded sqrt and the conversion to float are both monotone), so it equals s < match_below for the one double the host
// finds by bisection over the bit patterns (mot_b200.cu: assoc_match_below) -- no 64-bit square root on the critical path.
__global__ void __launch_bounds__(ASSOC_THREADS, 1) k_associate_fast(const float4* __restrict__ centroids, int K, int L, int max_tracks, double match_below,
                                                                      float f_lo, float f_hi,
                                                                      int* __restrict__ ids, const float4* __restrict__ rings, int* __restrict__ meta,
                                                                      int* __restrict__ slot_of_entry, int* __restrict__ occurrence,
                                                                      int* __restrict__ next_entry, int* __restrict__ entry_new,
                                                                      float4* __restrict__ prev_last, int cap) {
    extern __shared__ __align__(16) unsigned char af_smem[];
    float4* s_last = reinterpret_cast<float4*>(af_smem);                // [cap] last observation of every track
    int* s_lastentry = reinterpret_cast<int*>(s_last + cap);            // [cap] this frame's latest entry of the track (or -1)
    unsigned short* s_occ = reinterpret_cast<unsigned short*>(s_lastentry + cap);  // [cap] matches of the track in this frame
    // The per-centroid step is a latency chain, so it is cut to: search (shared memory, all warps but the last) -> warp minimum ->
    // shared-memory atomicMin -> ONE CTA barrier.  Every thread derives the outcome (slot, new track or not, table size) from the
    // minimum for itself; the record keeping -- the table entry, occurrence counter, links -- is done by one thread of the LAST warp
    // while the other warps already search for the next centroid with the one pending table update applied from registers (the
    // bookkeeper's stores become visible at the next barrier, when the update after them is the pending one).  Centroids come in
    // and records go out in chunks of AF_CHUNK, by all threads; nothing in the loop touches global memory.
    __shared__ float4 s_cen[AF_CHUNK], s_prev[AF_CHUNK];
    __shared__ int s_slot[AF_CHUNK], s_occv[AF_CHUNK], s_next[AF_CHUNK], s_new[AF_CHUNK];
    __shared__ int s_min[3];
    const int tid = threadIdx.x;
    const int n_search = blockDim.x - 32;            // threads that search (the CTA has at least two warps)
    const bool keeper = tid == n_search;             // first lane of the last warp
    const bool searcher = tid < n_search;
    const int n0 = meta[TM_NTRACKS], next_id0 = meta[TM_NEXT_ID];
    for (int t = tid; t < cap; t += blockDim.x) {
        if (t < n0) s_last[t] = rings[(size_t)t * L + (L - 1)];
        s_lastentry[t] = -1;
        s_occ[t] = 0;
    }
    if (tid < 3) s_min[tid] = 0x7fffffff;
    int n = n0;                                       // table size, tracked by every thread
    int pend_slot = -1;                               // the one table update the shared-memory copy does not show yet
    float4 pend_obj = make_float4(0.f, 0.f, 0.f, 0.f);
    int max_occ = 0, overflow = 0;                    // keeper
    int step = 0;                                     // centroid counter modulo 3 (buffer of the minimum)
    for (int base = 0; base < K; base += AF_CHUNK) {
        const int cn = min(AF_CHUNK, K - base);
        __syncthreads();  // the previous chunk's records are out, the keeper's last update is visible
        if (tid < cn) s_cen[tid] = centroids[base + tid];
        __syncthreads();
        for (int i = 0; i < cn; ++i) {
            const float4 obj = s_cen[i];
            if (searcher) {
                // first registered track whose last observation is within id_threshold (MOT.cpp:184-207)
                int best = 0x7fffffff;
                for (int t = tid; t < n; t += n_search) {
                    float4 last = s_last[t];
                    if (t == pend_slot) last = pend_obj;
                    // fp32 first: fl(a - b) is the correctly rounded difference, so fs is within 4 ulp-steps (2.4e-7 relative) of the
                    // exact squared distance; outside the +-2e-6 band around the threshold the fp64 value falls on the same side.  The
                    // band itself (and NaN) takes the reference's double expression -- five dependent fp64 operations that otherwise sit
                    // on the critical path of every centroid (ncu: the step is a latency chain, not issue bound)
                    const float fx = __fsub_rn(obj.x, last.x), fy = __fsub_rn(obj.y, last.y);
                    const float fs = __fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy));
                    if (fs > f_hi) continue;
                    if (!(fs < f_lo)) {
                        const double dx = __dsub_rn((double)obj.x, (double)last.x), dy = __dsub_rn((double)obj.y, (double)last.y);
                        if (!(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 0.0) < match_below)) continue;
                    }
                    best = t;  // t ascending per thread
                    break;
                }
                if (best != 0x7fffffff) atomicMin(&s_min[step], best);  // (the compiler aggregates the warp's lanes: one redux + one shared-memory atomic)
            }
            __syncthreads();
            const int b = s_min[step];
            const bool matched = b != 0x7fffffff, fresh = !matched && n < max_tracks;
            const int slot = matched ? b : (fresh ? n : -1);
            if (keeper) {
                s_min[step == 0 ? 2 : step - 1] = 0x7fffffff;  // read by everybody one barrier ago, used again two steps from now
                const int k = base + i;
                if (matched) s_prev[i] = s_last[slot];  // (the keeper's own earlier stores are in program order: its view of the table is current)
                if (slot >= 0) s_last[slot] = obj;  // updateObstacleQueue / registerNewObstacle: the centroid is the track's last observation
                s_slot[i] = slot;
                s_new[i] = fresh ? 1 : 0;
                s_next[i] = -1;
                int occ = 0x7fffffff;
                if (slot >= 0) {
                    occ = s_occ[slot];
                    s_occ[slot] = (unsigned short)(occ + 1);
                    const int pe = s_lastentry[slot];
                    if (pe >= base) s_next[pe - base] = k;
                    else if (pe >= 0) next_entry[pe] = k;  // the same track again, more than a chunk later: its record is already out
                    s_lastentry[slot] = k;
                    if (occ > max_occ) max_occ = occ;
                } else {
                    ++overflow;
                }
                s_occv[i] = occ;
            }
            if (slot >= 0) { pend_slot = slot; pend_obj = obj; }
            n += fresh ? 1 : 0;
            step = step == 2 ? 0 : step + 1;
        }
        __syncthreads();  // the keeper's records of this chunk are complete
        if (tid < cn) {
            slot_of_entry[base + tid] = s_slot[tid];
            occurrence[base + tid] = s_occv[tid];
            next_entry[base + tid] = s_next[tid];
            entry_new[base + tid] = s_new[tid];
            prev_last[base + tid] = s_prev[tid];
        }
    }
    for (int t = n0 + tid; t < n; t += blockDim.x) ids[t] = next_id0 + (t - n0);  // new tracks take consecutive ids in registration order
    if (keeper) {
        meta[TM_NTRACKS] = n;
        meta[TM_MAX_OCC] = max_occ;
        meta[TM_NEXT_ID] = next_id0 + (n - n0);
        meta[TM_OVERFLOW] = overflow;
    }
}

// What the matches imply, one warp per track: the entries of a track are applied in centroid order (the chain next_entry starts at
// the entry with occurrence 0).  A match shifts the ring left by (lost + 1) -- `lost` interpolated observations when the track was
// not seen for more than three periods (fill_with_linear_interpolation, MOT.cpp:593-619: every filled point is computed from the
// previous one, in the reference's own float / double order) -- and appends the centroid; a new track's ring is the centroid L times
// and its IHGP carry starts at zero.
constexpr int TA_THREADS = 256;
__global__ void __launch_bounds__(TA_THREADS) k_tracks_apply(const float4* __restrict__ centroids, int K, int L, float dt_gp, const int* __restrict__ ids,
                                                              float4* __restrict__ rings, double* __restrict__ m_state,
                                                              const int* __restrict__ slot_of_entry, const int* __restrict__ occurrence,
                                                              const int* __restrict__ next_entry, const int* __restrict__ entry_new,
                                                              const float4* __restrict__ prev_last, int* __restrict__ this_ids) {
    const int k0 = blockIdx.x * (TA_THREADS / 32) + warp_id();
    if (k0 >= K) return;
    const int lane = lane_id();
    const int slot0 = slot_of_entry[k0];
    if (slot0 < 0) {  // dropped (table full)
        if (lane == 0) this_ids[k0] = -1;
        return;
    }
    if (occurrence[k0] != 0) return;  // applied by the warp of the track's first entry
    float4* ring = rings + (size_t)slot0 * L;
    const int id = ids[slot0];
    for (int e = k0; e >= 0; e = next_entry[e]) {
        const float4 obj = centroids[e];
        if (lane == 0) this_ids[e] = id;
        if (entry_new[e]) {
            for (int i = lane; i < L; i += 32) ring[i] = obj;
            if (lane < 4) m_state[(size_t)slot0 * 4 + lane] = 0.0;
        } else {
            const float4 last0 = prev_last[e];
            int lost = 0;
            double dx_total = 0.0, dy_total = 0.0;
            if (__fsub_rn(obj.w, last0.w) > __fmul_rn(3.0f, dt_gp)) {
                dx_total = (double)__fsub_rn(obj.x, last0.x);
                dy_total = (double)__fsub_rn(obj.y, last0.y);
                const double dt_total = (double)__fsub_rn(obj.w, last0.w);
                lost = (int)round(__ddiv_rn(dt_total, (double)dt_gp)) - 1;
                if (lost < 0) lost = 0;
            }
            const int shift = lost + 1;
            // the new ring: old[i + shift] while that exists, then the filled points c_0 .. c_{lost-1}, then the centroid.  Chunks of 32
            // in ascending order, read before write: a chunk only reads positions nobody has written yet.
            for (int c = 0; c < L; c += 32) {
                const int i = c + lane;
                float4 v = obj;
                if (i < L) {
                    const long long src = (long long)i + shift;
                    if (src < L) {
                        v = ring[src];
                    } else {
                        const long long j = src - L;  // index into [c_0 .. c_{lost-1}, centroid]
                        if (j < lost) {
                            float4 lc = last0;
                            for (long long q = 0; q <= j; ++q) {
                                float4 f;
                                f.x = __double2float_rn(__dadd_rn((double)lc.x, __ddiv_rn(dx_total, (double)lost)));
                                f.y = __double2float_rn(__dadd_rn((double)lc.y, __ddiv_rn(dy_total, (double)lost)));
                                f.z = __double2float_rn(__dadd_rn((double)lc.z, __ddiv_rn(0.0, (double)lost)));
                                f.w = __fadd_rn(lc.w, dt_gp);
                                lc = f;
                            }
                            v = lc;
                        }
                    }
                }
                __syncwarp();
                if (i < L) ring[i] = v;
                __syncwarp();
            }
        }
        __syncwarp();
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

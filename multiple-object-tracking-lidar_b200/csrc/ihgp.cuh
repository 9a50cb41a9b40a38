// ihgp.cuh -- K9: batched Infinite-Horizon GP track filter, one warp per track.
//
// Replaces the per-track loop of ObstacleTrack::callIHGP (reference MOT.cpp:621-662): LPF_pos
// (MOT.cpp:824-833) and IHGP_fixed_vel (MOT.cpp:871-920), i.e. InfiniteHorizonGP::init_step + (L-1) x update
// + getEft (IHGP.cpp:108-196) for the x and the y axis, then the +-1.5 m/s clamp (MOT.cpp:649-654).
// The stationary matrices (A, AKHA, K, G) are per-axis constants computed once on the host
// (mot_ihgp_configure; IHGP.cpp:12-37, 168-170) and passed by value.  The likelihood / gradient
// recursion of update() (IHGP.cpp:138-154) feeds nothing the tracker reads and is not computed.
//
// Warp layout (round 2): 16 tracks per warp, the x and the y recursion of a track on neighbouring lanes; every lane runs its
// 2x2 recursions in the reference's own operation order in fp64 (bit-faithful sequential order) and keeps the filter means
// the backward pass re-reads in shared memory.
#pragma once
#include "common.cuh"

namespace mot {

struct IhgpAxis {
    double A[4], AKHA[4], K[2], G[4];
};

// one row per track with exactly what ObstacleTrack::publishObstacles (MOT.cpp:253-295) writes into a
// costmap_converter::ObstacleMsg, so a ROS shim is a field-by-field copy.  == mot_obstacle in mot_b200.h
struct ObstacleRow {
    int id;             // obstacle.id            (MOT.cpp:266)
    float radius;       // 0.3                    (MOT.cpp:267)
    float x, y;         // polygon.points[0].x/y  (MOT.cpp:288-290), z = 0
    float vx, vy;       // velocities.twist.linear.x/y (MOT.cpp:272-273), the other twist components = 0
    float vel_cov[6];   // diagonal of velocities.covariance: .1, .1, 1e9, 1e9, 1e9, .1 (MOT.cpp:279-284)
};

constexpr int IHGP_WARPS = 4;
constexpr int IHGP_MAX_EPW = 16;  // entries (tracks) per warp: two lanes each, x and y

// shared memory per lane: the finite-difference velocities (float, staged by the whole warp with coalesced ring loads) and the
// filter means MF[k] (two doubles per sample) that the backward pass (getEft) reads again
inline size_t ihgp_mf_stride(int L) { return 2 * (size_t)(L - 1) + 1; }  // doubles per lane, odd: neighbouring lanes fall into different banks
inline size_t ihgp_smem_bytes(int L, int epw) { return (size_t)IHGP_WARPS * epw * 2 * (ihgp_mf_stride(L) * sizeof(double) + (size_t)(L - 1) * sizeof(float)); }
// entries per warp that fit `budget` bytes of shared memory per CTA (16 at the reference's data_length of 10..40)
// ... and no more than it takes to give every SM a few warps: 1,000 tracks are filtered one per warp (all SMs busy, the
// shortest critical path), 100,000 tracks sixteen per warp (all lanes busy)
inline int ihgp_entries_per_warp(int L, size_t budget, int n_tracks = 1 << 30, int num_sms = 148) {
    int epw = IHGP_MAX_EPW;
    while (epw > 1 && (ihgp_smem_bytes(L, epw) > budget || (long long)n_tracks < (long long)epw * num_sms * IHGP_WARPS * 2)) epw >>= 1;
    return epw;
}

// Lane layout: lane = 2 * e + axis -- a warp filters `epw` tracks at once, the x and the y recursion of a track on neighbouring
// lanes (round 1 ran one track per warp with 2 active lanes).  Every lane runs its axis' 2x2 recursions in the reference's own
// operation order in fp64; the arithmetic per lane is unchanged, so results are bit-identical to the one-track-per-warp layout.
__global__ void __launch_bounds__(IHGP_WARPS * 32) k_ihgp_step(const float4* __restrict__ rings, int T, int L, float dt_gp, float lpf_tau,
                                                                IhgpAxis ax, IhgpAxis ay, double* __restrict__ m_state,
                                                                float4* __restrict__ pos_vel, const int* __restrict__ ids,
                                                                ObstacleRow* __restrict__ obstacles, const int* __restrict__ slot_of_entry,
                                                                const int* __restrict__ occurrence, int round, int epw) {
    extern __shared__ double ihgp_smem[];
    const int n = L - 1;
    const int lane = lane_id();
    const int axis = lane & 1, el = lane >> 1;
    const int slot_l = (warp_id() * epw + (el < epw ? el : 0)) * 2 + axis;  // this lane's arrays
    const size_t mf_stride = 2 * (size_t)n + 1;
    double* mf0 = ihgp_smem + (size_t)slot_l * mf_stride;
    double* mf1 = mf0 + n;
    float* vall = reinterpret_cast<float*>(ihgp_smem + (size_t)IHGP_WARPS * epw * 2 * mf_stride);  // behind all MF arrays
    float* vw = vall + (size_t)warp_id() * epw * 2 * n;                                            // this warp's velocities
    const float* v = vall + (size_t)slot_l * n;
    const IhgpAxis& q = axis == 0 ? ax : ay;
    // Entry e of the call uses the ring / carried state of track slot t (identity unless the on-device association
    // supplies slot_of_entry; an entry whose track already appeared earlier in the same frame runs in a later round,
    // exactly as the reference's sequential callIHGP loop would advance that track's state twice).
    const int n_groups = (T + epw - 1) / epw;
    for (int grp = blockIdx.x * IHGP_WARPS + warp_id(); grp < n_groups; grp += gridDim.x * IHGP_WARPS) {
        // stage: all lanes walk the rings of the group's entries (coalesced 16-byte loads, nothing depends on anything)
        for (int x = 0; x < epw; ++x) {
            const int ex = grp * epw + x;
            if (ex >= T || (occurrence && occurrence[ex] != round)) continue;  // warp uniform
            const float4* cx = rings + (size_t)(slot_of_entry ? slot_of_entry[ex] : ex) * L;
            for (int k = lane; k < n; k += 32) {
                const float4 c0 = cx[k], c1 = cx[k + 1];
                vw[(size_t)(2 * x) * n + k] = __fdiv_rn(__fsub_rn(c1.x, c0.x), dt_gp);      // MOT.cpp:889 (float arithmetic)
                vw[(size_t)(2 * x + 1) * n + k] = __fdiv_rn(__fsub_rn(c1.y, c0.y), dt_gp);  // MOT.cpp:893
            }
        }
        __syncwarp();
        const int e = grp * epw + el;
        const bool active = el < epw && e < T && !(occurrence && occurrence[e] != round);
        float vel = 0.0f;
        int t = 0;
        if (active) {
            t = slot_of_entry ? slot_of_entry[e] : e;
            double mean = 0.0;  // uninitialised in the reference (MOT.cpp:879-880); policy: 0
            for (int k = 0; k < n; ++k) mean = __dadd_rn(mean, (double)v[k]);
            mean = __ddiv_rn(mean, (double)n);
            double m0 = m_state[(size_t)t * 4 + 2 * axis], m1 = m_state[(size_t)t * 4 + 2 * axis + 1];
            for (int k = 0; k < n; ++k) {  // update(): m = AKHA*m + K*y  (IHGP.cpp:157)
                const double y = __dsub_rn((double)v[k], mean);
                const double n0 = __dadd_rn(__dadd_rn(__dmul_rn(q.AKHA[0], m0), __dmul_rn(q.AKHA[1], m1)), __dmul_rn(q.K[0], y));
                const double n1 = __dadd_rn(__dadd_rn(__dmul_rn(q.AKHA[2], m0), __dmul_rn(q.AKHA[3], m1)), __dmul_rn(q.K[1], y));
                m0 = n0; m1 = n1;
                mf0[k] = m0; mf1[k] = m1;
            }
            const double eft_last = m0;  // H*MF.back(), H = [1 0]
            for (int k = n - 2; k >= 0; --k) {  // getEft(): m = MF[k] + G*(m - A*MF[k])  (IHGP.cpp:185-189)
                const double f0 = mf0[k], f1 = mf1[k];
                const double r0 = __dsub_rn(m0, __dadd_rn(__dmul_rn(q.A[0], f0), __dmul_rn(q.A[1], f1)));
                const double r1 = __dsub_rn(m1, __dadd_rn(__dmul_rn(q.A[2], f0), __dmul_rn(q.A[3], f1)));
                m0 = __dadd_rn(f0, __dadd_rn(__dmul_rn(q.G[0], r0), __dmul_rn(q.G[1], r1)));
                m1 = __dadd_rn(f1, __dadd_rn(__dmul_rn(q.G[2], r0), __dmul_rn(q.G[3], r1)));
            }
            m_state[(size_t)t * 4 + 2 * axis] = m0;      // smoothed state at k = 0 is the next frame's carry-in
            m_state[(size_t)t * 4 + 2 * axis + 1] = m1;
            vel = __double2float_rn(__dadd_rn(eft_last, mean));  // MOT.cpp:914-915
            vel = vel > 1.5f ? 1.5f : (vel < -1.5f ? -1.5f : vel);  // MOT.cpp:649-654 (NaN passes through, as in the reference)
        }
        const float vy = __shfl_down_sync(kFull, vel, 1);  // the y lane sits next to the x lane
        if (active && axis == 0) {
            const float4* c = rings + (size_t)t * L;
            const float4 a = c[L - 2], b = c[L - 1];
            const float wa = __fdiv_rn(lpf_tau, __fadd_rn(lpf_tau, dt_gp)), wb = __fdiv_rn(dt_gp, __fadd_rn(lpf_tau, dt_gp));
            float4 pos, v4;
            pos.x = __fadd_rn(__fmul_rn(wa, a.x), __fmul_rn(wb, b.x));  // MOT.cpp:827
            pos.y = __fadd_rn(__fmul_rn(wa, a.y), __fmul_rn(wb, b.y));  // MOT.cpp:828
            pos.z = 0.0f; pos.w = b.w;
            v4.x = vel; v4.y = vy; v4.z = 0.0f; v4.w = b.w;
            pos_vel[(size_t)e * 2] = pos;
            pos_vel[(size_t)e * 2 + 1] = v4;
            if (obstacles) {
                ObstacleRow o;
                o.id = ids ? ids[e] : e;
                o.radius = 0.3f;
                o.x = pos.x; o.y = pos.y;
                o.vx = vel; o.vy = vy;
                o.vel_cov[0] = 0.1f; o.vel_cov[1] = 0.1f; o.vel_cov[2] = 1e9f; o.vel_cov[3] = 1e9f; o.vel_cov[4] = 1e9f; o.vel_cov[5] = 0.1f;
                obstacles[e] = o;
            }
        }
        __syncwarp();
    }
}

}  // namespace mot

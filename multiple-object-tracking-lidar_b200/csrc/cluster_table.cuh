// cluster_table.cuh -- K6/K7/K8: component sizes, min/max-size filter, cluster ordering, CSR emission,
// per-cluster segmented reduction (count / mean / bbox) and the reference's circumcentre "centroid".
//
// Output contract = std::vector<pcl::PointIndices> as the reference sees it after ec.extract()
// (MOT.cpp:488): components outside [min,max] dropped whole, indices ascending inside a cluster, clusters by
// size descending; ties (unspecified in PCL's std::sort) are pinned to smallest-index-first.
#pragma once
#include "common.cuh"
#include "grid_uf.cuh"

namespace mot {

// Layout of the 64-bit cluster ordering key: frame | (size_cap - size) | local min index.
struct ClusterKeyCodec {
    int size_bits, idx_bits;  // frame occupies the bits above
    unsigned size_cap;        // (1 << size_bits) - 1
};
__device__ __forceinline__ uint64_t ckey_make(const ClusterKeyCodec& c, int frame, int size, int local_min) {
    return ((((uint64_t)frame << c.size_bits) | (uint64_t)(c.size_cap - (unsigned)size)) << c.idx_bits) | (uint64_t)local_min;
}
__device__ __forceinline__ int ckey_size(const ClusterKeyCodec& c, uint64_t k) {
    return (int)(c.size_cap - (unsigned)((k >> c.idx_bits) & c.size_cap));
}
__device__ __forceinline__ int ckey_frame(const ClusterKeyCodec& c, uint64_t k) { return (int)(k >> (c.idx_bits + c.size_bits)); }

// Component size and smallest original index, accumulated per fine cell (warp-aggregated so that the root of a
// giant component is not hit by one atomic per cell).
__global__ void __launch_bounds__(256) k_comp_accumulate(const int* __restrict__ fc_start, const int* __restrict__ root,
                                                          int* __restrict__ csize, int* __restrict__ cmin,
                                                          const int* __restrict__ d_counts) {
    const int n_fine = d_counts[CNT_FINE];
    const int stride = gridDim.x * blockDim.x;
    const int rounds = (n_fine + stride - 1) / stride;
    for (int it = 0; it < rounds; ++it) {
        const int c = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = c < n_fine;
        const int r = valid ? root[c] : -1;
        const int n = valid ? fc_start[c + 1] - fc_start[c] : 0;
        const int mi = valid ? cmin[c] : 0x7fffffff;
        const unsigned peers = __match_any_sync(kFull, r);
        const int nsum = __reduce_add_sync(peers, n);
        const int mmin = __reduce_min_sync(peers, mi);
        if (valid && lane_id() == __ffs(peers) - 1) {
            atomicAdd(&csize[r], nsum);
            atomicMin(&cmin[r], mmin);
        }
    }
}

// Kept roots -> unordered list of (ordering key, root fine cell).  The sort that follows makes the order
// deterministic (keys are unique: a point index belongs to one component).
__global__ void __launch_bounds__(256) k_kept_list(const int* __restrict__ root, const int* __restrict__ csize, const int* __restrict__ cmin,
                                                    const int* __restrict__ frame_offsets, int n_frames, int min_size, int max_size,
                                                    ClusterKeyCodec kc, uint64_t* __restrict__ ckeys, uint32_t* __restrict__ croots,
                                                    int* __restrict__ d_counts) {
    const int n_fine = d_counts[CNT_FINE];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_fine; c += gridDim.x * blockDim.x) {
        if (root[c] != c) continue;
        const int s = csize[c];
        if (s < min_size || s > max_size) continue;
        const int mi = cmin[c];
        int frame = 0, local = mi;
        if (n_frames > 1) {
            frame = frame_of(frame_offsets, n_frames, mi);
            local = mi - frame_offsets[frame];
        }
        const int slot = atomicAdd(&d_counts[CNT_K], 1);
        ckeys[slot] = ckey_make(kc, frame, s, local);
        croots[slot] = (uint32_t)c;
        atomicAdd(&d_counts[CNT_TOTAL], s);
    }
}

// Small-K path: one CTA sorts up to CL_SMALL_MAX clusters with a bitonic network in shared memory, then emits
// rank -> root, sizes and CSR offsets in the same launch.
constexpr int CL_SMALL_MAX = 8192;
constexpr int CL_SMALL_THREADS = 1024;
constexpr size_t CL_SMALL_SMEM = (size_t)CL_SMALL_MAX * 12;  // dynamic: u64 keys + u32 roots

__global__ void __launch_bounds__(CL_SMALL_THREADS) k_clusters_small(const uint64_t* __restrict__ ckeys_in, const uint32_t* __restrict__ croots_in,
                                                                      int K, ClusterKeyCodec kc, uint64_t* __restrict__ ckeys_out,
                                                                      int* __restrict__ crank, int* __restrict__ cl_offsets) {
    extern __shared__ __align__(16) unsigned char cl_smem[];
    uint64_t* sk = reinterpret_cast<uint64_t*>(cl_smem);
    uint32_t* sr = reinterpret_cast<uint32_t*>(sk + CL_SMALL_MAX);
    __shared__ int scratch[36];
    int n2 = 1;
    while (n2 < K) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += CL_SMALL_THREADS) {
        sk[i] = i < K ? ckeys_in[i] : ~0ull;
        sr[i] = i < K ? croots_in[i] : 0u;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += CL_SMALL_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = (i & k) == 0;
                    const uint64_t a = sk[i], b = sk[ixj];
                    if ((a > b) == up) {
                        sk[i] = b; sk[ixj] = a;
                        const uint32_t t = sr[i]; sr[i] = sr[ixj]; sr[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
    // sizes -> exclusive scan -> offsets; each thread owns a contiguous run of <= 8 clusters
    const int per = (K + CL_SMALL_THREADS - 1) / CL_SMALL_THREADS;
    const int s = threadIdx.x * per, e = min(K, s + per);
    int local = 0;
    for (int i = s; i < e; ++i) local += ckey_size(kc, sk[i]);
    int total;
    int run = block_exclusive_scan(local, scratch, &total);
    for (int i = s; i < e; ++i) {
        cl_offsets[i] = run;
        run += ckey_size(kc, sk[i]);
        crank[sr[i]] = i;
        ckeys_out[i] = sk[i];
    }
    if (threadIdx.x == 0) cl_offsets[K] = total;
}

// Large-K path, after the 64-bit radix sort of (key, root): per-block size sums, then offsets + ranks.
constexpr int CLF_THREADS = 256;
constexpr int CLF_MAX_GRID = 592;
__global__ void __launch_bounds__(CLF_THREADS) k_clusters_count(const uint64_t* __restrict__ ckeys, int K, int chunk, ClusterKeyCodec kc,
                                                                 int* __restrict__ block_sums) {
    __shared__ int red[CLF_THREADS / 32];
    const int begin = blockIdx.x * chunk, end = min(K, begin + chunk);
    int s = 0;
    for (int i = begin + threadIdx.x; i < end; i += CLF_THREADS) s += ckey_size(kc, ckeys[i]);
    s = warp_sum(s);
    if (lane_id() == 0) red[warp_id()] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < CLF_THREADS / 32; ++w) t += red[w];
        block_sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(CLF_THREADS) k_clusters_finalize(const uint64_t* __restrict__ ckeys, const uint32_t* __restrict__ croots, int K,
                                                                    int chunk, ClusterKeyCodec kc, const int* __restrict__ block_sums,
                                                                    int* __restrict__ crank, int* __restrict__ cl_offsets) {
    __shared__ int scratch[36];
    int base = block_prefix_of(block_sums, blockIdx.x, scratch);
    const int begin = blockIdx.x * chunk, end = min(K, begin + chunk);
    for (int tb = begin; tb < end; tb += CLF_THREADS) {
        const int i = tb + threadIdx.x;
        const int sz = i < end ? ckey_size(kc, ckeys[i]) : 0;
        int total;
        const int excl = block_exclusive_scan(sz, scratch, &total);
        if (i < end) {
            cl_offsets[i] = base + excl;
            crank[croots[i]] = i;
        }
        base += total;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) cl_offsets[K] = base;
}

// frame_cluster_offsets[f] = first cluster of frame f in the sorted order (batch mode; K+1 sentinel at F).
__global__ void k_frame_cluster_offsets(const uint64_t* __restrict__ ckeys, int K, ClusterKeyCodec kc, int n_frames,
                                        int* __restrict__ frame_cluster_offsets) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > n_frames) return;
    int lo = 0, hi = K;  // first k with frame(k) >= f
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ckey_frame(kc, ckeys[mid]) < f) lo = mid + 1; else hi = mid;
    }
    frame_cluster_offsets[f] = lo;
}

// Cluster rank of every point, scattered back to original point order (key of the final stable partition),
// plus the component label (smallest original index of the component) for partition checks.
__global__ void __launch_bounds__(256) k_point_rank(const float4* __restrict__ spts, const uint32_t* __restrict__ svals, const int* __restrict__ root,
                                                     const int* __restrict__ crank, const int* __restrict__ cmin, int m, int K,
                                                     uint32_t* __restrict__ pkey, int* __restrict__ labels) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int orig = (int)svals[j];                          // sorted position -> original index
    const int r = root[__float_as_int(spts[j].w)];           // .w = fine cell id
    const int k = crank[r];
    pkey[orig] = k < 0 ? (uint32_t)K : (uint32_t)k;
    labels[orig] = cmin[r];
}

// Batch mode: turn global point indices into positions inside the owning frame.
__global__ void __launch_bounds__(256) k_localize_indices(uint32_t* __restrict__ idx, int total, const int* __restrict__ frame_offsets, int n_frames) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int g = (int)idx[t];
    idx[t] = (uint32_t)(g - frame_offsets[frame_of(frame_offsets, n_frames, g)]);
}

// ---- K7: segmented reduction over the CSR index list ----------------------------------------------------------
// Point-parallel: a warp walks 32 consecutive CSR entries at a time.  While the whole group lies inside one cluster
// the lanes just accumulate in registers; a group that straddles cluster boundaries is reduced with a segmented
// warp scan.  Partial results go to per-cluster accumulators (fp64 sums, ordered-int min/max), so the cost is
// independent of the cluster-size distribution (one 800k-point cluster or 100k tiny ones).
struct ClusterStat {  // == mot_cluster_stat
    int count;
    float mean[3], bmin[3], bmax[3];
};
struct StatAcc {       // 48 bytes per cluster
    double sum[3];
    int mn[3], mx[3];  // float_to_ordered encoded
};
__global__ void __launch_bounds__(256) k_stats_init(StatAcc* __restrict__ acc, int K) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    StatAcc a;
    for (int d = 0; d < 3; ++d) { a.sum[d] = 0.0; a.mn[d] = 0x7fffffff; a.mx[d] = (int)0x80000000; }
    acc[k] = a;
}

__device__ __forceinline__ void stat_flush(StatAcc* acc, int cid, double s0, double s1, double s2, float n0, float n1, float n2, float x0,
                                           float x1, float x2) {
    atomicAdd(&acc[cid].sum[0], s0); atomicAdd(&acc[cid].sum[1], s1); atomicAdd(&acc[cid].sum[2], s2);
    atomicMin(&acc[cid].mn[0], float_to_ordered(n0)); atomicMin(&acc[cid].mn[1], float_to_ordered(n1)); atomicMin(&acc[cid].mn[2], float_to_ordered(n2));
    atomicMax(&acc[cid].mx[0], float_to_ordered(x0)); atomicMax(&acc[cid].mx[1], float_to_ordered(x1)); atomicMax(&acc[cid].mx[2], float_to_ordered(x2));
}

constexpr int STAT_THREADS = 256;
// 32-entry groups a warp walks: as many as it takes to give every SM ~16 warps, at most STAT_GROUPS_PER_WARP
inline int stat_groups(long long total, int num_sms) {
    long long g = total / (32ll * 16 * num_sms);
    return (int)(g < 1 ? 1 : (g > 16 ? 16 : g));
}
constexpr int STAT_GROUPS_PER_WARP = 16;  // at most 512 CSR entries per warp; small tables use fewer (stat_groups) so that every SM gets warps
__global__ void __launch_bounds__(STAT_THREADS) k_stats_accumulate(const float4* __restrict__ pts, const int* __restrict__ cl_offsets,
                                                                    const uint32_t* __restrict__ indices, int K, int total,
                                                                    StatAcc* __restrict__ acc, int groups = STAT_GROUPS_PER_WARP) {
    const int lane = lane_id();
    const int warp_global = blockIdx.x * (STAT_THREADS / 32) + warp_id();
    const int t_begin = warp_global * (32 * groups);
    if (t_begin >= total) return;
    int cur = -1, cur_end = 0;  // warp-uniform: cluster whose points the lane accumulators currently hold
    double s0 = 0, s1 = 0, s2 = 0;
    float n0 = INFINITY, n1 = INFINITY, n2 = INFINITY, x0 = -INFINITY, x1 = -INFINITY, x2 = -INFINITY;
    for (int gidx = 0; gidx < groups; ++gidx) {
        const int t0 = t_begin + gidx * 32;
        if (t0 >= total) break;
        const int t = t0 + lane;
        const bool valid = t < total;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) p = pts[indices[t]];
        const int last = min(t0 + 31, total - 1);
        if (cur >= 0 && last < cur_end) {  // whole group inside the current cluster
            if (valid) {
                s0 += (double)p.x; s1 += (double)p.y; s2 += (double)p.z;
                n0 = fminf(n0, p.x); n1 = fminf(n1, p.y); n2 = fminf(n2, p.z);
                x0 = fmaxf(x0, p.x); x1 = fmaxf(x1, p.y); x2 = fmaxf(x2, p.z);
            }
            continue;
        }
        // flush what the lanes hold for `cur`
        if (cur >= 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(kFull, s0, o); s1 += __shfl_xor_sync(kFull, s1, o); s2 += __shfl_xor_sync(kFull, s2, o);
                n0 = fminf(n0, __shfl_xor_sync(kFull, n0, o)); n1 = fminf(n1, __shfl_xor_sync(kFull, n1, o)); n2 = fminf(n2, __shfl_xor_sync(kFull, n2, o));
                x0 = fmaxf(x0, __shfl_xor_sync(kFull, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(kFull, x1, o)); x2 = fmaxf(x2, __shfl_xor_sync(kFull, x2, o));
            }
            if (lane == 0) stat_flush(acc, cur, s0, s1, s2, n0, n1, n2, x0, x1, x2);
        }
        // cluster of every entry of the group: largest k with cl_offsets[k] <= t
        int cid = -1;
        if (valid) {
            int lo = cur >= 0 ? cur : 0, hi = K - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (cl_offsets[mid] <= t) lo = mid; else hi = mid - 1;
            }
            cid = lo;
        }
        // segmented inclusive scan over lanes (segments = runs of equal cid; cid is non-decreasing across lanes)
        double a0 = (double)p.x, a1 = (double)p.y, a2 = (double)p.z;
        float m0 = valid ? p.x : INFINITY, m1 = valid ? p.y : INFINITY, m2 = valid ? p.z : INFINITY;
        float y0 = valid ? p.x : -INFINITY, y1 = valid ? p.y : -INFINITY, y2 = valid ? p.z : -INFINITY;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int oc = __shfl_up_sync(kFull, cid, o);
            const double b0 = __shfl_up_sync(kFull, a0, o), b1 = __shfl_up_sync(kFull, a1, o), b2 = __shfl_up_sync(kFull, a2, o);
            const float c0 = __shfl_up_sync(kFull, m0, o), c1 = __shfl_up_sync(kFull, m1, o), c2 = __shfl_up_sync(kFull, m2, o);
            const float d0 = __shfl_up_sync(kFull, y0, o), d1 = __shfl_up_sync(kFull, y1, o), d2 = __shfl_up_sync(kFull, y2, o);
            if (lane >= o && oc == cid) {
                a0 += b0; a1 += b1; a2 += b2;
                m0 = fminf(m0, c0); m1 = fminf(m1, c1); m2 = fminf(m2, c2);
                y0 = fmaxf(y0, d0); y1 = fmaxf(y1, d1); y2 = fmaxf(y2, d2);
            }
        }
        const int next_cid = __shfl_down_sync(kFull, cid, 1);
        const bool tail = valid && (lane == 31 || next_cid != cid);
        const int last_cid = __shfl_sync(kFull, cid, last - t0);
        // the last segment stays in registers (lane `last - t0` holds its reduction) if the cluster continues
        const bool keep = tail && cid == last_cid;
        if (tail && !keep) stat_flush(acc, cid, a0, a1, a2, m0, m1, m2, y0, y1, y2);
        cur = last_cid;
        cur_end = cl_offsets[cur + 1];
        s0 = keep ? a0 : 0.0; s1 = keep ? a1 : 0.0; s2 = keep ? a2 : 0.0;
        n0 = keep ? m0 : INFINITY; n1 = keep ? m1 : INFINITY; n2 = keep ? m2 : INFINITY;
        x0 = keep ? y0 : -INFINITY; x1 = keep ? y1 : -INFINITY; x2 = keep ? y2 : -INFINITY;
    }
    if (cur >= 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(kFull, s0, o); s1 += __shfl_xor_sync(kFull, s1, o); s2 += __shfl_xor_sync(kFull, s2, o);
            n0 = fminf(n0, __shfl_xor_sync(kFull, n0, o)); n1 = fminf(n1, __shfl_xor_sync(kFull, n1, o)); n2 = fminf(n2, __shfl_xor_sync(kFull, n2, o));
            x0 = fmaxf(x0, __shfl_xor_sync(kFull, x0, o)); x1 = fmaxf(x1, __shfl_xor_sync(kFull, x1, o)); x2 = fmaxf(x2, __shfl_xor_sync(kFull, x2, o));
        }
        if (lane == 0) stat_flush(acc, cur, s0, s1, s2, n0, n1, n2, x0, x1, x2);
    }
}

__global__ void __launch_bounds__(256) k_stats_finalize(const StatAcc* __restrict__ acc, const int* __restrict__ cl_offsets, int K,
                                                         ClusterStat* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const StatAcc a = acc[k];
    ClusterStat st;
    st.count = cl_offsets[k + 1] - cl_offsets[k];
    for (int d = 0; d < 3; ++d) {
        st.mean[d] = (float)(a.sum[d] / (double)st.count);
        st.bmin[d] = ordered_to_float_bits(a.mn[d]);
        st.bmax[d] = ordered_to_float_bits(a.mx[d]);
    }
    out[k] = st;
}

// ---- VoxelGrid (reference call site MOT.cpp:452-456; PCL VoxelGrid::applyFilter restated) -----------------
// key = ijk0 + ijk1*div0 + ijk2*div0*div1 with ijk = floor(p * inv_leaf) - min_b in fp32, exactly PCL's arithmetic.
struct VoxelParams {
    float inv[3];
    int minb[3];
    int mul1, mul2;
};
__global__ void __launch_bounds__(256) k_voxel_keys(const float4* __restrict__ pts, int n, VoxelParams vp, uint32_t* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = ld_stream(pts + i);
    const int i0 = (int)floorf(__fmul_rn(p.x, vp.inv[0])) - vp.minb[0];
    const int i1 = (int)floorf(__fmul_rn(p.y, vp.inv[1])) - vp.minb[1];
    const int i2 = (int)floorf(__fmul_rn(p.z, vp.inv[2])) - vp.minb[2];
    keys[i] = (uint32_t)(i0 + i1 * vp.mul1 + i2 * vp.mul2);
}
// segment heads of a sorted key array: per-block counts, then starts[v] = first sorted position of voxel v
__global__ void __launch_bounds__(256) k_seg_count(const uint32_t* __restrict__ skeys, int n, int chunk, int* __restrict__ counts) {
    __shared__ int red[8];
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    int c = 0;
    for (int j = begin + threadIdx.x; j < end; j += 256) c += (j == 0 || skeys[j] != skeys[j - 1]);
    c = warp_sum(c);
    if (lane_id() == 0) red[warp_id()] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        counts[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) k_seg_write(const uint32_t* __restrict__ skeys, int n, int chunk, const int* __restrict__ counts,
                                                    int* __restrict__ starts, int* __restrict__ n_segments) {
    __shared__ int scratch[36];
    int base = block_prefix_of(counts, blockIdx.x, scratch);
    const int begin = blockIdx.x * chunk, end = min(n, begin + chunk);
    for (int tb = begin; tb < end; tb += 256) {
        const int j = tb + threadIdx.x;
        const int head = j < end && (j == 0 || skeys[j] != skeys[j - 1]);
        int total;
        const int excl = block_exclusive_scan(head, scratch, &total);
        if (head) starts[base + excl] = j;
        base += total;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        starts[base] = n;
        *n_segments = base;
    }
}
// centroid of every voxel from the fp64 accumulators of k_stats_accumulate (PCL divides an fp32 sum by the count)
__global__ void __launch_bounds__(256) k_voxel_finalize(const StatAcc* __restrict__ acc, const int* __restrict__ starts, int V,
                                                         float4* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double cnt = (double)(starts[v + 1] - starts[v]);
    out[v] = make_float4((float)(acc[v].sum[0] / cnt), (float)(acc[v].sum[1] / cnt), (float)(acc[v].sum[2] / cnt), 1.0f);
}

// ---- K8: the reference's getCentroid (MOT.cpp:708-822) ----------------------------------------------------------
// Step 1 (farthest pair, O(n^2)) is spread over `slabs` CTAs per cluster (rows i = slab, slab+slabs, ...); each
// writes its best candidate; step 2/3 (farthest point from the XY line, circumcentre) reduce them per cluster.
// The reference's scan keeps the FIRST pair, in (i, j) lexicographic order, whose float distance is strictly
// the largest; candidates are therefore compared as (dist desc, i asc, j asc).
struct PairCand {
    float dist;
    int i, j;
};
__device__ __forceinline__ bool cand_better(const PairCand& a, const PairCand& b) {  // a beats b
    if (a.dist != b.dist) return a.dist > b.dist;
    if (a.i != b.i) return a.i < b.i;
    return a.j < b.j;
}
__device__ __forceinline__ double ref_euc_sq(const float4& a, const float4& b) {  // MOT.cpp:1025-1028: the fp64 sum under the root
    const double dx = __dsub_rn((double)a.x, (double)b.x), dy = __dsub_rn((double)a.y, (double)b.y), dz = __dsub_rn((double)a.z, (double)b.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ float ref_euc_dist(const float4& a, const float4& b) {  // ... fp64 root, then float
    return __double2float_rn(__dsqrt_rn(ref_euc_sq(a, b)));
}
// One step of the reference's pair scan: the pair (i, j) replaces `best` iff its float distance is strictly larger.  float(sqrt(s))
// is monotone in s, so a pair whose s does not exceed the largest s seen so far (`seen_s`) cannot win: no root for it.
__device__ __forceinline__ void pair_scan_step(const float4& pi, const float4& pj, int i, int j, PairCand& best, double& seen_s) {
    const double s = ref_euc_sq(pi, pj);
    if (s > seen_s) {
        seen_s = s;
        const float d = __double2float_rn(__dsqrt_rn(s));
        if (d > best.dist) { best.dist = d; best.i = i; best.j = j; }
    }
}
// best candidate of cands[0 .. slabs) (slabs <= 64), computed by the first two warps of the CTA and returned to every thread
__device__ __forceinline__ PairCand reduce_cands(const PairCand* cands, int slabs, PairCand* s2 /* shared [2] */) {
    PairCand best;
    best.dist = -1.0f; best.i = 0x7fffffff; best.j = 0x7fffffff;
    if (threadIdx.x < 64) {
        if ((int)threadIdx.x < slabs) best = cands[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            PairCand other;
            other.dist = __shfl_xor_sync(kFull, best.dist, o);
            other.i = __shfl_xor_sync(kFull, best.i, o);
            other.j = __shfl_xor_sync(kFull, best.j, o);
            if (cand_better(other, best)) best = other;
        }
        if (lane_id() == 0) s2[warp_id()] = best;
    }
    __syncthreads();
    best = s2[0];
    if (cand_better(s2[1], best)) best = s2[1];
    return best;
}

constexpr int FP_THREADS = 256;
constexpr int FP_SMEM_POINTS = 2048;
__global__ void __launch_bounds__(FP_THREADS) k_farthest_pair(const float4* __restrict__ pts, const int* __restrict__ cl_offsets,
                                                               const uint32_t* __restrict__ indices, int K, int slabs,
                                                               PairCand* __restrict__ cands /* [K][slabs] */) {
    __shared__ float4 sp[FP_SMEM_POINTS];
    __shared__ PairCand sbest[FP_THREADS / 32];
    const int c = blockIdx.x / slabs, slab = blockIdx.x % slabs;
    if (c >= K) return;
    const int s = cl_offsets[c], n = cl_offsets[c + 1] - s;
    const bool staged = n <= FP_SMEM_POINTS;
    if (staged) {
        for (int t = threadIdx.x; t < n; t += FP_THREADS) sp[t] = pts[indices[s + t]];
        __syncthreads();
    }
    PairCand best;
    best.dist = -1.0f; best.i = 0x7fffffff; best.j = 0x7fffffff;
    double seen_s = -1.0;
    // rows are dealt to (slab, warp) round-robin; the lanes of a warp sweep j
    const int row_stride = slabs * (FP_THREADS / 32);
    for (int i = slab * (FP_THREADS / 32) + warp_id(); i < n - 1; i += row_stride) {
        const float4 pi = staged ? sp[i] : pts[indices[s + i]];
        for (int j = i + 1 + lane_id(); j < n; j += 32) {
            const float4 pj = staged ? sp[j] : pts[indices[s + j]];
            pair_scan_step(pi, pj, i, j, best, seen_s);  // (i, j) ascending within a lane: strict > keeps the first
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        PairCand other;
        other.dist = __shfl_xor_sync(kFull, best.dist, o);
        other.i = __shfl_xor_sync(kFull, best.i, o);
        other.j = __shfl_xor_sync(kFull, best.j, o);
        if (cand_better(other, best)) best = other;
    }
    if (lane_id() == 0) sbest[warp_id()] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < FP_THREADS / 32; ++w)
            if (cand_better(sbest[w], best)) best = sbest[w];
        cands[(size_t)c * slabs + slab] = best;
    }
}

// The arithmetic of getCentroid's steps 2 and 3, shared by k_circumcentre and the single-launch small-frame path
// (frame_small.cuh) so that both produce the same bits.
struct CcLine {
    double Pi[3], Pj[3];   // the farthest pair as the reference leaves it after the pair loop (zero if the cluster has < 2 points)
    double V0, V1, V2, denom;
};
__device__ __forceinline__ void cc_line_from_pair(const float4& a, const float4& b, bool have_pair, CcLine& L) {
    for (int d = 0; d < 3; ++d) { L.Pi[d] = 0.0; L.Pj[d] = 0.0; }
    L.V0 = 0.0; L.V1 = 0.0; L.V2 = 0.0;
    if (have_pair) {
        L.Pi[0] = a.x; L.Pi[1] = a.y; L.Pi[2] = a.z;
        L.Pj[0] = b.x; L.Pj[1] = b.y; L.Pj[2] = b.z;
        L.V0 = __ddiv_rn(__dsub_rn(L.Pj[1], L.Pi[1]), __dsub_rn(L.Pj[0], L.Pi[0]));  // MOT.cpp:753
        L.V1 = -1.0;
        L.V2 = __dadd_rn(__dmul_rn(L.V0, -L.Pi[0]), L.Pi[1]);                         // MOT.cpp:755
    }
    L.denom = __dsqrt_rn(__dadd_rn(__dmul_rn(L.V0, L.V0), __dmul_rn(L.V1, L.V1)));
}
// float distance of p from the XY line; `skip` = p equals Pi or Pj (the reference skips those without raising the bar)
__device__ __forceinline__ float cc_line_dist(const CcLine& L, const float4& p, bool& skip) {
    const double px = p.x, py = p.y, pz = p.z;
    const double num = fabs(__dadd_rn(__dadd_rn(__dmul_rn(L.V0, px), __dmul_rn(L.V1, py)), L.V2));
    const bool eqi = px == L.Pi[0] && py == L.Pi[1] && pz == L.Pi[2];
    const bool eqj = px == L.Pj[0] && py == L.Pj[1] && pz == L.Pj[2];
    skip = eqi || eqj;
    return __double2float_rn(__ddiv_rn(num, L.denom));
}
// step 3, MOT.cpp:787-809: float A..G from double expressions, float final arithmetic, no FMA
__device__ __forceinline__ float4 cc_finish(const CcLine& L, const double Pk[3], float intensity) {
    const double* Pi = L.Pi;
    const double* Pj = L.Pj;
    const float A = __double2float_rn(__dsub_rn(Pj[0], Pi[0]));
    const float B = __double2float_rn(__dsub_rn(Pj[1], Pi[1]));
    const float C = __double2float_rn(__dsub_rn(Pk[0], Pi[0]));
    const float D = __double2float_rn(__dsub_rn(Pk[1], Pi[1]));
    const float E = __double2float_rn(__dadd_rn(__dmul_rn((double)A, __dadd_rn(Pi[0], Pj[0])), __dmul_rn((double)B, __dadd_rn(Pi[1], Pj[1]))));
    const float F = __double2float_rn(__dadd_rn(__dmul_rn((double)C, __dadd_rn(Pi[0], Pk[0])), __dmul_rn((double)D, __dadd_rn(Pi[1], Pk[1]))));
    const float G = __double2float_rn(__dmul_rn(2.0, __dsub_rn(__dmul_rn((double)A, __dsub_rn(Pk[1], Pj[1])), __dmul_rn((double)B, __dsub_rn(Pk[0], Pj[0])))));
    float4 o;
    if (G == 0.0f) { o.x = __double2float_rn(Pi[0]); o.y = __double2float_rn(Pi[1]); }
    else {
        o.x = __fdiv_rn(__fsub_rn(__fmul_rn(D, E), __fmul_rn(B, F)), G);
        o.y = __fdiv_rn(__fsub_rn(__fmul_rn(A, F), __fmul_rn(C, E)), G);
    }
    o.z = 0.0f;
    o.w = intensity;
    return o;
}

constexpr int CC_THREADS = 128;
__global__ void __launch_bounds__(CC_THREADS) k_circumcentre(const float4* __restrict__ pts, const int* __restrict__ cl_offsets,
                                                              const uint32_t* __restrict__ indices, int K, int slabs,
                                                              const PairCand* __restrict__ cands, float intensity, float4* __restrict__ out,
                                                              const float* __restrict__ frame_stamps = nullptr,
                                                              const int* __restrict__ frame_cl_offsets = nullptr, int n_frames = 1) {
    __shared__ float sdist[CC_THREADS / 32];
    __shared__ int sk[CC_THREADS / 32];
    __shared__ PairCand s2[2];
    for (int c = blockIdx.x; c < K; c += gridDim.x) {
        const int s = cl_offsets[c], n = cl_offsets[c + 1] - s;
        const PairCand best = reduce_cands(cands + (size_t)c * slabs, slabs, s2);
        CcLine L;
        {
            const bool have = best.dist >= 0.0f;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            cc_line_from_pair(have ? pts[indices[s + best.i]] : z, have ? pts[indices[s + best.j]] : z, have, L);
        }
        // step 2: first k with strictly largest float line distance, skipping points equal to Pi / Pj.
        // (the reference only updates dist_max when the point is accepted, so skipped points do not raise the bar)
        float bd = -1.0f;
        int bk = 0x7fffffff;
        for (int k = threadIdx.x; k < n; k += CC_THREADS) {
            bool skip;
            const float d = cc_line_dist(L, pts[indices[s + k]], skip);
            if (d > bd && !skip) { bd = d; bk = k; }  // NaN distances never satisfy >, as in the reference
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(kFull, bd, o);
            const int ok = __shfl_xor_sync(kFull, bk, o);
            if (od > bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
        }
        __syncthreads();
        if (lane_id() == 0) { sdist[warp_id()] = bd; sk[warp_id()] = bk; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < CC_THREADS / 32; ++w)
                if (sdist[w] > bd || (sdist[w] == bd && sk[w] < bk)) { bd = sdist[w]; bk = sk[w]; }
            double Pk[3] = {0, 0, 0};  // UB policy: zero-initialised
            if (bk != 0x7fffffff && bd >= 0.0f) {
                const float4 p = pts[indices[s + bk]];
                Pk[0] = p.x; Pk[1] = p.y; Pk[2] = p.z;
            }
            float4 o = cc_finish(L, Pk, intensity);
            if (frame_stamps) {  // batch: the stamp of the frame that owns cluster c (largest f with frame_cl_offsets[f] <= c)
                int lo = 0, hi = n_frames - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (frame_cl_offsets[mid] <= c) lo = mid; else hi = mid - 1;
                }
                o.w = frame_stamps[lo];
            }
            out[c] = o;
        }
        __syncthreads();
    }
}

}  // namespace mot
